/* keygen.h — launch interface of keygen.cu (GPU key generation, encryption, phase). */
#ifndef IEACHE_KEYGEN_H
#define IEACHE_KEYGEN_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace ieache {

/* deterministic key bits from (seed, stream): stream 1 = LWE key, 2 = TLWE key */
void host_random_bits(uint64_t seed, uint64_t stream, int32_t *out, int count);

/* fills bkfft / ksk (device layouts of kernels.h); the *_export pointers (device, may be null)
 * receive libtfhe-order coefficient arrays: bk [n][kpl][2][1024], ksk [1024][t][base][n+1] (zeroed by the caller) */
cudaError_t launch_keygen(uint64_t seed, const DevParams &p, double ks_stdev, double bk_stdev, const int32_t *d_lwe_key,
                          const int32_t *d_tlwe_key, double2 *d_shat /*512*/, double2 *bkfft, int32_t *ksk,
                          int32_t *bk_coef_export, int32_t *ksk_export, cudaStream_t s);
cudaError_t launch_encrypt(uint64_t seed, int n, double stdev, int32_t mu, const int32_t *d_lwe_key, const int32_t *d_bits,
                           int32_t *out, long long count, cudaStream_t s);
cudaError_t launch_phase(int n, const int32_t *d_lwe_key, const int32_t *samples, int32_t *phases, long long count, cudaStream_t s);

} // namespace ieache
#endif
