/* keygen.h — launch interface of keygen.cu (GPU key generation, encryption, phase). */
#ifndef IEACHE_KEYGEN_H
#define IEACHE_KEYGEN_H
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "csprng.h"

namespace ieache {

/* secret key bits from the secret key stream: stream RNG_LWE_KEY or RNG_TLWE_KEY */
void host_key_bits(const RngKey &secret, uint64_t stream, int32_t *out, int count);

/* fills bkfft / ksk (device layouts of kernels.h); the *_export pointers (device, may be null)
 * receive libtfhe-order coefficient arrays: bk [n][kpl][2][1024], ksk [1024][t][base][n+1] (zeroed by the caller) */
cudaError_t launch_keygen(const RngKeys &keys, const DevParams &p, double ks_stdev, double bk_stdev, const int32_t *d_lwe_key,
                          const int32_t *d_tlwe_key, double2 *d_shat /*512*/, double2 *bkfft, int32_t *ksk,
                          int32_t *bk_coef_export, int32_t *ksk_export, cudaStream_t s);
cudaError_t launch_encrypt(const RngKeys &keys, int n, double stdev, int32_t mu, const int32_t *d_lwe_key, const int32_t *d_bits,
                           int32_t *out, long long count, cudaStream_t s);
cudaError_t launch_phase(int n, const int32_t *d_lwe_key, const int32_t *samples, int32_t *phases, long long count, cudaStream_t s);

} // namespace ieache
#endif
