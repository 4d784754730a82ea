/*
 * keygen.cu — key generation, encryption and decryption on the GPU.
 *
 * B200-native counterparts of the callers on either side of the hot path:
 *   Keygen/keygen.c:22-51   new_random_gate_bootstrapping_secret_keyset  -> keygen_bk_kernel, keygen_ksk_kernel
 *   Client1/alice.c:117     bootsSymEncrypt                              -> encrypt_kernel
 *   Output/verif.c:93       bootsSymDecrypt                              -> phase_kernel
 * They also give bench.py a synthetic key and synthetic ciphertexts without touching the oracle.
 *
 * Randomness: counter-mode ChaCha20 (csprng.h) under two independent 256-bit keys per key set — `secret` for the key
 * bits and every noise term, `mask` for the published uniform masks — taken from the operating system unless a
 * seed is given (tests only).  Box-Muller Gaussians.  libtfhe's own generator (std::default_random_engine behind
 * tfhe_random_generator_setSeed, Keygen/keygen.c:30-36) is not reproducible across standard libraries, so no
 * attempt is made to reproduce its streams (SURVEY.md App. A "Randomness").
 */
#include "kernels.h"
#include "br_core.h"
#include "keygen.h"
#include "csprng.h"

#include <sys/random.h>
#include <string.h>

namespace ieache {

/* defined in kernels.cu: device addresses of the pass-2 / pass-3 twiddle tables */
cudaError_t twiddle_ptrs(const Tw **tw2, const Tw **tw3);

int rng_keys_from_os(RngKeys &out)
{
    unsigned char *p = reinterpret_cast<unsigned char *>(&out);
    size_t got = 0;
    while (got < sizeof(out)) {
        const ssize_t r = getrandom(p + got, sizeof(out) - got, 0);
        if (r <= 0) return -1;
        got += (size_t)r;
    }
    return 0;
}
void rng_keys_from_seed(uint64_t seed, RngKeys &out)
{
    /* the seed keys a ChaCha20 stream whose first two blocks become the two keys: reproducible, and no stronger than
     * the 64-bit seed — tests only */
    RngKey base{};
    base.k[0] = (uint32_t)seed; base.k[1] = (uint32_t)(seed >> 32); base.k[2] = 0x69656163u; base.k[3] = 0x68655f62u; /* "ieache_b" */
    uint32_t blk[16];
    chacha20_block(base, 0, RNG_DERIVE, blk);
    memcpy(out.secret.k, blk, 32);
    chacha20_block(base, 1, RNG_DERIVE, blk);
    memcpy(out.mask.k, blk, 32);
}

void host_key_bits(const RngKey &secret, uint64_t stream, int32_t *out, int count)
{
    uint32_t blk[16];
    for (int i = 0; i < count; i++) {
        if ((i & 511) == 0) chacha20_block(secret, (uint64_t)(i >> 9), stream, blk);
        out[i] = (int32_t)((blk[(i & 511) >> 5] >> (i & 31)) & 1u);
    }
}

__device__ __forceinline__ void gsync() { __syncthreads(); }

/* 64-thread transforms on a private shared buffer (same passes as the blind-rotation kernel) */
__device__ __forceinline__ void fwd64(double (&xr)[8], double (&xi)[8], cd *buf, int tid, const Tw &w1, const Tw &w2, const Tw &w3)
{
    pass_fwd(xr, xi, w1); st_pass1(buf, tid, xr, xi); gsync();
    ld_pass2(buf, tid, xr, xi); pass_fwd(xr, xi, w2); st_pass2(buf, tid, xr, xi); gsync();
    ld_pass3(buf, tid, xr, xi); pass_fwd(xr, xi, w3); gsync();
}
__device__ __forceinline__ void inv64(double (&xr)[8], double (&xi)[8], cd *buf, int tid, const Tw &w1, const Tw &w2, const Tw &w3)
{
    pass_inv(xr, xi, w3); st_ipass3(buf, tid, xr, xi); gsync();
    ld_ipass2(buf, tid, xr, xi); pass_inv(xr, xi, w2); st_ipass2(buf, tid, xr, xi); gsync();
    ld_ipass1(buf, tid, xr, xi); pass_inv(xr, xi, w1); gsync();
}

/* transform of the TLWE key polynomial (binary), unscaled, layout [r][t3] */
__global__ void __launch_bounds__(64) keygen_sfft_kernel(const int32_t *__restrict__ tlwe_key, double2 *__restrict__ shat,
                                                         const Tw *__restrict__ d_tw2, const Tw *__restrict__ d_tw3)
{
    __shared__ cd buf[kBufElems];
    const int tid = threadIdx.x;
    double xr[8], xi[8];
#pragma unroll
    for (int m = 0; m < 8; m++) { xr[m] = (double)tlwe_key[tid + 64 * m]; xi[m] = (double)tlwe_key[tid + 64 * m + 512]; }
    fwd64(xr, xi, buf, tid, tw_pass1(), d_tw2[tid >> 3], d_tw3[tid]);
#pragma unroll
    for (int r = 0; r < 8; r++) shat[r * 64 + tid] = make_double2(xr[r], xi[r]);
}

/* one CTA per TGSW row (i, r): TLWE encryption of 0 plus s_i * h_p on the diagonal, written
 * straight into the transform-domain layout (and optionally in coefficient form for export) */
__global__ void __launch_bounds__(64)
keygen_bk_kernel(RngKeys keys, int l, int Bgbit, double bk_stdev, const int32_t *__restrict__ lwe_key,
                 const double2 *__restrict__ shat, double2 *__restrict__ bkfft, int32_t *__restrict__ bk_coef,
                 const Tw *__restrict__ d_tw2, const Tw *__restrict__ d_tw3)
{
    __shared__ cd buf[kBufElems];
    const int row = blockIdx.x, tid = threadIdx.x;
    const int kpl = 2 * l, i = row / kpl, r = row % kpl, q = r / l, pp = r % l;
    const Tw w1 = tw_pass1(), w2 = d_tw2[tid >> 3], w3 = d_tw3[tid];
    const int32_t msg = lwe_key[i] * (int32_t)(1u << (32 - (pp + 1) * Bgbit));
    int32_t a[16];
    double xr[8], xi[8];
    {   /* the thread's 16 mask coefficients are the 16 words of one key-stream block */
        uint32_t blk[16];
        chacha20_block(keys.mask, (uint64_t)row * 64 + tid, RNG_BK_MASK, blk);
#pragma unroll
        for (int h = 0; h < 16; h++) a[h] = (int32_t)blk[h];
    }
#pragma unroll
    for (int m = 0; m < 8; m++) { xr[m] = (double)a[m]; xi[m] = (double)a[8 + m]; }
    fwd64(xr, xi, buf, tid, w1, w2, w3);
    /* a * s' in the transform domain */
    double pr[8], pi[8];
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
        const double2 s = shat[rr * 64 + tid];
        pr[rr] = xr[rr] * s.x - xi[rr] * s.y;
        pi[rr] = xr[rr] * s.y + xi[rr] * s.x;
    }
    inv64(pr, pi, buf, tid, w1, w2, w3);
    int32_t b[16];
    {
        uint32_t n0[16], n1[16]; /* 16 Gaussians = 32 words = two blocks of the secret stream */
        chacha20_block(keys.secret, ((uint64_t)row * 64 + tid) * 2, RNG_BK_NOISE, n0);
        chacha20_block(keys.secret, ((uint64_t)row * 64 + tid) * 2 + 1, RNG_BK_NOISE, n1);
#pragma unroll
        for (int h = 0; h < 16; h++) {
            const double v = ((h < 8) ? pr[h & 7] : pi[h & 7]) * (1.0 / 512.0);
            b[h] = round_to_torus(v) + gauss_torus(h < 8 ? n0[2 * h] : n1[2 * (h - 8)], h < 8 ? n0[2 * h + 1] : n1[2 * (h - 8) + 1], bk_stdev);
        }
    }
    if (tid == 0) { if (q == 0) a[0] += msg; else b[0] += msg; }
    if (bk_coef) {
        int32_t *oa = bk_coef + (size_t)row * 2 * kN, *ob = oa + kN;
#pragma unroll
        for (int h = 0; h < 16; h++) { const int j = tid + 64 * (h & 7) + 512 * (h >> 3); oa[j] = a[h]; ob[j] = b[h]; }
    }
    /* transform-domain rows, scaled by 1/512 (the blind rotation's inverse is unnormalised) */
    const double sc = 1.0 / 512.0;
    double2 *o = bkfft + (size_t)row * 2 * kHalfN;
#pragma unroll
    for (int m = 0; m < 8; m++) { xr[m] = (double)a[m]; xi[m] = (double)a[8 + m]; }
    fwd64(xr, xi, buf, tid, w1, w2, w3);
#pragma unroll
    for (int rr = 0; rr < 8; rr++) o[rr * 64 + tid] = make_double2(xr[rr] * sc, xi[rr] * sc);
#pragma unroll
    for (int m = 0; m < 8; m++) { xr[m] = (double)b[m]; xi[m] = (double)b[8 + m]; }
    fwd64(xr, xi, buf, tid, w1, w2, w3);
#pragma unroll
    for (int rr = 0; rr < 8; rr++) o[kHalfN + rr * 64 + tid] = make_double2(xr[rr] * sc, xi[rr] * sc);
}

/* one warp per key-switch row (i, j, d), d = 1..base-1 */
__global__ void keygen_ksk_kernel(RngKeys keys, int n, int t, int basebit, double ks_stdev, const int32_t *__restrict__ lwe_key,
                                  const int32_t *__restrict__ tlwe_key, int32_t *__restrict__ ksk, int32_t *__restrict__ ksk_export,
                                  int rows)
{
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int basem1 = (1 << basebit) - 1;
    const int ij = row / basem1, d = row % basem1 + 1, i = ij / t, j = ij % t;
    int32_t *o = ksk + (size_t)row * kLweStride;
    int32_t *e = ksk_export ? ksk_export + ((size_t)ij * (basem1 + 1) + d) * (n + 1) : nullptr;
    int32_t acc = 0;
    uint32_t blk[16];
    for (int q = lane, k = 0; q < kLweStride; q += 32, k++) { /* word k of the lane: block k / 16 of its (row, lane) counter pair */
        if ((k & 15) == 0) chacha20_block(keys.mask, ((uint64_t)row * 32 + lane) * 2 + (k >> 4), RNG_KS_MASK, blk);
        int32_t v = 0;
        if (q < n) { v = (int32_t)blk[k & 15]; acc += v * lwe_key[q]; }
        if (q != n) { o[q] = v; if (e && q < n) e[q] = v; }
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        chacha20_block(keys.secret, (uint64_t)row, RNG_KS_NOISE, blk);
        const int32_t msg = (int32_t)((uint32_t)(d * tlwe_key[i]) << (32 - (j + 1) * basebit));
        const int32_t b = acc + msg + gauss_torus(blk[0], blk[1], ks_stdev);
        o[n] = b;
        if (e) e[n] = b;
    }
}

/* bootsSymEncrypt of `count` bits, one warp per sample, written with stride kLweStride */
__global__ void encrypt_kernel(RngKeys keys, int n, double stdev, int32_t mu, const int32_t *__restrict__ lwe_key,
                               const int32_t *__restrict__ bits, int32_t *__restrict__ out, long long count)
{
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= count) return;
    int32_t *o = out + (size_t)s * kLweStride;
    int32_t acc = 0;
    uint32_t blk[16];
    for (int q = lane, k = 0; q < kLweStride; q += 32, k++) {
        if ((k & 15) == 0) chacha20_block(keys.mask, ((uint64_t)s * 32 + lane) * 2 + (k >> 4), RNG_ENC_MASK, blk);
        int32_t v = 0;
        if (q < n) { v = (int32_t)blk[k & 15]; acc += v * lwe_key[q]; }
        if (q != n) o[q] = v;
    }
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        chacha20_block(keys.secret, (uint64_t)s, RNG_ENC_NOISE, blk);
        o[n] = acc + (bits[s] ? mu : -mu) + gauss_torus(blk[0], blk[1], stdev);
    }
}

/* phase = b - <a, s>, one warp per sample */
__global__ void phase_kernel(int n, const int32_t *__restrict__ lwe_key, const int32_t *__restrict__ samples, int32_t *__restrict__ phases,
                             long long count)
{
    const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (s >= count) return;
    const int32_t *p = samples + (size_t)s * kLweStride;
    int32_t acc = 0;
    for (int q = lane; q < n; q += 32) acc += p[q] * lwe_key[q];
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) phases[s] = p[n] - acc;
}

cudaError_t launch_keygen(const RngKeys &keys, const DevParams &p, double ks_stdev, double bk_stdev, const int32_t *d_lwe_key,
                          const int32_t *d_tlwe_key, double2 *d_shat, double2 *bkfft, int32_t *ksk, int32_t *bk_coef_export,
                          int32_t *ksk_export, cudaStream_t s)
{
    const Tw *tw2 = nullptr, *tw3 = nullptr;
    cudaError_t e = twiddle_ptrs(&tw2, &tw3);
    if (e != cudaSuccess) return e;
    keygen_sfft_kernel<<<1, 64, 0, s>>>(d_tlwe_key, d_shat, tw2, tw3);
    keygen_bk_kernel<<<p.n * 2 * p.l, 64, 0, s>>>(keys, p.l, p.Bgbit, bk_stdev, d_lwe_key, d_shat, bkfft, bk_coef_export, tw2, tw3);
    const int rows = kN * p.ks_t * ((1 << p.ks_basebit) - 1);
    keygen_ksk_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(keys, p.n, p.ks_t, p.ks_basebit, ks_stdev, d_lwe_key, d_tlwe_key, ksk,
                                                             ksk_export, rows);
    return cudaGetLastError();
}
cudaError_t launch_encrypt(const RngKeys &keys, int n, double stdev, int32_t mu, const int32_t *d_lwe_key, const int32_t *d_bits,
                           int32_t *out, long long count, cudaStream_t s)
{
    if (count <= 0) return cudaSuccess;
    const long long blocks = (count * 32 + 255) / 256;
    encrypt_kernel<<<(unsigned)blocks, 256, 0, s>>>(keys, n, stdev, mu, d_lwe_key, d_bits, out, count);
    return cudaGetLastError();
}
cudaError_t launch_phase(int n, const int32_t *d_lwe_key, const int32_t *samples, int32_t *phases, long long count, cudaStream_t s)
{
    if (count <= 0) return cudaSuccess;
    const long long blocks = (count * 32 + 255) / 256;
    phase_kernel<<<(unsigned)blocks, 256, 0, s>>>(n, d_lwe_key, samples, phases, count);
    return cudaGetLastError();
}

} // namespace ieache
