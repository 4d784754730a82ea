/*
 * csprng.h — counter-mode ChaCha20 for key generation and encryption (host and device).
 *
 * Replaces libtfhe's generator on the paths Keygen/keygen.c:30-36 (new_random_gate_bootstrapping_secret_keyset) and
 * Client1/alice.c:117 (bootsSymEncrypt) reach.  libtfhe draws from std::default_random_engine, which is neither
 * reproducible across standard libraries nor cryptographic; here every random word is a word of a ChaCha20 key
 * stream (RFC 8439 block function, 64-bit block counter + 64-bit stream id), keyed by 256 bits that come from the
 * operating system (getrandom) unless the caller asks for a seeded, reproducible stream — which is for tests only:
 * a seed is 64 bits, and whoever knows it regenerates the secret key.
 *
 * Two independent keys per key set: `secret` drives the LWE / TLWE key bits and every noise term, `mask` drives the
 * uniform masks that are published in cloud.key and in ciphertexts.  Nothing published is an output of the key
 * stream that produced a secret.
 */
#ifndef IEACHE_CSPRNG_H
#define IEACHE_CSPRNG_H

#include <stdint.h>

#ifdef __CUDACC__
#define IE_RNG_HD __host__ __device__ __forceinline__
#else
#define IE_RNG_HD inline
#endif

namespace ieache {

struct RngKey { uint32_t k[8]; };
struct RngKeys { RngKey secret, mask; };

IE_RNG_HD uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

/* one ChaCha20 block: 16 key-stream words for (key, counter, stream) */
IE_RNG_HD void chacha20_block(const RngKey &key, uint64_t counter, uint64_t stream, uint32_t (&out)[16])
{
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                      key.k[0], key.k[1], key.k[2], key.k[3], key.k[4], key.k[5], key.k[6], key.k[7],
                      (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#define IE_QR(a, b, c, d)                                                  \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
    x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        IE_QR(0, 4, 8, 12) IE_QR(1, 5, 9, 13) IE_QR(2, 6, 10, 14) IE_QR(3, 7, 11, 15)
        IE_QR(0, 5, 10, 15) IE_QR(1, 6, 11, 12) IE_QR(2, 7, 8, 13) IE_QR(3, 4, 9, 14)
    }
#undef IE_QR
#pragma unroll
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

/* streams (the 64-bit nonce): one per kind of value, so that no two draws share a (key, stream, counter) triple */
enum : uint64_t { RNG_LWE_KEY = 1, RNG_TLWE_KEY = 2, RNG_BK_MASK = 3, RNG_BK_NOISE = 4, RNG_KS_MASK = 5, RNG_KS_NOISE = 6,
                  RNG_ENC_MASK = 7, RNG_ENC_NOISE = 8, RNG_DERIVE = 9 };

/* Gaussian of standard deviation sigma (torus units) from two key-stream words, as a Torus32 (libtfhe gaussian32) */
#ifdef __CUDACC__
__device__ __forceinline__ int32_t gauss_torus(uint32_t w0, uint32_t w1, double sigma)
{
    const double u1 = ((double)w0 + 1.0) * (1.0 / 4294967296.0), u2 = (double)w1 * (1.0 / 4294967296.0);
    const double g = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2) * sigma;
    return (int32_t)(int64_t)((g - floor(g + 0.5)) * 4294967296.0);
}
#endif

/* host side (keygen.cu): 256-bit keys from the operating system, or derived from a 64-bit seed (tests only) */
int rng_keys_from_os(RngKeys &out);                 /* 0 on success */
void rng_keys_from_seed(uint64_t seed, RngKeys &out);

} // namespace ieache
#endif
