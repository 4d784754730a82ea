/*
 * oracle/tfhe_oracle.c — CPU restatement of libtfhe's gate bootstrapping (see tfhe_oracle.h).
 * TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (no libtfhe, no golden vectors in the reference).
 *
 * Reference call sites restated here (libtfhe itself is absent):
 *   boots{XOR,AND,NOT,COPY,CONSTANT}            Cloud/cloud.c:24,30,32,38,40,43,46,56,62
 *   bootsSymEncrypt / bootsSymDecrypt           Client1/alice.c:117, Cloud/cloud.c:712,823, Output/verif.c:93
 *   new_default_gate_bootstrapping_parameters   Keygen/keygen.c:22-23
 *   new_random_gate_bootstrapping_secret_keyset Keygen/keygen.c:30-36
 *   export_/import_ key and ciphertext files    Keygen/keygen.c:39-51, Cloud/cloud.c:656-766,826,900
 * Algorithm: SURVEY.md Appendix A (recalled from libtfhe master):
 *   gate = linear pre-combination + tfhe_bootstrap_FFT(mu = 1/8)
 *   bootstrap = modSwitch -> test vector rotate -> n x CMux(BK_i, X^{abar_i}) -> SampleExtract -> KeySwitch
 */
#include "tfhe_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ RNG */
typedef struct { uint64_t s[4]; int has_spare; double spare; } ORng;

static uint64_t splitmix64(uint64_t *x)
{
    uint64_t z = (*x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void rng_seed(ORng *r, uint64_t seed)
{
    for (int i = 0; i < 4; i++) r->s[i] = splitmix64(&seed);
    r->has_spare = 0;
}
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t rng_next(ORng *r) /* xoshiro256** */
{
    uint64_t *s = r->s;
    uint64_t result = rotl64(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3];
    s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
static inline Torus32 rng_torus(ORng *r) { return (Torus32)(uint32_t)(rng_next(r) >> 32); }
static inline int32_t rng_bit(ORng *r) { return (int32_t)(rng_next(r) >> 63); }
static double rng_gauss(ORng *r)
{
    if (r->has_spare) { r->has_spare = 0; return r->spare; }
    double u1, u2;
    do { u1 = (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); } while (u1 <= 0.0);
    u2 = (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0);
    double m = sqrt(-2.0 * log(u1));
    r->spare = m * sin(2.0 * M_PI * u2);
    r->has_spare = 1;
    return m * cos(2.0 * M_PI * u2);
}
/* libtfhe dtot32: real mod 1 -> Torus32 */
static inline Torus32 dtot32(double d)
{
    return (Torus32)(int64_t)((d - (double)(int64_t)d) * 4294967296.0);
}
static inline Torus32 gaussian32(ORng *r, Torus32 message, double sigma)
{
    return message + dtot32(sigma * rng_gauss(r));
}

/* ------------------------------------------------------------------ torus helpers (App. A) */
static inline Torus32 modSwitchToTorus32(int32_t mu, int32_t Msize)
{
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)Msize) * 2;
    uint64_t phase64 = (uint64_t)(int64_t)mu * interv;
    return (Torus32)(phase64 >> 32);
}
static inline int32_t modSwitchFromTorus32(Torus32 phase, int32_t Msize)
{
    uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)Msize) * 2;
    uint64_t half_interval = interv / 2;
    uint64_t phase64 = ((uint64_t)(uint32_t)phase << 32) + half_interval;
    return (int32_t)(phase64 / interv);
}

/* ------------------------------------------------------------------ parameters */
void o_params_default(OParams *p)
{
    /* tfhe master, 80 < lambda <= 128 branch (corroborated by the
     * 162304 = 64 x 2536 constant at Cloud/dragonfly_cipher_cloud.py:1295). */
    p->n = 630; p->N = 1024; p->k = 1; p->bk_l = 3; p->bk_Bgbit = 7;
    p->ks_t = 8; p->ks_basebit = 2;
    p->ks_stdev = pow(2.0, -15); p->bk_stdev = pow(2.0, -25); p->max_stdev = 0.012467;
}
void o_params_small(OParams *p, int32_t n_small)
{
    o_params_default(p);
    p->n = n_small;
}

/* ------------------------------------------------------------------ negacyclic FFT, double precision */
typedef struct {
    int N, M;          /* M = N/2 complex points */
    double *tw_re, *tw_im;   /* psi^j, j<M, psi = exp(i*pi/N) */
    double *w_re, *w_im;     /* exp(+2*pi*i*j/M), j<M/2 */
    double *st_re, *st_im;   /* per-stage contiguous twiddles for len = 8, 16, ..., M */
    int *rev;
} FftCtx;

static FftCtx g_ctx[8];
static int g_nctx = 0;

static const FftCtx *fft_ctx(int N)
{
    for (int i = 0; i < g_nctx; i++) if (g_ctx[i].N == N) return &g_ctx[i];
    const FftCtx *ret = NULL;
#ifdef _OPENMP
#pragma omp critical(o_fft_ctx)
#endif
    {
        for (int i = 0; i < g_nctx; i++) if (g_ctx[i].N == N) ret = &g_ctx[i];
        if (!ret && g_nctx < 8) {
            FftCtx *c = &g_ctx[g_nctx];
            int M = N / 2;
            c->N = N; c->M = M;
            c->tw_re = (double *)malloc(sizeof(double) * M); c->tw_im = (double *)malloc(sizeof(double) * M);
            c->w_re = (double *)malloc(sizeof(double) * (M / 2 + 1)); c->w_im = (double *)malloc(sizeof(double) * (M / 2 + 1));
            c->rev = (int *)malloc(sizeof(int) * M);
            for (int j = 0; j < M; j++) { c->tw_re[j] = cos(M_PI * j / N); c->tw_im[j] = sin(M_PI * j / N); }
            for (int j = 0; j < M / 2; j++) { c->w_re[j] = cos(2.0 * M_PI * j / M); c->w_im[j] = sin(2.0 * M_PI * j / M); }
            c->st_re = (double *)malloc(sizeof(double) * M); c->st_im = (double *)malloc(sizeof(double) * M);
            { int off = 0; for (int len = 8; len <= M; len <<= 1) { int half = len >> 1; for (int j = 0; j < half; j++) { c->st_re[off + j] = cos(2.0 * M_PI * j / len); c->st_im[off + j] = sin(2.0 * M_PI * j / len); } off += half; } }
            int lg = 0; while ((1 << lg) < M) lg++;
            for (int j = 0; j < M; j++) { int r = 0; for (int b = 0; b < lg; b++) if (j & (1 << b)) r |= 1 << (lg - 1 - b); c->rev[j] = r; }
            g_nctx++;
            ret = c;
        }
    }
    return ret;
}

/* in-place radix-2 DIT on bit-reversed input; sign = +1 uses exp(+2 pi i jk/M).
 * Twiddles of every stage are stored contiguously (st_re/st_im) so the inner loop vectorises (AVX2 via
 * -march=x86-64-v3): this is the CPU baseline of bench.py, it should not be needlessly slow. */
static void cfft(const FftCtx *c, double *restrict re, double *restrict im, int sign)
{
    const int M = c->M;
    /* first two stages (half = 1, 2) without multiplications by non-trivial twiddles */
    for (int base = 0; base < M; base += 2) {
        double ur = re[base], ui = im[base], xr = re[base + 1], xi = im[base + 1];
        re[base] = ur + xr; im[base] = ui + xi; re[base + 1] = ur - xr; im[base + 1] = ui - xi;
    }
    for (int base = 0; base < M; base += 4) {
        double ur = re[base], ui = im[base], xr = re[base + 2], xi = im[base + 2];
        re[base] = ur + xr; im[base] = ui + xi; re[base + 2] = ur - xr; im[base + 2] = ui - xi;
        /* twiddle exp(sign * i pi/2) = sign * i */
        ur = re[base + 1]; ui = im[base + 1]; xr = -sign * im[base + 3]; xi = sign * re[base + 3];
        re[base + 1] = ur + xr; im[base + 1] = ui + xi; re[base + 3] = ur - xr; im[base + 3] = ui - xi;
    }
    int off = 0;
    for (int len = 8; len <= M; len <<= 1) {
        const int half = len >> 1;
        const double *restrict wr = c->st_re + off, *restrict wi0 = c->st_im + off;
        for (int base = 0; base < M; base += len) {
            double *restrict ar = re + base, *restrict ai = im + base, *restrict br = re + base + half, *restrict bi = im + base + half;
            for (int j = 0; j < half; j++) {
                const double wi = sign * wi0[j];
                const double tr = br[j] * wr[j] - bi[j] * wi, ti = br[j] * wi + bi[j] * wr[j];
                const double ur = ar[j], ui = ai[j];
                ar[j] = ur + tr; ai[j] = ui + ti;
                br[j] = ur - tr; bi[j] = ui - ti;
            }
        }
        off += half;
    }
}
/* real coefficients p[0..N) -> M evaluations at psi^(4k+1) (libtfhe "LagrangeHalfC") */
static void nfft_fwd(const FftCtx *c, const double *p, double *ore, double *oim)
{
    int M = c->M;
    for (int j = 0; j < M; j++) {
        double zr = p[j], zi = p[j + M];
        int r = c->rev[j];
        ore[r] = zr * c->tw_re[j] - zi * c->tw_im[j];
        oim[r] = zr * c->tw_im[j] + zi * c->tw_re[j];
    }
    cfft(c, ore, oim, +1);
}
/* inverse: evaluations -> real coefficients (unscaled by 1/M inside; we scale here) */
static void nfft_inv(const FftCtx *c, const double *ire, const double *iim, double *p, double *sre, double *sim)
{
    int M = c->M;
    for (int j = 0; j < M; j++) { int r = c->rev[j]; sre[r] = ire[j]; sim[r] = iim[j]; }
    cfft(c, sre, sim, -1);
    double inv = 1.0 / M;
    for (int j = 0; j < M; j++) {
        double zr = sre[j] * inv, zi = sim[j] * inv;
        p[j] = zr * c->tw_re[j] + zi * c->tw_im[j];
        p[j + M] = zi * c->tw_re[j] - zr * c->tw_im[j];
    }
}

/* ------------------------------------------------------------------ key set */
struct OKeySet {
    OParams p;
    int has_secret;
    int32_t *lwe_key;   /* n */
    int32_t *tlwe_key;  /* k*N */
    Torus32 *bk;        /* [n][kpl][k+1][N] */
    Torus32 *ksk;       /* [kN][t][base][n+1] */
    double *bkfft_re, *bkfft_im; /* [n][kpl][k+1][N/2] */
};

const OParams *o_keyset_params(const OKeySet *ks) { return &ks->p; }
int o_keyset_has_secret(const OKeySet *ks) { return ks->has_secret; }
const int32_t *o_lwe_key(const OKeySet *ks) { return ks->lwe_key; }
const int32_t *o_tlwe_key(const OKeySet *ks) { return ks->tlwe_key; }
const Torus32 *o_bk_coef(const OKeySet *ks) { return ks->bk; }
const Torus32 *o_ksk(const OKeySet *ks) { return ks->ksk; }

static size_t bk_words(const OParams *p) { return (size_t)p->n * (p->k + 1) * p->bk_l * (p->k + 1) * p->N; }
static size_t ksk_words(const OParams *p) { return (size_t)p->k * p->N * p->ks_t * (1u << p->ks_basebit) * (p->n + 1); }

static OKeySet *keyset_alloc(const OParams *p, int with_secret)
{
    OKeySet *ks = (OKeySet *)calloc(1, sizeof(OKeySet));
    ks->p = *p;
    ks->has_secret = with_secret;
    if (with_secret) {
        ks->lwe_key = (int32_t *)calloc(p->n, sizeof(int32_t));
        ks->tlwe_key = (int32_t *)calloc((size_t)p->k * p->N, sizeof(int32_t));
    }
    ks->bk = (Torus32 *)calloc(bk_words(p), sizeof(Torus32));
    ks->ksk = (Torus32 *)calloc(ksk_words(p), sizeof(Torus32));
    return ks;
}
void o_keyset_free(OKeySet *ks)
{
    if (!ks) return;
    free(ks->lwe_key); free(ks->tlwe_key); free(ks->bk); free(ks->ksk);
    free(ks->bkfft_re); free(ks->bkfft_im);
    free(ks);
}

/* transform-domain copy of BK (libtfhe: new_LweBootstrappingKeyFFT, done at key load) */
static void keyset_build_fft(OKeySet *ks)
{
    const OParams *p = &ks->p;
    const FftCtx *c = fft_ctx(p->N);
    int M = p->N / 2;
    size_t npoly = (size_t)p->n * (p->k + 1) * p->bk_l * (p->k + 1);
    free(ks->bkfft_re); free(ks->bkfft_im);
    ks->bkfft_re = (double *)malloc(sizeof(double) * npoly * M);
    ks->bkfft_im = (double *)malloc(sizeof(double) * npoly * M);
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        double *tmp = (double *)malloc(sizeof(double) * p->N);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (long q = 0; q < (long)npoly; q++) {
            const Torus32 *src = ks->bk + (size_t)q * p->N;
            for (int j = 0; j < p->N; j++) tmp[j] = (double)src[j];
            nfft_fwd(c, tmp, ks->bkfft_re + (size_t)q * M, ks->bkfft_im + (size_t)q * M);
        }
        free(tmp);
    }
}

static void keyset_build_fft(OKeySet *ks);
OKeySet *o_keyset_from_arrays(const OParams *p, const int32_t *lwe_key, const int32_t *tlwe_key,
                              const Torus32 *bk, const Torus32 *ksk)
{
    OKeySet *ks = keyset_alloc(p, lwe_key != NULL);
    if (lwe_key) memcpy(ks->lwe_key, lwe_key, sizeof(int32_t) * p->n);
    if (lwe_key && tlwe_key) memcpy(ks->tlwe_key, tlwe_key, sizeof(int32_t) * p->k * p->N);
    memcpy(ks->bk, bk, sizeof(Torus32) * bk_words(p));
    memcpy(ks->ksk, ksk, sizeof(Torus32) * ksk_words(p));
    keyset_build_fft(ks);
    return ks;
}

OKeySet *o_keygen(const OParams *p, uint64_t seed)
{
    OKeySet *ks = keyset_alloc(p, 1);
    ORng rng; rng_seed(&rng, seed);
    const int n = p->n, N = p->N, k = p->k, l = p->bk_l, kpl = (k + 1) * l;
    const FftCtx *c = fft_ctx(N);
    const int M = N / 2;
    for (int i = 0; i < n; i++) ks->lwe_key[i] = rng_bit(&rng);
    for (int i = 0; i < k * N; i++) ks->tlwe_key[i] = rng_bit(&rng);

    /* transform of the TLWE key polynomials, for a * s' products */
    double *sre = (double *)malloc(sizeof(double) * k * M), *sim = (double *)malloc(sizeof(double) * k * M);
    double *tmp = (double *)malloc(sizeof(double) * N), *are = (double *)malloc(sizeof(double) * M), *aim = (double *)malloc(sizeof(double) * M);
    double *pre = (double *)malloc(sizeof(double) * M), *pim = (double *)malloc(sizeof(double) * M);
    double *s1 = (double *)malloc(sizeof(double) * M), *s2 = (double *)malloc(sizeof(double) * M);
    for (int q = 0; q < k; q++) {
        for (int j = 0; j < N; j++) tmp[j] = (double)ks->tlwe_key[q * N + j];
        nfft_fwd(c, tmp, sre + q * M, sim + q * M);
    }
    /* BK_i = TGSW_{s'}(s_i): kpl TLWE encryptions of 0, plus s_i * h_p on the diagonal */
    for (int i = 0; i < n; i++) {
        for (int r = 0; r < kpl; r++) {
            Torus32 *row = ks->bk + ((size_t)i * kpl + r) * (k + 1) * N;
            Torus32 *b = row + (size_t)k * N;
            for (int j = 0; j < N; j++) b[j] = gaussian32(&rng, 0, p->bk_stdev);
            for (int q = 0; q < k; q++) {
                Torus32 *a = row + (size_t)q * N;
                for (int j = 0; j < N; j++) { a[j] = rng_torus(&rng); tmp[j] = (double)a[j]; }
                nfft_fwd(c, tmp, are, aim);
                for (int f = 0; f < M; f++) {
                    pre[f] = are[f] * sre[q * M + f] - aim[f] * sim[q * M + f];
                    pim[f] = are[f] * sim[q * M + f] + aim[f] * sre[q * M + f];
                }
                nfft_inv(c, pre, pim, tmp, s1, s2);
                for (int j = 0; j < N; j++) b[j] += (Torus32)(uint32_t)(int64_t)llrint(tmp[j]);
            }
            int q = r / l, pp = r % l;
            Torus32 h = (Torus32)(1u << (32 - (pp + 1) * p->bk_Bgbit));
            row[(size_t)q * N + 0] += ks->lwe_key[i] * h;
        }
    }
    /* KSK[i][j][d] = LWE_s( d * s'_i * 2^(32-(j+1)basebit) ); d = 0 is a noiseless trivial 0 */
    const int t = p->ks_t, base = 1 << p->ks_basebit;
    for (int i = 0; i < k * N; i++)
        for (int j = 0; j < t; j++)
            for (int d = 0; d < base; d++) {
                Torus32 *s = ks->ksk + (((size_t)i * t + j) * base + d) * (n + 1);
                if (d == 0) { memset(s, 0, sizeof(Torus32) * (n + 1)); continue; }
                Torus32 msg = (Torus32)((uint32_t)(d * ks->tlwe_key[i]) << (32 - (j + 1) * p->ks_basebit));
                Torus32 acc = gaussian32(&rng, msg, p->ks_stdev);
                for (int q = 0; q < n; q++) { s[q] = rng_torus(&rng); acc += s[q] * ks->lwe_key[q]; }
                s[n] = acc;
            }
    free(sre); free(sim); free(tmp); free(are); free(aim); free(pre); free(pim); free(s1); free(s2);
    keyset_build_fft(ks);
    return ks;
}

/* ------------------------------------------------------------------ encrypt / decrypt */
void o_sym_encrypt(const OKeySet *ks, const int32_t *bits, size_t count, Torus32 *out, uint64_t seed)
{
    /* bootsSymEncrypt: mu = +-1/8, noise alpha_min of in_out_params (Client1/alice.c:117) */
    const int n = ks->p.n;
    const Torus32 mu = modSwitchToTorus32(1, 8);
    ORng rng; rng_seed(&rng, seed);
    for (size_t c = 0; c < count; c++) {
        Torus32 *s = out + c * (n + 1);
        Torus32 b = gaussian32(&rng, bits[c] ? mu : -mu, ks->p.ks_stdev);
        for (int i = 0; i < n; i++) { s[i] = rng_torus(&rng); b += s[i] * ks->lwe_key[i]; }
        s[n] = b;
    }
}
void o_phase(const OKeySet *ks, const Torus32 *samples, size_t count, Torus32 *phases)
{
    const int n = ks->p.n;
    for (size_t c = 0; c < count; c++) {
        const Torus32 *s = samples + c * (n + 1);
        Torus32 ph = s[n];
        for (int i = 0; i < n; i++) ph -= s[i] * ks->lwe_key[i];
        phases[c] = ph;
    }
}
void o_sym_decrypt(const OKeySet *ks, const Torus32 *samples, size_t count, int32_t *bits)
{
    /* bootsSymDecrypt: phase > 0 (Output/verif.c:93) */
    for (size_t c = 0; c < count; c++) {
        Torus32 ph;
        o_phase(ks, samples + c * (ks->p.n + 1), 1, &ph);
        bits[c] = ph > 0 ? 1 : 0;
    }
}
void o_phase_extracted(const OKeySet *ks, const Torus32 *samples, size_t count, Torus32 *phases)
{
    const int kN = ks->p.k * ks->p.N;
    for (size_t c = 0; c < count; c++) {
        const Torus32 *s = samples + c * (kN + 1);
        Torus32 ph = s[kN];
        for (int i = 0; i < kN; i++) ph -= s[i] * ks->tlwe_key[i];
        phases[c] = ph;
    }
}

/* ------------------------------------------------------------------ bootstrap */
static uint64_t g_bootstraps = 0;
void o_stats_reset(void) { g_bootstraps = 0; }
uint64_t o_stats_bootstraps(void) { return g_bootstraps; }

/* X^a * P for a in [0, 2N) (torusPolynomialMulByXai) */
static void mul_by_xai(Torus32 *out, int a, const Torus32 *in, int N)
{
    if (a < N) {
        for (int i = 0; i < a; i++) out[i] = -in[i - a + N];
        for (int i = a; i < N; i++) out[i] = in[i - a];
    } else {
        int aa = a - N;
        for (int i = 0; i < aa; i++) out[i] = in[i - aa + N];
        for (int i = aa; i < N; i++) out[i] = -in[i - aa];
    }
}
/* (X^a - 1) * P (torusPolynomialMulByXaiMinusOne) */
static void mul_by_xai_minus_one(Torus32 *out, int a, const Torus32 *in, int N)
{
    if (a < N) {
        for (int i = 0; i < a; i++) out[i] = -in[i - a + N] - in[i];
        for (int i = a; i < N; i++) out[i] = in[i - a] - in[i];
    } else {
        int aa = a - N;
        for (int i = 0; i < aa; i++) out[i] = in[i - aa + N] - in[i];
        for (int i = aa; i < N; i++) out[i] = -in[i - aa] - in[i];
    }
}

void o_bootstrap_woks(const OKeySet *ks, Torus32 *out, Torus32 mu, const Torus32 *x)
{
    const OParams *p = &ks->p;
    const int n = p->n, N = p->N, k = p->k, l = p->bk_l, kpl = (k + 1) * l, Bgbit = p->bk_Bgbit;
    const int M = N / 2, Nx2 = 2 * N;
    const FftCtx *c = fft_ctx(N);
    const uint32_t Bg = 1u << Bgbit, maskMod = Bg - 1, halfBg = Bg / 2;
    uint32_t offset = 0;
    for (int i = 1; i <= l; i++) offset += halfBg << (32 - i * Bgbit); /* tGswParams::offset */

    Torus32 *acc = (Torus32 *)malloc(sizeof(Torus32) * (k + 1) * N);
    Torus32 *rot = (Torus32 *)malloc(sizeof(Torus32) * (k + 1) * N);
    double *dec = (double *)malloc(sizeof(double) * N);
    double *dre = (double *)malloc(sizeof(double) * kpl * M), *dim = (double *)malloc(sizeof(double) * kpl * M);
    double *tre = (double *)malloc(sizeof(double) * M), *tim = (double *)malloc(sizeof(double) * M);
    double *s1 = (double *)malloc(sizeof(double) * M), *s2 = (double *)malloc(sizeof(double) * M);
    double *res = (double *)malloc(sizeof(double) * N);

    /* test vector mu * (1 + X + ... + X^{N-1}) rotated by X^{2N - bbar}; mask = 0 */
    int barb = modSwitchFromTorus32(x[n], Nx2);
    memset(acc, 0, sizeof(Torus32) * k * N);
    {
        Torus32 *tv = rot; /* scratch */
        for (int j = 0; j < N; j++) tv[j] = mu;
        if (barb != 0) mul_by_xai(acc + (size_t)k * N, Nx2 - barb, tv, N);
        else memcpy(acc + (size_t)k * N, tv, sizeof(Torus32) * N);
    }
    /* blind rotate: ACC <- ACC + BK_i (.) ((X^{abar_i} - 1) ACC) */
    for (int i = 0; i < n; i++) {
        int bara = modSwitchFromTorus32(x[i], Nx2);
        if (bara == 0) continue;
        for (int q = 0; q <= k; q++) mul_by_xai_minus_one(rot + (size_t)q * N, bara, acc + (size_t)q * N, N);
        /* gadget decomposition (tGswTorus32PolynomialDecompH) + forward transform */
        for (int q = 0; q <= k; q++)
            for (int pp = 0; pp < l; pp++) {
                int sh = 32 - (pp + 1) * Bgbit;
                for (int j = 0; j < N; j++) {
                    uint32_t v = (uint32_t)rot[(size_t)q * N + j] + offset;
                    dec[j] = (double)((int32_t)((v >> sh) & maskMod) - (int32_t)halfBg);
                }
                nfft_fwd(c, dec, dre + (size_t)(q * l + pp) * M, dim + (size_t)(q * l + pp) * M);
            }
        /* tLweFFTAddMulRTo over the kpl rows, per output polynomial */
        for (int q = 0; q <= k; q++) {
            for (int f = 0; f < M; f++) { tre[f] = 0; tim[f] = 0; }
            for (int r = 0; r < kpl; r++) {
                const double *bre = ks->bkfft_re + (((size_t)i * kpl + r) * (k + 1) + q) * M;
                const double *bim = ks->bkfft_im + (((size_t)i * kpl + r) * (k + 1) + q) * M;
                const double *xr = dre + (size_t)r * M, *xi = dim + (size_t)r * M;
                for (int f = 0; f < M; f++) {
                    tre[f] += xr[f] * bre[f] - xi[f] * bim[f];
                    tim[f] += xr[f] * bim[f] + xi[f] * bre[f];
                }
            }
            nfft_inv(c, tre, tim, res, s1, s2);
            for (int j = 0; j < N; j++)
                acc[(size_t)q * N + j] += (Torus32)(uint32_t)(int64_t)llrint(res[j]);
        }
    }
    /* tLweExtractLweSample, index 0 */
    for (int q = 0; q < k; q++) {
        out[q * N] = acc[(size_t)q * N];
        for (int j = 1; j < N; j++) out[q * N + j] = -acc[(size_t)q * N + N - j];
    }
    out[k * N] = acc[(size_t)k * N];
    free(acc); free(rot); free(dec); free(dre); free(dim); free(tre); free(tim); free(s1); free(s2); free(res);
}

void o_keyswitch(const OKeySet *ks, Torus32 *out, const Torus32 *u)
{
    const OParams *p = &ks->p;
    const int n = p->n, kN = p->k * p->N, t = p->ks_t, basebit = p->ks_basebit, base = 1 << basebit;
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));
    const uint32_t mask = base - 1;
    Torus32 *res = (Torus32 *)calloc(n + 1, sizeof(Torus32));
    res[n] = u[kN];
    for (int i = 0; i < kN; i++) {
        uint32_t aibar = (uint32_t)u[i] + prec_offset;
        for (int j = 0; j < t; j++) {
            uint32_t d = (aibar >> (32 - (j + 1) * basebit)) & mask;
            if (d != 0) {
                const Torus32 *row = ks->ksk + (((size_t)i * t + j) * base + d) * (n + 1);
                for (int q = 0; q <= n; q++) res[q] -= row[q];
            }
        }
    }
    memcpy(out, res, sizeof(Torus32) * (n + 1));
    free(res);
}

static void bootstrap(const OKeySet *ks, Torus32 *out, Torus32 mu, const Torus32 *x)
{
    const int kN = ks->p.k * ks->p.N;
    Torus32 *u = (Torus32 *)malloc(sizeof(Torus32) * (kN + 1));
    o_bootstrap_woks(ks, u, mu, x);
    o_keyswitch(ks, out, u);
    free(u);
#ifdef _OPENMP
#pragma omp atomic
#endif
    g_bootstraps++;
}

/* linear pre-combination table: temp = (0, cst*mu) + sa*ca + sb*cb */
static const int8_t k_gate_lin[10][3] = {
    /* NAND  */ {+1, -1, -1}, /* OR    */ {+1, +1, +1}, /* AND   */ {-1, +1, +1},
    /* XOR   */ {+2, +2, +2}, /* XNOR  */ {-2, -2, -2}, /* NOR   */ {-1, -1, -1},
    /* ANDNY */ {-1, -1, +1}, /* ANDYN */ {-1, +1, -1}, /* ORNY  */ {+1, -1, +1},
    /* ORYN  */ {+1, +1, -1}};

void o_gate(const OKeySet *ks, int op, Torus32 *out, const Torus32 *a, const Torus32 *b, const Torus32 *c, int32_t imm)
{
    const int n = ks->p.n;
    const Torus32 mu = modSwitchToTorus32(1, 8);
    if (op == O_CONST) { memset(out, 0, sizeof(Torus32) * n); out[n] = imm ? mu : -mu; return; }
    if (op == O_COPY) { if (out != a) memmove(out, a, sizeof(Torus32) * (n + 1)); return; }
    if (op == O_NOT) { for (int i = 0; i <= n; i++) out[i] = -a[i]; return; }
    Torus32 *tmp = (Torus32 *)malloc(sizeof(Torus32) * (n + 1));
    if (op == O_MUX) {
        /* bootsMUX: two blind rotations, one key switch */
        const int kN = ks->p.k * ks->p.N;
        Torus32 *u1 = (Torus32 *)malloc(sizeof(Torus32) * (kN + 1)), *u2 = (Torus32 *)malloc(sizeof(Torus32) * (kN + 1));
        for (int i = 0; i <= n; i++) tmp[i] = a[i] + b[i];
        tmp[n] += -mu;
        o_bootstrap_woks(ks, u1, mu, tmp);
        for (int i = 0; i <= n; i++) tmp[i] = -a[i] + c[i];
        tmp[n] += -mu;
        o_bootstrap_woks(ks, u2, mu, tmp);
        for (int i = 0; i <= kN; i++) u1[i] += u2[i];
        u1[kN] += mu;
        o_keyswitch(ks, out, u1);
        free(u1); free(u2); free(tmp);
#ifdef _OPENMP
#pragma omp atomic
#endif
        g_bootstraps += 2;
        return;
    }
    const int8_t *g = k_gate_lin[op];
    for (int i = 0; i <= n; i++) tmp[i] = g[1] * a[i] + g[2] * b[i];
    tmp[n] += g[0] * mu;
    bootstrap(ks, out, mu, tmp);
    free(tmp);
}

int o_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void o_gate_batch(const OKeySet *ks, int op, Torus32 *out, const Torus32 *a, const Torus32 *b, const Torus32 *c,
                  size_t count, int threads)
{
    const size_t w = ks->p.n + 1;
    (void)fft_ctx(ks->p.N);
#ifdef _OPENMP
    if (threads <= 0) threads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (long g = 0; g < (long)count; g++)
        o_gate(ks, op, out + g * w, a ? a + g * w : NULL, b ? b + g * w : NULL, c ? c + g * w : NULL, 0);
}

/* ------------------------------------------------------------------ libtfhe-format files (recalled) */
/* type ids as recalled from libtfhe's tfhe_io; the reader reports, not rejects, other values */
enum { UID_LWE_KEY = 43, UID_LWE_SAMPLE = 42, UID_KS_KEY = 200, UID_TLWE_SAMPLE = 45, UID_TGSW_KEY = 47,
       UID_TGSW_SAMPLE = 48, UID_BK = 201 };

static void write_params(FILE *f, const OParams *p)
{
    fprintf(f, "-----BEGIN GATEBOOTSPARAMS-----\nks_basebit: %d\nks_t: %d\n-----END GATEBOOTSPARAMS-----\n", p->ks_basebit, p->ks_t);
    fprintf(f, "-----BEGIN LWEPARAMS-----\nalpha_max: %.17g\nalpha_min: %.17g\nn: %d\n-----END LWEPARAMS-----\n", p->max_stdev, p->ks_stdev, p->n);
    fprintf(f, "-----BEGIN TLWEPARAMS-----\nN: %d\nalpha_max: %.17g\nalpha_min: %.17g\nk: %d\n-----END TLWEPARAMS-----\n", p->N, p->max_stdev, p->bk_stdev, p->k);
    fprintf(f, "-----BEGIN TGSWPARAMS-----\nBgbit: %d\nl: %d\n-----END TGSWPARAMS-----\n", p->bk_Bgbit, p->bk_l);
}
static int read_params(FILE *f, OParams *p)
{
    char line[256], title[64] = "";
    int seen = 0;
    memset(p, 0, sizeof(*p));
    for (;;) {
        long pos = ftell(f);
        int ch = fgetc(f);
        if (ch == EOF) break;
        ungetc(ch, f);
        if (ch != '-' && title[0] == 0) { fseek(f, pos, SEEK_SET); break; } /* binary part starts */
        if (!fgets(line, sizeof line, f)) break;
        if (!strncmp(line, "-----BEGIN ", 11)) { sscanf(line + 11, "%63[A-Z]", title); continue; }
        if (!strncmp(line, "-----END ", 9)) { title[0] = 0; seen++; continue; }
        char key[64]; double val;
        if (sscanf(line, " %63[^:]: %lf", key, &val) != 2) continue;
        if (!strcmp(title, "GATEBOOTSPARAMS")) { if (!strcmp(key, "ks_basebit")) p->ks_basebit = (int)val; if (!strcmp(key, "ks_t")) p->ks_t = (int)val; }
        else if (!strcmp(title, "LWEPARAMS")) { if (!strcmp(key, "n")) p->n = (int)val; if (!strcmp(key, "alpha_min")) p->ks_stdev = val; if (!strcmp(key, "alpha_max")) p->max_stdev = val; }
        else if (!strcmp(title, "TLWEPARAMS")) { if (!strcmp(key, "N")) p->N = (int)val; if (!strcmp(key, "k")) p->k = (int)val; if (!strcmp(key, "alpha_min")) p->bk_stdev = val; }
        else if (!strcmp(title, "TGSWPARAMS")) { if (!strcmp(key, "l")) p->bk_l = (int)val; if (!strcmp(key, "Bgbit")) p->bk_Bgbit = (int)val; }
    }
    return (seen >= 4 && p->n > 0 && p->N > 0 && p->bk_l > 0 && p->ks_t > 0) ? 0 : -1;
}
static void wr32(FILE *f, int32_t v) { fwrite(&v, 4, 1, f); }
static int rd32(FILE *f, int32_t *v) { return fread(v, 4, 1, f) == 1 ? 0 : -1; }

static void write_bk(FILE *f, const OKeySet *ks)
{
    const OParams *p = &ks->p;
    const int n = p->n, N = p->N, k = p->k, kpl = (k + 1) * p->bk_l, t = p->ks_t, base = 1 << p->ks_basebit;
    const double var_ks = p->ks_stdev * p->ks_stdev, var_bk = p->bk_stdev * p->bk_stdev;
    wr32(f, UID_BK);
    wr32(f, UID_KS_KEY);
    for (size_t s = 0; s < (size_t)k * N * t * base; s++) {
        wr32(f, UID_LWE_SAMPLE);
        fwrite(ks->ksk + s * (n + 1), 4, n + 1, f);
        double v = (s % base) ? var_ks : 0.0;
        fwrite(&v, 8, 1, f);
    }
    for (int i = 0; i < n; i++) {
        wr32(f, UID_TGSW_SAMPLE);
        for (int r = 0; r < kpl; r++) {
            wr32(f, UID_TLWE_SAMPLE);
            fwrite(ks->bk + ((size_t)i * kpl + r) * (k + 1) * N, 4, (size_t)(k + 1) * N, f);
            fwrite(&var_bk, 8, 1, f);
        }
    }
}
static int read_bk(FILE *f, OKeySet *ks)
{
    const OParams *p = &ks->p;
    const int n = p->n, N = p->N, k = p->k, kpl = (k + 1) * p->bk_l, t = p->ks_t, base = 1 << p->ks_basebit;
    int32_t id; double v;
    if (rd32(f, &id) || rd32(f, &id)) return -1;
    for (size_t s = 0; s < (size_t)k * N * t * base; s++) {
        if (rd32(f, &id)) return -1;
        if (fread(ks->ksk + s * (n + 1), 4, n + 1, f) != (size_t)(n + 1)) return -1;
        if (fread(&v, 8, 1, f) != 1) return -1;
    }
    for (int i = 0; i < n; i++) {
        if (rd32(f, &id)) return -1;
        for (int r = 0; r < kpl; r++) {
            if (rd32(f, &id)) return -1;
            if (fread(ks->bk + ((size_t)i * kpl + r) * (k + 1) * N, 4, (size_t)(k + 1) * N, f) != (size_t)(k + 1) * N) return -1;
            if (fread(&v, 8, 1, f) != 1) return -1;
        }
    }
    return 0;
}
int o_write_cloud_key(const OKeySet *ks, const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    write_params(f, &ks->p);
    write_bk(f, ks);
    fclose(f);
    return 0;
}
int o_write_secret_key(const OKeySet *ks, const char *path)
{
    if (!ks->has_secret) return -1;
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    write_params(f, &ks->p);
    write_bk(f, ks);
    wr32(f, UID_LWE_KEY);
    fwrite(ks->lwe_key, 4, ks->p.n, f);
    wr32(f, UID_TGSW_KEY);
    fwrite(ks->tlwe_key, 4, (size_t)ks->p.k * ks->p.N, f);
    fclose(f);
    return 0;
}
OKeySet *o_read_key(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    OParams p;
    if (read_params(f, &p)) { fclose(f); return NULL; }
    OKeySet *ks = keyset_alloc(&p, 1);
    if (read_bk(f, ks)) { fclose(f); o_keyset_free(ks); return NULL; }
    int32_t id;
    if (rd32(f, &id) == 0) {
        if (fread(ks->lwe_key, 4, p.n, f) != (size_t)p.n || rd32(f, &id) ||
            fread(ks->tlwe_key, 4, (size_t)p.k * p.N, f) != (size_t)p.k * p.N) { fclose(f); o_keyset_free(ks); return NULL; }
    } else {
        free(ks->lwe_key); free(ks->tlwe_key); ks->lwe_key = NULL; ks->tlwe_key = NULL; ks->has_secret = 0;
    }
    fclose(f);
    keyset_build_fft(ks);
    return ks;
}
int o_write_samples(const OKeySet *ks, const Torus32 *samples, size_t count, const char *path, int append)
{
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return -1;
    const int n = ks->p.n;
    const double var = ks->p.ks_stdev * ks->p.ks_stdev;
    for (size_t c = 0; c < count; c++) {
        wr32(f, UID_LWE_SAMPLE);
        fwrite(samples + c * (n + 1), 4, n + 1, f);
        fwrite(&var, 8, 1, f);
    }
    fclose(f);
    return 0;
}
long o_read_samples(const OKeySet *ks, Torus32 *samples, size_t max_count, const char *path, size_t skip)
{
    FILE *f = fopen(path, "rb");
    if (!f) return -1;
    const int n = ks->p.n;
    const size_t rec = 4 + 4 * (size_t)(n + 1) + 8;
    if (fseek(f, (long)(skip * rec), SEEK_SET)) { fclose(f); return -1; }
    long got = 0;
    for (size_t c = 0; c < max_count; c++) {
        int32_t id; double v;
        if (rd32(f, &id)) break;
        if (fread(samples + c * (n + 1), 4, n + 1, f) != (size_t)(n + 1)) break;
        if (fread(&v, 8, 1, f) != 1) break;
        got++;
    }
    fclose(f);
    return got;
}
