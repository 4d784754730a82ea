/*
 * oracle/exact_ref.c — blind rotation in exact integer arithmetic.  TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h).
 *
 * tfhe_bootstrap_woKS_FFT as SURVEY.md Appendix A states it (the algorithm Cloud/cloud.c reaches through
 * bootsAND / bootsXOR, cloud.c:30-43), with the external product tGswFFTExternMulToTLwe computed as schoolbook
 * negacyclic products of the gadget digits with the COEFFICIENT-domain bootstrapping key, in wrap-around 32-bit
 * integer arithmetic.  No transform, no floating point: this is what every FFT implementation of the step (libtfhe's,
 * the oracle's, the CUDA kernels') approximates, and equals as long as its rounding error stays below 1/2.  A
 * transform-based blind rotation that matches this function bit for bit has computed every one of its
 * n x (k+1) l x (k+1) polynomial products exactly; it is checked against arithmetic, not against another FFT.
 * (libtfhe's own last conversion is a truncating cast, SURVEY App. A, so the library itself may sit one unit of
 * 2^-32 below this value in some coefficients; that is 2^7 times below its own noise floor.)
 *
 * Cost: n (k+1)^2 l N^2 = 12.6 M multiply-adds per CMux step at the default ring; used at n <= 32, and for single
 * gates at n = 630.
 */
#include "tfhe_oracle.h"

#include <stdlib.h>
#include <string.h>

static int modswitch_2N(Torus32 phase, int N)
{
    /* modSwitchFromTorus32(phase, 2N) of App. A, for N a power of two */
    const uint64_t interv = ((UINT64_C(1) << 63) / (uint64_t)(2 * N)) * 2;
    const uint64_t phase64 = ((uint64_t)(uint32_t)phase << 32) + interv / 2;
    return (int)(phase64 / interv);
}

/* out += d (*) b  mod (X^N + 1, 2^32): d small signed digits, b torus coefficients */
static void negacyclic_mac(uint32_t *out, const int32_t *d, const Torus32 *b, int N)
{
    for (int a = 0; a < N; a++) {
        const uint32_t da = (uint32_t)d[a];
        if (!da) continue;
        uint32_t *lo = out + a;
        for (int j = 0; j < N - a; j++) lo[j] += da * (uint32_t)b[j];
        const Torus32 *hi = b + (N - a);
        for (int j = 0; j < a; j++) out[j] -= da * (uint32_t)hi[j];
    }
}

void o_bootstrap_woks_exact(const OKeySet *ks, Torus32 *out, Torus32 mu, const Torus32 *x)
{
    const OParams *p = o_keyset_params(ks);
    const int n = p->n, N = p->N, k = p->k, l = p->bk_l, kpl = (k + 1) * l, Bgbit = p->bk_Bgbit;
    const Torus32 *bk = o_bk_coef(ks); /* [n][kpl][k+1][N] */
    const uint32_t Bg = 1u << Bgbit, maskMod = Bg - 1, halfBg = Bg / 2;
    uint32_t offset = 0;
    for (int i = 1; i <= l; i++) offset += halfBg << (32 - i * Bgbit);

    Torus32 *acc = (Torus32 *)calloc((size_t)(k + 1) * N, sizeof(Torus32));
    Torus32 *rot = (Torus32 *)malloc(sizeof(Torus32) * (size_t)(k + 1) * N);
    int32_t *dec = (int32_t *)malloc(sizeof(int32_t) * (size_t)kpl * N);
    uint32_t *res = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(k + 1) * N);

    /* ACC = (0, X^{2N - bbar} * mu (1 + X + ... + X^{N-1})) */
    {
        const int a = (2 * N - modswitch_2N(x[n], N)) % (2 * N), ar = a % N;
        for (int j = 0; j < N; j++) {
            /* coefficient j of X^a * tv: tv[j - a] with a sign flip per wrap around X^N = -1 */
            const int neg = (j < ar) != (a >= N);
            acc[(size_t)k * N + j] = neg ? -mu : mu;
        }
    }
    for (int i = 0; i < n; i++) {
        const int a = modswitch_2N(x[i], N);
        if (a == 0) continue;
        const int ar = a % N, flip = a >= N;
        for (int q = 0; q <= k; q++) {
            const Torus32 *in = acc + (size_t)q * N;
            Torus32 *o = rot + (size_t)q * N;
            for (int j = 0; j < N; j++) {
                const Torus32 v = in[(j - ar + N) % N];
                o[j] = (((j < ar) != flip) ? -v : v) - in[j];
            }
        }
        for (int q = 0; q <= k; q++)
            for (int pp = 0; pp < l; pp++) {
                const int sh = 32 - (pp + 1) * Bgbit;
                int32_t *d = dec + (size_t)(q * l + pp) * N;
                for (int j = 0; j < N; j++)
                    d[j] = (int32_t)((((uint32_t)rot[(size_t)q * N + j] + offset) >> sh) & maskMod) - (int32_t)halfBg;
            }
        memset(res, 0, sizeof(uint32_t) * (size_t)(k + 1) * N);
        for (int r = 0; r < kpl; r++)
            for (int q = 0; q <= k; q++)
                negacyclic_mac(res + (size_t)q * N, dec + (size_t)r * N, bk + (((size_t)i * kpl + r) * (k + 1) + q) * N, N);
        for (size_t j = 0; j < (size_t)(k + 1) * N; j++) acc[j] = (Torus32)((uint32_t)acc[j] + res[j]);
    }
    for (int q = 0; q < k; q++) {
        out[q * N] = acc[(size_t)q * N];
        for (int j = 1; j < N; j++) out[q * N + j] = -acc[(size_t)q * N + N - j];
    }
    out[k * N] = acc[(size_t)k * N];
    free(acc); free(rot); free(dec); free(res);
}
