/*
 * br_warp.h — per-lane building blocks of the warp-per-gate blind rotation: the same negacyclic transform as
 * br_core.h (512 complex points z_j = p_j + i p_{j+512} evaluated at psi^(4K+1), psi = exp(i pi/1024)), laid out for
 * ONE warp with 16 points per lane, so that a transform needs one shared-memory exchange and one shuffle stage
 * instead of two shared-memory exchanges (160 instead of 256 LSU wavefronts) and no CTA barrier.
 *
 *   pass 1 (registers): lane L holds z[L + 32 m], m = 0..15; radix-16 evaluation of P_L(Z) = sum_m z[L+32m] Z^m at the
 *                       16 roots of Z^16 = i, Z_r = s1 w16^brev4(r), s1 = exp(i pi/32) (immediates)
 *   exchange          : element r*33 + L written by lane L; lane l = 16 Lpar + k reads k*33 + Lpar + 2 m'
 *   pass 2 (registers): radix-16 evaluation of R_Lpar(V) = sum_m' Q[Lpar+2m'][k] V^m' at the roots of V^16 = Z_k,
 *                       base exp(i pi (1 + 4 kappa)/512), kappa = brev4(k)
 *   final stage       : P(Y) = R_0(Y^2) +- Y R_1(Y^2) between lanes l and l^16; each lane does 8 of the 16
 *                       butterflies (lane with Lpar computes r' = s + 8 Lpar, s = 0..7) after receiving 8 values
 *   result            : lane l, register s (+8 for the minus root) holds evaluation
 *                       K = brev4(k) + 16 brev4(s + 8 Lpar) + 256 sigma
 *
 * Everything is __host__ __device__: tests/emul runs the identical arithmetic on the CPU against the oracle.
 * Replaces the same libtfhe code as br_core.h (tfhe_blindRotate_FFT / tGswFFTExternMulToTLwe, reached from
 * Cloud/cloud.c:30-43).
 */
#ifndef IEACHE_BR_WARP_H
#define IEACHE_BR_WARP_H

#include "br_core.h"

namespace ieache {

constexpr int kWarpBufElems = 16 * 33; /* exchange buffer of one warp: 16 rows of 32 + 1 padding element */

IE_HD int brev4(int r) { return ((r & 1) << 3) | ((r & 2) << 1) | ((r & 4) >> 1) | ((r >> 3) & 1); }

/* twiddles of one radix-16 pass with base s: s^8, s^4, s^2, s^2 e^{i pi/4}, s, s e^{i pi/4}, s e^{i pi/8}, s e^{i 3pi/8} */
struct Tw16 { double z8r, z8i, z4r, z4i, z2r, z2i, z2qr, z2qi, z1r, z1i, z1qr, z1qi, z1hr, z1hi, z1hqr, z1hqi; };

/* forward radix-16 pass: x[m] natural order -> x[r] = value at the root s w16^brev4(r) */
IE_HD void pass16_fwd(double (&xr)[16], double (&xi)[16], const Tw16 &w)
{
#pragma unroll
    for (int m = 0; m < 8; m++) bf(xr[m], xi[m], xr[m + 8], xi[m + 8], w.z8r, w.z8i);
#pragma unroll
    for (int m = 0; m < 4; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4r, w.z4i);
#pragma unroll
    for (int m = 8; m < 12; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], -w.z4i, w.z4r);
#pragma unroll
    for (int m = 0; m < 2; m++) {
        bf(xr[m], xi[m], xr[m + 2], xi[m + 2], w.z2r, w.z2i);
        bf(xr[4 + m], xi[4 + m], xr[6 + m], xi[6 + m], -w.z2i, w.z2r);
        bf(xr[8 + m], xi[8 + m], xr[10 + m], xi[10 + m], w.z2qr, w.z2qi);
        bf(xr[12 + m], xi[12 + m], xr[14 + m], xi[14 + m], -w.z2qi, w.z2qr);
    }
    bf(xr[0], xi[0], xr[1], xi[1], w.z1r, w.z1i);
    bf(xr[2], xi[2], xr[3], xi[3], -w.z1i, w.z1r);
    bf(xr[4], xi[4], xr[5], xi[5], w.z1qr, w.z1qi);
    bf(xr[6], xi[6], xr[7], xi[7], -w.z1qi, w.z1qr);
    bf(xr[8], xi[8], xr[9], xi[9], w.z1hr, w.z1hi);
    bf(xr[10], xi[10], xr[11], xi[11], -w.z1hi, w.z1hr);
    bf(xr[12], xi[12], xr[13], xi[13], w.z1hqr, w.z1hqi);
    bf(xr[14], xi[14], xr[15], xi[15], -w.z1hqi, w.z1hqr);
}
/* stages 2..4 of pass16_fwd (everything after the m / m+8 butterflies) */
IE_HD void pass16_fwd_tail(double (&xr)[16], double (&xi)[16], const Tw16 &w)
{
#pragma unroll
    for (int m = 0; m < 4; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4r, w.z4i);
#pragma unroll
    for (int m = 8; m < 12; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], -w.z4i, w.z4r);
#pragma unroll
    for (int m = 0; m < 2; m++) {
        bf(xr[m], xi[m], xr[m + 2], xi[m + 2], w.z2r, w.z2i);
        bf(xr[4 + m], xi[4 + m], xr[6 + m], xi[6 + m], -w.z2i, w.z2r);
        bf(xr[8 + m], xi[8 + m], xr[10 + m], xi[10 + m], w.z2qr, w.z2qi);
        bf(xr[12 + m], xi[12 + m], xr[14 + m], xi[14 + m], -w.z2qi, w.z2qr);
    }
    bf(xr[0], xi[0], xr[1], xi[1], w.z1r, w.z1i);
    bf(xr[2], xi[2], xr[3], xi[3], -w.z1i, w.z1r);
    bf(xr[4], xi[4], xr[5], xi[5], w.z1qr, w.z1qi);
    bf(xr[6], xi[6], xr[7], xi[7], -w.z1qi, w.z1qr);
    bf(xr[8], xi[8], xr[9], xi[9], w.z1hr, w.z1hi);
    bf(xr[10], xi[10], xr[11], xi[11], -w.z1hi, w.z1hr);
    bf(xr[12], xi[12], xr[13], xi[13], w.z1hqr, w.z1hqi);
    bf(xr[14], xi[14], xr[15], xi[15], -w.z1hqi, w.z1hqr);
}
/* Pass 1 of a forward transform straight from the gadget digits (the 16-point form of br_core.h's
 * pass1_fwd_from_digits).  cc[h] = coefficient + decomposition offset, h = m for the real part of point m and 16 + m
 * for its imaginary part; digit = ((cc >> shift) & mask) - half.  The first stage multiplies by (1 + i)/sqrt 2, so the
 * difference and the sum of the two small integers are taken before the conversion: 4 FMAs per butterfly instead of
 * 6, and the -half of the two digits cancels in the difference. */
IE_HD void pass16_fwd_from_digits(const uint32_t (&cc)[32], int shift, uint32_t mask, int32_t half, double (&xr)[16], double (&xi)[16],
                                  const Tw16 &w)
{
    const double c = 0.70710678118654752440;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int32_t ar_i = (int32_t)((cc[m] >> shift) & mask) - half, ai_i = (int32_t)((cc[16 + m] >> shift) & mask) - half;
        const int32_t br_i = (int32_t)((cc[m + 8] >> shift) & mask), bi_i = (int32_t)((cc[24 + m] >> shift) & mask);
        const double ar = (double)ar_i, ai = (double)ai_i;
        const double u = (double)(br_i - bi_i), v = (double)(br_i + bi_i - 2 * half);
        xr[m] = fma(c, u, ar); xi[m] = fma(c, v, ai);
        xr[m + 8] = fma(-c, u, ar); xi[m + 8] = fma(-c, v, ai);
    }
    pass16_fwd_tail(xr, xi, w);
}
/* The same pass from digits that are left IN PLACE in the word: ch[h] = (coefficient + decomposition offset) >> 1, so
 * digit p sits at bits [sp, sp + Bgbit) with sp = 31 - (p + 1) Bgbit and every sum / difference of two fields fits an
 * int32.  field = ch & (mask << sp) is the digit times 2^sp: no shift per digit, and the power of two is taken out
 * again by the key (the rows of digit p are stored times 2^-sp, w12 layout) — scaling by a power of two commutes with
 * every rounding on the way, so the result is bit-identical to the unscaled transform.
 * M = mask << sp, H = half << sp. */
IE_HD void pass16_fwd_from_fields(const uint32_t (&ch)[32], uint32_t M, uint32_t H, double (&xr)[16], double (&xi)[16], const Tw16 &w)
{
    const double c = 0.70710678118654752440;
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const int32_t ar_i = (int32_t)((ch[m] & M) - H), ai_i = (int32_t)((ch[16 + m] & M) - H);
        const uint32_t br_u = ch[m + 8] & M, bi_u = ch[24 + m] & M;
        const double ar = (double)ar_i, ai = (double)ai_i;
        const double u = (double)(int32_t)(br_u - bi_u), v = (double)(int32_t)(br_u + bi_u - 2u * H);
        xr[m] = fma(c, u, ar); xi[m] = fma(c, v, ai);
        xr[m + 8] = fma(-c, u, ar); xi[m + 8] = fma(-c, v, ai);
    }
    pass16_fwd_tail(xr, xi, w);
}
/* bit position of digit p in ch, and the factor the key rows of digit p carry */
IE_HD int w12_field_shift(int p, int Bgbit) { return 31 - (p + 1) * Bgbit; }
/* inverse pass (x16) */
IE_HD void pass16_inv(double (&xr)[16], double (&xi)[16], const Tw16 &w)
{
    ibf(xr[0], xi[0], xr[1], xi[1], w.z1r, w.z1i);
    ibf(xr[2], xi[2], xr[3], xi[3], -w.z1i, w.z1r);
    ibf(xr[4], xi[4], xr[5], xi[5], w.z1qr, w.z1qi);
    ibf(xr[6], xi[6], xr[7], xi[7], -w.z1qi, w.z1qr);
    ibf(xr[8], xi[8], xr[9], xi[9], w.z1hr, w.z1hi);
    ibf(xr[10], xi[10], xr[11], xi[11], -w.z1hi, w.z1hr);
    ibf(xr[12], xi[12], xr[13], xi[13], w.z1hqr, w.z1hqi);
    ibf(xr[14], xi[14], xr[15], xi[15], -w.z1hqi, w.z1hqr);
#pragma unroll
    for (int m = 0; m < 2; m++) {
        ibf(xr[m], xi[m], xr[m + 2], xi[m + 2], w.z2r, w.z2i);
        ibf(xr[4 + m], xi[4 + m], xr[6 + m], xi[6 + m], -w.z2i, w.z2r);
        ibf(xr[8 + m], xi[8 + m], xr[10 + m], xi[10 + m], w.z2qr, w.z2qi);
        ibf(xr[12 + m], xi[12 + m], xr[14 + m], xi[14 + m], -w.z2qi, w.z2qr);
    }
#pragma unroll
    for (int m = 0; m < 4; m++) ibf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4r, w.z4i);
#pragma unroll
    for (int m = 8; m < 12; m++) ibf(xr[m], xi[m], xr[m + 4], xi[m + 4], -w.z4i, w.z4r);
#pragma unroll
    for (int m = 0; m < 8; m++) ibf(xr[m], xi[m], xr[m + 8], xi[m + 8], w.z8r, w.z8i);
}

/* The same map as pass16_inv(x, tw16_pass1()), rearranged so that every butterfly is "a +- rho b" (6 FMAs, or 4 adds
 * where rho is 1 or -i) instead of "(a - b) conj(z)" (8 operations): the conj(z) factors of pass16_inv are carried as a
 * pending unit factor per position, a butterfly only needs the ratio rho of the two pending factors — which for the
 * pass-1 twiddles (exp(i pi (1 + 4 brev)/32) ...) is a 16th root of unity — and the product of all pending factors,
 * exp(-i pi pos/32), is applied once at the end.  48 FP64 instructions fewer per pass. */
IE_HD void bf_add(double &ar, double &ai, double &br, double &bi)       /* (a + b, a - b) */
{
    const double tr = ar + br, ti = ai + bi;
    br = ar - br; bi = ai - bi; ar = tr; ai = ti;
}
IE_HD void bf_mi(double &ar, double &ai, double &br, double &bi)        /* (a - i b, a + i b) */
{
    const double tr = ar + bi, ti = ai - br;
    const double ur = ar - bi, ui = ai + br;
    ar = tr; ai = ti; br = ur; bi = ui;
}
struct TwInvP1 { double c4, c8, s8, pad, tw[16][2]; };                   /* tw[pos] = (cos, sin)(pi pos/32) */
IE_HD void pass16_inv_p1(double (&xr)[16], double (&xi)[16], const TwInvP1 &t)
{
    const double c4 = t.c4;                 /* exp(-i pi/4) = c4 (1 - i) */
    const double c8 = t.c8, s8 = t.s8;      /* exp(-i pi/8) = c8 - i s8  */
#pragma unroll
    for (int k = 0; k < 8; k++) bf_add(xr[2 * k], xi[2 * k], xr[2 * k + 1], xi[2 * k + 1]);
#pragma unroll
    for (int b = 0; b < 4; b++) {
        bf_add(xr[4 * b], xi[4 * b], xr[4 * b + 2], xi[4 * b + 2]);
        bf_mi(xr[4 * b + 1], xi[4 * b + 1], xr[4 * b + 3], xi[4 * b + 3]);
    }
#pragma unroll
    for (int c = 0; c < 2; c++) {
        bf_add(xr[8 * c], xi[8 * c], xr[8 * c + 4], xi[8 * c + 4]);
        bf(xr[8 * c + 1], xi[8 * c + 1], xr[8 * c + 5], xi[8 * c + 5], c4, -c4);
        bf_mi(xr[8 * c + 2], xi[8 * c + 2], xr[8 * c + 6], xi[8 * c + 6]);
        bf(xr[8 * c + 3], xi[8 * c + 3], xr[8 * c + 7], xi[8 * c + 7], -c4, -c4);
    }
    bf_add(xr[0], xi[0], xr[8], xi[8]);
    bf(xr[1], xi[1], xr[9], xi[9], c8, -s8);      /* exp(-i  pi/8) */
    bf(xr[2], xi[2], xr[10], xi[10], c4, -c4);    /* exp(-i 2pi/8) */
    bf(xr[3], xi[3], xr[11], xi[11], s8, -c8);    /* exp(-i 3pi/8) */
    bf_mi(xr[4], xi[4], xr[12], xi[12]);          /* exp(-i 4pi/8) */
    bf(xr[5], xi[5], xr[13], xi[13], -s8, -c8);   /* exp(-i 5pi/8) */
    bf(xr[6], xi[6], xr[14], xi[14], -c4, -c4);   /* exp(-i 6pi/8) */
    bf(xr[7], xi[7], xr[15], xi[15], -c8, -s8);   /* exp(-i 7pi/8) */
#pragma unroll
    for (int m = 1; m < 16; m++) {                /* pending factors: exp(-i pi pos/32) */
        const double r = xr[m], i = xi[m];
        xr[m] = fma(i, t.tw[m][1], r * t.tw[m][0]);
        xi[m] = fma(-r, t.tw[m][1], i * t.tw[m][0]);
    }
}
#define IE_TW_INV_P1_INIT { 0.70710678118654752440, 0.92387953251128675613, 0.38268343236508977173, 0.0, { \
    {1.00000000000000000000, 0.00000000000000000000}, \
    {0.99518472667219692873, 0.09801714032956060363}, \
    {0.98078528040323043058, 0.19509032201612824808}, \
    {0.95694033573220882438, 0.29028467725446233105}, \
    {0.92387953251128673848, 0.38268343236508978178}, \
    {0.88192126434835504956, 0.47139673682599764204}, \
    {0.83146961230254523567, 0.55557023301960217765}, \
    {0.77301045336273699338, 0.63439328416364548779}, \
    {0.70710678118654757274, 0.70710678118654746172}, \
    {0.63439328416364548779, 0.77301045336273699338}, \
    {0.55557023301960228867, 0.83146961230254523567}, \
    {0.47139673682599780857, 0.88192126434835493853}, \
    {0.38268343236508983729, 0.92387953251128673848}, \
    {0.29028467725446233105, 0.95694033573220893540}, \
    {0.19509032201612833135, 0.98078528040323043058}, \
    {0.09801714032956077016, 0.99518472667219681771} } }

/* pass-1 twiddles, base exp(i pi/32), identical for every lane */
IE_HD Tw16 tw16_pass1()
{
    Tw16 w;
    w.z8r = 0.70710678118654752440;  w.z8i = 0.70710678118654752440;    /* exp(i pi/4)     */
    w.z4r = 0.92387953251128675613;  w.z4i = 0.38268343236508977173;    /* exp(i pi/8)     */
    w.z2r = 0.98078528040323044913;  w.z2i = 0.19509032201612826785;    /* exp(i pi/16)    */
    w.z2qr = 0.55557023301960222474; w.z2qi = 0.83146961230254523708;   /* exp(i 5pi/16)   */
    w.z1r = 0.99518472667219688624;  w.z1i = 0.09801714032956060199;    /* exp(i pi/32)    */
    w.z1qr = 0.63439328416364549822; w.z1qi = 0.77301045336273696081;   /* exp(i 9pi/32)   */
    w.z1hr = 0.88192126434835502971; w.z1hi = 0.47139673682599764856;   /* exp(i 5pi/32)   */
    w.z1hqr = 0.29028467725446236764; w.z1hqi = 0.95694033573220886494; /* exp(i 13pi/32)  */
    return w;
}

/* ---- exchange buffer moves (lane = 0..31) ---- */
IE_HD void st16_pass1(cd *buf, int lane, const double (&xr)[16], const double (&xi)[16])
{
#pragma unroll
    for (int r = 0; r < 16; r++) { cd v; v.x = xr[r]; v.y = xi[r]; buf[r * 33 + lane] = v; }
}
IE_HD void ld16_pass2(const cd *buf, int lane, double (&xr)[16], double (&xi)[16])
{
    const int base = (lane & 15) * 33 + (lane >> 4);
#pragma unroll
    for (int m = 0; m < 16; m++) { cd v = buf[base + 2 * m]; xr[m] = v.x; xi[m] = v.y; }
}
IE_HD void st16_ipass2(cd *buf, int lane, const double (&xr)[16], const double (&xi)[16])
{
    const int base = (lane & 15) * 33 + (lane >> 4);
#pragma unroll
    for (int m = 0; m < 16; m++) { cd v; v.x = xr[m]; v.y = xi[m]; buf[base + 2 * m] = v; }
}
IE_HD void ld16_ipass1(const cd *buf, int lane, double (&xr)[16], double (&xi)[16])
{
#pragma unroll
    for (int r = 0; r < 16; r++) { cd v = buf[r * 33 + lane]; xr[r] = v.x; xi[r] = v.y; }
}

/* ---- final stage between lanes l and l^16 (forward) ---- */
/* the 8 values a lane hands to its partner: R_1[s] from the Lpar = 1 lane, R_0[8+s] from the Lpar = 0 lane */
IE_HD void fin_fwd_send(const double (&yr)[16], const double (&yi)[16], int lpar, double (&sr)[8], double (&si)[8])
{
#pragma unroll
    for (int s = 0; s < 8; s++) { sr[s] = lpar ? yr[s] : yr[8 + s]; si[s] = lpar ? yi[s] : yi[8 + s]; }
}
/* butterflies r' = s + 8 Lpar with Y = exp(i phi): y[s] <- R_0 + Y R_1 (evaluation K), y[8+s] <- R_0 - Y R_1 (K + 256) */
IE_HD void fin_fwd_apply(double (&yr)[16], double (&yi)[16], int lpar, const double (&rr)[8], const double (&ri)[8],
                         const double (&zr)[8], const double (&zi)[8])
{
#pragma unroll
    for (int s = 0; s < 8; s++) {
        double ar = lpar ? rr[s] : yr[s], ai = lpar ? ri[s] : yi[s];
        double br = lpar ? yr[8 + s] : rr[s], bi = lpar ? yi[8 + s] : ri[s];
        bf(ar, ai, br, bi, zr[s], zi[s]);
        yr[s] = ar; yi[s] = ai; yr[8 + s] = br; yi[8 + s] = bi;
    }
}
/* ---- inverse of the final stage: local half, then the values to hand over, then their placement ---- */
IE_HD void fin_inv_local(double (&xr)[16], double (&xi)[16], const double (&zr)[8], const double (&zi)[8])
{
#pragma unroll
    for (int s = 0; s < 8; s++) ibf(xr[s], xi[s], xr[8 + s], xi[8 + s], zr[s], zi[s]); /* x[s] = 2 R_0[r'], x[8+s] = 2 R_1[r'] */
}
IE_HD void fin_inv_send(const double (&xr)[16], const double (&xi)[16], int lpar, double (&sr)[8], double (&si)[8])
{
#pragma unroll
    for (int s = 0; s < 8; s++) { sr[s] = lpar ? xr[s] : xr[8 + s]; si[s] = lpar ? xi[s] : xi[8 + s]; }
}
IE_HD void fin_inv_place(double (&xr)[16], double (&xi)[16], int lpar, const double (&rr)[8], const double (&ri)[8])
{
#pragma unroll
    for (int s = 0; s < 8; s++) {
        if (lpar) { xr[s] = rr[s]; xi[s] = ri[s]; } else { xr[8 + s] = rr[s]; xi[8 + s] = ri[s]; }
    }
}

/* ---- "folded" forward variant: no lane-dependent selects in the final stage ----
 * The Lpar = 1 lane runs pass 2 with the twiddles of its two 8-point blocks exchanged (and -z8), which leaves its 16
 * results with the register halves swapped: register s holds root s + 8, register 8 + s holds root s.  Both lanes of a
 * pair then send registers 8..15 and keep 0..7, and both compute keep +- w * received:
 *   Lpar = 0: keep = R_0[s],   recv = R_1[s],   w = Y           -> out_+ , out_-           (as before)
 *   Lpar = 1: keep = R_1[8+s], recv = R_0[8+s], w = conj(Y)     -> out_+ / Y , -out_- / Y
 * The unit-modulus factors of the Lpar = 1 lane are multiplied into the bootstrapping key when it is laid out
 * (folded_bk_factor), so the products that reach the accumulators are the true ones and the inverse is unchanged. */
struct Tw16g { double z8r, z8i, z4ar, z4ai, z4br, z4bi, z2ar, z2ai, z2br, z2bi, z1ar, z1ai, z1aqr, z1aqi, z1br, z1bi, z1bqr, z1bqi; };

IE_HD void pass16_fwd_g(double (&xr)[16], double (&xi)[16], const Tw16g &w)
{
#pragma unroll
    for (int m = 0; m < 8; m++) bf(xr[m], xi[m], xr[m + 8], xi[m + 8], w.z8r, w.z8i);
#pragma unroll
    for (int m = 0; m < 4; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4ar, w.z4ai);
#pragma unroll
    for (int m = 8; m < 12; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4br, w.z4bi);
#pragma unroll
    for (int m = 0; m < 2; m++) {
        bf(xr[m], xi[m], xr[m + 2], xi[m + 2], w.z2ar, w.z2ai);
        bf(xr[4 + m], xi[4 + m], xr[6 + m], xi[6 + m], -w.z2ai, w.z2ar);
        bf(xr[8 + m], xi[8 + m], xr[10 + m], xi[10 + m], w.z2br, w.z2bi);
        bf(xr[12 + m], xi[12 + m], xr[14 + m], xi[14 + m], -w.z2bi, w.z2br);
    }
    bf(xr[0], xi[0], xr[1], xi[1], w.z1ar, w.z1ai);
    bf(xr[2], xi[2], xr[3], xi[3], -w.z1ai, w.z1ar);
    bf(xr[4], xi[4], xr[5], xi[5], w.z1aqr, w.z1aqi);
    bf(xr[6], xi[6], xr[7], xi[7], -w.z1aqi, w.z1aqr);
    bf(xr[8], xi[8], xr[9], xi[9], w.z1br, w.z1bi);
    bf(xr[10], xi[10], xr[11], xi[11], -w.z1bi, w.z1br);
    bf(xr[12], xi[12], xr[13], xi[13], w.z1bqr, w.z1bqi);
    bf(xr[14], xi[14], xr[15], xi[15], -w.z1bqi, w.z1bqr);
}
/* stages 2..4 of pass16_fwd_g */
IE_HD void pass16_fwd_g_tail(double (&xr)[16], double (&xi)[16], const Tw16g &w)
{
#pragma unroll
    for (int m = 0; m < 4; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4ar, w.z4ai);
#pragma unroll
    for (int m = 8; m < 12; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4br, w.z4bi);
#pragma unroll
    for (int m = 0; m < 2; m++) {
        bf(xr[m], xi[m], xr[m + 2], xi[m + 2], w.z2ar, w.z2ai);
        bf(xr[4 + m], xi[4 + m], xr[6 + m], xi[6 + m], -w.z2ai, w.z2ar);
        bf(xr[8 + m], xi[8 + m], xr[10 + m], xi[10 + m], w.z2br, w.z2bi);
        bf(xr[12 + m], xi[12 + m], xr[14 + m], xi[14 + m], -w.z2bi, w.z2br);
    }
    bf(xr[0], xi[0], xr[1], xi[1], w.z1ar, w.z1ai);
    bf(xr[2], xi[2], xr[3], xi[3], -w.z1ai, w.z1ar);
    bf(xr[4], xi[4], xr[5], xi[5], w.z1aqr, w.z1aqi);
    bf(xr[6], xi[6], xr[7], xi[7], -w.z1aqi, w.z1aqr);
    bf(xr[8], xi[8], xr[9], xi[9], w.z1br, w.z1bi);
    bf(xr[10], xi[10], xr[11], xi[11], -w.z1bi, w.z1br);
    bf(xr[12], xi[12], xr[13], xi[13], w.z1bqr, w.z1bqi);
    bf(xr[14], xi[14], xr[15], xi[15], -w.z1bqi, w.z1bqr);
}
/* uniform final stage: y[s] <- keep + w recv, y[8+s] <- keep - w recv with keep = y[s]; recv is the partner's y[8+s] */
IE_HD void fin_fwd_apply_folded(double (&yr)[16], double (&yi)[16], const double (&rr)[8], const double (&ri)[8],
                                const double (&wr)[8], const double (&wi)[8])
{
#pragma unroll
    for (int s = 0; s < 8; s++) {
        double br = rr[s], bi = ri[s];
        bf(yr[s], yi[s], br, bi, wr[s], wi[s]);
        yr[8 + s] = br; yi[8 + s] = bi;
    }
}

/* ---- select-free inverse of the folded forward transform (persistent 12-warp kernel, br_w12.cu) ----
 * The folded forward leaves the Lpar = 1 lanes' evaluations multiplied by unit factors: conj(Y) on register s and
 * -conj(Y) on register 8 + s (Y the butterfly's true twiddle).  Those factors pass through the pointwise product with
 * the PLAIN key, so the accumulators of an Lpar = 1 lane arrive as u' = conj(Y) u, v' = -conj(Y) v, and with the
 * twiddle w = conj(Y) the same inverse butterfly every lane runs gives
 *     x[s] = u' + v' = 2 R_1[8 + s],      x[8 + s] = (u' - v') conj(w) = 2 R_0[8 + s]
 * where the Lpar = 0 lane gets x[s] = 2 R_0[s], x[8 + s] = 2 R_1[s]: both lanes send registers 8..15 and receive into
 * them, the Lpar = 1 lane ends with its two halves exchanged, which the inverse of pass16_fwd_g undoes.  No key
 * factor is left over (Y conj(Y) = 1).
 * The twiddle table holds w for even s only: w[s + 1] = i w[s] is exact on Lpar = 0 lanes and has the wrong sign on
 * Lpar = 1 lanes (conj(i Y) = -i conj(Y)); the wrong sign swaps the two outputs of that forward butterfly and the
 * two inputs of its inverse, i.e. registers s and 8 + s hold each other's evaluation point: a permutation of the
 * key layout (w12_slot_to_K), nothing else. */
IE_HD void pass16_inv_g(double (&xr)[16], double (&xi)[16], const Tw16g &w)
{
    ibf(xr[0], xi[0], xr[1], xi[1], w.z1ar, w.z1ai);
    ibf(xr[2], xi[2], xr[3], xi[3], -w.z1ai, w.z1ar);
    ibf(xr[4], xi[4], xr[5], xi[5], w.z1aqr, w.z1aqi);
    ibf(xr[6], xi[6], xr[7], xi[7], -w.z1aqi, w.z1aqr);
    ibf(xr[8], xi[8], xr[9], xi[9], w.z1br, w.z1bi);
    ibf(xr[10], xi[10], xr[11], xi[11], -w.z1bi, w.z1br);
    ibf(xr[12], xi[12], xr[13], xi[13], w.z1bqr, w.z1bqi);
    ibf(xr[14], xi[14], xr[15], xi[15], -w.z1bqi, w.z1bqr);
#pragma unroll
    for (int m = 0; m < 2; m++) {
        ibf(xr[m], xi[m], xr[m + 2], xi[m + 2], w.z2ar, w.z2ai);
        ibf(xr[4 + m], xi[4 + m], xr[6 + m], xi[6 + m], -w.z2ai, w.z2ar);
        ibf(xr[8 + m], xi[8 + m], xr[10 + m], xi[10 + m], w.z2br, w.z2bi);
        ibf(xr[12 + m], xi[12 + m], xr[14 + m], xi[14 + m], -w.z2bi, w.z2br);
    }
#pragma unroll
    for (int m = 0; m < 4; m++) ibf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4ar, w.z4ai);
#pragma unroll
    for (int m = 8; m < 12; m++) ibf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.z4br, w.z4bi);
#pragma unroll
    for (int m = 0; m < 8; m++) ibf(xr[m], xi[m], xr[m + 8], xi[m + 8], w.z8r, w.z8i);
}
/* butterflies s = 4 h .. 4 h + 3 of the uniform final stage; za, zb = table entries of s = 4 h and 4 h + 2 */
IE_HD void fin4_tw(double zar, double zai, double zbr, double zbi, double (&z)[8])
{
    z[0] = zar; z[1] = zai; z[2] = -zai; z[3] = zar;
    z[4] = zbr; z[5] = zbi; z[6] = -zbi; z[7] = zbr;
}
/* key slot p of lane l in the layout the 12-warp kernel reads: the plain warp layout with registers s and 8 + s
 * exchanged for odd s on the Lpar = 1 lanes */
IE_HD int w12_slot_to_K(int p, int lane)
{
    const int q = ((lane >> 4) && (p & 1)) ? (p ^ 8) : p;
    return brev4(lane & 15) + 16 * brev4((q & 7) + 8 * (lane >> 4)) + 256 * (q >> 3);
}

/* evaluation index of register p (0..15) of lane l */
IE_HD int warp_slot_to_K(int p, int lane)
{
    return brev4(lane & 15) + 16 * brev4((p & 7) + 8 * (lane >> 4)) + 256 * (p >> 3);
}

/* (X^a - 1) * ACC_q for the 32 coefficients of a lane: c[m] = coefficient lane+32m, c[16+m] = coefficient lane+32m+512 */
IE_HD void rot_minus_one32(const int32_t *acc /*1024*/, int lane, int a, int32_t (&c)[32])
{
    const int t0 = lane - a;
#pragma unroll
    for (int h = 0; h < 32; h++) {
        const int off = 32 * (h & 15) + 512 * (h >> 4);
        const int t = t0 + off;
        const int32_t v = acc[t & (kN - 1)];
        c[h] = ((t & kN) ? -v : v) - acc[lane + off];
    }
}

/* host-side generation of the per-lane twiddles: pass 2 (indexed by k = lane & 15) and the final stage (lane, s) */
inline Tw16 make_tw16(long double ang)
{
    const long double q = 0.78539816339744830961566084581988L; /* pi/4 */
    Tw16 w;
    w.z8r = (double)cosl(8 * ang); w.z8i = (double)sinl(8 * ang);
    w.z4r = (double)cosl(4 * ang); w.z4i = (double)sinl(4 * ang);
    w.z2r = (double)cosl(2 * ang); w.z2i = (double)sinl(2 * ang);
    w.z2qr = (double)cosl(2 * ang + q); w.z2qi = (double)sinl(2 * ang + q);
    w.z1r = (double)cosl(ang); w.z1i = (double)sinl(ang);
    w.z1qr = (double)cosl(ang + q); w.z1qi = (double)sinl(ang + q);
    w.z1hr = (double)cosl(ang + q / 2); w.z1hi = (double)sinl(ang + q / 2);
    w.z1hqr = (double)cosl(ang + 3 * q / 2); w.z1hqi = (double)sinl(ang + 3 * q / 2);
    return w;
}
struct FinTw { double zr[8], zi[8]; };
inline void host_twiddles_warp(Tw16 *tw2 /*16*/, FinTw *fin /*32*/)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int k = 0; k < 16; k++) tw2[k] = make_tw16(pi * (1 + 4 * brev4(k)) / 512.0L);
    for (int lane = 0; lane < 32; lane++)
        for (int s = 0; s < 8; s++) {
            const long double phi = pi * (1 + 4 * brev4(lane & 15) + 64 * brev4(s + 8 * (lane >> 4))) / 1024.0L;
            fin[lane].zr[s] = (double)cosl(phi); fin[lane].zi[s] = (double)sinl(phi);
        }
}

/* host-side tables of the folded variant: pass-2 twiddles per LANE, final-stage w per lane, and the factor that the
 * key layout multiplies into slot p of lane l (1 for Lpar = 0) */
inline void host_twiddles_warp_folded(Tw16g *tw2 /*32*/, FinTw *fin /*32*/)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int lane = 0; lane < 32; lane++) {
        const Tw16 t = make_tw16(pi * (1 + 4 * brev4(lane & 15)) / 512.0L);
        Tw16g g;
        const bool sw = (lane >> 4) != 0;
        g.z8r = sw ? -t.z8r : t.z8r; g.z8i = sw ? -t.z8i : t.z8i;
        /* block twiddles: natural lanes give block a = (z4, z2, z1, z1q) and block b = (i z4, z2q, z1h, z1hq) */
        const double a4r = t.z4r, a4i = t.z4i, b4r = -t.z4i, b4i = t.z4r;
        g.z4ar = sw ? b4r : a4r; g.z4ai = sw ? b4i : a4i; g.z4br = sw ? a4r : b4r; g.z4bi = sw ? a4i : b4i;
        g.z2ar = sw ? t.z2qr : t.z2r; g.z2ai = sw ? t.z2qi : t.z2i; g.z2br = sw ? t.z2r : t.z2qr; g.z2bi = sw ? t.z2i : t.z2qi;
        g.z1ar = sw ? t.z1hr : t.z1r; g.z1ai = sw ? t.z1hi : t.z1i; g.z1aqr = sw ? t.z1hqr : t.z1qr; g.z1aqi = sw ? t.z1hqi : t.z1qi;
        g.z1br = sw ? t.z1r : t.z1hr; g.z1bi = sw ? t.z1i : t.z1hi; g.z1bqr = sw ? t.z1qr : t.z1hqr; g.z1bqi = sw ? t.z1qi : t.z1hqi;
        tw2[lane] = g;
        for (int s = 0; s < 8; s++) {
            const long double phi = pi * (1 + 4 * brev4(lane & 15) + 64 * brev4(s + 8 * (lane >> 4))) / 1024.0L;
            fin[lane].zr[s] = (double)cosl(phi); fin[lane].zi[s] = sw ? -(double)sinl(phi) : (double)sinl(phi);
        }
    }
}
/* factor of key slot p (0..15) of lane l in the folded layout: Y for p < 8, -Y for p >= 8 on Lpar = 1 lanes, Y = conj(fin w) */
IE_HD void folded_bk_factor(const FinTw &fin_lane, int p, int lane, double &fr, double &fi)
{
    if ((lane >> 4) == 0) { fr = 1.0; fi = 0.0; return; }
    const double yr = fin_lane.zr[p & 7], yi = -fin_lane.zi[p & 7];
    fr = (p < 8) ? yr : -yr; fi = (p < 8) ? yi : -yi;
}

} // namespace ieache
#endif
