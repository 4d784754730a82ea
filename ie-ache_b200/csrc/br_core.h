/*
 * br_core.h — per-thread building blocks of the blind-rotation kernel (N = 1024, k = 1).
 *
 * Replaces libtfhe's tfhe_blindRotate_FFT / tGswFFTExternMulToTLwe / spqlios FFT that
 * Cloud/cloud.c reaches through bootsAND/bootsXOR (Cloud/cloud.c:30-43).
 *
 * A torus polynomial mod X^1024+1 is folded to 512 complex points z_j = p_j + i p_{j+512} and
 * evaluated at the 512 roots psi^(4K+1), psi = exp(i pi/1024), by three radix-8 passes of a
 * "negacyclic-native" decimation (no separate twist pass): pass p reduces
 *      sum_m x_m Y^m  mod (Y^8 - s^8)   ->   y_k = sum_m x_m (s w8^k)^m
 * with 12 twiddled radix-2 butterflies of 6 FMAs each.  64 threads own 8 points each:
 *   pass 1: thread t holds z[t+64m];      s = exp(i pi/16) for every thread (immediates)
 *   pass 2: thread (b,u) = (t>>3, t&7);   s = exp(i pi (1+4b)/128)
 *   pass 3: thread t3 = 8b+k';            s = exp(i pi (1+4b+32k')/1024)
 * Between passes the points go through one 576-element padded shared buffer
 * (phys(i) = i + i/8, conflict-free for 16-byte accesses).  After pass 3, register r of
 * thread t3 holds evaluation K = b + 8k' + 64*bitrev3(r); the bootstrapping key is stored in
 * exactly that [r][t3] order, so the pointwise product needs no permutation.
 *
 * Every function is __host__ __device__ so tests/ can run the identical arithmetic on the CPU
 * (csrc/host_emul.cpp) — there is no GPU in the build container.
 */
#ifndef IEACHE_BR_CORE_H
#define IEACHE_BR_CORE_H

#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define IE_HD __host__ __device__ __forceinline__
#else
#define IE_HD inline
#endif

namespace ieache {

constexpr int kN = 1024;          /* ring degree */
constexpr int kHalfN = 512;       /* complex points */
constexpr int kGroup = 64;        /* threads per transform */
constexpr int kBufElems = 576;    /* 512 + 512/8 padding */

struct cd { double x, y; };       /* complex double, 16-byte element of the exchange buffer */

IE_HD int phys(int i) { return i + (i >> 3); }

/* (a, b) <- (a + z b, a - z b) : 6 FMAs */
IE_HD void bf(double &ar, double &ai, double &br, double &bi, double zr, double zi)
{
    double tr = fma(zr, br, ar); tr = fma(-zi, bi, tr);
    double ti = fma(zr, bi, ai); ti = fma(zi, br, ti);
    br = fma(2.0, ar, -tr); bi = fma(2.0, ai, -ti);
    ar = tr; ai = ti;
}
/* (a, b) <- (a + b, (a - b) conj(z)) : inverse of bf up to a factor 2 */
IE_HD void ibf(double &ar, double &ai, double &br, double &bi, double zr, double zi)
{
    double dr = ar - br, di = ai - bi;
    ar += br; ai += bi;
    br = fma(di, zi, dr * zr);
    bi = fma(-dr, zi, di * zr);
}

/* twiddles of one radix-8 pass: s^4, s^2, s, s*exp(i pi/4) */
struct Tw { double s4r, s4i, s2r, s2i, s1r, s1i, sqr, sqi; };

/* forward radix-8 pass; input x[m] natural order, output x[r] = y_{bitrev3(r)} */
IE_HD void pass_fwd(double (&xr)[8], double (&xi)[8], const Tw &w)
{
#pragma unroll
    for (int m = 0; m < 4; m++) bf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.s4r, w.s4i);
    bf(xr[0], xi[0], xr[2], xi[2], w.s2r, w.s2i);
    bf(xr[1], xi[1], xr[3], xi[3], w.s2r, w.s2i);
    bf(xr[4], xi[4], xr[6], xi[6], -w.s2i, w.s2r);
    bf(xr[5], xi[5], xr[7], xi[7], -w.s2i, w.s2r);
    bf(xr[0], xi[0], xr[1], xi[1], w.s1r, w.s1i);
    bf(xr[2], xi[2], xr[3], xi[3], -w.s1i, w.s1r);
    bf(xr[4], xi[4], xr[5], xi[5], w.sqr, w.sqi);
    bf(xr[6], xi[6], xr[7], xi[7], -w.sqi, w.sqr);
}
/* inverse pass (x8): input x[r] = y_{bitrev3(r)}, output x[m] natural order */
IE_HD void pass_inv(double (&xr)[8], double (&xi)[8], const Tw &w)
{
    ibf(xr[0], xi[0], xr[1], xi[1], w.s1r, w.s1i);
    ibf(xr[2], xi[2], xr[3], xi[3], -w.s1i, w.s1r);
    ibf(xr[4], xi[4], xr[5], xi[5], w.sqr, w.sqi);
    ibf(xr[6], xi[6], xr[7], xi[7], -w.sqi, w.sqr);
    ibf(xr[0], xi[0], xr[2], xi[2], w.s2r, w.s2i);
    ibf(xr[1], xi[1], xr[3], xi[3], w.s2r, w.s2i);
    ibf(xr[4], xi[4], xr[6], xi[6], -w.s2i, w.s2r);
    ibf(xr[5], xi[5], xr[7], xi[7], -w.s2i, w.s2r);
#pragma unroll
    for (int m = 0; m < 4; m++) ibf(xr[m], xi[m], xr[m + 4], xi[m + 4], w.s4r, w.s4i);
}

/* Pass 1 of a forward transform straight from the integer gadget digits.  Its first stage multiplies by
 * s^4 = (1 + i)/sqrt 2, so a + s^4 b = (a_r + c (b_r - b_i)) + i (a_i + c (b_r + b_i)) with c = 1/sqrt 2: taking the
 * difference and the sum of the two small integers before the conversion leaves 4 FMAs per butterfly instead of 6
 * (48 fewer FP64 instructions per CMux step and thread).  dr[m], di[m]: digits of coefficients m and m+8 of the 16
 * a thread owns (the real and imaginary part of its point m). */
IE_HD void pass1_fwd_from_digits(const int32_t (&dr)[8], const int32_t (&di)[8], double (&xr)[8], double (&xi)[8], const Tw &w)
{
    const double c = 0.70710678118654752440;
#pragma unroll
    for (int m = 0; m < 4; m++) {
        const double ar = (double)dr[m], ai = (double)di[m];
        const double u = (double)(dr[m + 4] - di[m + 4]), v = (double)(dr[m + 4] + di[m + 4]);
        xr[m] = fma(c, u, ar); xi[m] = fma(c, v, ai);
        xr[m + 4] = fma(-c, u, ar); xi[m + 4] = fma(-c, v, ai);
    }
    bf(xr[0], xi[0], xr[2], xi[2], w.s2r, w.s2i);
    bf(xr[1], xi[1], xr[3], xi[3], w.s2r, w.s2i);
    bf(xr[4], xi[4], xr[6], xi[6], -w.s2i, w.s2r);
    bf(xr[5], xi[5], xr[7], xi[7], -w.s2i, w.s2r);
    bf(xr[0], xi[0], xr[1], xi[1], w.s1r, w.s1i);
    bf(xr[2], xi[2], xr[3], xi[3], -w.s1i, w.s1r);
    bf(xr[4], xi[4], xr[5], xi[5], w.sqr, w.sqi);
    bf(xr[6], xi[6], xr[7], xi[7], -w.sqi, w.sqr);
}
/* digit p of the signed base-2^Bgbit decomposition as an integer */
IE_HD int32_t digit_i32(int32_t c, uint32_t offset, int shift, uint32_t mask, int32_t halfBg)
{
    return (int32_t)((((uint32_t)c + offset) >> shift) & mask) - halfBg;
}

IE_HD int brev3(int r) { return ((r & 1) << 2) | (r & 2) | ((r >> 2) & 1); }

/* pass-1 twiddles: s = exp(i pi/16), identical for every thread */
IE_HD Tw tw_pass1()
{
    Tw w;
    w.s4r = 0.70710678118654752440;  w.s4i = 0.70710678118654752440;   /* exp(i pi/4)   */
    w.s2r = 0.92387953251128675613;  w.s2i = 0.38268343236508977173;   /* exp(i pi/8)   */
    w.s1r = 0.98078528040323044913;  w.s1i = 0.19509032201612826785;   /* exp(i pi/16)  */
    w.sqr = 0.55557023301960222474;  w.sqi = 0.83146961230254523708;   /* exp(i 5pi/16) */
    return w;
}

/* ---- exchange-buffer moves (tid = thread index inside the 64-thread group) ---- */
IE_HD void st_pass1(cd *buf, int tid, const double (&xr)[8], const double (&xi)[8])
{
#pragma unroll
    for (int r = 0; r < 8; r++) { cd v; v.x = xr[r]; v.y = xi[r]; buf[phys(64 * brev3(r) + tid)] = v; }
}
IE_HD void ld_pass2(const cd *buf, int tid, double (&xr)[8], double (&xi)[8])
{
    const int base = 72 * (tid >> 3) + (tid & 7);   /* phys(64b + u + 8m) = 72b + 9m + u */
#pragma unroll
    for (int m = 0; m < 8; m++) { cd v = buf[base + 9 * m]; xr[m] = v.x; xi[m] = v.y; }
}
IE_HD void st_pass2(cd *buf, int tid, const double (&xr)[8], const double (&xi)[8])
{
    const int base = 72 * (tid >> 3) + (tid & 7);
#pragma unroll
    for (int r = 0; r < 8; r++) { cd v; v.x = xr[r]; v.y = xi[r]; buf[base + 9 * brev3(r)] = v; }
}
IE_HD void ld_pass3(const cd *buf, int tid, double (&xr)[8], double (&xi)[8])
{
#pragma unroll
    for (int m = 0; m < 8; m++) { cd v = buf[9 * tid + m]; xr[m] = v.x; xi[m] = v.y; }
}
/* inverse direction */
IE_HD void st_ipass3(cd *buf, int tid, const double (&xr)[8], const double (&xi)[8])
{
#pragma unroll
    for (int m = 0; m < 8; m++) { cd v; v.x = xr[m]; v.y = xi[m]; buf[9 * tid + m] = v; }
}
IE_HD void ld_ipass2(const cd *buf, int tid, double (&xr)[8], double (&xi)[8])
{
    const int base = 72 * (tid >> 3) + (tid & 7);
#pragma unroll
    for (int r = 0; r < 8; r++) { cd v = buf[base + 9 * brev3(r)]; xr[r] = v.x; xi[r] = v.y; }
}
IE_HD void st_ipass2(cd *buf, int tid, const double (&xr)[8], const double (&xi)[8])
{
    const int base = 72 * (tid >> 3) + (tid & 7);
#pragma unroll
    for (int m = 0; m < 8; m++) { cd v; v.x = xr[m]; v.y = xi[m]; buf[base + 9 * m] = v; }
}
IE_HD void ld_ipass1(const cd *buf, int tid, double (&xr)[8], double (&xi)[8])
{
#pragma unroll
    for (int r = 0; r < 8; r++) { cd v = buf[phys(64 * brev3(r) + tid)]; xr[r] = v.x; xi[r] = v.y; }
}

/* ---- gadget decomposition of (X^a - 1) * ACC_q for the 16 coefficients a thread owns ---- */
/* c[m] = coefficient tid+64m, c[8+m] = coefficient tid+64m+512 of (X^a - 1) * acc, 0 < a < 2N */
IE_HD void rot_minus_one(const int32_t *acc /*1024*/, int tid, int a, int32_t (&c)[16])
{
    /* j - a lies in (-2N, N): its low 10 bits are the source index and bit 10 says whether the term picked up a sign
     * from X^N = -1 (once for a borrow, once more for a >= N) */
    const int t0 = tid - a;
#pragma unroll
    for (int h = 0; h < 16; h++) {
        const int off = 64 * (h & 7) + 512 * (h >> 3);
        const int t = t0 + off;
        const int32_t v = acc[t & (kN - 1)];
        c[h] = ((t & kN) ? -v : v) - acc[tid + off];
    }
}
/* digit p of the signed base-2^Bgbit decomposition (tGswTorus32PolynomialDecompH) */
IE_HD double digit_f64(int32_t c, uint32_t offset, int shift, uint32_t mask, int32_t halfBg)
{
    const uint32_t v = (uint32_t)c + offset;
    return (double)((int32_t)((v >> shift) & mask) - halfBg);
}

/* same value built without the int->double conversion unit: 2^52 + u has u in its low mantissa word, and the
 * subtraction is exact.  One FP64 add instead of an XU-pipe I2F (8 issue cycles per warp): used where a single
 * gate's latency matters; the throughput kernel keeps I2F because its FP64 pipe is the scarce one. */
IE_HD double digit_f64_magic(int32_t c, uint32_t offset, int shift, uint32_t mask, int32_t halfBg)
{
    const uint32_t u = (((uint32_t)c + offset) >> shift) & mask;
#ifdef __CUDA_ARCH__
    return __hiloint2double(0x43300000, (int)u) - (4503599627370496.0 + (double)halfBg);
#else
    return (double)((int32_t)u - halfBg);
#endif
}

/* round a double (|v| < 2^51) to the nearest integer and keep the low 32 bits */
IE_HD int32_t round_to_torus(double v)
{
    const double t = v + 6755399441055744.0; /* 1.5 * 2^52 */
#ifdef __CUDA_ARCH__
    return __double2loint(t);
#else
    union { double d; uint64_t u; } cv; cv.d = t;
    return (int32_t)(uint32_t)cv.u;
#endif
}

/* modSwitchFromTorus32(x, 2N) for N = 1024 */
IE_HD int modswitch_2N(int32_t x) { return (int)((((uint32_t)x) + (1u << 20)) >> 21); }

/* complex multiply-accumulate: acc += y * b */
IE_HD void cmac(double &ar, double &ai, double yr, double yi, double br, double bi)
{
    ar = fma(yr, br, ar); ar = fma(-yi, bi, ar);
    ai = fma(yr, bi, ai); ai = fma(yi, br, ai);
}

/* host-side generation of the per-thread twiddles of pass 2 (indexed by b) and pass 3 (by t3) */
inline Tw make_tw(long double ang)
{
    Tw w;
    const long double q = 0.78539816339744830961566084581988L; /* pi/4 */
    w.s4r = (double)cosl(4 * ang); w.s4i = (double)sinl(4 * ang);
    w.s2r = (double)cosl(2 * ang); w.s2i = (double)sinl(2 * ang);
    w.s1r = (double)cosl(ang);     w.s1i = (double)sinl(ang);
    w.sqr = (double)cosl(ang + q); w.sqi = (double)sinl(ang + q);
    return w;
}
inline void host_twiddles(Tw *tw2 /*8*/, Tw *tw3 /*64*/)
{
    const long double pi = 3.14159265358979323846264338327950288L;
    for (int b = 0; b < 8; b++) tw2[b] = make_tw(pi * (1 + 4 * b) / 128.0L);
    for (int t = 0; t < 64; t++) {
        const int b = t >> 3, k1 = t & 7;
        tw3[t] = make_tw(pi * (1 + 4 * b + 32 * k1) / 1024.0L);
    }
}

} // namespace ieache
#endif
