/*
 * kernels.cu — sm_100a kernels of the gate-bootstrapping engine.
 *
 *   blind_rotate_kernel  fused: linear gate pre-combination -> modSwitch -> test-vector init ->
 *                        n x { (X^abar-1)*ACC, gadget decomposition, (k+1)l forward transforms,
 *                        pointwise MAC with BK_i, k+1 inverse transforms, ACC += } -> SampleExtract
 *                        (replaces libtfhe tfhe_bootstrap_woKS_FFT reached from
 *                        Cloud/cloud.c:30,32,38,40,43 — SURVEY.md §8 a14, App. A)
 *   keyswitch_kernel     lweKeySwitch as a vectorised gather-accumulate over a compacted digit list
 *   bk_fft_kernel        key load: coefficient BK -> transform-domain layout (libtfhe does this on
 *                        the CPU inside new_tfheGateBootstrappingCloudKeySet_fromFile, Cloud/cloud.c:657)
 *
 * One 64-thread group owns one gate: ACC (8 KB), two exchange buffers (2 x 9 KB) and the
 * mod-switched mask live in shared memory; the transform-domain accumulators (2 x 8 complex)
 * live in registers.  Groups only use their own named barrier, so a CTA is just a container
 * that makes neighbouring gates share BK_i lines in L1.
 */
#include "kernels.h"
#include "br_core.h"
#include "br_warp.h"

#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

namespace ieache {

__device__ Tw d_tw2[8];
__device__ Tw d_tw3[64];
__device__ Tw16 d_tw16[16];   /* warp layout: pass-2 twiddles by lane & 15 */
__device__ FinTw d_fin[32];   /* warp layout: final-stage twiddles by lane */
__device__ Tw16g d_tw16g[32]; /* folded forward variant: pass-2 twiddles by lane */
__device__ FinTw d_finf[32];  /* folded forward variant: final-stage w by lane */

cudaError_t upload_twiddles()
{
    Tw tw2[8], tw3[64];
    host_twiddles(tw2, tw3);
    cudaError_t e = cudaMemcpyToSymbol(d_tw2, tw2, sizeof(tw2));
    if (e != cudaSuccess) return e;
    Tw16 tw16[16];
    FinTw fin[32];
    host_twiddles_warp(tw16, fin);
    e = cudaMemcpyToSymbol(d_tw16, tw16, sizeof(tw16));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(d_fin, fin, sizeof(fin));
    if (e != cudaSuccess) return e;
    Tw16g tw16g[32];
    FinTw finf[32];
    host_twiddles_warp_folded(tw16g, finf);
    e = cudaMemcpyToSymbol(d_tw16g, tw16g, sizeof(tw16g));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(d_finf, finf, sizeof(finf));
    if (e != cudaSuccess) return e;
    e = upload_twiddles_w12();
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(d_tw3, tw3, sizeof(tw3));
}

cudaError_t twiddle_ptrs(const Tw **tw2, const Tw **tw3)
{
    cudaError_t e = cudaGetSymbolAddress((void **)tw2, d_tw2);
    if (e != cudaSuccess) return e;
    return cudaGetSymbolAddress((void **)tw3, d_tw3);
}

/* named barrier of one 64-thread group.  With one group per CTA the id is a compile-time constant (ptxas then
 * reserves 2 barriers instead of 16, worth ~5 % in the throughput kernel); a switch over immediate ids was
 * measured slower than the register form for the multi-group kernels. */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int kCtaBarrier = -1; /* group id meaning "the gate's threads are spread over every warp of the CTA" */
__device__ __forceinline__ void group_sync(int grp)
{
    if (grp < 0) __syncthreads();
    else asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
}

/* forward transform of the 8 points in registers; leaves evaluations in x[r] of thread t3 = tid */
__device__ __forceinline__ void fwd_transform(double (&xr)[8], double (&xi)[8], cd *buf, int tid, int grp,
                                              const Tw &w1, const Tw &w2, const Tw &w3)
{
    pass_fwd(xr, xi, w1);
    st_pass1(buf, tid, xr, xi);
    group_sync(grp);
    ld_pass2(buf, tid, xr, xi);
    pass_fwd(xr, xi, w2);
    st_pass2(buf, tid, xr, xi);
    group_sync(grp);
    ld_pass3(buf, tid, xr, xi);
    pass_fwd(xr, xi, w3);
}
/* same, starting from the integer digits (pass 1 with the cheaper first stage, br_core.h) */
__device__ __forceinline__ void fwd_transform_digits(const int32_t (&dr)[8], const int32_t (&di)[8], double (&xr)[8], double (&xi)[8],
                                                     cd *buf, int tid, int grp, const Tw &w1, const Tw &w2, const Tw &w3)
{
    pass1_fwd_from_digits(dr, di, xr, xi, w1);
    st_pass1(buf, tid, xr, xi);
    group_sync(grp);
    ld_pass2(buf, tid, xr, xi);
    pass_fwd(xr, xi, w2);
    st_pass2(buf, tid, xr, xi);
    group_sync(grp);
    ld_pass3(buf, tid, xr, xi);
    pass_fwd(xr, xi, w3);
}
/* inverse (x512): evaluations in x[r] -> z[tid+64m] in x[m] */
__device__ __forceinline__ void inv_transform(double (&xr)[8], double (&xi)[8], cd *buf, int tid, int grp,
                                              const Tw &w1, const Tw &w2, const Tw &w3)
{
    pass_inv(xr, xi, w3);
    st_ipass3(buf, tid, xr, xi);
    group_sync(grp);
    ld_ipass2(buf, tid, xr, xi);
    pass_inv(xr, xi, w2);
    st_ipass2(buf, tid, xr, xi);
    group_sync(grp);
    ld_ipass1(buf, tid, xr, xi);
    pass_inv(xr, xi, w1);
}

/* read-only 16-byte load whose position in the instruction stream is pinned (volatile): the compiler would
 * otherwise hoist BK_i requests to the top of a step, in front of the ACC reads the transform is waiting for */
__device__ __forceinline__ double2 ldg_pinned(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
/* ------------------------------------------------------------------ key load */
__global__ void __launch_bounds__(64) bk_fft_kernel(const int32_t *__restrict__ coef, double2 *__restrict__ out, int npoly)
{
    __shared__ cd buf[kBufElems];
    const int q = blockIdx.x, tid = threadIdx.x;
    if (q >= npoly) return;
    const int32_t *p = coef + (size_t)q * kN;
    double xr[8], xi[8];
#pragma unroll
    for (int m = 0; m < 8; m++) { xr[m] = (double)p[tid + 64 * m]; xi[m] = (double)p[tid + 64 * m + 512]; }
    const Tw w1 = tw_pass1(), w2 = d_tw2[tid >> 3], w3 = d_tw3[tid];
    fwd_transform(xr, xi, buf, tid, 0, w1, w2, w3);
    double2 *o = out + (size_t)q * kHalfN;
    const double sc = 1.0 / 512.0;
#pragma unroll
    for (int r = 0; r < 8; r++) o[r * 64 + tid] = make_double2(xr[r] * sc, xi[r] * sc);
}

cudaError_t launch_bk_fft(const int32_t *bk_coef, double2 *bkfft, int npoly, cudaStream_t s)
{
    bk_fft_kernel<<<npoly, 64, 0, s>>>(bk_coef, bkfft, npoly);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ blind rotation */
constexpr int kAccBytes = 2 * kN * 4;
constexpr int kBufBytes = kBufElems * 16;
constexpr int kAbarBytes = 2080;
constexpr int kGroupSmem = kAccBytes + 2 * kBufBytes + kAbarBytes; /* 28704 */

/* L = gadget length, G = gates (64-thread groups) per CTA, MINB = CTAs per SM the register
 * allocation is tuned for, ROLL = 0 unrolled step body, 1 rolled over both loops, 2 rolled over digits only, 3 over polynomials only: the
 * step body fits the 32 KB instruction cache (the fully unrolled body is ~60 KB of SASS) */
template <int L, int G, int MINB, int ROLL, int NOBK = 0, bool LOCK = false, bool SPREAD = false, bool ACCREG = false>
__device__ __forceinline__ void blind_rotate_body(DevParams p, const double2 *__restrict__ bkfft, GateAddr ga, const int32_t *__restrict__ baseA,
                    const int32_t *__restrict__ baseB, int32_t *__restrict__ ext);

template <int L, int G, int MINB, int ROLL, int NOBK = 0, bool LOCK = false, bool SPREAD = false, bool ACCREG = false>
__global__ void __launch_bounds__(64 * G, MINB)
blind_rotate_kernel(DevParams p, const double2 *__restrict__ bkfft, GateAddr ga, const int32_t *__restrict__ baseA,
                    const int32_t *__restrict__ baseB, int32_t *__restrict__ ext, int stagger, int sms)
{
    /* CTAs that start together run the same program in phase: their FP64 bursts collide and their exchange phases
     * leave the pipe idle together.  The first wave starts staggered; later CTAs inherit the offsets. */
    if (stagger > 0 && (int)blockIdx.x < MINB * sms) {
        const long long wait = (long long)(blockIdx.x / sms) * stagger, t0 = clock64();
        while (clock64() - t0 < wait) { }
    }
    blind_rotate_body<L, G, MINB, ROLL, NOBK, LOCK, SPREAD, ACCREG>(p, bkfft, ga, baseA, baseB, ext);
}
template <int L, int G, int MINB, int ROLL, int NOBK, bool LOCK, bool SPREAD, bool ACCREG>
__device__ __forceinline__ void blind_rotate_body(DevParams p, const double2 *__restrict__ bkfft, GateAddr ga, const int32_t *__restrict__ baseA,
                    const int32_t *__restrict__ baseB, int32_t *__restrict__ ext)
{
    static_assert(!SPREAD || (LOCK && (G == 2 || G == 4)), "SPREAD: 2 or 4 lock-step gates per CTA");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    /* SPREAD: every warp holds 32/G lanes of each of the CTA's G gates (gate thread tid = warp * 32/G + lane % (32/G)),
     * so the lanes of different gates that own the same transform slots read the same BK_i addresses in the same
     * instruction: one 128-byte wavefront serves G gates (784 -> 784/G LSU wavefronts per gate-step for BK, and
     * 1/G of the L2->L1 traffic).  The price: the G gates run in lock step on CTA-wide barriers. */
    constexpr int LPG = 32 / (SPREAD ? G : 1);
    const int grp = SPREAD ? (int)((threadIdx.x & 31) / LPG) : ((G == 1) ? 0 : (threadIdx.x >> 6));
    const int tid = SPREAD ? (int)((threadIdx.x >> 5) * LPG + (threadIdx.x & (LPG - 1))) : ((G == 1) ? threadIdx.x : (threadIdx.x & 63));
    const int bar = SPREAD ? kCtaBarrier : grp; /* barrier domain of one transform */
    int g = blockIdx.x * G + grp;
    const bool active = g < ga.ntempl * ga.n_inst;
    if (!LOCK && !active) return; /* whole group leaves; groups never share a barrier */
    if (!active) g = ga.ntempl * ga.n_inst - 1; /* lock-step CTAs: idle groups shadow the last gate, write nothing */

    unsigned char *base = smem_raw + (size_t)grp * kGroupSmem;
    int32_t *acc = reinterpret_cast<int32_t *>(base);
    cd *bufA = reinterpret_cast<cd *>(base + kAccBytes);
    cd *bufB = reinterpret_cast<cd *>(base + kAccBytes + kBufBytes);
    uint16_t *abar = reinterpret_cast<uint16_t *>(base + kAccBytes + 2 * kBufBytes);

    const int n = p.n;
    /* 1. linear pre-combination + modSwitch to Z_{2N} */
    {
        const int e = g / ga.ntempl, t = g - e * ga.ntempl;
        GateT gt = ga.uni;
        if (ga.tmpl) gt = ga.tmpl[t]; else { gt.in0 = (gt.in0 >= 0) ? t : -1; gt.in1 = (gt.in1 >= 0) ? t : -1; }
        const size_t blk = (size_t)e * ga.inst_samples;
        const int32_t *in0 = gt.in0 >= 0 ? baseA + (blk + gt.in0) * ga.stride : nullptr;
        const int32_t *in1 = gt.in1 >= 0 ? baseB + (blk + gt.in1) * ga.stride : nullptr;
        const int32_t c0 = gt.c0, c1 = gt.c1, cst = gt.cst_mu * p.mu;
        for (int i = tid; i <= n; i += 64) {
            int32_t v = (i == n) ? cst : 0;
            if (in0) v += c0 * __ldg(in0 + i);
            if (in1) v += c1 * __ldg(in1 + i);
            abar[i] = (uint16_t)modswitch_2N(v);
        }
    }
    group_sync(bar);
    /* 2. ACC = (0, X^{2N-bbar} * mu * (1 + X + ... + X^{N-1})) */
    {
        const int bbar = abar[n];
        const int a = (2 * kN - bbar) & (2 * kN - 1);
        const int ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = tid; j < kN; j += 64) {
            acc[j] = 0;
            acc[kN + j] = ((j < ar) != flip) ? -p.mu : p.mu;
        }
    }
    group_sync(bar);

    const Tw w1 = tw_pass1();
    const Tw w2 = d_tw2[tid >> 3];
    const Tw w3 = d_tw3[tid];
    const int Bgbit = p.Bgbit;
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
#pragma unroll
    for (int i = 1; i <= L; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);

    constexpr int kRowElems = 2 * kHalfN;        /* one BK row: 2 output polys x 512 points */
    constexpr int kBkStride = 2 * L * kRowElems; /* elements per BK_i */
    int toggle = 0;
    /* ACCREG: the 32 ACC coefficients this thread reads unrotated and later updates stay in registers, so a step
     * only loads the rotated operand and stores the updated value (-128 of 3 135 LSU wavefronts per gate-step) */
    int32_t areg[2][16];
    if (ACCREG) {
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int h = 0; h < 16; h++) areg[q][h] = acc[q * kN + tid + 64 * (h & 7) + 512 * (h >> 3)];
    }

    /* 3. n CMux steps */
    for (int i = 0; i < n; i++) {
        /* LOCK: all gates of the CTA enter step i together, so BK_i is fetched from L2 once per CTA and the
         * other groups hit it in L1 */
        if (LOCK && !SPREAD) __syncthreads();
        const int a = abar[i];
        /* a = 0 adds exactly zero (all digits of (X^0 - 1) ACC are 0); the skip is only a shortcut, and the gates
         * of a SPREAD CTA share barriers, so they all run the step */
        if (!SPREAD && a == 0) continue; /* uniform inside the group */
        double s0r[8], s0i[8], s1r[8], s1i[8];
#pragma unroll
        for (int r = 0; r < 8; r++) { s0r[r] = 0.0; s0i[r] = 0.0; s1r[r] = 0.0; s1i[r] = 0.0; }
        const double2 *bk_r = bkfft + (size_t)i * kBkStride + tid;

#pragma unroll((ROLL == 1 || ROLL == 3) ? 1 : 2)
        for (int q = 0; q < 2; q++) {
            int32_t c[16];
            if (ACCREG) {
                const int t0 = tid - a;
#pragma unroll
                for (int h = 0; h < 16; h++) {
                    const int t = t0 + 64 * (h & 7) + 512 * (h >> 3);
                    const int32_t v = acc[q * kN + (t & (kN - 1))];
                    c[h] = ((t & kN) ? -v : v) - areg[q][h];
                }
            } else {
                rot_minus_one(acc + q * kN, tid, a, c);
            }
#pragma unroll((ROLL == 1 || ROLL == 2) ? 1 : L)
            for (int pp = 0; pp < L; pp++) {
                const int shift = 32 - (pp + 1) * Bgbit;
                double xr[8], xi[8];
                cd *buf = toggle ? bufB : bufA;
                toggle ^= 1;
                if (ACCREG) { /* the default variant also starts the transform from the integer digits */
                    int32_t dr[8], di[8];
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        dr[m] = digit_i32(c[m], offset, shift, maskBg, halfBg);
                        di[m] = digit_i32(c[8 + m], offset, shift, maskBg, halfBg);
                    }
                    fwd_transform_digits(dr, di, xr, xi, buf, tid, bar, w1, w2, w3);
                } else {
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    xr[m] = digit_f64(c[m], offset, shift, maskBg, halfBg);
                    xi[m] = digit_f64(c[8 + m], offset, shift, maskBg, halfBg);
                }
                if (NOBK == 3) { /* timing experiment (wrong results): forward transform without the butterflies of pass 2 */
                    pass_fwd(xr, xi, w1);
                    st_pass1(buf, tid, xr, xi);
                    group_sync(bar);
                    ld_pass2(buf, tid, xr, xi);
                    st_pass2(buf, tid, xr, xi);
                    group_sync(bar);
                    ld_pass3(buf, tid, xr, xi);
                    pass_fwd(xr, xi, w3);
                } else {
                    fwd_transform(xr, xi, buf, tid, bar, w1, w2, w3);
                }
                }
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    /* NOBK (timing experiments, wrong results): 1 = no loads at all, 2 = the same 16-byte loads from a
                     * 16 KB shared-memory row (what a TMA-staged BK_i would cost the LSU, without its L2 latency) */
                    const double2 *fake = reinterpret_cast<const double2 *>(smem_raw + (size_t)G * kGroupSmem) + tid;
                    const double2 b0 = NOBK == 1 ? make_double2(1.0 + r, 0.5) : (NOBK == 2 ? fake[r * 64] : __ldg(bk_r + r * 64));
                    const double2 b1 = NOBK == 1 ? make_double2(0.25, 2.0 - r) : (NOBK == 2 ? fake[kHalfN + r * 64] : __ldg(bk_r + kHalfN + r * 64));
                    cmac(s0r[r], s0i[r], xr[r], xi[r], b0.x, b0.y);
                    cmac(s1r[r], s1i[r], xr[r], xi[r], b1.x, b1.y);
                }
                bk_r += kRowElems;
            }
        }
        /* inverse transforms and ACC update; the second pass reuses the code of the first */
#pragma unroll((ROLL != 0 && !ACCREG) ? 1 : 2)
        for (int j = 0; j < 2; j++) {
            cd *buf = toggle ? bufB : bufA;
            toggle ^= 1;
            inv_transform(s0r, s0i, buf, tid, bar, w1, w2, w3);
            int32_t *accj = acc + j * kN;
            if (ACCREG) {
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    areg[j][m] += round_to_torus(s0r[m]);
                    areg[j][8 + m] += round_to_torus(s0i[m]);
                    accj[tid + 64 * m] = areg[j][m];
                    accj[tid + 64 * m + 512] = areg[j][8 + m];
                }
            } else {
#pragma unroll
                for (int m = 0; m < 8; m++) {
                    accj[tid + 64 * m] += round_to_torus(s0r[m]);
                    accj[tid + 64 * m + 512] += round_to_torus(s0i[m]);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; r++) { s0r[r] = s1r[r]; s0i[r] = s1i[r]; }
        }
        group_sync(bar);
    }

    /* 4. SampleExtract at index 0 */
    if (!active) return;
    int32_t *o = ext + (size_t)g * kExtStride;
    for (int j = tid; j < kN; j += 64) o[j] = (j == 0) ? acc[0] : -acc[kN - j];
    if (tid == 0) o[kN] = acc[kN];
}

/* ---- throughput variant with the transform-domain accumulators in tensor memory ----
 * The 784 LSU wavefronts a gate-step spends on BK_i are per gate because a thread has no registers left to serve a
 * second gate: 64 of its 252 registers are accumulators.  Here they live in TMEM (tcgen05.st / tcgen05.ld, 32x32b:
 * every thread owns its lane's columns; ~5 KB/clk/SM of read bandwidth measured by tools/tmem_probe.cu), one
 * 64-thread group owns GP gates, and each BK_i row is loaded into registers once and used for the GP forward
 * transforms of that row: BK wavefronts and L1/L2 traffic per gate drop by GP.
 * CTA = one 64-thread group (TMEM lanes 0..63), 4 CTAs per SM, GP * 64 columns per CTA (512 in all at GP = 2). */
template <int GP> __host__ __device__ constexpr int tmem_group_smem() { return GP * (kAccBytes + kAbarBytes) + 2 * kBufBytes; }

/* 32 columns = 16 doubles per thread (one polynomial's 8 slots x {re, im}); the b32 halves are packed inside the asm
 * so that ptxas allocates them as the register pairs of the doubles (no move instructions) */
#define IE_TMEM_LD16D(taddr, d) asm volatile("{\n\t.reg .b32 t<32>;\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15,t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31}, [%16];\n\t" \
    "tcgen05.wait::ld.sync.aligned;\n\t" \
    "mov.b64 %0, {t0,t1};\n\tmov.b64 %1, {t2,t3};\n\tmov.b64 %2, {t4,t5};\n\tmov.b64 %3, {t6,t7};\n\tmov.b64 %4, {t8,t9};\n\tmov.b64 %5, {t10,t11};\n\tmov.b64 %6, {t12,t13};\n\tmov.b64 %7, {t14,t15};\n\tmov.b64 %8, {t16,t17};\n\tmov.b64 %9, {t18,t19};\n\tmov.b64 %10, {t20,t21};\n\tmov.b64 %11, {t22,t23};\n\tmov.b64 %12, {t24,t25};\n\tmov.b64 %13, {t26,t27};\n\tmov.b64 %14, {t28,t29};\n\tmov.b64 %15, {t30,t31};\n\t}" \
    : "=d"(d[0]),"=d"(d[1]),"=d"(d[2]),"=d"(d[3]),"=d"(d[4]),"=d"(d[5]),"=d"(d[6]),"=d"(d[7]),"=d"(d[8]),"=d"(d[9]),"=d"(d[10]),"=d"(d[11]),"=d"(d[12]),"=d"(d[13]),"=d"(d[14]),"=d"(d[15]) : "r"(taddr) : "memory")
#define IE_TMEM_ST16D(taddr, d) asm volatile("{\n\t.reg .b32 t<32>;\n\t" \
    "mov.b64 {t0,t1}, %1;\n\tmov.b64 {t2,t3}, %2;\n\tmov.b64 {t4,t5}, %3;\n\tmov.b64 {t6,t7}, %4;\n\tmov.b64 {t8,t9}, %5;\n\tmov.b64 {t10,t11}, %6;\n\tmov.b64 {t12,t13}, %7;\n\tmov.b64 {t14,t15}, %8;\n\tmov.b64 {t16,t17}, %9;\n\tmov.b64 {t18,t19}, %10;\n\tmov.b64 {t20,t21}, %11;\n\tmov.b64 {t22,t23}, %12;\n\tmov.b64 {t24,t25}, %13;\n\tmov.b64 {t26,t27}, %14;\n\tmov.b64 {t28,t29}, %15;\n\tmov.b64 {t30,t31}, %16;\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15,t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31};\n\t}" \
    :: "r"(taddr), "d"(d[0]),"d"(d[1]),"d"(d[2]),"d"(d[3]),"d"(d[4]),"d"(d[5]),"d"(d[6]),"d"(d[7]),"d"(d[8]),"d"(d[9]),"d"(d[10]),"d"(d[11]),"d"(d[12]),"d"(d[13]),"d"(d[14]),"d"(d[15]) : "memory")

template <int L, int GP, int MINB = 4>
__global__ void __launch_bounds__(64, MINB)
blind_rotate_tmem_kernel(DevParams p, const double2 *__restrict__ bkfft, GateAddr ga, const int32_t *__restrict__ baseA,
                         const int32_t *__restrict__ baseB, int32_t *__restrict__ ext)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    constexpr int kCols = GP * 64;                        /* per thread: GP gates x 2 polynomials x 8 slots x 4 words */
    constexpr int grp = 0;                                /* one group per CTA: compile-time barrier id */
    const int tid = threadIdx.x, warp = threadIdx.x >> 5;
    unsigned char *base = smem_raw;
    int32_t *acc = reinterpret_cast<int32_t *>(base);                                    /* [GP][2][1024] */
    uint16_t *abar = reinterpret_cast<uint16_t *>(base + GP * kAccBytes);                /* [GP][1040] */
    cd *bufA = reinterpret_cast<cd *>(base + GP * (kAccBytes + kAbarBytes)), *bufB = bufA + kBufElems;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "n"(kCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const int n = p.n;
    const long long total = (long long)ga.ntempl * ga.n_inst;
    const long long g_first = (long long)blockIdx.x * GP;
    /* 1. linear pre-combination + modSwitch, 2. ACC init, for the GP gates of this group (idle slots shadow the last gate) */
#pragma unroll
    for (int k = 0; k < GP; k++) {
        long long g = g_first + k;
        if (g >= total) g = total - 1;
        const int e = (int)(g / ga.ntempl), t = (int)(g - (long long)e * ga.ntempl);
        GateT gt = ga.uni;
        if (ga.tmpl) gt = ga.tmpl[t]; else { gt.in0 = (gt.in0 >= 0) ? t : -1; gt.in1 = (gt.in1 >= 0) ? t : -1; }
        const size_t blk = (size_t)e * ga.inst_samples;
        const int32_t *in0 = gt.in0 >= 0 ? baseA + (blk + gt.in0) * ga.stride : nullptr;
        const int32_t *in1 = gt.in1 >= 0 ? baseB + (blk + gt.in1) * ga.stride : nullptr;
        const int32_t c0 = gt.c0, c1 = gt.c1, cst = gt.cst_mu * p.mu;
        for (int i = tid; i <= n; i += 64) {
            int32_t v = (i == n) ? cst : 0;
            if (in0) v += c0 * __ldg(in0 + i);
            if (in1) v += c1 * __ldg(in1 + i);
            abar[k * 1040 + i] = (uint16_t)modswitch_2N(v);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tmem_base_slot + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int k = 0; k < GP; k++) {
        const int bbar = abar[k * 1040 + n];
        const int a = (2 * kN - bbar) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = tid; j < kN; j += 64) {
            acc[k * 2 * kN + j] = 0;
            acc[k * 2 * kN + kN + j] = ((j < ar) != flip) ? -p.mu : p.mu;
        }
    }
    group_sync(grp);

    const Tw w1 = tw_pass1(), w2 = d_tw2[tid >> 3], w3 = d_tw3[tid];
    const int Bgbit = p.Bgbit;
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
#pragma unroll
    for (int i = 1; i <= L; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    constexpr int kRowElems = 2 * kHalfN, kBkStride = 2 * L * kRowElems;
    int toggle = 0;

    for (int i = 0; i < n; i++) {
        const double2 *bk_r = bkfft + (size_t)i * kBkStride + tid;
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int32_t c[GP][16];
#pragma unroll
            for (int k = 0; k < GP; k++) rot_minus_one(acc + (k * 2 + q) * kN, tid, abar[k * 1040 + i], c[k]);
#pragma unroll 1
            for (int pp = 0; pp < L; pp++) {
                const int shift = 32 - (pp + 1) * Bgbit;
                /* this row of BK_i: once for the GP gates of the group.  With more than 4 CTAs per SM there are no
                 * registers to hold it across a transform: it is then requested per polynomial right before use */
                double2 b0[8], b1[8];
                if (MINB <= 4) {
#pragma unroll
                    for (int r = 0; r < 8; r++) { b0[r] = __ldg(bk_r + r * 64); b1[r] = __ldg(bk_r + kHalfN + r * 64); }
                }
                const double2 *bk_row = bk_r;
                bk_r += kRowElems;
                const bool first = (q == 0) && (pp == 0);
#pragma unroll
                for (int k = 0; k < GP; k++) {
                    double xr[8], xi[8];
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        xr[m] = digit_f64(c[k][m], offset, shift, maskBg, halfBg);
                        xi[m] = digit_f64(c[k][8 + m], offset, shift, maskBg, halfBg);
                    }
                    cd *buf = toggle ? bufB : bufA;
                    toggle ^= 1;
                    fwd_transform(xr, xi, buf, tid, grp, w1, w2, w3);
                    /* accumulate into TMEM (layout per gate and polynomial: 8 slots x {re, im}), one output polynomial
                     * at a time; stores of the previous row to the same columns are long complete, the wait only
                     * orders them */
                    if (!first) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        double sacc[16];
                        const uint32_t tj = taddr + (uint32_t)((k * 2 + j) * 32);
                        double2 bj[8];
#pragma unroll
                        for (int r = 0; r < 8; r++) bj[r] = (MINB <= 4) ? (j ? b1[r] : b0[r]) : __ldg(bk_row + j * kHalfN + r * 64);
                        if (!first) {
                            IE_TMEM_LD16D(tj, sacc);
                        } else {
#pragma unroll
                            for (int r = 0; r < 16; r++) sacc[r] = 0.0;
                        }
#pragma unroll
                        for (int r = 0; r < 8; r++) cmac(sacc[2 * r], sacc[2 * r + 1], xr[r], xi[r], bj[r].x, bj[r].y);
                        IE_TMEM_ST16D(tj, sacc);
                    }
                }
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        /* inverse transforms and ACC update */
#pragma unroll 1
        for (int kj = 0; kj < 2 * GP; kj++) {
            double sv[16];
            const uint32_t t0 = taddr + (uint32_t)(kj * 32);
            IE_TMEM_LD16D(t0, sv);
            double sr[8], si[8];
#pragma unroll
            for (int r = 0; r < 8; r++) { sr[r] = sv[2 * r]; si[r] = sv[2 * r + 1]; }
            cd *buf = toggle ? bufB : bufA;
            toggle ^= 1;
            inv_transform(sr, si, buf, tid, grp, w1, w2, w3);
            int32_t *accj = acc + (size_t)kj * kN;
#pragma unroll
            for (int m = 0; m < 8; m++) {
                accj[tid + 64 * m] += round_to_torus(sr[m]);
                accj[tid + 64 * m + 512] += round_to_torus(si[m]);
            }
        }
        group_sync(grp);
    }

    /* SampleExtract */
#pragma unroll
    for (int k = 0; k < GP; k++) {
        const long long g = g_first + k;
        if (g < total) {
            int32_t *o = ext + (size_t)g * kExtStride;
            const int32_t *ak = acc + k * 2 * kN;
            for (int j = tid; j < kN; j += 64) o[j] = (j == 0) ? ak[0] : -ak[kN - j];
            if (tid == 0) o[kN] = ak[kN];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_slot), "n"(kCols));
}

template <int L, int GP, int MINB = 4>
static cudaError_t launch_br_tmem(const DevParams &p, const double2 *bkfft, const GateAddr &ga, const int32_t *baseA,
                                  const int32_t *baseB, int32_t *ext, long long count, cudaStream_t s)
{
    constexpr int smem = tmem_group_smem<GP>();
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_tmem_kernel<L, GP, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    const int grid = (int)((count + GP - 1) / GP);
    blind_rotate_tmem_kernel<L, GP, MINB><<<grid, 64, smem, s>>>(p, bkfft, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

/* ---- throughput variant with one WARP per gate (br_warp.h): 16 points per lane, one shared-memory exchange and one
 * shuffle stage per transform (160 LSU wavefronts instead of 256), no CTA barrier in the loop, accumulators in TMEM
 * (2 polynomials x 16 slots x 4 words = 128 columns per lane).  CTA = 4 warps = 4 gates = the four lane quadrants of
 * the CTA's 128 TMEM columns; 2 CTAs per SM.  Reads the bootstrapping key in the layout [row][poly][slot 16][lane 32]
 * that bk_relayout_warp_kernel derives from the [slot 8][thread 64] one at key load (same values, permuted). */
constexpr int kWarpGateSmem = kAccBytes + kWarpBufElems * 16 + kAbarBytes; /* 18 720 B per gate */

__global__ void __launch_bounds__(512) bk_relayout_warp_kernel(const double2 *__restrict__ old, double2 *__restrict__ neu, int npoly, int folded)
{
    const int q = blockIdx.x, idx = threadIdx.x;
    if (q >= npoly) return;
    const int p = idx >> 5, lane = idx & 31;
    const int K = folded == 3 ? w12_slot_to_K(p, lane) : warp_slot_to_K(p, lane); /* 3: plain values in the 12-warp kernel's slot order */
    const int t3 = 8 * (K & 7) + ((K >> 3) & 7), r8 = brev3(K >> 6); /* br_core.h: K = b + 8k' + 64 brev3(r), t3 = 8b + k' */
    double2 v = old[(size_t)q * kHalfN + r8 * 64 + t3];
    if (folded == 2) { /* the unit factor the select-free forward transform leaves on the Lpar = 1 lanes (br_warp.h) */
        double fr, fi;
        folded_bk_factor(d_finf[lane], p, lane, fr, fi);
        v = make_double2(v.x * fr - v.y * fi, v.x * fi + v.y * fr);
    }
    neu[(size_t)q * kHalfN + idx] = v;
}
cudaError_t launch_bk_relayout_warp(const double2 *bkfft, double2 *bkfft_w, int npoly, int folded, cudaStream_t s)
{
    bk_relayout_warp_kernel<<<npoly, 512, 0, s>>>(bkfft, bkfft_w, npoly, folded);
    return cudaGetLastError();
}

__device__ __forceinline__ void warp_exchange8(const double (&sr)[8], const double (&si)[8], double (&rr)[8], double (&ri)[8])
{
#pragma unroll
    for (int s = 0; s < 8; s++) { rr[s] = __shfl_xor_sync(0xffffffffu, sr[s], 16); ri[s] = __shfl_xor_sync(0xffffffffu, si[s], 16); }
}

template <int L, bool FOLD = false>
__global__ void __launch_bounds__(128, 2)
blind_rotate_warp_kernel(DevParams p, const double2 *__restrict__ bkw, GateAddr ga, const int32_t *__restrict__ baseA,
                         const int32_t *__restrict__ baseB, int32_t *__restrict__ ext)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    __shared__ double s_finr[8][32], s_fini[8][32]; /* final-stage twiddles [s][lane]: 32 registers per thread otherwise */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, lpar = lane >> 4;
    __shared__ double s_finfr[FOLD ? 8 : 1][32], s_finfi[FOLD ? 8 : 1][32]; /* folded variant: forward-stage w */
    __shared__ double s_tw2[FOLD ? 16 : 1][16];                             /* folded variant: plain pass-2 twiddles for the inverse */
    if (warp == 1) {
#pragma unroll
        for (int s8 = 0; s8 < 8; s8++) { s_finr[s8][lane] = d_fin[lane].zr[s8]; s_fini[s8][lane] = d_fin[lane].zi[s8]; }
        if (FOLD) {
#pragma unroll
            for (int s8 = 0; s8 < 8; s8++) { s_finfr[s8][lane] = d_finf[lane].zr[s8]; s_finfi[s8][lane] = d_finf[lane].zi[s8]; }
        }
    }
    if (FOLD && warp == 2 && lane < 16) {
        const double *src = reinterpret_cast<const double *>(&d_tw16[lane]);
#pragma unroll
        for (int v = 0; v < 16; v++) s_tw2[v][lane] = src[v];
    }
    unsigned char *base = smem_raw + (size_t)warp * kWarpGateSmem;
    int32_t *acc = reinterpret_cast<int32_t *>(base);
    cd *buf = reinterpret_cast<cd *>(base + kAccBytes);
    uint16_t *abar = reinterpret_cast<uint16_t *>(base + kAccBytes + kWarpBufElems * 16);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_base_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const int n = p.n;
    const long long total = (long long)ga.ntempl * ga.n_inst;
    long long g = (long long)blockIdx.x * 4 + warp;
    const bool active = g < total;
    if (!active) g = total - 1; /* idle warps shadow the last gate and write nothing: the CTA frees its TMEM together */
    {
        const int e = (int)(g / ga.ntempl), t = (int)(g - (long long)e * ga.ntempl);
        GateT gt = ga.uni;
        if (ga.tmpl) gt = ga.tmpl[t]; else { gt.in0 = (gt.in0 >= 0) ? t : -1; gt.in1 = (gt.in1 >= 0) ? t : -1; }
        const size_t blk = (size_t)e * ga.inst_samples;
        const int32_t *in0 = gt.in0 >= 0 ? baseA + (blk + gt.in0) * ga.stride : nullptr;
        const int32_t *in1 = gt.in1 >= 0 ? baseB + (blk + gt.in1) * ga.stride : nullptr;
        const int32_t c0 = gt.c0, c1 = gt.c1, cst = gt.cst_mu * p.mu;
        for (int i = lane; i <= n; i += 32) {
            int32_t v = (i == n) ? cst : 0;
            if (in0) v += c0 * __ldg(in0 + i);
            if (in1) v += c1 * __ldg(in1 + i);
            abar[i] = (uint16_t)modswitch_2N(v);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tmem_base_slot + ((uint32_t)(warp * 32) << 16);
    {
        const int bbar = abar[n];
        const int a = (2 * kN - bbar) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = lane; j < kN; j += 32) { acc[j] = 0; acc[kN + j] = ((j < ar) != flip) ? -p.mu : p.mu; }
    }
    __syncwarp();

    const Tw16 w1 = tw16_pass1();
    const Tw16 w2 = d_tw16[FOLD ? 0 : (lane & 15)]; /* FOLD: unused (the forward uses w2g, the inverse reloads from shared memory) */
    const Tw16g w2g = d_tw16g[FOLD ? lane : 0];
    const int Bgbit = p.Bgbit;
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
#pragma unroll
    for (int i = 1; i <= L; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    constexpr int kRowElems = 2 * kHalfN, kBkStride = 2 * L * kRowElems;

    for (int i = 0; i < n; i++) {
        const int a = abar[i];
        const double2 *bk_r = bkw + (size_t)i * kBkStride + lane;
        bool first = true;
#pragma unroll 1
        for (int q = 0; q < 2; q++) {
            int32_t c[32];
            rot_minus_one32(acc + q * kN, lane, a, c);
#pragma unroll 1
            for (int pp = 0; pp < L; pp++) {
                const int shift = 32 - (pp + 1) * Bgbit;
                double xr[16], xi[16];
#pragma unroll
                for (int m = 0; m < 16; m++) {
                    xr[m] = digit_f64_magic(c[m], offset, shift, maskBg, halfBg);
                    xi[m] = digit_f64_magic(c[16 + m], offset, shift, maskBg, halfBg);
                }
                /* forward transform */
                pass16_fwd(xr, xi, w1);
                __syncwarp();                      /* every lane has finished reading the buffer of the previous transform */
                st16_pass1(buf, lane, xr, xi);
                __syncwarp();
                ld16_pass2(buf, lane, xr, xi);
                if (FOLD) {
                    /* select-free form: both lanes of a pair send registers 8..15, keep 0..7 and compute keep +- w recv;
                     * the unit factors this leaves on the Lpar = 1 lanes are in the key layout (br_warp.h) */
                    pass16_fwd_g(xr, xi, w2g);
                    double sr[8], si[8], rr[8], ri[8], fzr[8], fzi[8];
#pragma unroll
                    for (int s8 = 0; s8 < 8; s8++) { sr[s8] = xr[8 + s8]; si[s8] = xi[8 + s8]; }
                    warp_exchange8(sr, si, rr, ri);
#pragma unroll
                    for (int s8 = 0; s8 < 8; s8++) { fzr[s8] = s_finfr[s8][lane]; fzi[s8] = s_finfi[s8][lane]; }
                    fin_fwd_apply_folded(xr, xi, rr, ri, fzr, fzi);
                } else {
                    pass16_fwd(xr, xi, w2);
                    double sr[8], si[8], rr[8], ri[8], fzr[8], fzi[8];
                    fin_fwd_send(xr, xi, lpar, sr, si);
                    warp_exchange8(sr, si, rr, ri);
#pragma unroll
                    for (int s8 = 0; s8 < 8; s8++) { fzr[s8] = s_finr[s8][lane]; fzi[s8] = s_fini[s8][lane]; }
                    fin_fwd_apply(xr, xi, lpar, rr, ri, fzr, fzi);
                }
                /* multiply-accumulate into TMEM: 8 slots of one output polynomial at a time; the BK_i values of chunk
                 * c+1 are requested before chunk c is computed (the registers come from keeping the final-stage
                 * twiddles in shared memory: 98.9 k -> 107.8 k gates/s).  Measured and rejected: 8 pipelined chunks of 4
                 * slots (72 k: twice the tcgen05 round trips, spills); zeroing the accumulators with stores at the
                 * start of a step instead of the `first` special case (89 k: 116 bytes of spills - the kernel sits at
                 * 254 registers). */
                if (!first) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                double2 bj[8];
#pragma unroll
                for (int r = 0; r < 8; r++) bj[r] = __ldg(bk_r + r * 32);
#pragma unroll
                for (int cidx = 0; cidx < 4; cidx++) {
                    double sacc[16];
                    const uint32_t tj = taddr + (uint32_t)(32 * cidx);
                    if (!first) {
                        IE_TMEM_LD16D(tj, sacc);
                    } else {
#pragma unroll
                        for (int r = 0; r < 16; r++) sacc[r] = 0.0;
                    }
                    const int sl = 8 * (cidx & 1); /* slots 8h..8h+7 of polynomial cidx >> 1 */
#pragma unroll
                    for (int r = 0; r < 8; r++) cmac(sacc[2 * r], sacc[2 * r + 1], xr[sl + r], xi[sl + r], bj[r].x, bj[r].y);
                    if (cidx + 1 < 4) { /* the next chunk's BK_i values go into the registers just consumed, under the TMEM round trip */
#pragma unroll
                        for (int r = 0; r < 8; r++) bj[r] = ldg_pinned(bk_r + (8 * (cidx + 1) + r) * 32);
                    }
                    IE_TMEM_ST16D(tj, sacc);
                }
                first = false;
                bk_r += kRowElems;
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        /* inverse transforms and ACC update */
#pragma unroll 1
        for (int j = 0; j < 2; j++) {
            double xr[16], xi[16];
            {
                double s0[16], s1[16];
                IE_TMEM_LD16D(taddr + (uint32_t)(j * 64), s0);
                IE_TMEM_LD16D(taddr + (uint32_t)(j * 64 + 32), s1);
#pragma unroll
                for (int r = 0; r < 8; r++) { xr[r] = s0[2 * r]; xi[r] = s0[2 * r + 1]; xr[8 + r] = s1[2 * r]; xi[8 + r] = s1[2 * r + 1]; }
            }
            {
                double fzr[8], fzi[8];
#pragma unroll
                for (int s8 = 0; s8 < 8; s8++) { fzr[s8] = s_finr[s8][lane]; fzi[s8] = s_fini[s8][lane]; }
                fin_inv_local(xr, xi, fzr, fzi);
            }
            {
                double sr[8], si[8], rr[8], ri[8];
                fin_inv_send(xr, xi, lpar, sr, si);
                warp_exchange8(sr, si, rr, ri);
                fin_inv_place(xr, xi, lpar, rr, ri);
            }
            if (FOLD) {
                Tw16 wi;
                double *dst = reinterpret_cast<double *>(&wi);
#pragma unroll
                for (int v = 0; v < 16; v++) dst[v] = s_tw2[v][lane & 15];
                pass16_inv(xr, xi, wi);
            } else {
                pass16_inv(xr, xi, w2);
            }
            __syncwarp();
            st16_ipass2(buf, lane, xr, xi);
            __syncwarp();
            ld16_ipass1(buf, lane, xr, xi);
            pass16_inv(xr, xi, w1);
            int32_t *accj = acc + j * kN;
#pragma unroll
            for (int m = 0; m < 16; m++) {
                accj[lane + 32 * m] += round_to_torus(xr[m]);
                accj[lane + 32 * m + 512] += round_to_torus(xi[m]);
            }
        }
        __syncwarp();
    }

    if (active) {
        int32_t *o = ext + (size_t)g * kExtStride;
        for (int j = lane; j < kN; j += 32) o[j] = (j == 0) ? acc[0] : -acc[kN - j];
        if (lane == 0) o[kN] = acc[kN];
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base_slot));
}

template <int L, bool FOLD>
static cudaError_t launch_br_warp(const DevParams &p, const double2 *bkw, const GateAddr &ga, const int32_t *baseA,
                                  const int32_t *baseB, int32_t *ext, long long count, cudaStream_t s)
{
    constexpr int smem = 4 * kWarpGateSmem;
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_warp_kernel<L, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    blind_rotate_warp_kernel<L, FOLD><<<(int)((count + 3) / 4), 128, smem, s>>>(p, bkw, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

/* ---- latency variant with two groups per gate: group q owns ACC polynomial q, runs its l forward transforms
 * with register accumulators for both output polynomials, hands the partial sum of the *other* polynomial to
 * the other group through 8 KB of shared memory, and inverts / updates its own polynomial.  Per step: 4
 * transform latencies instead of 8, and only 32 KB of extra shared-memory traffic (an earlier kernel with one
 * group per forward transform moved 192 KB per step and was LSU-bound at 72 %). */
constexpr int pair_smem_bytes() { return kAccBytes + 2 * 2 * kBufBytes + 2 * kHalfN * 16 + kAbarBytes; }

template <int L, int MINB = 2>
__global__ void __launch_bounds__(128, MINB)
blind_rotate_pair_kernel(DevParams p, const double2 *__restrict__ bkfft, GateAddr ga, const int32_t *__restrict__ baseA,
                         const int32_t *__restrict__ baseB, int32_t *__restrict__ ext)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x >> 6, tid = threadIdx.x & 63;
    const int g = blockIdx.x;
    int32_t *acc = reinterpret_cast<int32_t *>(smem_raw);
    cd *bufA = reinterpret_cast<cd *>(smem_raw + kAccBytes) + (size_t)grp * 2 * kBufElems, *bufB = bufA + kBufElems;
    cd *xch = reinterpret_cast<cd *>(smem_raw + kAccBytes + 4 * kBufBytes); /* [2][512] */
    uint16_t *abar = reinterpret_cast<uint16_t *>(smem_raw + kAccBytes + 4 * kBufBytes + 2 * kHalfN * 16);

    const int n = p.n;
    {
        const int e = g / ga.ntempl, t = g - e * ga.ntempl;
        GateT gt = ga.uni;
        if (ga.tmpl) gt = ga.tmpl[t]; else { gt.in0 = (gt.in0 >= 0) ? t : -1; gt.in1 = (gt.in1 >= 0) ? t : -1; }
        const size_t blk = (size_t)e * ga.inst_samples;
        const int32_t *in0 = gt.in0 >= 0 ? baseA + (blk + gt.in0) * ga.stride : nullptr;
        const int32_t *in1 = gt.in1 >= 0 ? baseB + (blk + gt.in1) * ga.stride : nullptr;
        const int32_t c0 = gt.c0, c1 = gt.c1, cst = gt.cst_mu * p.mu;
        for (int i = threadIdx.x; i <= n; i += 128) {
            int32_t v = (i == n) ? cst : 0;
            if (in0) v += c0 * __ldg(in0 + i);
            if (in1) v += c1 * __ldg(in1 + i);
            abar[i] = (uint16_t)modswitch_2N(v);
        }
    }
    __syncthreads();
    {
        const int bbar = abar[n];
        const int a = (2 * kN - bbar) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = threadIdx.x; j < kN; j += 128) { acc[j] = 0; acc[kN + j] = ((j < ar) != flip) ? -p.mu : p.mu; }
    }
    __syncthreads();

    const Tw w1 = tw_pass1(), w2 = d_tw2[tid >> 3], w3 = d_tw3[tid];
    const int Bgbit = p.Bgbit;
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
#pragma unroll
    for (int i = 1; i <= L; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    constexpr int kRowElems = 2 * kHalfN, kBkStride = 2 * L * kRowElems;
    int32_t *myacc = acc + grp * kN;
    int toggle = 0;

    for (int i = 0; i < n; i++) {
        const int a = abar[i];
        if (a == 0) continue; /* uniform in the CTA */
        double mr[8], mi[8], orr[8], oi[8]; /* partial sums of my polynomial / of the other group's */
#pragma unroll
        for (int r = 0; r < 8; r++) { mr[r] = 0.0; mi[r] = 0.0; orr[r] = 0.0; oi[r] = 0.0; }
        const double2 *bk_r = bkfft + (size_t)i * kBkStride + (size_t)grp * L * kRowElems + tid;
        const double2 *bk_mine = bk_r + grp * kHalfN, *bk_other = bk_r + (1 - grp) * kHalfN;
        int32_t c[16];
        rot_minus_one(myacc, tid, a, c);
#pragma unroll
        for (int pp = 0; pp < L; pp++) {
            const int shift = 32 - (pp + 1) * Bgbit;
            double xr[8], xi[8];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                xr[m] = digit_f64_magic(c[m], offset, shift, maskBg, halfBg);
                xi[m] = digit_f64_magic(c[8 + m], offset, shift, maskBg, halfBg);
            }
            cd *buf = toggle ? bufB : bufA;
            toggle ^= 1;
            /* the row's 16 KB of BK_i are requested before the transform: with one gate per SM nothing else hides
             * the L2 latency (the barriers inside the transform keep the compiler from sinking the loads) */
            double2 b0[8], b1[8];
            if (MINB <= 2) {
#pragma unroll
                for (int r = 0; r < 8; r++) { b0[r] = __ldg(bk_mine + pp * kRowElems + r * 64); b1[r] = __ldg(bk_other + pp * kRowElems + r * 64); }
            }
            fwd_transform(xr, xi, buf, tid, grp, w1, w2, w3);
            if (MINB > 2) { /* three CTAs per SM leave 168 registers: no room to hold a row across the transform */
#pragma unroll
                for (int r = 0; r < 8; r++) { b0[r] = __ldg(bk_mine + pp * kRowElems + r * 64); b1[r] = __ldg(bk_other + pp * kRowElems + r * 64); }
            }
#pragma unroll
            for (int r = 0; r < 8; r++) {
                cmac(mr[r], mi[r], xr[r], xi[r], b0[r].x, b0[r].y);
                cmac(orr[r], oi[r], xr[r], xi[r], b1[r].x, b1[r].y);
            }
        }
        /* hand the other polynomial's partial sum over */
        cd *out = xch + (size_t)grp * kHalfN + tid;
#pragma unroll
        for (int r = 0; r < 8; r++) { cd v; v.x = orr[r]; v.y = oi[r]; out[r * 64] = v; }
        __syncthreads();
        const cd *in = xch + (size_t)(1 - grp) * kHalfN + tid;
#pragma unroll
        for (int r = 0; r < 8; r++) { const cd v = in[r * 64]; mr[r] += v.x; mi[r] += v.y; }
        cd *buf = toggle ? bufB : bufA;
        toggle ^= 1;
        inv_transform(mr, mi, buf, tid, grp, w1, w2, w3);
#pragma unroll
        for (int m = 0; m < 8; m++) {
            myacc[tid + 64 * m] += round_to_torus(mr[m]);
            myacc[tid + 64 * m + 512] += round_to_torus(mi[m]);
        }
        __syncthreads();
    }
    int32_t *o = ext + (size_t)g * kExtStride;
    for (int j = threadIdx.x; j < kN; j += 128) o[j] = (j == 0) ? acc[0] : -acc[kN - j];
    if (threadIdx.x == 0) o[kN] = acc[kN];
}

/* ---- latency variant on a 2-CTA cluster: one gate on two SMs ----
 * CTA q of the pair owns ACC polynomial q (decomposition input q and CMux output q).  Its l groups run the l
 * forward transforms of that polynomial's digits at the same time and multiply by their BK_i rows for both output
 * polynomials.  Each group parks its product for polynomial q in local shared memory and sends its product for
 * polynomial 1-q to the peer CTA with one 8 KB bulk copy (shared::cta -> shared::cluster, completion counted on
 * the peer's mbarrier); group 0 then adds the l local and the l received products, inverts and updates ACC_q.
 * Critical path per step: one forward and one inverse transform (the two-group kernel above has l + 1), at the
 * cost of two SMs per gate: a circuit level of one expression (<= 54 gates for a*b+c) fits the 74
 * cluster slots of a B200 in one wave.
 * Measured while building this (B200, cycles per step): 512 separate 16-byte remote stores + release-arrive, or
 * st.async, or a pre-summed single bulk copy all delivered ~1200 cycles after the send; mbarrier.try_wait on the
 * receiving side added ~500 cycles over a test_wait poll. */
constexpr int kClPartialBytes = kHalfN * 16; /* one partial product: 512 complex */
__host__ __device__ constexpr int cluster_smem_bytes(int L)
{
    return kN * 4 + (L + 1) * kBufBytes + L * kClPartialBytes /*mine*/ + 2 * L * kClPartialBytes /*other, by parity*/ +
           2 * L * kClPartialBytes /*received, by parity*/ + kAbarBytes + 64;
}

__device__ __forceinline__ uint32_t map_to_peer(uint32_t local_addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int L, bool PROF = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 * L, 1)
blind_rotate_cluster_kernel(DevParams p, const double2 *__restrict__ bkfft, GateAddr ga, const int32_t *__restrict__ baseA,
                            const int32_t *__restrict__ baseB, int32_t *__restrict__ ext)
{
    constexpr int NT = 64 * L;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x >> 6, tid = threadIdx.x & 63;
    const int g = blockIdx.x >> 1;
    uint32_t q;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(q));
    int32_t *acc = reinterpret_cast<int32_t *>(smem_raw);                                  /* ACC_q */
    cd *fbuf = reinterpret_cast<cd *>(smem_raw + kN * 4) + (size_t)grp * kBufElems;        /* forward exchange, per group */
    cd *ibuf = reinterpret_cast<cd *>(smem_raw + kN * 4) + (size_t)L * kBufElems;          /* inverse exchange (group 0) */
    cd *pmine = reinterpret_cast<cd *>(smem_raw + kN * 4 + (L + 1) * kBufBytes);           /* [L][512] products for polynomial q */
    cd *pother = pmine + (size_t)L * kHalfN;                                               /* [2][L][512] products for polynomial 1-q: bulk-copy sources */
    cd *recv = pother + (size_t)2 * L * kHalfN;                                            /* [2][L][512] written by the peer */
    uint16_t *abar = reinterpret_cast<uint16_t *>(recv + (size_t)2 * L * kHalfN);
    uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw + cluster_smem_bytes(L) - 64);  /* [2], one per step parity */
    constexpr uint32_t kStepTx = (uint32_t)L * kClPartialBytes;                            /* bytes the peer sends per step */

    const int n = p.n;
    if (threadIdx.x == 0) {
        /* one local arrival per use, which also announces the bytes the peer's l bulk copies will deliver */
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + 1)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(kStepTx) : "memory"); /* step 0 */
    }
    {
        const int e = g / ga.ntempl, t = g - e * ga.ntempl;
        GateT gt = ga.uni;
        if (ga.tmpl) gt = ga.tmpl[t]; else { gt.in0 = (gt.in0 >= 0) ? t : -1; gt.in1 = (gt.in1 >= 0) ? t : -1; }
        const size_t blk = (size_t)e * ga.inst_samples;
        const int32_t *in0 = gt.in0 >= 0 ? baseA + (blk + gt.in0) * ga.stride : nullptr;
        const int32_t *in1 = gt.in1 >= 0 ? baseB + (blk + gt.in1) * ga.stride : nullptr;
        const int32_t c0 = gt.c0, c1 = gt.c1, cst = gt.cst_mu * p.mu;
        for (int i = threadIdx.x; i <= n; i += NT) {
            int32_t v = (i == n) ? cst : 0;
            if (in0) v += c0 * __ldg(in0 + i);
            if (in1) v += c1 * __ldg(in1 + i);
            abar[i] = (uint16_t)modswitch_2N(v);
        }
    }
    __syncthreads();
    {
        const int bbar = abar[n];
        const int a = (2 * kN - bbar) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = threadIdx.x; j < kN; j += NT) acc[j] = (q == 0) ? 0 : (((j < ar) != flip) ? -p.mu : p.mu);
    }
    cluster_sync_all(); /* both CTAs' mbarriers are initialised and armed before anyone copies into the peer */

    const Tw w1 = tw_pass1(), w2 = d_tw2[tid >> 3], w3 = d_tw3[tid];
    const int Bgbit = p.Bgbit;
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
#pragma unroll
    for (int i = 1; i <= L; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    const int shift = 32 - (grp + 1) * Bgbit;
    constexpr int kRowElems = 2 * kHalfN, kBkStride = 2 * L * kRowElems;
    const uint32_t peer = q ^ 1u;
    const uint32_t peer_recv = map_to_peer(smem_u32(recv), peer), peer_mbar = map_to_peer(smem_u32(mbar), peer);

    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
#define CL_TICK(k) do { if (PROF) { const long long tn = clock64(); tacc[k] += tn - tprev; tprev = tn; } } while (0)
    int a = abar[0];
    /* BK_i rows of this group (16 KB) live in registers and are requested one step ahead, right after the products
     * of the current step have left: the LSU takes ~7 cycles per 512-byte global load, ~700 cycles per SM and step,
     * which then overlap the wait for the peer instead of sitting in front of the transform */
    double2 bm[8], bo[8];
    {
        const double2 *bk_r = bkfft + ((size_t)q * L + grp) * kRowElems + tid;
#pragma unroll
        for (int r = 0; r < 8; r++) { bm[r] = ldg_pinned(bk_r + q * kHalfN + r * 64); bo[r] = ldg_pinned(bk_r + (1 - q) * kHalfN + r * 64); }
    }
    for (int i = 0; i < n; i++) {
        if (PROF) tprev = clock64();
        const int par = i & 1;
        /* arm the other barrier for step i+1: the peer cannot send that step's products before this CTA has sent
         * step i's, which happens after this point in program order of thread 0's group */
        if (threadIdx.x == 0 && i + 1 < n)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar + (par ^ 1))), "r"(kStepTx) : "memory");
        double mr[8], mi[8];
        {
            double orr[8], oi[8];
            int32_t c[16];
            rot_minus_one(acc, tid, a, c); /* a = 0 gives all-zero digits: the step adds exactly zero, and the pair stays in step */
            a = abar[i + 1];               /* abar has n + 1 entries; the value is only used by the next step */
            CL_TICK(0);
            double xr[8], xi[8];
#pragma unroll
            for (int m = 0; m < 8; m++) {
                xr[m] = digit_f64_magic(c[m], offset, shift, maskBg, halfBg);
                xi[m] = digit_f64_magic(c[8 + m], offset, shift, maskBg, halfBg);
            }
            CL_TICK(1);
            fwd_transform(xr, xi, fbuf, tid, grp, w1, w2, w3);
            CL_TICK(2);
#pragma unroll
            for (int r = 0; r < 8; r++) {
                mr[r] = xr[r] * bm[r].x - xi[r] * bm[r].y; mi[r] = fma(xr[r], bm[r].y, xi[r] * bm[r].x);
                orr[r] = xr[r] * bo[r].x - xi[r] * bo[r].y; oi[r] = fma(xr[r], bo[r].y, xi[r] * bo[r].x);
            }
            /* the product for the peer's polynomial leaves first: park, make it visible to the async proxy, one
             * thread starts the copy once the whole group has parked */
            cd *po = pother + ((size_t)par * L + grp) * kHalfN;
#pragma unroll
            for (int r = 0; r < 8; r++) { cd v; v.x = orr[r]; v.y = oi[r]; po[r * 64 + tid] = v; }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            group_sync(grp);
            if (tid == 0)
                asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(peer_recv + (uint32_t)(par * L + grp) * kClPartialBytes), "r"(smem_u32(po)), "r"(kClPartialBytes),
                               "r"(peer_mbar + (uint32_t)par * 8u) : "memory");
            if (grp != 0) {
                cd *o = pmine + (size_t)grp * kHalfN + tid;
#pragma unroll
                for (int r = 0; r < 8; r++) { cd v; v.x = mr[r]; v.y = mi[r]; o[r * 64] = v; }
            }
            if (i + 1 < n) {
                const double2 *bk_r = bkfft + (size_t)(i + 1) * kBkStride + ((size_t)q * L + grp) * kRowElems + tid;
#pragma unroll
                for (int r = 0; r < 8; r++) { bm[r] = ldg_pinned(bk_r + q * kHalfN + r * 64); bo[r] = ldg_pinned(bk_r + (1 - q) * kHalfN + r * 64); }
            }
        }
        CL_TICK(3);
        __syncthreads();
        if (grp == 0) {
#pragma unroll
            for (int gg = 1; gg < L; gg++) {
                const cd *in = pmine + (size_t)gg * kHalfN + tid;
#pragma unroll
                for (int r = 0; r < 8; r++) { const cd v = in[r * 64]; mr[r] += v.x; mi[r] += v.y; }
            }
            CL_TICK(4);
            /* the peer's products of this step: mbar[par] is used every other step.  test_wait poll: try_wait's
             * suspend was measured ~500 cycles slower here */
            const uint32_t mb = smem_u32(mbar + par), phase = (uint32_t)(i >> 1) & 1u;
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mb), "r"(phase) : "memory");
            CL_TICK(5);
#pragma unroll
            for (int gg = 0; gg < L; gg++) {
                const cd *in = recv + ((size_t)par * L + gg) * kHalfN + tid;
#pragma unroll
                for (int r = 0; r < 8; r++) { const cd v = in[r * 64]; mr[r] += v.x; mi[r] += v.y; }
            }
            inv_transform(mr, mi, ibuf, tid, 0, w1, w2, w3);
            CL_TICK(6);
#pragma unroll
            for (int m = 0; m < 8; m++) {
                acc[tid + 64 * m] += round_to_torus(mr[m]);
                acc[tid + 64 * m + 512] += round_to_torus(mi[m]);
            }
            CL_TICK(7);
        }
        __syncthreads();
    }
#undef CL_TICK
    if (PROF && blockIdx.x < 2 && tid == 0)
        printf("cluster prof cta %d grp %d cycles/step: bk+rot %lld digits %lld fwd %lld mac+send %lld localsum %lld wait %lld add+inv %lld update %lld\n",
               (int)blockIdx.x, grp, tacc[0] / n, tacc[1] / n, tacc[2] / n, tacc[3] / n, tacc[4] / n, tacc[5] / n, tacc[6] / n, tacc[7] / n);
    /* SampleExtract: the mask comes from ACC_0, the body from ACC_1[0] */
    int32_t *o = ext + (size_t)g * kExtStride;
    if (q == 0) { for (int j = threadIdx.x; j < kN; j += NT) o[j] = (j == 0) ? acc[0] : -acc[kN - j]; }
    else if (threadIdx.x == 0) o[kN] = acc[0];
    cluster_sync_all(); /* nobody leaves while the peer could still be copying into this CTA */
}

template <int L>
static cudaError_t launch_br_cluster(const DevParams &p, const double2 *bkfft, const GateAddr &ga, const int32_t *baseA,
                                     const int32_t *baseB, int32_t *ext, long long count, cudaStream_t s)
{
    constexpr int smem = cluster_smem_bytes(L);
    static const bool prof = getenv("IEACHE_CLUSTER_PROF") != nullptr; /* developer aid: per-phase cycle counts of cluster 0 */
    if (prof) {
        cudaError_t e = cudaFuncSetAttribute(blind_rotate_cluster_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        blind_rotate_cluster_kernel<L, true><<<(int)count * 2, 64 * L, smem, s>>>(p, bkfft, ga, baseA, baseB, ext);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_cluster_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    blind_rotate_cluster_kernel<L><<<(int)count * 2, 64 * L, smem, s>>>(p, bkfft, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

template <int L>
static cudaError_t launch_br_pair(const DevParams &p, const double2 *bkfft, const GateAddr &ga, const int32_t *baseA,
                                  const int32_t *baseB, int32_t *ext, long long count, cudaStream_t s)
{
    constexpr int smem = pair_smem_bytes();
    static const bool three = getenv("IEACHE_PAIR_3CTA") != nullptr; /* experiment: 3 CTAs per SM at 168 registers */
    if (three) {
        cudaError_t e = cudaFuncSetAttribute(blind_rotate_pair_kernel<L, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        blind_rotate_pair_kernel<L, 3><<<(int)count, 128, smem, s>>>(p, bkfft, ga, baseA, baseB, ext);
        return cudaGetLastError();
    }
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_pair_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    blind_rotate_pair_kernel<L><<<(int)count, 128, smem, s>>>(p, bkfft, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

/* launches of at most this many gates use the latency kernel: 3 waves of 148 one-gate CTAs take about as
 * long as one wave of the throughput kernel (4 gates per SM) */
static long long g_wide_max = [] { const char *e = getenv("IEACHE_WIDE_MAX"); return e ? atoll(e) : 296LL; }();
void set_wide_max(long long v) { g_wide_max = v; }
long long get_wide_max() { return g_wide_max; }
/* launches of at most this many gates use the 2-CTA-cluster kernel: 74 pairs of SMs run in one wave */
static long long g_cluster_max = [] { const char *e = getenv("IEACHE_CLUSTER_MAX"); return e ? atoll(e) : 74LL; }();
void set_cluster_max(long long v) { g_cluster_max = v; }
long long get_cluster_max() { return g_cluster_max; }

/* launch configuration: IEACHE_BR_VARIANT selects among the compiled variants (tuning aid) */
static int g_br_variant = -1;
static int br_variant()
{
    if (g_br_variant < 0) { const char *e = getenv("IEACHE_BR_VARIANT"); g_br_variant = e ? atoi(e) : 41; }
    return g_br_variant;
}
int set_throughput_variant(int v) { const int old = br_variant(); if (v >= 0) g_br_variant = v; return old; }
int blind_rotate_groups_per_cta() { const int v = br_variant(); return (v == 4 || v == 11 || v == 13 || v == 31 || v == 34) ? 4 : ((v == 0 || v == 1 || v == 3 || v == 5 || v == 6 || v == 12 || v == 30 || v == 32 || v == 33) ? 2 : 1); }
int blind_rotate_smem_bytes(int groups) { return groups * kGroupSmem; }

template <int L, int G, int MINB, int ROLL, int NOBK = 0, bool LOCK = false, bool SPREAD = false, bool ACCREG = false>
static cudaError_t launch_br_variant(const DevParams &p, const double2 *bkfft, const GateAddr &ga, const int32_t *baseA,
                                     const int32_t *baseB, int32_t *ext, long long count, cudaStream_t s)
{
    const int smem = G * kGroupSmem + (NOBK == 2 ? 2 * kHalfN * 16 : 0);
    const int grid = (int)((count + G - 1) / G);
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_kernel<L, G, MINB, ROLL, NOBK, LOCK, SPREAD, ACCREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    static const int stagger = [] { const char *e = getenv("IEACHE_BR_STAGGER"); return e ? atoi(e) : 0; }();
    static const int sms = [] { int d = 0, v = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, d); return v; }();
    blind_rotate_kernel<L, G, MINB, ROLL, NOBK, LOCK, SPREAD, ACCREG><<<grid, 64 * G, smem, s>>>(p, bkfft, ga, baseA, baseB, ext, stagger, sms);
    return cudaGetLastError();
}

int blind_rotate_warp_layout(long long count) /* 0 = not needed, 1 = plain warp layout, 2 = folded (variant 61) */
{
    if (count <= 0) return 0;
    return br_variant() == 60 ? 1 : (br_variant() == 61 ? 2 : (br_variant() == 70 ? 3 : 0));
}

cudaError_t launch_blind_rotate(const DevParams &p, const double2 *bkfft, const double2 *bkfft_w, const GateAddr &ga, const int32_t *baseA,
                                const int32_t *baseB, int32_t *ext, int ext_base, cudaStream_t s)
{
    const long long count = (long long)ga.ntempl * ga.n_inst;
    if (count <= 0) return cudaSuccess;
    ext += (size_t)ext_base * kExtStride;
    /* narrow launches (a circuit level of a few expressions): per-gate latency is what matters */
    if (count <= g_cluster_max && count <= g_wide_max) {
        if (p.l == 3) return launch_br_cluster<3>(p, bkfft, ga, baseA, baseB, ext, count, s);
        if (p.l == 2) return launch_br_cluster<2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    }
    /* The throughput kernel holds 4 gates per SM, so its time moves in steps of 4 x SMs gates (592: 5.8 ms, 720:
     * 9.3 ms); the two-group kernel holds 2 gates per SM at 97 % of the throughput kernel's full-wave rate and
     * finer steps (720 gates: 7.4 ms).  Launches that would leave more than 10 % of the throughput kernel's last
     * wave empty therefore also use it (measured table in DESIGN.md 4). */
    bool two_group = count <= g_wide_max;
    if (!two_group && g_wide_max > 0) {
        static const long long slots = [] { int d = 0, sms = 148; cudaGetDevice(&d); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d); return 4LL * sms; }();
        const long long waves = (count + slots - 1) / slots;
        two_group = count * 10 < waves * slots * 9;
    }
    if (two_group) {
        if (p.l == 3) return launch_br_pair<3>(p, bkfft, ga, baseA, baseB, ext, count, s);
        if (p.l == 2) return launch_br_pair<2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    }
    if (bkfft_w && br_variant() == 70) return launch_blind_rotate_w12(p, bkfft_w, ga, baseA, baseB, ext, count, s);
    if (bkfft_w && br_variant() == 60) {
        if (p.l == 3) return launch_br_warp<3, false>(p, bkfft_w, ga, baseA, baseB, ext, count, s);
        if (p.l == 2) return launch_br_warp<2, false>(p, bkfft_w, ga, baseA, baseB, ext, count, s);
    }
    if (bkfft_w && br_variant() == 61) {
        if (p.l == 3) return launch_br_warp<3, true>(p, bkfft_w, ga, baseA, baseB, ext, count, s);
        if (p.l == 2) return launch_br_warp<2, true>(p, bkfft_w, ga, baseA, baseB, ext, count, s);
    }
    if (p.l == 2) return launch_br_variant<2, 1, 4, 2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    if (p.l != 3) return cudaErrorInvalidValue;
    switch (br_variant()) {
    case 52: return launch_br_tmem<3, 2>(p, bkfft, ga, baseA, baseB, ext, count, s); /* accumulators in TMEM, 2 gates per group */
    case 51: return launch_br_tmem<3, 1>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 55: return launch_br_tmem<3, 1, 5>(p, bkfft, ga, baseA, baseB, ext, count, s); /* 5 / 6 CTAs per SM: the registers the */
    case 56: return launch_br_tmem<3, 1, 6>(p, bkfft, ga, baseA, baseB, ext, count, s); /* accumulators no longer occupy      */
    case 0: return launch_br_variant<3, 2, 2, 0>(p, bkfft, ga, baseA, baseB, ext, count, s); /* round-1 first version */
    case 2: return launch_br_variant<3, 1, 6, 1>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 3: return launch_br_variant<3, 2, 2, 1>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 4: return launch_br_variant<3, 4, 1, 1>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 5: return launch_br_variant<3, 2, 4, 1>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 6: return launch_br_variant<3, 2, 3, 0>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 8: return launch_br_variant<3, 1, 5, 0>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 9: return launch_br_variant<3, 1, 6, 0>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 10: return launch_br_variant<3, 1, 5, 1>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 21: return launch_br_variant<3, 1, 6, 2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 22: return launch_br_variant<3, 1, 6, 3>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 24: return launch_br_variant<3, 1, 4, 3>(p, bkfft, ga, baseA, baseB, ext, count, s);
    /* SPREAD: lanes of 2 / 4 gates interleaved in every warp (shared BK_i wavefronts) */
    case 30: return launch_br_variant<3, 2, 2, 2, false, true, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 31: return launch_br_variant<3, 4, 1, 2, false, true, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 32: return launch_br_variant<3, 2, 3, 2, false, true, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 33: return launch_br_variant<3, 2, 2, 0, false, true, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 34: return launch_br_variant<3, 4, 1, 0, false, true, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 11: return launch_br_variant<3, 4, 1, 0, false, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 12: return launch_br_variant<3, 2, 2, 0, false, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 13: return launch_br_variant<3, 4, 1, 0, false, false>(p, bkfft, ga, baseA, baseB, ext, count, s);
    /* timing experiments only (wrong results): no BK loads */
    case 107: return launch_br_variant<3, 1, 4, 0, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 40: return launch_br_variant<3, 1, 4, 2, 0, false, false, true>(p, bkfft, ga, baseA, baseB, ext, count, s); /* ACC in registers */
    case 41: return launch_br_variant<3, 1, 4, 0, 0, false, false, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 307: return launch_br_variant<3, 1, 4, 0, 3, false, false, true>(p, bkfft, ga, baseA, baseB, ext, count, s); /* -19 % FP64, same LSU */
    case 207: return launch_br_variant<3, 1, 4, 2, 2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 208: return launch_br_variant<3, 1, 5, 2, 2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 109: return launch_br_variant<3, 1, 6, 0, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 102: return launch_br_variant<3, 1, 6, 1, true>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 1: return launch_br_variant<3, 2, 3, 1>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 7: return launch_br_variant<3, 1, 4, 0>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case 23: return launch_br_variant<3, 1, 4, 2>(p, bkfft, ga, baseA, baseB, ext, count, s); /* as fast as the fully unrolled body (7) at half the code size */
    default: return launch_br_variant<3, 1, 4, 0, 0, false, false, true>(p, bkfft, ga, baseA, baseB, ext, count, s); /* 41: ACC coefficients in registers, +1.2 % over 23 */
    }
}

/* ------------------------------------------------------------------ key switch */
constexpr int kKsThreads = 160; /* 158 int4 lanes cover a 632-word row */
constexpr int kKsMaxRows = 1024 * 16;

__global__ void __launch_bounds__(kKsThreads)
keyswitch_kernel(DevParams p, const int32_t *__restrict__ ksk, GateAddr ga, int32_t *__restrict__ out_base,
                 const int32_t *__restrict__ ext, int pair_offset, int32_t cst_post)
{
    extern __shared__ __align__(16) unsigned char ks_smem[];
    int32_t *rows = reinterpret_cast<int32_t *>(ks_smem); /* compacted row indices */
    __shared__ int s_nrows;
    const int g = blockIdx.x, tid = threadIdx.x;
    if (g >= ga.ntempl * ga.n_inst) return;
    const int e = g / ga.ntempl, tt = g - e * ga.ntempl;
    const int out_idx = ga.tmpl ? ga.tmpl[tt].out : tt;
    int32_t *outp = out_base + ((size_t)e * ga.inst_samples + out_idx) * ga.stride;
    const int t = p.ks_t, basebit = p.ks_basebit, basem1 = (1 << basebit) - 1;
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));
    const int32_t *u0 = ext + (size_t)g * kExtStride;
    const int32_t *u1 = pair_offset > 0 ? ext + ((size_t)g + pair_offset) * kExtStride : nullptr;
    if (tid == 0) s_nrows = 0;
    __syncthreads();
    for (int i = tid; i < kN; i += kKsThreads) {
        uint32_t a = (uint32_t)u0[i];
        if (u1) a += (uint32_t)u1[i];
        a += prec_offset;
        for (int j = 0; j < t; j++) {
            const int d = (a >> (32 - (j + 1) * basebit)) & basem1;
            if (d) {
                const int slot = atomicAdd(&s_nrows, 1);
                rows[slot] = (i * t + j) * basem1 + (d - 1);
            }
        }
    }
    __syncthreads();
    const int nrows = s_nrows;
    int4 accv = make_int4(0, 0, 0, 0);
    if (tid < kLweStride / 4) {
        const int4 *kv = reinterpret_cast<const int4 *>(ksk) + tid;
        int r = 0;
        for (; r + 8 <= nrows; r += 8) {
            int4 v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = __ldg(kv + (size_t)rows[r + q] * (kLweStride / 4));
#pragma unroll
            for (int q = 0; q < 8; q++) { accv.x -= v[q].x; accv.y -= v[q].y; accv.z -= v[q].z; accv.w -= v[q].w; }
        }
        for (; r < nrows; r++) {
            const int4 v = __ldg(kv + (size_t)rows[r] * (kLweStride / 4));
            accv.x -= v.x; accv.y -= v.y; accv.z -= v.z; accv.w -= v.w;
        }
        /* b lives at word n of the row */
        int32_t bval = u0[kN] + (u1 ? u1[kN] : 0) + cst_post;
        const int w0 = tid * 4;
        if (p.n >= w0 && p.n < w0 + 4) {
            if (p.n == w0) accv.x += bval; else if (p.n == w0 + 1) accv.y += bval;
            else if (p.n == w0 + 2) accv.z += bval; else accv.w += bval;
        }
        reinterpret_cast<int4 *>(outp)[tid] = accv;
    }
}

/* latency variant for narrow launches: a cluster of 8 CTAs per gate, each gathering the rows of 128 of the
 * 1024 positions; the partial sums meet in CTA 0 through distributed shared memory.  Integer adds commute,
 * so the result is bit-identical to keyswitch_kernel's. */
constexpr int kKsCluster = 8;
__global__ void __cluster_dims__(kKsCluster, 1, 1) __launch_bounds__(kKsThreads)
keyswitch_cluster_kernel(DevParams p, const int32_t *__restrict__ ksk, GateAddr ga, int32_t *__restrict__ out_base,
                         const int32_t *__restrict__ ext, int pair_offset, int32_t cst_post)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ int32_t rows[(kN / kKsCluster) * 16];
    __shared__ __align__(16) int4 partial[kLweStride / 4];
    __shared__ int s_nrows;
    const int g = blockIdx.x / kKsCluster, part = (int)cluster.block_rank(), tid = threadIdx.x;
    const int e = g / ga.ntempl, tt = g - e * ga.ntempl;
    const int out_idx = ga.tmpl ? ga.tmpl[tt].out : tt;
    int32_t *outp = out_base + ((size_t)e * ga.inst_samples + out_idx) * ga.stride;
    const int t = p.ks_t, basebit = p.ks_basebit, basem1 = (1 << basebit) - 1;
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));
    const int32_t *u0 = ext + (size_t)g * kExtStride;
    const int32_t *u1 = pair_offset > 0 ? ext + ((size_t)g + pair_offset) * kExtStride : nullptr;
    if (tid == 0) s_nrows = 0;
    __syncthreads();
    constexpr int kPer = kN / kKsCluster;
    for (int ii = tid; ii < kPer; ii += kKsThreads) {
        const int i = part * kPer + ii;
        uint32_t a = (uint32_t)u0[i];
        if (u1) a += (uint32_t)u1[i];
        a += prec_offset;
        for (int j = 0; j < t; j++) {
            const int d = (a >> (32 - (j + 1) * basebit)) & basem1;
            if (d) rows[atomicAdd(&s_nrows, 1)] = (i * t + j) * basem1 + (d - 1);
        }
    }
    __syncthreads();
    const int nrows = s_nrows;
    int4 accv = make_int4(0, 0, 0, 0);
    if (tid < kLweStride / 4) {
        const int4 *kv = reinterpret_cast<const int4 *>(ksk) + tid;
        int r = 0;
        for (; r + 8 <= nrows; r += 8) {
            int4 v[8];
#pragma unroll
            for (int q = 0; q < 8; q++) v[q] = __ldg(kv + (size_t)rows[r + q] * (kLweStride / 4));
#pragma unroll
            for (int q = 0; q < 8; q++) { accv.x -= v[q].x; accv.y -= v[q].y; accv.z -= v[q].z; accv.w -= v[q].w; }
        }
        for (; r < nrows; r++) {
            const int4 v = __ldg(kv + (size_t)rows[r] * (kLweStride / 4));
            accv.x -= v.x; accv.y -= v.y; accv.z -= v.z; accv.w -= v.w;
        }
        partial[tid] = accv;
    }
    cluster.sync();
    if (part == 0 && tid < kLweStride / 4) {
        for (int r = 1; r < kKsCluster; r++) {
            const int4 v = cluster.map_shared_rank(partial, r)[tid];
            accv.x += v.x; accv.y += v.y; accv.z += v.z; accv.w += v.w;
        }
        const int32_t bval = u0[kN] + (u1 ? u1[kN] : 0) + cst_post;
        const int w0 = tid * 4;
        if (p.n >= w0 && p.n < w0 + 4) {
            if (p.n == w0) accv.x += bval; else if (p.n == w0 + 1) accv.y += bval;
            else if (p.n == w0 + 2) accv.z += bval; else accv.w += bval;
        }
        reinterpret_cast<int4 *>(outp)[tid] = accv;
    }
    cluster.sync(); /* keep every CTA's shared memory alive until CTA 0 has read it */
}

/* throughput variant: the (base-1)*t key rows of one input position i are contiguous in the packed key
 * (24 rows = 60 672 B at t = 8, base = 4), so ONE bulk copy per position stages them in shared memory for the
 * kKsGates gates of a CTA; every gate then picks the <= t rows its digits select with 16-byte shared loads.  The
 * gather kernel above pulls each row from L2 once per gate (15.5 MB per gate, 17 TB/s: the SMs' L2 ports are the
 * limit); here a CTA pulls the whole key once per kKsGates gates (5.2 MB per gate at 12) and the shared-memory
 * pipe (128 B/clk) is the limit instead.  Integer adds commute: results are bit-identical.
 * Roles: kKsGates groups of 64 consumer threads (3 int4 lanes of the 158-lane row each) + one producer warp that
 * keeps a two-deep ring of row blocks full (mbarrier full/empty pairs). */
constexpr int kKsGates = 12;
constexpr int kKsStageThreads = kKsGates * 64 + 32;
constexpr int kKsRowBytes = kLweStride * 4;
constexpr int kKsMaxBlockRows = 24;
constexpr int kKsStageBlockBytes = kKsMaxBlockRows * kKsRowBytes;
constexpr int kKsStageUWords = 1028;
constexpr int kKsRing = 3;           /* row blocks in flight: 3 x 60 672 B + 12 samples = 226 KB of the 227 KB a CTA may use */
constexpr int kKsStageSmem = kKsRing * kKsStageBlockBytes + kKsGates * kKsStageUWords * 4 + 64;

__device__ __forceinline__ void mbar_wait(uint32_t mb, uint32_t parity)
{
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mb), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(kKsStageThreads, 1)
keyswitch_staged_kernel(DevParams p, const int32_t *__restrict__ ksk, GateAddr ga, int32_t *__restrict__ out_base,
                        const int32_t *__restrict__ ext, int pair_offset, int32_t cst_post)
{
    extern __shared__ __align__(128) unsigned char ks_smem[];
    int4 *stage = reinterpret_cast<int4 *>(ks_smem);                                   /* [kKsRing][block rows][158] */
    int32_t *u_all = reinterpret_cast<int32_t *>(ks_smem + kKsRing * kKsStageBlockBytes);    /* [kKsGates][1028] */
    uint64_t *mbar = reinterpret_cast<uint64_t *>(ks_smem + kKsStageSmem - 64);        /* full[kKsRing], empty[kKsRing] */
    const int t = p.ks_t, basebit = p.ks_basebit, basem1 = (1 << basebit) - 1;
    const int block_rows = t * basem1;
    const uint32_t block_bytes = (uint32_t)block_rows * kKsRowBytes;
    const int grp = threadIdx.x >> 6, tid = threadIdx.x & 63;
    const bool producer = grp == kKsGates;
    const long long total = (long long)ga.ntempl * ga.n_inst;
    const long long g = (long long)blockIdx.x * kKsGates + grp;
    const bool active = !producer && g < total;

    if (threadIdx.x == 0) {
        for (int r = 0; r < kKsRing; r++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(mbar + r)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar + kKsRing + r)), "r"(kKsGates * 64));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    /* the extracted sample(s) of this group's gate; idle groups use zeros (no digit selects a row) */
    if (!producer) {
        int32_t *u = u_all + grp * kKsStageUWords;
        const int32_t *u0 = active ? ext + (size_t)g * kExtStride : nullptr;
        const int32_t *u1 = (active && pair_offset > 0) ? ext + ((size_t)g + pair_offset) * kExtStride : nullptr;
        for (int i = tid; i <= kN; i += 64) u[i] = active ? (u0[i] + (u1 ? u1[i] : 0)) : 0;
    }
    __syncthreads();

    if (producer) {
        if (threadIdx.x == kKsGates * 64) {
            for (int i = 0; i < kN; i++) {
                const int b = i % kKsRing, use = i / kKsRing;
                if (use >= 1) mbar_wait(smem_u32(mbar + kKsRing + b), (uint32_t)(use - 1) & 1u); /* every consumer released the slot */
                const uint32_t full = smem_u32(mbar + b);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(block_bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(stage) + (uint32_t)b * kKsStageBlockBytes),
                               "l"(reinterpret_cast<const unsigned char *>(ksk) + (size_t)i * block_bytes), "r"(block_bytes), "r"(full) : "memory");
            }
        }
        return;
    }

    const int32_t *u = u_all + grp * kKsStageUWords;
    const uint32_t prec_offset = 1u << (32 - (1 + basebit * t));
    int4 a0 = make_int4(0, 0, 0, 0), a1 = a0, a2 = a0;
    const bool third = tid + 128 < kLweStride / 4;
    for (int i = 0; i < kN; i++) {
        const int b = i % kKsRing;
        const uint32_t av = (uint32_t)u[i] + prec_offset;
        mbar_wait(smem_u32(mbar + b), (uint32_t)(i / kKsRing) & 1u);
        const int4 *blk = stage + (size_t)b * (kKsStageBlockBytes / 16) + tid;
        /* digits two at a time: when both select a row the two subtractions fold into one three-input add per word
         * (the INT32 pipe issues at half rate on this part and was the limit with one add per word and row) */
#pragma unroll 2
        for (int j = 0; j + 1 < t; j += 2) {
            const int d0 = (av >> (32 - (j + 1) * basebit)) & basem1, d1 = (av >> (32 - (j + 2) * basebit)) & basem1;
            const int4 *r0 = blk + (size_t)(j * basem1 + d0 - 1) * (kLweStride / 4);
            const int4 *r1 = blk + (size_t)((j + 1) * basem1 + d1 - 1) * (kLweStride / 4);
            if (d0 && d1) { /* uniform in the group */
                const int4 p0 = r0[0], q0 = r1[0], p1 = r0[64], q1 = r1[64];
                a0.x = a0.x - p0.x - q0.x; a0.y = a0.y - p0.y - q0.y; a0.z = a0.z - p0.z - q0.z; a0.w = a0.w - p0.w - q0.w;
                a1.x = a1.x - p1.x - q1.x; a1.y = a1.y - p1.y - q1.y; a1.z = a1.z - p1.z - q1.z; a1.w = a1.w - p1.w - q1.w;
                if (third) { const int4 p2 = r0[128], q2 = r1[128]; a2.x = a2.x - p2.x - q2.x; a2.y = a2.y - p2.y - q2.y; a2.z = a2.z - p2.z - q2.z; a2.w = a2.w - p2.w - q2.w; }
            } else if (d0 | d1) {
                const int4 *row = d0 ? r0 : r1;
                const int4 v0 = row[0], v1 = row[64];
                a0.x -= v0.x; a0.y -= v0.y; a0.z -= v0.z; a0.w -= v0.w;
                a1.x -= v1.x; a1.y -= v1.y; a1.z -= v1.z; a1.w -= v1.w;
                if (third) { const int4 v2 = row[128]; a2.x -= v2.x; a2.y -= v2.y; a2.z -= v2.z; a2.w -= v2.w; }
            }
        }
        if (t & 1) {
            const int j = t - 1, d = (av >> (32 - (j + 1) * basebit)) & basem1;
            if (d) {
                const int4 *row = blk + (size_t)(j * basem1 + d - 1) * (kLweStride / 4);
                const int4 v0 = row[0], v1 = row[64];
                a0.x -= v0.x; a0.y -= v0.y; a0.z -= v0.z; a0.w -= v0.w;
                a1.x -= v1.x; a1.y -= v1.y; a1.z -= v1.z; a1.w -= v1.w;
                if (third) { const int4 v2 = row[128]; a2.x -= v2.x; a2.y -= v2.y; a2.z -= v2.z; a2.w -= v2.w; }
            }
        }
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(mbar + kKsRing + b)) : "memory");
    }
    if (!active) return;
    const int e = (int)(g / ga.ntempl), tt = (int)(g - (long long)e * ga.ntempl);
    const int out_idx = ga.tmpl ? ga.tmpl[tt].out : tt;
    int32_t *outp = out_base + ((size_t)e * ga.inst_samples + out_idx) * ga.stride;
    const int32_t bval = u[kN] + cst_post;
    auto put = [&](int lane, int4 v) {
        const int w0 = lane * 4;
        if (p.n >= w0 && p.n < w0 + 4) {
            if (p.n == w0) v.x += bval; else if (p.n == w0 + 1) v.y += bval;
            else if (p.n == w0 + 2) v.z += bval; else v.w += bval;
        }
        reinterpret_cast<int4 *>(outp)[lane] = v;
    };
    put(tid, a0);
    put(tid + 64, a1);
    if (third) put(tid + 128, a2);
}

/* launches of at least this many gates use the staged kernel: a CTA needs ~0.96 ms for its 12 gates whatever the
 * launch size, the gather kernel ~0.95 us per gate, so they cross near 1000 gates */
static long long g_ks_staged_min = [] { const char *e = getenv("IEACHE_KS_STAGED_MIN"); return e ? atoll(e) : 1000LL; }();
void set_ks_staged_min(long long v) { g_ks_staged_min = v; }
long long get_ks_staged_min() { return g_ks_staged_min; }

cudaError_t launch_keyswitch(const DevParams &p, const int32_t *ksk, const GateAddr &ga, int32_t *out_base,
                             const int32_t *ext, int pair_offset, int32_t cst_post, cudaStream_t s)
{
    const long long count = (long long)ga.ntempl * ga.n_inst;
    if (count <= 0) return cudaSuccess;
    const int smem = kN * p.ks_t * 4;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(keyswitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (smem > 64 * 1024) return cudaErrorInvalidValue;
    if (count <= g_wide_max && p.ks_t <= 16) {
        keyswitch_cluster_kernel<<<(unsigned)count * kKsCluster, kKsThreads, 0, s>>>(p, ksk, ga, out_base, ext, pair_offset, cst_post);
        return cudaGetLastError();
    }
    /* wide launches: row blocks staged once per kKsGates gates */
    const int block_rows = p.ks_t * ((1 << p.ks_basebit) - 1);
    if (block_rows <= kKsMaxBlockRows && count >= g_ks_staged_min) {
        static bool attr2 = false;
        if (!attr2) {
            cudaError_t e = cudaFuncSetAttribute(keyswitch_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKsStageSmem);
            if (e != cudaSuccess) return e;
            attr2 = true;
        }
        const unsigned grid = (unsigned)((count + kKsGates - 1) / kKsGates);
        keyswitch_staged_kernel<<<grid, kKsStageThreads, kKsStageSmem, s>>>(p, ksk, ga, out_base, ext, pair_offset, cst_post);
        return cudaGetLastError();
    }
    keyswitch_kernel<<<(unsigned)count, kKsThreads, smem, s>>>(p, ksk, ga, out_base, ext, pair_offset, cst_post);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ small helpers */
__global__ void linear_kernel(int32_t *out, const int32_t *a, int count, int stride, int n, int coef, int cst)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)count * stride;
    if (idx >= total) return;
    const int w = (int)(idx % stride);
    int32_t v = 0;
    if (w <= n) {
        v = (a && coef) ? coef * a[idx] : 0;
        if (w == n) v += cst;
    }
    out[idx] = v;
}
cudaError_t launch_linear(int32_t *out, const int32_t *a, int count, int stride, int n, int coef, int cst, cudaStream_t s)
{
    if (count <= 0) return cudaSuccess;
    const size_t total = (size_t)count * stride;
    linear_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(out, a, count, stride, n, coef, cst);
    return cudaGetLastError();
}

__global__ void pack_ksk_kernel(const int32_t *__restrict__ src, int32_t *__restrict__ dst, int rows_in, int base, int n)
{
    /* one CTA per destination row */
    const int row = blockIdx.x;            /* (i*t + j)*(base-1) + (d-1) */
    const int ij = row / (base - 1), d = row % (base - 1) + 1;
    const int32_t *s = src + ((size_t)ij * base + d) * (n + 1);
    int32_t *o = dst + (size_t)row * kLweStride;
    for (int w = threadIdx.x; w < kLweStride; w += blockDim.x) o[w] = (w <= n) ? s[w] : 0;
    (void)rows_in;
}
cudaError_t launch_pack_ksk(const int32_t *src, int32_t *dst, int kNdim, int t, int base, int n, cudaStream_t s)
{
    const int rows = kNdim * t * (base - 1);
    pack_ksk_kernel<<<rows, 128, 0, s>>>(src, dst, rows, base, n);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ circuit wire blocks */
/* instance block = [const0, const1, inputs..., gate slots...] of kLweStride-word samples */
__global__ void circuit_scatter_kernel(int32_t *__restrict__ wires, const int32_t *__restrict__ inputs, int n_expr, int n_inputs,
                                       int n_slots, int n, int32_t mu)
{
    const int per = n_inputs + 2;
    const long long sample = blockIdx.x; /* (e, s) */
    const int e = (int)(sample / per), s = (int)(sample % per);
    if (e >= n_expr) return;
    int32_t *dst = wires + ((size_t)e * n_slots + s) * kLweStride;
    if (s < 2) {
        for (int w = threadIdx.x; w < kLweStride; w += blockDim.x) dst[w] = (w == n) ? (s ? mu : -mu) : 0;
    } else {
        const int32_t *src = inputs + ((size_t)e * n_inputs + (s - 2)) * kLweStride;
        for (int w = threadIdx.x; w < kLweStride; w += blockDim.x) dst[w] = src[w];
    }
}
cudaError_t launch_circuit_scatter_inputs(int32_t *wires, const int32_t *inputs, int n_expr, int n_inputs, int n_slots, int n,
                                          int32_t mu, cudaStream_t s)
{
    const long long blocks = (long long)n_expr * (n_inputs + 2);
    circuit_scatter_kernel<<<(unsigned)blocks, 160, 0, s>>>(wires, inputs, n_expr, n_inputs, n_slots, n, mu);
    return cudaGetLastError();
}
__global__ void circuit_gather_kernel(int32_t *__restrict__ outputs, const int32_t *__restrict__ wires,
                                      const int32_t *__restrict__ out_slots, int n_expr, int n_outputs, int n_slots)
{
    const long long sample = blockIdx.x;
    const int e = (int)(sample / n_outputs), o = (int)(sample % n_outputs);
    if (e >= n_expr) return;
    const int32_t *src = wires + ((size_t)e * n_slots + out_slots[o]) * kLweStride;
    int32_t *dst = outputs + ((size_t)e * n_outputs + o) * kLweStride;
    for (int w = threadIdx.x; w < kLweStride; w += blockDim.x) dst[w] = src[w];
}
cudaError_t launch_circuit_gather_outputs(int32_t *outputs, const int32_t *wires, const int32_t *out_slots, int n_expr,
                                          int n_outputs, int n_slots, int n, cudaStream_t s)
{
    (void)n;
    const long long blocks = (long long)n_expr * n_outputs;
    circuit_gather_kernel<<<(unsigned)blocks, 160, 0, s>>>(outputs, wires, out_slots, n_expr, n_outputs, n_slots);
    return cudaGetLastError();
}

/* ------------------------------------------------------------------ FP64 pipe peak (roofline denominator) */
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters)
{
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    if (s == 12345.678) out[0] = s; /* keep the chain alive */
}
cudaError_t launch_fp64_peak(cudaStream_t s, double *tflops)
{
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *d = nullptr;
    cudaError_t e = cudaMalloc((void **)&d, 8);
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4096, blocks = sms * 8, threads = 256;
    double best = 0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(a, s);
        fp64_peak_kernel<<<blocks, threads, 0, s>>>(d, iters);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        if (rep > 0) best = fmax(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(d);
    *tflops = best;
    return cudaGetLastError();
}

} // namespace ieache
