/*
 * kernels.cu — sm_100a kernels of the gate-bootstrapping engine (all but the persistent blind rotation, br_w12.cu).
 *
 *   blind_rotate_{cluster,pair,}_kernel   fused: linear gate pre-combination -> modSwitch -> test-vector init ->
 *                        n x { (X^abar-1)*ACC, gadget decomposition, (k+1)l forward transforms, pointwise MAC with
 *                        BK_i, k+1 inverse transforms, ACC += } -> SampleExtract (replaces libtfhe
 *                        tfhe_bootstrap_woKS_FFT reached from Cloud/cloud.c:30,32,38,40,43 — SURVEY.md §8 a14, App. A)
 *                        in three shapes by launch size: one gate on a 2-CTA cluster, one gate on two groups of a
 *                        CTA, one gate per 64-thread CTA
 *   keyswitch_{cluster,staged,}_kernel    lweKeySwitch: 8-CTA cluster per gate / TMA-staged row blocks shared by 12
 *                        gates / per-gate gather over a compacted digit list
 *   bk_fft_kernel        key load: coefficient BK -> transform-domain layout (libtfhe does this on the CPU inside
 *                        new_tfheGateBootstrappingCloudKeySet_fromFile, Cloud/cloud.c:657)
 *   pick_blind_rotate / pick_keyswitch    which shape a launch of a given size uses (LaunchPolicy, per context)
 */
#include "kernels.h"
#include "br_core.h"
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

namespace ieache {

__device__ Tw d_tw2[8];
__device__ Tw d_tw3[64];

cudaError_t upload_twiddles()
{
    Tw tw2[8], tw3[64];
    host_twiddles(tw2, tw3);
    cudaError_t e = cudaMemcpyToSymbol(d_tw2, tw2, sizeof(tw2));
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
__device__ __forceinline__ void group_sync(int grp)
{
    asm volatile("bar.sync %0, 64;" ::"r"(grp + 1) : "memory");
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

/* One 64-thread group (CTA) per gate, 4 CTAs per SM: the transform-domain accumulators (2 x 8 complex) and, with
 * ACCREG, the 32 ACC coefficients a thread reads unrotated and updates live in registers.  L = gadget length;
 * ROLL = 0 unrolled step body, 2 rolled over the digits (half the code).  Used for launches of 297..~900 gates, where
 * its 592-gate wave is a better fit than the 1 776-gate wave of the persistent kernel (br_w12.cu). */
template <int L, int ROLL, bool ACCREG>
__global__ void __launch_bounds__(64, 4)
blind_rotate_kernel(DevParams p, const double2 *__restrict__ bkfft, GateAddr ga, const int32_t *__restrict__ baseA,
                    const int32_t *__restrict__ baseB, int32_t *__restrict__ ext)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int grp = 0;
    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    if (g >= ga.ntempl * ga.n_inst) return;

    unsigned char *base = smem_raw;
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
    group_sync(grp);
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
    group_sync(grp);

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
     * only loads the rotated operand and stores the updated value */
    int32_t areg[2][16];
    if (ACCREG) {
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int h = 0; h < 16; h++) areg[q][h] = acc[q * kN + tid + 64 * (h & 7) + 512 * (h >> 3)];
    }

    /* 3. n CMux steps */
    for (int i = 0; i < n; i++) {
        const int a = abar[i];
        if (a == 0) continue; /* (X^0 - 1) ACC = 0: the step adds exactly zero; uniform inside the group */
        double s0r[8], s0i[8], s1r[8], s1i[8];
#pragma unroll
        for (int r = 0; r < 8; r++) { s0r[r] = 0.0; s0i[r] = 0.0; s1r[r] = 0.0; s1i[r] = 0.0; }
        const double2 *bk_r = bkfft + (size_t)i * kBkStride + tid;

#pragma unroll
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
#pragma unroll(ROLL == 2 ? 1 : L)
            for (int pp = 0; pp < L; pp++) {
                const int shift = 32 - (pp + 1) * Bgbit;
                double xr[8], xi[8];
                cd *buf = toggle ? bufB : bufA;
                toggle ^= 1;
                if (ACCREG) { /* the transform starts from the integer digits (cheaper first stage, br_core.h) */
                    int32_t dr[8], di[8];
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        dr[m] = digit_i32(c[m], offset, shift, maskBg, halfBg);
                        di[m] = digit_i32(c[8 + m], offset, shift, maskBg, halfBg);
                    }
                    fwd_transform_digits(dr, di, xr, xi, buf, tid, grp, w1, w2, w3);
                } else {
#pragma unroll
                    for (int m = 0; m < 8; m++) {
                        xr[m] = digit_f64(c[m], offset, shift, maskBg, halfBg);
                        xi[m] = digit_f64(c[8 + m], offset, shift, maskBg, halfBg);
                    }
                    fwd_transform(xr, xi, buf, tid, grp, w1, w2, w3);
                }
#pragma unroll
                for (int r = 0; r < 8; r++) {
                    const double2 b0 = __ldg(bk_r + r * 64), b1 = __ldg(bk_r + kHalfN + r * 64);
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
            inv_transform(s0r, s0i, buf, tid, grp, w1, w2, w3);
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
        group_sync(grp);
    }

    /* 4. SampleExtract at index 0 */
    int32_t *o = ext + (size_t)g * kExtStride;
    for (int j = tid; j < kN; j += 64) o[j] = (j == 0) ? acc[0] : -acc[kN - j];
    if (tid == 0) o[kN] = acc[kN];
}

/* ---- latency variant with two groups per gate: group q owns ACC polynomial q, runs its l forward transforms
 * with register accumulators for both output polynomials, hands the partial sum of the *other* polynomial to
 * the other group through 8 KB of shared memory, and inverts / updates its own polynomial.  Per step: 4
 * transform latencies instead of 8, and only 32 KB of extra shared-memory traffic (an earlier kernel with one
 * group per forward transform moved 192 KB per step and was LSU-bound at 72 %). */
constexpr int pair_smem_bytes() { return kAccBytes + 2 * 2 * kBufBytes + 2 * kHalfN * 16 + kAbarBytes; }

template <int L>
__global__ void __launch_bounds__(128, 2)
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
#pragma unroll
            for (int r = 0; r < 8; r++) { b0[r] = __ldg(bk_mine + pp * kRowElems + r * 64); b1[r] = __ldg(bk_other + pp * kRowElems + r * 64); }
            fwd_transform(xr, xi, buf, tid, grp, w1, w2, w3);
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
#ifdef IEACHE_CLUSTER_PROF /* developer build: per-phase cycle counts of cluster 0 on stdout */
    constexpr bool prof = true;
#else
    constexpr bool prof = false;
#endif
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_cluster_kernel<L, prof>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    blind_rotate_cluster_kernel<L, prof><<<(int)count * 2, 64 * L, smem, s>>>(p, bkfft, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

template <int L>
static cudaError_t launch_br_pair(const DevParams &p, const double2 *bkfft, const GateAddr &ga, const int32_t *baseA,
                                  const int32_t *baseB, int32_t *ext, long long count, cudaStream_t s)
{
    constexpr int smem = pair_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_pair_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    blind_rotate_pair_kernel<L><<<(int)count, 128, smem, s>>>(p, bkfft, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

template <int L, int ROLL, bool ACCREG>
static cudaError_t launch_br_group(const DevParams &p, const double2 *bkfft, const GateAddr &ga, const int32_t *baseA,
                                   const int32_t *baseB, int32_t *ext, long long count, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_kernel<L, ROLL, ACCREG>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGroupSmem);
    if (e != cudaSuccess) return e;
    blind_rotate_kernel<L, ROLL, ACCREG><<<(int)count, 64, kGroupSmem, s>>>(p, bkfft, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

/* Which kernel a launch of `count` gates uses (measured table in DESIGN.md 4):
 *   <= cluster_max            one gate on a 2-CTA cluster (1.6 ms per gate, 74 clusters per wave)
 *   <= pair_max               one gate per CTA, two groups (2.1 ms, 148-gate steps)
 *   >= w12_min                persistent kernel, 12 gates per SM (1 776-gate rounds)
 *   between                   group kernel (592-gate waves), or the two-group kernel when the group kernel's last
 *                             wave would be less than 90 % full */
int pick_blind_rotate(const LaunchPolicy &pol, long long count)
{
    if (count <= pol.cluster_max && count <= pol.pair_max) return BR_CLUSTER;
    if (count <= pol.pair_max) return BR_PAIR;
    if (pol.throughput == BR_GROUP || pol.throughput == BR_W12) return pol.throughput;
    if (count >= pol.w12_min) return BR_W12;
    if (pol.pair_max > 0) {
        const long long slots = 4LL * pol.sms, waves = (count + slots - 1) / slots;
        if (count * 10 < waves * slots * 9) return BR_PAIR;
    }
    return BR_GROUP;
}

cudaError_t launch_blind_rotate(const DevParams &p, const LaunchPolicy &pol, const double2 *bkfft, const double2 *bkfft_w, const GateAddr &ga,
                                const int32_t *baseA, const int32_t *baseB, int32_t *ext, int ext_base, cudaStream_t s)
{
    const long long count = (long long)ga.ntempl * ga.n_inst;
    if (count <= 0) return cudaSuccess;
    if (p.l != 2 && p.l != 3) return cudaErrorInvalidValue;
    ext += (size_t)ext_base * kExtStride;
    switch (pick_blind_rotate(pol, count)) {
    case BR_CLUSTER:
        return p.l == 3 ? launch_br_cluster<3>(p, bkfft, ga, baseA, baseB, ext, count, s) : launch_br_cluster<2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case BR_PAIR:
        return p.l == 3 ? launch_br_pair<3>(p, bkfft, ga, baseA, baseB, ext, count, s) : launch_br_pair<2>(p, bkfft, ga, baseA, baseB, ext, count, s);
    case BR_W12:
        if (!bkfft_w) return cudaErrorInvalidValue; /* the caller lays the key out when pick_blind_rotate says BR_W12 */
        return launch_blind_rotate_w12(p, bkfft_w, ga, baseA, baseB, ext, count, pol.sms, s);
    default:
        /* l = 3: unrolled step body with the ACC coefficients in registers; l = 2 keeps the rolled form it was validated in */
        return p.l == 3 ? launch_br_group<3, 0, true>(p, bkfft, ga, baseA, baseB, ext, count, s)
                        : launch_br_group<2, 2, false>(p, bkfft, ga, baseA, baseB, ext, count, s);
    }
}

/* ------------------------------------------------------------------ key switch */
constexpr int kKsThreads = 160; /* 158 int4 lanes cover a 632-word row */

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
    extern __shared__ __align__(128) unsigned char ks_stage_smem[];
    unsigned char *ks_smem = ks_stage_smem;
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

/* Which key-switch kernel: launches of at most pair_max gates use the 8-CTA cluster kernel; launches of at least
 * ks_staged_min gates the staged kernel (a CTA needs ~0.96 ms for its 12 gates whatever the launch size, the gather
 * kernel ~0.95 us per gate, so they cross near 1000 gates) */
int pick_keyswitch(const LaunchPolicy &pol, const DevParams &p, long long count)
{
    if (count <= pol.pair_max && p.ks_t <= 16) return KS_CLUSTER;
    const int block_rows = p.ks_t * ((1 << p.ks_basebit) - 1);
    if (block_rows <= 24 && count >= pol.ks_staged_min) return KS_STAGED;
    return KS_GATHER;
}

cudaError_t launch_keyswitch(const DevParams &p, const LaunchPolicy &pol, const int32_t *ksk, const GateAddr &ga, int32_t *out_base,
                             const int32_t *ext, int pair_offset, int32_t cst_post, cudaStream_t s)
{
    const long long count = (long long)ga.ntempl * ga.n_inst;
    if (count <= 0) return cudaSuccess;
    const int smem = kN * p.ks_t * 4;
    if (smem > 64 * 1024) return cudaErrorInvalidValue;
    /* function attributes are per device: set on every launch (as the blind-rotation launchers do) so that a process
     * with contexts on several GPUs gets the opt-in on each of them */
    switch (pick_keyswitch(pol, p, count)) {
    case KS_CLUSTER:
        keyswitch_cluster_kernel<<<(unsigned)count * kKsCluster, kKsThreads, 0, s>>>(p, ksk, ga, out_base, ext, pair_offset, cst_post);
        return cudaGetLastError();
    case KS_STAGED: {
        static_assert(kKsMaxBlockRows == 24, "pick_keyswitch assumes 24-row blocks");
        cudaError_t e = cudaFuncSetAttribute(keyswitch_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kKsStageSmem);
        if (e != cudaSuccess) return e;
        const unsigned grid = (unsigned)((count + kKsGates - 1) / kKsGates);
        keyswitch_staged_kernel<<<grid, kKsStageThreads, kKsStageSmem, s>>>(p, ksk, ga, out_base, ext, pair_offset, cst_post);
        return cudaGetLastError();
    }
    default: {
        cudaError_t e = cudaFuncSetAttribute(keyswitch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        keyswitch_kernel<<<(unsigned)count, kKsThreads, smem, s>>>(p, ksk, ga, out_base, ext, pair_offset, cst_post);
        return cudaGetLastError();
    }
    }
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

/* ------------------------------------------------------------------ operand blocks of the session layer */
/* copies `nblocks` blocks of 32 samples (32 x kLweStride words) between device arrays: src[b] / dst[b] are absolute
 * device addresses.  Used to assemble circuit inputs from, and to place circuit outputs into, the device-resident
 * operand batches of engine.cu (the splice Cloud/dragonfly_cipher_cloud.py:1306-1315 does on files). */
__global__ void __launch_bounds__(256) copy_blocks_kernel(const int4 *const *__restrict__ src, int4 *const *__restrict__ dst, int nblocks)
{
    const int b = blockIdx.x;
    if (b >= nblocks) return;
    const int4 *s = src[b];
    int4 *d = dst[b];
    constexpr int kVec = 32 * kLweStride / 4;
    for (int i = threadIdx.x; i < kVec; i += 256) d[i] = s[i];
}
cudaError_t launch_copy_blocks(const void *const *d_src, void *const *d_dst, int nblocks, cudaStream_t s)
{
    if (nblocks <= 0) return cudaSuccess;
    copy_blocks_kernel<<<nblocks, 256, 0, s>>>(reinterpret_cast<const int4 *const *>(d_src), reinterpret_cast<int4 *const *>(d_dst), nblocks);
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
