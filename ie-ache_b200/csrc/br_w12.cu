/*
 * br_w12.cu — persistent warp-per-gate blind rotation: 12 gates resident per SM.
 *
 * Same arithmetic as the other blind-rotation kernels (replaces libtfhe's tfhe_blindRotate_FFT /
 * tGswFFTExternMulToTLwe reached from Cloud/cloud.c:30-43), laid out for what bounds it on this
 * part: instruction issue — three warps per scheduler instead of two, and as few instructions next to
 * the FP64 work as the layout allows.  One WARP owns one gate with 16 points per lane (br_warp.h): a transform needs
 * one shared-memory exchange and one shuffle stage instead of two shared-memory exchanges, and nothing in the CMux
 * loop crosses a warp.  What makes 12 warps fit on an SM (the 64-thread kernel holds 8) is tensor memory:
 *
 *   TMEM columns (per 32-lane quadrant)   [160 s, 160 s + 128)        transform-domain accumulators of warp slot s = warp / 4
 *                                         [160 s + 128, 160 s + 160)  the 32 rotated differences (X^a - 1) ACC_q of the lane
 *                                         [480, 512)                  pass-2 twiddles of the lane (shared by the 3 slots)
 *   registers (<= 168 per thread)         the 16 points in flight, one 4-slot chunk of accumulators and key values
 *   shared memory (17 920 B per warp)     ACC (8 KB), one exchange buffer (16 x 33 complex), the mod-switched mask;
 *                                         per CTA: the final-stage twiddles (4 complex per lane: w[s + 1] = i w[s]) and
 *                                         z8 of pass 2
 * Forward and inverse transforms are the select-free ("folded") forms of br_warp.h: every lane runs the same
 * butterflies and sends the same registers; the key is read in the w12_slot_to_K order with its plain values.
 *
 * One CTA of 12 warps per SM, launched once; warp w of CTA b takes gates b + grid * (w + 12 k), so a launch of any
 * size is spread over all SMs first and has no tail wave of half-empty CTAs.
 */
#include "kernels.h"
#include "br_core.h"
#include "br_warp.h"

#include <stdlib.h>

/* Compile-time switches of the A/B measurements in DESIGN.md 4.3 (tools/w12_variants.sh builds one library per
 * combination; the shipped library is the defaults): W12_TM2 pair-wise (1) or quad (0) tensor-memory accumulate,
 * W12_NOINLINE phases as functions, W12_ROT_ASM select-free rotation, W12_INV_DIT 6-FMA last inverse pass,
 * W12_ATOMS shared-memory atomic ACC update. */
#ifndef W12_INV_DIT
#define W12_INV_DIT 1
#endif
#ifndef W12_ATOMS
#define W12_ATOMS 1
#endif
#ifndef W12_ROT_ASM
#define W12_ROT_ASM 1
#endif
#ifndef W12_ROUND
#define W12_ROUND(v) ((int32_t)__double2ll_rn(v)) /* F2I on the conversion pipe: one FP64 issue less per coefficient than the 1.5 * 2^52 trick */
#endif

namespace ieache {

/* pass-1 twiddles (br_warp.h tw16_pass1) in constant memory: DFMA takes them as c[bank][offset] operands; as literals
 * the compiler rebuilt every one of them in uniform registers inside the loop (62 UMOV per transform) */
__constant__ TwInvP1 c12_inv1 = IE_TW_INV_P1_INIT;
__constant__ int c12_m1 = -1; /* a run-time -1: keeps "offset - x" on the multiply-add pipe */
__constant__ Tw16 c12_w1 = {0.70710678118654752440, 0.70710678118654752440, 0.92387953251128675613, 0.38268343236508977173,
                            0.98078528040323044913, 0.19509032201612826785, 0.55557023301960222474, 0.83146961230254523708,
                            0.99518472667219688624, 0.09801714032956060199, 0.63439328416364549822, 0.77301045336273696081,
                            0.88192126434835502971, 0.47139673682599764856, 0.29028467725446236764, 0.95694033573220886494};
__device__ Tw16g d12_tw16g[32]; /* pass-2 twiddles by lane (block twiddles exchanged on the Lpar = 1 lanes, br_warp.h) */
__device__ FinTw d12_finf[32];  /* final-stage twiddles by lane (conjugated on the Lpar = 1 lanes) */

cudaError_t upload_twiddles_w12()
{
    Tw16g tw16g[32];
    FinTw finf[32];
    host_twiddles_warp_folded(tw16g, finf);
    cudaError_t e = cudaMemcpyToSymbol(d12_tw16g, tw16g, sizeof(tw16g));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(d12_finf, finf, sizeof(finf));
}

constexpr int kW12Warps = 12;
constexpr int kW12AbarBytes = 1280;                                       /* 632 x uint16, padded */
constexpr int kW12GateSmem = 2 * kN * 4 + kWarpBufElems * 16 + kW12AbarBytes; /* 17 920 B */
constexpr int kW12Smem = kW12Warps * kW12GateSmem;                        /* 215 040 B */
constexpr uint32_t kW12SlotCols = 160, kW12ColCc = 128, kW12ColTw2 = 480;

__device__ __forceinline__ uint32_t smem_u32_w12(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* 16 columns = 8 doubles, 32 columns = 16 doubles of the calling lane.  The b32 halves are packed inside the asm so
 * that ptxas allocates them as the register pairs of the doubles. */
#define W12_LD16D(taddr, d) asm volatile("{\n\t.reg .b32 t<32>;\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15,t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31}, [%16];\n\t" \
    "tcgen05.wait::ld.sync.aligned;\n\t" \
    "mov.b64 %0, {t0,t1};\n\tmov.b64 %1, {t2,t3};\n\tmov.b64 %2, {t4,t5};\n\tmov.b64 %3, {t6,t7};\n\tmov.b64 %4, {t8,t9};\n\tmov.b64 %5, {t10,t11};\n\tmov.b64 %6, {t12,t13};\n\tmov.b64 %7, {t14,t15};\n\tmov.b64 %8, {t16,t17};\n\tmov.b64 %9, {t18,t19};\n\tmov.b64 %10, {t20,t21};\n\tmov.b64 %11, {t22,t23};\n\tmov.b64 %12, {t24,t25};\n\tmov.b64 %13, {t26,t27};\n\tmov.b64 %14, {t28,t29};\n\tmov.b64 %15, {t30,t31};\n\t}" \
    : "=d"(d[0]),"=d"(d[1]),"=d"(d[2]),"=d"(d[3]),"=d"(d[4]),"=d"(d[5]),"=d"(d[6]),"=d"(d[7]),"=d"(d[8]),"=d"(d[9]),"=d"(d[10]),"=d"(d[11]),"=d"(d[12]),"=d"(d[13]),"=d"(d[14]),"=d"(d[15]) : "r"(taddr) : "memory")
#define W12_ST16D(taddr, d) asm volatile("{\n\t.reg .b32 t<32>;\n\t" \
    "mov.b64 {t0,t1}, %1;\n\tmov.b64 {t2,t3}, %2;\n\tmov.b64 {t4,t5}, %3;\n\tmov.b64 {t6,t7}, %4;\n\tmov.b64 {t8,t9}, %5;\n\tmov.b64 {t10,t11}, %6;\n\tmov.b64 {t12,t13}, %7;\n\tmov.b64 {t14,t15}, %8;\n\tmov.b64 {t16,t17}, %9;\n\tmov.b64 {t18,t19}, %10;\n\tmov.b64 {t20,t21}, %11;\n\tmov.b64 {t22,t23}, %12;\n\tmov.b64 {t24,t25}, %13;\n\tmov.b64 {t26,t27}, %14;\n\tmov.b64 {t28,t29}, %15;\n\tmov.b64 {t30,t31}, %16;\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15,t16,t17,t18,t19,t20,t21,t22,t23,t24,t25,t26,t27,t28,t29,t30,t31};\n\t}" \
    :: "r"(taddr), "d"(d[0]),"d"(d[1]),"d"(d[2]),"d"(d[3]),"d"(d[4]),"d"(d[5]),"d"(d[6]),"d"(d[7]),"d"(d[8]),"d"(d[9]),"d"(d[10]),"d"(d[11]),"d"(d[12]),"d"(d[13]),"d"(d[14]),"d"(d[15]) : "memory")
/* 32 columns = 32 words */
#define W12_LD32W(taddr, v) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t" \
    "tcgen05.wait::ld.sync.aligned;" \
    : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]), \
      "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr) : "memory")
#define W12_ST32W(taddr, v) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
    :: "r"(taddr), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]), \
       "r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory")
/* 16 columns of zeros */
#define W12_ST_ZERO16(taddr) asm volatile("{\n\t.reg .b32 z;\n\tmov.b32 z, 0;\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {z,z,z,z,z,z,z,z,z,z,z,z,z,z,z,z};\n\t}" :: "r"(taddr) : "memory")

/* 4 slots (16 columns) as four 4-register accesses: a slot's {re, im} is one aligned register quad, the same
 * constraint ptxas already solves for 16-byte shared-memory accesses, so the accumulate needs no register moves
 * (a 16-register vector made ptxas copy 12-16 registers per chunk between its load block and its store block) */
#define W12_LD4x4(taddr, d) asm volatile("{\n\t.reg .b32 t<16>;\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x4.b32 {t0,t1,t2,t3}, [%8];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x4.b32 {t4,t5,t6,t7}, [%8 + 4];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x4.b32 {t8,t9,t10,t11}, [%8 + 8];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x4.b32 {t12,t13,t14,t15}, [%8 + 12];\n\t" \
    "tcgen05.wait::ld.sync.aligned;\n\t" \
    "mov.b64 %0, {t0,t1};\n\tmov.b64 %1, {t2,t3};\n\tmov.b64 %2, {t4,t5};\n\tmov.b64 %3, {t6,t7};\n\tmov.b64 %4, {t8,t9};\n\tmov.b64 %5, {t10,t11};\n\tmov.b64 %6, {t12,t13};\n\tmov.b64 %7, {t14,t15};\n\t}" \
    : "=d"(d[0]),"=d"(d[1]),"=d"(d[2]),"=d"(d[3]),"=d"(d[4]),"=d"(d[5]),"=d"(d[6]),"=d"(d[7]) : "r"(taddr) : "memory")
#define W12_ST4x4(taddr, d) asm volatile("{\n\t.reg .b32 t<16>;\n\t" \
    "mov.b64 {t0,t1}, %1;\n\tmov.b64 {t2,t3}, %2;\n\tmov.b64 {t4,t5}, %3;\n\tmov.b64 {t6,t7}, %4;\n\tmov.b64 {t8,t9}, %5;\n\tmov.b64 {t10,t11}, %6;\n\tmov.b64 {t12,t13}, %7;\n\tmov.b64 {t14,t15}, %8;\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {t0,t1,t2,t3};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0 + 4], {t4,t5,t6,t7};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0 + 8], {t8,t9,t10,t11};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0 + 12], {t12,t13,t14,t15};\n\t}" \
    :: "r"(taddr), "d"(d[0]),"d"(d[1]),"d"(d[2]),"d"(d[3]),"d"(d[4]),"d"(d[5]),"d"(d[6]),"d"(d[7]) : "memory")

#ifndef W12_TM2
#define W12_TM2 1
#endif
#if W12_TM2
/* ... and as eight 2-register accesses: a double is an aligned register pair wherever it lives, so the accumulate has
 * no placement constraint at all (with quads ptxas still moved 60-110 registers per transform between the load, the
 * two dependent FMAs and the store); 32 more tensor-memory instructions per transform, each a single issue slot */
#undef W12_LD4x4
#undef W12_ST4x4
#define W12_LD4x4(taddr, d) asm volatile("{\n\t.reg .b32 t<16>;\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t0,t1}, [%8];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t2,t3}, [%8 + 2];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t4,t5}, [%8 + 4];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t6,t7}, [%8 + 6];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t8,t9}, [%8 + 8];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t10,t11}, [%8 + 10];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t12,t13}, [%8 + 12];\n\t" \
    "tcgen05.ld.sync.aligned.32x32b.x2.b32 {t14,t15}, [%8 + 14];\n\t" \
    "tcgen05.wait::ld.sync.aligned;\n\t" \
    "mov.b64 %0, {t0,t1};\n\tmov.b64 %1, {t2,t3};\n\tmov.b64 %2, {t4,t5};\n\tmov.b64 %3, {t6,t7};\n\tmov.b64 %4, {t8,t9};\n\tmov.b64 %5, {t10,t11};\n\tmov.b64 %6, {t12,t13};\n\tmov.b64 %7, {t14,t15};\n\t}" \
    : "=d"(d[0]),"=d"(d[1]),"=d"(d[2]),"=d"(d[3]),"=d"(d[4]),"=d"(d[5]),"=d"(d[6]),"=d"(d[7]) : "r"(taddr) : "memory")
#define W12_ST4x4(taddr, d) asm volatile("{\n\t.reg .b32 t<16>;\n\t" \
    "mov.b64 {t0,t1}, %1;\n\tmov.b64 {t2,t3}, %2;\n\tmov.b64 {t4,t5}, %3;\n\tmov.b64 {t6,t7}, %4;\n\tmov.b64 {t8,t9}, %5;\n\tmov.b64 {t10,t11}, %6;\n\tmov.b64 {t12,t13}, %7;\n\tmov.b64 {t14,t15}, %8;\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {t0,t1};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0 + 2], {t2,t3};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0 + 4], {t4,t5};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0 + 6], {t6,t7};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0 + 8], {t8,t9};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0 + 10], {t10,t11};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0 + 12], {t12,t13};\n\t" \
    "tcgen05.st.sync.aligned.32x32b.x2.b32 [%0 + 14], {t14,t15};\n\t" \
    "}" \
    :: "r"(taddr), "d"(d[0]),"d"(d[1]),"d"(d[2]),"d"(d[3]),"d"(d[4]),"d"(d[5]),"d"(d[6]),"d"(d[7]) : "memory")
#endif

/* one coefficient of the rotated difference, halved with the decomposition offset added:
 * cc = (sign * ACC[(j - a) mod N] - ACC[j] + offset) >> 1 */
#define W12_ROT(out, aR, aJ, lim, offs, sN, sW, m1, OFF) asm volatile("{\n\t.reg .pred p;\n\t.reg .s32 v, u;\n\t" \
    "setp.le.s32 p, %3, %8;\n\t" \
    "@!p ld.shared.s32 v, [%1 + %9];\n\t" \
    "@p ld.shared.s32 v, [%1 + %10];\n\t" \
    "ld.shared.s32 u, [%2 + %9];\n\t" \
    "mad.lo.s32 u, u, %7, %4;\n\t" \
    "@!p mad.lo.s32 u, v, %5, u;\n\t" \
    "@p mad.lo.s32 u, v, %6, u;\n\t" \
    "shr.u32 %0, u, 1;\n\t}" \
    : "=r"(out) : "r"(aR), "r"(aJ), "r"(lim), "r"(offs), "r"(sN), "r"(sW), "r"(m1), "n"(OFF), "n"(4 * (OFF)), "n"(4 * (OFF) - 4 * kN) : "memory")

template <int H>
__device__ __forceinline__ void w12_rot_all(uint32_t (&cc)[32], uint32_t aR, uint32_t aJ, int lim, uint32_t offs, int sN, int sW, int m1)
{
    if constexpr (H < 32) {
        constexpr int off = 32 * (H & 15) + 512 * (H >> 4);
        W12_ROT(cc[H], aR, aJ, lim, offs, sN, sW, m1, off);
        w12_rot_all<H + 1>(cc, aR, aJ, lim, offs, sN, sW, m1);
    }
}

__device__ __forceinline__ double2 ldg_nc_pinned(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

/* final-stage twiddles of butterflies s = 4 h .. 4 h + 3 from the lane's 4-entry table (even s; w[s + 1] = i w[s]) */
template <int H>
__device__ __forceinline__ void w12_fin_tw(const double2 *s_tab, int lane, double (&z)[8])
{
    const double2 za = s_tab[(2 * H) * 32 + lane], zb = s_tab[(2 * H + 1) * 32 + lane];
    fin4_tw(za.x, za.y, zb.x, zb.y, z);
}
/* final stage of the forward transform, butterflies s = 4 h .. 4 h + 3: every lane sends registers 8 + s, keeps s
 * and computes keep +- w recv (br_warp.h, "folded" form: no lane-dependent selects) */
template <int H>
__device__ __forceinline__ void w12_fin_fwd_half(double (&yr)[16], double (&yi)[16], int lane, const double2 *s_tab)
{
    double rr[4], ri[4], z[8];
#pragma unroll
    for (int s = 0; s < 4; s++) { rr[s] = __shfl_xor_sync(0xffffffffu, yr[8 + 4 * H + s], 16); ri[s] = __shfl_xor_sync(0xffffffffu, yi[8 + 4 * H + s], 16); }
    w12_fin_tw<H>(s_tab, lane, z);
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int k = 4 * H + s;
        bf(yr[k], yi[k], rr[s], ri[s], z[2 * s], z[2 * s + 1]);
        yr[8 + k] = rr[s]; yi[8 + k] = ri[s];
    }
}
/* its inverse: the same butterfly on every lane, registers 8 + s travel */
template <int H>
__device__ __forceinline__ void w12_fin_inv_half(double (&xr)[16], double (&xi)[16], int lane, const double2 *s_tab)
{
    double z[8];
    w12_fin_tw<H>(s_tab, lane, z);
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int k = 4 * H + s;
        ibf(xr[k], xi[k], xr[8 + k], xi[8 + k], z[2 * s], z[2 * s + 1]);
    }
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int k = 4 * H + s;
        xr[8 + k] = __shfl_xor_sync(0xffffffffu, xr[8 + k], 16); xi[8 + k] = __shfl_xor_sync(0xffffffffu, xi[8 + k], 16);
    }
}
/* pass-2 twiddles of the lane: 8 complex values from the quadrant's shared TMEM columns, z8 from shared memory */
__device__ __forceinline__ void w12_load_tw2(Tw16g &w, uint32_t t_tw2, const double2 *s_tab, int lane);

__device__ __forceinline__ void w12_load_tw2(Tw16g &w, uint32_t t_tw2, const double2 *s_tab, int lane)
{
    double tw[16];
    W12_LD16D(t_tw2, tw);
    const double2 z8 = s_tab[4 * 32 + lane];
    w.z8r = z8.x; w.z8i = z8.y;
    w.z4ar = tw[0]; w.z4ai = tw[1]; w.z4br = tw[2]; w.z4bi = tw[3];
    w.z2ar = tw[4]; w.z2ai = tw[5]; w.z2br = tw[6]; w.z2bi = tw[7];
    w.z1ar = tw[8]; w.z1ai = tw[9]; w.z1aqr = tw[10]; w.z1aqi = tw[11];
    w.z1br = tw[12]; w.z1bi = tw[13]; w.z1bqr = tw[14]; w.z1bqi = tw[15];
}

#ifndef W12_NOINLINE
#define W12_NOINLINE 1
#endif
#if W12_NOINLINE
#define W12_PHASE __device__ __noinline__
#else
#define W12_PHASE __device__ __forceinline__
#endif
/* phase 1 of a CMux step: cc = ((X^a - 1) ACC_q + offset) >> 1 for the lane's 32 coefficients.  The phases are separate
 * functions so that each gets its own register allocation: inlined, a change in one phase reshuffles the register quads
 * of the forward transform and costs it up to 90 register moves */
W12_PHASE void w12_rotate(const int32_t *acc, int q, int a, int lane, uint32_t offset, uint32_t t_cc)
{
    uint32_t cc[32];
#if W12_ROT_ASM
    /* coefficient j = lane + off of (X^a - 1) ACC_q: the rotated source sits at R0 + off, wrapped once the
     * sum reaches N — so both candidate addresses are "register + immediate" and one compare picks the load
     * and the sign.  IMAD issues at full rate, the INT32 ALU at half rate: the sign is a multiply-add */
    const uint32_t sa = smem_u32_w12(acc) + (uint32_t)q * (kN * 4);
    int aq = a;
    asm volatile("" : "+r"(aq)); /* recomputed per polynomial: nothing of this stays live across the transforms */
    const int t0 = (lane - aq) & (2 * kN - 1);
    const int r0 = t0 & (kN - 1);
    const int sN = (t0 & kN) ? -1 : 1, sW = -sN;
    const int lim = kN - r0;
    const uint32_t aR = sa + 4u * (uint32_t)r0, aJ = sa + 4u * (uint32_t)lane;
    w12_rot_all<0>(cc, aR, aJ, lim, offset, sN, sW, c12_m1);
#else
    const int32_t *accq = acc + q * kN;
    int aq = a;
    asm volatile("" : "+r"(aq));
    const int t0 = lane - aq;
#pragma unroll
    for (int h = 0; h < 32; h++) {
        const int off = 32 * (h & 15) + 512 * (h >> 4);
        const int t = t0 + off;
        const int32_t v = accq[t & (kN - 1)];
        cc[h] = ((uint32_t)(((t & kN) ? -v : v) - accq[lane + off]) + offset) >> 1;
    }
#endif
    W12_ST32W(t_cc, cc);   /* digit fields stay in place (pass16_fwd_from_fields); read back once per digit */
}

/* phase 3: inverse transforms of both accumulator polynomials (which leaves them zero) and the ACC update */
W12_PHASE void w12_inverse(int32_t *acc, cd *buf, const double2 *s_tab, uint32_t t_acc, uint32_t t_tw2, int lane)
{
#if !W12_INV_DIT
    const Tw16 &w1 = c12_w1;
#endif
    /* inverse transforms and ACC update */
#pragma unroll 1
    for (int j = 0; j < 2; j++) {
        double xr[16], xi[16];
        {
            double s0[16], s1[16];
            W12_LD16D(t_acc + (uint32_t)(j * 64), s0);
            W12_LD16D(t_acc + (uint32_t)(j * 64 + 32), s1);
#pragma unroll
            for (int r = 0; r < 8; r++) { xr[r] = s0[2 * r]; xi[r] = s0[2 * r + 1]; xr[8 + r] = s1[2 * r]; xi[8 + r] = s1[2 * r + 1]; }
        }
#pragma unroll
        for (int c16 = 0; c16 < 4; c16++) W12_ST_ZERO16(t_acc + (uint32_t)(j * 64 + 16 * c16));
        w12_fin_inv_half<0>(xr, xi, lane, s_tab);
        w12_fin_inv_half<1>(xr, xi, lane, s_tab);
        {
            Tw16g w2;
            w12_load_tw2(w2, t_tw2, s_tab, lane);
            pass16_inv_g(xr, xi, w2);
        }
        __syncwarp();
        st16_ipass2(buf, lane, xr, xi);
        __syncwarp();
        ld16_ipass1(buf, lane, xr, xi);
#if W12_INV_DIT
        pass16_inv_p1(xr, xi, c12_inv1);
#else
        pass16_inv(xr, xi, w1);
#endif
        int32_t *accj = acc + j * kN;
#pragma unroll
        for (int m = 0; m < 16; m++) {
#if W12_ATOMS
            atomicAdd(&accj[lane + 32 * m], W12_ROUND(xr[m]));
            atomicAdd(&accj[lane + 32 * m + 512], W12_ROUND(xi[m]));
#else
            accj[lane + 32 * m] += W12_ROUND(xr[m]);
            accj[lane + 32 * m + 512] += W12_ROUND(xi[m]);
#endif
        }
    }
}

template <int L>
__global__ void __launch_bounds__(32 * kW12Warps, 1)
blind_rotate_w12_kernel(DevParams p, const double2 *__restrict__ bkw, GateAddr ga, const int32_t *__restrict__ baseA,
                        const int32_t *__restrict__ baseB, int32_t *__restrict__ ext)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t tmem_base_slot;
    __shared__ double2 s_tab[5 * 32]; /* per lane: final-stage twiddles of s = 0, 2, 4, 6, then z8 of pass 2 */
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char *base = smem_raw + (size_t)warp * kW12GateSmem;
    int32_t *acc = reinterpret_cast<int32_t *>(base);
    cd *buf = reinterpret_cast<cd *>(base + 2 * kN * 4);
    uint16_t *abar = reinterpret_cast<uint16_t *>(base + 2 * kN * 4 + kWarpBufElems * 16);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32_w12(&tmem_base_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t t_lane = tmem_base_slot + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t t_acc = t_lane + (uint32_t)(warp >> 2) * kW12SlotCols;
    const uint32_t t_cc = t_acc + kW12ColCc, t_tw2 = t_lane + kW12ColTw2;
    if (warp < 4) { /* the pass-2 twiddles of this quadrant's lanes (all but z8), read by its three warp slots */
        double tw[16];
        const double *src = reinterpret_cast<const double *>(&d12_tw16g[lane]) + 2;
#pragma unroll
        for (int v = 0; v < 16; v++) tw[v] = src[v];
        W12_ST16D(t_tw2, tw);
    } else if (warp == 4) {
#pragma unroll
        for (int h = 0; h < 4; h++) s_tab[h * 32 + lane] = make_double2(d12_finf[lane].zr[2 * h], d12_finf[lane].zi[2 * h]);
        s_tab[4 * 32 + lane] = make_double2(d12_tw16g[lane].z8r, d12_tw16g[lane].z8i);
    }
    /* accumulators start at zero; every step leaves them at zero again (the inverse transform clears what it reads) */
#pragma unroll
    for (int c16 = 0; c16 < 8; c16++) W12_ST_ZERO16(t_acc + 16 * c16);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    const int n = p.n;
    const Tw16 &w1 = c12_w1;
    const int Bgbit = p.Bgbit;
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
#pragma unroll
    for (int i = 1; i <= L; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    constexpr int kRowElems = 2 * kHalfN, kBkStride = 2 * L * kRowElems;
    const long long total = (long long)ga.ntempl * ga.n_inst;
    const long long slots = (long long)gridDim.x * kW12Warps;

    for (long long g = blockIdx.x + (long long)gridDim.x * warp; g < total; g += slots) {
        /* 1. linear pre-combination + modSwitch to Z_{2N} */
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
        __syncwarp();
        /* 2. ACC = (0, X^{2N-bbar} * mu * (1 + X + ... + X^{N-1})) */
        {
            const int bbar = abar[n];
            const int a = (2 * kN - bbar) & (2 * kN - 1), ar = a & (kN - 1);
            const bool flip = a >= kN;
            for (int j = lane; j < kN; j += 32) { acc[j] = 0; acc[kN + j] = ((j < ar) != flip) ? -p.mu : p.mu; }
        }
        __syncwarp();

        /* 3. n CMux steps */
        for (int i = 0; i < n; i++) {
            const int a = abar[i];
            if (a == 0) continue; /* (X^0 - 1) ACC = 0: the step adds exactly zero */
            const double2 *bk_r = bkw + (size_t)i * kBkStride + lane;
#pragma unroll 1
            for (int q = 0; q < 2; q++) {
                w12_rotate(acc, q, a, lane, offset, t_cc);
#pragma unroll 1
                for (int pp = 0; pp < L; pp++) {
                    const int sp = w12_field_shift(pp, Bgbit);
                    double xr[16], xi[16];
                    {
                        uint32_t cd[32];
                        if (pp == 0) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        W12_LD32W(t_cc, cd);
                        pass16_fwd_from_fields(cd, maskBg << sp, (uint32_t)halfBg << sp, xr, xi, w1);
                    }
                    __syncwarp();                      /* every lane has finished reading the buffer of the previous transform */
                    st16_pass1(buf, lane, xr, xi);
                    __syncwarp();
                    {
                        Tw16g w2;
                        w12_load_tw2(w2, t_tw2, s_tab, lane);
                        ld16_pass2(buf, lane, xr, xi);
                        pass16_fwd_g(xr, xi, w2);
                    }
                    w12_fin_fwd_half<0>(xr, xi, lane, s_tab);
                    w12_fin_fwd_half<1>(xr, xi, lane, s_tab);
                    /* multiply-accumulate into TMEM, 4 slots of one output polynomial at a time; the key values of the
                     * next chunk are requested before the current one is computed */
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    double2 bj[4];
#pragma unroll
                    for (int r = 0; r < 4; r++) bj[r] = ldg_nc_pinned(bk_r + r * 32);
#pragma unroll
                    for (int ch = 0; ch < 8; ch++) {
                        const int sl = 4 * (ch & 3);                      /* slots sl .. sl + 3 of polynomial ch >> 2 */
                        const uint32_t tj = t_acc + (uint32_t)(16 * ch);
                        double2 bn[4];
                        if (ch + 1 < 8) {
#pragma unroll
                            for (int r = 0; r < 4; r++) bn[r] = ldg_nc_pinned(bk_r + ((ch + 1) >> 2) * kHalfN + (4 * ((ch + 1) & 3) + r) * 32);
                        }
                        double sacc[8];
                        W12_LD4x4(tj, sacc);
#pragma unroll
                        for (int r = 0; r < 4; r++) cmac(sacc[2 * r], sacc[2 * r + 1], xr[sl + r], xi[sl + r], bj[r].x, bj[r].y);
                        W12_ST4x4(tj, sacc);
                        if (ch + 1 < 8) {
#pragma unroll
                            for (int r = 0; r < 4; r++) bj[r] = bn[r];
                        }
                    }
                    bk_r += kRowElems;
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            w12_inverse(acc, buf, s_tab, t_acc, t_tw2, lane);
            __syncwarp();
        }

        /* 4. SampleExtract at index 0 */
        int32_t *o = ext + (size_t)g * kExtStride;
        for (int j = lane; j < kN; j += 32) o[j] = (j == 0) ? acc[0] : -acc[kN - j];
        if (lane == 0) o[kN] = acc[kN];
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base_slot));
}

template <int L>
static cudaError_t launch_w12_t(const DevParams &p, const double2 *bkw, const GateAddr &ga, const int32_t *baseA, const int32_t *baseB,
                                int32_t *ext, long long count, int sms, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(blind_rotate_w12_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, kW12Smem);
    if (e != cudaSuccess) return e;
    const int grid = (int)(count < sms ? count : sms);
    blind_rotate_w12_kernel<L><<<grid, 32 * kW12Warps, kW12Smem, s>>>(p, bkw, ga, baseA, baseB, ext);
    return cudaGetLastError();
}

cudaError_t launch_blind_rotate_w12(const DevParams &p, const double2 *bkw, const GateAddr &ga, const int32_t *baseA,
                                    const int32_t *baseB, int32_t *ext, long long count, int sms, cudaStream_t s)
{
    if (p.l == 3) return launch_w12_t<3>(p, bkw, ga, baseA, baseB, ext, count, sms, s);
    if (p.l == 2) return launch_w12_t<2>(p, bkw, ga, baseA, baseB, ext, count, sms, s);
    return cudaErrorInvalidValue;
}

/* key load: the [slot 8][thread 64] layout of bk_fft_kernel -> [slot 16][lane 32] in the w12_slot_to_K order; the rows
 * of gadget digit p also take the factor 2^-sp that the kernel's in-place digit fields carry (an exact scaling) */
__global__ void __launch_bounds__(512) bk_relayout_w12_kernel(const double2 *__restrict__ old, double2 *__restrict__ neu, int npoly, int l, int Bgbit)
{
    const int q = blockIdx.x, idx = threadIdx.x;   /* q = ((i kpl + r) 2 + j): digit p = r mod l */
    if (q >= npoly) return;
    const int K = w12_slot_to_K(idx >> 5, idx & 31);
    const int t3 = 8 * (K & 7) + ((K >> 3) & 7), r8 = brev3(K >> 6); /* br_core.h: K = b + 8k' + 64 brev3(r), t3 = 8b + k' */
    const double sc = ldexp(1.0, -w12_field_shift((q >> 1) % l, Bgbit));
    const double2 v = old[(size_t)q * kHalfN + r8 * 64 + t3];
    neu[(size_t)q * kHalfN + idx] = make_double2(v.x * sc, v.y * sc);
}
cudaError_t launch_bk_relayout_w12(const double2 *bkfft, double2 *bkfft_w, int npoly, int l, int Bgbit, cudaStream_t s)
{
    bk_relayout_w12_kernel<<<npoly, 512, 0, s>>>(bkfft, bkfft_w, npoly, l, Bgbit);
    return cudaGetLastError();
}

} // namespace ieache
