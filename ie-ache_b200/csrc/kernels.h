/*
 * kernels.h — launch interface of the sm_100a kernels (plain C++ types, no torch).
 * Device data layout (DESIGN.md "Data layout in HBM"):
 *   LWE sample      int32[stride], a[0..n) then b at word n; internal stride = 632 words
 *   extracted LWE   int32[1028]:   a[0..1024) then b at word 1024
 *   bkfft           double2[n][kpl][2][8][64]   evaluation K = b+8k'+64*brev3(r) at [r][8b+k'],
 *                   pre-scaled by 1/512 (the inverse transform is unnormalised)
 *   ksk             int32[1024][t][base-1][632] rows for digits d = 1..base-1 (d = 0 never read)
 */
#ifndef IEACHE_KERNELS_H
#define IEACHE_KERNELS_H
#include <cuda_runtime.h>
#include <stdint.h>

namespace ieache {

constexpr int kLweStride = 632;   /* words per LWE sample in device buffers (n <= 631) */
constexpr int kExtStride = 1028;  /* words per extracted sample */
constexpr int kMaxN = 1024;       /* max LWE dimension supported by the kernels */

/* One bootstrapped gate of a level template: x = (0, cst_mu*mu) + c0*S[in0] + c1*S[in1],
 * result written to S[out]; in/out are sample indices inside one instance's wire block. */
struct GateT {
    int32_t in0, in1;     /* sample index, or -1 */
    int32_t out;
    int8_t c0, c1;        /* coefficients in {+-1, +-2} */
    int16_t cst_mu;       /* constant added to b, in units of mu */
};

/* Addressing of a launch: gate g = e*ntempl + t reads sample (e*inst_samples + idx) of its base
 * array.  tmpl == nullptr means a homogeneous batch: idx = t for both inputs and the output,
 * coefficients taken from `uni`. */
struct GateAddr {
    const GateT *tmpl;
    int32_t ntempl;
    int32_t n_inst;
    int32_t inst_samples; /* samples per instance block */
    int32_t stride;       /* words per sample */
    GateT uni;
};

struct DevParams {
    int32_t n;            /* LWE dimension */
    int32_t l;            /* gadget length (2 or 3) */
    int32_t Bgbit;
    int32_t ks_t, ks_basebit;
    int32_t mu;           /* bootstrap message, 2^29 */
};

/* Launch policy of one context: which kernel shape a launch of a given size uses.  Held by ieache_ctx (no process-wide
 * state), initialised from the defaults below and from IEACHE_* environment overrides at context creation. */
enum { BR_CLUSTER = 1, BR_PAIR = 2, BR_GROUP = 41, BR_W12 = 70 };
enum { KS_CLUSTER = 1, KS_GATHER = 2, KS_STAGED = 3 };
struct LaunchPolicy {
    long long cluster_max = 74;     /* <= : one gate on a 2-CTA cluster (74 clusters in one wave) */
    long long pair_max = 296;       /* <= : one gate per CTA, two groups; also the 8-CTA-cluster key switch */
    long long w12_min = 900;        /* >= : persistent 12-warp kernel (1 776 gates per round) */
    long long ks_staged_min = 1000; /* >= : staged key switch */
    int throughput = 0;             /* 0 = by size; BR_GROUP / BR_W12 = that kernel for every launch above pair_max */
    int sms = 148;
};
int pick_blind_rotate(const LaunchPolicy &pol, long long count);
int pick_keyswitch(const LaunchPolicy &pol, const DevParams &p, long long count);

/* one-time: upload the twiddle tables to the current device */
cudaError_t upload_twiddles();

/* key load: coefficient-domain BK polys int32[npoly][1024] -> bkfft layout.
 * poly index q = ((i*kpl + r)*2 + j). */
cudaError_t launch_bk_fft(const int32_t *bk_coef, double2 *bkfft, int npoly, cudaStream_t s);

/* blind rotation + sample extraction of ntempl*n_inst gates -> ext[ext_base + g][1028].
 * bkfft_w: the same key in the persistent kernel's layout [row][poly][slot 16][lane 32] (launch_bk_relayout_w12);
 * required when pick_blind_rotate(pol, count) == BR_W12 */
cudaError_t launch_blind_rotate(const DevParams &p, const LaunchPolicy &pol, const double2 *bkfft, const double2 *bkfft_w, const GateAddr &ga,
                                const int32_t *baseA, const int32_t *baseB, int32_t *ext, int ext_base, cudaStream_t s);
/* br_w12.cu: persistent warp-per-gate blind rotation (12 gates per SM) */
cudaError_t upload_twiddles_w12();
cudaError_t launch_blind_rotate_w12(const DevParams &p, const double2 *bkfft_w, const GateAddr &ga, const int32_t *baseA,
                                    const int32_t *baseB, int32_t *ext, long long count, int sms, cudaStream_t s);
cudaError_t launch_bk_relayout_w12(const double2 *bkfft, double2 *bkfft_w, int npoly, int l, int Bgbit, cudaStream_t s);
/* key switch of ext[g] (+ ext[g + pair_offset] if pair_offset > 0) + (0, cst_post) into sample out of out_base */
cudaError_t launch_keyswitch(const DevParams &p, const LaunchPolicy &pol, const int32_t *ksk, const GateAddr &ga, int32_t *out_base,
                             const int32_t *ext, int pair_offset, int32_t cst_post, cudaStream_t s);

/* free linear ops on whole arrays (bootsNOT / bootsCOPY / bootsCONSTANT) */
cudaError_t launch_linear(int32_t *out, const int32_t *a, int count, int stride, int n, int coef, int cst, cudaStream_t s);

/* repack n-word key-switch rows: src [kN][t][base][n+1] (libtfhe order) -> dst [kN][t][base-1][632] */
cudaError_t launch_pack_ksk(const int32_t *src, int32_t *dst, int kN, int t, int base, int n, cudaStream_t s);

/* circuit wire blocks: [const0, const1, inputs..., gate slots...] per instance */
cudaError_t launch_circuit_scatter_inputs(int32_t *wires, const int32_t *inputs, int n_expr, int n_inputs, int n_slots, int n,
                                          int32_t mu, cudaStream_t s);
cudaError_t launch_circuit_gather_outputs(int32_t *outputs, const int32_t *wires, const int32_t *out_slots, int n_expr,
                                          int n_outputs, int n_slots, int n, cudaStream_t s);

/* session layer: copy blocks of 32 samples between device arrays given as arrays of device addresses */
cudaError_t launch_copy_blocks(const void *const *d_src, void *const *d_dst, int nblocks, cudaStream_t s);

/* dense FP64 FMA microbenchmark (best of 3), TFLOP/s */
cudaError_t launch_fp64_peak(cudaStream_t s, double *tflops);

} // namespace ieache
#endif
