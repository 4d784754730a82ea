// Developer probe: FP64 cost of the real butterfly code (br_core.h pass_fwd / pass_inv / cmac) with everything in
// registers, at the occupancy of the blind-rotation kernel (64-thread CTAs, MINB per SM).  Prints cycles per FP64
// instruction per SM sub-partition; 2.0 is the pipe's nominal rate.
#include <cstdio>
#include <cuda_runtime.h>
#include "../ie-ache_b200/csrc/br_core.h"
using namespace ieache;
template <int MODE, int MINB>
__global__ void __launch_bounds__(64, MINB) probe(double *out, const Tw *tw, int iters)
{
    double xr[8], xi[8], s0r[8], s0i[8], s1r[8], s1i[8];
    const int tid = threadIdx.x;
    for (int m = 0; m < 8; m++) { xr[m] = tid + m; xi[m] = tid - m; s0r[m] = s0i[m] = s1r[m] = s1i[m] = 0.0; }
    const Tw w1 = tw_pass1(), w2 = tw[tid >> 3], w3 = tw[8 + tid];
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { pass_fwd(xr, xi, w1); pass_fwd(xr, xi, w2); pass_fwd(xr, xi, w3); }
        if (MODE == 1) { pass_inv(xr, xi, w3); pass_inv(xr, xi, w2); pass_inv(xr, xi, w1); }
        if (MODE == 2) {
#pragma unroll
            for (int r = 0; r < 8; r++) { cmac(s0r[r], s0i[r], xr[r], xi[r], w2.s1r, w3.s1i); cmac(s1r[r], s1i[r], xr[r], xi[r], w3.s2r, w2.s2i); }
        }
        if (MODE == 3) { /* forward transform + MAC, like one row of a CMux step without the exchanges */
            pass_fwd(xr, xi, w1); pass_fwd(xr, xi, w2); pass_fwd(xr, xi, w3);
#pragma unroll
            for (int r = 0; r < 8; r++) { cmac(s0r[r], s0i[r], xr[r], xi[r], w2.s1r, w3.s1i); cmac(s1r[r], s1i[r], xr[r], xi[r], w3.s2r, w2.s2i); }
#pragma unroll
            for (int r = 0; r < 8; r++) { xr[r] *= 1e-3; xi[r] *= 1e-3; }
        }
    }
    double s = 0;
    for (int m = 0; m < 8; m++) s += xr[m] + xi[m] + s0r[m] + s0i[m] + s1r[m] + s1i[m];
    if (s == 1.2345) out[0] = s;
}
template <int MODE, int MINB> void run(const char *name, int fp64_per_iter)
{
    double *d; cudaMalloc(&d, 8);
    Tw h[72]; host_twiddles(h, h + 8);
    Tw *dtw; cudaMalloc(&dtw, sizeof(h)); cudaMemcpy(dtw, h, sizeof(h), cudaMemcpyHostToDevice);
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); probe<MODE, MINB><<<sms * MINB, 64>>>(d, dtw, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double cyc_sm = best * 1e-3 * clk * 1e3 / iters;                 // SM cycles per iteration of all resident warps
    const double per = cyc_sm / (fp64_per_iter * (2.0 * MINB / 4.0));       // warps per SMSP = 2*MINB/4
    printf("%-34s CTAs/SM=%d  %.2f cycles per FP64 instruction per SMSP\n", name, MINB, per);
    cudaFree(d); cudaFree(dtw);
}
int main()
{
    run<0, 4>("3 x pass_fwd (216 DFMA)", 216);   run<0, 2>("3 x pass_fwd (216 DFMA)", 216);
    run<1, 4>("3 x pass_inv (288 FP64)", 288);
    run<2, 4>("MAC (64 DFMA)", 64);
    run<3, 4>("fwd + MAC + scale (296)", 296);   run<3, 6>("fwd + MAC + scale (296)", 296);
    return 0;
}
