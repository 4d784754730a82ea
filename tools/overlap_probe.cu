// Developer probe: do FP64 instructions and other instruction classes of DIFFERENT warps of one scheduler overlap?
// 12 warps per SM (3 per sub-partition), each running bursts: RF x 8 DFMA, then RA x 8 LOP3, RM x 8 IMAD, RL x 8 LDS —
// the shape of the persistent blind rotation (FP64-dense butterflies, then integer / shared-memory phases).  Warps start
// the burst sequence at different points, so at any time the three warps of a scheduler are in different phases.
// Reports cycles per loop iteration and sub-partition against the sum and the maximum of the single-class costs.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/overlap_probe tools/overlap_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int RF, int RA, int RM, int RL>
__global__ void __launch_bounds__(384, 1) probe(double *out, int iters, int stagger)
{
    __shared__ int sh[1024 + 64];
    double a[8]; int b[8], c[8], d[8];
    for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; b[i] = threadIdx.x + i; c[i] = threadIdx.x * 3 + i; d[i] = i; }
    for (int i = threadIdx.x; i < 1024 + 64; i += blockDim.x) sh[i] = i;
    __syncthreads();
    const double m = 1.0000001, k = 1e-7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int phase = stagger ? (warp >> 2) : 0;   // the 3 warps of a scheduler start in different bursts
    for (int it = 0; it < iters; it++) {
        for (int ph = 0; ph < 4; ph++) {
            const int which = (ph + phase) & 3;
            if (which == 0) {
#pragma unroll 4
                for (int r = 0; r < RF; r++) {
#pragma unroll
                    for (int u = 0; u < 8; u++) a[u] = fma(a[u], m, k);
                }
            } else if (which == 1) {
#pragma unroll 4
                for (int r = 0; r < RA; r++) {
#pragma unroll
                    for (int u = 0; u < 8; u++) b[u] = (b[u] ^ (b[u] >> 3)) & (0x7fffffff ^ it);   // LOP3 + SHF: two half-rate ops
                }
            } else if (which == 2) {
#pragma unroll 4
                for (int r = 0; r < RM; r++) {
#pragma unroll
                    for (int u = 0; u < 8; u++) c[u] = c[u] * 3 + it;                              // IMAD
                }
            } else {
#pragma unroll 4
                for (int r = 0; r < RL; r++) {
#pragma unroll
                    for (int u = 0; u < 8; u++) d[u] += sh[lane + 32 * u + (r & 63)];                  // LDS (independent addresses) + one add
                }
            }
        }
    }
    double s = 0; int t = 0;
    for (int i = 0; i < 8; i++) { s += a[i]; t += b[i] + c[i] + d[i]; }
    if (s == 1.2345 || t == 12345) out[0] = s + t;
}

template <int RF, int RA, int RM, int RL> double run(const char *name, int stagger)
{
    double *dptr; cudaMalloc(&dptr, 8);
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 2000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); probe<RF, RA, RM, RL><<<sms, 384>>>(dptr, iters, stagger); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * clk * 1e3 / iters / 3.0;   // cycles per warp-iteration and sub-partition (3 warps each)
    printf("%-44s stagger=%d  %8.1f cycles per warp-iteration per sub-partition  (%s)\n", name, stagger, cyc, cudaGetErrorString(cudaGetLastError()));
    cudaFree(dptr);
    return cyc;
}

int main()
{
    // 64 x 8 = 512 DFMA per iteration; 32 x 8 x 2 = 512 ALU ops; 32 x 8 = 256 IMAD; 32 x 8 = 256 LDS (+ 256 adds)
    const double f = run<64, 0, 0, 0>("512 DFMA", 0);
    const double a = run<0, 32, 0, 0>("256 x (LOP3 + SHF)", 0);
    const double m = run<0, 0, 32, 0>("256 IMAD", 0);
    const double l = run<0, 0, 0, 32>("256 x (LDS + add)", 0);
    const double fa = run<64, 32, 0, 0>("512 DFMA | 256 x (LOP3 + SHF), staggered", 1);
    const double fm = run<64, 0, 32, 0>("512 DFMA | 256 IMAD, staggered", 1);
    const double fl = run<64, 0, 0, 32>("512 DFMA | 256 x (LDS + add), staggered", 1);
    printf("DFMA+ALU: sum %.1f measured %.1f | DFMA+IMAD: sum %.1f measured %.1f | DFMA+LDS: sum %.1f measured %.1f\n", f + a, fa, f + m, fm, f + l, fl);
    const double all0 = run<64, 32, 32, 32>("all four bursts, warps in lockstep", 0);
    const double all1 = run<64, 32, 32, 32>("all four bursts, warps staggered", 1);
    printf("sum of the parts %.1f, max %.1f; lockstep %.1f, staggered %.1f\n", f + a + m + l, f, all0, all1);
    return 0;
}
