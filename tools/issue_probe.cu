// Developer probe: does a half-rate FP64 instruction block the scheduler's dispatch port for 2 cycles
// (so integer/LSU instructions cannot be issued in its shadow)?  Times DFMA-only, INT-only and mixed loops.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NI, int NS>
__global__ void probe(double *out, int *iout, int iters)
{
    __shared__ double sh[1024];
    double a[8]; int b[8];
    for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; b[i] = threadIdx.x + i; }
    sh[threadIdx.x] = threadIdx.x;
    __syncthreads();
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (u < NF) a[u] = fma(a[u], m, c);
            if (u < NI) b[u] = (b[u] ^ (b[u] >> 3)) + it;
            if (u < NS) a[u] += sh[(threadIdx.x + u * 32 + it) & 1023];
        }
    }
    double s = 0; int t = 0;
    for (int i = 0; i < 8; i++) { s += a[i]; t += b[i]; }
    if (s == 1.2345) out[0] = s;
    if (t == 12345) iout[0] = t;
}
template <int NF, int NI, int NS>
void run(const char *name, int warps_per_sm)
{
    double *d; int *di; cudaMalloc(&d, 8); cudaMalloc(&di, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000, threads = 32 * warps_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        probe<NF, NI, NS><<<sms, threads>>>(d, di, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * clk * 1e3 / iters;   // cycles per loop iteration (per SM, all warps)
    printf("%-28s warps/SM=%2d  %.1f cycles/iter  (per SMSP: %.2f cycles per warp-iteration-instruction group)\n", name, warps_per_sm, cyc, cyc / (warps_per_sm / 4.0));
}
int main()
{
    for (int w : {4, 8, 16}) {
        if (w == 4) { run<8,0,0>("8 DFMA", 4); run<0,8,0>("8 INT(3 ops)", 4); run<8,8,0>("8 DFMA + 8 INT", 4); run<8,0,8>("8 DFMA + 8 LDS+DADD", 4); run<0,0,8>("8 LDS+DADD", 4); }
        if (w == 8) { run<8,0,0>("8 DFMA", 8); run<0,8,0>("8 INT(3 ops)", 8); run<8,8,0>("8 DFMA + 8 INT", 8); run<8,0,8>("8 DFMA + 8 LDS+DADD", 8); run<0,0,8>("8 LDS+DADD", 8); }
        if (w == 16) { run<8,0,0>("8 DFMA", 16); run<0,8,0>("8 INT(3 ops)", 16); run<8,8,0>("8 DFMA + 8 INT", 16); }
    }
    return 0;
}
