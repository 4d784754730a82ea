// Developer probe: DFMA issue rate vs number of distinct 64-bit source operands (register-file bandwidth).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void probe(double *out, int iters, double m0, double c0)
{
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; i++) { a[i] = 1.0 + threadIdx.x * 1e-9 + i; b[i] = 1.0 + 1e-9 * i + 1e-12 * threadIdx.x; c[i] = 1e-9 * (i + 1); }
    double m = m0, k = c0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (MODE == 0) a[u] = fma(a[u], m, k);          // 1 varying + 2 loop constants
            if (MODE == 1) a[u] = fma(b[u], c[u], a[u]);    // 3 distinct varying operands
            if (MODE == 2) a[u] = fma(b[u], m, a[u]);       // 2 distinct + 1 constant
            if (MODE == 3) a[u] = fma(b[u], b[u], a[u]);    // 2 distinct (b twice)
            if (MODE == 4) a[u] = a[u] + b[u];              // DADD 2 operands
            if (MODE == 5) a[u] = fma(b[(u + 1) & 7], c[u], a[u]);  // 3 distinct, permuted
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + b[i] + c[i];
    if (s == 1.2345) out[0] = s;
}
template <int MODE> void run(const char *name, int warps)
{
    double *d; cudaMalloc(&d, 8);
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); probe<MODE><<<sms, 32 * warps>>>(d, iters, 1.0000001, 1e-7); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * clk * 1e3 / iters / 8.0 / (warps / 4.0);
    printf("%-40s warps/SM=%2d  %.2f cycles per DFMA per SMSP\n", name, warps, cyc);
}
int main()
{
    for (int w : {8, 16}) {
        if (w == 8) { run<0>("fma(a,m,k) 1 var", 8); run<1>("fma(b,c,a) 3 distinct", 8); run<2>("fma(b,m,a) 2 distinct+const", 8); run<3>("fma(b,b,a)", 8); run<4>("dadd(a,b)", 8); run<5>("fma(b',c,a) 3 distinct permuted", 8); }
        else { run<0>("fma(a,m,k) 1 var", 16); run<1>("fma(b,c,a) 3 distinct", 16); run<2>("fma(b,m,a) 2 distinct+const", 16); run<4>("dadd(a,b)", 16); }
    }
    return 0;
}
