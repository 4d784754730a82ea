// Developer probe: shared-memory instruction cost for 64/128-bit loads and stores with the kernel's access patterns.
#include <cstdio>
#include <cuda_runtime.h>
struct cd { double x, y; };
template <int MODE>
__global__ void probe(double *out, int iters)
{
    __shared__ __align__(16) cd buf[2304];   // 4 x 576
    const int tid = threadIdx.x & 63, grp = threadIdx.x >> 6;
    cd *b = buf + (grp & 3) * 576;
    for (int i = threadIdx.x; i < 2304; i += blockDim.x) { buf[i].x = i; buf[i].y = -i; }
    __syncthreads();
    double ax = 0, ay = 0;
    cd v[8];
    for (int m = 0; m < 8; m++) { v[m].x = tid + m; v[m].y = tid - m; }
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // 8 x LDS.128, stride-9 pattern (pass 3 read)
#pragma unroll
            for (int m = 0; m < 8; m++) { cd t = b[9 * tid + m]; ax += t.x; ay += t.y; }
        } else if (MODE == 1) {     // 8 x STS.128, contiguous per lane (pass 1 write)
#pragma unroll
            for (int m = 0; m < 8; m++) { v[m].x += 1.0; b[72 * m + tid + (tid >> 3)] = v[m]; }
        } else if (MODE == 2) {     // 8 x LDS.64 pairs (same bytes as MODE 0 but 64-bit accesses)
            const double *d = reinterpret_cast<const double *>(b);
#pragma unroll
            for (int m = 0; m < 8; m++) { ax += d[2 * (9 * tid + m)]; ay += d[2 * (9 * tid + m) + 1]; }
        } else if (MODE == 3) {     // 8 x LDS.128 contiguous (pass 2 read pattern: base + 9m)
            const int base = 72 * (tid >> 3) + (tid & 7);
#pragma unroll
            for (int m = 0; m < 8; m++) { cd t = b[base + 9 * m]; ax += t.x; ay += t.y; }
        } else if (MODE == 4) {     // 8 STS.128 + 8 LDS.128 (one exchange)
            const int base = 72 * (tid >> 3) + (tid & 7);
#pragma unroll
            for (int m = 0; m < 8; m++) { v[m].x += ax; b[base + 9 * m] = v[m]; }
            __syncwarp();
#pragma unroll
            for (int m = 0; m < 8; m++) { cd t = b[9 * tid + m]; ax += t.x; ay += t.y; }
        }
    }
    if (ax + ay + v[3].x == 1.2345) out[0] = ax;
}
template <int MODE> void run(const char *name, int warps, int insts)
{
    double *d; cudaMalloc(&d, 8);
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); probe<MODE><<<sms, 32 * warps>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double cyc = best * 1e-3 * clk * 1e3 / iters / (double)(insts * warps);
    printf("%-44s warps/SM=%2d  %.2f SM-cycles per warp-instruction\n", name, warps, cyc);
}
int main()
{
    run<0>("LDS.128 stride 9 (pass-3 read)", 8, 8);   run<0>("LDS.128 stride 9 (pass-3 read)", 16, 8);
    run<3>("LDS.128 8-lane contiguous (pass-2 read)", 8, 8);
    run<1>("STS.128 contiguous (pass-1 write)", 8, 8); run<1>("STS.128 contiguous (pass-1 write)", 16, 8);
    run<2>("2 x LDS.64 stride 9", 8, 16);
    run<4>("8 STS.128 + 8 LDS.128", 8, 16);           run<4>("8 STS.128 + 8 LDS.128", 16, 16);
    return 0;
}
