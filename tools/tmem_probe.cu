// Developer probe: can tensor memory serve as per-thread scratch for FP64 accumulators?  Measures the cost of
// tcgen05.ld / tcgen05.st (32x32b.x32: 32 columns x 32 lanes x 4 B = 4 KB per warp instruction) with 8 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(128, 2) probe(uint32_t *out, int iters)
{
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = threadIdx.x + i;
#define ST32(col) asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" \
        ::"r"(taddr + (col)), "r"(v[0]),"r"(v[1]),"r"(v[2]),"r"(v[3]),"r"(v[4]),"r"(v[5]),"r"(v[6]),"r"(v[7]),"r"(v[8]),"r"(v[9]),"r"(v[10]),"r"(v[11]),"r"(v[12]),"r"(v[13]),"r"(v[14]),"r"(v[15]), \
          "r"(v[16]),"r"(v[17]),"r"(v[18]),"r"(v[19]),"r"(v[20]),"r"(v[21]),"r"(v[22]),"r"(v[23]),"r"(v[24]),"r"(v[25]),"r"(v[26]),"r"(v[27]),"r"(v[28]),"r"(v[29]),"r"(v[30]),"r"(v[31]) : "memory")
#define LD32(col) asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];" \
        : "=r"(v[0]),"=r"(v[1]),"=r"(v[2]),"=r"(v[3]),"=r"(v[4]),"=r"(v[5]),"=r"(v[6]),"=r"(v[7]),"=r"(v[8]),"=r"(v[9]),"=r"(v[10]),"=r"(v[11]),"=r"(v[12]),"=r"(v[13]),"=r"(v[14]),"=r"(v[15]), \
          "=r"(v[16]),"=r"(v[17]),"=r"(v[18]),"=r"(v[19]),"=r"(v[20]),"=r"(v[21]),"=r"(v[22]),"=r"(v[23]),"=r"(v[24]),"=r"(v[25]),"=r"(v[26]),"=r"(v[27]),"=r"(v[28]),"=r"(v[29]),"=r"(v[30]),"=r"(v[31]) : "r"(taddr + (col)) : "memory")
    ST32(0); ST32(32); ST32(64); ST32(96);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    for (int it = 0; it < iters; it++) {
        const uint32_t col = (uint32_t)(it & 3) * 32;
        if (MODE == 0 || MODE == 2) { LD32(col); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] += 1;
        if (MODE == 1 || MODE == 2) { ST32(col); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase));
}
template <int MODE> void run(const char *name)
{
    int sms, clk; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0); cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t *d; cudaMalloc(&d, sms * 2 * 128 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); probe<MODE><<<sms * 2, 128>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double cyc = best * 1e-3 * clk * 1e3 / iters;          // SM cycles per iteration of all 8 resident warps
    const double bytes = 8.0 * 4096 * (MODE == 2 ? 2 : 1);
    printf("%-22s %.1f cycles per iteration of 8 warps -> %.0f B/clk/SM  (%s)\n", name, cyc, bytes / cyc, cudaGetErrorString(err));
    cudaFree(d);
}
int main() { run<0>("tcgen05.ld x32"); run<1>("tcgen05.st x32"); run<2>("ld + add + st"); return 0; }
