// Developer probe: what would an exact integer NTT cost next to the FP64 transform?  (north_star: "negacyclic FFT (or an
// exact integer NTT, whichever ncu shows is faster)").
//
// Measures the issue cost, per SM sub-partition, of one radix-2 butterfly (a, b) -> (a + w b, a - w b)
//   * over the Goldilocks prime p = 2^64 - 2^32 + 1 (one 64-bit residue holds the 52-bit products of this path exactly),
//   * over a 31-bit prime with Shoup multiplication (two such residues + CRT would be needed),
// next to the FP64 complex butterfly of br_core.h (6 FMAs), all on register-resident data with 8 independent butterflies
// in flight per thread, 8 / 12 warps per SM.  A 1024-point negacyclic NTT has 512 x 10 butterflies; the FP64 transform
// works on 512 complex points and has 256 x 9.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ntt_probe tools/ntt_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t gl_reduce128(uint64_t lo, uint64_t hi)
{
    // x = hi * 2^64 + lo mod p, with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p)
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & 0xffffffffull;
    uint64_t t = lo - hi_hi;
    if (lo < hi_hi) t -= 0xffffffffull;                 // borrow: subtract 2^32 - 1 once more (i.e. add p)
    const uint64_t m = hi_lo * 0xffffffffull;            // hi_lo * (2^32 - 1)
    uint64_t r = t + m;
    if (r < m) r += 0xffffffffull;                       // carry: 2^64 = 2^32 - 1
    return r;
}
__device__ __forceinline__ uint64_t gl_mul(uint64_t a, uint64_t b) { return gl_reduce128(a * b, __umul64hi(a, b)); }
__device__ __forceinline__ uint64_t gl_add(uint64_t a, uint64_t b) { uint64_t r = a + b; if (r < a) r += 0xffffffffull; return r; }
__device__ __forceinline__ uint64_t gl_sub(uint64_t a, uint64_t b) { uint64_t r = a - b; if (a < b) r -= 0xffffffffull; return r; }

// 31-bit prime, Shoup: w' = floor(w 2^32 / p) precomputed; a + w b with lazy reduction to [0, 2p)
__device__ __forceinline__ uint32_t shoup_mul(uint32_t b, uint32_t w, uint32_t wp, uint32_t p)
{
    const uint32_t q = __umulhi(b, wp);
    return b * w - q * p;                                // in [0, 2p)
}

template <int MODE>
__global__ void __launch_bounds__(128) probe(uint64_t *out, int iters, uint64_t seed)
{
    constexpr int B = 8;
    if (MODE == 0) {            // Goldilocks
        uint64_t a[B], b[B];
        const uint64_t w = seed | 3;
#pragma unroll
        for (int i = 0; i < B; i++) { a[i] = seed + threadIdx.x * 977 + i; b[i] = seed * 31 + threadIdx.x + 7 * i; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < B; i++) {
                const uint64_t t = gl_mul(b[i], w);
                b[i] = gl_sub(a[i], t);
                a[i] = gl_add(a[i], t);
            }
        }
        uint64_t s = 0;
#pragma unroll
        for (int i = 0; i < B; i++) s ^= a[i] ^ b[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 1) {     // 31-bit prime, Shoup, lazy [0, 4p)
        uint32_t a[B], b[B];
        const uint32_t p = 2013265921u, w = (uint32_t)seed % p, wp = (uint32_t)(((uint64_t)w << 32) / p);
#pragma unroll
        for (int i = 0; i < B; i++) { a[i] = (uint32_t)(seed + threadIdx.x * 977 + i) % p; b[i] = (uint32_t)(seed * 31 + threadIdx.x + 7 * i) % p; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < B; i++) {
                uint32_t x = a[i]; if (x >= 2 * p) x -= 2 * p;
                const uint32_t t = shoup_mul(b[i], w, wp, p);
                a[i] = x + t;
                b[i] = x - t + 2 * p;
            }
        }
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < B; i++) s ^= a[i] ^ b[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {                    // FP64 complex butterfly (br_core.h bf): 6 FMAs
        double ar[B], ai[B], br[B], bi[B];
        const double zr = 0.98078528040323044913, zi = 0.19509032201612826785;
#pragma unroll
        for (int i = 0; i < B; i++) { ar[i] = 1.0 + 1e-3 * (threadIdx.x + i); ai[i] = 0.5 + 1e-3 * i; br[i] = 0.25 + 1e-3 * threadIdx.x; bi[i] = 0.125 * i; }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < B; i++) {
                double tr = fma(zr, br[i], ar[i]); tr = fma(-zi, bi[i], tr);
                double ti = fma(zr, bi[i], ai[i]); ti = fma(zi, br[i], ti);
                br[i] = fma(2.0, ar[i], -tr) * 0.5; bi[i] = fma(2.0, ai[i], -ti) * 0.5;   // the 0.5 keeps the values bounded
                ar[i] = tr * 0.5; ai[i] = ti * 0.5;
            }
        }
        double s = 0;
#pragma unroll
        for (int i = 0; i < B; i++) s += ar[i] + ai[i] + br[i] + bi[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (uint64_t)__double_as_longlong(s);
    }
}

template <int MODE> void run(const char *name, int ctas_per_sm, double ops_note)
{
    int sms, clk;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint64_t *d;
    cudaMalloc(&d, (size_t)sms * ctas_per_sm * 128 * 8);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 4; r++) {
        cudaEventRecord(e0);
        probe<MODE><<<sms * ctas_per_sm, 128>>>(d, iters, 0x9E3779B97F4A7C15ull);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms < best) best = ms;
    }
    const double warps_per_smsp = ctas_per_sm * 4 / 4.0;
    const double cyc_per_bfly = best * 1e-3 * clk * 1e3 / ((double)iters * 8 * warps_per_smsp);   // SMSP cycles per warp-butterfly
    printf("%-44s warps/SM=%2d  %6.2f cycles per butterfly and sub-partition  %s (%s)\n", name, ctas_per_sm * 4, cyc_per_bfly,
           ops_note > 0 ? "" : "", cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main()
{
    for (int c : {2, 3}) {
        run<2>("FP64 complex butterfly (6 FMA + 4 scaling mul)", c, 0);
        run<0>("Goldilocks 64-bit butterfly (mul + add + sub)", c, 0);
        run<1>("31-bit Shoup butterfly (one of two CRT residues)", c, 0);
    }
    printf("a 1024-point negacyclic NTT has 5120 butterflies (x2 residues for the 31-bit form); the FP64 transform has 2304\n");
    return 0;
}
