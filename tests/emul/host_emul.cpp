/*
 * tests/emul/host_emul.cpp — TEST-ONLY host emulation of the blind-rotation kernel.
 *
 * Runs the exact per-thread arithmetic of ie-ache_b200/csrc/br_core.h (the same __host__
 * __device__ functions the CUDA kernel calls) for the 64 threads of one group, phase by phase,
 * with each barrier of the kernel becoming the end of a loop over threads.  It exists because
 * the build container has no GPU: it lets the CPU test suite check the transform index math,
 * the exchange-buffer layout, the BK layout and the decomposition against the oracle before a
 * GPU run is spent.  It is NOT part of the product library and is never a fallback.
 */
#include "../../ie-ache_b200/csrc/br_core.h"
#include "../../ie-ache_b200/csrc/br_warp.h"

#include <cstdlib>
#include <cstring>
#include <vector>

using namespace ieache;

namespace {
struct Regs { double xr[8], xi[8]; };

struct Emul {
    Tw w1, w2[8], w3[64];
    Emul() { w1 = tw_pass1(); host_twiddles(w2, w3); }

    void fwd(Regs *t, cd *buf) const
    {
        for (int tid = 0; tid < 64; tid++) { pass_fwd(t[tid].xr, t[tid].xi, w1); st_pass1(buf, tid, t[tid].xr, t[tid].xi); }
        for (int tid = 0; tid < 64; tid++) ld_pass2(buf, tid, t[tid].xr, t[tid].xi);
        for (int tid = 0; tid < 64; tid++) { pass_fwd(t[tid].xr, t[tid].xi, w2[tid >> 3]); st_pass2(buf, tid, t[tid].xr, t[tid].xi); }
        for (int tid = 0; tid < 64; tid++) { ld_pass3(buf, tid, t[tid].xr, t[tid].xi); pass_fwd(t[tid].xr, t[tid].xi, w3[tid]); }
    }
    void inv(Regs *t, cd *buf) const
    {
        for (int tid = 0; tid < 64; tid++) { pass_inv(t[tid].xr, t[tid].xi, w3[tid]); st_ipass3(buf, tid, t[tid].xr, t[tid].xi); }
        for (int tid = 0; tid < 64; tid++) ld_ipass2(buf, tid, t[tid].xr, t[tid].xi);
        for (int tid = 0; tid < 64; tid++) { pass_inv(t[tid].xr, t[tid].xi, w2[tid >> 3]); st_ipass2(buf, tid, t[tid].xr, t[tid].xi); }
        for (int tid = 0; tid < 64; tid++) { ld_ipass1(buf, tid, t[tid].xr, t[tid].xi); pass_inv(t[tid].xr, t[tid].xi, w1); }
    }
};
const Emul &emul() { static Emul e; return e; }
} // namespace

static const TwInvP1 inv1 = IE_TW_INV_P1_INIT;
extern "C" {

/* forward transform of one int32 polynomial into the device BK layout [r][t3] (re,im), scaled */
void emul_poly_fft(const int32_t *coef, double *out /*512*2*/, double scale)
{
    const Emul &e = emul();
    Regs t[64];
    cd buf[kBufElems];
    for (int tid = 0; tid < 64; tid++)
        for (int m = 0; m < 8; m++) { t[tid].xr[m] = (double)coef[tid + 64 * m]; t[tid].xi[m] = (double)coef[tid + 64 * m + 512]; }
    e.fwd(t, buf);
    for (int tid = 0; tid < 64; tid++)
        for (int r = 0; r < 8; r++) { out[2 * (r * 64 + tid)] = t[tid].xr[r] * scale; out[2 * (r * 64 + tid) + 1] = t[tid].xi[r] * scale; }
}

/* inverse of emul_poly_fft (x512 unnormalised, like the kernel) */
void emul_poly_ifft(const double *in /*512*2*/, double *coef_out /*1024*/)
{
    const Emul &e = emul();
    Regs t[64];
    cd buf[kBufElems];
    for (int tid = 0; tid < 64; tid++)
        for (int r = 0; r < 8; r++) { t[tid].xr[r] = in[2 * (r * 64 + tid)]; t[tid].xi[r] = in[2 * (r * 64 + tid) + 1]; }
    e.inv(t, buf);
    for (int tid = 0; tid < 64; tid++)
        for (int m = 0; m < 8; m++) { coef_out[tid + 64 * m] = t[tid].xr[m]; coef_out[tid + 64 * m + 512] = t[tid].xi[m]; }
}

/* evaluation index K (root psi^(4K+1)) held at slot [r][t3] */
int emul_slot_to_K(int r, int t3) { return (t3 >> 3) + 8 * (t3 & 7) + 64 * brev3(r); }

/* whole blind rotation + sample extract of one gate, as the kernel does it.
 * bk_coef: [n][2l][2][1024] int32 ; x: n+1 words (already pre-combined) ; ext: 1025 words */
void emul_blind_rotate(int n, int l, int Bgbit, int32_t mu, const int32_t *bk_coef, const int32_t *x, int32_t *ext)
{
    const Emul &e = emul();
    const int kpl = 2 * l;
    /* key load */
    std::vector<double> bkfft((size_t)n * kpl * 2 * 1024);
    for (long q = 0; q < (long)n * kpl * 2; q++) emul_poly_fft(bk_coef + (size_t)q * kN, &bkfft[(size_t)q * 1024], 1.0 / 512.0);

    std::vector<int32_t> acc(2 * kN);
    std::vector<int> abar(n + 1);
    for (int i = 0; i <= n; i++) abar[i] = modswitch_2N(x[i]);
    {
        const int a = (2 * kN - abar[n]) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = 0; j < kN; j++) { acc[j] = 0; acc[kN + j] = ((j < ar) != flip) ? -mu : mu; }
    }
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
    for (int i = 1; i <= l; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);

    Regs t[64];
    cd bufA[kBufElems], bufB[kBufElems];
    int toggle = 0;
    std::vector<double> sr(64 * 2 * 8), si(64 * 2 * 8);
    for (int i = 0; i < n; i++) {
        const int a = abar[i];
        if (a == 0) continue;
        std::fill(sr.begin(), sr.end(), 0.0); std::fill(si.begin(), si.end(), 0.0);
        for (int q = 0; q < 2; q++) {
            int32_t c[64][16];
            for (int tid = 0; tid < 64; tid++) rot_minus_one(&acc[q * kN], tid, a, c[tid]);
            for (int pp = 0; pp < l; pp++) {
                const int shift = 32 - (pp + 1) * Bgbit;
                {   /* as the default kernel: pass 1 straight from the integer digits, then passes 2 and 3 */
                    cd *buf = toggle ? bufB : bufA; toggle ^= 1;
                    for (int tid = 0; tid < 64; tid++) {
                        int32_t dr[8], di[8];
                        for (int m = 0; m < 8; m++) {
                            dr[m] = digit_i32(c[tid][m], offset, shift, maskBg, halfBg);
                            di[m] = digit_i32(c[tid][8 + m], offset, shift, maskBg, halfBg);
                        }
                        pass1_fwd_from_digits(dr, di, t[tid].xr, t[tid].xi, e.w1);
                        st_pass1(buf, tid, t[tid].xr, t[tid].xi);
                    }
                    for (int tid = 0; tid < 64; tid++) ld_pass2(buf, tid, t[tid].xr, t[tid].xi);
                    for (int tid = 0; tid < 64; tid++) { pass_fwd(t[tid].xr, t[tid].xi, e.w2[tid >> 3]); st_pass2(buf, tid, t[tid].xr, t[tid].xi); }
                    for (int tid = 0; tid < 64; tid++) { ld_pass3(buf, tid, t[tid].xr, t[tid].xi); pass_fwd(t[tid].xr, t[tid].xi, e.w3[tid]); }
                }
                const double *bk_r = &bkfft[(((size_t)i * kpl + q * l + pp) * 2) * 1024];
                for (int tid = 0; tid < 64; tid++)
                    for (int j = 0; j < 2; j++)
                        for (int r = 0; r < 8; r++) {
                            const double *b = bk_r + (size_t)j * 1024 + 2 * (r * 64 + tid);
                            cmac(sr[(tid * 2 + j) * 8 + r], si[(tid * 2 + j) * 8 + r], t[tid].xr[r], t[tid].xi[r], b[0], b[1]);
                        }
            }
        }
        for (int j = 0; j < 2; j++) {
            for (int tid = 0; tid < 64; tid++)
                for (int r = 0; r < 8; r++) { t[tid].xr[r] = sr[(tid * 2 + j) * 8 + r]; t[tid].xi[r] = si[(tid * 2 + j) * 8 + r]; }
            e.inv(t, toggle ? bufB : bufA); toggle ^= 1;
            for (int tid = 0; tid < 64; tid++)
                for (int m = 0; m < 8; m++) {
                    acc[j * kN + tid + 64 * m] += round_to_torus(t[tid].xr[m]);
                    acc[j * kN + tid + 64 * m + 512] += round_to_torus(t[tid].xi[m]);
                }
        }
    }
    for (int j = 0; j < kN; j++) ext[j] = (j == 0) ? acc[0] : -acc[kN - j];
    ext[kN] = acc[kN];
}

} // extern "C"

/* ------------------------------------------------------------------ warp-per-gate layout (br_warp.h) */
namespace {
struct Regs16 { double xr[16], xi[16]; };
struct EmulW {
    Tw16 w1, w2[16];
    FinTw fin[32];
    EmulW() { w1 = tw16_pass1(); host_twiddles_warp(w2, fin); }
    void fwd(Regs16 *t, cd *buf) const
    {
        for (int l = 0; l < 32; l++) { pass16_fwd(t[l].xr, t[l].xi, w1); st16_pass1(buf, l, t[l].xr, t[l].xi); }
        for (int l = 0; l < 32; l++) { ld16_pass2(buf, l, t[l].xr, t[l].xi); pass16_fwd(t[l].xr, t[l].xi, w2[l & 15]); }
        double sr[32][8], si[32][8];
        for (int l = 0; l < 32; l++) fin_fwd_send(t[l].xr, t[l].xi, l >> 4, sr[l], si[l]);
        for (int l = 0; l < 32; l++) fin_fwd_apply(t[l].xr, t[l].xi, l >> 4, sr[l ^ 16], si[l ^ 16], fin[l].zr, fin[l].zi); /* shfl_xor 16 */
    }
    void inv(Regs16 *t, cd *buf) const
    {
        double sr[32][8], si[32][8];
        for (int l = 0; l < 32; l++) { fin_inv_local(t[l].xr, t[l].xi, fin[l].zr, fin[l].zi); fin_inv_send(t[l].xr, t[l].xi, l >> 4, sr[l], si[l]); }
        for (int l = 0; l < 32; l++) fin_inv_place(t[l].xr, t[l].xi, l >> 4, sr[l ^ 16], si[l ^ 16]);
        for (int l = 0; l < 32; l++) { pass16_inv(t[l].xr, t[l].xi, w2[l & 15]); st16_ipass2(buf, l, t[l].xr, t[l].xi); }
        for (int l = 0; l < 32; l++) { ld16_ipass1(buf, l, t[l].xr, t[l].xi); pass16_inv(t[l].xr, t[l].xi, w1); }
    }
};
const EmulW &emulw() { static EmulW e; return e; }
} // namespace

extern "C" {

/* forward transform into the warp layout [p][lane] (re,im) */
void emul_warp_fft(const int32_t *coef, double *out /*512*2*/, double scale)
{
    const EmulW &e = emulw();
    Regs16 t[32];
    cd buf[kWarpBufElems];
    for (int l = 0; l < 32; l++)
        for (int m = 0; m < 16; m++) { t[l].xr[m] = (double)coef[l + 32 * m]; t[l].xi[m] = (double)coef[l + 32 * m + 512]; }
    e.fwd(t, buf);
    for (int l = 0; l < 32; l++)
        for (int p = 0; p < 16; p++) { out[2 * (p * 32 + l)] = t[l].xr[p] * scale; out[2 * (p * 32 + l) + 1] = t[l].xi[p] * scale; }
}
void emul_warp_ifft(const double *in /*512*2*/, double *coef_out /*1024*/)
{
    const EmulW &e = emulw();
    Regs16 t[32];
    cd buf[kWarpBufElems];
    for (int l = 0; l < 32; l++)
        for (int p = 0; p < 16; p++) { t[l].xr[p] = in[2 * (p * 32 + l)]; t[l].xi[p] = in[2 * (p * 32 + l) + 1]; }
    e.inv(t, buf);
    for (int l = 0; l < 32; l++)
        for (int m = 0; m < 16; m++) { coef_out[l + 32 * m] = t[l].xr[m]; coef_out[l + 32 * m + 512] = t[l].xi[m]; }
}
int emul_warp_slot_to_K(int p, int lane) { return warp_slot_to_K(p, lane); }

/* whole blind rotation + sample extract with the warp layout (BK relaid out through emul_slot_to_K / warp_slot_to_K,
 * exactly what the key-load kernel does on the device) */
void emul_warp_blind_rotate(int n, int l, int Bgbit, int32_t mu, const int32_t *bk_coef, const int32_t *x, int32_t *ext)
{
    const EmulW &e = emulw();
    const int kpl = 2 * l;
    std::vector<double> bkw((size_t)n * kpl * 2 * 1024);
    {
        /* old layout first, then the permutation old [r8][t64] -> new [p16][lane32] */
        int pos_of_K[512];
        for (int r = 0; r < 8; r++) for (int t3 = 0; t3 < 64; t3++) pos_of_K[emul_slot_to_K(r, t3)] = r * 64 + t3;
        std::vector<double> old(1024);
        for (long q = 0; q < (long)n * kpl * 2; q++) {
            emul_poly_fft(bk_coef + (size_t)q * kN, old.data(), 1.0 / 512.0);
            double *dst = &bkw[(size_t)q * 1024];
            for (int p = 0; p < 16; p++) for (int lane = 0; lane < 32; lane++) {
                const int src = pos_of_K[warp_slot_to_K(p, lane)];
                dst[2 * (p * 32 + lane)] = old[2 * src]; dst[2 * (p * 32 + lane) + 1] = old[2 * src + 1];
            }
        }
    }
    std::vector<int32_t> acc(2 * kN);
    std::vector<int> abar(n + 1);
    for (int i = 0; i <= n; i++) abar[i] = modswitch_2N(x[i]);
    {
        const int a = (2 * kN - abar[n]) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = 0; j < kN; j++) { acc[j] = 0; acc[kN + j] = ((j < ar) != flip) ? -mu : mu; }
    }
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
    for (int i = 1; i <= l; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    Regs16 t[32];
    cd buf[kWarpBufElems];
    std::vector<double> sr(32 * 2 * 16), si(32 * 2 * 16);
    for (int i = 0; i < n; i++) {
        const int a = abar[i];
        std::fill(sr.begin(), sr.end(), 0.0); std::fill(si.begin(), si.end(), 0.0);
        for (int q = 0; q < 2; q++) {
            int32_t c[32][32];
            for (int lane = 0; lane < 32; lane++) rot_minus_one32(&acc[q * kN], lane, a, c[lane]);
            for (int pp = 0; pp < l; pp++) {
                const int shift = 32 - (pp + 1) * Bgbit;
                for (int lane = 0; lane < 32; lane++)
                    for (int m = 0; m < 16; m++) {
                        t[lane].xr[m] = digit_f64(c[lane][m], offset, shift, maskBg, halfBg);
                        t[lane].xi[m] = digit_f64(c[lane][16 + m], offset, shift, maskBg, halfBg);
                    }
                e.fwd(t, buf);
                const double *bk_r = &bkw[(((size_t)i * kpl + q * l + pp) * 2) * 1024];
                for (int lane = 0; lane < 32; lane++)
                    for (int j = 0; j < 2; j++)
                        for (int p = 0; p < 16; p++) {
                            const double *b = bk_r + (size_t)j * 1024 + 2 * (p * 32 + lane);
                            cmac(sr[(lane * 2 + j) * 16 + p], si[(lane * 2 + j) * 16 + p], t[lane].xr[p], t[lane].xi[p], b[0], b[1]);
                        }
            }
        }
        for (int j = 0; j < 2; j++) {
            for (int lane = 0; lane < 32; lane++)
                for (int p = 0; p < 16; p++) { t[lane].xr[p] = sr[(lane * 2 + j) * 16 + p]; t[lane].xi[p] = si[(lane * 2 + j) * 16 + p]; }
            e.inv(t, buf);
            for (int lane = 0; lane < 32; lane++)
                for (int m = 0; m < 16; m++) {
                    acc[j * kN + lane + 32 * m] += round_to_torus(t[lane].xr[m]);
                    acc[j * kN + lane + 32 * m + 512] += round_to_torus(t[lane].xi[m]);
                }
        }
    }
    for (int j = 0; j < kN; j++) ext[j] = (j == 0) ? acc[0] : -acc[kN - j];
    ext[kN] = acc[kN];
}

} // extern "C"

/* ------------------------------------------------------------------ persistent 12-warp kernel (br_w12.cu)
 * Same layout as the warp kernel; differs in pass 1 (straight from the integer digits, pass16_fwd_from_digits) and in
 * the final-stage twiddle table (4 complex per lane, z[s + 1] = i z[s] for even s). */
extern "C" {
void emul_w12_blind_rotate(int n, int l, int Bgbit, int32_t mu, const int32_t *bk_coef, const int32_t *x, int32_t *ext)
{
    const EmulW &e = emulw();
    const int kpl = 2 * l;
    std::vector<double> bkw((size_t)n * kpl * 2 * 1024);
    {
        int pos_of_K[512];
        for (int r = 0; r < 8; r++) for (int t3 = 0; t3 < 64; t3++) pos_of_K[emul_slot_to_K(r, t3)] = r * 64 + t3;
        std::vector<double> old(1024);
        for (long q = 0; q < (long)n * kpl * 2; q++) {
            emul_poly_fft(bk_coef + (size_t)q * kN, old.data(), 1.0 / 512.0);
            double *dst = &bkw[(size_t)q * 1024];
            for (int p = 0; p < 16; p++) for (int lane = 0; lane < 32; lane++) {
                const int src = pos_of_K[warp_slot_to_K(p, lane)];
                dst[2 * (p * 32 + lane)] = old[2 * src]; dst[2 * (p * 32 + lane) + 1] = old[2 * src + 1];
            }
        }
    }
    /* the kernel's 4-entry table and the full one it stands for */
    FinTw fin4[32];
    for (int lane = 0; lane < 32; lane++)
        for (int h = 0; h < 4; h++) {
            fin4[lane].zr[2 * h] = e.fin[lane].zr[2 * h]; fin4[lane].zi[2 * h] = e.fin[lane].zi[2 * h];
            fin4[lane].zr[2 * h + 1] = -e.fin[lane].zi[2 * h]; fin4[lane].zi[2 * h + 1] = e.fin[lane].zr[2 * h];
        }
    std::vector<int32_t> acc(2 * kN);
    std::vector<int> abar(n + 1);
    for (int i = 0; i <= n; i++) abar[i] = modswitch_2N(x[i]);
    {
        const int a = (2 * kN - abar[n]) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = 0; j < kN; j++) { acc[j] = 0; acc[kN + j] = ((j < ar) != flip) ? -mu : mu; }
    }
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
    for (int i = 1; i <= l; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    Regs16 t[32];
    cd buf[kWarpBufElems];
    std::vector<double> sr(32 * 2 * 16), si(32 * 2 * 16);
    for (int i = 0; i < n; i++) {
        const int a = abar[i];
        if (a == 0) continue;
        std::fill(sr.begin(), sr.end(), 0.0); std::fill(si.begin(), si.end(), 0.0);
        for (int q = 0; q < 2; q++) {
            uint32_t cc[32][32];
            for (int lane = 0; lane < 32; lane++) {
                int32_t c[32];
                rot_minus_one32(&acc[q * kN], lane, a, c);
                for (int h = 0; h < 32; h++) cc[lane][h] = (uint32_t)c[h] + offset;
            }
            for (int pp = 0; pp < l; pp++) {
                const int shift = 32 - (pp + 1) * Bgbit;
                for (int lane = 0; lane < 32; lane++) { pass16_fwd_from_digits(cc[lane], shift, maskBg, halfBg, t[lane].xr, t[lane].xi, e.w1); st16_pass1(buf, lane, t[lane].xr, t[lane].xi); }
                for (int lane = 0; lane < 32; lane++) { ld16_pass2(buf, lane, t[lane].xr, t[lane].xi); pass16_fwd(t[lane].xr, t[lane].xi, e.w2[lane & 15]); }
                double ssr[32][8], ssi[32][8];
                for (int lane = 0; lane < 32; lane++) fin_fwd_send(t[lane].xr, t[lane].xi, lane >> 4, ssr[lane], ssi[lane]);
                for (int lane = 0; lane < 32; lane++) fin_fwd_apply(t[lane].xr, t[lane].xi, lane >> 4, ssr[lane ^ 16], ssi[lane ^ 16], fin4[lane].zr, fin4[lane].zi);
                const double *bk_r = &bkw[(((size_t)i * kpl + q * l + pp) * 2) * 1024];
                for (int lane = 0; lane < 32; lane++)
                    for (int j = 0; j < 2; j++)
                        for (int p = 0; p < 16; p++) {
                            const double *b = bk_r + (size_t)j * 1024 + 2 * (p * 32 + lane);
                            cmac(sr[(lane * 2 + j) * 16 + p], si[(lane * 2 + j) * 16 + p], t[lane].xr[p], t[lane].xi[p], b[0], b[1]);
                        }
            }
        }
        for (int j = 0; j < 2; j++) {
            for (int lane = 0; lane < 32; lane++)
                for (int p = 0; p < 16; p++) { t[lane].xr[p] = sr[(lane * 2 + j) * 16 + p]; t[lane].xi[p] = si[(lane * 2 + j) * 16 + p]; }
            double ssr[32][8], ssi[32][8];
            for (int lane = 0; lane < 32; lane++) { fin_inv_local(t[lane].xr, t[lane].xi, fin4[lane].zr, fin4[lane].zi); fin_inv_send(t[lane].xr, t[lane].xi, lane >> 4, ssr[lane], ssi[lane]); }
            for (int lane = 0; lane < 32; lane++) fin_inv_place(t[lane].xr, t[lane].xi, lane >> 4, ssr[lane ^ 16], ssi[lane ^ 16]);
            for (int lane = 0; lane < 32; lane++) { pass16_inv(t[lane].xr, t[lane].xi, e.w2[lane & 15]); st16_ipass2(buf, lane, t[lane].xr, t[lane].xi); }
            for (int lane = 0; lane < 32; lane++) { ld16_ipass1(buf, lane, t[lane].xr, t[lane].xi); pass16_inv(t[lane].xr, t[lane].xi, e.w1); }
            for (int lane = 0; lane < 32; lane++)
                for (int m = 0; m < 16; m++) {
                    acc[j * kN + lane + 32 * m] += round_to_torus(t[lane].xr[m]);
                    acc[j * kN + lane + 32 * m + 512] += round_to_torus(t[lane].xi[m]);
                }
        }
    }
    for (int j = 0; j < kN; j++) ext[j] = (j == 0) ? acc[0] : -acc[kN - j];
    ext[kN] = acc[kN];
}
} // extern "C"

/* ------------------------------------------------------------------ folded forward variant of the warp layout */
namespace {
struct EmulWF {
    Tw16 w1;
    Tw16g w2[32];
    FinTw fin[32];
    EmulWF() { w1 = tw16_pass1(); host_twiddles_warp_folded(w2, fin); }
    void fwd(Regs16 *t, cd *buf) const
    {
        for (int l = 0; l < 32; l++) { pass16_fwd(t[l].xr, t[l].xi, w1); st16_pass1(buf, l, t[l].xr, t[l].xi); }
        for (int l = 0; l < 32; l++) { ld16_pass2(buf, l, t[l].xr, t[l].xi); pass16_fwd_g(t[l].xr, t[l].xi, w2[l]); }
        double sr[32][8], si[32][8];
        for (int l = 0; l < 32; l++) for (int s = 0; s < 8; s++) { sr[l][s] = t[l].xr[8 + s]; si[l][s] = t[l].xi[8 + s]; }
        for (int l = 0; l < 32; l++) fin_fwd_apply_folded(t[l].xr, t[l].xi, sr[l ^ 16], si[l ^ 16], fin[l].zr, fin[l].zi);
    }
};
const EmulWF &emulwf() { static EmulWF e; return e; }
} // namespace

extern "C" {
/* forward transform of the folded variant, multiplied back by the key factor: must equal the plain warp layout */
void emul_warpf_fft_unfolded(const int32_t *coef, double *out /*512*2*/)
{
    const EmulWF &e = emulwf();
    Regs16 t[32];
    cd buf[kWarpBufElems];
    for (int l = 0; l < 32; l++)
        for (int m = 0; m < 16; m++) { t[l].xr[m] = (double)coef[l + 32 * m]; t[l].xi[m] = (double)coef[l + 32 * m + 512]; }
    e.fwd(t, buf);
    for (int l = 0; l < 32; l++)
        for (int p = 0; p < 16; p++) {
            double fr, fi;
            folded_bk_factor(e.fin[l], p, l, fr, fi);
            out[2 * (p * 32 + l)] = t[l].xr[p] * fr - t[l].xi[p] * fi;
            out[2 * (p * 32 + l) + 1] = t[l].xr[p] * fi + t[l].xi[p] * fr;
        }
}
} // extern "C"

/* ------------------------------------------------------------------ 12-warp kernel, select-free form (br_w12.cu)
 * folded forward AND folded inverse, 4-entry final-stage table with the uniform rule w[s + 1] = i w[s], plain key
 * values in the w12_slot_to_K order */
extern "C" {
int emul_w12_slot_to_K(int p, int lane) { return w12_slot_to_K(p, lane); }
void emul_w12f_blind_rotate(int n, int l, int Bgbit, int32_t mu, const int32_t *bk_coef, const int32_t *x, int32_t *ext)
{
    const EmulWF &e = emulwf();
    const int kpl = 2 * l;
    std::vector<double> bkw((size_t)n * kpl * 2 * 1024);
    {
        int pos_of_K[512];
        for (int r = 0; r < 8; r++) for (int t3 = 0; t3 < 64; t3++) pos_of_K[emul_slot_to_K(r, t3)] = r * 64 + t3;
        std::vector<double> old(1024);
        for (long q = 0; q < (long)n * kpl * 2; q++) {
            emul_poly_fft(bk_coef + (size_t)q * kN, old.data(), 1.0 / 512.0);
            double *dst = &bkw[(size_t)q * 1024];
            const int dig = (int)((q / 2) % l);                       /* q = ((i kpl + r) 2 + j), digit p = r mod l */
            const double sc = ldexp(1.0, -w12_field_shift(dig, Bgbit));   /* rows of digit p carry 2^-sp (pass16_fwd_from_fields) */
            for (int p = 0; p < 16; p++) for (int lane = 0; lane < 32; lane++) {
                const int src = pos_of_K[w12_slot_to_K(p, lane)];
                dst[2 * (p * 32 + lane)] = old[2 * src] * sc; dst[2 * (p * 32 + lane) + 1] = old[2 * src + 1] * sc;
            }
        }
    }
    std::vector<int32_t> acc(2 * kN);
    std::vector<int> abar(n + 1);
    for (int i = 0; i <= n; i++) abar[i] = modswitch_2N(x[i]);
    {
        const int a = (2 * kN - abar[n]) & (2 * kN - 1), ar = a & (kN - 1);
        const bool flip = a >= kN;
        for (int j = 0; j < kN; j++) { acc[j] = 0; acc[kN + j] = ((j < ar) != flip) ? -mu : mu; }
    }
    const uint32_t maskBg = (1u << Bgbit) - 1;
    const int32_t halfBg = 1 << (Bgbit - 1);
    uint32_t offset = 0;
    for (int i = 1; i <= l; i++) offset += (uint32_t)halfBg << (32 - i * Bgbit);
    Regs16 t[32];
    cd buf[kWarpBufElems];
    std::vector<double> sr(32 * 2 * 16), si(32 * 2 * 16);
    auto fin_z = [&](int lane, int h, double (&z)[8]) {
        fin4_tw(e.fin[lane].zr[4 * h], e.fin[lane].zi[4 * h], e.fin[lane].zr[4 * h + 2], e.fin[lane].zi[4 * h + 2], z);
    };
    for (int i = 0; i < n; i++) {
        const int a = abar[i];
        if (a == 0) continue;
        std::fill(sr.begin(), sr.end(), 0.0); std::fill(si.begin(), si.end(), 0.0);
        for (int q = 0; q < 2; q++) {
            uint32_t cc[32][32];
            for (int lane = 0; lane < 32; lane++) {
                int32_t c[32];
                rot_minus_one32(&acc[q * kN], lane, a, c);
                for (int h = 0; h < 32; h++) cc[lane][h] = ((uint32_t)c[h] + offset) >> 1;
            }
            for (int pp = 0; pp < l; pp++) {
                const int sp = w12_field_shift(pp, Bgbit);
                for (int lane = 0; lane < 32; lane++) { pass16_fwd_from_fields(cc[lane], maskBg << sp, (uint32_t)halfBg << sp, t[lane].xr, t[lane].xi, e.w1); st16_pass1(buf, lane, t[lane].xr, t[lane].xi); }
                for (int lane = 0; lane < 32; lane++) { ld16_pass2(buf, lane, t[lane].xr, t[lane].xi); pass16_fwd_g(t[lane].xr, t[lane].xi, e.w2[lane]); }
                double ssr[32][8], ssi[32][8];
                for (int lane = 0; lane < 32; lane++) for (int s8 = 0; s8 < 8; s8++) { ssr[lane][s8] = t[lane].xr[8 + s8]; ssi[lane][s8] = t[lane].xi[8 + s8]; }
                for (int lane = 0; lane < 32; lane++)
                    for (int h = 0; h < 2; h++) {
                        double z[8];
                        fin_z(lane, h, z);
                        for (int s4 = 0; s4 < 4; s4++) {
                            const int k = 4 * h + s4;
                            double br = ssr[lane ^ 16][k], bi = ssi[lane ^ 16][k];
                            bf(t[lane].xr[k], t[lane].xi[k], br, bi, z[2 * s4], z[2 * s4 + 1]);
                            t[lane].xr[8 + k] = br; t[lane].xi[8 + k] = bi;
                        }
                    }
                const double *bk_r = &bkw[(((size_t)i * kpl + q * l + pp) * 2) * 1024];
                for (int lane = 0; lane < 32; lane++)
                    for (int j = 0; j < 2; j++)
                        for (int p = 0; p < 16; p++) {
                            const double *b = bk_r + (size_t)j * 1024 + 2 * (p * 32 + lane);
                            cmac(sr[(lane * 2 + j) * 16 + p], si[(lane * 2 + j) * 16 + p], t[lane].xr[p], t[lane].xi[p], b[0], b[1]);
                        }
            }
        }
        for (int j = 0; j < 2; j++) {
            for (int lane = 0; lane < 32; lane++)
                for (int p = 0; p < 16; p++) { t[lane].xr[p] = sr[(lane * 2 + j) * 16 + p]; t[lane].xi[p] = si[(lane * 2 + j) * 16 + p]; }
            double ssr[32][8], ssi[32][8];
            for (int lane = 0; lane < 32; lane++)
                for (int h = 0; h < 2; h++) {
                    double z[8];
                    fin_z(lane, h, z);
                    for (int s4 = 0; s4 < 4; s4++) {
                        const int k = 4 * h + s4;
                        ibf(t[lane].xr[k], t[lane].xi[k], t[lane].xr[8 + k], t[lane].xi[8 + k], z[2 * s4], z[2 * s4 + 1]);
                        ssr[lane][k] = t[lane].xr[8 + k]; ssi[lane][k] = t[lane].xi[8 + k];
                    }
                }
            for (int lane = 0; lane < 32; lane++) for (int s8 = 0; s8 < 8; s8++) { t[lane].xr[8 + s8] = ssr[lane ^ 16][s8]; t[lane].xi[8 + s8] = ssi[lane ^ 16][s8]; }
            for (int lane = 0; lane < 32; lane++) { pass16_inv_g(t[lane].xr, t[lane].xi, e.w2[lane]); st16_ipass2(buf, lane, t[lane].xr, t[lane].xi); }
            for (int lane = 0; lane < 32; lane++) { ld16_ipass1(buf, lane, t[lane].xr, t[lane].xi); pass16_inv_p1(t[lane].xr, t[lane].xi, inv1); /* as the kernel: 6-FMA butterflies + one final twist */ }
            for (int lane = 0; lane < 32; lane++)
                for (int m = 0; m < 16; m++) {
                    acc[j * kN + lane + 32 * m] += round_to_torus(t[lane].xr[m]);
                    acc[j * kN + lane + 32 * m + 512] += round_to_torus(t[lane].xi[m]);
                }
        }
    }
    for (int j = 0; j < kN; j++) ext[j] = (j == 0) ? acc[0] : -acc[kN - j];
    ext[kN] = acc[kN];
}
/* the two forms of the last inverse pass (br_warp.h): mirrored butterflies with conj(z) factors, and 6-FMA butterflies
 * with the pending factors applied at the end */
void emul_pass16_inv_both(const double *in /*32*/, double *out_mirror /*32*/, double *out_pending /*32*/)
{
    double ar[16], ai[16], br[16], bi[16];
    for (int m = 0; m < 16; m++) { ar[m] = br[m] = in[2 * m]; ai[m] = bi[m] = in[2 * m + 1]; }
    pass16_inv(ar, ai, tw16_pass1());
    pass16_inv_p1(br, bi, inv1);
    for (int m = 0; m < 16; m++) { out_mirror[2 * m] = ar[m]; out_mirror[2 * m + 1] = ai[m]; out_pending[2 * m] = br[m]; out_pending[2 * m + 1] = bi[m]; }
}
} // extern "C"

