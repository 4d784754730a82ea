/*
 * engine.cu — host side of the C ABI in include/ieache_b200.h: device key management, batched
 * gates, the levelised circuit executor and the Cloud node's process contract.
 *
 * No CPU compute path exists here: every gate goes through launch_blind_rotate /
 * launch_keyswitch.  Host code only parses files, moves bytes and handles the nbit metadata
 * that Cloud/cloud.c itself handles on the host (cloud.c:709-855).
 */
#include <cuda_runtime.h>

#include <chrono>
#include <thread>
#include <atomic>
#include <map>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/ieache_b200.h"
#include "circuit.h"
#include "kernels.h"
#include "keygen.h"
#include "tfhe_io.h"

using namespace ieache;

static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) return fail(IEACHE_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

/* device allocation that is released on every exit path (the CU() macro returns early on errors) */
struct DevBuf {
    void *p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

extern "C" const char *ieache_last_error(void) { return g_err.c_str(); }
extern "C" const char *ieache_version(void) { return "ieache_b200 0.1 (sm_100a)"; }

/* ------------------------------------------------------------------ context */
struct TimedLaunch { cudaEvent_t a, b; int kind; };

struct ieache_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t ks_stream = nullptr;   /* key switches of chunk c overlap the blind rotation of chunk c+1 */
    cudaEvent_t ev_br[2] = {nullptr, nullptr}, ev_ks[2] = {nullptr, nullptr};
    /* host-buffer calls: chunk k+1 is copied in and chunk k-1 copied out while chunk k computes */
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    /* measured on B200 (profiles/README.md): running the key switch under the next blind rotation is
     * SLOWER (92k vs 98k gates/s) — its gather stream evicts the bootstrapping key from L2 and its CTAs
     * displace blind-rotation CTAs — so the overlap is off unless IEACHE_OVERLAP_KS=1 */
    bool overlap_ks = false;
    const void *persist_key = nullptr;
    uint64_t launches = 0;
    int32_t *d_ext = nullptr; size_t ext_cap = 0;      /* extracted samples scratch */
    int32_t *d_stage[4] = {nullptr, nullptr, nullptr, nullptr}; size_t stage_cap[4] = {0, 0, 0, 0};
    int32_t *d_wires = nullptr; size_t wires_cap = 0;
    LaunchPolicy policy;                 /* which kernel shape a launch of a given size uses: per context, no process-wide state */
    uint64_t h2d_value_copies = 0, d2h_value_copies = 0; /* operand / result blocks moved by the session calls (tests read them) */
    bool timing = false;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    std::vector<TimedLaunch> timed;
    double br_ms = 0, ks_ms = 0; uint64_t br_n = 0, ks_n = 0;
};

static int ensure(ieache_ctx *ctx, int32_t **buf, size_t *cap, size_t words)
{
    if (*cap >= words) return IEACHE_OK;
    if (*buf) {
        cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->ks_stream);
        if (ctx->h2d_stream) cudaStreamSynchronize(ctx->h2d_stream);
        if (ctx->d2h_stream) cudaStreamSynchronize(ctx->d2h_stream);
        cudaFree(*buf); *buf = nullptr; *cap = 0;
    }
    CU(cudaMalloc((void **)buf, words * sizeof(int32_t)));
    *cap = words;
    return IEACHE_OK;
}

extern "C" int ieache_ctx_create(int device, ieache_ctx **out)
{
    if (!out) return fail(IEACHE_ERR_ARG, "null out");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(IEACHE_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(IEACHE_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    std::unique_ptr<ieache_ctx> ctx(new ieache_ctx());
    ctx->device = device;
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CU(cudaStreamCreateWithPriority(&ctx->ks_stream, cudaStreamNonBlocking, prio_hi));
    CU(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&ctx->ev_br[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_ks[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_done[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
    }
    CU(upload_twiddles());
    const char *ov = getenv("IEACHE_OVERLAP_KS");
    ctx->overlap_ks = ov && atoi(ov) != 0;
    ctx->policy.sms = prop.multiProcessorCount;
    /* developer overrides of the launch policy (tools/); the per-context setter is ieache_ctx_set_tuning */
    if (const char *e = getenv("IEACHE_CLUSTER_MAX")) ctx->policy.cluster_max = atoll(e);
    if (const char *e = getenv("IEACHE_WIDE_MAX")) ctx->policy.pair_max = atoll(e);
    if (const char *e = getenv("IEACHE_W12_MIN")) ctx->policy.w12_min = atoll(e);
    if (const char *e = getenv("IEACHE_KS_STAGED_MIN")) ctx->policy.ks_staged_min = atoll(e);
    if (const char *e = getenv("IEACHE_BR_VARIANT")) { const int v = atoi(e); if (v == 0 || v == BR_GROUP || v == BR_W12) ctx->policy.throughput = v; }
    *out = ctx.release();
    return IEACHE_OK;
}
extern "C" void ieache_ctx_destroy(ieache_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &t : ctx->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    cudaFree(ctx->d_ext);
    for (int i = 0; i < 4; i++) cudaFree(ctx->d_stage[i]);
    cudaFree(ctx->d_wires);
    cudaStreamSynchronize(ctx->ks_stream);
    cudaStreamSynchronize(ctx->h2d_stream); cudaStreamSynchronize(ctx->d2h_stream);
    for (int i = 0; i < 2; i++) {
        cudaEventDestroy(ctx->ev_br[i]); cudaEventDestroy(ctx->ev_ks[i]);
        cudaEventDestroy(ctx->ev_in[i]); cudaEventDestroy(ctx->ev_done[i]); cudaEventDestroy(ctx->ev_out[i]);
    }
    cudaStreamDestroy(ctx->h2d_stream); cudaStreamDestroy(ctx->d2h_stream);
    cudaStreamDestroy(ctx->ks_stream);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}
extern "C" int ieache_ctx_sync(ieache_ctx *ctx)
{
    if (!ctx) return fail(IEACHE_ERR_ARG, "null ctx");
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->ks_stream));
    return IEACHE_OK;
}
extern "C" uint64_t ieache_ctx_launch_count(const ieache_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int ieache_ctx_set_timing(ieache_ctx *ctx, int enabled)
{
    if (!ctx) return fail(IEACHE_ERR_ARG, "null ctx");
    ctx->timing = enabled != 0;
    return IEACHE_OK;
}
static int drain_timed(ieache_ctx *ctx)
{
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->ks_stream));
    for (auto &t : ctx->timed) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, t.a, t.b));
        if (t.kind == 0) { ctx->br_ms += ms; ctx->br_n++; } else { ctx->ks_ms += ms; ctx->ks_n++; }
        cudaEventDestroy(t.a); cudaEventDestroy(t.b);
    }
    ctx->timed.clear();
    return IEACHE_OK;
}
extern "C" int ieache_ctx_kernel_times(ieache_ctx *ctx, double *br_ms, double *ks_ms, uint64_t *br_n, uint64_t *ks_n, int reset)
{
    if (!ctx) return fail(IEACHE_ERR_ARG, "null ctx");
    int rc = drain_timed(ctx);
    if (rc) return rc;
    if (br_ms) *br_ms = ctx->br_ms;
    if (ks_ms) *ks_ms = ctx->ks_ms;
    if (br_n) *br_n = ctx->br_n;
    if (ks_n) *ks_n = ctx->ks_n;
    if (reset) { ctx->br_ms = ctx->ks_ms = 0; ctx->br_n = ctx->ks_n = 0; }
    return IEACHE_OK;
}

extern "C" int ieache_ctx_set_tuning(ieache_ctx *ctx, int which, int64_t value, int64_t *old_value)
{
    if (!ctx) return fail(IEACHE_ERR_ARG, "null ctx");
    LaunchPolicy &pol = ctx->policy;
    long long *slot = nullptr;
    switch (which) {
    case IEACHE_TUNE_CLUSTER_MAX: slot = &pol.cluster_max; break;
    case IEACHE_TUNE_PAIR_MAX: slot = &pol.pair_max; break;
    case IEACHE_TUNE_W12_MIN: slot = &pol.w12_min; break;
    case IEACHE_TUNE_KS_STAGED_MIN: slot = &pol.ks_staged_min; break;
    case IEACHE_TUNE_THROUGHPUT_KERNEL:
        if (old_value) *old_value = pol.throughput;
        if (value != 0 && value != IEACHE_KERNEL_GROUP && value != IEACHE_KERNEL_W12)
            return fail(IEACHE_ERR_ARG, "throughput kernel must be 0 (by size), %d or %d (got %lld)", IEACHE_KERNEL_GROUP, IEACHE_KERNEL_W12, (long long)value);
        pol.throughput = (int)value;
        return IEACHE_OK;
    default: return fail(IEACHE_ERR_ARG, "unknown tuning parameter %d", which);
    }
    if (old_value) *old_value = *slot;
    if (value < 0) return fail(IEACHE_ERR_ARG, "negative threshold");
    *slot = value;
    return IEACHE_OK;
}
extern "C" int ieache_ctx_copy_counts(const ieache_ctx *ctx, uint64_t *h2d_value_blocks, uint64_t *d2h_value_blocks)
{
    if (!ctx) return fail(IEACHE_ERR_ARG, "null ctx");
    if (h2d_value_blocks) *h2d_value_blocks = ctx->h2d_value_copies;
    if (d2h_value_blocks) *d2h_value_blocks = ctx->d2h_value_copies;
    return IEACHE_OK;
}
extern "C" int ieache_ctx_timer_start(ieache_ctx *ctx)
{
    if (!ctx) return fail(IEACHE_ERR_ARG, "null ctx");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->t0) { CU(cudaEventCreate(&ctx->t0)); CU(cudaEventCreate(&ctx->t1)); }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventRecord(ctx->t0, ctx->stream));
    return IEACHE_OK;
}
extern "C" int ieache_ctx_timer_stop(ieache_ctx *ctx, double *elapsed_ms)
{
    if (!ctx || !ctx->t0 || !elapsed_ms) return fail(IEACHE_ERR_ARG, "timer not started");
    CU(cudaEventRecord(ctx->t1, ctx->stream));
    CU(cudaEventSynchronize(ctx->t1));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->t0, ctx->t1));
    *elapsed_ms = ms;
    return IEACHE_OK;
}
extern "C" int ieache_measure_fp64_peak(ieache_ctx *ctx, double *tflops)
{
    if (!ctx || !tflops) return fail(IEACHE_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    double best = 0;
    CU(launch_fp64_peak(ctx->stream, &best));
    ctx->launches += 4;
    *tflops = best;
    return IEACHE_OK;
}
extern "C" int ieache_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(IEACHE_ERR_ARG, "null out");
    CU(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return IEACHE_OK;
}
extern "C" int ieache_host_free(void *ptr)
{
    CU(cudaFreeHost(ptr));
    return IEACHE_OK;
}

/* ------------------------------------------------------------------ cloud key */
struct ieache_cloudkey {
    ieache_ctx *ctx = nullptr;
    ieache_params p{};
    DevParams dp{};
    double2 *bkfft = nullptr; size_t bkfft_bytes = 0;
    mutable double2 *bkfft_w = nullptr;                /* the persistent kernel's layout of the same values, made on its first use */
    int32_t *ksk = nullptr; size_t ksk_bytes = 0;
    bool owns = true;
};

static int check_params(const ieache_params *p)
{
    if (!p) return fail(IEACHE_ERR_ARG, "null params");
    if (p->N != 1024 || p->k != 1) return fail(IEACHE_ERR_UNSUPPORTED, "kernels support N=1024,k=1 (got N=%d,k=%d)", p->N, p->k);
    if (p->bk_l != 2 && p->bk_l != 3) return fail(IEACHE_ERR_UNSUPPORTED, "kernels support bk_l in {2,3} (got %d)", p->bk_l);
    if (p->n < 1 || p->n > kLweStride - 1) return fail(IEACHE_ERR_UNSUPPORTED, "kernels support 1 <= n <= %d (got %d)", kLweStride - 1, p->n);
    if (p->bk_Bgbit < 1 || p->bk_Bgbit * p->bk_l > 31) return fail(IEACHE_ERR_UNSUPPORTED, "unsupported Bgbit=%d", p->bk_Bgbit);
    if (p->ks_t < 1 || p->ks_basebit < 1 || p->ks_basebit * p->ks_t > 31 || p->ks_t > 16) return fail(IEACHE_ERR_UNSUPPORTED, "unsupported key-switch parameters t=%d basebit=%d", p->ks_t, p->ks_basebit);
    return IEACHE_OK;
}
static void fill_dev_params(const ieache_params &p, DevParams &dp)
{
    dp.n = p.n; dp.l = p.bk_l; dp.Bgbit = p.bk_Bgbit; dp.ks_t = p.ks_t; dp.ks_basebit = p.ks_basebit;
    dp.mu = 1 << 29; /* modSwitchToTorus32(1, 8) */
}
extern "C" int ieache_ctx_pick_kernels(const ieache_ctx *ctx, const ieache_cloudkey *key, int64_t count, int *blind_rotate, int *keyswitch)
{
    if (!ctx || !key) return fail(IEACHE_ERR_ARG, "null argument");
    if (blind_rotate) *blind_rotate = pick_blind_rotate(ctx->policy, count);
    if (keyswitch) *keyswitch = pick_keyswitch(ctx->policy, key->dp, count);
    return IEACHE_OK;
}
extern "C" int ieache_cloudkey_device_sizes(const ieache_params *p, size_t *bkfft_bytes, size_t *ksk_bytes)
{
    int rc = check_params(p);
    if (rc) return rc;
    if (bkfft_bytes) *bkfft_bytes = (size_t)p->n * 2 * p->bk_l * 2 * 512 * sizeof(double2);
    if (ksk_bytes) *ksk_bytes = (size_t)p->N * p->ks_t * ((1u << p->ks_basebit) - 1) * kLweStride * sizeof(int32_t);
    return IEACHE_OK;
}

extern "C" void ieache_cloudkey_destroy(ieache_cloudkey *key);

extern "C" int ieache_cloudkey_create(ieache_ctx *ctx, const ieache_params *p, const int32_t *bk, const int32_t *ksk,
                                      ieache_cloudkey **out)
{
    if (!ctx || !bk || !ksk || !out) return fail(IEACHE_ERR_ARG, "null argument");
    int rc = check_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    /* the key object owns its arrays from the moment they exist: an early return frees them through the deleter */
    std::unique_ptr<ieache_cloudkey, void (*)(ieache_cloudkey *)> key(new ieache_cloudkey(), ieache_cloudkey_destroy);
    key->ctx = ctx; key->p = *p; fill_dev_params(*p, key->dp);
    ieache_cloudkey_device_sizes(p, &key->bkfft_bytes, &key->ksk_bytes);
    const int kpl = 2 * p->bk_l, base = 1 << p->ks_basebit;
    const size_t bk_words = (size_t)p->n * kpl * 2 * 1024;
    const size_t ksk_words = (size_t)1024 * p->ks_t * base * (p->n + 1);
    DevBuf tmp;
    CU(cudaMalloc((void **)&key->bkfft, key->bkfft_bytes));
    CU(cudaMalloc((void **)&key->ksk, key->ksk_bytes));
    CU(tmp.alloc(std::max(bk_words, ksk_words) * sizeof(int32_t)));
    /* bkFFT on the GPU (libtfhe builds it on the CPU at key load) */
    CU(cudaMemcpyAsync(tmp.p, bk, bk_words * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_bk_fft(tmp.as<int32_t>(), key->bkfft, p->n * kpl * 2, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaMemcpyAsync(tmp.p, ksk, ksk_words * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(launch_pack_ksk(tmp.as<int32_t>(), key->ksk, 1024, p->ks_t, base, p->n, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->launches += 2;
    *out = key.release();
    return IEACHE_OK;
}

extern "C" int ieache_cloudkey_load_file(ieache_ctx *ctx, const char *path, ieache_cloudkey **out)
{
    if (!ctx || !path || !out) return fail(IEACHE_ERR_ARG, "null argument");
    HostKeySet hk;
    std::string msg;
    int rc = read_keyset(path, hk, true, msg);
    if (rc) return fail(rc, "%s", msg.c_str());
    return ieache_cloudkey_create(ctx, &hk.p, hk.bk.data(), hk.ksk.data(), out);
}
extern "C" void ieache_cloudkey_destroy(ieache_cloudkey *key)
{
    if (!key) return;
    if (key->owns || key->bkfft_w) {
        cudaSetDevice(key->ctx->device);
        cudaStreamSynchronize(key->ctx->stream);
    }
    if (key->owns) { cudaFree(key->bkfft); cudaFree(key->ksk); }
    cudaFree(key->bkfft_w); /* always owned: derived from bkfft on first use */
    delete key;
}
extern "C" int ieache_cloudkey_params(const ieache_cloudkey *key, ieache_params *out)
{
    if (!key || !out) return fail(IEACHE_ERR_ARG, "null argument");
    *out = key->p;
    return IEACHE_OK;
}
extern "C" int ieache_cloudkey_device_arrays(const ieache_cloudkey *key, void **bkfft, size_t *bkfft_bytes, void **ksk, size_t *ksk_bytes)
{
    if (!key) return fail(IEACHE_ERR_ARG, "null key");
    if (bkfft) *bkfft = key->bkfft;
    if (bkfft_bytes) *bkfft_bytes = key->bkfft_bytes;
    if (ksk) *ksk = key->ksk;
    if (ksk_bytes) *ksk_bytes = key->ksk_bytes;
    return IEACHE_OK;
}
extern "C" int ieache_cloudkey_adopt_device(ieache_ctx *ctx, const ieache_params *p, void *bkfft, void *ksk, ieache_cloudkey **out)
{
    if (!ctx || !bkfft || !ksk || !out) return fail(IEACHE_ERR_ARG, "null argument");
    int rc = check_params(p);
    if (rc) return rc;
    ieache_cloudkey *key = new ieache_cloudkey();
    key->ctx = ctx; key->p = *p; fill_dev_params(*p, key->dp);
    ieache_cloudkey_device_sizes(p, &key->bkfft_bytes, &key->ksk_bytes);
    key->bkfft = (double2 *)bkfft; key->ksk = (int32_t *)ksk; key->owns = false;
    *out = key;
    return IEACHE_OK;
}

/* ------------------------------------------------------------------ secret-key side on the GPU */
struct ieache_secretkey {
    ieache_ctx *ctx = nullptr;
    ieache_params p{};
    std::vector<int32_t> lwe_key, tlwe_key;
    int32_t *d_lwe_key = nullptr, *d_tlwe_key = nullptr;
};
extern "C" void ieache_secretkey_destroy(ieache_secretkey *sk)
{
    if (!sk) return;
    cudaSetDevice(sk->ctx->device);
    cudaFree(sk->d_lwe_key); cudaFree(sk->d_tlwe_key);
    delete sk;
}
extern "C" int ieache_secretkey_import(ieache_ctx *ctx, const ieache_params *p, const int32_t *lwe_key, const int32_t *tlwe_key,
                                       ieache_secretkey **out)
{
    if (!ctx || !lwe_key || !out) return fail(IEACHE_ERR_ARG, "null argument");
    int rc = check_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    std::unique_ptr<ieache_secretkey, void (*)(ieache_secretkey *)> sk(new ieache_secretkey(), ieache_secretkey_destroy);
    sk->ctx = ctx; sk->p = *p;
    sk->lwe_key.assign(lwe_key, lwe_key + p->n);
    sk->tlwe_key.assign(1024, 0);
    if (tlwe_key) sk->tlwe_key.assign(tlwe_key, tlwe_key + 1024);
    CU(cudaMalloc((void **)&sk->d_lwe_key, kLweStride * 4));
    CU(cudaMalloc((void **)&sk->d_tlwe_key, 1024 * 4));
    CU(cudaMemset(sk->d_lwe_key, 0, kLweStride * 4));
    CU(cudaMemcpy(sk->d_lwe_key, sk->lwe_key.data(), (size_t)p->n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sk->d_tlwe_key, sk->tlwe_key.data(), 1024 * 4, cudaMemcpyHostToDevice));
    *out = sk.release();
    return IEACHE_OK;
}
extern "C" int ieache_secretkey_export(const ieache_secretkey *sk, int32_t *lwe_key, int32_t *tlwe_key)
{
    if (!sk || !lwe_key) return fail(IEACHE_ERR_ARG, "null argument");
    memcpy(lwe_key, sk->lwe_key.data(), sk->lwe_key.size() * 4);
    if (tlwe_key) memcpy(tlwe_key, sk->tlwe_key.data(), sk->tlwe_key.size() * 4);
    return IEACHE_OK;
}
extern "C" int ieache_keygen(ieache_ctx *ctx, const ieache_params *p, uint64_t seed, ieache_secretkey **sk_out, ieache_cloudkey **ck_out,
                             int32_t *bk_export, int32_t *ksk_export)
{
    if (!ctx || !sk_out || !ck_out) return fail(IEACHE_ERR_ARG, "null argument");
    int rc = check_params(p);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    /* seed 0: both 256-bit stream keys come from the operating system; a non-zero seed gives a reproducible key set
     * for tests (and no more secrecy than the 64-bit seed) */
    RngKeys rk;
    if (seed == 0) { if (rng_keys_from_os(rk)) return fail(IEACHE_ERR_IO, "getrandom failed"); }
    else rng_keys_from_seed(seed, rk);
    std::vector<int32_t> lwe(p->n), tlwe(1024);
    host_key_bits(rk.secret, RNG_LWE_KEY, lwe.data(), p->n);
    host_key_bits(rk.secret, RNG_TLWE_KEY, tlwe.data(), 1024);
    ieache_secretkey *sk = nullptr;
    if ((rc = ieache_secretkey_import(ctx, p, lwe.data(), tlwe.data(), &sk))) return rc;
    std::unique_ptr<ieache_secretkey, void (*)(ieache_secretkey *)> skg(sk, ieache_secretkey_destroy);
    std::unique_ptr<ieache_cloudkey, void (*)(ieache_cloudkey *)> key(new ieache_cloudkey(), ieache_cloudkey_destroy);
    key->ctx = ctx; key->p = *p; fill_dev_params(*p, key->dp);
    ieache_cloudkey_device_sizes(p, &key->bkfft_bytes, &key->ksk_bytes);
    CU(cudaMalloc((void **)&key->bkfft, key->bkfft_bytes));
    CU(cudaMalloc((void **)&key->ksk, key->ksk_bytes));
    const int kpl = 2 * p->bk_l, base = 1 << p->ks_basebit;
    const size_t bk_words = (size_t)p->n * kpl * 2 * 1024, ksk_words = (size_t)1024 * p->ks_t * base * (p->n + 1);
    DevBuf shat, dbk, dks;
    CU(shat.alloc(512 * sizeof(double2)));
    if (bk_export) CU(dbk.alloc(bk_words * 4));
    if (ksk_export) { CU(dks.alloc(ksk_words * 4)); CU(cudaMemsetAsync(dks.p, 0, ksk_words * 4, ctx->stream)); }
    CU(launch_keygen(rk, key->dp, p->ks_stdev, p->bk_stdev, sk->d_lwe_key, sk->d_tlwe_key, shat.as<double2>(), key->bkfft, key->ksk,
                     dbk.as<int32_t>(), dks.as<int32_t>(), ctx->stream));
    ctx->launches += 3;
    if (bk_export) CU(cudaMemcpyAsync(bk_export, dbk.p, bk_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (ksk_export) CU(cudaMemcpyAsync(ksk_export, dks.p, ksk_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *sk_out = skg.release();
    *ck_out = key.release();
    return IEACHE_OK;
}
extern "C" int ieache_sym_encrypt_device(ieache_ctx *ctx, const ieache_secretkey *sk, const int32_t *bits, size_t count, int32_t *out_dev,
                                         uint64_t seed)
{
    if (!ctx || !sk || !bits || !out_dev) return fail(IEACHE_ERR_ARG, "null argument");
    if (count == 0) return IEACHE_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf dbits;
    CU(dbits.alloc(count * 4));
    int32_t *d_bits = dbits.as<int32_t>();
    CU(cudaMemcpyAsync(d_bits, bits, count * 4, cudaMemcpyHostToDevice, ctx->stream));
    RngKeys rk;
    if (seed == 0) { if (rng_keys_from_os(rk)) return fail(IEACHE_ERR_IO, "getrandom failed"); } /* fresh masks and noise per call */
    else rng_keys_from_seed(seed, rk);                                                            /* tests only: the same seed repeats the masks */
    CU(launch_encrypt(rk, sk->p.n, sk->p.ks_stdev, 1 << 29, sk->d_lwe_key, d_bits, out_dev, (long long)count, ctx->stream));
    ctx->launches++;
    CU(cudaStreamSynchronize(ctx->stream));
    return IEACHE_OK;
}
extern "C" int ieache_sym_decrypt_device(ieache_ctx *ctx, const ieache_secretkey *sk, const int32_t *samples_dev, size_t count,
                                         int32_t *bits, int32_t *phases)
{
    if (!ctx || !sk || !samples_dev || (!bits && !phases)) return fail(IEACHE_ERR_ARG, "null argument");
    if (count == 0) return IEACHE_OK;
    CU(cudaSetDevice(ctx->device));
    DevBuf dph;
    CU(dph.alloc(count * 4));
    int32_t *d_ph = dph.as<int32_t>();
    CU(launch_phase(sk->p.n, sk->d_lwe_key, samples_dev, d_ph, (long long)count, ctx->stream));
    ctx->launches++;
    std::vector<int32_t> ph(count);
    CU(cudaMemcpyAsync(ph.data(), d_ph, count * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (phases) memcpy(phases, ph.data(), count * 4);
    if (bits) for (size_t i = 0; i < count; i++) bits[i] = ph[i] > 0 ? 1 : 0;
    return IEACHE_OK;
}

/* Optional (IEACHE_L2_PERSIST=1): pin the transform-domain bootstrapping key in L2 for launches on the
 * blind-rotation stream, so that key-switch traffic cannot evict it. */
static int apply_l2_persist(ieache_ctx *ctx, const ieache_cloudkey *key)
{
    static const int want = [] { const char *e = getenv("IEACHE_L2_PERSIST"); return e ? atoi(e) : 0; }();
    if (!want || ctx->persist_key == key->bkfft) return IEACHE_OK;
    int max_win = 0, max_persist = 0;
    CU(cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device));
    CU(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device));
    const size_t bytes = std::min<size_t>(key->bkfft_bytes, (size_t)max_win);
    CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)max_persist)));
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = key->bkfft;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = std::min(1.0f, (float)max_persist / (float)bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CU(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    ctx->persist_key = key->bkfft;
    if (want > 1) fprintf(stderr, "ieache: L2 persist window %zu B, max persisting %d B, max window %d B\n", bytes, max_persist, max_win);
    return IEACHE_OK;
}

/* ------------------------------------------------------------------ launches with optional timing */
constexpr size_t kHalfNBytes = 512 * sizeof(double2);
static int run_br(ieache_ctx *ctx, const ieache_cloudkey *key, const GateAddr &ga, const int32_t *A, const int32_t *B, int ext_base,
                  int32_t *ext = nullptr)
{
    TimedLaunch t{};
    { int rcp = apply_l2_persist(ctx, key); if (rcp) return rcp; }
    if (ctx->timing) { CU(cudaEventCreate(&t.a)); CU(cudaEventCreate(&t.b)); t.kind = 0; CU(cudaEventRecord(t.a, ctx->stream)); }
    if (!key->bkfft_w && pick_blind_rotate(ctx->policy, (long long)ga.ntempl * ga.n_inst) == BR_W12) {
        CU(cudaMalloc((void **)&key->bkfft_w, key->bkfft_bytes));
        CU(launch_bk_relayout_w12(key->bkfft, key->bkfft_w, (int)(key->bkfft_bytes / (kHalfNBytes)), key->dp.l, key->dp.Bgbit, ctx->stream));
        ctx->launches++;
    }
    CU(launch_blind_rotate(key->dp, ctx->policy, key->bkfft, key->bkfft_w, ga, A, B, ext ? ext : ctx->d_ext, ext_base, ctx->stream));
    if (ctx->timing) { CU(cudaEventRecord(t.b, ctx->stream)); ctx->timed.push_back(t); }
    ctx->launches++;
    return IEACHE_OK;
}
static int run_ks(ieache_ctx *ctx, const ieache_cloudkey *key, const GateAddr &ga, int32_t *out, int pair_offset, int32_t cst_post,
                  const int32_t *ext = nullptr, cudaStream_t st = nullptr)
{
    TimedLaunch t{};
    if (!st) st = ctx->stream;
    if (ctx->timing) { CU(cudaEventCreate(&t.a)); CU(cudaEventCreate(&t.b)); t.kind = 1; CU(cudaEventRecord(t.a, st)); }
    CU(launch_keyswitch(key->dp, ctx->policy, key->ksk, ga, out, ext ? ext : ctx->d_ext, pair_offset, cst_post, st));
    if (ctx->timing) { CU(cudaEventRecord(t.b, st)); ctx->timed.push_back(t); }
    ctx->launches++;
    if (ctx->timed.size() > 4096) return drain_timed(ctx);
    return IEACHE_OK;
}

/* ------------------------------------------------------------------ batched gates */
static const int8_t k_lin[10][3] = {{+1, -1, -1}, {+1, +1, +1}, {-1, +1, +1}, {+2, +2, +2}, {-2, -2, -2},
                                    {-1, -1, -1}, {-1, -1, +1}, {-1, +1, -1}, {+1, -1, +1}, {+1, +1, -1}};
constexpr size_t kChunk = 1u << 16; /* gates per launch: bounds the extracted-sample scratch to 2 x 135 MB (x2 for MUX) */

extern "C" int ieache_gate_batch_device(ieache_ctx *ctx, const ieache_cloudkey *key, int op, int32_t *out, const int32_t *a,
                                        const int32_t *b, const int32_t *c, int32_t imm, size_t count)
{
    if (!ctx || !key || !out) return fail(IEACHE_ERR_ARG, "null argument");
    if (key->ctx->device != ctx->device) return fail(IEACHE_ERR_ARG, "key lives on another device");
    if (count == 0) return IEACHE_OK;
    CU(cudaSetDevice(ctx->device));
    const int n = key->p.n, mu = key->dp.mu;
    if (op == IEACHE_OP_CONST) { CU(launch_linear(out, nullptr, (int)count, kLweStride, n, 0, imm ? mu : -mu, ctx->stream)); ctx->launches++; return IEACHE_OK; }
    if (!a) return fail(IEACHE_ERR_ARG, "null input a");
    if (op == IEACHE_OP_COPY || op == IEACHE_OP_NOT) {
        CU(launch_linear(out, a, (int)count, kLweStride, n, op == IEACHE_OP_NOT ? -1 : 1, 0, ctx->stream));
        ctx->launches++;
        return IEACHE_OK;
    }
    if (op < 0 || op > IEACHE_OP_MUX) return fail(IEACHE_ERR_ARG, "unknown op %d", op);
    if (!b || (op == IEACHE_OP_MUX && !c)) return fail(IEACHE_ERR_ARG, "null input operand");
    /* chunks of <= 32768 gates, double-buffered extracted samples: the (memory-bound) key switch of
     * chunk c runs on a second, higher-priority stream under the (FP64-bound) blind rotation of c+1 */
    const size_t per = std::min(count, kChunk);
    const size_t slot_words = per * kExtStride * (op == IEACHE_OP_MUX ? 2 : 1);
    int rc = ensure(ctx, &ctx->d_ext, &ctx->ext_cap, 2 * slot_words);
    if (rc) return rc;
    size_t chunk = 0;
    for (size_t off = 0; off < count; off += per, chunk++) {
        const size_t m = std::min(per, count - off);
        const int slot = (int)(chunk & 1);
        int32_t *ext = ctx->d_ext + slot * slot_words;
        if (chunk >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_ks[slot], 0));
        GateAddr ga{};
        ga.tmpl = nullptr; ga.ntempl = (int)m; ga.n_inst = 1; ga.inst_samples = 0; ga.stride = kLweStride;
        const int32_t *pa = a + off * kLweStride, *pb = b + off * kLweStride;
        int32_t *po = out + off * kLweStride;
        if (op == IEACHE_OP_MUX) {
            /* bootsMUX: u1 = BR((0,-mu)+a+b), u2 = BR((0,-mu)-a+c), out = KS((0,mu)+u1+u2) */
            const int32_t *pc = c + off * kLweStride;
            ga.uni = GateT{0, 0, 0, +1, +1, -1};
            if ((rc = run_br(ctx, key, ga, pa, pb, 0, ext))) return rc;
            ga.uni = GateT{0, 0, 0, -1, +1, -1};
            if ((rc = run_br(ctx, key, ga, pa, pc, (int)m, ext))) return rc;
        } else {
            ga.uni = GateT{0, 0, 0, k_lin[op][1], k_lin[op][2], k_lin[op][0]};
            if ((rc = run_br(ctx, key, ga, pa, pb, 0, ext))) return rc;
        }
        cudaStream_t kst = ctx->overlap_ks ? ctx->ks_stream : ctx->stream;
        CU(cudaEventRecord(ctx->ev_br[slot], ctx->stream));
        if (ctx->overlap_ks) CU(cudaStreamWaitEvent(kst, ctx->ev_br[slot], 0));
        if ((rc = run_ks(ctx, key, ga, po, op == IEACHE_OP_MUX ? (int)m : 0, op == IEACHE_OP_MUX ? mu : 0, ext, kst))) return rc;
        CU(cudaEventRecord(ctx->ev_ks[slot], kst));
    }
    /* join: later work on the context stream sees every key switch */
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_ks[0], 0));
    if (chunk >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_ks[1], 0));
    return IEACHE_OK;
}

extern "C" int ieache_samples_to_device(ieache_ctx *ctx, int32_t *dev, const int32_t *host, size_t count, int32_t n)
{
    if (!ctx || !dev || !host) return fail(IEACHE_ERR_ARG, "null argument");
    if (count == 0) return IEACHE_OK;
    CU(cudaMemcpy2DAsync(dev, kLweStride * 4, host, (size_t)(n + 1) * 4, (size_t)(n + 1) * 4, count, cudaMemcpyHostToDevice, ctx->stream));
    return IEACHE_OK;
}
extern "C" int ieache_samples_to_host(ieache_ctx *ctx, int32_t *host, const int32_t *dev, size_t count, int32_t n)
{
    if (!ctx || !dev || !host) return fail(IEACHE_ERR_ARG, "null argument");
    if (count == 0) return IEACHE_OK;
    CU(cudaMemcpy2DAsync(host, (size_t)(n + 1) * 4, dev, kLweStride * 4, (size_t)(n + 1) * 4, count, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return IEACHE_OK;
}
extern "C" int ieache_device_alloc(ieache_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return fail(IEACHE_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(out, bytes));
    return IEACHE_OK;
}
extern "C" int ieache_device_copy(ieache_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx || !dst || !src) return fail(IEACHE_ERR_ARG, "null argument");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return IEACHE_OK;
}
extern "C" int ieache_device_free(ieache_ctx *ctx, void *ptr)
{
    if (!ctx) return fail(IEACHE_ERR_ARG, "null ctx");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(ptr));
    return IEACHE_OK;
}

extern "C" int ieache_gate_batch(ieache_ctx *ctx, const ieache_cloudkey *key, int op, int32_t *out, const int32_t *a,
                                 const int32_t *b, const int32_t *c, int32_t imm, size_t count)
{
    if (!ctx || !key || !out) return fail(IEACHE_ERR_ARG, "null argument");
    if (count == 0) return IEACHE_OK;
    CU(cudaSetDevice(ctx->device));
    const int n = key->p.n;
    const size_t row = (size_t)(n + 1) * 4;
    const int32_t *src[3] = {a, b, c};
    /* chunks of <= 65536 gates through two staging slots: the inputs of chunk k+1 go up on their own stream and the
     * results of chunk k-1 come down on a third one while chunk k computes (pinned host buffers make the copies
     * asynchronous; pageable ones still work, serialised by the driver) */
    const size_t per = std::min(count, kChunk);
    int rc;
    for (int i = 0; i < 4; i++)
        if (i == 3 || src[i])
            if ((rc = ensure(ctx, &ctx->d_stage[i], &ctx->stage_cap[i], 2 * per * kLweStride))) return rc;
    /* everything queued earlier on the context stream (and its use of the staging buffers) comes first */
    CU(cudaEventRecord(ctx->ev_done[0], ctx->stream));
    CU(cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_done[0], 0));
    size_t chunk = 0;
    for (size_t off = 0; off < count; off += per, chunk++) {
        const size_t m = std::min(per, count - off);
        const int slot = (int)(chunk & 1);
        const size_t so = (size_t)slot * per * kLweStride;
        if (chunk >= 2) CU(cudaStreamWaitEvent(ctx->h2d_stream, ctx->ev_done[slot], 0)); /* chunk-2 no longer reads these inputs */
        for (int i = 0; i < 3; i++)
            if (src[i])
                CU(cudaMemcpy2DAsync(ctx->d_stage[i] + so, kLweStride * 4, src[i] + off * (size_t)(n + 1), row, row, m,
                                     cudaMemcpyHostToDevice, ctx->h2d_stream));
        CU(cudaEventRecord(ctx->ev_in[slot], ctx->h2d_stream));
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_in[slot], 0));
        if (chunk >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_out[slot], 0));      /* chunk-2's results have left the slot */
        rc = ieache_gate_batch_device(ctx, key, op, ctx->d_stage[3] + so, a ? ctx->d_stage[0] + so : nullptr,
                                      b ? ctx->d_stage[1] + so : nullptr, c ? ctx->d_stage[2] + so : nullptr, imm, m);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev_done[slot], ctx->stream));
        CU(cudaStreamWaitEvent(ctx->d2h_stream, ctx->ev_done[slot], 0));
        CU(cudaMemcpy2DAsync(out + off * (size_t)(n + 1), row, ctx->d_stage[3] + so, kLweStride * 4, row, m, cudaMemcpyDeviceToHost,
                             ctx->d2h_stream));
        CU(cudaEventRecord(ctx->ev_out[slot], ctx->d2h_stream));
    }
    CU(cudaStreamSynchronize(ctx->d2h_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return IEACHE_OK;
}

extern "C" int ieache_bootstrap_woks(ieache_ctx *ctx, const ieache_cloudkey *key, int32_t *ext_out, const int32_t *x, size_t count)
{
    if (!ctx || !key || !ext_out || !x) return fail(IEACHE_ERR_ARG, "null argument");
    if (count == 0) return IEACHE_OK;
    if (count > kChunk) return fail(IEACHE_ERR_ARG, "count > %zu", kChunk);
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, &ctx->d_stage[0], &ctx->stage_cap[0], count * kLweStride))) return rc;
    if ((rc = ensure(ctx, &ctx->d_ext, &ctx->ext_cap, count * kExtStride))) return rc;
    if ((rc = ieache_samples_to_device(ctx, ctx->d_stage[0], x, count, key->p.n))) return rc;
    GateAddr ga{};
    ga.tmpl = nullptr; ga.ntempl = (int)count; ga.n_inst = 1; ga.stride = kLweStride;
    ga.uni = GateT{0, -1, 0, +1, 0, 0};
    if ((rc = run_br(ctx, key, ga, ctx->d_stage[0], ctx->d_stage[0], 0))) return rc;
    CU(cudaMemcpy2DAsync(ext_out, 1025 * 4, ctx->d_ext, kExtStride * 4, 1025 * 4, count, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return IEACHE_OK;
}
extern "C" int ieache_keyswitch(ieache_ctx *ctx, const ieache_cloudkey *key, int32_t *out, const int32_t *ext, size_t count)
{
    if (!ctx || !key || !out || !ext) return fail(IEACHE_ERR_ARG, "null argument");
    if (count == 0) return IEACHE_OK;
    if (count > kChunk) return fail(IEACHE_ERR_ARG, "count > %zu", kChunk);
    CU(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, &ctx->d_stage[3], &ctx->stage_cap[3], count * kLweStride))) return rc;
    if ((rc = ensure(ctx, &ctx->d_ext, &ctx->ext_cap, count * kExtStride))) return rc;
    CU(cudaMemcpy2DAsync(ctx->d_ext, kExtStride * 4, ext, 1025 * 4, 1025 * 4, count, cudaMemcpyHostToDevice, ctx->stream));
    GateAddr ga{};
    ga.tmpl = nullptr; ga.ntempl = (int)count; ga.n_inst = 1; ga.stride = kLweStride;
    if ((rc = run_ks(ctx, key, ga, ctx->d_stage[3], 0, 0))) return rc;
    return ieache_samples_to_host(ctx, out, ctx->d_stage[3], count, key->p.n);
}

/* ------------------------------------------------------------------ circuits */
struct ieache_circuit {
    Circuit c;
    /* device copies of the level templates, per device, created lazily */
    int device = -1;
    GateT *d_tmpl = nullptr;
    std::vector<size_t> level_off;
    int32_t *d_out_slots = nullptr; /* wire slot of every output, uploaded with the templates */
};

extern "C" int ieache_circuit_build(int kind, int width, ieache_circuit **out)
{
    if (!out) return fail(IEACHE_ERR_ARG, "null out");
    std::unique_ptr<ieache_circuit> c(new ieache_circuit());
    if (!build_circuit(kind, width, c->c)) return fail(IEACHE_ERR_UNSUPPORTED, "no circuit for kind %d at width %d", kind, width);
    *out = c.release();
    return IEACHE_OK;
}
extern "C" void ieache_circuit_destroy(ieache_circuit *c)
{
    if (!c) return;
    if (c->d_tmpl) { cudaSetDevice(c->device); cudaFree(c->d_tmpl); cudaFree(c->d_out_slots); }
    delete c;
}
extern "C" int ieache_circuit_stats(const ieache_circuit *c, uint64_t *bootstraps, uint64_t *and_gates, uint64_t *xor_gates,
                                    uint32_t *levels, uint32_t *max_width, uint32_t *n_inputs, uint32_t *n_outputs)
{
    if (!c) return fail(IEACHE_ERR_ARG, "null circuit");
    if (bootstraps) *bootstraps = c->c.gates.size();
    if (and_gates) *and_gates = c->c.n_and;
    if (xor_gates) *xor_gates = c->c.n_xor;
    if (levels) *levels = (uint32_t)c->c.levels.size();
    if (max_width) *max_width = c->c.max_width;
    if (n_inputs) *n_inputs = (uint32_t)c->c.n_inputs;
    if (n_outputs) *n_outputs = (uint32_t)c->c.outputs.size();
    return IEACHE_OK;
}

static int circuit_upload(ieache_ctx *ctx, ieache_circuit *c)
{
    if (c->d_tmpl && c->device == ctx->device) return IEACHE_OK;
    if (c->d_tmpl) { cudaSetDevice(c->device); cudaFree(c->d_tmpl); cudaFree(c->d_out_slots); c->d_tmpl = nullptr; c->d_out_slots = nullptr; }
    CU(cudaSetDevice(ctx->device));
    /* output slot table (sign folded: outputs of cloud.c circuits are never negated refs) */
    std::vector<int32_t> out_slots(c->c.outputs.size());
    for (size_t i = 0; i < c->c.outputs.size(); i++) {
        if (c->c.outputs[i].neg) return fail(IEACHE_ERR_UNSUPPORTED, "negated output reference");
        out_slots[i] = c->c.slot_of_wire[c->c.outputs[i].wire];
    }
    std::vector<GateT> all;
    c->level_off.clear();
    for (const Level &lv : c->c.levels) { c->level_off.push_back(all.size()); all.insert(all.end(), lv.tmpl.begin(), lv.tmpl.end()); }
    CU(cudaMalloc((void **)&c->d_tmpl, all.size() * sizeof(GateT)));
    CU(cudaMemcpy(c->d_tmpl, all.data(), all.size() * sizeof(GateT), cudaMemcpyHostToDevice));
    CU(cudaMalloc((void **)&c->d_out_slots, out_slots.size() * 4));
    CU(cudaMemcpy(c->d_out_slots, out_slots.data(), out_slots.size() * 4, cudaMemcpyHostToDevice));
    c->device = ctx->device;
    return IEACHE_OK;
}

/* launches everything on the context's stream and returns without waiting */
static int circuit_eval_device_async(ieache_ctx *ctx, const ieache_cloudkey *key, const ieache_circuit *cc, const int32_t *inputs,
                                     int32_t *outputs, size_t n_expr);
extern "C" int ieache_circuit_eval_device(ieache_ctx *ctx, const ieache_cloudkey *key, const ieache_circuit *cc, const int32_t *inputs,
                                          int32_t *outputs, size_t n_expr)
{
    int rc = circuit_eval_device_async(ctx, key, cc, inputs, outputs, n_expr);
    if (rc) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    return IEACHE_OK;
}
static int circuit_eval_device_async(ieache_ctx *ctx, const ieache_cloudkey *key, const ieache_circuit *cc, const int32_t *inputs,
                                     int32_t *outputs, size_t n_expr)
{
    if (!ctx || !key || !cc || !inputs || !outputs) return fail(IEACHE_ERR_ARG, "null argument");
    if (n_expr == 0) return IEACHE_OK;
    ieache_circuit *c = const_cast<ieache_circuit *>(cc);
    int rc = circuit_upload(ctx, c);
    if (rc) return rc;
    const Circuit &C = c->c;
    const int n = key->p.n;
    /* instances per pass, bounded by the wire memory budget (8 GiB) and the scratch for one level */
    const size_t bytes_per_inst = (size_t)C.n_slots * kLweStride * 4;
    size_t per = std::max<size_t>(1, std::min<size_t>(n_expr, (8ull << 30) / bytes_per_inst));
    per = std::max<size_t>(1, std::min<size_t>(per, (size_t)(1u << 20) / std::max<uint32_t>(1, C.max_width)));
    if ((rc = ensure(ctx, &ctx->d_wires, &ctx->wires_cap, per * C.n_slots * kLweStride))) return rc;
    if ((rc = ensure(ctx, &ctx->d_ext, &ctx->ext_cap, per * C.max_width * kExtStride))) return rc;
    for (size_t off = 0; off < n_expr; off += per) {
        const int m = (int)std::min(per, n_expr - off);
        CU(launch_circuit_scatter_inputs(ctx->d_wires, inputs + off * C.n_inputs * kLweStride, m, C.n_inputs, C.n_slots, n, key->dp.mu, ctx->stream));
        ctx->launches++;
        for (size_t L = 0; L < C.levels.size(); L++) {
            GateAddr ga{};
            ga.tmpl = c->d_tmpl + c->level_off[L];
            ga.ntempl = (int)C.levels[L].tmpl.size();
            ga.n_inst = m; ga.inst_samples = C.n_slots; ga.stride = kLweStride;
            if ((rc = run_br(ctx, key, ga, ctx->d_wires, ctx->d_wires, 0))) return rc;
            if ((rc = run_ks(ctx, key, ga, ctx->d_wires, 0, 0))) return rc;
        }
        CU(launch_circuit_gather_outputs(outputs + off * C.outputs.size() * kLweStride, ctx->d_wires, c->d_out_slots, m,
                                         (int)C.outputs.size(), C.n_slots, n, ctx->stream));
        ctx->launches++;
    }
    return IEACHE_OK;
}

extern "C" int ieache_circuit_eval(ieache_ctx *ctx, const ieache_cloudkey *key, const ieache_circuit *c, const int32_t *inputs,
                                   int32_t *outputs, size_t n_expr)
{
    if (!ctx || !key || !c || !inputs || !outputs) return fail(IEACHE_ERR_ARG, "null argument");
    if (n_expr == 0) return IEACHE_OK;
    CU(cudaSetDevice(ctx->device));
    const int n = key->p.n;
    const size_t nin = (size_t)c->c.n_inputs * n_expr, nout = c->c.outputs.size() * n_expr;
    int rc;
    if ((rc = ensure(ctx, &ctx->d_stage[0], &ctx->stage_cap[0], nin * kLweStride))) return rc;
    if ((rc = ensure(ctx, &ctx->d_stage[3], &ctx->stage_cap[3], nout * kLweStride))) return rc;
    if ((rc = ieache_samples_to_device(ctx, ctx->d_stage[0], inputs, nin, n))) return rc;
    if ((rc = ieache_circuit_eval_device(ctx, key, c, ctx->d_stage[0], ctx->d_stage[3], n_expr))) return rc;
    return ieache_samples_to_host(ctx, outputs, ctx->d_stage[3], nout, n);
}

/* ------------------------------------------------------------------ Cloud node process contract */
static int32_t dec32(const HostKeySet &k, const int32_t *blk)
{
    int32_t bits[32], v = 0;
    sym_decrypt_bits(k, blk, 32, bits);
    for (int i = 0; i < 32; i++) v |= bits[i] << i;
    return v;
}
static void enc32(const HostKeySet &k, int32_t v, int32_t *blk)
{
    int32_t bits[32];
    for (int i = 0; i < 32; i++) bits[i] = (v >> i) & 1;
    sym_encrypt_bits(k, bits, 32, blk);
}

/* ---- session: keys loaded once, any number of operators ---------------------------
 * The reference starts ./cloud once per operator and each start re-parses cloud.key (114 MB) and rebuilds
 * bkFFT (Cloud/cloud.c:656-663); intermediate results travel through answer.data -> cloud.data on disk
 * (Cloud/dragonfly_cipher_cloud.py:1306-1327).  A session keeps the device key and the nbit key, evaluates
 * batches of requests (grouped by circuit) in one levelised pass, and chains operators in memory. */
struct ieache_session {
    ieache_ctx *ctx = nullptr;
    ieache_cloudkey *key = nullptr;
    bool owns_key = false;
    HostKeySet nbit;
    std::map<std::pair<int, int>, ieache_circuit *> circuits;
    size_t pass = 256;   /* requests evaluated together by the batch calls (ieache_session_set_pass) */
};

extern "C" void ieache_session_close(ieache_session *s)
{
    if (!s) return;
    for (auto &kv : s->circuits) ieache_circuit_destroy(kv.second);
    if (s->owns_key) ieache_cloudkey_destroy(s->key);
    delete s;
}
extern "C" int ieache_session_open_keys(ieache_ctx *ctx, ieache_cloudkey *key, const int32_t *nbit_lwe_key, ieache_session **out)
{
    if (!ctx || !key || !nbit_lwe_key || !out) return fail(IEACHE_ERR_ARG, "null argument");
    ieache_session *s = new ieache_session();
    s->ctx = ctx; s->key = key; s->owns_key = false;
    s->nbit.p = key->p; s->nbit.has_secret = true;
    s->nbit.lwe_key.assign(nbit_lwe_key, nbit_lwe_key + key->p.n);
    *out = s;
    return IEACHE_OK;
}
extern "C" int ieache_session_open(ieache_ctx *ctx, const char *cloud_key_path, const char *nbit_key_path, ieache_session **out)
{
    if (!ctx || !cloud_key_path || !nbit_key_path || !out) return fail(IEACHE_ERR_ARG, "null argument");
    std::unique_ptr<ieache_session, void (*)(ieache_session *)> s(new ieache_session(), ieache_session_close);
    s->ctx = ctx;
    int rc = ieache_cloudkey_load_file(ctx, cloud_key_path, &s->key);          /* cloud.c:656-658 */
    if (rc) return rc;
    s->owns_key = true;
    std::string msg;
    if ((rc = read_keyset(nbit_key_path, s->nbit, false, msg))) return fail(rc, "%s", msg.c_str());   /* cloud.c:661-663 */
    if (!s->nbit.has_secret) return fail(IEACHE_ERR_FORMAT, "nbit.key holds no secret key");
    if (s->nbit.p.n != s->key->p.n) return fail(IEACHE_ERR_UNSUPPORTED, "nbit.key and cloud.key use different n (%d vs %d)", s->nbit.p.n, s->key->p.n);
    *out = s.release();
    return IEACHE_OK;
}
extern "C" int ieache_session_params(const ieache_session *s, ieache_params *out)
{
    if (!s || !out) return fail(IEACHE_ERR_ARG, "null argument");
    *out = s->key->p;
    return IEACHE_OK;
}

/* ---- operand batches resident on the device -------------------------------------------------------------------
 * The value part of `count` operand (or answer) blocks: 8 value chunks, least significant first, and the carry block
 * — 9 blocks of 32 samples each, device stride.  The two metadata blocks (sign code, width) are encrypted under
 * nbit.key, which the Cloud holds and cloud.c decrypts on the host anyway (cloud.c:709-796): they travel as plain
 * integers next to the device array and are only encrypted again when an answer leaves the engine. */
constexpr int kValBlocks = 9;
constexpr size_t kBlockWords = (size_t)32 * kLweStride;
struct DevValues {
    DevBuf buf;
    size_t count = 0;
    std::vector<int32_t> sign, width;
    std::vector<int32_t> exit_code;      /* per element: 0, or 126 once a 256-bit multiply was refused (cloud.c:860-864) */
    std::vector<uint8_t> computed;       /* per element: a circuit produced the value blocks (352-sample answer) */
    int32_t *d() const { return buf.as<int32_t>(); }
    int32_t *block(size_t e, int blk) const { return d() + (e * kValBlocks + blk) * kBlockWords; }
    int alloc(size_t n)
    {
        count = n;
        sign.assign(n, 0); width.assign(n, 0); exit_code.assign(n, 0); computed.assign(n, 1);
        if (n == 0) return IEACHE_OK;
        CU(buf.alloc(n * kValBlocks * kBlockWords * sizeof(int32_t)));
        return IEACHE_OK;
    }
};

/* pinned staging: two slots so that slot k+1 is packed and copied while slot k is in use */
struct PinnedSlots {
    int32_t *h[2] = {nullptr, nullptr};
    size_t words = 0;
    ~PinnedSlots() { for (int i = 0; i < 2; i++) if (h[i]) cudaFreeHost(h[i]); }
    int ensure(size_t w)
    {
        if (w <= words) return IEACHE_OK;
        for (int i = 0; i < 2; i++) { if (h[i]) cudaFreeHost(h[i]); h[i] = nullptr; }
        words = 0;
        for (int i = 0; i < 2; i++) CU(cudaHostAlloc((void **)&h[i], w * sizeof(int32_t), cudaHostAllocDefault));
        words = w;
        return IEACHE_OK;
    }
};
constexpr size_t kSessionChunk = 64; /* operands per staging copy: bounds pinned memory to 2 slots x 64 x 9 x 32 x (n + 1) x 4 B
                                        = 93 MB at n = 630 whatever the pass size */

/* Requests evaluated together by compute_batch / eval_postfix / compute_dirs.  A pass is what fills the GPU: a 32-bit add
 * has one or two gates per level, so 256 requests give levels of a few hundred gates and 2 048 reach the persistent
 * kernel; host memory of compute_dirs is 5.3 MB and device memory 2.2 MB per request of a pass. */
extern "C" int ieache_session_set_pass(ieache_session *s, size_t requests, size_t *old_value)
{
    if (!s) return fail(IEACHE_ERR_ARG, "null session");
    if (old_value) *old_value = s->pass;
    if (requests < 1 || requests > 65536) return fail(IEACHE_ERR_ARG, "pass size must be 1..65536 requests");
    s->pass = requests;
    return IEACHE_OK;
}

/* host operand blocks (352 samples of n + 1 words; element e at base + e * stride words) -> device values + metadata */
static int upload_values(ieache_session *s, size_t count, const int32_t *base, size_t stride_words, DevValues &out, PinnedSlots &pin)
{
    ieache_ctx *ctx = s->ctx;
    const int n = s->key->p.n;
    const size_t w = n + 1, B = 32 * w, vwords = kValBlocks * B;
    int rc = out.alloc(count);
    if (rc) return rc;
    if ((rc = pin.ensure(kSessionChunk * vwords))) return rc;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; i++) CU(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    size_t chunk = 0;
    for (size_t off = 0; off < count; off += kSessionChunk, chunk++) {
        const size_t m = std::min(kSessionChunk, count - off);
        const int slot = (int)(chunk & 1);
        if (chunk >= 2) CU(cudaEventSynchronize(ev[slot]));                       /* the copy out of this slot two chunks ago is done */
        for (size_t e = 0; e < m; e++) {
            const int32_t *blk = base + (off + e) * stride_words;
            out.sign[off + e] = dec32(s->nbit, blk);                              /* cloud.c:780-796 */
            out.width[off + e] = dec32(s->nbit, blk + B);                         /* cloud.c:709-746 */
            memcpy(pin.h[slot] + e * vwords, blk + 2 * B, vwords * sizeof(int32_t));
        }
        CU(cudaMemcpy2DAsync(out.block(off, 0), kLweStride * 4, pin.h[slot], w * 4, w * 4, m * kValBlocks * 32, cudaMemcpyHostToDevice, ctx->h2d_stream));
        CU(cudaEventRecord(ev[slot], ctx->h2d_stream));
        ctx->h2d_value_copies += m;
    }
    CU(cudaStreamSynchronize(ctx->h2d_stream));
    for (int i = 0; i < 2; i++) cudaEventDestroy(ev[i]);
    return IEACHE_OK;
}

/* device values + metadata -> host answer blocks: sign[32] (nbit key) || width[32] (nbit key) || 8 result blocks || carry
 * (cloud.c:812-855, 899-916).  answer_counts[e] = 352, or 64 when nothing was computed for the element. */
static int download_values(ieache_session *s, const DevValues &v, int32_t *answers, size_t stride_words, size_t *answer_counts, PinnedSlots &pin)
{
    ieache_ctx *ctx = s->ctx;
    const int n = s->key->p.n;
    const size_t w = n + 1, B = 32 * w, vwords = kValBlocks * B;
    int rc;
    if ((rc = pin.ensure(kSessionChunk * vwords))) return rc;
    CU(cudaStreamSynchronize(ctx->stream));
    for (size_t off = 0; off < v.count; off += kSessionChunk) {
        const size_t m = std::min(kSessionChunk, v.count - off);
        CU(cudaMemcpy2DAsync(pin.h[0], w * 4, v.block(off, 0), kLweStride * 4, w * 4, m * kValBlocks * 32, cudaMemcpyDeviceToHost, ctx->d2h_stream));
        CU(cudaStreamSynchronize(ctx->d2h_stream));
        ctx->d2h_value_copies += m;
        for (size_t e = 0; e < m; e++) {
            int32_t *ans = answers + (off + e) * stride_words;
            enc32(s->nbit, v.sign[off + e], ans);
            enc32(s->nbit, v.width[off + e], ans + B);
            if (v.computed[off + e]) memcpy(ans + 2 * B, pin.h[0] + e * vwords, vwords * sizeof(int32_t));
            if (answer_counts) answer_counts[off + e] = v.computed[off + e] ? 352 : 64;
        }
    }
    return IEACHE_OK;
}

struct Request { int kind = 0, width = 0; bool swap = false; };

/* One operator over `count` element pairs, everything on the device.  Metadata handling is Cloud/cloud.c:780-864 per
 * element; elements that need the same circuit are evaluated together, level by level.  R receives the answers
 * (value blocks on the device, metadata in plain). */
static int apply_operator(ieache_session *s, const int32_t *ops, const DevValues &A, const DevValues &Bv, DevValues &R, double *seconds)
{
    ieache_ctx *ctx = s->ctx;
    const size_t count = A.count;
    int rc = R.alloc(count);
    if (rc) return rc;
    std::vector<Request> req(count);
    std::map<std::pair<int, int>, std::vector<size_t>> groups; /* (kind | swap<<8, width) -> element indices */
    for (size_t i = 0; i < count; i++) {
        const int int_op = ops[i];
        const int32_t int_bit1 = A.width[i], int_bit2 = Bv.width[i];
        int32_t n1 = A.sign[i];
        const int32_t n2 = Bv.sign[i];
        if (n1 == 2) n1 = 1;                                                                          /* cloud.c:788-789 */
        const int32_t int_negative = n1 + n2;
        /* cloud.c:812-826: 1 -> 1, 2 -> 2, 3 -> 4, anything else (an operand that already carries code 4 from an earlier
         * operator of a chain) -> 0 */
        R.sign[i] = int_negative == 1 ? 1 : int_negative == 2 ? 2 : int_negative == 3 ? 4 : 0;
        const int32_t int_bit = std::max(int_bit1, int_bit2);
        R.width[i] = int_op == 4 ? int_bit * 2 : int_bit;                                             /* cloud.c:833-855 */
        R.exit_code[i] = std::max(A.exit_code[i], Bv.exit_code[i]);
        R.computed[i] = 0;
        Request &r = req[i];
        r.width = int_bit;
        if (R.exit_code[i]) continue;                                                                 /* the chain stopped earlier */
        if (int_op == 4 && int_bit >= 256) { R.exit_code[i] = 126; continue; }                        /* cloud.c:860-864 */
        if ((int_op == 1 && int_negative != 1 && int_negative != 2) || (int_op == 2 && (int_negative == 1 || int_negative == 2)))
            r.kind = IEACHE_CIRC_ADD;                                                                 /* cloud.c:870 */
        else if (int_op == 2 || (int_op == 1 && (int_negative == 1 || int_negative == 2))) {
            r.kind = IEACHE_CIRC_SUB;                                                                 /* cloud.c:1194 */
            r.swap = !((int_op == 2 && int_negative == 0) || (int_op == 1 && int_negative == 2));     /* cloud.c:1196,1809 */
        } else if (int_op == 4) r.kind = IEACHE_CIRC_MUL;                                             /* cloud.c:2368 */
        if (!r.kind) continue;
        const bool ok_width = (r.kind == IEACHE_CIRC_MUL) ? (int_bit == 32 || int_bit == 64 || int_bit == 128)
                                                          : (int_bit == 32 || int_bit == 64 || int_bit == 128 || int_bit == 256);
        if (!ok_width) { r.kind = 0; continue; }          /* the reference computes nothing for other widths */
        groups[{r.kind | (r.swap ? 256 : 0), int_bit}].push_back(i);
    }
    double secs = 0;
    for (auto &kv : groups) {
        const int kind = kv.first.first & 255, width = kv.first.second;
        const bool swap_ops = (kv.first.first & 256) != 0;
        const std::vector<size_t> &idx = kv.second;
        ieache_circuit *&circ = s->circuits[{kind, width}];
        if (!circ && (rc = ieache_circuit_build(kind, width, &circ))) return rc;
        const int nc = width / 32;
        const size_t nin = circ->c.n_inputs / 32, nout = circ->c.outputs.size() / 32, m = idx.size();   /* in blocks of 32 samples */
        DevBuf din, dout, dptr;
        CU(din.alloc(m * nin * kBlockWords * 4));
        CU(dout.alloc(m * nout * kBlockWords * 4));
        /* assemble the circuit inputs on the device: [first operand chunks][second operand chunks][operand 1's carry block] */
        std::vector<const void *> src;
        std::vector<void *> dst;
        for (size_t e = 0; e < m; e++) {
            const size_t i = idx[e];
            const DevValues &X = swap_ops ? Bv : A, &Y = swap_ops ? A : Bv;
            int32_t *in = din.as<int32_t>() + e * nin * kBlockWords;
            for (int c = 0; c < nc; c++) { src.push_back(X.block(i, c)); dst.push_back(in + (size_t)c * kBlockWords); }
            for (int c = 0; c < nc; c++) { src.push_back(Y.block(i, c)); dst.push_back(in + (size_t)(nc + c) * kBlockWords); }
            src.push_back(A.block(i, 8)); dst.push_back(in + (size_t)2 * nc * kBlockWords);          /* ciphertextcarry1 */
        }
        /* and where the results go: result blocks, then copies of operand 1's carry block as padding and as block 10 (cloud.c:899-916) */
        const size_t n_in_copies = src.size();
        for (size_t e = 0; e < m; e++) {
            const size_t i = idx[e];
            for (size_t q = 0; q < 8; q++) {
                src.push_back(q < nout ? (const void *)(dout.as<int32_t>() + (e * nout + q) * kBlockWords) : (const void *)A.block(i, 8));
                dst.push_back(R.block(i, (int)q));
            }
            src.push_back(A.block(i, 8)); dst.push_back(R.block(i, 8));
            R.computed[i] = 1;
        }
        CU(dptr.alloc(src.size() * 2 * sizeof(void *)));
        const void **d_src = dptr.as<const void *>();
        void **d_dst = reinterpret_cast<void **>(dptr.as<void *>() + src.size());
        CU(cudaMemcpyAsync(d_src, src.data(), src.size() * sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(d_dst, dst.data(), dst.size() * sizeof(void *), cudaMemcpyHostToDevice, ctx->stream));
        CU(launch_copy_blocks(d_src, d_dst, (int)n_in_copies, ctx->stream));
        ctx->launches++;
        const auto t0 = std::chrono::steady_clock::now();
        if ((rc = circuit_eval_device_async(ctx, s->key, circ, din.as<int32_t>(), dout.as<int32_t>(), m))) return rc;
        CU(launch_copy_blocks(d_src + n_in_copies, d_dst + n_in_copies, (int)(src.size() - n_in_copies), ctx->stream));
        ctx->launches++;
        CU(cudaStreamSynchronize(ctx->stream));        /* din / dout / dptr are released on leaving the scope */
        secs += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (seconds) *seconds += secs;
    return IEACHE_OK;
}

/* `count` requests (operator + two 352-sample client blocks each) -> `count` answer blocks, in passes of s->pass
 * requests through pinned staging slots (bounded pinned and device memory whatever `count` is).
 * exit_codes[i] = 0 or 126; answer_counts[i] = 352 or 64 samples. */
extern "C" int ieache_session_compute_batch(ieache_session *s, size_t count, const int32_t *ops, const int32_t *operands1,
                                            const int32_t *operands2, int32_t *answers, int32_t *exit_codes, size_t *answer_counts,
                                            double *seconds)
{
    if (!s || !ops || !operands1 || !operands2 || !answers) return fail(IEACHE_ERR_ARG, "null argument");
    CU(cudaSetDevice(s->ctx->device));
    const size_t blk = 352 * (size_t)(s->key->p.n + 1);
    PinnedSlots pin;
    double secs = 0;
    for (size_t off = 0; off < count; off += s->pass) {
        const size_t m = std::min(s->pass, count - off);
        DevValues A, B, R;
        int rc;
        if ((rc = upload_values(s, m, operands1 + off * blk, blk, A, pin))) return rc;
        if ((rc = upload_values(s, m, operands2 + off * blk, blk, B, pin))) return rc;
        if ((rc = apply_operator(s, ops + off, A, B, R, &secs))) return rc;
        if ((rc = download_values(s, R, answers + off * blk, blk, answer_counts ? answer_counts + off : nullptr, pin))) return rc;
        if (exit_codes) for (size_t e = 0; e < m; e++) exit_codes[off + e] = R.exit_code[e];
    }
    if (seconds) *seconds = secs;
    return IEACHE_OK;
}

extern "C" int ieache_session_compute(ieache_session *s, int op, const int32_t *operand1, const int32_t *operand2, int32_t *answer,
                                      size_t *answer_count, double *seconds)
{
    int32_t code = 0;
    const int32_t op32 = op;
    int rc = ieache_session_compute_batch(s, 1, &op32, operand1, operand2, answer, &code, answer_count, seconds);
    return rc ? rc : code;
}

/* A whole postfix expression per instance, e.g. "AB*C+" (Output/output_dynamic.py builds it; the Cloud walks
 * it at Cloud/dragonfly_cipher_cloud.py:685-729).  Letters index the operand blocks (A = 0), operators are
 * + - * (opcodes 1, 2, 4).  Operand order follows the infix expression exactly as the reference's `flip`
 * logic does.  All instances share the postfix string; each operator is one batched evaluation over all instances of a
 * pass.  Operands go to the device once, intermediate results STAY there (only their sign code and width, which the
 * Cloud decrypts anyway, are host integers), and only the final answers come back: n_operands uploads and one download
 * per instance whatever the number of operators.  Returns 126 if any instance hit the 256-bit multiply abort. */
extern "C" int ieache_session_eval_postfix(ieache_session *s, const char *postfix, size_t n_expr, const int32_t *operands,
                                           int n_operands, int32_t *answers, size_t *answer_counts, double *seconds)
{
    if (!s || !postfix || !operands || !answers) return fail(IEACHE_ERR_ARG, "null argument");
    if (n_operands < 1 || n_operands > 26) return fail(IEACHE_ERR_ARG, "1..26 operands");
    CU(cudaSetDevice(s->ctx->device));
    const size_t blk = 352 * (size_t)(s->key->p.n + 1);
    /* validate the expression once */
    {
        int depth = 0;
        for (const char *c = postfix; *c; ++c) {
            if (*c == ' ') continue;
            if (*c >= 'A' && *c <= 'Z') {
                if (*c - 'A' >= n_operands) return fail(IEACHE_ERR_ARG, "postfix names operand %c but only %d were given", *c, n_operands);
                depth++;
            } else if (*c == '+' || *c == '-' || *c == '*') {
                if (depth < 2) return fail(IEACHE_ERR_ARG, "malformed postfix expression");
                depth--;
            } else return fail(IEACHE_ERR_ARG, "unknown token '%c' in postfix expression", *c);
        }
        if (depth < 1) return fail(IEACHE_ERR_ARG, "empty postfix expression");
    }
    PinnedSlots pin;
    double secs = 0;
    int last_code = 0;
    for (size_t off = 0; off < n_expr; off += s->pass) {
        const size_t m = std::min(s->pass, n_expr - off);
        std::vector<std::unique_ptr<DevValues>> loaded(n_operands);   /* operand k of this pass, uploaded on first use */
        std::vector<std::unique_ptr<DevValues>> temps;                /* intermediate results */
        std::vector<DevValues *> stack;
        int rc;
        for (const char *c = postfix; *c; ++c) {
            if (*c == ' ') continue;
            if (*c >= 'A' && *c <= 'Z') {
                const int k = *c - 'A';
                if (!loaded[k]) {
                    loaded[k].reset(new DevValues());
                    if ((rc = upload_values(s, m, operands + (off * n_operands + k) * blk, (size_t)n_operands * blk, *loaded[k], pin))) return rc;
                }
                stack.push_back(loaded[k].get());
                continue;
            }
            const int32_t op = *c == '+' ? 1 : *c == '-' ? 2 : 4;
            DevValues *b = stack.back(); stack.pop_back();
            DevValues *a = stack.back(); stack.pop_back();
            std::vector<int32_t> ops(m, op);
            temps.emplace_back(new DevValues());
            if ((rc = apply_operator(s, ops.data(), *a, *b, *temps.back(), &secs))) return rc;
            stack.push_back(temps.back().get());
        }
        DevValues *res = stack.back();
        if ((rc = download_values(s, *res, answers + off * blk, blk, answer_counts ? answer_counts + off : nullptr, pin))) return rc;
        for (size_t e = 0; e < m; e++) if (res->exit_code[e]) last_code = res->exit_code[e];
    }
    if (seconds) *seconds = secs;
    return last_code;
}

/* Batched ingest: `count` request directories, each holding what ./cloud reads (cloud.data =
 * two 352-record client blocks, operator.txt), evaluated as levelised batches with the session's keys; every
 * directory gets the answer.data (and, on multiply, the averagestandard.txt line) ./cloud would have written.
 * Directories are taken in passes of s->pass: files are parsed and written by a small pool of host threads, the
 * requests of a pass form one batch, and host memory stays bounded by the pass size whatever `count` is.  The files of
 * pass k + 1 are read and the answers of pass k - 1 written while pass k is on the GPU.  `seconds` is the circuit time of
 * the whole call.  exit_codes[i] = 0 / 126 like ./cloud, or a negative IEACHE_ERR_* when that directory could not be
 * read or written (the others still run). */
namespace {
struct DirPassIn { size_t off = 0, m = 0; std::vector<int32_t> o1, o2, ops; std::vector<int> io_err; };
struct DirPassOut { size_t off = 0, m = 0; std::vector<int32_t> answers, codes, ops; std::vector<size_t> counts; std::vector<int> io_err; double secs = 0; };
template <class Fn> void dir_pool(size_t m, Fn &&fn)
{
    const unsigned hw = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> pool;
    std::atomic<size_t> next{0};
    for (unsigned t = 0; t < std::min<size_t>(hw, m); t++)
        pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < m;) fn(i); });
    for (auto &th : pool) th.join();
}
} // namespace
extern "C" int ieache_session_compute_dirs(ieache_session *s, size_t count, const char *const *dirs, int32_t *exit_codes, double *seconds)
{
    if (!s || !dirs || !exit_codes) return fail(IEACHE_ERR_ARG, "null argument");
    if (seconds) *seconds = 0;
    if (count == 0) return IEACHE_OK;
    const int n = s->key->p.n;
    const size_t w = n + 1, blk = 352 * w, pass = s->pass;
    const double var = s->key->p.ks_stdev * s->key->p.ks_stdev;
    auto read_pass = [&](DirPassIn &in, size_t off) {
        in.off = off; in.m = std::min(pass, count - off);
        in.o1.resize(in.m * blk); in.o2.resize(in.m * blk); in.ops.assign(in.m, 0); in.io_err.assign(in.m, 0);
        dir_pool(in.m, [&](size_t i) {
            const std::string d(dirs[off + i] ? dirs[off + i] : "");
            std::vector<int32_t> data(704 * w);
            FILE *f = fopen((d + "/cloud.data").c_str(), "rb");
            if (!f) { in.io_err[i] = IEACHE_ERR_IO; return; }
            const int rc = read_samples(f, n, data.data(), 704);
            fclose(f);
            if (rc) { in.io_err[i] = rc; return; }
            memcpy(&in.o1[i * blk], data.data(), blk * 4);
            memcpy(&in.o2[i * blk], data.data() + blk, blk * 4);
            int op = 0;
            f = fopen((d + "/operator.txt").c_str(), "r");
            if (!f) { in.io_err[i] = IEACHE_ERR_IO; return; }
            if (fscanf(f, "%d", &op) != 1) op = 0;
            fclose(f);
            in.ops[i] = op;
        });
        /* unreadable directories are left out of the batch (operator 0 selects no circuit, cloud.c computes nothing) */
        for (size_t i = 0; i < in.m; i++) if (in.io_err[i]) in.ops[i] = 0;
    };
    auto write_pass = [&](DirPassOut &out) {
        dir_pool(out.m, [&](size_t i) {
            if (out.io_err[i]) { exit_codes[out.off + i] = out.io_err[i]; return; }
            exit_codes[out.off + i] = out.codes[i];
            const std::string d(dirs[out.off + i]);
            if (out.counts[i] == 352 && out.ops[i] == 4) {                           /* cloud.c:2468-2471 */
                FILE *t = fopen((d + "/averagestandard.txt").c_str(), "a");
                if (t) { fprintf(t, "%lf\n", out.secs); fclose(t); }
            }
            FILE *f = fopen((d + "/answer.data").c_str(), "wb");
            if (!f) { exit_codes[out.off + i] = IEACHE_ERR_IO; return; }
            if (write_samples(f, n, &out.answers[i * blk], out.counts[i], var)) exit_codes[out.off + i] = IEACHE_ERR_IO;
            if (fclose(f) != 0) exit_codes[out.off + i] = IEACHE_ERR_IO;
        });
    };
    DirPassIn in[2];
    DirPassOut out[2];
    std::thread reader, writer;
    auto join = [](std::thread &t) { if (t.joinable()) t.join(); };
    double total = 0;
    int rc = IEACHE_OK;
    read_pass(in[0], 0);
    size_t k = 0;
    for (size_t off = 0; off < count; off += pass, k++) {
        DirPassIn &cur = in[k & 1];
        DirPassOut &res = out[k & 1];
        if (off + pass < count) reader = std::thread([&, off] { read_pass(in[(k + 1) & 1], off + pass); });
        res.off = cur.off; res.m = cur.m; res.ops = cur.ops; res.io_err = cur.io_err;
        res.answers.resize(cur.m * blk); res.codes.assign(cur.m, 0); res.counts.assign(cur.m, 0); res.secs = 0;
        rc = ieache_session_compute_batch(s, cur.m, cur.ops.data(), cur.o1.data(), cur.o2.data(), res.answers.data(), res.codes.data(),
                                          res.counts.data(), &res.secs);
        join(writer);                 /* the previous pass's answers are on disk before its buffers are reused two passes on */
        join(reader);
        if (rc) break;
        total += res.secs;
        writer = std::thread([&write_pass, &res] { write_pass(res); });
    }
    join(reader); join(writer);
    if (rc) return rc;
    if (seconds) *seconds = total;
    return IEACHE_OK;
}

extern "C" int ieache_cloud_run(ieache_ctx *ctx, const char *dir, double *seconds)
{
    if (!ctx || !dir) return fail(IEACHE_ERR_ARG, "null argument");
    const std::string d(dir);
    int rc;
    ieache_session *sess = nullptr;
    if ((rc = ieache_session_open(ctx, (d + "/cloud.key").c_str(), (d + "/nbit.key").c_str(), &sess))) return rc;
    std::unique_ptr<ieache_session, void (*)(ieache_session *)> guard(sess, ieache_session_close);
    const int n = sess->key->p.n;
    const size_t w = n + 1;
    /* cloud.c:703-766: two client blocks of 11 x 32 samples */
    std::vector<int32_t> data(704 * w);
    {
        FILE *f = fopen((d + "/cloud.data").c_str(), "rb");
        if (!f) return fail(IEACHE_ERR_IO, "cannot open %s/cloud.data", dir);
        rc = read_samples(f, n, data.data(), 704);
        fclose(f);
        if (rc) return fail(rc, "cloud.data: short read (expected 704 samples of %zu bytes)", sample_record_bytes(n));
    }
    int int_op = 0;                                                             /* cloud.c:770-773 */
    {
        FILE *f = fopen((d + "/operator.txt").c_str(), "r");
        if (!f) return fail(IEACHE_ERR_IO, "cannot open %s/operator.txt", dir);
        if (fscanf(f, "%d", &int_op) != 1) int_op = 0;
        fclose(f);
    }
    std::vector<int32_t> answer(352 * w);
    size_t out_count = 64;
    double secs = 0;
    const int exit_code = ieache_session_compute(sess, int_op, &data[0], &data[352 * w], answer.data(), &out_count, &secs);
    if (exit_code < 0) return exit_code;
    if (out_count == 352) {
        printf("Computation Time: %lf[sec]\n", secs);                           /* cloud.c:896 */
        if (int_op == 4) {                                                       /* cloud.c:2468-2471 */
            FILE *t = fopen((d + "/averagestandard.txt").c_str(), "a");
            if (t) { fprintf(t, "%lf\n", secs); fclose(t); }
        }
    }
    if (seconds) *seconds = secs;
    FILE *f = fopen((d + "/answer.data").c_str(), "wb");
    if (!f) return fail(IEACHE_ERR_IO, "cannot write %s/answer.data", dir);
    rc = write_samples(f, n, answer.data(), out_count, sess->key->p.ks_stdev * sess->key->p.ks_stdev);
    fclose(f);
    if (rc) return fail(rc, "short write on answer.data");
    return exit_code;
}

/* ------------------------------------------------------------------ the callers on either side of the path
 * (SURVEY.md §8 f-2, f-3): Keygen/keygen.c, Client1/alice.c, Output/verif.c on the reference's files. */

/* Keygen/keygen.c:22-51: two independent secret key sets (main key: seed {314,1592,657}; nbit key: seed
 * {314,1592,888} in the reference); writes secret.key, cloud.key, nbit.key.  Keys are generated on the GPU. */
extern "C" int ieache_keygen_files(ieache_ctx *ctx, const char *dir, const ieache_params *p, uint64_t seed_key, uint64_t seed_nbit)
{
    if (!ctx || !dir) return fail(IEACHE_ERR_ARG, "null argument");
    ieache_params def{};
    if (!p) {
        def.n = 630; def.N = 1024; def.k = 1; def.bk_l = 3; def.bk_Bgbit = 7; def.ks_t = 8; def.ks_basebit = 2;
        def.ks_stdev = 1.0 / 32768.0; def.bk_stdev = 1.0 / 33554432.0; def.max_stdev = 0.012467;
        p = &def;
    }
    int rc = check_params(p);
    if (rc) return rc;
    const std::string d(dir);
    const size_t bk_words = (size_t)p->n * 2 * p->bk_l * 2 * 1024, ks_words = (size_t)1024 * p->ks_t * (1u << p->ks_basebit) * (p->n + 1);
    for (int which = 0; which < 2; which++) {
        HostKeySet hk;
        hk.p = *p; hk.bk.resize(bk_words); hk.ksk.resize(ks_words);
        hk.lwe_key.resize(p->n); hk.tlwe_key.resize(1024); hk.has_secret = true;
        ieache_secretkey *sk = nullptr;
        ieache_cloudkey *ck = nullptr;
        if ((rc = ieache_keygen(ctx, p, which ? seed_nbit : seed_key, &sk, &ck, hk.bk.data(), hk.ksk.data()))) return rc;
        std::unique_ptr<ieache_secretkey, void (*)(ieache_secretkey *)> skg(sk, ieache_secretkey_destroy);
        std::unique_ptr<ieache_cloudkey, void (*)(ieache_cloudkey *)> ckg(ck, ieache_cloudkey_destroy);
        ieache_secretkey_export(sk, hk.lwe_key.data(), hk.tlwe_key.data());
        std::string msg;
        if (which == 0) {
            if ((rc = write_keyset((d + "/secret.key").c_str(), hk, true, msg))) return fail(rc, "%s", msg.c_str());   /* keygen.c:39-41 */
            if ((rc = write_keyset((d + "/cloud.key").c_str(), hk, false, msg))) return fail(rc, "%s", msg.c_str());   /* keygen.c:44-46 */
        } else if ((rc = write_keyset((d + "/nbit.key").c_str(), hk, true, msg))) return fail(rc, "%s", msg.c_str());  /* keygen.c:49-51 */
    }
    return IEACHE_OK;
}

/* Client1/alice.c:116-189: sign code and width under the nbit key, 8 value chunks (least significant first)
 * and a zero carry block under the main key -> 352 samples appended to or written at out_path */
extern "C" int ieache_alice_encrypt(const char *dir, int32_t sign_code, int32_t width, const uint32_t *chunks, const char *out_path,
                                    int append)
{
    if (!dir || !chunks || !out_path) return fail(IEACHE_ERR_ARG, "null argument");
    const std::string d(dir);
    HostKeySet key, nbit;
    std::string msg;
    int rc;
    if ((rc = read_keyset((d + "/secret.key").c_str(), key, false, msg))) return fail(rc, "%s", msg.c_str());   /* alice.c:36-38 */
    if ((rc = read_keyset((d + "/nbit.key").c_str(), nbit, false, msg))) return fail(rc, "%s", msg.c_str());    /* alice.c:40-42 */
    if (!key.has_secret || !nbit.has_secret) return fail(IEACHE_ERR_FORMAT, "secret.key / nbit.key hold no secret key");
    const int n = key.p.n;
    if (nbit.p.n != n) return fail(IEACHE_ERR_UNSUPPORTED, "secret.key and nbit.key use different n");
    const size_t w = n + 1, B = 32 * w;
    std::vector<int32_t> blk(352 * w);
    enc32(nbit, sign_code, &blk[0]);
    enc32(nbit, width, &blk[B]);
    for (int c = 0; c < 8; c++) enc32(key, c < width / 32 ? (int32_t)chunks[c] : 0, &blk[(2 + c) * B]);   /* unused chunks: plaintext3 = 0 */
    enc32(key, 0, &blk[10 * B]);
    FILE *f = fopen(out_path, append ? "ab" : "wb");
    if (!f) return fail(IEACHE_ERR_IO, "cannot write %s", out_path);
    rc = write_samples(f, n, blk.data(), 352, key.p.ks_stdev * key.p.ks_stdev);
    fclose(f);
    return rc ? fail(rc, "short write on %s", out_path) : IEACHE_OK;
}

/* Client1/alice.c main(): values.txt (lines of 32 binary digits: sign code, bit count, chunks) -> cloud.data */
extern "C" int ieache_alice_run(const char *dir)
{
    if (!dir) return fail(IEACHE_ERR_ARG, "null argument");
    const std::string d(dir);
    FILE *f = fopen((d + "/values.txt").c_str(), "r");
    if (!f) return fail(IEACHE_ERR_IO, "cannot open %s/values.txt", dir);
    char line[128];
    std::vector<uint32_t> vals;
    while (fgets(line, sizeof line, f)) {
        uint32_t v = 0;
        int digits = 0;
        for (const char *c = line; *c == '0' || *c == '1'; ++c, ++digits) v = (v << 1) | (uint32_t)(*c - '0');
        if (digits) vals.push_back(v);
    }
    fclose(f);
    if (vals.size() < 3) return fail(IEACHE_ERR_FORMAT, "values.txt: expected sign code, bit count and chunks");
    uint32_t chunks[8] = {0};
    for (size_t i = 2; i < vals.size() && i < 10; i++) chunks[i - 2] = vals[i];
    return ieache_alice_encrypt(dir, (int32_t)vals[0], (int32_t)vals[1], chunks, (d + "/cloud.data").c_str(), 0);
}

static std::string u256_to_decimal(const uint32_t *limbs_le, int nlimbs, bool negative)
{
    std::vector<uint32_t> v(limbs_le, limbs_le + nlimbs);
    std::string out;
    bool nonzero = false;
    for (uint32_t x : v) nonzero |= x != 0;
    if (!nonzero) return "0";
    while (true) {
        bool any = false;
        uint64_t rem = 0;
        for (int i = nlimbs - 1; i >= 0; i--) {
            const uint64_t cur = (rem << 32) | v[i];
            v[i] = (uint32_t)(cur / 1000000000u);
            rem = cur % 1000000000u;
            any |= v[i] != 0;
        }
        char buf[16];
        snprintf(buf, sizeof buf, any ? "%09u" : "%u", (unsigned)rem);
        out = std::string(buf) + out;
        if (!any) break;
    }
    return negative ? "-" + out : out;
}

/* Output/verif.c: decrypt answer.data (sign code and width under the nbit key, width/32 chunks under the main
 * key, :46-95), reassemble chunks most significant first (:229) and apply the sign rules of each operator
 * (:120-173 add, :733-789 subtract, :1409-1435 multiply).  result receives the decimal string verif prints. */
extern "C" int ieache_verif_run(const char *dir, char *result, size_t result_cap, int32_t *sign_code, int32_t *width)
{
    if (!dir) return fail(IEACHE_ERR_ARG, "null argument");
    const std::string d(dir);
    HostKeySet key, nbit;
    std::string msg;
    int rc;
    if ((rc = read_keyset((d + "/secret.key").c_str(), key, false, msg))) return fail(rc, "%s", msg.c_str());   /* verif.c:22-24 */
    if ((rc = read_keyset((d + "/nbit.key").c_str(), nbit, false, msg))) return fail(rc, "%s", msg.c_str());    /* verif.c:27-29 */
    if (!key.has_secret || !nbit.has_secret) return fail(IEACHE_ERR_FORMAT, "secret.key / nbit.key hold no secret key");
    const int n = key.p.n;
    const size_t w = n + 1, B = 32 * w;
    int int_op = 0;
    {
        FILE *f = fopen((d + "/operator.txt").c_str(), "r");                     /* verif.c:66-69 */
        if (!f) return fail(IEACHE_ERR_IO, "cannot open %s/operator.txt", dir);
        if (fscanf(f, "%d", &int_op) != 1) int_op = 0;
        fclose(f);
    }
    std::vector<int32_t> ans(352 * w);
    FILE *f = fopen((d + "/answer.data").c_str(), "rb");
    if (!f) return fail(IEACHE_ERR_IO, "cannot open %s/answer.data", dir);
    rc = read_samples(f, n, ans.data(), 64);
    if (rc) { fclose(f); return fail(rc, "answer.data: short read"); }
    const int32_t code = dec32(nbit, &ans[0]), bits = dec32(nbit, &ans[B]);
    if (sign_code) *sign_code = code;
    if (width) *width = bits;
    const int nchunks = bits / 32;
    if (nchunks < 1 || nchunks > 8) { fclose(f); return fail(IEACHE_ERR_FORMAT, "answer.data: implausible width %d", bits); }
    rc = read_samples(f, n, &ans[2 * B], (size_t)nchunks * 32);
    fclose(f);
    if (rc) return fail(rc, "answer.data holds no result chunks (abort path of Cloud/cloud.c:860-864?)");
    uint32_t limbs[9] = {0};
    for (int c = 0; c < nchunks; c++) limbs[c] = (uint32_t)dec32(key, &ans[(2 + c) * B]);
    /* two's-complement reinterpretation happens only for 32-character strings (verif.c:148,747) */
    bool negative = false;
    const bool twos = bits == 32 && (limbs[0] >> 31);
    auto negate32 = [&]() { limbs[0] = (uint32_t)(-(int32_t)limbs[0]); negative = !negative; };
    if (int_op == 1) {
        if (code != 0 && code != 4 && twos) negate32();
        if (code == 4) negative = !negative;
    } else if (int_op == 2) {
        if (code != 2 && twos) negate32();
        if (code == 1) negative = !negative;
    } else if (int_op == 4) {
        negative = (code == 1 || code == 2);
    } else return fail(IEACHE_ERR_ARG, "operator.txt holds %d", int_op);
    const std::string dec = u256_to_decimal(limbs, 8, negative);
    if (result && result_cap) snprintf(result, result_cap, "%s", dec.c_str());
    return IEACHE_OK;
}
