/*
 * tfhe_compat.cpp — libtfhe's gate-level C API (include/tfhe/tfhe.h, tfhe_io.h) on top of the
 * engine.  These are the symbols Cloud/cloud.c imports.  libtfhe has no error
 * returns and aborts on failure; so does this layer, after printing the engine's message.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/ieache_b200.h"
#include "../../include/tfhe/tfhe.h"
#include "../../include/tfhe/tfhe_io.h"
#include "tfhe_io.h"

using namespace ieache;

namespace {
[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "ieache_b200 (libtfhe-compatible API): %s: %s\n", what, ieache_last_error());
    abort();
}
std::mutex g_mu;
ieache_ctx *g_ctx = nullptr;
ieache_ctx *ctx()
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) {
        const char *dev = getenv("IEACHE_DEVICE");
        if (ieache_ctx_create(dev ? atoi(dev) : 0, &g_ctx) != IEACHE_OK) die("no usable B200 (there is no CPU fallback)");
    }
    return g_ctx;
}

struct ParamBundle {
    TFheGateBootstrappingParameterSet set;
    LweParams lwe;
    TLweParams tlwe;
    TGswParams tgsw;
    ieache_params raw;
};
ParamBundle *make_params(const ieache_params &p)
{
    ParamBundle *b = new ParamBundle();
    b->raw = p;
    b->lwe = LweParams{p.n, p.ks_stdev, p.max_stdev};
    b->tlwe = TLweParams{p.N, p.k, p.bk_stdev, p.max_stdev, LweParams{p.N * p.k, p.bk_stdev, p.max_stdev}};
    b->tgsw.l = p.bk_l; b->tgsw.Bgbit = p.bk_Bgbit; b->tgsw.Bg = 1 << p.bk_Bgbit; b->tgsw.halfBg = b->tgsw.Bg / 2;
    b->tgsw.maskMod = b->tgsw.Bg - 1; b->tgsw.tlwe_params = &b->tlwe; b->tgsw.kpl = (p.k + 1) * p.bk_l; b->tgsw.h = nullptr;
    uint32_t off = 0;
    for (int i = 1; i <= p.bk_l; i++) off += (uint32_t)b->tgsw.halfBg << (32 - i * p.bk_Bgbit);
    b->tgsw.offset = off;
    b->set.ks_t = p.ks_t; b->set.ks_basebit = p.ks_basebit; b->set.in_out_params = &b->lwe; b->set.tgsw_params = &b->tgsw;
    return b;
}
/* the parameter set is the first member of the bundle, so a set pointer identifies its bundle */
const ParamBundle *bundle_of(const TFheGateBootstrappingParameterSet *s) { return reinterpret_cast<const ParamBundle *>(s); }

struct CloudWrap { TFheGateBootstrappingCloudKeySet pub; ParamBundle *params; ieache_cloudkey *key; };
struct SecretWrap { TFheGateBootstrappingSecretKeySet pub; ParamBundle *params; HostKeySet host; LweKey lwe; };

const CloudWrap *cw(const TFheGateBootstrappingCloudKeySet *k) { return reinterpret_cast<const CloudWrap *>(k); }
const SecretWrap *sw(const TFheGateBootstrappingSecretKeySet *k) { return reinterpret_cast<const SecretWrap *>(k); }

void pack(const LweSample *s, int n, int32_t *dst) { memcpy(dst, s->a, 4 * (size_t)n); dst[n] = s->b; }
void unpack(LweSample *s, int n, const int32_t *src) { memcpy(s->a, src, 4 * (size_t)n); s->b = src[n]; }

void gate(int op, LweSample *result, const LweSample *a, const LweSample *b, const LweSample *c, int32_t imm, int32_t count,
          const TFheGateBootstrappingCloudKeySet *bk)
{
    /* libtfhe code sometimes passes &secret_keyset->cloud.  Key sets read from a secret-key file by this layer carry no
     * bootstrapping key on the device (Cloud/cloud.c only decrypts with nbit.key), and that member is not a CloudWrap:
     * refuse it by name instead of reinterpreting whatever lies behind it */
    if (!bk || !bk->bkFFT) {
        fprintf(stderr, "ieache_b200 (libtfhe-compatible API): this key set holds no bootstrapping key (a secret key set's .cloud "
                        "member?); load cloud.key with new_tfheGateBootstrappingCloudKeySet_fromFile\n");
        abort();
    }
    const CloudWrap *w = cw(bk);
    if (reinterpret_cast<const LweBootstrappingKeyFFT *>(w->key) != bk->bkFFT) {
        fprintf(stderr, "ieache_b200 (libtfhe-compatible API): not a cloud key set made by this library\n");
        abort();
    }
    const int n = w->params->raw.n;
    const size_t rec = (size_t)n + 1;
    /* cloud.c calls gates from OpenMP sections on a shared key (cloud.c:27-41): the engine context's staging
     * buffers are not shared between concurrent calls, so calls are serialised here */
    static std::mutex gate_mu;
    std::lock_guard<std::mutex> lk(gate_mu);
    std::vector<int32_t> buf(4 * rec * count);
    int32_t *pa = buf.data(), *pb = pa + rec * count, *pc = pb + rec * count, *po = pc + rec * count;
    for (int32_t g = 0; g < count; g++) {
        if (a) pack(a + g, n, pa + g * rec);
        if (b) pack(b + g, n, pb + g * rec);
        if (c) pack(c + g, n, pc + g * rec);
    }
    if (ieache_gate_batch(ctx(), w->key, op, po, a ? pa : nullptr, b ? pb : nullptr, c ? pc : nullptr, imm, count) != IEACHE_OK)
        die("gate evaluation failed");
    for (int32_t g = 0; g < count; g++) unpack(result + g, n, po + g * rec);
}
} // namespace

extern "C" {

LweSample *new_LweSample_array(int32_t nbelems, const LweParams *params)
{
    /* one block: the LweSample records, then all mask vectors back to back */
    const size_t n = params->n;
    char *blk = (char *)malloc(sizeof(LweSample) * nbelems + sizeof(Torus32) * n * nbelems);
    if (!blk) return nullptr;
    LweSample *arr = (LweSample *)blk;
    Torus32 *masks = (Torus32 *)(blk + sizeof(LweSample) * nbelems);
    memset(masks, 0, sizeof(Torus32) * n * nbelems);
    for (int32_t i = 0; i < nbelems; i++) { arr[i].a = masks + (size_t)i * n; arr[i].b = 0; arr[i].current_variance = 0.0; }
    return arr;
}
void delete_LweSample_array(int32_t, LweSample *samples) { free(samples); }
LweSample *new_gate_bootstrapping_ciphertext_array(int32_t nbelems, const TFheGateBootstrappingParameterSet *params)
{
    return new_LweSample_array(nbelems, params->in_out_params);
}
void delete_gate_bootstrapping_ciphertext_array(int32_t, LweSample *samples) { free(samples); }
LweSample *new_gate_bootstrapping_ciphertext(const TFheGateBootstrappingParameterSet *params) { return new_LweSample_array(1, params->in_out_params); }
void delete_gate_bootstrapping_ciphertext(LweSample *sample) { free(sample); }

TFheGateBootstrappingParameterSet *new_default_gate_bootstrapping_parameters(int32_t minimum_lambda)
{
    /* tfhe master: lambda in (80,128] -> n=630, l=3, Bgbit=7 (Keygen/keygen.c:22-23) */
    if (minimum_lambda > 128) { fprintf(stderr, "ieache_b200: no parameter set for lambda=%d\n", minimum_lambda); abort(); }
    ieache_params p{};
    p.n = 630; p.N = 1024; p.k = 1; p.bk_l = 3; p.bk_Bgbit = 7; p.ks_t = 8; p.ks_basebit = 2;
    p.ks_stdev = 1.0 / 32768.0; p.bk_stdev = 1.0 / 33554432.0; p.max_stdev = 0.012467;
    return &make_params(p)->set;
}
void delete_gate_bootstrapping_parameters(TFheGateBootstrappingParameterSet *params) { delete const_cast<ParamBundle *>(bundle_of(params)); }

TFheGateBootstrappingCloudKeySet *new_tfheGateBootstrappingCloudKeySet_fromFile(FILE *f)
{
    HostKeySet hk;
    std::string msg;
    if (read_keyset_stream(f, hk, true, msg) != IEACHE_OK) { fprintf(stderr, "ieache_b200: cloud key: %s\n", msg.c_str()); abort(); }
    CloudWrap *w = new CloudWrap();
    w->params = make_params(hk.p);
    if (ieache_cloudkey_create(ctx(), &hk.p, hk.bk.data(), hk.ksk.data(), &w->key) != IEACHE_OK) die("cloud key upload failed");
    w->pub.params = &w->params->set;
    w->pub.bk = nullptr;
    w->pub.bkFFT = reinterpret_cast<const LweBootstrappingKeyFFT *>(w->key);
    return &w->pub;
}
void delete_gate_bootstrapping_cloud_keyset(TFheGateBootstrappingCloudKeySet *keyset)
{
    if (!keyset) return;
    CloudWrap *w = const_cast<CloudWrap *>(cw(keyset));
    ieache_cloudkey_destroy(w->key);
    delete w->params;
    delete w;
}
TFheGateBootstrappingSecretKeySet *new_tfheGateBootstrappingSecretKeySet_fromFile(FILE *f)
{
    SecretWrap *w = new SecretWrap();
    std::string msg;
    if (read_keyset_stream(f, w->host, false, msg) != IEACHE_OK || !w->host.has_secret) {
        fprintf(stderr, "ieache_b200: secret key: %s\n", msg.empty() ? "no secret part" : msg.c_str());
        abort();
    }
    w->params = make_params(w->host.p);
    w->lwe.params = &w->params->lwe; w->lwe.key = w->host.lwe_key.data();
    w->pub.params = &w->params->set; w->pub.lwe_key = &w->lwe; w->pub.tgsw_key = nullptr;
    w->pub.cloud.params = &w->params->set; w->pub.cloud.bk = nullptr; w->pub.cloud.bkFFT = nullptr;
    return &w->pub;
}
void delete_gate_bootstrapping_secret_keyset(TFheGateBootstrappingSecretKeySet *keyset)
{
    if (!keyset) return;
    SecretWrap *w = const_cast<SecretWrap *>(sw(keyset));
    delete w->params;
    delete w;
}

void import_gate_bootstrapping_ciphertext_fromFile(FILE *f, LweSample *sample, const TFheGateBootstrappingParameterSet *params)
{
    const int n = params->in_out_params->n;
    std::vector<int32_t> rec(n + 1);
    if (read_samples(f, n, rec.data(), 1) != IEACHE_OK) { fprintf(stderr, "ieache_b200: short read on ciphertext file\n"); abort(); }
    unpack(sample, n, rec.data());
}
void export_gate_bootstrapping_ciphertext_toFile(FILE *f, const LweSample *sample, const TFheGateBootstrappingParameterSet *params)
{
    const int n = params->in_out_params->n;
    std::vector<int32_t> rec(n + 1);
    pack(sample, n, rec.data());
    if (write_samples(f, n, rec.data(), 1, sample->current_variance) != IEACHE_OK) { fprintf(stderr, "ieache_b200: short write on ciphertext file\n"); abort(); }
}

void bootsSymEncrypt(LweSample *result, int32_t message, const TFheGateBootstrappingSecretKeySet *key)
{
    const SecretWrap *w = sw(key);
    const int n = w->host.p.n;
    std::vector<int32_t> rec(n + 1);
    int32_t bit = message ? 1 : 0;
    sym_encrypt_bits(w->host, &bit, 1, rec.data());
    unpack(result, n, rec.data());
    result->current_variance = w->host.p.ks_stdev * w->host.p.ks_stdev;
}
int32_t bootsSymDecrypt(const LweSample *sample, const TFheGateBootstrappingSecretKeySet *key)
{
    const SecretWrap *w = sw(key);
    const int n = w->host.p.n;
    std::vector<int32_t> rec(n + 1);
    pack(sample, n, rec.data());
    int32_t bit;
    sym_decrypt_bits(w->host, rec.data(), 1, &bit);
    return bit;
}

void bootsCONSTANT(LweSample *result, int32_t value, const TFheGateBootstrappingCloudKeySet *bk)
{
    /* free linear op (trivial sample); no GPU round trip, exactly (0, +-mu) as libtfhe */
    const int n = cw(bk)->params->raw.n;
    memset(result->a, 0, 4 * (size_t)n);
    result->b = value ? (1 << 29) : -(1 << 29);
    result->current_variance = 0.0;
}
void bootsNOT(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk)
{
    const int n = cw(bk)->params->raw.n;
    for (int i = 0; i < n; i++) result->a[i] = -ca->a[i];
    result->b = -ca->b;
    result->current_variance = ca->current_variance;
}
void bootsCOPY(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk)
{
    const int n = cw(bk)->params->raw.n;
    if (result != ca) { memmove(result->a, ca->a, 4 * (size_t)n); result->b = ca->b; result->current_variance = ca->current_variance; }
}
#define BIN_GATE(NAME, OP)                                                                                              \
    void NAME(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk)   \
    {                                                                                                                    \
        gate(OP, result, ca, cb, nullptr, 0, 1, bk);                                                                     \
    }
BIN_GATE(bootsNAND, IEACHE_OP_NAND)
BIN_GATE(bootsOR, IEACHE_OP_OR)
BIN_GATE(bootsAND, IEACHE_OP_AND)
BIN_GATE(bootsXOR, IEACHE_OP_XOR)
BIN_GATE(bootsXNOR, IEACHE_OP_XNOR)
BIN_GATE(bootsNOR, IEACHE_OP_NOR)
BIN_GATE(bootsANDNY, IEACHE_OP_ANDNY)
BIN_GATE(bootsANDYN, IEACHE_OP_ANDYN)
BIN_GATE(bootsORNY, IEACHE_OP_ORNY)
BIN_GATE(bootsORYN, IEACHE_OP_ORYN)
void bootsMUX(LweSample *result, const LweSample *a, const LweSample *b, const LweSample *c, const TFheGateBootstrappingCloudKeySet *bk)
{
    gate(IEACHE_OP_MUX, result, a, b, c, 0, 1, bk);
}
int ieache_boots_batch(int op, LweSample *result, const LweSample *ca, const LweSample *cb, const LweSample *cc, int32_t count,
                       const TFheGateBootstrappingCloudKeySet *bk)
{
    if (count <= 0) return 0;
    gate(op, result, ca, cb, cc, 0, count, bk);
    return 0;
}

} // extern "C"
