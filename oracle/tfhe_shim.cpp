/*
 * oracle/tfhe_shim.cpp — TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h).
 *
 * Serves the 17 libtfhe symbols the reference's Cloud/cloud.c imports with the CPU oracle's
 * gates, so that the UNMODIFIED /root/reference/Cloud/cloud.c can be compiled (oracle/Makefile, target
 * _ref/cloud_ref_oracle) and run on the same files as the product.  What this pins: the circuits, the dispatch, the
 * metadata handling and the file layout of the oracle's restatement (cloud_oracle.c) and of the product
 * (csrc/circuit.cpp, engine.cu) against the reference's own source, gate call by gate call.  What it cannot pin is
 * the gate itself: libtfhe is absent, the gates here are the oracle's (tfhe_oracle.c).
 *
 * The structs come from include/tfhe/tfhe.h, the header the product ships for the same call sites; nothing of the
 * product's code is linked here.
 */
#include <tfhe/tfhe.h>
#include <tfhe/tfhe_io.h>

#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <atomic>
#include <string>
#include <vector>

#include "tfhe_oracle.h"

namespace {

struct ShimKey {                        /* owns everything a key-set pointer of the reference reaches */
    OKeySet *ks = nullptr;
    LweParams lwe{};
    TLweParams tlwe{};
    TGswParams tgsw{};
    TFheGateBootstrappingParameterSet params{};
    TFheGateBootstrappingCloudKeySet cloud{};
    TFheGateBootstrappingSecretKeySet secret{};
};

std::vector<ShimKey *> &registry() { static std::vector<ShimKey *> r; return r; }

ShimKey *find_by_params(const TFheGateBootstrappingParameterSet *p)
{
    for (ShimKey *k : registry()) if (&k->params == p) return k;
    return nullptr;
}
ShimKey *find_cloud(const TFheGateBootstrappingCloudKeySet *c)
{
    for (ShimKey *k : registry()) if (&k->cloud == c || &k->secret.cloud == c) return k;
    fprintf(stderr, "tfhe_shim: unknown cloud key set\n");
    abort();
}
ShimKey *find_secret(const TFheGateBootstrappingSecretKeySet *s)
{
    for (ShimKey *k : registry()) if (&k->secret == s) return k;
    fprintf(stderr, "tfhe_shim: unknown secret key set\n");
    abort();
}

ShimKey *load(FILE *f)
{
    /* the oracle reads by path: recover it from the descriptor (Linux) */
    char link[64], path[4096];
    snprintf(link, sizeof link, "/proc/self/fd/%d", fileno(f));
    const ssize_t len = readlink(link, path, sizeof path - 1);
    if (len <= 0) { fprintf(stderr, "tfhe_shim: cannot resolve the key file's path\n"); abort(); }
    path[len] = 0;
    ShimKey *k = new ShimKey();
    k->ks = o_read_key(path);
    if (!k->ks) { fprintf(stderr, "tfhe_shim: cannot parse %s\n", path); abort(); }
    const OParams *p = o_keyset_params(k->ks);
    k->lwe = LweParams{p->n, p->ks_stdev, p->max_stdev};
    k->tlwe.N = p->N; k->tlwe.k = p->k; k->tlwe.alpha_min = p->bk_stdev; k->tlwe.alpha_max = p->max_stdev;
    k->tlwe.extracted_lweparams = LweParams{p->N * p->k, p->bk_stdev, p->max_stdev};
    k->tgsw.l = p->bk_l; k->tgsw.Bgbit = p->bk_Bgbit; k->tgsw.Bg = 1 << p->bk_Bgbit; k->tgsw.halfBg = k->tgsw.Bg / 2;
    k->tgsw.maskMod = (uint32_t)k->tgsw.Bg - 1; k->tgsw.tlwe_params = &k->tlwe; k->tgsw.kpl = (p->k + 1) * p->bk_l;
    k->params.ks_t = p->ks_t; k->params.ks_basebit = p->ks_basebit;
    k->params.in_out_params = &k->lwe; k->params.tgsw_params = &k->tgsw;
    k->cloud.params = &k->params;
    k->secret.params = &k->params;
    k->secret.cloud.params = &k->params;
    registry().push_back(k);
    return k;
}

void unload(ShimKey *k)
{
    auto &r = registry();
    for (size_t i = 0; i < r.size(); i++) if (r[i] == k) { r.erase(r.begin() + i); break; }
    o_keyset_free(k->ks);
    delete k;
}

std::vector<int32_t> flat(const LweSample *s, int n)
{
    std::vector<int32_t> v(n + 1);
    memcpy(v.data(), s->a, (size_t)n * 4);
    v[n] = s->b;
    return v;
}
void unflat(LweSample *s, const std::vector<int32_t> &v, int n, double var)
{
    memcpy(s->a, v.data(), (size_t)n * 4);
    s->b = v[n];
    s->current_variance = var;
}

void gate(int op, LweSample *result, const LweSample *ca, const LweSample *cb, int32_t imm, const TFheGateBootstrappingCloudKeySet *bk)
{
    ShimKey *k = find_cloud(bk);
    const int n = k->lwe.n;
    std::vector<int32_t> a, b, out(n + 1);
    if (ca) a = flat(ca, n);
    if (cb) b = flat(cb, n);
    o_gate(k->ks, op, out.data(), ca ? a.data() : nullptr, cb ? b.data() : nullptr, nullptr, imm);
    unflat(result, out, n, k->lwe.alpha_min * k->lwe.alpha_min);
}

std::atomic<uint64_t> g_enc_seed{0x5EED0001};
constexpr int32_t kLweSampleTypeId = 42; /* tfhe_oracle.c UID_LWE_SAMPLE */

} // namespace

extern "C" {

LweSample *new_LweSample_array(int32_t nbelems, const LweParams *params)
{
    /* one block: the structs, then the masks, so that `x + i` walks contiguous samples as in libtfhe */
    const int n = params->n;
    char *blk = (char *)calloc(1, sizeof(int64_t) + (size_t)nbelems * sizeof(LweSample) + (size_t)nbelems * n * sizeof(Torus32));
    if (!blk) abort();
    LweSample *arr = (LweSample *)(blk + sizeof(int64_t));
    Torus32 *masks = (Torus32 *)(arr + nbelems);
    for (int32_t i = 0; i < nbelems; i++) arr[i].a = masks + (size_t)i * n;
    return arr;
}
void delete_LweSample_array(int32_t, LweSample *samples) { if (samples) free((char *)samples - sizeof(int64_t)); }
LweSample *new_gate_bootstrapping_ciphertext_array(int32_t nbelems, const TFheGateBootstrappingParameterSet *params)
{
    return new_LweSample_array(nbelems, params->in_out_params);
}
void delete_gate_bootstrapping_ciphertext_array(int32_t nbelems, LweSample *samples) { delete_LweSample_array(nbelems, samples); }

TFheGateBootstrappingCloudKeySet *new_tfheGateBootstrappingCloudKeySet_fromFile(FILE *f) { return &load(f)->cloud; }
TFheGateBootstrappingSecretKeySet *new_tfheGateBootstrappingSecretKeySet_fromFile(FILE *f)
{
    ShimKey *k = load(f);
    if (!o_keyset_has_secret(k->ks)) { fprintf(stderr, "tfhe_shim: the file holds no secret key\n"); abort(); }
    return &k->secret;
}
void delete_gate_bootstrapping_cloud_keyset(TFheGateBootstrappingCloudKeySet *keyset) { if (keyset) unload(find_cloud(keyset)); }
void delete_gate_bootstrapping_secret_keyset(TFheGateBootstrappingSecretKeySet *keyset) { if (keyset) unload(find_secret(keyset)); }

void import_gate_bootstrapping_ciphertext_fromFile(FILE *f, LweSample *sample, const TFheGateBootstrappingParameterSet *params)
{
    const int n = params->in_out_params->n;
    int32_t id = 0;
    if (fread(&id, 4, 1, f) != 1 || fread(sample->a, 4, (size_t)n, f) != (size_t)n || fread(&sample->b, 4, 1, f) != 1 ||
        fread(&sample->current_variance, 8, 1, f) != 1) {
        fprintf(stderr, "tfhe_shim: short read on a ciphertext\n");
        abort(); /* libtfhe aborts too */
    }
}
void export_gate_bootstrapping_ciphertext_toFile(FILE *f, const LweSample *sample, const TFheGateBootstrappingParameterSet *params)
{
    const int n = params->in_out_params->n;
    const ShimKey *k = find_by_params(params);
    const double var = k ? k->lwe.alpha_min * k->lwe.alpha_min : sample->current_variance; /* as o_write_samples */
    fwrite(&kLweSampleTypeId, 4, 1, f);
    fwrite(sample->a, 4, (size_t)n, f);
    fwrite(&sample->b, 4, 1, f);
    fwrite(&var, 8, 1, f);
}

void bootsSymEncrypt(LweSample *result, int32_t message, const TFheGateBootstrappingSecretKeySet *key)
{
    ShimKey *k = find_secret(key);
    const int n = k->lwe.n;
    std::vector<int32_t> out(n + 1);
    const int32_t bit = message ? 1 : 0;
    o_sym_encrypt(k->ks, &bit, 1, out.data(), g_enc_seed.fetch_add(1));
    unflat(result, out, n, k->lwe.alpha_min * k->lwe.alpha_min);
}
int32_t bootsSymDecrypt(const LweSample *sample, const TFheGateBootstrappingSecretKeySet *key)
{
    ShimKey *k = find_secret(key);
    const std::vector<int32_t> v = flat(sample, k->lwe.n);
    int32_t bit = 0;
    o_sym_decrypt(k->ks, v.data(), 1, &bit);
    return bit;
}

void bootsCONSTANT(LweSample *result, int32_t value, const TFheGateBootstrappingCloudKeySet *bk) { gate(O_CONST, result, nullptr, nullptr, value, bk); }
void bootsNOT(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk) { gate(O_NOT, result, ca, nullptr, 0, bk); }
void bootsCOPY(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk) { gate(O_COPY, result, ca, nullptr, 0, bk); }
void bootsAND(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk) { gate(O_AND, result, ca, cb, 0, bk); }
void bootsXOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk) { gate(O_XOR, result, ca, cb, 0, bk); }

} // extern "C"
