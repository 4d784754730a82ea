/*
 * tfhe/tfhe.h — libtfhe-compatible gate-level API, served by the B200 engine.
 *
 * Drop-in for the header Cloud/cloud.c includes at cloud.c:5 (and tfhe_io.h at cloud.c:6):
 * the struct members cloud.c dereferences (keyset->params->in_out_params cloud.c:20, bk->params
 * cloud.c:666, nbitkey->params cloud.c:669) and every function it calls keep libtfhe's names,
 * argument order and ownership rules.  Link with -lieache_b200 instead of
 * -ltfhe-spqlios-fma (compile_c.py:63,65).
 *
 * Each boots* call is synchronous and runs one gate on the GPU; the fast path for whole
 * circuits is ieache_circuit_eval / ieache_cloud_run in ieache_b200.h.
 */
#ifndef IEACHE_TFHE_COMPAT_H
#define IEACHE_TFHE_COMPAT_H
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t Torus32;

typedef struct LweParams { int32_t n; double alpha_min; double alpha_max; } LweParams;
typedef struct LweSample { Torus32 *a; Torus32 b; double current_variance; } LweSample;
typedef struct LweKey { const LweParams *params; int32_t *key; } LweKey;
typedef struct TLweParams { int32_t N; int32_t k; double alpha_min; double alpha_max; LweParams extracted_lweparams; } TLweParams;
typedef struct TGswParams { int32_t l; int32_t Bgbit; int32_t Bg; int32_t halfBg; uint32_t maskMod; const TLweParams *tlwe_params; int32_t kpl; Torus32 *h; uint32_t offset; } TGswParams;
typedef struct TGswKey TGswKey;                         /* opaque */
typedef struct LweBootstrappingKey LweBootstrappingKey; /* opaque: coefficient-domain key is not kept on the host */
typedef struct LweBootstrappingKeyFFT LweBootstrappingKeyFFT; /* opaque: device-resident ieache_cloudkey */

typedef struct TFheGateBootstrappingParameterSet {
    int32_t ks_t;
    int32_t ks_basebit;
    const LweParams *in_out_params;
    const TGswParams *tgsw_params;
} TFheGateBootstrappingParameterSet;

typedef struct TFheGateBootstrappingCloudKeySet {
    const TFheGateBootstrappingParameterSet *params;
    const LweBootstrappingKey *bk;
    const LweBootstrappingKeyFFT *bkFFT;
} TFheGateBootstrappingCloudKeySet;

typedef struct TFheGateBootstrappingSecretKeySet {
    const TFheGateBootstrappingParameterSet *params;
    const LweKey *lwe_key;
    const TGswKey *tgsw_key;
    TFheGateBootstrappingCloudKeySet cloud;
} TFheGateBootstrappingSecretKeySet;

/* ---- allocation (cloud.c:21-23,672-698; freed at cloud.c:48-50,922-933) ---- */
LweSample *new_LweSample_array(int32_t nbelems, const LweParams *params);
void delete_LweSample_array(int32_t nbelems, LweSample *samples);
LweSample *new_gate_bootstrapping_ciphertext_array(int32_t nbelems, const TFheGateBootstrappingParameterSet *params);
void delete_gate_bootstrapping_ciphertext_array(int32_t nbelems, LweSample *samples);
LweSample *new_gate_bootstrapping_ciphertext(const TFheGateBootstrappingParameterSet *params);
void delete_gate_bootstrapping_ciphertext(LweSample *sample);
void delete_gate_bootstrapping_cloud_keyset(TFheGateBootstrappingCloudKeySet *keyset);
void delete_gate_bootstrapping_secret_keyset(TFheGateBootstrappingSecretKeySet *keyset);
void delete_gate_bootstrapping_parameters(TFheGateBootstrappingParameterSet *params);
TFheGateBootstrappingParameterSet *new_default_gate_bootstrapping_parameters(int32_t minimum_lambda);

/* ---- secret-key side (cloud.c:712,745,783,794,823,837) ---- */
void bootsSymEncrypt(LweSample *result, int32_t message, const TFheGateBootstrappingSecretKeySet *key);
int32_t bootsSymDecrypt(const LweSample *sample, const TFheGateBootstrappingSecretKeySet *key);

/* ---- gates; result may alias an input (cloud.c:40,43) ---- */
void bootsCONSTANT(LweSample *result, int32_t value, const TFheGateBootstrappingCloudKeySet *bk);
void bootsNOT(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk);
void bootsCOPY(LweSample *result, const LweSample *ca, const TFheGateBootstrappingCloudKeySet *bk);
void bootsNAND(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsAND(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsXOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsXNOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsNOR(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsANDNY(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsANDYN(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsORNY(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsORYN(LweSample *result, const LweSample *ca, const LweSample *cb, const TFheGateBootstrappingCloudKeySet *bk);
void bootsMUX(LweSample *result, const LweSample *a, const LweSample *b, const LweSample *c, const TFheGateBootstrappingCloudKeySet *bk);

/* ---- extension: `count` independent gates in one launch (arrays of LweSample) ---- */
int ieache_boots_batch(int op, LweSample *result, const LweSample *ca, const LweSample *cb, const LweSample *cc,
                       int32_t count, const TFheGateBootstrappingCloudKeySet *bk);

#ifdef __cplusplus
}
#endif
#endif
