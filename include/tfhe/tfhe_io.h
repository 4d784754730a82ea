/*
 * tfhe/tfhe_io.h — libtfhe-compatible file I/O (Cloud/cloud.c:6), served by the B200 engine.
 * Call sites replaced: cloud.c:657,662 (key sets), cloud.c:705-766 (import), cloud.c:826,839,900 (export).
 */
#ifndef IEACHE_TFHE_IO_COMPAT_H
#define IEACHE_TFHE_IO_COMPAT_H
#include "tfhe.h"

#ifdef __cplusplus
extern "C" {
#endif

/* parses cloud.key, uploads it and builds the transform-domain key on the GPU */
TFheGateBootstrappingCloudKeySet *new_tfheGateBootstrappingCloudKeySet_fromFile(FILE *f);
/* parses a secret key set; only the LWE key and the parameters are kept (that is all cloud.c uses) */
TFheGateBootstrappingSecretKeySet *new_tfheGateBootstrappingSecretKeySet_fromFile(FILE *f);
void import_gate_bootstrapping_ciphertext_fromFile(FILE *f, LweSample *sample, const TFheGateBootstrappingParameterSet *params);
void export_gate_bootstrapping_ciphertext_toFile(FILE *f, const LweSample *sample, const TFheGateBootstrappingParameterSet *params);

#ifdef __cplusplus
}
#endif
#endif
