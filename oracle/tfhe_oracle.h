/*
 * oracle/tfhe_oracle.h — CPU restatement of the libtfhe gate-bootstrapping path used by
 * IE-ACHE's Cloud node.  TEST INFRASTRUCTURE ONLY: nothing under ie-ache_b200/ links,
 * imports or executes this code.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * PARITY UNPINNED: the algorithm lives in the third-party library github.com/tfhe/tfhe
 * (branch master, unpinned — /root/reference/README.md:36-48), which is absent from
 * /root/reference and from this image; the reference ships no golden vectors, no
 * known-answer tests and no recorded keys.  This file restates the
 * published TFHE algorithm (Chillotti-Gama-Georgieva-Izabachène, "TFHE: Fast Fully
 * Homomorphic Encryption over the Torus") as recalled in SURVEY.md Appendix A and is
 * anchored on the reference's own call sites: Cloud/cloud.c:18-647 (circuits),
 * Cloud/cloud.c:650-866 (metadata, file layout), Keygen/keygen.c:22-51, Client1/alice.c:58-189,
 * Output/verif.c:46-95.
 */
#ifndef IEACHE_TFHE_ORACLE_H
#define IEACHE_TFHE_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t Torus32;

/* Gate opcodes — shared numbering with include/ieache_b200.h (IEACHE_OP_*). */
enum {
    O_NAND = 0, O_OR = 1, O_AND = 2, O_XOR = 3, O_XNOR = 4, O_NOR = 5,
    O_ANDNY = 6, O_ANDYN = 7, O_ORNY = 8, O_ORYN = 9, O_MUX = 10,
    O_NOT = 11, O_COPY = 12, O_CONST = 13
};

typedef struct OParams {
    int32_t n;          /* LWE dimension (630 for lambda=110 on tfhe master) */
    int32_t N;          /* ring degree (1024) */
    int32_t k;          /* TLWE mask polynomials (1) */
    int32_t bk_l;       /* gadget length (3) */
    int32_t bk_Bgbit;   /* gadget base bits (7) */
    int32_t ks_t;       /* key-switch length (8) */
    int32_t ks_basebit; /* key-switch base bits (2) */
    double ks_stdev;    /* LWE / key-switch noise (2^-15) */
    double bk_stdev;    /* TLWE / bootstrapping-key noise (2^-25) */
    double max_stdev;   /* 0.012467 */
} OParams;

typedef struct OKeySet OKeySet;

/* Keygen/keygen.c:22-23 -> new_default_gate_bootstrapping_parameters(110). */
void o_params_default(OParams *p);
/* Reduced set for fast CPU tests of circuit logic (same N, l, Bgbit; n = n_small). */
void o_params_small(OParams *p, int32_t n_small);

/* Keygen/keygen.c:30-36: one random secret key set (LWE key, TLWE key, BK, KSK). */
OKeySet *o_keygen(const OParams *p, uint64_t seed);
/* wrap externally generated key material (copied); lwe_key / tlwe_key may be NULL (cloud key only) */
OKeySet *o_keyset_from_arrays(const OParams *p, const int32_t *lwe_key, const int32_t *tlwe_key,
                              const Torus32 *bk, const Torus32 *ksk);
void o_keyset_free(OKeySet *ks);
const OParams *o_keyset_params(const OKeySet *ks);
int o_keyset_has_secret(const OKeySet *ks);
/* raw views (host memory, owned by the key set) */
const int32_t *o_lwe_key(const OKeySet *ks);    /* n  */
const int32_t *o_tlwe_key(const OKeySet *ks);   /* k*N */
const Torus32 *o_bk_coef(const OKeySet *ks);    /* [n][(k+1)l][k+1][N] */
const Torus32 *o_ksk(const OKeySet *ks);        /* [kN][t][2^basebit][n+1] */

/* Samples are flat int32 records of (n+1) words: a[0..n) then b. */
void o_sym_encrypt(const OKeySet *ks, const int32_t *bits, size_t count, Torus32 *out, uint64_t seed);
void o_sym_decrypt(const OKeySet *ks, const Torus32 *samples, size_t count, int32_t *bits);
void o_phase(const OKeySet *ks, const Torus32 *samples, size_t count, Torus32 *phases);
/* phase of an extracted (dimension k*N) sample under the TLWE key */
void o_phase_extracted(const OKeySet *ks, const Torus32 *samples, size_t count, Torus32 *phases);

/* One gate (libtfhe boots*). c is only read for O_MUX; b ignored for NOT/COPY;
 * for O_CONST the value is passed in `imm`. out may alias inputs. */
void o_gate(const OKeySet *ks, int op, Torus32 *out, const Torus32 *a, const Torus32 *b,
            const Torus32 *c, int32_t imm);
/* count independent gates, OpenMP over `threads` host threads (0 = all). */
void o_gate_batch(const OKeySet *ks, int op, Torus32 *out, const Torus32 *a, const Torus32 *b,
                  const Torus32 *c, size_t count, int threads);
/* stages, exposed for per-stage parity tests */
void o_bootstrap_woks(const OKeySet *ks, Torus32 *out_extracted /*kN+1*/, Torus32 mu,
                      const Torus32 *x /*n+1*/);
void o_keyswitch(const OKeySet *ks, Torus32 *out /*n+1*/, const Torus32 *u /*kN+1*/);
/* exact_ref.c: o_bootstrap_woks with the external products in exact integer arithmetic (no transform, no rounding) */
void o_bootstrap_woks_exact(const OKeySet *ks, Torus32 *out_extracted /*kN+1*/, Torus32 mu, const Torus32 *x /*n+1*/);
int o_max_threads(void);

/* libtfhe-format files (SURVEY App. A "Serialisation", recalled). */
int o_write_cloud_key(const OKeySet *ks, const char *path);
int o_write_secret_key(const OKeySet *ks, const char *path);
OKeySet *o_read_key(const char *path); /* cloud or secret key set; NULL on error */
int o_write_samples(const OKeySet *ks, const Torus32 *samples, size_t count, const char *path, int append);
long o_read_samples(const OKeySet *ks, Torus32 *samples, size_t max_count, const char *path, size_t skip);

/* ---- circuits of Cloud/cloud.c, one libtfhe-style gate call at a time ---- */
/* arrays are 32 samples, bit i in sample i (LSB first); carry-in = c[0] */
void o_add(const OKeySet *ck, Torus32 *sum, Torus32 *carryover, const Torus32 *x, const Torus32 *y,
           const Torus32 *c, int nb_bits);
void o_mul32(const OKeySet *ck, Torus32 *res_hi, Torus32 *res_lo, const Torus32 *a, const Torus32 *b,
             const Torus32 *carry);
void o_mul64(const OKeySet *ck, Torus32 *r1, Torus32 *r2, Torus32 *r3, const Torus32 *a, const Torus32 *b,
             const Torus32 *c, const Torus32 *carry);
void o_mul128(const OKeySet *ck, Torus32 *r[5], const Torus32 *a, const Torus32 *b, const Torus32 *c,
              const Torus32 *d, const Torus32 *e, const Torus32 *carry);
void o_split(const OKeySet *ck, Torus32 *f1, Torus32 *f2, Torus32 *f3, const Torus32 *a, const Torus32 *b,
             const Torus32 *c, const Torus32 *d, const Torus32 *e, const Torus32 *carry);

/* Client1/alice.c: encrypt one operand into the 11x32-sample client layout
 * (sign code + width under nbit key, 8 value chunks + zero carry block under main key). */
void o_alice(const OKeySet *key, const OKeySet *nbitkey, int32_t sign_code, int32_t width,
             const uint32_t chunks[8], Torus32 *out_352, uint64_t seed);
/* Cloud/cloud.c main(): operands are two 352-sample client blocks; writes the 352-sample
 * answer block (or 64 samples on the abort path). Returns 0, or 126 on the abort path.
 * *out_count receives the number of samples written. */
int o_cloud_main(const OKeySet *cloudkey, const OKeySet *nbitkey, int32_t int_op,
                 const Torus32 *cloud_data_704, Torus32 *answer_352, size_t *out_count, uint64_t seed);
/* Output/verif.c:46-95: decrypt the answer block into sign code, width and 8 chunks. */
void o_verif_decrypt(const OKeySet *key, const OKeySet *nbitkey, const Torus32 *answer_352,
                     int32_t *sign_code, int32_t *width, uint32_t chunks[8]);

/* gate counter for circuit statistics tests (bootstrapped gates since last reset) */
void o_stats_reset(void);
uint64_t o_stats_bootstraps(void);

#ifdef __cplusplus
}
#endif
#endif
