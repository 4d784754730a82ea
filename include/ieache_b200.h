/*
 * ieache_b200.h — C ABI of the B200-native TFHE gate-bootstrapping engine for IE-ACHE's Cloud node.
 *
 * Plain pointers and sizes only (no C++ or torch types): this is what a ctypes replacement for
 * Cloud/cloud_dynamic.py / Cloud/dragonfly_cipher_cloud.py:compute() binds, and
 * what include/tfhe/tfhe.h (the libtfhe-compatible gate API cloud.c links against) is built on.
 *
 * Every compute entry point runs on the GPU; there is no CPU fallback.  All functions return
 * IEACHE_OK (0) or a negative error code; ieache_last_error() gives the message of the last
 * failure on the calling thread.
 *
 * An LWE sample is a record of (n+1) int32 words: a[0..n) then b  (libtfhe LweSample{a,b},
 * SURVEY.md §8 a12).  Host arrays of samples are contiguous records; device arrays use a stride
 * of IEACHE_DEVICE_STRIDE words.
 */
#ifndef IEACHE_B200_H
#define IEACHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IEACHE_OK 0
#define IEACHE_ERR_CUDA (-1)
#define IEACHE_ERR_ARG (-2)
#define IEACHE_ERR_IO (-3)
#define IEACHE_ERR_FORMAT (-4)
#define IEACHE_ERR_UNSUPPORTED (-5)
#define IEACHE_ERR_NOMEM (-6)

#define IEACHE_DEVICE_STRIDE 632 /* int32 words per LWE sample in device-resident arrays */

/* gate opcodes (libtfhe boots* names; Cloud/cloud.c uses XOR, AND, NOT, COPY, CONSTANT) */
enum ieache_op {
    IEACHE_OP_NAND = 0, IEACHE_OP_OR = 1, IEACHE_OP_AND = 2, IEACHE_OP_XOR = 3, IEACHE_OP_XNOR = 4,
    IEACHE_OP_NOR = 5, IEACHE_OP_ANDNY = 6, IEACHE_OP_ANDYN = 7, IEACHE_OP_ORNY = 8, IEACHE_OP_ORYN = 9,
    IEACHE_OP_MUX = 10, IEACHE_OP_NOT = 11, IEACHE_OP_COPY = 12, IEACHE_OP_CONST = 13
};

/* new_default_gate_bootstrapping_parameters(110) = {630,1024,1,3,7,8,2,2^-15,2^-25,0.012467}
 * (Keygen/keygen.c:22-23).  The engine reads these from the key, it does not assume them;
 * the kernels support N = 1024, k = 1, l in {2,3}, n <= 631. */
typedef struct ieache_params {
    int32_t n, N, k, bk_l, bk_Bgbit, ks_t, ks_basebit, reserved;
    double ks_stdev, bk_stdev, max_stdev;
} ieache_params;

typedef struct ieache_ctx ieache_ctx;           /* one per GPU (several per process are fine: no process-wide state) */
typedef struct ieache_cloudkey ieache_cloudkey; /* device-resident TFheGateBootstrappingCloudKeySet */
typedef struct ieache_circuit ieache_circuit;   /* levelised gate DAG (Cloud/cloud.c circuits) */

const char *ieache_last_error(void);
const char *ieache_version(void);

/* ---- context ---- */
int ieache_ctx_create(int device, ieache_ctx **out);
void ieache_ctx_destroy(ieache_ctx *ctx);
int ieache_ctx_sync(ieache_ctx *ctx);
/* number of engine kernels launched by this context so far (bench.py "gpu_launches") */
uint64_t ieache_ctx_launch_count(const ieache_ctx *ctx);
/* device time of the blind-rotation / key-switch kernels since the last reset, measured with
 * CUDA events on the engine's stream (milliseconds) and number of launches timed */
int ieache_ctx_kernel_times(ieache_ctx *ctx, double *blind_rotate_ms, double *keyswitch_ms,
                            uint64_t *blind_rotate_launches, uint64_t *keyswitch_launches, int reset);
int ieache_ctx_set_timing(ieache_ctx *ctx, int enabled);

/* Kernel selection (launch policy) of this context.  Every blind-rotation and key-switch kernel gives the same
 * results (same parity tests); the policy only chooses the shape that is fastest for the size of a launch:
 *   blind rotation   <= CLUSTER_MAX gates: one gate on a 2-CTA cluster (default 74: one wave of SM pairs);
 *                    <= PAIR_MAX: one gate per CTA on two thread groups (default 296);
 *                    >= W12_MIN: persistent kernel, one warp per gate, 12 gates per SM (default 900);
 *                    between: one gate per 64-thread CTA, or the two-group kernel when that kernel's last wave of
 *                    4 x SMs gates would be less than 90 % full
 *   key switch       <= PAIR_MAX: 8-CTA cluster per gate; >= KS_STAGED_MIN (default 1000): row blocks staged by
 *                    bulk copies and shared by 12 gates; between: per-gate gather
 * THROUGHPUT_KERNEL pins the blind rotation of every launch above PAIR_MAX to IEACHE_KERNEL_GROUP or
 * IEACHE_KERNEL_W12 (0 = by size, the default); any other value is IEACHE_ERR_ARG.  Thresholds must be >= 0.
 * The setting belongs to the context: two contexts (two sessions, two GPUs) never see each other's.  old_value (may be
 * NULL) receives the previous setting.  Tuning / test aid; the defaults are the measured crossovers. */
enum { IEACHE_TUNE_CLUSTER_MAX = 1, IEACHE_TUNE_PAIR_MAX = 2, IEACHE_TUNE_W12_MIN = 3, IEACHE_TUNE_KS_STAGED_MIN = 4,
       IEACHE_TUNE_THROUGHPUT_KERNEL = 5 };
enum { IEACHE_KERNEL_CLUSTER = 1, IEACHE_KERNEL_PAIR = 2, IEACHE_KERNEL_GROUP = 41, IEACHE_KERNEL_W12 = 70 };
enum { IEACHE_KS_CLUSTER = 1, IEACHE_KS_GATHER = 2, IEACHE_KS_STAGED = 3 };
int ieache_ctx_set_tuning(ieache_ctx *ctx, int which, int64_t value, int64_t *old_value);
/* which kernels a launch of `count` gates would use under the context's policy (IEACHE_KERNEL_*, IEACHE_KS_*) */
int ieache_ctx_pick_kernels(const ieache_ctx *ctx, const ieache_cloudkey *key, int64_t count, int *blind_rotate, int *keyswitch);
/* 352-sample operand / answer blocks the session calls have copied host->device and device->host on this context */
int ieache_ctx_copy_counts(const ieache_ctx *ctx, uint64_t *h2d_value_blocks, uint64_t *d2h_value_blocks);
/* step timer: CUDA events on the engine's stream (torch events only see torch's stream) */
int ieache_ctx_timer_start(ieache_ctx *ctx);
int ieache_ctx_timer_stop(ieache_ctx *ctx, double *elapsed_ms); /* records, synchronises, returns the elapsed device time */
/* dense FP64 FMA throughput of this GPU (TFLOP/s, 2 flops per FMA), measured live: the roofline
 * denominator of the transform + multiply-accumulate work (MEASURED_PEAKS.json has no FP64 figure) */
int ieache_measure_fp64_peak(ieache_ctx *ctx, double *tflops);
/* pinned host memory for the end-to-end path */
int ieache_host_alloc(size_t bytes, void **out);
int ieache_host_free(void *ptr);

/* ---- cloud key: replaces new_tfheGateBootstrappingCloudKeySet_fromFile (Cloud/cloud.c:656-658) ---- */
/* bk:  int32 [n][(k+1)l][k+1][N]  coefficient-domain TGSW samples (libtfhe bk->bk[i].all_sample[r].a[j])
 * ksk: int32 [kN][t][2^basebit][n+1]  key-switch samples (libtfhe bk->ks->ks[i][j][d])
 * The transform-domain copy (libtfhe bkFFT) is computed on the GPU. */
int ieache_cloudkey_create(ieache_ctx *ctx, const ieache_params *p, const int32_t *bk, const int32_t *ksk,
                           ieache_cloudkey **out);
int ieache_cloudkey_load_file(ieache_ctx *ctx, const char *path, ieache_cloudkey **out);
void ieache_cloudkey_destroy(ieache_cloudkey *key);
int ieache_cloudkey_params(const ieache_cloudkey *key, ieache_params *out);
/* device arrays of the key, for the one-time NCCL broadcast to the other GPUs */
int ieache_cloudkey_device_arrays(const ieache_cloudkey *key, void **bkfft, size_t *bkfft_bytes, void **ksk,
                                  size_t *ksk_bytes);
/* bytes the arrays of a key with parameters p occupy on the device */
int ieache_cloudkey_device_sizes(const ieache_params *p, size_t *bkfft_bytes, size_t *ksk_bytes);
/* wrap arrays that already hold a key (filled by a broadcast); the caller keeps ownership */
int ieache_cloudkey_adopt_device(ieache_ctx *ctx, const ieache_params *p, void *bkfft, void *ksk,
                                 ieache_cloudkey **out);

/* ---- secret-key side on the GPU: Keygen/keygen.c, Client1/alice.c:117, Output/verif.c:93 ---- */
typedef struct ieache_secretkey ieache_secretkey; /* LWE key (n bits) + TLWE key (N bits), host and device copies */
/* new_random_gate_bootstrapping_secret_keyset (Keygen/keygen.c:30-36): draws both secret keys and builds the
 * bootstrapping and key-switch keys directly in device layout.  If bk_export / ksk_export are non-NULL they receive
 * the libtfhe-order coefficient arrays ([n][(k+1)l][k+1][N] and [kN][t][2^basebit][n+1]) so the key can be written to
 * cloud.key.
 * Randomness: ChaCha20 key streams under two independent 256-bit keys, one for the secret bits and the noise, one
 * for the published masks.  seed == 0 (use this): both keys come from the operating system (getrandom).
 * seed != 0: both keys are derived from the seed — reproducible key sets FOR TESTS ONLY; such a key set is exactly as
 * secret as its 64-bit seed, and two key sets made from one seed are identical. */
int ieache_keygen(ieache_ctx *ctx, const ieache_params *p, uint64_t seed, ieache_secretkey **sk, ieache_cloudkey **ck,
                  int32_t *bk_export, int32_t *ksk_export);
int ieache_secretkey_import(ieache_ctx *ctx, const ieache_params *p, const int32_t *lwe_key, const int32_t *tlwe_key,
                            ieache_secretkey **sk);
int ieache_secretkey_export(const ieache_secretkey *sk, int32_t *lwe_key /*n*/, int32_t *tlwe_key /*N, may be NULL*/);
void ieache_secretkey_destroy(ieache_secretkey *sk);
/* bootsSymEncrypt of `count` bits (host array) into device samples (stride IEACHE_DEVICE_STRIDE).
 * seed == 0 (use this): fresh masks and noise from the operating system's entropy on every call.
 * seed != 0: reproducible, FOR TESTS ONLY — two calls with the same seed use the same masks, which reveals the
 * difference of their plaintexts. */
int ieache_sym_encrypt_device(ieache_ctx *ctx, const ieache_secretkey *sk, const int32_t *bits, size_t count, int32_t *out_dev,
                              uint64_t seed);
/* bootsSymDecrypt of device samples: bits (phase > 0) and, optionally, the raw phases, to host arrays */
int ieache_sym_decrypt_device(ieache_ctx *ctx, const ieache_secretkey *sk, const int32_t *samples_dev, size_t count,
                              int32_t *bits, int32_t *phases);

/* ---- batched gates: `count` independent boots<OP>(out[g], a[g], b[g] (, c[g])) ---- */
/* host buffers, count x (n+1) words each; out may alias an input; b/c NULL where unused.
 * For IEACHE_OP_CONST, a is NULL and imm is the plaintext bit. */
int ieache_gate_batch(ieache_ctx *ctx, const ieache_cloudkey *key, int op, int32_t *out, const int32_t *a,
                      const int32_t *b, const int32_t *c, int32_t imm, size_t count);
/* device-resident buffers with stride IEACHE_DEVICE_STRIDE; asynchronous on the context stream */
int ieache_gate_batch_device(ieache_ctx *ctx, const ieache_cloudkey *key, int op, int32_t *out, const int32_t *a,
                             const int32_t *b, const int32_t *c, int32_t imm, size_t count);
/* stages, for per-stage parity tests: pre-combined x -> extracted (kN+1 words, stride 1028 on device) */
int ieache_bootstrap_woks(ieache_ctx *ctx, const ieache_cloudkey *key, int32_t *ext_out /*count x 1025*/,
                          const int32_t *x /*count x (n+1)*/, size_t count);
int ieache_keyswitch(ieache_ctx *ctx, const ieache_cloudkey *key, int32_t *out /*count x (n+1)*/,
                     const int32_t *ext /*count x 1025*/, size_t count);

/* device memory helpers for callers without a CUDA binding (ctypes) */
int ieache_device_alloc(ieache_ctx *ctx, size_t bytes, void **out);
int ieache_device_free(ieache_ctx *ctx, void *ptr);
/* device-to-device copy on the context stream, synchronous on return (key replication plumbing) */
int ieache_device_copy(ieache_ctx *ctx, void *dst, const void *src, size_t bytes);
/* copy count samples between packed host records (n+1 words) and strided device records */
int ieache_samples_to_device(ieache_ctx *ctx, int32_t *dev, const int32_t *host, size_t count, int32_t n);
int ieache_samples_to_host(ieache_ctx *ctx, int32_t *host, const int32_t *dev, size_t count, int32_t n);

/* ---- circuits of Cloud/cloud.c as levelised DAGs ---- */
/* kinds: the dispatch of Cloud/cloud.c main() */
enum ieache_circuit_kind {
    IEACHE_CIRC_ADD = 1,      /* chained add()            cloud.c:870-1190 ; inputs A chunks, B chunks, carry */
    IEACHE_CIRC_SUB = 2,      /* A - B via two's complement cloud.c:1194-1800 */
    IEACHE_CIRC_MUL = 4,      /* mul32 / 2xmul64+split / 4xmul128+15 adds  cloud.c:2368-2719 */
    IEACHE_CIRC_MULADD = 5    /* a*b+c: mul32 then 64-bit add (BASELINE.json config 3) */
};
/* width in bits of each operand: 32, 64, 128 or 256 (MUL: 32, 64, 128; MULADD: 32) */
int ieache_circuit_build(int kind, int width, ieache_circuit **out);
void ieache_circuit_destroy(ieache_circuit *c);
/* statistics checked against SURVEY.md App. B */
int ieache_circuit_stats(const ieache_circuit *c, uint64_t *bootstraps, uint64_t *and_gates, uint64_t *xor_gates,
                         uint32_t *levels, uint32_t *max_width, uint32_t *n_inputs, uint32_t *n_outputs);
/* evaluate n_expr independent instances, level by level, one blind-rotation launch and one
 * key-switch launch per level over all instances.
 * inputs : host, [n_expr][n_inputs][n+1]; outputs: host, [n_expr][n_outputs][n+1]
 * Input order: operand-1 chunks (32 samples each, LSB first), operand-2 chunks, (MULADD: operand-3
 * chunk,) then the 32-sample carry block of operand 1 (only sample 0 is read, cloud.c:24). */
int ieache_circuit_eval(ieache_ctx *ctx, const ieache_cloudkey *key, const ieache_circuit *c, const int32_t *inputs,
                        int32_t *outputs, size_t n_expr);
/* same with device-resident inputs/outputs (stride IEACHE_DEVICE_STRIDE), asynchronous */
int ieache_circuit_eval_device(ieache_ctx *ctx, const ieache_cloudkey *key, const ieache_circuit *c,
                               const int32_t *inputs, int32_t *outputs, size_t n_expr);

/* ---- the Cloud node's process contract (Cloud/cloud.c main()) ---- */
/* Reads cloud.key, nbit.key, cloud.data, operator.txt in `dir`, writes answer.data (and appends
 * averagestandard.txt on multiply).  Returns the exit code of ./cloud: 0, or 126 on the
 * 256-bit-multiply abort path; negative on engine errors.  seconds (optional) receives the
 * circuit time the reference prints as "Computation Time". */
int ieache_cloud_run(ieache_ctx *ctx, const char *dir, double *seconds);

/* ---- the callers on either side of the path, on the reference's files ---- */
/* Keygen/keygen.c: writes secret.key, cloud.key, nbit.key into dir (p == NULL: lambda = 110 defaults).
 * seed_key / seed_nbit: 0 = operating-system entropy (use this); non-zero = reproducible, tests only (see ieache_keygen).
 * Returns IEACHE_ERR_IO when a file cannot be written completely. */
int ieache_keygen_files(ieache_ctx *ctx, const char *dir, const ieache_params *p, uint64_t seed_key, uint64_t seed_nbit);
/* Client1/alice.c: one operand (sign code 0/2, width, 8 chunks least significant first) -> 352 records */
int ieache_alice_encrypt(const char *dir, int32_t sign_code, int32_t width, const uint32_t *chunks, const char *out_path,
                         int append);
int ieache_alice_run(const char *dir); /* values.txt -> cloud.data, like ./alice */
/* Output/verif.c: decrypts answer.data with secret.key / nbit.key, applies the sign rules of operator.txt and
 * returns the decimal string verif prints */
int ieache_verif_run(const char *dir, char *result, size_t result_cap, int32_t *sign_code, int32_t *width);

/* ---- sessions: keys loaded once, operators chained in memory, requests batched ---- */
typedef struct ieache_session ieache_session;
/* loads cloud.key onto the GPU and the LWE key of nbit.key (Cloud/cloud.c:656-663) once */
int ieache_session_open(ieache_ctx *ctx, const char *cloud_key_path, const char *nbit_key_path, ieache_session **out);
/* same from objects already in memory; the cloud key is borrowed */
int ieache_session_open_keys(ieache_ctx *ctx, ieache_cloudkey *key, const int32_t *nbit_lwe_key, ieache_session **out);
void ieache_session_close(ieache_session *s);
int ieache_session_params(const ieache_session *s, ieache_params *out);
/* Cloud/cloud.c main() on memory blocks: operand blocks are the 352-sample client layout (11 x 32 samples,
 * packed n+1 words); answer is 352 samples (64 on the abort path).  Returns 0 / 126 / negative error. */
int ieache_session_compute(ieache_session *s, int op, const int32_t *operand1, const int32_t *operand2, int32_t *answer,
                           size_t *answer_count, double *seconds);
/* `count` independent requests; requests needing the same circuit share every level launch */
int ieache_session_compute_batch(ieache_session *s, size_t count, const int32_t *ops, const int32_t *operands1,
                                 const int32_t *operands2, int32_t *answers, int32_t *exit_codes, size_t *answer_counts,
                                 double *seconds);
/* whole postfix expressions ("AB*C+"), the walk of Cloud/dragonfly_cipher_cloud.py:685-729 without the
 * answer.data -> cloud.data round trips; operands: [n_expr][n_operands][352][n+1] */
int ieache_session_eval_postfix(ieache_session *s, const char *postfix, size_t n_expr, const int32_t *operands,
                                int n_operands, int32_t *answers, size_t *answer_counts, double *seconds);
/* Batched ingest (replaces `count` runs of ./cloud, Cloud/dragonfly_cipher_cloud.py:1233): every dirs[i] holds
 * cloud.data (2 x 352 records) and operator.txt; all requests are evaluated as one levelised batch and every
 * directory receives its answer.data (plus the averagestandard.txt line on multiply).  exit_codes[i] = 0 / 126 as
 * ./cloud would exit, or a negative IEACHE_ERR_* if that directory could not be read or written. */
int ieache_session_compute_dirs(ieache_session *s, size_t count, const char *const *dirs, int32_t *exit_codes, double *seconds);
/* Requests evaluated together by compute_batch / eval_postfix / compute_dirs (default 256, 1..65536).  Larger passes fill
 * the GPU better for shallow circuits (a 32-bit add has 1-2 gates per level and request); memory per request of a pass:
 * 2.2 MB on the device, and in compute_dirs 5.3 MB on the host (the files of the next pass are read and the answers of
 * the previous one written while a pass computes).  Pinned staging stays at 2 x 47 MB whatever the pass size. */
int ieache_session_set_pass(ieache_session *s, size_t requests, size_t *old_value);

#ifdef __cplusplus
}
#endif
#endif
