/*
 * oracle/cloud_oracle.c — restatement of IE-ACHE's encrypted arithmetic circuits and of the
 * Cloud node's main(), one libtfhe-style gate call at a time, in the reference's own order.
 * TEST INFRASTRUCTURE ONLY (see tfhe_oracle.h).  PARITY UNPINNED at the ciphertext level;
 * circuit structure follows the reference line by line:
 *   add     Cloud/cloud.c:18-51       zero  :53-57      NOT   :59-63     split :65-113
 *   mul32   Cloud/cloud.c:115-218     mul64 :220-385    mul128 :387-647
 *   main    Cloud/cloud.c:650-2720    (metadata :775-855, abort :860-864, dispatch :870/1194/2368)
 *   alice   Client1/alice.c:58-189    verif decrypt Output/verif.c:46-95
 */
#include "tfhe_oracle.h"

#include <stdlib.h>
#include <string.h>

#define W(ks) ((size_t)o_keyset_params(ks)->n + 1)

static Torus32 *arr32(const OKeySet *ck) { return (Torus32 *)calloc(32 * W(ck), sizeof(Torus32)); }
static void g2(const OKeySet *ck, int op, Torus32 *out, const Torus32 *a, const Torus32 *b) { o_gate(ck, op, out, a, b, NULL, 0); }
static void g_const(const OKeySet *ck, Torus32 *out, int v) { o_gate(ck, O_CONST, out, NULL, NULL, NULL, v); }
static void g_copy(const OKeySet *ck, Torus32 *out, const Torus32 *a) { o_gate(ck, O_COPY, out, a, NULL, NULL, 0); }

/* Cloud/cloud.c:18-51 */
void o_add(const OKeySet *ck, Torus32 *sum, Torus32 *carryover, const Torus32 *x, const Torus32 *y,
           const Torus32 *c, int nb_bits)
{
    const size_t w = W(ck);
    Torus32 *carry = (Torus32 *)calloc(w, sizeof(Torus32)), *axc = (Torus32 *)calloc(w, sizeof(Torus32)),
            *bxc = (Torus32 *)calloc(w, sizeof(Torus32));
    g_copy(ck, carry, c);
    for (int i = 0; i < nb_bits; i++) {
        g2(ck, O_XOR, axc, x + i * w, carry);
        g2(ck, O_XOR, bxc, y + i * w, carry);
        g2(ck, O_XOR, sum + i * w, x + i * w, bxc);
        g2(ck, O_AND, axc, axc, bxc);
        g2(ck, O_XOR, carry, carry, axc);
    }
    g_copy(ck, carryover, carry);
    free(carry); free(axc); free(bxc);
}

static void zero32(const OKeySet *ck, Torus32 *r) { for (int i = 0; i < 32; i++) g_const(ck, r + i * W(ck), 0); }
static void not32(const OKeySet *ck, Torus32 *r, const Torus32 *x)
{
    for (int i = 0; i < 32; i++) o_gate(ck, O_NOT, r + i * W(ck), x + i * W(ck), NULL, NULL, 0);
}
static void copy32(const OKeySet *ck, Torus32 *r, const Torus32 *x) { memcpy(r, x, 32 * W(ck) * sizeof(Torus32)); }

/* Cloud/cloud.c:65-113 */
void o_split(const OKeySet *ck, Torus32 *f1, Torus32 *f2, Torus32 *f3, const Torus32 *a, const Torus32 *b,
             const Torus32 *c, const Torus32 *d, const Torus32 *e, const Torus32 *carry)
{
    Torus32 *sum = arr32(ck), *sum2 = arr32(ck), *sum3 = arr32(ck);
    Torus32 *co = arr32(ck), *co2 = arr32(ck), *co3 = arr32(ck);
    zero32(ck, sum); zero32(ck, sum2); zero32(ck, sum3); zero32(ck, co); zero32(ck, co2); zero32(ck, co3);
    o_add(ck, sum, co, e, b, carry, 32);
    o_add(ck, sum2, co2, d, a, co, 32);
    o_add(ck, sum3, co3, c, co2, carry, 32);
    copy32(ck, f1, sum3); copy32(ck, f2, sum2); copy32(ck, f3, sum);
    free(sum); free(sum2); free(sum3); free(co); free(co2); free(co3);
}

/* shared shape of mul32/mul64/mul128 (Cloud/cloud.c:148-198, :265-359, :443-612): K input chunks
 * x[0..K) (least significant first) times the 32-bit multiplier m; sums[0..K] least significant first. */
static void mulK(const OKeySet *ck, int K, Torus32 **sums /*K+1*/, const Torus32 *const *x, const Torus32 *m,
                 const Torus32 *carry)
{
    const size_t w = W(ck);
    Torus32 *tmp[4], *word[5], *cy[5];
    for (int q = 0; q < K; q++) { tmp[q] = arr32(ck); zero32(ck, tmp[q]); }
    for (int q = 0; q <= K; q++) { word[q] = arr32(ck); zero32(ck, word[q]); cy[q] = arr32(ck); zero32(ck, cy[q]); zero32(ck, sums[q]); }
    for (int round = 0; round < 32; round++) {
        for (int kk = 0; kk < 32; kk++)
            for (int q = 0; q < K; q++) g2(ck, O_AND, tmp[q] + kk * w, x[q] + kk * w, m + round * w);
        const int c1 = 32 - round, c2 = round;
        for (int i = 0; i < round; i++) g_const(ck, word[0] + i * w, 0);
        for (int i = 0; i < c1; i++) g_copy(ck, word[0] + (i + round) * w, tmp[0] + i * w);
        for (int q = 1; q <= K; q++) {
            for (int i = 0; i < c2; i++) g_copy(ck, word[q] + i * w, tmp[q - 1] + (i + c1) * w);
            if (q < K) for (int i = 0; i < c1; i++) g_copy(ck, word[q] + (i + c2) * w, tmp[q] + i * w);
        }
        for (int q = 0; q <= K; q++) o_add(ck, sums[q], cy[q], sums[q], word[q], q ? cy[q - 1] : carry, 32);
    }
    for (int q = 0; q < K; q++) free(tmp[q]);
    for (int q = 0; q <= K; q++) { free(word[q]); free(cy[q]); }
}

/* Cloud/cloud.c:115-218: result = high word, result2 = low word */
void o_mul32(const OKeySet *ck, Torus32 *res_hi, Torus32 *res_lo, const Torus32 *a, const Torus32 *b, const Torus32 *carry)
{
    Torus32 *sums[2] = {arr32(ck), arr32(ck)};
    const Torus32 *x[1] = {a};
    mulK(ck, 1, sums, x, b, carry);
    copy32(ck, res_hi, sums[1]); copy32(ck, res_lo, sums[0]);
    free(sums[0]); free(sums[1]);
}
/* Cloud/cloud.c:220-385: result = top, result3 = lowest */
void o_mul64(const OKeySet *ck, Torus32 *r1, Torus32 *r2, Torus32 *r3, const Torus32 *a, const Torus32 *b,
             const Torus32 *c, const Torus32 *carry)
{
    Torus32 *sums[3] = {arr32(ck), arr32(ck), arr32(ck)};
    const Torus32 *x[2] = {a, b};
    mulK(ck, 2, sums, x, c, carry);
    copy32(ck, r1, sums[2]); copy32(ck, r2, sums[1]); copy32(ck, r3, sums[0]);
    for (int q = 0; q < 3; q++) free(sums[q]);
}
/* Cloud/cloud.c:387-647: r[0] = top ... r[4] = lowest */
void o_mul128(const OKeySet *ck, Torus32 *r[5], const Torus32 *a, const Torus32 *b, const Torus32 *c,
              const Torus32 *d, const Torus32 *e, const Torus32 *carry)
{
    Torus32 *sums[5];
    for (int q = 0; q < 5; q++) sums[q] = arr32(ck);
    const Torus32 *x[4] = {a, b, c, d};
    mulK(ck, 4, sums, x, e, carry);
    for (int q = 0; q < 5; q++) copy32(ck, r[q], sums[4 - q]);
    for (int q = 0; q < 5; q++) free(sums[q]);
}

/* Client1/alice.c:58-66,116-149,167-189 */
void o_alice(const OKeySet *key, const OKeySet *nbitkey, int32_t sign_code, int32_t width,
             const uint32_t chunks[8], Torus32 *out, uint64_t seed)
{
    const size_t w = W(key);
    int32_t bits[32];
    for (int i = 0; i < 32; i++) bits[i] = (sign_code >> i) & 1;
    o_sym_encrypt(nbitkey, bits, 32, out, seed * 16 + 1);
    for (int i = 0; i < 32; i++) bits[i] = (width >> i) & 1;
    o_sym_encrypt(nbitkey, bits, 32, out + 32 * w, seed * 16 + 2);
    for (int cidx = 0; cidx < 8; cidx++) {
        /* alice.c only encrypts width/32 chunks; the remaining blocks are never read by cloud.c
         * for that width, so encrypting their zero value is equivalent. */
        uint32_t v = (cidx < width / 32) ? chunks[cidx] : 0;
        for (int i = 0; i < 32; i++) bits[i] = (v >> i) & 1;
        o_sym_encrypt(key, bits, 32, out + (size_t)(2 + cidx) * 32 * w, seed * 16 + 3 + cidx);
    }
    for (int i = 0; i < 32; i++) bits[i] = 0;
    o_sym_encrypt(key, bits, 32, out + (size_t)10 * 32 * w, seed * 16 + 11);
}

static int32_t dec32(const OKeySet *k, const Torus32 *blk)
{
    int32_t bits[32], v = 0;
    o_sym_decrypt(k, blk, 32, bits);
    for (int i = 0; i < 32; i++) v |= bits[i] << i;
    return v;
}
static void enc32(const OKeySet *k, int32_t v, Torus32 *blk, uint64_t seed)
{
    int32_t bits[32];
    for (int i = 0; i < 32; i++) bits[i] = (v >> i) & 1;
    o_sym_encrypt(k, bits, 32, blk, seed);
}

/* Output/verif.c:46-95 */
void o_verif_decrypt(const OKeySet *key, const OKeySet *nbitkey, const Torus32 *ans, int32_t *sign_code,
                     int32_t *width, uint32_t chunks[8])
{
    const size_t w = W(key);
    *sign_code = dec32(nbitkey, ans);
    *width = dec32(nbitkey, ans + 32 * w);
    for (int c = 0; c < 8; c++) chunks[c] = (uint32_t)dec32(key, ans + (size_t)(2 + c) * 32 * w);
}

/* Cloud/cloud.c:650-2720 */
int o_cloud_main(const OKeySet *ck, const OKeySet *nbitkey, int32_t int_op, const Torus32 *data,
                 Torus32 *answer, size_t *out_count, uint64_t seed)
{
    const size_t w = W(ck), B = 32 * w;
    const Torus32 *neg1 = data, *bit1 = data + B, *carry1 = data + 10 * B;
    const Torus32 *neg2 = data + 11 * B, *bit2 = data + 12 * B;
    const Torus32 *ct[16];
    for (int j = 0; j < 8; j++) { ct[j] = data + (size_t)(2 + j) * B; ct[8 + j] = data + (size_t)(13 + j) * B; }

    int32_t int_bit1 = dec32(nbitkey, bit1), int_bit2 = dec32(nbitkey, bit2);          /* :709-746 */
    int32_t n1 = dec32(nbitkey, neg1), n2 = dec32(nbitkey, neg2);                      /* :780-796 */
    if (n1 == 2) n1 = 1;                                                               /* :788-789 */
    const int32_t int_negative = n1 + n2;                                              /* :804 */
    /* :812-821: 1 -> 1, 2 -> 2, 3 -> 4 and 0 for every other sum (e.g. a chained operand that already carries code 4) */
    const int32_t code = int_negative == 1 ? 1 : int_negative == 2 ? 2 : int_negative == 3 ? 4 : 0;
    enc32(nbitkey, code, answer, seed * 4 + 1);                                        /* :822-826 */
    int32_t int_bit;
    if (int_op == 4) {                                                                 /* :833-843 */
        int_bit = (int_bit1 >= int_bit2 ? int_bit1 : int_bit2);
        enc32(nbitkey, int_bit * 2, answer + B, seed * 4 + 2);
    } else if (int_bit1 >= int_bit2) { int_bit = int_bit1; memcpy(answer + B, bit1, B * sizeof(Torus32)); }
    else { int_bit = int_bit2; memcpy(answer + B, bit2, B * sizeof(Torus32)); }
    *out_count = 64;
    if (int_op == 4 && int_bit >= 256) return 126;                                     /* :860-864 */

    Torus32 *res[8] = {0};
    int nres = 0;
    const int nc = int_bit / 32;
    const int width_ok = (int_bit == 32 || int_bit == 64 || int_bit == 128 || int_bit == 256);

    if ((int_op == 1 && int_negative != 1 && int_negative != 2) ||
        (int_op == 2 && (int_negative == 1 || int_negative == 2))) {                   /* :870 */
        if (width_ok) {
            Torus32 *cy_prev = NULL;
            for (int j = 0; j < nc; j++) {
                res[j] = arr32(ck);
                Torus32 *cy = arr32(ck);
                o_add(ck, res[j], cy, ct[j], ct[8 + j], j ? cy_prev : carry1, 32);     /* :891,951-952,1020-1023,1109-1116 */
                free(cy_prev); cy_prev = cy;
            }
            free(cy_prev);
            nres = nc;
        }
    } else if (int_op == 2 || (int_op == 1 && (int_negative == 1 || int_negative == 2))) {   /* :1194 */
        /* b_from_a: A - B (:1196); otherwise B - A (:1809) */
        const int b_from_a = (int_op == 2 && int_negative == 0) || (int_op == 1 && int_negative == 2);
        if (width_ok) {
            const Torus32 *const *minuend = b_from_a ? ct : ct + 8;
            const Torus32 *const *subtrahend = b_from_a ? ct + 8 : ct;
            Torus32 *temp = arr32(ck), *zeros = arr32(ck);
            zero32(ck, temp); zero32(ck, zeros);
            g_const(ck, temp, 1);                                                      /* :1233 */
            Torus32 *tw[8], *twc_prev = NULL;
            for (int j = 0; j < nc; j++) {                                             /* :1225-1236,1325-1341 */
                Torus32 *inv = arr32(ck), *twc = arr32(ck);
                tw[j] = arr32(ck);
                not32(ck, inv, subtrahend[j]);
                o_add(ck, tw[j], twc, inv, j ? zeros : temp, j ? twc_prev : zeros, 32);
                free(inv); free(twc_prev); twc_prev = twc;
            }
            free(twc_prev);
            Torus32 *cy_prev = NULL;
            for (int j = 0; j < nc; j++) {                                             /* :1245,1352-1353 */
                res[j] = arr32(ck);
                Torus32 *cy = arr32(ck);
                o_add(ck, res[j], cy, minuend[j], tw[j], j ? cy_prev : carry1, 32);
                free(cy_prev); cy_prev = cy; free(tw[j]);
            }
            free(cy_prev); free(temp); free(zeros);
            nres = nc;
        }
    } else if (int_op == 4) {                                                          /* :2368 */
        if (int_bit == 32) {                                                           /* :2655-2686 */
            res[0] = arr32(ck); res[1] = arr32(ck);
            o_mul32(ck, res[1], res[0], ct[0], ct[8], carry1);
            nres = 2;
        } else if (int_bit == 64) {                                                    /* :2568-2616 */
            Torus32 *r[6];
            for (int q = 0; q < 6; q++) r[q] = arr32(ck);
            o_mul64(ck, r[0], r[1], r[2], ct[0], ct[1], ct[8], carry1);
            o_mul64(ck, r[3], r[4], r[5], ct[0], ct[1], ct[9], carry1);
            res[0] = r[2];
            res[1] = arr32(ck); res[2] = arr32(ck); res[3] = arr32(ck);
            o_split(ck, res[3], res[2], res[1], r[0], r[1], r[3], r[4], r[5], carry1);
            free(r[0]); free(r[1]); free(r[3]); free(r[4]); free(r[5]);
            nres = 4;
        } else if (int_bit == 128) {                                                   /* :2371-2491 */
            Torus32 *r[21], *s[16], *co[16];
            for (int q = 1; q <= 20; q++) r[q] = arr32(ck);
            for (int q = 1; q <= 15; q++) { s[q] = arr32(ck); co[q] = arr32(ck); }
            for (int y = 0; y < 4; y++) {
                Torus32 *rr[5] = {r[5 * y + 1], r[5 * y + 2], r[5 * y + 3], r[5 * y + 4], r[5 * y + 5]};
                o_mul128(ck, rr, ct[0], ct[1], ct[2], ct[3], ct[8 + y], carry1);       /* :2434-2443 */
            }
            o_add(ck, s[1], co[1], r[10], r[4], carry1, 32);                           /* :2445-2461 */
            o_add(ck, s[2], co[2], r[9], r[3], co[1], 32);
            o_add(ck, s[3], co[3], r[8], r[2], co[2], 32);
            o_add(ck, s[4], co[4], r[7], r[1], co[3], 32);
            o_add(ck, s[5], co[5], r[6], carry1, co[4], 32);
            o_add(ck, s[6], co[6], s[2], r[15], co[5], 32);
            o_add(ck, s[7], co[7], s[3], r[14], co[6], 32);
            o_add(ck, s[8], co[8], s[4], r[13], co[7], 32);
            o_add(ck, s[9], co[9], s[5], r[12], co[8], 32);
            o_add(ck, s[10], co[10], r[11], carry1, co[9], 32);
            o_add(ck, s[11], co[11], s[7], r[20], co[10], 32);
            o_add(ck, s[12], co[12], s[8], r[19], co[11], 32);
            o_add(ck, s[13], co[13], s[9], r[18], co[12], 32);
            o_add(ck, s[14], co[14], s[10], r[17], co[13], 32);
            o_add(ck, s[15], co[15], r[16], carry1, co[14], 32);
            Torus32 *order[8] = {r[5], s[1], s[6], s[11], s[12], s[13], s[14], s[15]};  /* :2476-2491 */
            for (int q = 0; q < 8; q++) { res[q] = arr32(ck); copy32(ck, res[q], order[q]); }
            for (int q = 1; q <= 20; q++) free(r[q]);
            for (int q = 1; q <= 15; q++) { free(s[q]); free(co[q]); }
            nres = 8;
        }
    }
    if (nres == 0) return 0; /* unsupported width: the reference writes nothing more */
    /* export: results, padding copies of operand 1's carry block, then the carry block (:899-916) */
    for (int q = 0; q < 8; q++) memcpy(answer + (size_t)(2 + q) * B, q < nres ? res[q] : carry1, B * sizeof(Torus32));
    memcpy(answer + (size_t)10 * B, carry1, B * sizeof(Torus32));
    for (int q = 0; q < nres; q++) free(res[q]);
    *out_count = 352;
    return 0;
}
