/*
 * tests/compat/add_like_cloud.cpp — a caller written the way Cloud/cloud.c is (its add() at
 * cloud.c:18-51, its file handling at cloud.c:656-733), compiled against include/tfhe/tfhe.h and
 * linked with libieache_b200.so in place of libtfhe.  Run from a directory holding cloud.key,
 * secret.key and cloud.data; prints the decrypted 32-bit sum and a MUX/OR/NOT check.
 */
#include <tfhe/tfhe.h>
#include <tfhe/tfhe_io.h>
#include <stdio.h>

static void add(LweSample *sum, LweSample *carryover, const LweSample *x, const LweSample *y, const LweSample *c,
                const int32_t nb_bits, const TFheGateBootstrappingCloudKeySet *keyset)
{
    const LweParams *in_out_params = keyset->params->in_out_params;
    LweSample *carry = new_LweSample_array(1, in_out_params);
    LweSample *axc = new_LweSample_array(1, in_out_params);
    LweSample *bxc = new_LweSample_array(1, in_out_params);
    bootsCOPY(carry, c, keyset);
    for (int32_t i = 0; i < nb_bits; i++) {
        bootsXOR(axc, x + i, carry, keyset);
        bootsXOR(bxc, y + i, carry, keyset);
        bootsXOR(sum + i, x + i, bxc, keyset);
        bootsAND(axc, axc, bxc, keyset);      /* result aliases an input */
        bootsXOR(carry, carry, axc, keyset);
    }
    bootsCOPY(carryover, carry, keyset);
    delete_LweSample_array(1, carry);
    delete_LweSample_array(1, axc);
    delete_LweSample_array(1, bxc);
}

int main()
{
    FILE *cloud_key = fopen("cloud.key", "rb");
    TFheGateBootstrappingCloudKeySet *bk = new_tfheGateBootstrappingCloudKeySet_fromFile(cloud_key);
    fclose(cloud_key);
    FILE *secret_key = fopen("secret.key", "rb");
    TFheGateBootstrappingSecretKeySet *key = new_tfheGateBootstrappingSecretKeySet_fromFile(secret_key);
    fclose(secret_key);
    const TFheGateBootstrappingParameterSet *params = bk->params;

    LweSample *skip = new_gate_bootstrapping_ciphertext_array(32, params);
    LweSample *ciphertext1 = new_gate_bootstrapping_ciphertext_array(32, params);
    LweSample *ciphertext9 = new_gate_bootstrapping_ciphertext_array(32, params);
    LweSample *ciphertextcarry1 = new_gate_bootstrapping_ciphertext_array(32, params);
    FILE *cloud_data = fopen("cloud.data", "rb");
    for (int blk = 0; blk < 22; blk++) {
        LweSample *dst = blk == 2 ? ciphertext1 : blk == 10 ? ciphertextcarry1 : blk == 13 ? ciphertext9 : skip;
        for (int i = 0; i < 32; i++) import_gate_bootstrapping_ciphertext_fromFile(cloud_data, &dst[i], params);
    }
    fclose(cloud_data);

    LweSample *result = new_gate_bootstrapping_ciphertext_array(32, params);
    LweSample *carry1 = new_gate_bootstrapping_ciphertext_array(32, params);
    add(result, carry1, ciphertext1, ciphertext9, ciphertextcarry1, 32, bk);

    uint32_t sum = 0;
    for (int i = 0; i < 32; i++) sum |= (uint32_t)(bootsSymDecrypt(&result[i], key) > 0) << i;
    printf("sum=%u\n", sum);

    /* the other gates of the API the north star names: OR, MUX, NOT, CONSTANT */
    LweSample *t = new_gate_bootstrapping_ciphertext_array(4, params);
    int ok = 1;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            bootsCONSTANT(&t[0], a, bk);
            bootsSymEncrypt(&t[1], b, key);
            bootsNOT(&t[2], &t[1], bk);
            bootsMUX(&t[3], &t[0], &t[1], &t[2], bk);                    /* a ? b : !b */
            ok &= (bootsSymDecrypt(&t[3], key) > 0) == (a ? b : !b);
            bootsOR(&t[3], &t[0], &t[1], bk);
            ok &= (bootsSymDecrypt(&t[3], key) > 0) == (a | b);
            bootsNAND(&t[3], &t[0], &t[1], bk);
            ok &= (bootsSymDecrypt(&t[3], key) > 0) == !(a & b);
        }
    printf("mux_ok=%d\n", ok);

    delete_gate_bootstrapping_ciphertext_array(4, t);
    delete_gate_bootstrapping_ciphertext_array(32, result);
    delete_gate_bootstrapping_ciphertext_array(32, carry1);
    delete_gate_bootstrapping_ciphertext_array(32, skip);
    delete_gate_bootstrapping_ciphertext_array(32, ciphertext1);
    delete_gate_bootstrapping_ciphertext_array(32, ciphertext9);
    delete_gate_bootstrapping_ciphertext_array(32, ciphertextcarry1);
    delete_gate_bootstrapping_cloud_keyset(bk);
    delete_gate_bootstrapping_secret_keyset(key);
    return 0;
}
