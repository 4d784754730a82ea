/*
 * tfhe_io.h — readers/writers for libtfhe's key and ciphertext files (host side).
 * Replaces new_tfheGateBootstrapping{Cloud,Secret}KeySet_fromFile and
 * import_/export_gate_bootstrapping_ciphertext_{from,to}File as used at
 * Cloud/cloud.c:656-663, 703-766, 826, 839, 900.  Format as recalled in SURVEY.md App. A
 * ("Serialisation"); the record size 4 + 4(n+1) + 8 = 2536 B for n = 630 is corroborated by the
 * reference's 162304-byte abort file (Cloud/dragonfly_cipher_cloud.py:1295).
 */
#ifndef IEACHE_TFHE_IO_H
#define IEACHE_TFHE_IO_H
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/ieache_b200.h"

namespace ieache {

struct HostKeySet {
    ieache_params p{};
    std::vector<int32_t> bk;   /* [n][kpl][k+1][N] */
    std::vector<int32_t> ksk;  /* [kN][t][base][n+1] */
    bool has_secret = false;
    std::vector<int32_t> lwe_key;  /* n */
    std::vector<int32_t> tlwe_key; /* k*N */
};

/* returns IEACHE_OK or an error code; msg receives a description on failure */
int read_keyset(const char *path, HostKeySet &ks, bool want_bk, std::string &msg);
int read_keyset_stream(FILE *f, HostKeySet &ks, bool want_bk, std::string &msg);
int write_keyset(const char *path, const HostKeySet &ks, bool with_secret, std::string &msg);

size_t sample_record_bytes(int n);
/* read `count` LWE samples (packed n+1 words each) from an open file */
int read_samples(FILE *f, int n, int32_t *dst, size_t count);
int write_samples(FILE *f, int n, const int32_t *src, size_t count, double variance);

/* bootsSymEncrypt / bootsSymDecrypt with a secret key (used by cloud.c for the nbit metadata,
 * Cloud/cloud.c:712,745,783,794,823,837) */
void sym_encrypt_bits(const HostKeySet &ks, const int32_t *bits, size_t count, int32_t *out);
void sym_decrypt_bits(const HostKeySet &ks, const int32_t *samples, size_t count, int32_t *bits);

} // namespace ieache
#endif
