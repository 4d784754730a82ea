/* tfhe_io.cpp — see tfhe_io.h */
#include "tfhe_io.h"
#include "csprng.h"

#include <cmath>
#include <cstring>
#include <map>
#include <new>

namespace ieache {

/* type ids as recalled from libtfhe's tfhe_io.cpp; unverified, so the reader
 * only insists on the framing being consistent with the parameter block, not on the id values */
enum : int32_t { UID_LWE_KEY = 43, UID_LWE_SAMPLE = 42, UID_KS_KEY = 200, UID_TLWE_SAMPLE = 45,
                 UID_TGSW_KEY = 47, UID_TGSW_SAMPLE = 48, UID_BK = 201 };

size_t sample_record_bytes(int n) { return 4 + 4 * (size_t)(n + 1) + 8; }

static bool read_param_blocks(FILE *f, ieache_params &p, std::string &msg)
{
    std::map<std::string, std::map<std::string, double>> blocks;
    std::string title;
    char line[512];
    for (;;) {
        long pos = ftell(f);
        int ch = fgetc(f);
        if (ch == EOF) break;
        ungetc(ch, f);
        if (title.empty() && ch != '-') { fseek(f, pos, SEEK_SET); break; } /* binary section */
        if (!fgets(line, sizeof line, f)) break;
        if (!strncmp(line, "-----BEGIN ", 11)) {
            char t[64] = "";
            sscanf(line + 11, "%63[A-Z]", t);
            title = t;
            continue;
        }
        if (!strncmp(line, "-----END ", 9)) { title.clear(); continue; }
        char key[64];
        double val;
        if (!title.empty() && sscanf(line, " %63[^:]: %lf", key, &val) == 2) blocks[title][key] = val;
    }
    auto get = [&](const char *b, const char *k, double &out) {
        auto it = blocks.find(b);
        if (it == blocks.end()) { msg = std::string("missing parameter block ") + b; return false; }
        auto jt = it->second.find(k);
        if (jt == it->second.end()) { msg = std::string("missing parameter ") + b + "." + k; return false; }
        out = jt->second;
        return true;
    };
    double v;
    memset(&p, 0, sizeof p);
    if (!get("GATEBOOTSPARAMS", "ks_basebit", v)) return false; p.ks_basebit = (int)v;
    if (!get("GATEBOOTSPARAMS", "ks_t", v)) return false; p.ks_t = (int)v;
    if (!get("LWEPARAMS", "n", v)) return false; p.n = (int)v;
    if (!get("LWEPARAMS", "alpha_min", v)) return false; p.ks_stdev = v;
    if (!get("LWEPARAMS", "alpha_max", v)) return false; p.max_stdev = v;
    if (!get("TLWEPARAMS", "N", v)) return false; p.N = (int)v;
    if (!get("TLWEPARAMS", "k", v)) return false; p.k = (int)v;
    if (!get("TLWEPARAMS", "alpha_min", v)) return false; p.bk_stdev = v;
    if (!get("TGSWPARAMS", "l", v)) return false; p.bk_l = (int)v;
    if (!get("TGSWPARAMS", "Bgbit", v)) return false; p.bk_Bgbit = (int)v;
    return true;
}

static void write_param_blocks(FILE *f, const ieache_params &p)
{
    fprintf(f, "-----BEGIN GATEBOOTSPARAMS-----\nks_basebit: %d\nks_t: %d\n-----END GATEBOOTSPARAMS-----\n", p.ks_basebit, p.ks_t);
    fprintf(f, "-----BEGIN LWEPARAMS-----\nalpha_max: %.17g\nalpha_min: %.17g\nn: %d\n-----END LWEPARAMS-----\n", p.max_stdev, p.ks_stdev, p.n);
    fprintf(f, "-----BEGIN TLWEPARAMS-----\nN: %d\nalpha_max: %.17g\nalpha_min: %.17g\nk: %d\n-----END TLWEPARAMS-----\n", p.N, p.max_stdev, p.bk_stdev, p.k);
    fprintf(f, "-----BEGIN TGSWPARAMS-----\nBgbit: %d\nl: %d\n-----END TGSWPARAMS-----\n", p.bk_Bgbit, p.bk_l);
}

static bool rd(FILE *f, void *dst, size_t bytes) { return fread(dst, 1, bytes, f) == bytes; }

int read_keyset(const char *path, HostKeySet &ks, bool want_bk, std::string &msg)
{
    FILE *f = fopen(path, "rb");
    if (!f) { msg = std::string("cannot open ") + path; return IEACHE_ERR_IO; }
    int rc = read_keyset_stream(f, ks, want_bk, msg);
    fclose(f);
    if (rc) msg += std::string(" (") + path + ")";
    return rc;
}

int read_keyset_stream(FILE *f, HostKeySet &ks, bool want_bk, std::string &msg)
{
    if (!read_param_blocks(f, ks.p, msg)) return IEACHE_ERR_FORMAT;
    const ieache_params &p = ks.p;
    /* bounds before anything is shifted or allocated from these numbers (a hostile or damaged header must not turn
     * into 1 << 40 or a terabyte resize): the widest values libtfhe's own parameter sets use are far inside them */
    if (p.n <= 0 || p.N <= 0 || p.k <= 0 || p.bk_l <= 0 || p.ks_t <= 0 || p.ks_basebit <= 0 || p.bk_Bgbit <= 0 || p.n > 4096 ||
        p.N > 4096 || p.k > 4 || p.bk_l > 16 || p.bk_Bgbit > 31 || p.bk_l * p.bk_Bgbit > 32 || p.ks_basebit > 8 || p.ks_t > 32 ||
        p.ks_basebit * p.ks_t > 32) {
        msg = "implausible parameters in key file"; return IEACHE_ERR_FORMAT;
    }
    const int n = p.n, N = p.N, k = p.k, kpl = (k + 1) * p.bk_l, t = p.ks_t, base = 1 << p.ks_basebit;
    const size_t nks = (size_t)k * N * t * base;
    int32_t id;
    double var;
    bool ok = rd(f, &id, 4) && rd(f, &id, 4);
    std::vector<int32_t> scratch;
    try {
        if (want_bk) { ks.ksk.resize(nks * (n + 1)); ks.bk.resize((size_t)n * kpl * (k + 1) * N); }
        scratch.resize((size_t)(k + 1) * N > (size_t)n + 1 ? (size_t)(k + 1) * N : (size_t)n + 1);
    } catch (const std::bad_alloc &) {
        msg = "out of memory for the key arrays"; return IEACHE_ERR_NOMEM;
    }
    for (size_t s = 0; ok && s < nks; s++) {
        int32_t *dst = want_bk ? &ks.ksk[s * (n + 1)] : scratch.data();
        ok = rd(f, &id, 4) && rd(f, dst, 4 * (size_t)(n + 1)) && rd(f, &var, 8);
    }
    for (int i = 0; ok && i < n; i++) {
        ok = rd(f, &id, 4);
        for (int r = 0; ok && r < kpl; r++) {
            int32_t *dst = want_bk ? &ks.bk[((size_t)i * kpl + r) * (k + 1) * N] : scratch.data();
            ok = rd(f, &id, 4) && rd(f, dst, 4 * (size_t)(k + 1) * N) && rd(f, &var, 8);
        }
    }
    if (!ok) { msg = "truncated key file"; return IEACHE_ERR_FORMAT; }
    /* optional secret part */
    ks.has_secret = false;
    if (rd(f, &id, 4)) {
        ks.lwe_key.resize(n); ks.tlwe_key.resize((size_t)k * N);
        if (!(rd(f, ks.lwe_key.data(), 4 * (size_t)n) && rd(f, &id, 4) && rd(f, ks.tlwe_key.data(), 4 * (size_t)k * N))) {
            msg = "truncated secret part of key file"; return IEACHE_ERR_FORMAT;
        }
        ks.has_secret = true;
    }
    return IEACHE_OK;
}

int write_keyset(const char *path, const HostKeySet &ks, bool with_secret, std::string &msg)
{
    FILE *f = fopen(path, "wb");
    if (!f) { msg = std::string("cannot open ") + path; return IEACHE_ERR_IO; }
    const ieache_params &p = ks.p;
    const int n = p.n, N = p.N, k = p.k, kpl = (k + 1) * p.bk_l, t = p.ks_t, base = 1 << p.ks_basebit;
    bool ok = true; /* every write is checked: a full disk must not leave a truncated 114 MB key reported as written */
    auto put = [&](const void *src, size_t bytes) { ok = ok && fwrite(src, 1, bytes, f) == bytes; };
    write_param_blocks(f, p);
    int32_t id = UID_BK; put(&id, 4);
    id = UID_KS_KEY; put(&id, 4);
    const double var_ks = p.ks_stdev * p.ks_stdev, var_bk = p.bk_stdev * p.bk_stdev, zero = 0.0;
    for (size_t s = 0; ok && s < (size_t)k * N * t * base; s++) {
        id = UID_LWE_SAMPLE; put(&id, 4);
        put(&ks.ksk[s * (n + 1)], 4 * (size_t)(n + 1));
        put((s % base) ? &var_ks : &zero, 8);
    }
    for (int i = 0; ok && i < n; i++) {
        id = UID_TGSW_SAMPLE; put(&id, 4);
        for (int r = 0; r < kpl; r++) {
            id = UID_TLWE_SAMPLE; put(&id, 4);
            put(&ks.bk[((size_t)i * kpl + r) * (k + 1) * N], 4 * (size_t)(k + 1) * N);
            put(&var_bk, 8);
        }
    }
    if (with_secret && ks.has_secret) {
        id = UID_LWE_KEY; put(&id, 4);
        put(ks.lwe_key.data(), 4 * (size_t)n);
        id = UID_TGSW_KEY; put(&id, 4);
        put(ks.tlwe_key.data(), 4 * (size_t)k * N);
    }
    ok = ok && !ferror(f);
    if (fclose(f) != 0) ok = false;
    if (!ok) { msg = std::string("short write on ") + path; return IEACHE_ERR_IO; }
    return IEACHE_OK;
}

int read_samples(FILE *f, int n, int32_t *dst, size_t count)
{
    for (size_t c = 0; c < count; c++) {
        int32_t id;
        double var;
        if (!(rd(f, &id, 4) && rd(f, dst + c * (n + 1), 4 * (size_t)(n + 1)) && rd(f, &var, 8))) return IEACHE_ERR_FORMAT;
    }
    return IEACHE_OK;
}
int write_samples(FILE *f, int n, const int32_t *src, size_t count, double variance)
{
    for (size_t c = 0; c < count; c++) {
        int32_t id = UID_LWE_SAMPLE;
        if (fwrite(&id, 4, 1, f) != 1 || fwrite(src + c * (n + 1), 4, n + 1, f) != (size_t)(n + 1) ||
            fwrite(&variance, 8, 1, f) != 1)
            return IEACHE_ERR_IO;
    }
    return IEACHE_OK;
}

void sym_encrypt_bits(const HostKeySet &ks, const int32_t *bits, size_t count, int32_t *out)
{
    /* fresh 256-bit stream keys from the operating system for every call (csprng.h): masks from one key stream,
     * noise from the other; Box-Muller as on the device */
    RngKeys rk;
    if (rng_keys_from_os(rk)) { fprintf(stderr, "ieache: getrandom failed\n"); abort(); }
    const int n = ks.p.n;
    const int32_t mu = 1 << 29;
    uint32_t blk[16], nz[16];
    uint64_t ctr = 0;
    for (size_t c = 0; c < count; c++) {
        int32_t *s = out + c * (n + 1);
        chacha20_block(rk.secret, (uint64_t)c, RNG_ENC_NOISE, nz);
        const double u1 = ((double)nz[0] + 1.0) * (1.0 / 4294967296.0), u2 = (double)nz[1] * (1.0 / 4294967296.0);
        const double e = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2) * ks.p.ks_stdev;
        int32_t b = (bits[c] ? mu : -mu) + (int32_t)(int64_t)((e - std::floor(e + 0.5)) * 4294967296.0);
        for (int i = 0; i < n; i++) {
            if ((i & 15) == 0) chacha20_block(rk.mask, ctr++, RNG_ENC_MASK, blk);
            s[i] = (int32_t)blk[i & 15];
            b += s[i] * ks.lwe_key[i];
        }
        s[n] = b;
    }
}
void sym_decrypt_bits(const HostKeySet &ks, const int32_t *samples, size_t count, int32_t *bits)
{
    const int n = ks.p.n;
    for (size_t c = 0; c < count; c++) {
        const int32_t *s = samples + c * (n + 1);
        int32_t ph = s[n];
        for (int i = 0; i < n; i++) ph -= s[i] * ks.lwe_key[i];
        bits[c] = ph > 0 ? 1 : 0;
    }
}

} // namespace ieache
