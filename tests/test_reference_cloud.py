"""Parity pins that do not depend on the oracle's own FFT or on our reading of the circuits:

1. oracle/exact_ref.c — blind rotation in exact integer arithmetic (schoolbook negacyclic products, no transform):
   the oracle's FP64 FFT path and the CPU emulations of the CUDA kernels must equal it bit for bit.
2. oracle/_ref/cloud_ref_oracle — the reference's own Cloud/cloud.c, compiled UNMODIFIED from /root/reference
   against include/tfhe/*.h with the 17 libtfhe symbols served by the oracle's gates (oracle/tfhe_shim.cpp).  Run on
   the same cloud.key / nbit.key / cloud.data / operator.txt as the oracle's restatement (cloud_oracle.c): because
   both issue the same gate calls in the same order on the same deterministic gate, the value blocks of answer.data
   must be IDENTICAL, not just decrypt-equal.  This pins the restated circuits, dispatch and file layout (SURVEY §8
   rows a1-a11) against the reference's source; what stays unpinned is the gate itself (libtfhe is absent).

The GPU half (oracle/_ref/cloud_ref_b200: the same cloud.c linked against libieache_b200.so) is in
tests/test_gpu_e2e_files.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_bind as ob

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ORACLE = os.path.join(ROOT, "oracle", "_ref", "cloud_ref_oracle")
EMUL = os.path.join(ROOT, "tests", "emul", "libbr_emul.so")


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------ 1. exact integer reference
@pytest.mark.parametrize("n,l,bgbit", [(8, 3, 7), (32, 3, 7), (12, 2, 10)])
def test_oracle_fft_equals_exact_integer_arithmetic(oracle, n, l, bgbit):
    p = ob.params_default(n)
    p.bk_l, p.bk_Bgbit = l, bgbit
    ks = oracle.keygen(p, seed=600 + n)
    rng = np.random.default_rng(n)
    bits = rng.integers(0, 2, 6).astype(np.int32)
    a, b = ks.encrypt(bits, 1), ks.encrypt(1 - bits, 2)
    x = np.concatenate([a + b, 2 * a + 2 * b]).astype(np.int32)
    x[:, n] -= np.int32(1 << 29)
    x[0, 0] = 0                                   # a zero mask coefficient: the step is skipped
    x[1, :] = 0
    assert (ks.bootstrap_woks(x) == ks.bootstrap_woks_exact(x)).all()
    ks.free()


@pytest.mark.parametrize("fn", ["emul_blind_rotate", "emul_warp_blind_rotate", "emul_w12f_blind_rotate"])
def test_kernel_arithmetic_equals_exact_integer_arithmetic(oracle, fn):
    """the per-thread code of the CUDA kernels (br_core.h / br_warp.h, run on the CPU by tests/emul) against exact
    arithmetic: bit-exact, so the transforms' rounding error stayed below 1/2 in every product"""
    E = ctypes.CDLL(EMUL)
    n = 24
    ks = oracle.keygen(ob.params_default(n), seed=611)
    bits = np.array([1, 0, 1, 1, 0, 0], dtype=np.int32)
    s = ks.encrypt(bits, 9)
    bk = ks.bk_coef()
    for g in range(6):
        x = (s[g] + s[(g + 1) % 6]).astype(np.int32)
        x[n] = np.int32((int(x[n]) - (1 << 29) + 2 ** 31) % 2 ** 32 - 2 ** 31)
        ext = np.zeros(1025, dtype=np.int32)
        getattr(E, fn)(n, 3, 7, ctypes.c_int32(1 << 29), vp(bk), vp(x), vp(ext))
        assert (ext == ks.bootstrap_woks_exact(x[None])[0]).all(), (fn, g)
    ks.free()


# ------------------------------------------------------------------ 2. the reference's cloud.c on the oracle's gates
needs_ref = pytest.mark.skipif(not os.path.exists(REF_ORACLE), reason="oracle/_ref/cloud_ref_oracle not built (needs /root/reference at build time)")

#        op  sign1 sign2 width a                    b
CASES = [(1, 0, 0, 32, 1 << 30, 1 << 30),          # Client1/process.c:96 stock operands, a + b
         (1, 2, 2, 32, 5, 7),                      # (-a) + (-b): sign code 1 + 2 = 3 -> 4
         (2, 0, 2, 32, 100, 23),                   # a - (-b): magnitudes add
         (2, 0, 0, 32, 1000, 1),                   # a - b
         (2, 0, 0, 32, 1, 1000),                   # a - b, negative result in two's complement
         (1, 2, 0, 32, 50, 20),                    # (-a) + b: operands swap (cloud.c:1809)
         (2, 2, 2, 32, 3, 10),                     # (-a) - (-b)
         (1, 0, 0, 64, (1 << 62) + 12345, (1 << 62) + 1),
         (2, 0, 0, 128, (1 << 100) + 5, (1 << 99) + 77),
         (1, 0, 0, 256, (1 << 255) - 19, (1 << 200) + 3),
         (2, 2, 0, 256, (1 << 250) + 1, (1 << 251) + 9),
         (4, 2, 0, 32, 77777, 99999),
         (4, 0, 0, 64, (1 << 62) + 99, (1 << 61) + 12345678901),
         (4, 0, 2, 128, (1 << 127) - 1, (1 << 126) + 987654321987654321)]


@pytest.fixture(scope="module")
def keys8(oracle):
    ks = oracle.keygen(ob.params_default(8), seed=31)
    nbit = oracle.keygen(ob.params_default(8), seed=32)
    yield ks, nbit
    ks.free(); nbit.free()


def write_request(d, oracle, ks, nbit, op, s1, s2, width, a, b):
    ks.write_cloud_key(os.path.join(d, "cloud.key"))
    nbit.write_secret_key(os.path.join(d, "nbit.key"))
    data = np.concatenate([oracle.alice(ks, nbit, s1, width, a, seed=1), oracle.alice(ks, nbit, s2, width, b, seed=2)])
    ks.write_samples(data, os.path.join(d, "cloud.data"))
    with open(os.path.join(d, "operator.txt"), "w") as f:
        f.write(str(op))
    return data


@needs_ref
@pytest.mark.parametrize("op,s1,s2,width,a,b", CASES)
def test_reference_cloud_c_equals_the_restatement(tmp_path, oracle, keys8, op, s1, s2, width, a, b):
    ks, nbit = keys8
    d = str(tmp_path)
    data = write_request(d, oracle, ks, nbit, op, s1, s2, width, a, b)
    r = subprocess.run([REF_ORACLE], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:]
    rec = 4 + 4 * 9 + 8
    assert os.path.getsize(os.path.join(d, "answer.data")) == 352 * rec          # Cloud/dragonfly_cipher_cloud.py:1292-1297
    ans_ref = ks.read_samples(os.path.join(d, "answer.data"), 352)
    rc, ans_o = oracle.cloud_main(ks, nbit, op, data)
    assert rc == 0 and len(ans_o) == 352
    # value blocks, padding blocks and the carry block: the same gate calls in the same order -> identical ciphertexts
    assert (ans_ref[64:] == ans_o[64:]).all()
    # metadata blocks are fresh encryptions under the nbit key: equal after decryption
    assert oracle.verif(ks, nbit, ans_ref) == oracle.verif(ks, nbit, ans_o)
    code, w, chunks = oracle.verif(ks, nbit, ans_ref)
    va, vb = (-a if s1 == 2 else a), (-b if s2 == 2 else b)
    assert ob.decode_result(op, code, w, chunks) == {1: va + vb, 2: va - vb, 4: va * vb}[op]


@needs_ref
@pytest.mark.parametrize("code1,code2", [(4, 0), (4, 2), (1, 2), (3, 0)])
def test_reference_cloud_c_sign_codes_of_chained_operands(tmp_path, oracle, keys8, code1, code2):
    """an operand block can carry a sign code the client never sends (4 = both factors negative, written by an earlier
    operator of a chain, Cloud/dragonfly_cipher_cloud.py:1306-1315): cloud.c:812-821 writes 1, 2 or 4 for the sums
    1, 2, 3 and 0 for every other sum; the dispatch sees the raw sum"""
    ks, nbit = keys8
    d = str(tmp_path)
    data = write_request(d, oracle, ks, nbit, 1, code1, code2, 32, 1234, 99)
    r = subprocess.run([REF_ORACLE], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:]
    ans_ref = ks.read_samples(os.path.join(d, "answer.data"), 352)
    rc, ans_o = oracle.cloud_main(ks, nbit, 1, data)
    assert rc == 0
    assert oracle.verif(ks, nbit, ans_ref) == oracle.verif(ks, nbit, ans_o)
    n1 = 1 if code1 == 2 else code1
    assert oracle.verif(ks, nbit, ans_ref)[0] == {1: 1, 2: 2, 3: 4}.get(n1 + code2, 0)
    if len(ans_ref) == len(ans_o):
        assert (ans_ref[64:] == ans_o[64:]).all()


@needs_ref
def test_reference_cloud_c_abort_path(tmp_path, oracle, keys8):
    """multiply at width 256: exit 126 and a 64-record answer.data (Cloud/cloud.c:860-864)"""
    ks, nbit = keys8
    d = str(tmp_path)
    data = write_request(d, oracle, ks, nbit, 4, 0, 0, 256, 3, 5)
    r = subprocess.run([REF_ORACLE], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 126
    assert os.path.getsize(os.path.join(d, "answer.data")) == 64 * (4 + 4 * 9 + 8)
    rc, ans_o = oracle.cloud_main(ks, nbit, 4, data)
    assert rc == 126 and len(ans_o) == 64
