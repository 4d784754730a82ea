"""BASELINE.json config 1 end to end on the product alone, at the repo's default parameters (n = 630):
keygen -> alice -> cloud -> verif through the reference's files, every stage by libieache_b200.so
(keys on the GPU, circuits on the GPU), cross-checked by the oracle reading the same files."""
import os

import numpy as np
import pytest

import oracle_bind as ob

pytestmark = pytest.mark.gpu


def test_keygen_alice_cloud_verif_default_parameters(tmp_path, pkg, oracle):
    d = str(tmp_path)
    eng = pkg.Engine(0)
    eng.keygen_files(d)                                              # Keygen/keygen.c
    sizes = {f: os.path.getsize(os.path.join(d, f)) for f in ("secret.key", "cloud.key", "nbit.key")}
    assert sizes["secret.key"] - sizes["cloud.key"] == 4 + 630 * 4 + 4 + 1024 * 4 and sizes["nbit.key"] == sizes["secret.key"]
    a = b = 1 << 30                                                  # Client1/process.c:96
    pkg.alice_encrypt(d, 0, 32, a, os.path.join(d, "cloud.data"))   # Client1/alice.c
    pkg.alice_encrypt(d, 0, 32, b, os.path.join(d, "cloud.data"), append=True)
    assert os.path.getsize(os.path.join(d, "cloud.data")) == 1785344
    open(os.path.join(d, "operator.txt"), "w").write("1")
    rc, secs = eng.cloud_run(d)                                      # Cloud/cloud.c
    assert rc == 0 and os.path.getsize(os.path.join(d, "answer.data")) == 892672
    value, code, width = pkg.verif_run(d)                            # Output/verif.c
    assert (value, code, width) == (2147483648, 0, 32)
    # the oracle reads the product's key files and agrees on every decrypted block
    ks, nbit = oracle.read_key(os.path.join(d, "secret.key")), oracle.read_key(os.path.join(d, "nbit.key"))
    ans = ks.read_samples(os.path.join(d, "answer.data"), 352)
    assert oracle.verif(ks, nbit, ans) == (0, 32, [1 << 31, 0, 0, 0, 0, 0, 0, 0])
    # and the oracle, running one gate with the GPU-made key, decrypts correctly (key validity on the CPU side)
    x, y = ks.encrypt([0, 1, 1], 1), ks.encrypt([1, 1, 0], 2)
    assert list(ks.decrypt(ks.gate_batch(ob.OPS["AND"], x, y))) == [0, 1, 0]
    # multiply at default parameters, then the 256-bit abort path with its 162304-byte file
    open(os.path.join(d, "operator.txt"), "w").write("4")
    rc, _ = eng.cloud_run(d)
    assert rc == 0 and pkg.verif_run(d)[0] == 1 << 60
    pkg.alice_encrypt(d, 0, 256, 3, os.path.join(d, "cloud.data"))
    pkg.alice_encrypt(d, 0, 256, 5, os.path.join(d, "cloud.data"), append=True)
    rc, _ = eng.cloud_run(d)
    assert rc == 126 and os.path.getsize(os.path.join(d, "answer.data")) == 162304
    ks.free(); nbit.free(); eng.close()


def test_batched_ingest_of_request_directories(tmp_path, pkg):
    """several Cloud request directories (cloud.data + operator.txt) evaluated as one batch; each
    gets the answer.data ./cloud would have written (decoded by the verifier), unreadable ones are reported"""
    import shutil
    keys = str(tmp_path / "keys")
    os.makedirs(keys)
    eng = pkg.Engine(0)
    eng.keygen_files(keys, pkg.Params.default(16))                   # small LWE dimension: the circuits are the point here
    cases = [(1, 0, 1000, 0, 234), (2, 0, 1000, 0, 234), (4, 0, 77, 0, 1001), (1, 2, 5, 0, 9), (2, 0, 3, 0, 10), (4, 2, 12, 0, 12)]
    want = [{1: va + vb, 2: va - vb, 4: va * vb}[op] for op, s1, a, s2, b in cases for va, vb in [(-a if s1 == 2 else a, -b if s2 == 2 else b)]]
    dirs = []
    for k, (op, s1, v1, s2, v2) in enumerate(cases):
        d = str(tmp_path / f"req{k}")
        os.makedirs(d)
        for f in ("secret.key", "nbit.key"):
            shutil.copy(os.path.join(keys, f), d)                    # the verifier's and alice's keys
        pkg.alice_encrypt(d, s1, 32, v1, os.path.join(d, "cloud.data"))
        pkg.alice_encrypt(d, s2, 32, v2, os.path.join(d, "cloud.data"), append=True)
        open(os.path.join(d, "operator.txt"), "w").write(str(op))
        dirs.append(d)
    dirs.append(str(tmp_path / "missing"))
    sess = eng.session(os.path.join(keys, "cloud.key"), os.path.join(keys, "nbit.key"))
    codes, secs = sess.compute_dirs(dirs)
    assert list(codes[:-1]) == [0] * len(cases) and codes[-1] == -3 and secs > 0      # IEACHE_ERR_IO for the missing directory
    for d, (op, *_), w in zip(dirs, cases, want):
        assert os.path.getsize(os.path.join(d, "answer.data")) == 352 * (4 + 17 * 4 + 8)
        value, code, width = pkg.verif_run(d)
        assert value == w and width == (64 if op == 4 else 32), (d, value, code, width, w)
    sess.close(); eng.close()


# ------------------------------------------------------------------ the reference's own cloud.c on this library
REF_B200 = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "cloud_ref_b200")
REF_CASES = [(1, 0, 0, 32, 1 << 30, 1 << 30), (1, 2, 2, 32, 5, 7), (2, 0, 0, 32, 1, 1000), (1, 2, 0, 32, 50, 20),
             (2, 0, 0, 64, (1 << 63) + 5, (1 << 62) + 77), (4, 2, 0, 32, 77777, 99999),
             # the wide multipliers through the reference's own main(): 2 x mul64 + split, and 4 x mul128 + 15 chained adds
             (4, 0, 0, 64, (1 << 62) + 3, (1 << 61) + 5), (4, 0, 2, 128, (1 << 127) - 1, (1 << 126) + 987654321)]


@pytest.mark.skipif(not os.path.exists(REF_B200), reason="oracle/_ref/cloud_ref_b200 not built (needs /root/reference at build time)")
@pytest.mark.parametrize("op,s1,s2,width,a,b", REF_CASES)
def test_unmodified_reference_cloud_c_on_the_gpu_library(tmp_path, pkg, oracle, op, s1, s2, width, a, b):
    """SURVEY 8 b1 executed: /root/reference/Cloud/cloud.c compiled UNMODIFIED against include/tfhe/*.h and linked with
    libieache_b200.so instead of libtfhe (oracle/Makefile), run as ./cloud is run (Cloud/dragonfly_cipher_cloud.py:1233)
    on oracle-written cloud.key / nbit.key / cloud.data / operator.txt.  Its answer.data must decrypt to what
    ieache_cloud_run writes for the same files (the levelised GPU circuits) and to what the oracle's cloud main() gives."""
    import subprocess
    from test_reference_cloud import write_request
    ks = oracle.keygen(ob.params_default(8), seed=31)
    nbit = oracle.keygen(ob.params_default(8), seed=32)
    d = str(tmp_path)
    data = write_request(d, oracle, ks, nbit, op, s1, s2, width, a, b)
    r = subprocess.run([REF_B200], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "Computation Time:" in r.stdout
    ans_ref = ks.read_samples(os.path.join(d, "answer.data"), 352)
    assert len(ans_ref) == 352
    eng = pkg.Engine(0)
    rc, _ = eng.cloud_run(d)
    assert rc == 0
    ans_gpu = ks.read_samples(os.path.join(d, "answer.data"), 352)
    rc_o, ans_o = oracle.cloud_main(ks, nbit, op, data)
    v = oracle.verif(ks, nbit, ans_ref)
    assert v == oracle.verif(ks, nbit, ans_gpu) == oracle.verif(ks, nbit, ans_o)
    va, vb = (-a if s1 == 2 else a), (-b if s2 == 2 else b)
    assert ob.decode_result(op, *v) == {1: va + vb, 2: va - vb, 4: va * vb}[op]
    # padding and carry blocks are verbatim copies of operand 1's carry block in all three (cloud.c:901-916)
    nres = (2 * width if op == 4 else width) // 32
    for ans in (ans_ref, ans_gpu, ans_o):
        assert (ans[(2 + nres) * 32:(3 + nres) * 32] == data[320:352]).all() and (ans[320:352] == data[320:352]).all()
    eng.close(); ks.free(); nbit.free()
