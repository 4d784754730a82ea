"""CPU tests of the product's host-side callers of the path: ieache_alice_* (Client1/alice.c)
and ieache_verif_run (Output/verif.c) against the oracle's alice / cloud main() / verif on the same key files —
exercises the product's libtfhe-format reader and writer without a GPU."""
import os

import numpy as np
import pytest

import oracle_bind as ob


@pytest.fixture(scope="module")
def keydir(tmp_path_factory, oracle):
    d = str(tmp_path_factory.mktemp("keys"))
    ks = oracle.keygen(ob.params_default(8), seed=31)
    nbit = oracle.keygen(ob.params_default(8), seed=32)
    ks.write_secret_key(os.path.join(d, "secret.key"))
    ks.write_cloud_key(os.path.join(d, "cloud.key"))
    nbit.write_secret_key(os.path.join(d, "nbit.key"))
    yield d, ks, nbit
    ks.free(); nbit.free()


def test_alice_matches_reference_layout(pkg, oracle, keydir):
    d, ks, nbit = keydir
    out = os.path.join(d, "cloud.data")
    pkg.alice_encrypt(d, 2, 64, (1 << 62) + 5, out)
    pkg.alice_encrypt(d, 0, 32, 77, out, append=True)
    assert os.path.getsize(out) == 704 * (4 + 4 * 9 + 8)
    blk = ks.read_samples(out, 704)
    assert oracle.verif(ks, nbit, blk[:352])[:2] == (2, 64)
    c = oracle.verif(ks, nbit, blk[:352])[2]
    assert c[0] | (c[1] << 32) == (1 << 62) + 5 and c[2:] == [0] * 6
    assert oracle.verif(ks, nbit, blk[352:]) == (0, 32, [77, 0, 0, 0, 0, 0, 0, 0])
    assert ks.decrypt_word(blk[320:352]) == 0      # carry block: 32 encryptions of 0 (alice.c:147-149)


def test_alice_run_reads_values_txt(pkg, oracle, keydir):
    d, ks, nbit = keydir
    # Client1/process.c:80-99: the stock 32-bit positive operand 2^30
    open(os.path.join(d, "values.txt"), "w").write("0" * 32 + "\n" + format(32, "032b") + "\n" + format(1 << 30, "032b") + "\n" + "0" * 32 + "\n")
    pkg.alice_run(d)
    blk = ks.read_samples(os.path.join(d, "cloud.data"), 352)
    assert oracle.verif(ks, nbit, blk) == (0, 32, [1 << 30, 0, 0, 0, 0, 0, 0, 0])


CASES = [(1, 0, 0, 32, 1 << 30, 1 << 30), (1, 2, 2, 32, 5, 7), (2, 0, 0, 32, 1, 1000), (2, 2, 0, 32, 9, 4), (1, 0, 2, 32, 50, 20),
         (2, 2, 2, 32, 3, 10), (1, 0, 0, 64, (1 << 62) + 12345, (1 << 62) + 1), (4, 2, 0, 32, 77777, 99999),
         (4, 0, 0, 64, (1 << 63) + 9, (1 << 63) + 7)]


@pytest.mark.parametrize("op,s1,s2,width,a,b", CASES)
def test_verif_run_against_integers(pkg, oracle, keydir, op, s1, s2, width, a, b):
    """alice (product) -> cloud main() (oracle) -> verif (product): every sign rule of Output/verif.c"""
    d, ks, nbit = keydir
    out = os.path.join(d, "cloud.data")
    pkg.alice_encrypt(d, s1, width, a, out)
    pkg.alice_encrypt(d, s2, width, b, out, append=True)
    rc, ans = oracle.cloud_main(ks, nbit, op, ks.read_samples(out, 704))
    assert rc == 0
    ks.write_samples(ans, os.path.join(d, "answer.data"))
    open(os.path.join(d, "operator.txt"), "w").write(str(op))
    value, code, w = pkg.verif_run(d)
    va, vb = (-a if s1 == 2 else a), (-b if s2 == 2 else b)
    assert value == {1: va + vb, 2: va - vb, 4: va * vb}[op]
    assert w == (2 * width if op == 4 else width)


def test_verif_run_rejects_abort_file(pkg, oracle, keydir):
    d, ks, nbit = keydir
    data = np.concatenate([oracle.alice(ks, nbit, 0, 256, 3, seed=1), oracle.alice(ks, nbit, 0, 256, 5, seed=2)])
    rc, ans = oracle.cloud_main(ks, nbit, 4, data)
    assert rc == 126
    ks.write_samples(ans, os.path.join(d, "answer.data"))
    open(os.path.join(d, "operator.txt"), "w").write("4")
    with pytest.raises(pkg.EngineError):
        pkg.verif_run(d)
