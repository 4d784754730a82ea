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
