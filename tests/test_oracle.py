"""CPU tests of the oracle (the checker): gate truth tables under decryption, noise statistics,
file round-trips, the reference's structural invariants (SURVEY.md §8c "what can still be pinned"),
and the committed golden vectors."""
import os

import numpy as np
import pytest

import oracle_bind as ob

TT = {
    "NAND": [1, 1, 1, 0], "OR": [0, 1, 1, 1], "AND": [0, 0, 0, 1], "XOR": [0, 1, 1, 0], "XNOR": [1, 0, 0, 1],
    "NOR": [1, 0, 0, 0], "ANDNY": [0, 1, 0, 0], "ANDYN": [0, 0, 1, 0], "ORNY": [1, 1, 0, 1], "ORYN": [1, 0, 1, 1],
}
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "gates_n16_seed4242.npz")


@pytest.fixture(scope="module")
def ks_small(oracle):
    ks = oracle.keygen(ob.params_default(24), seed=99)
    yield ks
    ks.free()


def test_default_parameters():
    p = ob.params_default()
    # Keygen/keygen.c:22-23 -> tfhe master lambda=110 set
    assert (p.n, p.N, p.k, p.bk_l, p.bk_Bgbit, p.ks_t, p.ks_basebit) == (630, 1024, 1, 3, 7, 8, 2)
    assert p.ks_stdev == 2.0 ** -15 and p.bk_stdev == 2.0 ** -25


@pytest.mark.parametrize("name", list(TT))
def test_truth_tables_small(ks_small, name):
    a = ks_small.encrypt([0, 0, 1, 1], 1)
    b = ks_small.encrypt([0, 1, 0, 1], 2)
    out = ks_small.gate_batch(ob.OPS[name], a, b, threads=4)
    assert list(ks_small.decrypt(out)) == TT[name]
    ph = ks_small.phase(out).astype(np.float64) / 2 ** 32
    assert np.abs(np.abs(ph) - 0.125).max() < 0.03  # fresh bootstrap output: +-1/8 plus small noise


def test_mux_not_copy_const(ks_small):
    bits = [(a, b, c) for a in (0, 1) for b in (0, 1) for c in (0, 1)]
    a = ks_small.encrypt([t[0] for t in bits], 3)
    b = ks_small.encrypt([t[1] for t in bits], 4)
    c = ks_small.encrypt([t[2] for t in bits], 5)
    out = ks_small.gate_batch(ob.OPS["MUX"], a, b, c, threads=4)
    assert list(ks_small.decrypt(out)) == [(t[1] if t[0] else t[2]) for t in bits]
    assert list(ks_small.decrypt(ks_small.gate_batch(ob.OPS["NOT"], a))) == [1 - t[0] for t in bits]
    assert (ks_small.gate_batch(ob.OPS["COPY"], a) == a).all()


def test_one_gate_default_parameters(oracle):
    ks = oracle.keygen(ob.params_default(630), seed=5)
    a, b = ks.encrypt([1, 1], 1), ks.encrypt([0, 1], 2)
    out = ks.gate_batch(ob.OPS["NAND"], a, b, threads=2)
    assert list(ks.decrypt(out)) == [1, 0]
    ks.free()


def test_bootstrap_noise_statistics(ks_small):
    # output noise of a bootstrapped gate stays far inside the 1/8 decision margin
    bits = np.tile([0, 1], 32).astype(np.int32)
    a, b = ks_small.encrypt(bits, 7), ks_small.encrypt(bits[::-1].copy(), 8)
    out = ks_small.gate_batch(ob.OPS["XOR"], a, b, threads=8)
    ph = ks_small.phase(out).astype(np.float64) / 2 ** 32
    assert np.std(np.abs(ph) - 0.125) < 0.01


def test_stage_composition(ks_small):
    # gate == keyswitch(bootstrap_woks(precombination)) (tfhe_bootstrap_FFT structure, App. A)
    a, b = ks_small.encrypt([1, 0, 1], 1), ks_small.encrypt([1, 1, 0], 2)
    x = (a + b).astype(np.int32)
    x[:, ks_small.n] -= 1 << 29  # AND: (0,-1/8) + ca + cb
    got = ks_small.keyswitch(ks_small.bootstrap_woks(x))
    want = ks_small.gate_batch(ob.OPS["AND"], a, b, threads=1)
    assert (got == want).all()


def test_file_roundtrip_and_sizes(tmp_path, oracle):
    """Structural invariants from the reference: 2536-byte LWE record (n=630), 352-record client
    file = 892672 B, two-operand cloud.data = 1785344 B, abort file = 162304 B
    (Cloud/dragonfly_cipher_cloud.py:1295, Output/output_dynamic.py:1018)."""
    ks = oracle.keygen(ob.params_default(630), seed=11)
    nbit = ks  # sizes do not depend on which key encrypts the metadata
    blk = oracle.alice(ks, nbit, 0, 32, 1 << 30, seed=3)
    path = str(tmp_path / "cloud.data")
    ks.write_samples(blk, path)
    assert os.path.getsize(path) == 352 * 2536 == 892672
    ks.write_samples(blk, path, append=True)
    assert os.path.getsize(path) == 704 * 2536 == 1785344
    back = ks.read_samples(path, 704)
    assert (back[:352] == blk).all() and (back[352:] == blk).all()
    ab = str(tmp_path / "answer.data")
    ks.write_samples(blk[:64], ab)
    assert os.path.getsize(ab) == 162304
    # key files round-trip through the libtfhe-style text+binary format
    kp = str(tmp_path / "secret.key")
    ks.write_secret_key(kp)
    ks2 = oracle.read_key(kp)
    assert (ks2.lwe_key() == ks.lwe_key()).all() and (ks2.bk_coef() == ks.bk_coef()).all() and (ks2.ksk() == ks.ksk()).all()
    cp = str(tmp_path / "cloud.key")
    ks.write_cloud_key(cp)
    assert os.path.getsize(kp) - os.path.getsize(cp) == 4 + 630 * 4 + 4 + 1024 * 4
    ks.free(); ks2.free()


def test_golden_vectors(oracle):
    """tests/golden/gates_n16_seed4242.npz (made by tests/golden/make_golden.py from this oracle):
    pins the oracle's keygen, encryption and every gate against silent change."""
    g = np.load(GOLDEN)
    ks = oracle.keygen(ob.params_default(int(g["n"])), seed=int(g["seed"]))
    assert (ks.lwe_key() == g["lwe_key"]).all()
    a, b, c = ks.encrypt(g["bits_a"], 1), ks.encrypt(g["bits_b"], 2), ks.encrypt(g["bits_c"], 3)
    assert (a == g["a"]).all() and (b == g["b"]).all()
    for name, op in ob.OPS.items():
        if name in ("NOT", "COPY", "CONST"):
            continue
        out = ks.gate_batch(op, a, b, c if name == "MUX" else None, threads=2)
        assert (out == g["out_" + name]).all(), name
    ks.free()
