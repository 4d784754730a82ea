"""Host-side circuit layer (no GPU): the levelised DAGs built by libieache_b200.so must have exactly the
gate counts and depths of Cloud/cloud.c's circuits, and the oracle's
one-gate-at-a-time restatement of the same circuits must agree with plain integer arithmetic."""
import numpy as np
import pytest

import oracle_bind as ob

# (kind, width) -> (bootstraps, AND, XOR, levels, max level width)   SURVEY.md App. B
APP_B = {
    (1, 32): (160, 32, 128, 96, 2), (1, 64): (320, 64, 256, 192, 2), (1, 128): (640, 128, 512, 384, 2),
    (1, 256): (1280, 256, 1024, 768, 2),
    (2, 32): (320, 64, 256, 98, 4), (2, 64): (640, 128, 512, 194, 4), (2, 128): (1280, 256, 1024, 386, 4),
    (2, 256): (2560, 512, 2048, 770, 4),
    (4, 32): (11264, 3072, 8192, 255, 1056), (4, 64): (35296, 10336, 24960, 449, 4160),
    (4, 128): (121184, 37344, 83840, 1601, 16512),
    (5, 32): (11584, 3136, 8448, 257, 1057),
}


@pytest.mark.parametrize("kind,width", list(APP_B))
def test_circuit_statistics(pkg, kind, width):
    c = pkg.Circuit(kind, width)
    assert (c.bootstraps, c.and_gates, c.xor_gates, c.levels, c.max_width) == APP_B[(kind, width)]
    nc = width // 32
    if kind == 5:
        assert (c.n_inputs, c.n_outputs) == (160, 64)
    else:
        assert c.n_inputs == (2 * nc + 1) * 32
        assert c.n_outputs == (2 * nc if kind == 4 else nc) * 32


def test_unsupported_circuits(pkg):
    for kind, width in ((4, 256), (1, 48), (1, 512), (5, 64), (9, 32)):
        with pytest.raises(pkg.EngineError):
            pkg.Circuit(kind, width)


@pytest.fixture(scope="module")
def keys(oracle):
    ks = oracle.keygen(ob.params_default(8), seed=31)
    nbit = oracle.keygen(ob.params_default(8), seed=32)
    yield ks, nbit
    ks.free(); nbit.free()


def test_oracle_add_counts_and_value(oracle, keys):
    ks, _ = keys
    x, y = 0x40000000, 0x40000000   # Client1/process.c:96: stock magnitude 2^30
    oracle.stats_reset()
    s, co = ks.add(ks.encrypt_word(x, 1), ks.encrypt_word(y, 2), ks.encrypt_word(0, 3))
    assert oracle.stats_bootstraps() == 160
    assert ks.decrypt_word(s) == 2147483648 and ks.decrypt(co[:1])[0] == 0
    s, co = ks.add(ks.encrypt_word(0xFFFFFFFF, 4), ks.encrypt_word(1, 5), ks.encrypt_word(0, 6))
    assert ks.decrypt_word(s) == 0 and ks.decrypt(co[:1])[0] == 1


def test_oracle_mul32(oracle, keys):
    ks, _ = keys
    a, b = 0x9E3779B9, 0x7F4A7C15
    oracle.stats_reset()
    hi, lo = ks.mul32(ks.encrypt_word(a, 1), ks.encrypt_word(b, 2), ks.encrypt_word(0, 3))
    assert oracle.stats_bootstraps() == 11264
    assert (ks.decrypt_word(hi) << 32) | ks.decrypt_word(lo) == a * b


CASES = [  # (op, sign1, sign2, width, A, B)   signs as the client sends them: 0 positive, 2 negative
    (1, 0, 0, 32, 1 << 30, 1 << 30), (1, 2, 2, 32, 5, 7), (2, 0, 2, 32, 100, 23), (2, 2, 0, 32, 9, 4),
    (2, 0, 0, 32, 1000, 1), (2, 0, 0, 32, 1, 1000), (1, 0, 2, 32, 50, 20), (1, 2, 0, 32, 50, 20), (2, 2, 2, 32, 3, 10),
    (1, 0, 0, 64, (1 << 62) + 12345, (1 << 62) + 1), (2, 0, 0, 64, (1 << 63), 1),
    (4, 0, 0, 32, 1 << 30, 1 << 30), (4, 2, 0, 32, 77777, 99999),
]


def _expected(op, s1, s2, a, b):
    va, vb = (-a if s1 == 2 else a), (-b if s2 == 2 else b)
    return {1: va + vb, 2: va - vb, 4: va * vb}[op]


@pytest.mark.parametrize("op,s1,s2,width,a,b", CASES)
def test_oracle_cloud_main(oracle, keys, op, s1, s2, width, a, b):
    """keygen -> alice -> cloud -> verif on the oracle (BASELINE.json config 1 and its siblings):
    every dispatch branch of Cloud/cloud.c main() against Python integers."""
    ks, nbit = keys
    data = np.concatenate([oracle.alice(ks, nbit, s1, width, a, seed=1), oracle.alice(ks, nbit, s2, width, b, seed=2)])
    rc, ans = oracle.cloud_main(ks, nbit, op, data)
    assert rc == 0 and len(ans) == 352
    code, w, chunks = oracle.verif(ks, nbit, ans)
    assert w == (2 * width if op == 4 else width)
    got = ob.decode_result(op, code, w, chunks)
    want = _expected(op, s1, s2, a, b)
    if op != 4 and width == 32 and not (-(1 << 31) <= want < (1 << 32)):
        pytest.skip("outside what verif.c can represent")
    assert got == want, (code, w, [hex(c) for c in chunks])


def test_oracle_abort_path(oracle, keys):
    ks, nbit = keys
    data = np.concatenate([oracle.alice(ks, nbit, 0, 256, 3, seed=1), oracle.alice(ks, nbit, 0, 256, 5, seed=2)])
    rc, ans = oracle.cloud_main(ks, nbit, 4, data)
    assert rc == 126 and len(ans) == 64   # Cloud/cloud.c:860-864; 64 x 2536 B = 162304 B at n = 630
