"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI of
libieache_b200.so, against the CPU oracle on the same seeded inputs, against the committed golden
vectors, and — at sizes the oracle cannot reach — against plain integer arithmetic and
size-independent properties.

Bar: decrypted bits identical; key switching (integer) bit-exact; per-gate output
phase within 2^-13 of the torus of the oracle's (FFT rounding may flip a key-switch digit, which costs
about one key-switch noise draw ~2^-15; the decision margin is 1/8)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_bind as ob

pytestmark = pytest.mark.gpu

PHASE_TOL = 1 << 19  # 2^-13 of the torus, in Torus32 units (one flipped key-switch digit costs about 2^-15 * sqrt 2)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN_OPS = ["NAND", "OR", "AND", "XOR", "XNOR", "NOR", "ANDNY", "ANDYN", "ORNY", "ORYN"]


def torus_dist(a, b):
    d = a.astype(np.int64) - b.astype(np.int64)
    return np.abs(((d + 2 ** 31) % 2 ** 32) - 2 ** 31)


@pytest.fixture(scope="module")
def eng(pkg):
    e = pkg.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def full(pkg, oracle, eng):
    """default parameters (n = 630): oracle key set + the same key on the GPU"""
    ks = oracle.keygen(ob.params_default(630), seed=2024)
    key = eng.cloud_key_from_arrays(pkg.Params.default(630), ks.bk_coef(), ks.ksk())
    yield ks, key
    key.close(); ks.free()


@pytest.fixture(scope="module")
def small(pkg, oracle, eng):
    """n = 8: the oracle finishes whole multiplier circuits in seconds"""
    ks = oracle.keygen(ob.params_default(8), seed=31)
    nbit = oracle.keygen(ob.params_default(8), seed=32)
    key = eng.cloud_key_from_arrays(pkg.Params.default(8), ks.bk_coef(), ks.ksk())
    yield ks, nbit, key
    key.close(); ks.free(); nbit.free()


@pytest.fixture(params=["cluster_kernel", "pair_kernel", "group_kernel", "w12_kernel"])
def kernel_mode(pkg, eng, request):
    """every blind-rotation kernel (one gate on a 2-SM cluster / on two groups of one CTA / on one 64-thread CTA with
    the accumulators in registers / on one warp of the persistent 12-warp kernel with the accumulators in tensor
    memory) must pass the same parity tests whatever the launch size"""
    throughput = request.param in ("group_kernel", "w12_kernel")
    old = eng.set_pair_max(0 if throughput else 1 << 40)
    oldc = eng.set_cluster_max(1 << 40 if request.param == "cluster_kernel" else 0)
    oldv = eng.set_throughput_kernel({"w12_kernel": pkg.KERNEL_W12, "group_kernel": pkg.KERNEL_GROUP}.get(request.param, 0))
    yield request.param
    eng.set_pair_max(old)
    eng.set_cluster_max(oldc)
    eng.set_throughput_kernel(oldv)


# ------------------------------------------------------------------ gates
@pytest.mark.parametrize("op", BIN_OPS + ["MUX"])
def test_gate_parity_default_params(eng, full, op, kernel_mode):
    ks, key = full
    rng = np.random.default_rng(abs(hash(op)) % 1000)
    ba, bb, bc = (rng.integers(0, 2, 16).astype(np.int32) for _ in range(3))
    a, b, c = ks.encrypt(ba, 1), ks.encrypt(bb, 2), ks.encrypt(bc, 3)
    got = eng.gate_batch(key, op, a, b, c if op == "MUX" else None)
    want = ks.gate_batch(ob.OPS[op], a, b, c if op == "MUX" else None)
    assert (ks.decrypt(got) == ks.decrypt(want)).all()
    assert torus_dist(ks.phase(got), ks.phase(want)).max() < PHASE_TOL


def test_free_gates_are_exact(eng, full):
    ks, key = full
    a = ks.encrypt([0, 1, 1, 0], 9)
    assert (eng.gate_batch(key, "NOT", a) == ks.gate_batch(ob.OPS["NOT"], a)).all()
    assert (eng.gate_batch(key, "COPY", a) == a).all()
    c1 = eng.gate_batch(key, "CONST", imm=1, count=3)
    assert (c1[:, :630] == 0).all() and (c1[:, 630] == 1 << 29).all()
    c0 = eng.gate_batch(key, "CONST", imm=0, count=2)
    assert (c0[:, 630] == -(1 << 29)).all()


def test_golden_vectors(pkg, oracle, eng, kernel_mode):
    g = np.load(os.path.join(ROOT, "tests", "golden", "gates_n16_seed4242.npz"))
    ks = oracle.keygen(ob.params_default(int(g["n"])), seed=int(g["seed"]))
    key = eng.cloud_key_from_arrays(pkg.Params.default(int(g["n"])), ks.bk_coef(), ks.ksk())
    a, b, c = (np.ascontiguousarray(g[k]) for k in ("a", "b", "c"))
    for op in BIN_OPS + ["MUX"]:
        got = eng.gate_batch(key, op, a, b, c if op == "MUX" else None)
        want = np.ascontiguousarray(g["out_" + op])
        assert (ks.decrypt(got) == ks.decrypt(want)).all(), op
        assert torus_dist(ks.phase(got), ks.phase(want)).max() < PHASE_TOL, op
    key.close(); ks.free()


def test_stage_parity(eng, full, kernel_mode):
    """blind rotation + extraction against the oracle (phase tolerance), key switch bit-exact"""
    ks, key = full
    a, b = ks.encrypt([1, 0, 1, 1, 0, 1, 0, 0], 1), ks.encrypt([1, 1, 0, 1, 0, 0, 1, 0], 2)
    x = (a + b).astype(np.int32)
    x[:, 630] -= 1 << 29
    ext_gpu = eng.bootstrap_woks(key, x)
    ext_cpu = ks.bootstrap_woks(x)
    assert torus_dist(ks.phase_extracted(ext_gpu), ks.phase_extracted(ext_cpu)).max() < 1 << 12
    assert torus_dist(ext_gpu, ext_cpu).max() < 1 << 12      # coefficient-wise, ~2^-20 of the torus
    assert (eng.keyswitch(key, ext_cpu) == ks.keyswitch(ext_cpu)).all()  # integer path: bit-exact


@pytest.mark.parametrize("n", [8, 32])
def test_blind_rotation_equals_exact_integer_arithmetic(pkg, oracle, eng, kernel_mode, n):
    """every blind-rotation kernel against oracle/exact_ref.c: schoolbook negacyclic products of the gadget digits with
    the coefficient-domain key in wrap-around integer arithmetic, no transform at all.  The CUDA transforms are
    checked against arithmetic here, not against another FP64 FFT: bound 0 (bit-exact) — every one of the
    n x 6 x 2 polynomial products came out with rounding error below 1/2."""
    ks = oracle.keygen(ob.params_default(n), seed=900 + n)
    key = eng.cloud_key_from_arrays(pkg.Params.default(n), ks.bk_coef(), ks.ksk())
    rng = np.random.default_rng(n)
    bits = rng.integers(0, 2, 12).astype(np.int32)
    a, b = ks.encrypt(bits, 1), ks.encrypt(1 - bits, 2)
    x = np.concatenate([(a + b), (2 * a + 2 * b), (-a - b)]).astype(np.int32)    # AND-, XOR- and NOR-shaped inputs
    x[:, n] += np.int32(1 << 29)
    # a mask with a zero and with both extremes of the mod-switched range, and an all-zero sample
    x[0, 0], x[1, 0], x[2, 0] = 0, np.int32(-(1 << 31)), np.int32((1 << 31) - 1)
    x[3, :] = 0
    want = ks.bootstrap_woks_exact(x)
    got = eng.bootstrap_woks(key, x)
    assert (got == want).all(), f"max coefficient difference {torus_dist(got, want).max()} (Torus32 units)"
    assert (ks.bootstrap_woks(x) == want).all()                                  # the oracle's own FP64 FFT, same bound
    key.close(); ks.free()


def test_one_default_parameter_gate_equals_exact_integer_arithmetic(eng, full, kernel_mode):
    """the same at n = 630 for one sample (8 G integer multiply-adds on the CPU): 630 chained CMux steps, bit-exact"""
    ks, key = full
    x = (ks.encrypt([1], 77)[0] + ks.encrypt([0], 78)[0]).astype(np.int32)[None]
    x[:, 630] -= np.int32(1 << 29)
    assert (eng.bootstrap_woks(key, x) == ks.bootstrap_woks_exact(x)).all()


def test_default_kernel_selection(pkg, eng, full):
    """the launch policy at its boundaries: cluster <= 74 < two-group <= 296 < group kernel (two-group
    when the last 592-gate wave would be under 90 % full) < 900 <= persistent kernel; staged key switch from 1000"""
    _, key = full
    K = pkg
    expect = {1: K.KERNEL_CLUSTER, 74: K.KERNEL_CLUSTER, 75: K.KERNEL_PAIR, 296: K.KERNEL_PAIR, 297: K.KERNEL_PAIR, 532: K.KERNEL_PAIR,
              533: K.KERNEL_GROUP, 592: K.KERNEL_GROUP, 593: K.KERNEL_PAIR, 899: K.KERNEL_PAIR, 900: K.KERNEL_W12, 1 << 16: K.KERNEL_W12}
    for count, kern in expect.items():
        assert eng.pick_kernels(key, count)[0] == kern, count
    assert [eng.pick_kernels(key, c)[1] for c in (296, 297, 999, 1000)] == [K.KS_CLUSTER, K.KS_GATHER, K.KS_GATHER, K.KS_STAGED]
    # the selection really is what runs: a 900-gate batch through the default policy, checked by decryption
    ks, _ = full
    bits = (np.arange(900) % 2).astype(np.int32)
    out = eng.gate_batch(key, "NAND", ks.encrypt(bits, 41), ks.encrypt(1 - bits, 42))
    assert (ks.decrypt(out) == 1).all()
    # settings the library must refuse
    with pytest.raises(pkg.EngineError):
        eng.set_throughput_kernel(60)
    with pytest.raises(pkg.EngineError):
        eng.set_pair_max(-5)
    assert eng.set_throughput_kernel(0) == 0


def test_two_contexts_do_not_share_tuning(pkg, full):
    """the launch policy belongs to the context (it used to be process-wide)"""
    e1, e2 = pkg.Engine(0), pkg.Engine(0)
    try:
        e1.set_pair_max(0); e1.set_throughput_kernel(pkg.KERNEL_GROUP)
        _, key = full
        assert e1.pick_kernels(key, 100)[0] == pkg.KERNEL_GROUP
        assert e2.pick_kernels(key, 100)[0] == pkg.KERNEL_PAIR
    finally:
        e1.close(); e2.close()


def test_two_devices_in_one_process(pkg, oracle):
    """contexts on GPUs 0 and 1 of one process, 2 000 NANDs each through every kernel size class: function attributes
    (dynamic shared memory opt-in) are per device and must be set on both.  Skips on a one-GPU box."""
    import ctypes
    cudart = ctypes.CDLL("libcudart.so", mode=ctypes.RTLD_GLOBAL) if False else None
    try:
        e1 = pkg.Engine(1)
    except pkg.EngineError:
        pytest.skip("needs two GPUs")
    e0 = pkg.Engine(0)
    try:
        for e in (e0, e1, e0):
            sk, key = e.keygen(pkg.Params.default(630), seed=5)
            for count in (2000, 50):
                rng = np.random.default_rng(count)
                ba, bb = rng.integers(0, 2, count).astype(np.int32), rng.integers(0, 2, count).astype(np.int32)
                da, db, do = (e.device_alloc(count * 632 * 4) for _ in range(3))
                sk.encrypt_to_device(ba, da, seed=1); sk.encrypt_to_device(bb, db, seed=2)
                e.gate_batch_device(key, "NAND", do, da, db, count=count)
                assert (sk.decrypt_from_device(do, count) == 1 - (ba & bb)).all()
                for ptr in (da, db, do):
                    e.device_free(ptr)
            key.close(); sk.close()
    finally:
        e0.close(); e1.close()


def test_keyswitch_kernels_bit_exact(pkg, eng, full):
    """the three key-switch kernels (8-CTA cluster, per-gate gather, staged row blocks) are integer sums of the
    same rows: identical to the oracle and to each other, also for a count that leaves idle groups in a CTA"""
    ks, key = full
    rng = np.random.default_rng(5)
    ext = rng.integers(-2 ** 31, 2 ** 31, size=(29, 1025), dtype=np.int64).astype(np.int32)
    want = ks.keyswitch(ext)
    old_w, old_s = eng.set_pair_max(1 << 40), eng.set_ks_staged_min(1 << 40)
    try:
        assert eng.pick_kernels(key, 29)[1] == pkg.KS_CLUSTER
        assert (eng.keyswitch(key, ext) == want).all()           # cluster kernel (narrow launch)
        eng.set_pair_max(0)
        assert eng.pick_kernels(key, 29)[1] == pkg.KS_GATHER
        assert (eng.keyswitch(key, ext) == want).all()           # per-gate gather
        eng.set_ks_staged_min(1)
        assert eng.pick_kernels(key, 29)[1] == pkg.KS_STAGED
        assert (eng.keyswitch(key, ext) == want).all()           # staged: 3 CTAs, the last one with 7 idle groups
    finally:
        eng.set_pair_max(old_w); eng.set_ks_staged_min(old_s)


def test_keyswitch_staged_small_lwe_dimension(pkg, eng, small):
    """staged key switch when the LWE dimension is far below the 632-word row (n = 8): rows are zero-padded"""
    ks, _, key = small
    rng = np.random.default_rng(6)
    ext = rng.integers(-2 ** 31, 2 ** 31, size=(13, 1025), dtype=np.int64).astype(np.int32)
    old_w, old_s = eng.set_pair_max(0), eng.set_ks_staged_min(1)
    try:
        assert (eng.keyswitch(key, ext) == ks.keyswitch(ext)).all()
    finally:
        eng.set_pair_max(old_w); eng.set_ks_staged_min(old_s)


def test_aliasing_and_empty(eng, full, kernel_mode):
    ks, key = full
    a, b = ks.encrypt([1, 1, 0], 1), ks.encrypt([1, 0, 0], 2)
    out = eng.gate_batch(key, "AND", a, b)
    assert list(ks.decrypt(out)) == [1, 0, 0]
    assert eng.gate_batch(key, "AND", a[:0], b[:0]).shape == (0, 631)
    # ragged batch sizes around the CTA packing (2 gates per CTA) and one above a wave
    for count in (1, 3, 5, 593):
        bits = np.arange(count) % 2
        x = ks.encrypt(bits, 50 + count)
        y = ks.encrypt(1 - bits, 60 + count)
        assert (ks.decrypt(eng.gate_batch(key, "OR", x, y)) == 1).all()
        assert (ks.decrypt(eng.gate_batch(key, "XOR", x, x.copy())) == 0).all()


def test_large_batch_truth_table(pkg, eng):
    """2^15 independent NANDs on GPU-generated key and ciphertexts (BASELINE.json config 2, scaled):
    every decrypted result must equal the truth table, and the phases must sit at +-1/8."""
    p = pkg.Params.default(630)
    sk, key = eng.keygen(p, seed=314_1592_657)
    count = 1 << 15
    rng = np.random.default_rng(5)
    ba, bb = rng.integers(0, 2, count).astype(np.int32), rng.integers(0, 2, count).astype(np.int32)
    da, db, do = (eng.device_alloc(count * 632 * 4) for _ in range(3))
    sk.encrypt_to_device(ba, da, seed=1)
    sk.encrypt_to_device(bb, db, seed=2)
    assert (sk.decrypt_from_device(da, count) == ba).all()
    eng.gate_batch_device(key, "NAND", do, da, db, count=count)
    bits, ph = sk.decrypt_from_device(do, count, want_phases=True)
    assert (bits == 1 - (ba & bb)).all()
    dev = np.abs(np.abs(ph.astype(np.float64) / 2 ** 32) - 0.125)
    assert dev.max() < 0.04 and dev.std() < 0.01
    for ptr in (da, db, do):
        eng.device_free(ptr)
    key.close(); sk.close()


@pytest.mark.parametrize("count", [13, 149, 590, 1000, 1789])
def test_every_kernel_shape_is_deterministic_run_to_run(pkg, eng, full, count):
    """The same batch evaluated twice must give the same ciphertext words, for every kernel shape the default policy picks
    (13 -> 2-CTA cluster, 149 -> pair, 590 -> one gate per CTA, 1000 and the ragged 1789 -> persistent kernel, whose
    warps exchange data through shared and tensor memory with warp-level synchronisation only).  A missing barrier or an
    access outside a warp's own buffers shows up as words that differ between runs; this pool offers no race checker, so
    this and the bit-equality with the integer reference are the evidence."""
    ks, key = full
    rng = np.random.default_rng(count)
    a = ks.encrypt(rng.integers(0, 2, count).astype(np.int32), 11)
    b = ks.encrypt(rng.integers(0, 2, count).astype(np.int32), 12)
    first = eng.gate_batch(key, "NAND", a, b)
    for _ in range(3):
        assert np.array_equal(eng.gate_batch(key, "NAND", a, b), first)


def test_gpu_keygen_is_a_valid_key_for_the_oracle(pkg, oracle, eng):
    """Keys made on the GPU (Keygen/keygen.c's role) must work in the CPU oracle and vice versa."""
    p = pkg.Params.default(20)
    sk, key, bk, ksk = eng.keygen(p, seed=77, export=True)
    lwe, tlwe = sk.export()
    ks = oracle.from_arrays(ob.params_default(20), lwe, tlwe, bk, ksk)
    a, b = ks.encrypt([0, 0, 1, 1], 1), ks.encrypt([0, 1, 0, 1], 2)
    want = ks.gate_batch(ob.OPS["NAND"], a, b)
    assert list(ks.decrypt(want)) == [1, 1, 1, 0]
    got = eng.gate_batch(key, "NAND", a, b)
    assert list(ks.decrypt(got)) == [1, 1, 1, 0]
    assert torus_dist(ks.phase(got), ks.phase(want)).max() < PHASE_TOL
    key.close(); sk.close(); ks.free()


# ------------------------------------------------------------------ circuits
def _inputs(ks, words, seed):
    return np.concatenate([ks.encrypt_word(w, seed + i) for i, w in enumerate(words)])


def _chunks(v, nc):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(nc)]


def _decode(ks, out, nwords):
    return sum(ks.decrypt_word(out[32 * i:32 * (i + 1)]) << (32 * i) for i in range(nwords))


@pytest.mark.parametrize("kind,width", [(1, 32), (1, 64), (1, 256), (2, 32), (2, 128), (4, 32), (5, 32)])
def test_circuits_default_params_vs_integers(pkg, eng, full, kind, width, kernel_mode):
    ks, key = full
    nc = width // 32
    circ = eng.circuit(kind, width)
    rng = np.random.default_rng(kind * 1000 + width)
    n_expr = 3 if kind in (1, 2) else 2
    ins, expect = [], []
    for e in range(n_expr):
        a = int(rng.integers(0, 2 ** 62)) ** 5 % (1 << width)
        b = int(rng.integers(0, 2 ** 62)) ** 5 % (1 << width)
        if e == 0 and kind in (1, 4):
            a = b = 1 << (width - 2)          # Client1/process.c:94-99 stock operand 2^(bits-2)
        if kind == 5:
            c = int(rng.integers(0, 2 ** 31))
            ins.append(_inputs(ks, [a, b, c, 0, 0], 100 * e))
            expect.append((a * b + c) % (1 << 64))
        else:
            ins.append(_inputs(ks, _chunks(a, nc) + _chunks(b, nc) + [0], 100 * e))
            expect.append({1: (a + b) % (1 << width), 2: (a - b) % (1 << width), 4: a * b}[kind])
    out = eng.eval(key, circ, np.ascontiguousarray(np.stack(ins)))
    for e in range(n_expr):
        assert _decode(ks, out[e], circ.n_outputs // 32) == expect[e], (kind, width, e)


@pytest.mark.parametrize("kind,width", [(4, 64), (4, 128), (2, 256), (1, 128)])
def test_wide_circuits_default_params_vs_integers(pkg, eng, full, kind, width):
    """the deep circuits at n = 630 under the default launch policy: 64-bit multiply (35 296 bootstraps, 449 levels),
    128-bit multiply (121 184 bootstraps, 1 601 levels: the circuit that leans hardest on the latency kernels),
    256-bit subtract (2 560 bootstraps, 770 levels)"""
    ks, key = full
    nc = width // 32
    circ = eng.circuit(kind, width)
    rng = np.random.default_rng(kind * 7 + width)
    ins, expect = [], []
    for e in range(2):
        a = int.from_bytes(rng.bytes(width // 8), "little")
        b = int.from_bytes(rng.bytes(width // 8), "little")
        if e == 0 and kind == 4:
            a = b = (1 << width) - 1                       # every partial product and every carry chain at full length
        ins.append(_inputs(ks, _chunks(a, nc) + _chunks(b, nc) + [0], 100 * e))
        expect.append({1: (a + b) % (1 << width), 2: (a - b) % (1 << width), 4: a * b}[kind])
    out = eng.eval(key, circ, np.ascontiguousarray(np.stack(ins)))
    for e in range(2):
        assert _decode(ks, out[e], circ.n_outputs // 32) == expect[e], (kind, width, e)


@pytest.mark.parametrize("width", [64, 128])
def test_wide_multipliers_small_params(pkg, eng, small, width):
    ks, _, key = small
    nc = width // 32
    circ = eng.circuit(4, width)
    rng = np.random.default_rng(width)
    a = int.from_bytes(rng.bytes(width // 8), "little")
    b = int.from_bytes(rng.bytes(width // 8), "little")
    ins = _inputs(ks, _chunks(a, nc) + _chunks(b, nc) + [0], 7)[None]
    out = eng.eval(key, circ, np.ascontiguousarray(ins))
    assert _decode(ks, out[0], 2 * nc) == a * b


def test_mul32_matches_oracle_circuit(eng, small, kernel_mode):
    """the same mul32 on the oracle (11 264 gate calls, in cloud.c's order) and on the GPU (255 levels)"""
    ks, _, key = small
    a, b = 0x9E3779B9, 0x7F4A7C15
    A, B, C = ks.encrypt_word(a, 1), ks.encrypt_word(b, 2), ks.encrypt_word(0, 3)
    hi, lo = ks.mul32(A, B, C)
    circ = eng.circuit(4, 32)
    out = eng.eval(key, circ, np.ascontiguousarray(np.concatenate([A, B, C])[None]))[0]
    assert (ks.decrypt(out[:32]) == ks.decrypt(lo)).all() and (ks.decrypt(out[32:]) == ks.decrypt(hi)).all()
    assert _decode(ks, out, 2) == a * b
    assert torus_dist(ks.phase(out[:32]), ks.phase(lo)).max() < PHASE_TOL


# ------------------------------------------------------------------ the ./cloud process contract
CLOUD_CASES = [(1, 0, 0, 32, 1 << 30, 1 << 30), (1, 2, 2, 32, 5, 7), (2, 0, 2, 32, 100, 23), (2, 0, 0, 32, 1000, 1),
               (2, 0, 0, 32, 1, 1000), (1, 2, 0, 32, 50, 20), (2, 2, 2, 32, 3, 10), (1, 0, 0, 64, (1 << 62) + 12345, (1 << 62) + 1),
               (4, 2, 0, 32, 77777, 99999), (4, 0, 0, 64, (1 << 62), (1 << 62)),
               # the branches round 1 never ran through cloud_run: 128-bit multiply (4 x mul128 + 15 chained adds,
               # 121 184 bootstraps, 1 601 levels; Cloud/cloud.c:2371-2491), 128- and 256-bit add / subtract
               (4, 0, 2, 128, (1 << 127) - 1, (1 << 126) + 987654321987654321), (4, 0, 0, 128, (1 << 128) - 1, (1 << 128) - 1),
               (2, 0, 0, 128, (1 << 100) + 5, (1 << 99) + 77), (1, 0, 0, 256, (1 << 255) - 19, (1 << 200) + 3),
               (2, 2, 0, 256, (1 << 250) + 1, (1 << 251) + 9)]


@pytest.mark.parametrize("op,s1,s2,width,a,b", CLOUD_CASES)
def test_cloud_run_files(tmp_path, oracle, eng, small, op, s1, s2, width, a, b):
    """keygen -> alice -> [GPU cloud] -> verif through the reference's files: cloud.key, nbit.key,
    cloud.data, operator.txt in; answer.data out (Cloud/cloud.c:656-916), against the oracle's
    cloud main() on the same files and against Python integers."""
    ks, nbit, _ = small
    d = str(tmp_path)
    ks.write_cloud_key(os.path.join(d, "cloud.key"))
    nbit.write_secret_key(os.path.join(d, "nbit.key"))
    data = np.concatenate([oracle.alice(ks, nbit, s1, width, a, seed=1), oracle.alice(ks, nbit, s2, width, b, seed=2)])
    ks.write_samples(data, os.path.join(d, "cloud.data"))
    open(os.path.join(d, "operator.txt"), "w").write(str(op))
    rc, secs = eng.cloud_run(d)
    assert rc == 0
    rec = 4 + 4 * 9 + 8
    assert os.path.getsize(os.path.join(d, "answer.data")) == 352 * rec
    ans = ks.read_samples(os.path.join(d, "answer.data"), 352)
    code, w, chunks = oracle.verif(ks, nbit, ans)
    rc_o, ans_o = oracle.cloud_main(ks, nbit, op, data)
    assert (code, w, chunks) == oracle.verif(ks, nbit, ans_o)          # decrypt-identical to the reference flow
    va, vb = (-a if s1 == 2 else a), (-b if s2 == 2 else b)
    assert ob.decode_result(op, code, w, chunks) == {1: va + vb, 2: va - vb, 4: va * vb}[op]
    # padding blocks are copies of operand 1's carry block (cloud.c:901-916)
    nres = (2 * width if op == 4 else width) // 32
    assert (ans[(2 + nres) * 32:(3 + nres) * 32] == data[320:352]).all() and (ans[320:352] == data[320:352]).all()
    if op == 4:
        assert os.path.exists(os.path.join(d, "averagestandard.txt"))


def test_cloud_run_abort_path(tmp_path, oracle, eng, small):
    ks, nbit, _ = small
    d = str(tmp_path)
    ks.write_cloud_key(os.path.join(d, "cloud.key"))
    nbit.write_secret_key(os.path.join(d, "nbit.key"))
    data = np.concatenate([oracle.alice(ks, nbit, 0, 256, 3, seed=1), oracle.alice(ks, nbit, 0, 256, 5, seed=2)])
    ks.write_samples(data, os.path.join(d, "cloud.data"))
    open(os.path.join(d, "operator.txt"), "w").write("4")
    rc, _ = eng.cloud_run(d)
    assert rc == 126                                                    # Cloud/cloud.c:860-864
    assert os.path.getsize(os.path.join(d, "answer.data")) == 64 * (4 + 4 * 9 + 8)


def test_cloud_executable_and_libtfhe_api(tmp_path, oracle, small):
    """(1) the drop-in ./cloud binary; (2) a C++ program written against include/tfhe/tfhe.h exactly like
    Cloud/cloud.c's add() (pointer arithmetic on LweSample arrays, aliasing result/input) linked
    against libieache_b200.so instead of libtfhe."""
    ks, nbit, _ = small
    d = str(tmp_path)
    ks.write_cloud_key(os.path.join(d, "cloud.key"))
    ks.write_secret_key(os.path.join(d, "secret.key"))
    nbit.write_secret_key(os.path.join(d, "nbit.key"))
    a, b = 123456789, 987654321
    data = np.concatenate([oracle.alice(ks, nbit, 0, 32, a, seed=1), oracle.alice(ks, nbit, 0, 32, b, seed=2)])
    ks.write_samples(data, os.path.join(d, "cloud.data"))
    open(os.path.join(d, "operator.txt"), "w").write("1")
    r = subprocess.run([os.path.join(ROOT, "ie-ache_b200", "cloud")], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert "Computation Time:" in r.stdout
    code, w, chunks = oracle.verif(ks, nbit, ks.read_samples(os.path.join(d, "answer.data"), 352))
    assert (code, w, chunks[0]) == (0, 32, a + b)
    # (2)
    src = os.path.join(ROOT, "tests", "compat", "add_like_cloud.cpp")
    exe = os.path.join(d, "add_like_cloud")
    cc = subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                         "-L", os.path.join(ROOT, "ie-ache_b200"), "-lieache_b200", "-Wl,-rpath," + os.path.join(ROOT, "ie-ache_b200")],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert cc.returncode == 0, cc.stdout
    r = subprocess.run([exe], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    assert f"sum={a + b}" in r.stdout and "mux_ok=1" in r.stdout, r.stdout
