"""The blind-rotation kernel's per-thread code (ie-ache_b200/csrc/br_core.h) executed on the CPU by
tests/emul/host_emul.cpp and compared with the oracle: transform layout, exactness of the negacyclic
product, and a whole blind rotation + sample extraction."""
import ctypes
import os

import numpy as np
import pytest

import oracle_bind as ob

EMUL = os.path.join(os.path.dirname(__file__), "emul", "libbr_emul.so")


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(EMUL):
        pytest.fail("tests/emul/libbr_emul.so missing: run __graft_entry__.build()")
    E = ctypes.CDLL(EMUL)
    E.emul_slot_to_K.restype = ctypes.c_int
    return E


def vp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_slot_order_is_evaluation_at_odd_roots(emul):
    rng = np.random.default_rng(3)
    a = rng.integers(-64, 64, 1024).astype(np.int32)
    fa = np.zeros(1024)
    emul.emul_poly_fft(vp(a), vp(fa), ctypes.c_double(1.0))
    ca = fa[0::2] + 1j * fa[1::2]
    psi = np.exp(1j * np.pi / 1024)
    seen = set()
    for r in range(8):
        for t3 in (0, 9, 31, 63):
            K = emul.emul_slot_to_K(r, t3)
            seen.add(K)
            x = psi ** (4 * K + 1)
            val = np.polyval(a[::-1].astype(np.complex128), x)
            assert abs(val - ca[r * 64 + t3]) < 1e-6
    assert len(seen) == 32
    assert sorted(emul.emul_slot_to_K(r, t) for r in range(8) for t in range(64)) == list(range(512))


def test_negacyclic_product_is_exact(emul):
    rng = np.random.default_rng(4)
    a = rng.integers(-64, 64, 1024).astype(np.int32)           # gadget digits
    b = rng.integers(-2 ** 31, 2 ** 31, 1024).astype(np.int32)  # torus coefficients
    fa, fb = np.zeros(1024), np.zeros(1024)
    emul.emul_poly_fft(vp(a), vp(fa), ctypes.c_double(1.0))
    emul.emul_poly_fft(vp(b), vp(fb), ctypes.c_double(1.0 / 512))
    prod = (fa[0::2] + 1j * fa[1::2]) * (fb[0::2] + 1j * fb[1::2])
    pin = np.zeros(1024)
    pin[0::2], pin[1::2] = prod.real, prod.imag
    out = np.zeros(1024)
    emul.emul_poly_ifft(vp(pin), vp(out))
    full = np.convolve(a.astype(object), b.astype(object))
    ref = [int(full[i]) - (int(full[i + 1024]) if i + 1024 < len(full) else 0) for i in range(1024)]
    assert max(abs(out[i] - ref[i]) for i in range(1024)) < 0.25


@pytest.mark.parametrize("n,l,bgbit", [(24, 3, 7), (12, 2, 10)])
def test_blind_rotate_matches_oracle(emul, oracle, n, l, bgbit):
    p = ob.params_default(n)
    p.bk_l, p.bk_Bgbit = l, bgbit
    ks = oracle.keygen(p, seed=777)
    bits = np.array([1, 0, 1, 1, 0, 0], dtype=np.int32)
    s = ks.encrypt(bits, 5)
    mu = 1 << 29
    bk = ks.bk_coef()
    for g in range(6):
        x = (s[g] + s[(g + 1) % 6]).astype(np.int32)
        x[n] -= mu
        ext_o = ks.bootstrap_woks(x[None])[0]
        ext_e = np.zeros(1025, dtype=np.int32)
        emul.emul_blind_rotate(n, l, bgbit, ctypes.c_int32(mu), vp(bk), vp(x), vp(ext_e))
        want = int(bits[g] & bits[(g + 1) % 6])
        ph = ks.phase_extracted(ext_e[None])[0]
        assert (ph > 0) == bool(want)
        d = ext_o.astype(np.int64) - ext_e.astype(np.int64)
        d = ((d + 2 ** 31) % 2 ** 32) - 2 ** 31
        assert np.abs(d).max() <= 2   # FP64 rounding may differ in the last unit; normally 0
    ks.free()


# ---- warp-per-gate layout (ie-ache_b200/csrc/br_warp.h): 16 points per lane, one exchange + one shuffle stage

def test_warp_slot_order_is_evaluation_at_odd_roots(emul):
    emul.emul_warp_slot_to_K.restype = ctypes.c_int
    rng = np.random.default_rng(13)
    a = rng.integers(-64, 64, 1024).astype(np.int32)
    fa = np.zeros(1024)
    emul.emul_warp_fft(vp(a), vp(fa), ctypes.c_double(1.0))
    ca = fa[0::2] + 1j * fa[1::2]
    psi = np.exp(1j * np.pi / 1024)
    for p in range(16):
        for lane in (0, 5, 15, 16, 22, 31):
            K = emul.emul_warp_slot_to_K(p, lane)
            val = np.polyval(a[::-1].astype(np.complex128), psi ** (4 * K + 1))
            assert abs(val - ca[p * 32 + lane]) < 1e-6, (p, lane, K)
    assert sorted(emul.emul_warp_slot_to_K(p, l) for p in range(16) for l in range(32)) == list(range(512))


def test_warp_negacyclic_product_is_exact(emul):
    rng = np.random.default_rng(14)
    a = rng.integers(-64, 64, 1024).astype(np.int32)
    b = rng.integers(-2 ** 31, 2 ** 31, 1024).astype(np.int32)
    fa, fb = np.zeros(1024), np.zeros(1024)
    emul.emul_warp_fft(vp(a), vp(fa), ctypes.c_double(1.0))
    emul.emul_warp_fft(vp(b), vp(fb), ctypes.c_double(1.0 / 512))
    prod = (fa[0::2] + 1j * fa[1::2]) * (fb[0::2] + 1j * fb[1::2])
    pin = np.zeros(1024)
    pin[0::2], pin[1::2] = prod.real, prod.imag
    out = np.zeros(1024)
    emul.emul_warp_ifft(vp(pin), vp(out))
    full = np.convolve(a.astype(object), b.astype(object))
    ref = [int(full[i]) - (int(full[i + 1024]) if i + 1024 < len(full) else 0) for i in range(1024)]
    assert max(abs(out[i] - ref[i]) for i in range(1024)) < 0.25


@pytest.mark.parametrize("n,l,bgbit", [(24, 3, 7), (12, 2, 10)])
def test_warp_blind_rotate_matches_oracle(emul, oracle, n, l, bgbit):
    p = ob.params_default(n)
    p.bk_l, p.bk_Bgbit = l, bgbit
    ks = oracle.keygen(p, seed=778)
    bits = np.array([1, 0, 1, 1, 0, 0], dtype=np.int32)
    s = ks.encrypt(bits, 6)
    mu = 1 << 29
    bk = ks.bk_coef()
    for g in range(6):
        x = (s[g] + s[(g + 1) % 6]).astype(np.int32)
        x[n] -= mu
        ext_o = ks.bootstrap_woks(x[None])[0]
        ext_e = np.zeros(1025, dtype=np.int32)
        emul.emul_warp_blind_rotate(n, l, bgbit, ctypes.c_int32(mu), vp(bk), vp(x), vp(ext_e))
        ph = ks.phase_extracted(ext_e[None])[0]
        assert (ph > 0) == bool(bits[g] & bits[(g + 1) % 6])
        d = ext_o.astype(np.int64) - ext_e.astype(np.int64)
        d = ((d + 2 ** 31) % 2 ** 32) - 2 ** 31
        assert np.abs(d).max() <= 2
    ks.free()


@pytest.mark.parametrize("n,l,bgbit", [(24, 3, 7), (12, 2, 10)])
def test_w12_blind_rotate_matches_oracle(emul, oracle, n, l, bgbit):
    """the persistent 12-warp kernel's arithmetic (br_w12.cu): pass 1 from the integer digits, 4-entry final-stage table"""
    p = ob.params_default(n)
    p.bk_l, p.bk_Bgbit = l, bgbit
    ks = oracle.keygen(p, seed=779)
    bits = np.array([1, 0, 1, 1, 0, 0], dtype=np.int32)
    s = ks.encrypt(bits, 7)
    mu = 1 << 29
    bk = ks.bk_coef()
    for g in range(6):
        x = (s[g] + s[(g + 1) % 6]).astype(np.int32)
        x[n] = np.int32((int(x[n]) - mu + 2 ** 31) % 2 ** 32 - 2 ** 31)
        ext_o = ks.bootstrap_woks(x[None])[0]
        ext_e = np.zeros(1025, dtype=np.int32)
        emul.emul_w12_blind_rotate(n, l, bgbit, ctypes.c_int32(mu), vp(bk), vp(x), vp(ext_e))
        ph = ks.phase_extracted(ext_e[None])[0]
        assert (ph > 0) == bool(bits[g] & bits[(g + 1) % 6])
        d = ext_o.astype(np.int64) - ext_e.astype(np.int64)
        d = ((d + 2 ** 31) % 2 ** 32) - 2 ** 31
        assert np.abs(d).max() <= 2
    ks.free()


@pytest.mark.parametrize("n,l,bgbit", [(24, 3, 7), (12, 2, 10)])
def test_w12_select_free_blind_rotate_matches_oracle(emul, oracle, n, l, bgbit):
    """br_w12.cu as shipped: folded forward and folded inverse (no lane-dependent selects), uniform rule for the odd
    final-stage twiddles, plain key values permuted by w12_slot_to_K"""
    emul.emul_w12_slot_to_K.restype = ctypes.c_int
    assert sorted(emul.emul_w12_slot_to_K(p, ln) for p in range(16) for ln in range(32)) == list(range(512))
    p = ob.params_default(n)
    p.bk_l, p.bk_Bgbit = l, bgbit
    ks = oracle.keygen(p, seed=780)
    bits = np.array([1, 0, 1, 1, 0, 0], dtype=np.int32)
    s = ks.encrypt(bits, 8)
    mu = 1 << 29
    bk = ks.bk_coef()
    for g in range(6):
        x = (s[g] + s[(g + 1) % 6]).astype(np.int32)
        x[n] = np.int32((int(x[n]) - mu + 2 ** 31) % 2 ** 32 - 2 ** 31)
        ext_o = ks.bootstrap_woks(x[None])[0]
        ext_e = np.zeros(1025, dtype=np.int32)
        emul.emul_w12f_blind_rotate(n, l, bgbit, ctypes.c_int32(mu), vp(bk), vp(x), vp(ext_e))
        ph = ks.phase_extracted(ext_e[None])[0]
        assert (ph > 0) == bool(bits[g] & bits[(g + 1) % 6])
        d = ext_o.astype(np.int64) - ext_e.astype(np.int64)
        d = ((d + 2 ** 31) % 2 ** 32) - 2 ** 31
        assert np.abs(d).max() <= 2
    ks.free()


def test_warp_folded_forward_equals_plain_after_key_factor(emul):
    """the select-free forward variant leaves the Lpar = 1 lanes' results scaled by unit factors that the key layout
    carries: multiplied back, it must reproduce the plain warp-layout spectrum"""
    rng = np.random.default_rng(21)
    a = rng.integers(-64, 64, 1024).astype(np.int32)
    plain, folded = np.zeros(1024), np.zeros(1024)
    emul.emul_warp_fft(vp(a), vp(plain), ctypes.c_double(1.0))
    emul.emul_warpf_fft_unfolded(vp(a), vp(folded))
    assert np.abs(plain - folded).max() < 1e-9


def test_last_inverse_pass_with_pending_factors_equals_mirrored_butterflies(emul):
    """pass16_inv_p1 (6-FMA butterflies, exp(-i pi pos/32) applied once at the end) is the same linear map as
    pass16_inv with the pass-1 twiddles: equal to a few ulp of the largest input on 52-bit-sized data."""
    import ctypes
    rng = np.random.default_rng(5)
    f = emul.emul_pass16_inv_both
    f.argtypes = [ctypes.c_void_p] * 3
    for scale in (1.0, 2.0 ** 40):
        x = (rng.standard_normal(32) * scale).astype(np.float64)
        a = np.zeros(32); b = np.zeros(32)
        f(x.ctypes.data, a.ctypes.data, b.ctypes.data)
        assert np.max(np.abs(a)) > 0
        assert np.max(np.abs(a - b)) <= 64 * np.finfo(np.float64).eps * np.max(np.abs(a))
