"""ctypes binding of oracle/liboracle.so — the CPU restatement of libtfhe used as the checker.
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_long, c_size_t, c_uint64, c_void_p

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")

OPS = {"NAND": 0, "OR": 1, "AND": 2, "XOR": 3, "XNOR": 4, "NOR": 5, "ANDNY": 6, "ANDYN": 7, "ORNY": 8, "ORYN": 9,
       "MUX": 10, "NOT": 11, "COPY": 12, "CONST": 13}


class OParams(ctypes.Structure):
    _fields_ = [(k, c_int32) for k in ("n", "N", "k", "bk_l", "bk_Bgbit", "ks_t", "ks_basebit")] + [("_pad", c_int32)] + [
        (k, c_double) for k in ("ks_stdev", "bk_stdev", "max_stdev")]


def _vp(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


_L = None


def _lib():
    global _L
    if _L is None:
        if not os.path.exists(LIB):
            raise RuntimeError(f"{LIB} missing; run `make -C oracle` (or __graft_entry__.build())")
        L = ctypes.CDLL(LIB)
        L.o_keygen.restype = c_void_p
        L.o_keygen.argtypes = [POINTER(OParams), c_uint64]
        L.o_keyset_from_arrays.restype = c_void_p
        L.o_keyset_from_arrays.argtypes = [POINTER(OParams), c_void_p, c_void_p, c_void_p, c_void_p]
        L.o_read_key.restype = c_void_p
        L.o_read_key.argtypes = [c_char_p]
        for f in ("o_lwe_key", "o_tlwe_key", "o_bk_coef", "o_ksk", "o_keyset_params"):
            getattr(L, f).restype = c_void_p
            getattr(L, f).argtypes = [c_void_p]
        L.o_keyset_free.argtypes = [c_void_p]
        L.o_sym_encrypt.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p, c_uint64]
        L.o_sym_decrypt.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        L.o_phase.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        L.o_phase_extracted.argtypes = [c_void_p, c_void_p, c_size_t, c_void_p]
        L.o_gate_batch.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int]
        L.o_bootstrap_woks.argtypes = [c_void_p, c_void_p, c_int32, c_void_p]
        L.o_bootstrap_woks_exact.argtypes = [c_void_p, c_void_p, c_int32, c_void_p]
        L.o_keyswitch.argtypes = [c_void_p, c_void_p, c_void_p]
        L.o_write_cloud_key.argtypes = [c_void_p, c_char_p]
        L.o_write_secret_key.argtypes = [c_void_p, c_char_p]
        L.o_write_samples.argtypes = [c_void_p, c_void_p, c_size_t, c_char_p, c_int]
        L.o_read_samples.restype = c_long
        L.o_read_samples.argtypes = [c_void_p, c_void_p, c_size_t, c_char_p, c_size_t]
        L.o_add.argtypes = [c_void_p] + [c_void_p] * 5 + [c_int]
        L.o_mul32.argtypes = [c_void_p] + [c_void_p] * 5
        L.o_alice.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_uint64]
        L.o_cloud_main.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, POINTER(c_size_t), c_uint64]
        L.o_verif_decrypt.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_int32), POINTER(c_int32), c_void_p]
        L.o_stats_bootstraps.restype = c_uint64
        _L = L
    return _L


def params_default(n: int = 630) -> OParams:
    p = OParams()
    _lib().o_params_default(byref(p))
    p.n = n
    return p


class KeySet:
    def __init__(self, handle):
        self._h = c_void_p(handle)
        self.p = OParams.from_address(_lib().o_keyset_params(self._h))
        self.n = self.p.n

    def free(self):
        if self._h:
            _lib().o_keyset_free(self._h)
            self._h = None

    # raw views (copies)
    def bk_coef(self) -> np.ndarray:
        p = self.p
        cnt = p.n * (p.k + 1) * p.bk_l * (p.k + 1) * p.N
        return np.ctypeslib.as_array((c_int32 * cnt).from_address(_lib().o_bk_coef(self._h))).copy()

    def ksk(self) -> np.ndarray:
        p = self.p
        cnt = p.k * p.N * p.ks_t * (1 << p.ks_basebit) * (p.n + 1)
        return np.ctypeslib.as_array((c_int32 * cnt).from_address(_lib().o_ksk(self._h))).copy()

    def lwe_key(self) -> np.ndarray:
        return np.ctypeslib.as_array((c_int32 * self.n).from_address(_lib().o_lwe_key(self._h))).copy()

    def encrypt(self, bits, seed: int) -> np.ndarray:
        b = np.ascontiguousarray(bits, dtype=np.int32)
        out = np.zeros((len(b), self.n + 1), dtype=np.int32)
        _lib().o_sym_encrypt(self._h, _vp(b), len(b), _vp(out), seed)
        return out

    def decrypt(self, samples: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(samples.reshape(-1, self.n + 1))
        out = np.zeros(len(s), dtype=np.int32)
        _lib().o_sym_decrypt(self._h, _vp(s), len(s), _vp(out))
        return out

    def phase(self, samples: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(samples.reshape(-1, self.n + 1))
        out = np.zeros(len(s), dtype=np.int32)
        _lib().o_phase(self._h, _vp(s), len(s), _vp(out))
        return out

    def phase_extracted(self, ext: np.ndarray) -> np.ndarray:
        s = np.ascontiguousarray(ext.reshape(-1, self.p.k * self.p.N + 1))
        out = np.zeros(len(s), dtype=np.int32)
        _lib().o_phase_extracted(self._h, _vp(s), len(s), _vp(out))
        return out

    def encrypt_word(self, value: int, seed: int) -> np.ndarray:
        return self.encrypt([(value >> i) & 1 for i in range(32)], seed)

    def decrypt_word(self, samples: np.ndarray) -> int:
        bits = self.decrypt(samples)
        return int(sum(int(b) << i for i, b in enumerate(bits)))

    def gate_batch(self, op: int, a, b=None, c=None, threads: int = 0) -> np.ndarray:
        out = np.zeros_like(a)
        _lib().o_gate_batch(self._h, op, _vp(out), _vp(a), _vp(b), _vp(c), len(a), threads)
        return out

    def bootstrap_woks(self, x: np.ndarray, mu: int = 1 << 29) -> np.ndarray:
        out = np.zeros((len(x), self.p.k * self.p.N + 1), dtype=np.int32)
        for i in range(len(x)):
            _lib().o_bootstrap_woks(self._h, _vp(out[i]), mu, _vp(np.ascontiguousarray(x[i])))
        return out

    def bootstrap_woks_exact(self, x: np.ndarray, mu: int = 1 << 29) -> np.ndarray:
        """oracle/exact_ref.c: the same blind rotation in exact integer arithmetic (schoolbook products, no transform)"""
        out = np.zeros((len(x), self.p.k * self.p.N + 1), dtype=np.int32)
        for i in range(len(x)):
            _lib().o_bootstrap_woks_exact(self._h, _vp(out[i]), mu, _vp(np.ascontiguousarray(x[i])))
        return out

    def keyswitch(self, ext: np.ndarray) -> np.ndarray:
        out = np.zeros((len(ext), self.n + 1), dtype=np.int32)
        for i in range(len(ext)):
            _lib().o_keyswitch(self._h, _vp(out[i]), _vp(np.ascontiguousarray(ext[i])))
        return out

    def write_cloud_key(self, path: str):
        assert _lib().o_write_cloud_key(self._h, path.encode()) == 0

    def write_secret_key(self, path: str):
        assert _lib().o_write_secret_key(self._h, path.encode()) == 0

    def write_samples(self, samples: np.ndarray, path: str, append: bool = False):
        s = np.ascontiguousarray(samples.reshape(-1, self.n + 1))
        assert _lib().o_write_samples(self._h, _vp(s), len(s), path.encode(), int(append)) == 0

    def read_samples(self, path: str, count: int, skip: int = 0) -> np.ndarray:
        out = np.zeros((count, self.n + 1), dtype=np.int32)
        got = _lib().o_read_samples(self._h, _vp(out), count, path.encode(), skip)
        return out[:max(got, 0)]

    # circuits (one gate at a time, like Cloud/cloud.c)
    def add(self, x, y, c, nb_bits: int = 32):
        s = np.zeros((32, self.n + 1), dtype=np.int32)
        co = np.zeros((32, self.n + 1), dtype=np.int32)
        _lib().o_add(self._h, _vp(s), _vp(co), _vp(np.ascontiguousarray(x)), _vp(np.ascontiguousarray(y)), _vp(np.ascontiguousarray(c)), nb_bits)
        return s, co

    def mul32(self, a, b, carry):
        hi = np.zeros((32, self.n + 1), dtype=np.int32)
        lo = np.zeros((32, self.n + 1), dtype=np.int32)
        _lib().o_mul32(self._h, _vp(hi), _vp(lo), _vp(np.ascontiguousarray(a)), _vp(np.ascontiguousarray(b)), _vp(np.ascontiguousarray(carry)))
        return hi, lo


class Oracle:
    def keygen(self, p: OParams, seed: int) -> KeySet:
        return KeySet(_lib().o_keygen(byref(p), seed))

    def from_arrays(self, p: OParams, lwe_key, tlwe_key, bk, ksk) -> KeySet:
        return KeySet(_lib().o_keyset_from_arrays(byref(p), _vp(lwe_key), _vp(tlwe_key), _vp(bk), _vp(ksk)))

    def read_key(self, path: str) -> KeySet:
        h = _lib().o_read_key(path.encode())
        if not h:
            raise RuntimeError(f"oracle could not read {path}")
        return KeySet(h)

    def alice(self, key: KeySet, nbit: KeySet, sign_code: int, width: int, value: int, seed: int) -> np.ndarray:
        chunks = np.array([(value >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
        out = np.zeros((352, key.n + 1), dtype=np.int32)
        _lib().o_alice(key._h, nbit._h, sign_code, width, _vp(chunks), _vp(out), seed)
        return out

    def cloud_main(self, ck: KeySet, nbit: KeySet, op: int, data704: np.ndarray, seed: int = 1):
        ans = np.zeros((352, ck.n + 1), dtype=np.int32)
        cnt = c_size_t()
        rc = _lib().o_cloud_main(ck._h, nbit._h, op, _vp(np.ascontiguousarray(data704)), _vp(ans), byref(cnt), seed)
        return rc, ans[:cnt.value]

    def verif(self, key: KeySet, nbit: KeySet, answer: np.ndarray):
        ans = np.zeros((352, key.n + 1), dtype=np.int32)
        ans[:len(answer)] = answer
        sc, w = c_int32(), c_int32()
        chunks = np.zeros(8, dtype=np.uint32)
        _lib().o_verif_decrypt(key._h, nbit._h, _vp(ans), byref(sc), byref(w), _vp(chunks))
        return sc.value, w.value, [int(c) for c in chunks]

    def stats_reset(self):
        _lib().o_stats_reset()

    def stats_bootstraps(self) -> int:
        return _lib().o_stats_bootstraps()

    def max_threads(self) -> int:
        return _lib().o_max_threads()


def decode_result(op: int, sign_code: int, width: int, chunks) -> int:
    """Output/verif.c's sign-magnitude decoding of the decrypted answer:
    add (op 1) :120-173, subtract (op 2) :733-789, multiply (op 4) :1409-1435."""
    nchunks = width // 32
    bits = 0
    for i in range(nchunks):
        bits |= chunks[i] << (32 * i)
    if op == 1:
        if sign_code in (0, 4):
            total = bits
        elif width == 32 and (bits >> 31) & 1:  # verif.c only re-interprets two's complement when length == 32
            total = bits - (1 << 32)
        else:
            total = bits
        return -total if sign_code == 4 else total
    if op == 2:
        if sign_code == 2:
            total = bits
        elif width == 32 and (bits >> 31) & 1:
            total = bits - (1 << 32)
        else:
            total = bits
        return -total if sign_code == 1 else total
    if op == 4:
        return -bits if sign_code in (1, 2) else bits
    raise ValueError(op)
