"""ieache_b200 — ctypes binding of libieache_b200.so, the B200-native TFHE gate-bootstrapping engine
that replaces libtfhe under IE-ACHE's Cloud node (Cloud/cloud.c).

This is the host-side mirror of the reference's interface for the hot path:

* ``Engine.gate_batch``    — `count` independent ``boots<OP>`` calls (Cloud/cloud.c:30-43) in one launch
* ``Engine.circuit``/``eval`` — the add / subtract / multiply circuits of Cloud/cloud.c:18-647, levelised
* ``Engine.cloud_run``     — the ``subprocess.call("./cloud")`` contract of
  Cloud/dragonfly_cipher_cloud.py:1233 (reads cloud.key, nbit.key, cloud.data, operator.txt; writes answer.data)

There is no CPU fallback: importing works anywhere (so the symbol table can be checked on a CPU
box), but every compute call needs the CUDA library and a B200 and raises ``EngineError`` otherwise.
The directory name carries a hyphen (``ie-ache_b200``); load it with ``__graft_entry__.load_package()``
or put the repo root on ``sys.path`` and use ``importlib`` — it registers itself as ``ieache_b200``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int, c_int32, c_size_t, c_uint32, c_uint64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libieache_b200.so")

DEVICE_STRIDE = 632

OPS = {
    "NAND": 0, "OR": 1, "AND": 2, "XOR": 3, "XNOR": 4, "NOR": 5, "ANDNY": 6, "ANDYN": 7,
    "ORNY": 8, "ORYN": 9, "MUX": 10, "NOT": 11, "COPY": 12, "CONST": 13,
}
CIRC_ADD, CIRC_SUB, CIRC_MUL, CIRC_MULADD = 1, 2, 4, 5


class EngineError(RuntimeError):
    pass


class Params(ctypes.Structure):
    _fields_ = [(k, c_int32) for k in ("n", "N", "k", "bk_l", "bk_Bgbit", "ks_t", "ks_basebit", "reserved")] + [
        (k, c_double) for k in ("ks_stdev", "bk_stdev", "max_stdev")
    ]

    @classmethod
    def default(cls, n: int = 630) -> "Params":
        """new_default_gate_bootstrapping_parameters(110) (Keygen/keygen.c:22-23) with n overridable for tests."""
        return cls(n, 1024, 1, 3, 7, 8, 2, 0, 2.0 ** -15, 2.0 ** -25, 0.012467)


_lib = None


def lib() -> ctypes.CDLL:
    """Load the CUDA extension; fail loudly if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    L = ctypes.CDLL(LIB_PATH)
    L.ieache_last_error.restype = c_char_p
    L.ieache_version.restype = c_char_p
    L.ieache_ctx_launch_count.restype = c_uint64
    L.ieache_ctx_launch_count.argtypes = [c_void_p]
    L.ieache_ctx_create.argtypes = [c_int, POINTER(c_void_p)]
    L.ieache_ctx_destroy.argtypes = [c_void_p]
    L.ieache_ctx_sync.argtypes = [c_void_p]
    L.ieache_ctx_set_timing.argtypes = [c_void_p, c_int]
    L.ieache_ctx_kernel_times.argtypes = [c_void_p, POINTER(c_double), POINTER(c_double), POINTER(c_uint64), POINTER(c_uint64), c_int]
    L.ieache_cloudkey_create.argtypes = [c_void_p, POINTER(Params), c_void_p, c_void_p, POINTER(c_void_p)]
    L.ieache_cloudkey_load_file.argtypes = [c_void_p, c_char_p, POINTER(c_void_p)]
    L.ieache_cloudkey_destroy.argtypes = [c_void_p]
    L.ieache_cloudkey_params.argtypes = [c_void_p, POINTER(Params)]
    L.ieache_cloudkey_device_arrays.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_size_t), POINTER(c_void_p), POINTER(c_size_t)]
    L.ieache_cloudkey_device_sizes.argtypes = [POINTER(Params), POINTER(c_size_t), POINTER(c_size_t)]
    L.ieache_cloudkey_adopt_device.argtypes = [c_void_p, POINTER(Params), c_void_p, c_void_p, POINTER(c_void_p)]
    L.ieache_gate_batch.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_size_t]
    L.ieache_gate_batch_device.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_size_t]
    L.ieache_bootstrap_woks.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]
    L.ieache_keyswitch.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]
    L.ieache_device_alloc.argtypes = [c_void_p, c_size_t, POINTER(c_void_p)]
    L.ieache_device_free.argtypes = [c_void_p, c_void_p]
    L.ieache_device_copy.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t]
    L.ieache_samples_to_device.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_int32]
    L.ieache_samples_to_host.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_int32]
    L.ieache_circuit_build.argtypes = [c_int, c_int, POINTER(c_void_p)]
    L.ieache_circuit_destroy.argtypes = [c_void_p]
    L.ieache_circuit_stats.argtypes = [c_void_p, POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint64), POINTER(c_uint32),
                                       POINTER(c_uint32), POINTER(c_uint32), POINTER(c_uint32)]
    L.ieache_circuit_eval.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]
    L.ieache_circuit_eval_device.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t]
    L.ieache_cloud_run.argtypes = [c_void_p, c_char_p, POINTER(c_double)]
    L.ieache_ctx_set_tuning.argtypes = [c_void_p, c_int, ctypes.c_int64, POINTER(ctypes.c_int64)]
    L.ieache_ctx_pick_kernels.argtypes = [c_void_p, c_void_p, ctypes.c_int64, POINTER(c_int), POINTER(c_int)]
    L.ieache_ctx_copy_counts.argtypes = [c_void_p, POINTER(c_uint64), POINTER(c_uint64)]
    L.ieache_ctx_timer_start.argtypes = [c_void_p]
    L.ieache_ctx_timer_stop.argtypes = [c_void_p, POINTER(c_double)]
    L.ieache_measure_fp64_peak.argtypes = [c_void_p, POINTER(c_double)]
    L.ieache_host_alloc.argtypes = [c_size_t, POINTER(c_void_p)]
    L.ieache_host_free.argtypes = [c_void_p]
    L.ieache_session_open.argtypes = [c_void_p, c_char_p, c_char_p, POINTER(c_void_p)]
    L.ieache_session_open_keys.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_void_p)]
    L.ieache_session_close.argtypes = [c_void_p]
    L.ieache_session_params.argtypes = [c_void_p, POINTER(Params)]
    L.ieache_session_compute.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, POINTER(c_size_t), POINTER(c_double)]
    L.ieache_session_compute_batch.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(c_double)]
    L.ieache_session_eval_postfix.argtypes = [c_void_p, c_char_p, c_size_t, c_void_p, c_int, c_void_p, c_void_p, POINTER(c_double)]
    L.ieache_session_compute_dirs.argtypes = [c_void_p, c_size_t, c_void_p, c_void_p, POINTER(c_double)]
    L.ieache_session_set_pass.argtypes = [c_void_p, c_size_t, c_void_p]
    L.ieache_keygen_files.argtypes = [c_void_p, c_char_p, POINTER(Params), c_uint64, c_uint64]
    L.ieache_alice_encrypt.argtypes = [c_char_p, c_int32, c_int32, c_void_p, c_char_p, c_int]
    L.ieache_alice_run.argtypes = [c_char_p]
    L.ieache_verif_run.argtypes = [c_char_p, c_char_p, c_size_t, POINTER(c_int32), POINTER(c_int32)]
    L.ieache_keygen.argtypes = [c_void_p, POINTER(Params), c_uint64, POINTER(c_void_p), POINTER(c_void_p), c_void_p, c_void_p]
    L.ieache_secretkey_import.argtypes = [c_void_p, POINTER(Params), c_void_p, c_void_p, POINTER(c_void_p)]
    L.ieache_secretkey_export.argtypes = [c_void_p, c_void_p, c_void_p]
    L.ieache_secretkey_destroy.argtypes = [c_void_p]
    L.ieache_sym_encrypt_device.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_uint64]
    L.ieache_sym_decrypt_device.argtypes = [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]
    _lib = L
    return L


def alice_encrypt(directory: str, sign_code: int, width: int, value: int, out_path: str, append: bool = False) -> None:
    """Client1/alice.c: one operand of `width` bits -> 352 ciphertext records (needs secret.key, nbit.key in directory)"""
    chunks = np.array([(value >> (32 * i)) & 0xFFFFFFFF for i in range(8)], dtype=np.uint32)
    _check(lib().ieache_alice_encrypt(directory.encode(), sign_code, width, _ptr(chunks), out_path.encode(), int(append)))


def alice_run(directory: str) -> None:
    """./alice: values.txt -> cloud.data"""
    _check(lib().ieache_alice_run(directory.encode()))


def verif_run(directory: str):
    """./verif: answer.data + operator.txt -> (decimal value, sign code, width)"""
    buf = ctypes.create_string_buffer(128)
    sc, w = c_int32(), c_int32()
    _check(lib().ieache_verif_run(directory.encode(), buf, 128, byref(sc), byref(w)))
    return int(buf.value.decode()), sc.value, w.value


TUNE_CLUSTER_MAX, TUNE_PAIR_MAX, TUNE_W12_MIN, TUNE_KS_STAGED_MIN, TUNE_THROUGHPUT_KERNEL = 1, 2, 3, 4, 5
KERNEL_CLUSTER, KERNEL_PAIR, KERNEL_GROUP, KERNEL_W12 = 1, 2, 41, 70
KS_CLUSTER, KS_GATHER, KS_STAGED = 1, 2, 3


def _check(rc: int) -> None:
    if rc < 0:
        raise EngineError(f"ieache_b200 error {rc}: {lib().ieache_last_error().decode()}")


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.dtype in (np.int32, np.uint32) and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(c_void_p)
    return c_void_p(int(a))  # raw device/host address


class Circuit:
    """Levelised DAG of one Cloud/cloud.c dispatch branch (kind x width)."""

    def __init__(self, kind: int, width: int):
        self._h = c_void_p()
        _check(lib().ieache_circuit_build(kind, width, byref(self._h)))
        b, a, x = c_uint64(), c_uint64(), c_uint64()
        lv, mw, ni, no = c_uint32(), c_uint32(), c_uint32(), c_uint32()
        _check(lib().ieache_circuit_stats(self._h, byref(b), byref(a), byref(x), byref(lv), byref(mw), byref(ni), byref(no)))
        self.kind, self.width = kind, width
        self.bootstraps, self.and_gates, self.xor_gates = b.value, a.value, x.value
        self.levels, self.max_width, self.n_inputs, self.n_outputs = lv.value, mw.value, ni.value, no.value

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ieache_circuit_destroy(self._h)
            self._h = None


class CloudKey:
    def __init__(self, engine: "Engine", handle: c_void_p):
        self.engine, self._h = engine, handle
        p = Params()
        _check(lib().ieache_cloudkey_params(handle, byref(p)))
        self.params = p

    def device_arrays(self):
        bk, bkb, ks, ksb = c_void_p(), c_size_t(), c_void_p(), c_size_t()
        _check(lib().ieache_cloudkey_device_arrays(self._h, byref(bk), byref(bkb), byref(ks), byref(ksb)))
        return bk.value, bkb.value, ks.value, ksb.value

    def close(self):
        if self._h:
            lib().ieache_cloudkey_destroy(self._h)
            self._h = None


def pinned_array(shape, dtype=np.int32) -> np.ndarray:
    """numpy view of page-locked host memory (cudaHostAlloc); freed when the array is collected."""
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = c_void_p()
    _check(lib().ieache_host_alloc(max(nbytes, 1), byref(p)))
    buf = (ctypes.c_byte * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    import weakref
    weakref.finalize(buf, lib().ieache_host_free, p)
    return arr


class SecretKey:
    """LWE + TLWE secret keys (host and device copies): the Keygen/Client/Output side of the protocol."""

    def __init__(self, engine: "Engine", handle: c_void_p, params: Params):
        self.engine, self._h, self.params = engine, handle, params

    def export(self):
        lwe = np.zeros(self.params.n, dtype=np.int32)
        tlwe = np.zeros(1024, dtype=np.int32)
        _check(lib().ieache_secretkey_export(self._h, _ptr(lwe), _ptr(tlwe)))
        return lwe, tlwe

    def encrypt_to_device(self, bits: np.ndarray, out_dev: int, seed: int):
        """bootsSymEncrypt (Client1/alice.c:117) of a host bit array into device samples."""
        b = np.ascontiguousarray(bits, dtype=np.int32)
        _check(lib().ieache_sym_encrypt_device(self.engine._h, self._h, _ptr(b), len(b), c_void_p(out_dev), seed))

    def decrypt_from_device(self, samples_dev: int, count: int, want_phases: bool = False):
        """bootsSymDecrypt (Output/verif.c:93) of device samples -> bits (and phases)."""
        bits = np.zeros(count, dtype=np.int32)
        ph = np.zeros(count, dtype=np.int32) if want_phases else None
        _check(lib().ieache_sym_decrypt_device(self.engine._h, self._h, c_void_p(samples_dev), count, _ptr(bits), _ptr(ph)))
        return (bits, ph) if want_phases else bits

    def close(self):
        if self._h:
            lib().ieache_secretkey_destroy(self._h)
            self._h = None


class Session:
    """Cloud-node session: cloud.key and nbit.key loaded once, then any number of operators
    (Cloud/cloud.c main() per operator, Cloud/dragonfly_cipher_cloud.py:685-729 for whole expressions)."""

    def __init__(self, engine: "Engine", handle: c_void_p):
        self.engine, self._h = engine, handle
        self.params = Params()
        _check(lib().ieache_session_params(handle, byref(self.params)))

    def compute(self, op: int, operand1: np.ndarray, operand2: np.ndarray):
        """one operator on two 352-sample client blocks -> (exit_code, answer block, seconds)"""
        w = self.params.n + 1
        ans = np.zeros((352, w), dtype=np.int32)
        cnt, secs = c_size_t(), c_double()
        rc = lib().ieache_session_compute(self._h, op, _ptr(np.ascontiguousarray(operand1)), _ptr(np.ascontiguousarray(operand2)),
                                          _ptr(ans), byref(cnt), byref(secs))
        _check(rc)
        return rc, ans[:cnt.value], secs.value

    def compute_batch(self, ops, operands1: np.ndarray, operands2: np.ndarray):
        """`len(ops)` independent requests, batched per circuit -> (exit_codes, answers, counts, seconds)"""
        n, w = len(ops), self.params.n + 1
        ops = np.ascontiguousarray(ops, dtype=np.int32)
        ans = np.zeros((n, 352, w), dtype=np.int32)
        codes = np.zeros(n, dtype=np.int32)
        counts = np.zeros(n, dtype=np.uint64)
        secs = c_double()
        _check(lib().ieache_session_compute_batch(self._h, n, _ptr(ops), _ptr(np.ascontiguousarray(operands1)),
                                                  _ptr(np.ascontiguousarray(operands2)), _ptr(ans), _ptr(codes),
                                                  counts.ctypes.data_as(c_void_p), byref(secs)))
        return codes, ans, counts, secs.value

    def eval_postfix(self, postfix: str, operands: np.ndarray):
        """operands: (n_expr, n_operands, 352, n+1); e.g. postfix "AB*C+" -> (exit_code, answers, counts, seconds)"""
        n_expr, n_ops = operands.shape[0], operands.shape[1]
        w = self.params.n + 1
        ans = np.zeros((n_expr, 352, w), dtype=np.int32)
        counts = np.zeros(n_expr, dtype=np.uint64)
        secs = c_double()
        rc = lib().ieache_session_eval_postfix(self._h, postfix.encode(), n_expr, _ptr(np.ascontiguousarray(operands)), n_ops,
                                               _ptr(ans), counts.ctypes.data_as(c_void_p), byref(secs))
        _check(rc)
        return rc, ans, counts, secs.value

    def compute_dirs(self, directories):
        """batched ingest: every directory holds cloud.data + operator.txt and receives answer.data, as if ./cloud
        had run there; one levelised batch for all of them -> (exit_codes, seconds)"""
        n = len(directories)
        arr = (ctypes.c_char_p * n)(*[d.encode() for d in directories])
        codes = np.zeros(n, dtype=np.int32)
        secs = c_double()
        _check(lib().ieache_session_compute_dirs(self._h, n, arr, _ptr(codes), byref(secs)))
        return codes, secs.value

    def set_pass(self, requests: int) -> int:
        """requests evaluated together by the batch calls (default 256); returns the previous value"""
        old = ctypes.c_size_t()
        _check(lib().ieache_session_set_pass(self._h, requests, byref(old)))
        return old.value

    def close(self):
        if self._h:
            lib().ieache_session_close(self._h)
            self._h = None


class Engine:
    """One engine context per GPU (a process may hold several: the launch policy lives in the context)."""

    def __init__(self, device: int = 0):
        self._h = c_void_p()
        _check(lib().ieache_ctx_create(device, byref(self._h)))
        self.device = device

    # ---- launch policy (which kernel shape a launch of a given size uses) --------------------
    def set_tuning(self, which: int, value: int) -> int:
        """ieache_ctx_set_tuning: returns the previous value; raises EngineError for a value the library refuses."""
        old = ctypes.c_int64()
        _check(lib().ieache_ctx_set_tuning(self._h, which, value, byref(old)))
        return old.value

    def set_cluster_max(self, max_gates: int) -> int:
        return self.set_tuning(TUNE_CLUSTER_MAX, max_gates)

    def set_pair_max(self, max_gates: int) -> int:
        return self.set_tuning(TUNE_PAIR_MAX, max_gates)

    def set_w12_min(self, min_gates: int) -> int:
        return self.set_tuning(TUNE_W12_MIN, min_gates)

    def set_ks_staged_min(self, min_gates: int) -> int:
        return self.set_tuning(TUNE_KS_STAGED_MIN, min_gates)

    def set_throughput_kernel(self, kernel: int) -> int:
        """0 = by size, KERNEL_GROUP or KERNEL_W12 for every launch above the two-group threshold"""
        return self.set_tuning(TUNE_THROUGHPUT_KERNEL, kernel)

    def pick_kernels(self, key: "CloudKey", count: int):
        """(blind-rotation kernel, key-switch kernel) a launch of `count` gates would use"""
        br, ks = c_int(), c_int()
        _check(lib().ieache_ctx_pick_kernels(self._h, key._h, count, byref(br), byref(ks)))
        return br.value, ks.value

    def copy_counts(self):
        """(host->device, device->host) 352-sample blocks moved by the session calls so far"""
        a, b = c_uint64(), c_uint64()
        _check(lib().ieache_ctx_copy_counts(self._h, byref(a), byref(b)))
        return a.value, b.value

    def close(self):
        if self._h:
            lib().ieache_ctx_destroy(self._h)
            self._h = None

    # ---- keys -------------------------------------------------------------------------------
    def cloud_key_from_arrays(self, params: Params, bk: np.ndarray, ksk: np.ndarray) -> CloudKey:
        h = c_void_p()
        _check(lib().ieache_cloudkey_create(self._h, byref(params), _ptr(bk), _ptr(ksk), byref(h)))
        return CloudKey(self, h)

    def cloud_key_from_file(self, path: str) -> CloudKey:
        h = c_void_p()
        _check(lib().ieache_cloudkey_load_file(self._h, path.encode(), byref(h)))
        return CloudKey(self, h)

    def cloud_key_adopt(self, params: Params, bkfft_dev: int, ksk_dev: int) -> CloudKey:
        h = c_void_p()
        _check(lib().ieache_cloudkey_adopt_device(self._h, byref(params), c_void_p(bkfft_dev), c_void_p(ksk_dev), byref(h)))
        return CloudKey(self, h)

    def keygen(self, params: Params, seed: int, export: bool = False):
        """Keygen/keygen.c on the GPU: returns (SecretKey, CloudKey[, bk_coef, ksk] when export)."""
        sk, ck = c_void_p(), c_void_p()
        bk = ks = None
        if export:
            bk = np.zeros(params.n * 2 * params.bk_l * 2 * 1024, dtype=np.int32)
            ks = np.zeros(1024 * params.ks_t * (1 << params.ks_basebit) * (params.n + 1), dtype=np.int32)
        _check(lib().ieache_keygen(self._h, byref(params), seed, byref(sk), byref(ck), _ptr(bk), _ptr(ks)))
        out = (SecretKey(self, sk, params), CloudKey(self, ck))
        return out + (bk, ks) if export else out

    def secret_key_import(self, params: Params, lwe_key: np.ndarray, tlwe_key: np.ndarray | None = None) -> SecretKey:
        h = c_void_p()
        _check(lib().ieache_secretkey_import(self._h, byref(params), _ptr(np.ascontiguousarray(lwe_key, dtype=np.int32)),
                                             _ptr(None if tlwe_key is None else np.ascontiguousarray(tlwe_key, dtype=np.int32)), byref(h)))
        return SecretKey(self, h, params)

    # ---- gates ------------------------------------------------------------------------------
    def gate_batch(self, key: CloudKey, op, a=None, b=None, c=None, imm: int = 0, count: int | None = None) -> np.ndarray:
        """`count` independent gates on host arrays of shape (count, n+1) int32."""
        opc = OPS[op] if isinstance(op, str) else int(op)
        n = key.params.n
        if count is None:
            count = len(a)
        out = np.empty((count, n + 1), dtype=np.int32)
        _check(lib().ieache_gate_batch(self._h, key._h, opc, _ptr(out), _ptr(a), _ptr(b), _ptr(c), imm, count))
        return out

    def gate_batch_device(self, key: CloudKey, op, out_dev: int, a_dev: int, b_dev: int = 0, c_dev: int = 0, imm: int = 0, count: int = 0):
        opc = OPS[op] if isinstance(op, str) else int(op)
        _check(lib().ieache_gate_batch_device(self._h, key._h, opc, c_void_p(out_dev), c_void_p(a_dev),
                                              c_void_p(b_dev) if b_dev else None, c_void_p(c_dev) if c_dev else None, imm, count))

    def bootstrap_woks(self, key: CloudKey, x: np.ndarray) -> np.ndarray:
        out = np.empty((len(x), 1025), dtype=np.int32)
        _check(lib().ieache_bootstrap_woks(self._h, key._h, _ptr(out), _ptr(x), len(x)))
        return out

    def keyswitch(self, key: CloudKey, ext: np.ndarray) -> np.ndarray:
        out = np.empty((len(ext), key.params.n + 1), dtype=np.int32)
        _check(lib().ieache_keyswitch(self._h, key._h, _ptr(out), _ptr(ext), len(ext)))
        return out

    # ---- device memory ----------------------------------------------------------------------
    def device_alloc(self, nbytes: int) -> int:
        p = c_void_p()
        _check(lib().ieache_device_alloc(self._h, nbytes, byref(p)))
        return p.value

    def device_free(self, ptr: int):
        _check(lib().ieache_device_free(self._h, c_void_p(ptr)))

    def device_copy(self, dst: int, src: int, nbytes: int):
        _check(lib().ieache_device_copy(self._h, c_void_p(dst), c_void_p(src), nbytes))

    def samples_to_device(self, dev: int, host, count: int, n: int):
        _check(lib().ieache_samples_to_device(self._h, c_void_p(dev), _ptr(host), count, n))

    def samples_to_host(self, host, dev: int, count: int, n: int):
        _check(lib().ieache_samples_to_host(self._h, _ptr(host), c_void_p(dev), count, n))

    def sync(self):
        _check(lib().ieache_ctx_sync(self._h))

    # ---- circuits ---------------------------------------------------------------------------
    def circuit(self, kind: int, width: int) -> Circuit:
        return Circuit(kind, width)

    def eval(self, key: CloudKey, circ: Circuit, inputs: np.ndarray) -> np.ndarray:
        """inputs: (n_expr, n_inputs, n+1) int32 -> (n_expr, n_outputs, n+1)."""
        n_expr = inputs.shape[0]
        assert inputs.shape[1] == circ.n_inputs
        out = np.empty((n_expr, circ.n_outputs, key.params.n + 1), dtype=np.int32)
        _check(lib().ieache_circuit_eval(self._h, key._h, circ._h, _ptr(inputs), _ptr(out), n_expr))
        return out

    def eval_device(self, key: CloudKey, circ: Circuit, in_dev: int, out_dev: int, n_expr: int):
        _check(lib().ieache_circuit_eval_device(self._h, key._h, circ._h, c_void_p(in_dev), c_void_p(out_dev), n_expr))

    def keygen_files(self, directory: str, params: Params | None = None, seed_key: int = 0, seed_nbit: int = 0):
        """Keygen/keygen.c: secret.key, cloud.key, nbit.key.  Seeds 0 (default) = operating-system entropy; a non-zero
        seed gives a reproducible key set for tests only (the key set is then no more secret than the seed)"""
        _check(lib().ieache_keygen_files(self._h, directory.encode(), byref(params) if params is not None else None, seed_key, seed_nbit))

    def session(self, cloud_key_path: str, nbit_key_path: str) -> Session:
        h = c_void_p()
        _check(lib().ieache_session_open(self._h, cloud_key_path.encode(), nbit_key_path.encode(), byref(h)))
        return Session(self, h)

    def session_from_keys(self, key: CloudKey, nbit_lwe_key: np.ndarray) -> Session:
        h = c_void_p()
        k = np.ascontiguousarray(nbit_lwe_key, dtype=np.int32)
        _check(lib().ieache_session_open_keys(self._h, key._h, _ptr(k), byref(h)))
        return Session(self, h)

    # ---- the ./cloud process contract ---------------------------------------------------------
    def cloud_run(self, directory: str) -> tuple[int, float]:
        """Cloud/dragonfly_cipher_cloud.py:1233 `subprocess.call("./cloud")` without the subprocess."""
        secs = c_double()
        rc = lib().ieache_cloud_run(self._h, directory.encode(), byref(secs))
        _check(rc)
        return rc, secs.value

    # ---- measurement --------------------------------------------------------------------------
    def set_timing(self, on: bool):
        _check(lib().ieache_ctx_set_timing(self._h, int(on)))

    def kernel_times(self, reset: bool = True):
        br, ks, nbr, nks = c_double(), c_double(), c_uint64(), c_uint64()
        _check(lib().ieache_ctx_kernel_times(self._h, byref(br), byref(ks), byref(nbr), byref(nks), int(reset)))
        return {"blind_rotate_ms": br.value, "keyswitch_ms": ks.value, "blind_rotate_launches": nbr.value, "keyswitch_launches": nks.value}

    def timer_start(self):
        _check(lib().ieache_ctx_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = c_double()
        _check(lib().ieache_ctx_timer_stop(self._h, byref(ms)))
        return ms.value

    def fp64_peak_tflops(self) -> float:
        t = c_double()
        _check(lib().ieache_measure_fp64_peak(self._h, byref(t)))
        return t.value

    @property
    def launch_count(self) -> int:
        return lib().ieache_ctx_launch_count(self._h)
