"""Host-side mirror of the Cloud node's driver for the hot path.

The reference's `Cloud/cloud_dynamic.py` is a supervisor loop around `dragonfly_cipher_cloud.py`, whose
`compute()` (Cloud/dragonfly_cipher_cloud.py:1219-1297) writes `operator.txt` and runs
`subprocess.call("./cloud")`, and whose postfix walk (:685-729) chains operators through
`answer.data` -> `cloud.data` on disk (`compute_final()`, :1300-1327).  This module keeps those names and
that file contract but reaches the GPU engine through ctypes: no subprocess, the 114 MB key is parsed and
transformed once per session instead of once per operator, and intermediates stay in memory.

Sockets, Dragonfly key exchange, AES wrapping and BER framing are out of scope."""
from __future__ import annotations

import os

import numpy as np

from . import Engine, EngineError, Session  # noqa: F401

OPCODES = {"+": 1, "-": 2, "*": 4}           # Output/output_dynamic.py opcode mapping
ABORT_FILE_SIZE = 162304                      # Cloud/dragonfly_cipher_cloud.py:1295 (64 records at n = 630)


def _record_bytes(n: int) -> int:
    return 4 + 4 * (n + 1) + 8


def read_block(path: str, n: int, count: int = 352, skip: int = 0) -> np.ndarray:
    """import_gate_bootstrapping_ciphertext_fromFile x count (Cloud/cloud.c:705-766)"""
    rec = _record_bytes(n)
    raw = np.fromfile(path, dtype=np.uint8, count=count * rec, offset=skip * rec)
    if raw.size != count * rec:
        raise EngineError(f"{path}: expected {count} records of {rec} bytes after {skip}, file is too short")
    return np.ascontiguousarray(raw.reshape(count, rec)[:, 4:4 + 4 * (n + 1)]).view(np.int32).reshape(count, n + 1)


def write_block(path: str, block: np.ndarray, variance: float, append: bool = False) -> None:
    """export_gate_bootstrapping_ciphertext_toFile x len(block) (Cloud/cloud.c:826,900)"""
    count, w = block.shape
    rec = np.zeros((count, _record_bytes(w - 1)), dtype=np.uint8)
    rec[:, 0:4] = np.frombuffer(np.int32(42).tobytes(), dtype=np.uint8)
    rec[:, 4:4 + 4 * w] = np.ascontiguousarray(block, dtype=np.int32).view(np.uint8).reshape(count, 4 * w)
    rec[:, 4 + 4 * w:] = np.frombuffer(np.float64(variance).tobytes(), dtype=np.uint8)
    with open(path, "ab" if append else "wb") as f:
        f.write(rec.tobytes())


class CloudNode:
    """One Cloud directory (cloud.key, nbit.key, cloud.data, operator.txt, answer.data)."""

    def __init__(self, directory: str = ".", device: int = 0, engine: Engine | None = None):
        self.dir = directory
        self.engine = engine or Engine(device)
        self.session = self.engine.session(os.path.join(directory, "cloud.key"), os.path.join(directory, "nbit.key"))
        self.n = self.session.params.n
        self.variance = self.session.params.ks_stdev ** 2

    def compute(self, operator: int) -> int:
        """dragonfly_cipher_cloud.py compute(): one operator on cloud.data -> answer.data; returns ./cloud's exit code"""
        with open(os.path.join(self.dir, "operator.txt"), "w") as f:
            f.write(str(operator))
        data = read_block(os.path.join(self.dir, "cloud.data"), self.n, 704)
        rc, ans, secs = self.session.compute(operator, data[:352], data[352:])
        write_block(os.path.join(self.dir, "answer.data"), ans, self.variance)
        with open(os.path.join(self.dir, "timings.txt"), "a") as f:
            f.write(f"\nComputation time: {round(secs, 3)}")
        if operator == 4 and rc == 0:
            with open(os.path.join(self.dir, "averagestandard.txt"), "a") as f:
                f.write(f"{secs:f}\n")
        return rc

    def evaluate(self, postfix: str, operand_files: list[str]) -> int:
        """the whole postfix walk (handshake() loop + computation() + compute_final()) in one call:
        operand_files[k] is client k's 352-record upload; writes answer.data"""
        ops = np.stack([read_block(p, self.n, 352) for p in operand_files])[None]
        rc, ans, counts, _ = self.session.eval_postfix(postfix, np.ascontiguousarray(ops))
        write_block(os.path.join(self.dir, "answer.data"), ans[0][:int(counts[0])], self.variance)
        return rc

    def compute_many(self, directories: list[str]):
        """many queued requests (one directory each: cloud.data + operator.txt) as one batch on the GPU; every
        directory gets its answer.data; returns the exit code ./cloud would have had in each"""
        codes, _ = self.session.compute_dirs(directories)
        return [int(c) for c in codes]

    def close(self):
        self.session.close()
