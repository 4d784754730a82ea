"""Generates tests/golden/gates_n16_seed4242.npz from the CPU oracle (oracle/liboracle.so).

The reference ships no golden vectors and libtfhe is not available, so these vectors
pin OUR oracle (keygen, encryption, all gates) and are what the CUDA path is compared against on the
GPU box, where the oracle's outputs for the same inputs are also recomputed live.
Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_bind as ob  # noqa: E402

N_SMALL, SEED = 16, 4242
orc = ob.Oracle()
ks = orc.keygen(ob.params_default(N_SMALL), seed=SEED)
rng = np.random.default_rng(SEED)
bits = {k: rng.integers(0, 2, 8).astype(np.int32) for k in ("bits_a", "bits_b", "bits_c")}
a, b, c = ks.encrypt(bits["bits_a"], 1), ks.encrypt(bits["bits_b"], 2), ks.encrypt(bits["bits_c"], 3)
out = {"n": N_SMALL, "seed": SEED, "lwe_key": ks.lwe_key(), "a": a, "b": b, "c": c, **bits}
for name, op in ob.OPS.items():
    if name in ("NOT", "COPY", "CONST"):
        continue
    out["out_" + name] = ks.gate_batch(op, a, b, c if name == "MUX" else None, threads=1)
np.savez_compressed(os.path.join(HERE, "gates_n16_seed4242.npz"), **out)
print("wrote", os.path.join(HERE, "gates_n16_seed4242.npz"))
