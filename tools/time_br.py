"""Developer probe: device time of the blind-rotation / key-switch kernels for one NAND batch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
eng = m.Engine(0)
count = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
p = m.Params.default(630)
sk, key = eng.keygen(p, seed=1)
rng = np.random.default_rng(1)
ba = rng.integers(0, 2, count).astype(np.int32); bb = rng.integers(0, 2, count).astype(np.int32)
da, db, do = (eng.device_alloc(count * 632 * 4) for _ in range(3))
sk.encrypt_to_device(ba, da, 1); sk.encrypt_to_device(bb, db, 2)
eng.set_timing(True)
best = None
for rep in range(4):
    eng.kernel_times(reset=True)
    eng.gate_batch_device(key, "NAND", do, da, db, count=count); eng.sync()
    kt = eng.kernel_times(reset=True)
    if rep and (best is None or kt["blind_rotate_ms"] < best["blind_rotate_ms"]): best = kt
ok = int((sk.decrypt_from_device(do, count) == 1 - (ba & bb)).sum())
print(f"kernels={eng.pick_kernels(key, count)} count={count} BR={best['blind_rotate_ms']:.2f} ms ({count/best['blind_rotate_ms']*1e3:.0f} gates/s) KS={best['keyswitch_ms']:.2f} ms total={count/(best['blind_rotate_ms']+best['keyswitch_ms'])*1e3:.0f} gates/s correct={ok}/{count}", flush=True)
