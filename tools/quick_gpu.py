"""Developer probe run under gpurun: smoke + raw timing of NAND batches (not the bench)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g
import oracle_bind as ob

t0 = time.time(); g.smoke(); print("smoke time", time.time() - t0, flush=True)
m = g.load_package()
eng = m.Engine(0)
orc = ob.Oracle()
n = 630
ks = orc.keygen(ob.params_default(n), seed=2024)
key = eng.cloud_key_from_arrays(m.Params.default(n), ks.bk_coef(), ks.ksk())
rng = np.random.default_rng(1)
for count in [int(x) for x in (sys.argv[1:] or ["592", "4736", "32768"])]:
    bits_a = rng.integers(0, 2, count).astype(np.int32); bits_b = rng.integers(0, 2, count).astype(np.int32)
    a = ks.encrypt(bits_a, 5); b = ks.encrypt(bits_b, 6)
    da, db, do = (eng.device_alloc(count * 632 * 4) for _ in range(3))
    eng.samples_to_device(da, a, count, n); eng.samples_to_device(db, b, count, n)
    eng.set_timing(True)
    for rep in range(3):
        eng.kernel_times(reset=True)
        t = time.time(); eng.gate_batch_device(key, "NAND", do, da, db, count=count); eng.sync(); dt = time.time() - t
        kt = eng.kernel_times(reset=True)
        print(f"count={count} rep={rep} wall={dt*1e3:.2f} ms  {count/dt:.0f} gates/s  BR={kt['blind_rotate_ms']:.2f} ms KS={kt['keyswitch_ms']:.2f} ms", flush=True)
    out = np.empty((count, n + 1), dtype=np.int32); eng.samples_to_host(out, do, count, n)
    got = ks.decrypt(out); ref = 1 - (bits_a & bits_b)
    print("  correct:", int((got == ref).sum()), "/", count, flush=True)
    for p in (da, db, do): eng.device_free(p)
