"""Developer probe: latency / throughput of whole circuits (BASELINE.json configs 3-5) on one GPU."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
m = g.load_package()
eng = m.Engine(0)
p = m.Params.default(630)
sk, key = eng.keygen(p, seed=1)
n = 630
def run(kind, width, n_expr, reps=2):
    circ = eng.circuit(kind, width)
    rng = np.random.default_rng(kind * 100 + width)
    nw = circ.n_inputs // 32
    vals = rng.integers(0, 2 ** 31, size=(n_expr, nw), dtype=np.int64)
    vals[:, -1] = 0                      # carry block = encrypted zeros
    if kind == m.CIRC_MULADD: vals[:, 3] = 0
    bits = ((vals[:, :, None] >> np.arange(32)) & 1).astype(np.int32).reshape(-1)
    d_in = eng.device_alloc(bits.size * 632 * 4); d_out = eng.device_alloc(n_expr * circ.n_outputs * 632 * 4)
    sk.encrypt_to_device(bits, d_in, seed=5)
    best = 1e9
    for r in range(reps + 1):
        eng.sync(); t = time.perf_counter()
        eng.eval_device(key, circ, d_in, d_out, n_expr); eng.sync()
        dt = time.perf_counter() - t
        if r: best = min(best, dt)
    out_bits = sk.decrypt_from_device(d_out, n_expr * circ.n_outputs).reshape(n_expr, -1, 32)
    words = (out_bits.astype(np.int64) << np.arange(32)).sum(axis=2)
    ok = 0
    for e in range(n_expr):
        got = sum(int(w) << (32 * i) for i, w in enumerate(words[e]))
        v = [int(x) for x in vals[e]]
        if kind == m.CIRC_MULADD: want = v[0] * v[1] + v[2]
        elif kind == m.CIRC_MUL:
            nc = width // 32
            a = sum(v[i] << (32 * i) for i in range(nc)); b = sum(v[nc + i] << (32 * i) for i in range(nc)); want = a * b
        else: want = None
        ok += (want is None) or (got == want)
    print(f"kind={kind} width={width} n_expr={n_expr}: {best*1e3:.1f} ms  -> {circ.bootstraps*n_expr/best:.0f} gates/s, {best/n_expr*1e3:.2f} ms/expr, levels={circ.levels} correct={ok}/{n_expr}", flush=True)
    eng.device_free(d_in); eng.device_free(d_out)
for spec in (sys.argv[1:] or ["5,32,1", "5,32,16", "5,32,128", "4,64,32", "1,32,1"]):
    k, w, e = (int(x) for x in spec.split(","))
    run(k, w, e)
