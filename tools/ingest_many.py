"""Developer run (GPU box): SURVEY 8 f-4 at the size it was asked for — N request directories (default 4096), each with the
two-operand cloud.data and operator.txt a client pair would have uploaded, evaluated by ONE session call
(ieache_session_compute_dirs: passes of 64 requests through two pinned staging slots), default parameters.
Reports requests/s, gates/s, the peak resident set of the process, and verifies a sample of the answers with the verifier.

usage: python tools/ingest_many.py [N] [workdir] [pass sizes, e.g. 64,256,2048]
"""
import os, resource, shutil, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g

m = g.load_package()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
work = tempfile.mkdtemp(prefix="ieache_ingest_", dir=sys.argv[2] if len(sys.argv) > 2 else None)
rss = lambda: resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0   # MiB
try:
    eng = m.Engine(0)
    keys = os.path.join(work, "keys"); os.makedirs(keys)
    eng.keygen_files(keys)                                   # secret.key, cloud.key, nbit.key (OS entropy)
    rng = np.random.default_rng(7)
    dirs, want, ops = [], [], []
    t0 = time.perf_counter()
    uniq = min(N, 128)                        # distinct operand pairs; the other directories hard-link their cloud.data
    for k in range(N):
        d = os.path.join(work, f"r{k:05d}"); os.makedirs(d)
        u = k % uniq
        if k < uniq:
            a, b = int(rng.integers(1 << 20, 1 << 30)), int(rng.integers(0, 1 << 20))
            m.alice_encrypt(keys, 0, 32, a, os.path.join(d, "cloud.data"))
            m.alice_encrypt(keys, 0, 32, b, os.path.join(d, "cloud.data"), append=True)
            vals = (a, b) if k == 0 else None
            pairs = [(a, b)] if k == 0 else pairs + [(a, b)]
        else:
            os.link(os.path.join(dirs[u], "cloud.data"), os.path.join(d, "cloud.data"))
        op = (1, 2)[(k // uniq + k) & 1]
        a, b = pairs[u]
        open(os.path.join(d, "operator.txt"), "w").write(str(op))
        dirs.append(d); ops.append(op); want.append(a + b if op == 1 else a - b)
    t_gen = time.perf_counter() - t0
    sess = eng.session(os.path.join(keys, "cloud.key"), os.path.join(keys, "nbit.key"))
    gates = sum(160 if o == 1 else 320 for o in ops)
    passes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [256]
    for ps in passes:
        sess.set_pass(ps)
        rss0 = rss()
        t0 = time.perf_counter()
        codes, secs = sess.compute_dirs(dirs)
        wall = time.perf_counter() - t0
        rss1 = rss()
        assert (codes == 0).all()
        print(f"pass={ps}: compute_dirs {wall:.2f} s wall ({secs:.2f} s in circuits) -> {N / wall:.1f} requests/s, {gates / wall:.0f} gate bootstraps/s; "
              f"peak resident set {rss0:.0f} -> {rss1:.0f} MiB", flush=True)
    sample = rng.choice(N, size=min(N, 64), replace=False)
    ok = 0
    for k in sample:
        for f in ("secret.key", "nbit.key"):
            os.symlink(os.path.join(keys, f), os.path.join(dirs[k], f))
        val, sc, w = m.verif_run(dirs[k])
        ok += (val == want[k])
    in_gb = N * 2 * 352 * 2536 / 1e9; out_gb = N * 352 * 2536 / 1e9
    print(f"requests={N} (add32 / sub32 alternating, default parameters) inputs written in {t_gen:.1f} s")
    print(f"compute_dirs: {wall:.2f} s wall ({secs:.2f} s in circuits) -> {N / wall:.1f} requests/s, {gates / wall:.0f} gate bootstraps/s; "
          f"{in_gb:.2f} GB of cloud.data read, {out_gb:.2f} GB of answer.data written")
    print(f"peak resident set: {rss0:.0f} MiB before the call, {rss1:.0f} MiB after ({rss1 - rss0:+.0f} MiB for {N} requests)")
    print(f"verified {ok}/{len(sample)} sampled answers with ieache_verif_run")
    sess.close()
finally:
    shutil.rmtree(work, ignore_errors=True)
