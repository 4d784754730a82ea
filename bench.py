#!/usr/bin/env python
"""bench.py — gate bootstraps/sec of the B200 engine on BASELINE.json's microbench config
(configs[1]: 2^20 independent bootsNAND on synthetic ciphertexts, per GPU).

A step is one pass of the hot path (linear pre-combination -> blind rotation -> sample extraction ->
key switch) over one batch of 2^20 NAND gates.  `value` is measured with inputs resident in HBM;
`e2e` goes through the host-buffer C-ABI call (ieache_gate_batch) with pinned host inputs and outputs,
host<->device copies inside the timed region.  N > 1 (torchrun): every rank owns one GPU, the cloud key
is generated on rank 0 and broadcast once over NCCL, each rank runs its own 2^20 gates (weak scaling,
no data-path collective), time = max over ranks.

Every line also carries `expression_batch`: BASELINE.json configs 4 and 5 time-boxed per GPU (512 independent a*b+c
expressions and 128 independent 64-bit multiplies per GPU through the levelised circuits, sharded by expression,
every result decrypted and checked on its rank), so that the scaling record holds the circuit workloads next to the
independent gates.

`--impl reference` times the CPU restatement of libtfhe (oracle/, kind "port": libtfhe itself is not in
/root/reference nor in the image) on all host threads, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "gate_bootstraps_per_sec"
UNIT = "gates/s"
N_LWE = 630
FLOP_PER_BOOTSTRAP = 163e6   # n*[(kpl+k+1)*F_T(1024) + kpl*(k+1)*(N/2)*8], FP64
BK_BYTES = 61_931_520        # transform-domain bootstrapping key
KS_BYTES_GATHERED = 15_507_456  # expected key-switch rows gathered per gate


PARAMS_STR = "n=630,N=1024,k=1,l=3,Bgbit=7,t=8,basebit=2"
LIBTFHE_ADVERTISED_MS_PER_GATE = 13.0   # tfhe.github.io: "about 13 ms per binary gate" (spqlios-fma, one core, other hardware)
KERNEL_NAMES = {1: "blind_rotate_cluster_kernel<3>", 2: "blind_rotate_pair_kernel<3>", 41: "blind_rotate_kernel<3,0,true>",
                70: "blind_rotate_w12_kernel<3>"}


def workload_config(log2_gates: int, world: int) -> dict:
    """identical in the b200 arm and in the reference arm: the workload, not the run"""
    return {"workload": f"bootsNAND_batch_2^{log2_gates}_per_gpu", "gates_per_step_per_gpu": 1 << log2_gates, "params": PARAMS_STR,
            "parallelism": f"dp{world} (key replicated, no data-path collective)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md 'clocks' line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=5)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    return rank, world, local


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def cpu_port_rate(seconds_budget: float, threads: int = 0):
    """Oracle (CPU restatement of libtfhe) on `threads` host threads over independent NAND gates."""
    import oracle_bind as ob
    orc = ob.Oracle()
    # torchrun exports OMP_NUM_THREADS=1; the baseline must still use every host core it may run on
    threads = threads or max(orc.max_threads(), len(os.sched_getaffinity(0)))
    ks = orc.keygen(ob.params_default(N_LWE), seed=2024)
    batch = max(threads * 2, 8)
    rng = np.random.default_rng(3)
    a = ks.encrypt(rng.integers(0, 2, batch).astype(np.int32), 1)
    b = ks.encrypt(rng.integers(0, 2, batch).astype(np.int32), 2)
    ks.gate_batch(ob.OPS["NAND"], a[:threads], b[:threads], threads=threads)  # warm caches / FFT tables
    done, t0 = 0, time.perf_counter()
    while True:
        ks.gate_batch(ob.OPS["NAND"], a, b, threads=threads)
        done += batch
        dt = time.perf_counter() - t0
        if dt >= seconds_budget:
            break
    ks.free()
    return done / dt, threads, done, dt


def cpu_port_single_thread_ms(n_gates: int = 100) -> float:
    """one host thread, mean over n_gates bootstrapped NAND gates of the oracle, in ms per gate"""
    import oracle_bind as ob
    orc = ob.Oracle()
    ks = orc.keygen(ob.params_default(N_LWE), seed=2025)
    rng = np.random.default_rng(4)
    a = ks.encrypt(rng.integers(0, 2, n_gates).astype(np.int32), 1)
    b = ks.encrypt(rng.integers(0, 2, n_gates).astype(np.int32), 2)
    ks.gate_batch(ob.OPS["NAND"], a[:2], b[:2], threads=1)
    t0 = time.perf_counter()
    ks.gate_batch(ob.OPS["NAND"], a, b, threads=1)
    dt = time.perf_counter() - t0
    ks.free()
    return 1e3 * dt / n_gates


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return 0
    per_step = args.ref_seconds or max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    rates = []
    for i in range(args.warmup + args.steps):
        r, threads, done, dt = cpu_port_rate(per_step)
        if i >= args.warmup:
            rates.append((r, done, dt))
    total = sum(d for _, d, _ in rates)
    secs = sum(t for _, _, t in rates)
    value = total / secs
    sample = f"{rates[0][1]} independent bootsNAND per step on {threads} host threads (bounded sample of the 2^{args.log2_gates}-gate workload)"
    single_ms = cpu_port_single_thread_ms(20)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, len(rates)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 torus + f64 transform", "data": "synthetic",
        "config": workload_config(args.log2_gates, max(1, world)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "single_thread_ms_per_gate": single_ms,
                         "port_vs_advertised_libtfhe": single_ms / LIBTFHE_ADVERTISED_MS_PER_GATE,
                         "note": "the port is slower per gate than libtfhe advertises for spqlios-fma (13 ms): ratios against this arm are "
                                 "ratios against our port, not against libtfhe"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


CIRCUIT_WORKLOADS = {
    "muladd": ("a*b+c on 32-bit operands (mul32 then 64-bit add, Cloud/cloud.c:115-218 + :18-51)", "4"),
    "mul64": ("64-bit multiply (2 x mul64 + split, Cloud/cloud.c:2568-2616)", "5"),
}


def expression_batch(m, eng, sk, key, rank, world, which: str, n_expr: int, steps: int = 1, warmup: int = 0):
    """`n_expr` independent expressions PER GPU through the levelised circuit (every level = one blind-rotation + one
    key-switch launch over all expressions of the rank); expression e of the whole job is drawn from seed (base, e) and
    lives on rank e // n_expr (contiguous shards, ie-ache_b200/dist.py shard_range); no data-path collective.  Every
    decrypted result is checked against the plaintext arithmetic on its rank.  Returns the aggregate figures."""
    import torch
    import torch.distributed as dist
    from ieache_b200.dist import shard_range
    kind, width = (m.CIRC_MULADD, 32) if which == "muladd" else (m.CIRC_MUL, 64)
    circ = eng.circuit(kind, width)
    lo, hi = shard_range(world * n_expr, rank, world)
    nw = circ.n_inputs // 32
    vals = np.stack([np.random.default_rng([4000, e]).integers(0, 2 ** 31, size=nw, dtype=np.int64) for e in range(lo, hi)])
    vals[:, -1] = 0                                   # carry block: encryptions of 0
    if kind == m.CIRC_MULADD:
        vals[:, 3] = 0                                # high chunk of c
    bits = ((vals[:, :, None] >> np.arange(32)) & 1).astype(np.int32).reshape(-1)
    d_in = eng.device_alloc(bits.size * m.DEVICE_STRIDE * 4)
    d_res = eng.device_alloc(n_expr * circ.n_outputs * m.DEVICE_STRIDE * 4)
    sk.encrypt_to_device(bits, d_in, seed=500 + rank)

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        eng.eval_device(key, circ, d_in, d_res, n_expr)
    barrier()
    launches0 = eng.launch_count
    eng.timer_start()
    for _ in range(steps):
        eng.eval_device(key, circ, d_in, d_res, n_expr)
    ms_total = eng.timer_stop()
    barrier()
    launches = eng.launch_count - launches0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    out_bits = sk.decrypt_from_device(d_res, n_expr * circ.n_outputs).reshape(n_expr, -1, 32).astype(np.int64)
    words = (out_bits << np.arange(32)).sum(axis=2)
    for e in range(n_expr):
        got = sum(int(w) << (32 * i) for i, w in enumerate(words[e]))
        v = [int(x) for x in vals[e]]
        want = v[0] * v[1] + v[2] if kind == m.CIRC_MULADD else (v[0] | v[1] << 32) * (v[2] | v[3] << 32)
        if got != want:
            raise SystemExit(f"rank {rank}: expression {lo + e} decrypts to {got}, expected {want}")
    eng.device_free(d_in); eng.device_free(d_res)
    name, cfg = CIRCUIT_WORKLOADS[which]
    return {"workload": f"{n_expr} independent {name} per GPU: {circ.bootstraps} bootstraps, {circ.levels} levels each "
                        f"(time-boxed subset of BASELINE.json config {cfg})",
            "n_expr_per_gpu": n_expr, "n_gpus": world, "gates_per_s": world * n_expr * circ.bootstraps * steps / (ms_total * 1e-3),
            "ms_per_step": ms_total / steps, "ms_per_expression": ms_total / steps / n_expr, "gpu_launches": launches,
            "verified": "every decrypted result equals the plaintext arithmetic on its rank"}


def run_circuit_workload(args, m, eng, sk, key, rank, world, local, bcast):
    """--workload muladd / mul64: the circuit workload as the whole bench line"""
    import torch.distributed as dist
    sampler = ClockSampler(local)
    sampler.start()
    r = expression_batch(m, eng, sk, key, rank, world, args.workload, args.n_expr, steps=args.steps, warmup=args.warmup)
    clocks = sampler.stop()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": r["gates_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 torus + f64 transform", "data": "synthetic",
            "config": {"workload": r["workload"], "params": PARAMS_STR, "parallelism": f"dp{world} by expression (key replicated, no data-path collective)"},
            "run_info": {"ms_per_expression": r["ms_per_expression"], "key_replication": bcast, "verified": r["verified"]},
            "e2e": None, "gpu_launches": r["gpu_launches"], "clocks": clocks, "roofline": None, "cpu_baseline": None}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-gates", type=int, default=20, help="gates per step and per GPU (default 2^20, BASELINE.json config 2)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--workload", default="nand", choices=["nand", "muladd", "mul64"],
                    help="nand: BASELINE.json config 2 (the metric's configuration, default); muladd / mul64: a time-boxed subset of "
                         "configs 4 / 5 (independent a*b+c expressions / 64-bit multiplies, --n-expr per GPU) through the levelised circuits")
    ap.add_argument("--n-expr", type=int, default=64, help="expressions per GPU and step for --workload muladd / mul64")
    ap.add_argument("--ref-seconds", type=float, default=0.0, help="--impl reference: seconds of CPU work per step (default: sized so the run ends within ~2 minutes)")
    ap.add_argument("--batch-muladd", type=int, default=512, help="expression_batch: a*b+c expressions per GPU (0 = skip)")
    ap.add_argument("--batch-mul64", type=int, default=128, help="expression_batch: 64-bit multiplies per GPU (0 = skip)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-expression", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    m = g.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the engine (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    nccl_init_ms = 0.0
    if world > 1:
        t0 = time.perf_counter()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        warm = torch.zeros(1, device="cuda")
        dist.all_reduce(warm)                      # communicator bring-up (rings / NVLS setup) happens on the first collective
        torch.cuda.synchronize()
        nccl_init_ms = 1e3 * (time.perf_counter() - t0)
    eng = m.Engine(local)
    params = m.Params.default(N_LWE)
    count = 1 << args.log2_gates

    # ---- cloud key: generated on rank 0's GPU (Keygen/keygen.c's role), replicated once over NCCL
    keep = None
    if rank == 0:
        sk, key = eng.keygen(params, seed=314_1592_657)      # the reference's seed triple {314,1592,657}
        lwe, tlwe = sk.export()
    else:
        sk = key = None
        lwe = np.zeros(N_LWE, dtype=np.int32)
        tlwe = np.zeros(1024, dtype=np.int32)
    bcast = {"nccl_init_ms": nccl_init_ms, "broadcast_ms": 0.0, "bytes": 0}
    if world > 1:
        from ieache_b200 import dist as idist
        timings = {}
        key, keep = idist.broadcast_cloud_key(eng, key, params, src=0, timings=timings)
        bcast.update(broadcast_ms=timings["broadcast_ms"], bytes=timings["bytes"],
                     GBps=timings["bytes"] / (timings["broadcast_ms"] * 1e-3) / 1e9 if timings["broadcast_ms"] else None,
                     note="nccl_init_ms = init_process_group + first collective (communicator bring-up); broadcast_ms = the two "
                          "key arrays, CUDA events around the NCCL broadcasts, max over ranks")
        kt = torch.from_numpy(np.concatenate([lwe, tlwe])).cuda()
        dist.broadcast(kt, src=0)                   # the bench's own decryption key (a test fixture, not part of the path)
        torch.cuda.synchronize()
        if rank != 0:
            both = kt.cpu().numpy()
            lwe, tlwe = both[:N_LWE].copy(), both[N_LWE:].copy()
            sk = eng.secret_key_import(params, lwe, tlwe)

    if args.workload != "nand":
        return run_circuit_workload(args, m, eng, sk, key, rank, world, local, bcast)

    # ---- synthetic ciphertexts, made on the GPU (Client/alice.c's role): 2 x count samples, resident in HBM
    rng = np.random.default_rng(1000 + rank)
    bits_a = rng.integers(0, 2, count).astype(np.int32)
    bits_b = rng.integers(0, 2, count).astype(np.int32)
    nbytes = count * m.DEVICE_STRIDE * 4
    d_a, d_b, d_out = eng.device_alloc(nbytes), eng.device_alloc(nbytes), eng.device_alloc(nbytes)
    sk.encrypt_to_device(bits_a, d_a, seed=11 + 2 * rank)
    sk.encrypt_to_device(bits_b, d_b, seed=12 + 2 * rank)

    def barrier():
        eng.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident arm -------------------------------------------------------------------
    for _ in range(args.warmup):
        eng.gate_batch_device(key, "NAND", d_out, d_a, d_b, count=count)
    eng.sync()
    eng.set_timing(True)
    eng.kernel_times(reset=True)
    sampler = ClockSampler(local)
    barrier()
    launches0 = eng.launch_count
    sampler.start()
    eng.timer_start()
    for _ in range(args.steps):
        eng.gate_batch_device(key, "NAND", d_out, d_a, d_b, count=count)
    ms_total = eng.timer_stop()
    clocks = sampler.stop()
    barrier()
    launches = eng.launch_count - launches0
    kt = eng.kernel_times(reset=True)
    eng.set_timing(False)
    ms_total = max_over_ranks(ms_total)
    value = world * count * args.steps / (ms_total * 1e-3)

    # correctness of what was timed: decrypt a random 2^14 subset of the last step against NAND's truth table
    sub = np.random.default_rng(7).choice(count, size=min(count, 1 << 14), replace=False)
    got = sk.decrypt_from_device(d_out, count)
    bad = int((got[sub] != 1 - (bits_a[sub] & bits_b[sub])).sum())
    if bad:
        raise SystemExit(f"rank {rank}: {bad} wrong NAND results in the verification subset")

    # ---- roofline of the dominant kernel (blind rotation) ----------------------------------------
    peaks = measured_peaks()
    fp64_peak = eng.fp64_peak_tflops()
    br_launches = max(1, kt["blind_rotate_launches"])
    br_ms = kt["blind_rotate_ms"] / br_launches
    gates_per_launch = count * args.steps / br_launches
    achieved_tflops = FLOP_PER_BOOTSTRAP * gates_per_launch / (br_ms * 1e-3) / 1e12
    br_kernel = eng.pick_kernels(key, int(gates_per_launch))[0]
    traffic, traffic_source = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("kernel") == KERNEL_NAMES.get(br_kernel):
            traffic = tj.get("blind_rotate_dram_bytes_per_launch")
            traffic_source = "static: ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch of this size, " + tj.get("source", "profiles/")
    except (OSError, ValueError):
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    # memory view: BK streamed once per launch (reuse across all gates of the launch) + per-gate inputs/outputs
    alg_bytes = BK_BYTES + gates_per_launch * (2 * 2524 + 4100)
    roofline = {
        "kernel": KERNEL_NAMES.get(br_kernel, str(br_kernel)), "bound": "fp64", "achieved": achieved_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved_tflops / fp64_peak if fp64_peak else None, "traffic": traffic, "traffic_source": traffic_source,
        "peak_source": "dense FP64 FMA microbenchmark run live by this bench (MEASURED_PEAKS.json has no FP64 figure)",
        "flop_per_gate": FLOP_PER_BOOTSTRAP, "gates_per_launch": gates_per_launch, "ms_per_launch": br_ms,
        "share_of_step": kt["blind_rotate_ms"] / ms_total if world == 1 else None,
        "hbm_view": {"bound": "hbm", "achieved": alg_bytes / (br_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": alg_bytes / (br_ms * 1e-3) / 1e9 / hbm_peak,
                     "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                     "note": "algorithmic bytes = BK once per launch + 2 inputs + 1 extracted sample per gate; the kernel is bound by FP64 issue, not by HBM"},
        "keyswitch": {"ms_per_launch": kt["keyswitch_ms"] / max(1, kt["keyswitch_launches"]),
                      "achieved_GBps": KS_BYTES_GATHERED * gates_per_launch / (kt["keyswitch_ms"] / max(1, kt["keyswitch_launches"]) * 1e-3) / 1e9,
                      "share_of_step": kt["keyswitch_ms"] / ms_total if world == 1 else None},
    }

    # ---- end-to-end arm: host buffers through ieache_gate_batch -------------------------------------
    e2e = None
    if not args.skip_e2e:
        n1 = N_LWE + 1
        h_a, h_b, h_out = (m.pinned_array((count, n1)) for _ in range(3))
        eng.samples_to_host(h_a, d_a, count, N_LWE)
        eng.samples_to_host(h_b, d_b, count, N_LWE)
        for p in (d_a, d_b, d_out):
            eng.device_free(p)
        d_a = d_b = d_out = None
        lib = m.lib()
        import ctypes

        def e2e_step():
            rc = lib.ieache_gate_batch(eng._h, key._h, m.OPS["NAND"], h_out.ctypes.data_as(ctypes.c_void_p),
                                       h_a.ctypes.data_as(ctypes.c_void_p), h_b.ctypes.data_as(ctypes.c_void_p), None, 0, count)
            if rc:
                raise SystemExit(lib.ieache_last_error().decode())

        for _ in range(max(1, min(args.warmup, 2))):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()          # returns after the device->host copy of the step's results
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        # the result read back by the user is checked too
        ph = np.empty(len(sub), dtype=np.int64)
        s_lwe = lwe.astype(np.int64)
        rows = h_out[sub].astype(np.int64)
        ph = (rows[:, N_LWE] - rows[:, :N_LWE] @ s_lwe) & 0xFFFFFFFF
        got_bits = ((ph ^ 0x80000000) - 0x80000000 > 0).astype(np.int32)
        if int((got_bits != 1 - (bits_a[sub] & bits_b[sub])).sum()):
            raise SystemExit(f"rank {rank}: wrong results on the end-to-end path")
        e2e = {"value": world * count * args.steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 2 * count * n1 * 4,
               "d2h_bytes_per_step": count * n1 * 4, "api": "ieache_gate_batch (C ABI, pinned host buffers)"}

    # ---- 3-operand expression latency (BASELINE.json metric, config 3), rank 0 at N = 1 only -----------
    expression = None
    if rank == 0 and world == 1 and not args.skip_expression:
        circ = eng.circuit(m.CIRC_MULADD, 32)        # a*b+c: mul32 then 64-bit add, 11 584 bootstraps, 257 levels
        def run_expr(n_expr, reps):
            rng2 = np.random.default_rng(99)
            vals = rng2.integers(0, 2 ** 31, size=(n_expr, 5), dtype=np.int64)
            vals[:, 3:] = 0                           # high chunk of c and the carry block are encryptions of 0
            ebits = ((vals[:, :, None] >> np.arange(32)) & 1).astype(np.int32).reshape(-1)
            d_in = eng.device_alloc(ebits.size * m.DEVICE_STRIDE * 4)
            d_res = eng.device_alloc(n_expr * circ.n_outputs * m.DEVICE_STRIDE * 4)
            sk.encrypt_to_device(ebits, d_in, seed=77)
            times = []
            for r in range(reps + 1):
                eng.sync(); t0 = time.perf_counter()
                eng.eval_device(key, circ, d_in, d_res, n_expr); eng.sync()
                if r:
                    times.append(time.perf_counter() - t0)
            ob_ = sk.decrypt_from_device(d_res, n_expr * circ.n_outputs).reshape(n_expr, 2, 32).astype(np.int64)
            words = (ob_ << np.arange(32)).sum(axis=2)
            good = all(int(words[e, 0]) | (int(words[e, 1]) << 32) == int(vals[e, 0]) * int(vals[e, 1]) + int(vals[e, 2]) for e in range(n_expr))
            eng.device_free(d_in); eng.device_free(d_res)
            if not good:
                raise SystemExit("a*b+c: wrong decrypted result")
            return float(np.median(times))
        lat = run_expr(1, 3)
        nb = 64
        tb = run_expr(nb, 1)
        expression = {"workload": "a*b+c on 32-bit operands: mul32 then 64-bit add (Cloud/cloud.c circuits), 11584 bootstraps, 257 levels",
                      "latency_ms": 1e3 * lat, "batch": {"n_expr": nb, "seconds": tb, "gates_per_s": nb * circ.bootstraps / tb,
                                                          "ms_per_expression": 1e3 * tb / nb},
                      "reference_paper": "A+B*C 329 s on a single-core i7 VM (AC058.pdf p.4, whole run, other hardware)",
                      "verified": "decrypted results equal a*b+c"}

    # ---- BASELINE.json configs 4 / 5, time-boxed per GPU, at every N ------------------------------------
    batches = {}
    if args.batch_muladd > 0:
        batches["muladd"] = expression_batch(m, eng, sk, key, rank, world, "muladd", args.batch_muladd)
    if args.batch_mul64 > 0:
        batches["mul64"] = expression_batch(m, eng, sk, key, rank, world, "mul64", args.batch_mul64)
    for b in batches.values():
        b["fraction_of_nand_rate"] = b["gates_per_s"] / value

    # ---- CPU baseline (rank 0, N = 1 only) ---------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        r, threads, done, dt = cpu_port_rate(args.cpu_seconds)
        cpu = {"value": r, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{done} independent bootsNAND in {dt:.1f} s on {threads} threads (oracle/: C restatement of libtfhe, default parameters)",
               "single_thread_ms_per_gate": cpu_port_single_thread_ms(100 if args.cpu_seconds >= 5 else 10)}
        cpu["port_vs_advertised_libtfhe"] = cpu["single_thread_ms_per_gate"] / LIBTFHE_ADVERTISED_MS_PER_GATE

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32 torus + f64 transform", "data": "synthetic",
            "config": workload_config(args.log2_gates, world),
            "run_info": {"l2": f"inputs {2 * count * 2528 / 1e9:.1f} GB per GPU, larger than L2 (126 MB); no flush needed",
                         "key_replication": bcast, "verified": f"{len(sub)} decrypted results per rank against the NAND truth table",
                         "blind_rotate_kernel": KERNEL_NAMES.get(br_kernel, str(br_kernel))},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "expression": expression,
            "expression_batch": batches,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
