"""The bench line the driver reads, produced on the GPU at a small size: every key of the contract must be there and
the numbers must be self-consistent (the full-size run is the driver's; this guards the plumbing)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract_small():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--log2-gates", "12", "--steps", "2", "--warmup", "3",
                          "--skip-expression", "--cpu-seconds", "1", "--batch-muladd", "2", "--batch-mul64", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["metric"] == "gate_bootstraps_per_sec" and d["unit"] == "gates/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["value"] > 1000 and abs(d["value"] - 4096 * 2 / (d["ms_per_step"] * 2e-3)) / d["value"] < 1e-6
    assert d["config"]["workload"].startswith("bootsNAND_batch_2^12") and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 2 * 4096 * 631 * 4 and e["d2h_bytes_per_step"] == 4096 * 631 * 4
    assert d["gpu_launches"] >= 4                      # a blind rotation and a key switch per step
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    r = d["roofline"]
    assert r["bound"] == "fp64" and r["unit"] == "TFLOP/s" and r["peak"] > 10 and 0 < r["frac"] < 1
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r and r["hbm_view"]["bound"] == "hbm"
    assert r["kernel"].startswith("blind_rotate_")
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c and c["port_vs_advertised_libtfhe"] > 0
    # BASELINE.json configs 4 / 5, time-boxed per GPU: both present, verified, rates self-consistent
    b = d["expression_batch"]
    assert set(b) == {"muladd", "mul64"}
    assert b["muladd"]["n_expr_per_gpu"] == 2 and b["mul64"]["n_expr_per_gpu"] == 1 and b["muladd"]["n_gpus"] == 1
    assert abs(b["muladd"]["gates_per_s"] - 2 * 11584 / (b["muladd"]["ms_per_step"] * 1e-3)) / b["muladd"]["gates_per_s"] < 1e-6
    assert abs(b["mul64"]["gates_per_s"] - 35296 / (b["mul64"]["ms_per_step"] * 1e-3)) / b["mul64"]["gates_per_s"] < 1e-6
    # the two arms name the same workload
    ref = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2-gates", "12", "--steps", "1",
                          "--warmup", "0", "--ref-seconds", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert ref.returncode == 0, ref.stderr[-2000:]
    assert json.loads(ref.stdout.strip().splitlines()[-1])["config"] == d["config"]
