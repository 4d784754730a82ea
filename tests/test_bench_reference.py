"""bench.py's reference arm (the CPU restatement of libtfhe on the host cores) runs without a GPU: the JSON
contract the driver reads is checked here; the B200 arm itself is exercised on the GPU box."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-seconds", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "gate_bootstraps_per_sec" and line["unit"] == "gates/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"] == {"workload": "bootsNAND_batch_2^20_per_gpu", "gates_per_step_per_gpu": 1 << 20,
                              "params": "n=630,N=1024,k=1,l=3,Bgbit=7,t=8,basebit=2",
                              "parallelism": "dp1 (key replicated, no data-path collective)"}
    assert line["cpu_baseline"]["port_vs_advertised_libtfhe"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--ref-seconds", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
