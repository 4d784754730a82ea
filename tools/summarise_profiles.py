"""Turn the scratch outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/:

  <name>_<round>_ncu_full.csv      selected metrics + stall samples of one `ncu --set full` capture (.ncu-rep)
  <name>_<round>_stalls_by_class.csv  the source page of the same capture, aggregated by instruction class:
                                   instructions per launch, share of the stall samples, top stall reasons
  sass_<round>.md                  registers / stack / shared memory (cuobjdump -res-usage) and a SASS opcode histogram
                                   (DFMA / DADD / DMUL / LDS / STS / LDG / SHFL / LDTM / STTM / UBLKCP / SYNCS ...) of
                                   every kernel shipped in libieache_b200.so
  traffic.json                     DRAM bytes of one blind-rotation launch (roofline.traffic)
  launches_<round>.csv             copied as is

usage: python tools/summarise_profiles.py r02   (run in the build container after `gpurun ... tools/profile_round.sh r02`)
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__sass_inst_executed_op_tmem_ldt.sum",
        "smsp__sass_inst_executed_op_tmem_stt.sum", "sass__inst_executed_register_spilling", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def raw_summary(rep, dst):
    rows = list(csv.reader(ncu(rep, "raw").splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(units, vals)))
    with open(dst, "w") as f:
        f.write("metric,unit,value\n")
        f.write('kernel,,"%s"\n' % d["Kernel Name"][1])
        for k in KEEP:
            if k in d:
                f.write(f"{k},{d[k][0]},{d[k][1]}\n")
        for k in sorted(d):
            if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k:
                f.write(f"{k},{d[k][0]},{d[k][1]}\n")


def klass(op):
    b = op.split(".")[0]
    if b in ("DFMA", "DADD", "DMUL"): return "FP64 (DFMA/DADD/DMUL)"
    if op.startswith("IMAD.MOV") or b == "MOV": return "register moves (IMAD.MOV/MOV)"
    if b in ("IMAD",): return "IMAD (address / integer)"
    if b in ("SEL", "FSEL", "LOP3", "SHF", "IADD3", "VIADD", "ISETP", "LEA", "PRMT", "IADD", "PLOP3"): return "INT32 ALU (half rate)"
    if b in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "LDTM", "STTM", "I2F", "F2I", "BAR", "UBLKCP", "SYNCS"): return b
    return "other"


def stalls_by_class(rep, dst):
    rows = list(csv.reader(ncu(rep, "source").splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    inst, samp = collections.Counter(), collections.defaultdict(collections.Counter)
    for r in rows[2:]:
        if len(r) < len(hdr): continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        c = klass(m.group(2) if m else "?")
        inst[c] += int(r[ix["Instructions Executed"]])
        for s in stalls:
            samp[c][s[6:]] += int(r[ix[s]])
    total = sum(sum(v.values()) for v in samp.values()) or 1
    with open(dst, "w") as f:
        f.write("instruction_class,warp_instructions_per_launch,share_of_stall_samples_pct,top_stall_reasons_pct_of_all_samples\n")
        for c, v in sorted(samp.items(), key=lambda kv: -sum(kv[1].values())):
            top = "; ".join(f"{k} {100 * n / total:.1f}" for k, n in v.most_common(4) if n)
            f.write(f'"{c}",{inst[c]},{100 * sum(v.values()) / total:.1f},"{top}"\n')


def sass_report(dst, rnd):
    so = os.path.join(ROOT, "ie-ache_b200", "libieache_b200.so")
    res = subprocess.run(["cuobjdump", "-res-usage", so], stdout=subprocess.PIPE, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True).stdout
    usage, cur = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m: cur = m.group(1); continue
        if cur and "REG:" in line:
            usage[cur] = line.strip(); cur = None
    hist, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m: cur = m.group(1); hist[cur] = collections.Counter(); continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur: hist[cur][m.group(1)] += 1
    demangle = lambda n: subprocess.run(["c++filt", n], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
    cols = ["DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "SHFL", "LDTM", "STTM", "UBLKCP", "SYNCS", "I2F", "F2I", "IMAD", "BAR", "LDL", "STL"]
    with open(dst, "w") as f:
        f.write(f"# Shipped kernels of libieache_b200.so, round {rnd}: resources and SASS opcode histogram\n\n")
        f.write("`cuobjdump -res-usage` and a static count of `cuobjdump -sass` mnemonics (whole kernel, not weighted by trip counts).\n")
        f.write("UBLKCP = bulk (TMA) copy, SYNCS = mbarrier, LDTM/STTM = tensor-memory load/store; no UTC*MMA anywhere: the path has no dense contraction.\n\n")
        f.write("| kernel | " + " | ".join(["REG", "STACK", "SHARED(static)"] + cols) + " |\n|---|" + "---|" * (3 + len(cols)) + "\n")
        for name in sorted(hist, key=demangle):
            u = usage.get(name, "")
            g = lambda k: (re.search(k + r":(\d+)", u) or [None, "?"])[1]
            f.write(f"| `{demangle(name)}` | {g('REG')} | {g('STACK')} | {g('SHARED')} | " + " | ".join(str(hist[name][c]) for c in cols) + " |\n")


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(PROF, exist_ok=True)
    for short, name in (("w12", "blind_rotate_w12"), ("group", "blind_rotate_group"), ("cluster", "blind_rotate_cluster"), ("ks", "keyswitch_staged")):
        rep = os.path.join(OUT, f"prof_{short}_{rnd}.ncu-rep")
        if os.path.exists(rep):
            raw_summary(rep, os.path.join(PROF, f"{name}_{rnd}_ncu_full.csv"))
            stalls_by_class(rep, os.path.join(PROF, f"{name}_{rnd}_stalls_by_class.csv"))
            print("summarised", rep)
    for f in (f"launches_{rnd}.csv", f"bench_{rnd}_n1.json", f"bench_{rnd}_reference.json"):
        if os.path.exists(os.path.join(OUT, f)):
            shutil.copy(os.path.join(OUT, f), os.path.join(PROF, f))
    tr = os.path.join(OUT, f"traffic_{rnd}.csv")
    if os.path.exists(tr):
        shutil.copy(tr, os.path.join(PROF, f"traffic_{rnd}_ncu.csv"))
        vals = {}
        for r in csv.reader(open(tr)):
            if len(r) > 14 and r[12] in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"):
                vals[r[12]] = float(r[14]); kern = r[4]
        if vals:
            k = re.sub(r"^void ", "", kern).split("(")[0]
            json.dump({"kernel": k, "blind_rotate_dram_bytes_per_launch": int(vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]),
                       "dram_bytes_read": int(vals["dram__bytes_read.sum"]), "dram_bytes_written": int(vals["dram__bytes_write.sum"]),
                       "gates_per_launch": 65536, "ms_under_ncu": vals["gpu__time_duration.sum"] / 1e6,
                       "source": f"profiles/traffic_{rnd}_ncu.csv (tools/profile_round.sh)",
                       "algorithmic_bytes_per_launch": 61931520 + 65536 * (2 * 2524 + 4100),
                       "note": "the key (62 MB in each of the two layouts in use) is re-read from DRAM several times per launch: it does not stay in one "
                               "63 MB L2 partition next to 0.6 GB of samples; a few GB/s, no effect on a kernel bound by FP64 issue"},
                      open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    sass_report(os.path.join(PROF, f"sass_{rnd}.md"), rnd)
    print("wrote profiles/ summaries for", rnd)


if __name__ == "__main__":
    main()
