"""Developer aid: static dispatch-cost estimate of a blind-rotation kernel from its SASS.

Measured on B200 (profiles/README.md, round 2): with 12 resident warps per SM the warp-per-gate kernel runs at
sum(2.2 x FP64 + 2 x other half-rate instructions + 1 x the rest) / 4 cycles per CMux step — FP64, the INT32 ALU ops
and IMAD/MOV each hold a sub-partition's dispatch port for two cycles (FP64 2.1-2.3 by operand pattern).  Code reached
through CALL (the phase functions of br_w12.cu) is only counted where it contains a loop.  This script finds the loops of a kernel in
`cuobjdump -sass` output (backward branches), prints the instruction mix of each and the cost of one step for given
trip counts, so that a restructuring can be judged before GPU time is spent.

usage: sass_cost.py <object-or-so> <kernel-name-regex> [weights in order of loop start address, e.g. 0,0,1,2,6,2 | auto]
       auto: warp-per-gate kernel structure — the loop with most DFMA is the forward transform (x6 per CMux step), the
       one with most DADD the inverse (x2), the forward's parent the per-polynomial loop (x2), their parent the step (x1)
"""
import collections, re, subprocess, sys

HALF = {"DFMA", "DADD", "DMUL", "IMAD", "MOV", "SEL", "FSEL", "LOP3", "SHF", "IADD3", "VIADD", "ISETP", "LEA", "PRMT", "IADD", "FMNMX", "IABS", "FLO", "POPC", "PLOP3"}

def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], stdout=subprocess.PIPE, text=True).stdout
    cur, res = None, {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); res[cur] = []; continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            res[cur].append((int(m.group(1), 16), m.group(2).strip()))
    return res

def opcode(txt):
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", txt)
    return m.group(2) if m else "?"

def main():
    path, pat = sys.argv[1], sys.argv[2]
    auto = len(sys.argv) > 3 and sys.argv[3] == "auto"
    weights = [float(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 and not auto else None
    for name, ins in kernels(path).items():
        if not re.search(pat, name): continue
        print(name, len(ins), "instructions")
        loops = []
        for addr, txt in ins:
            if opcode(txt).startswith("BRA") and not opcode(txt).startswith("BRA.DIV"):
                hexes = re.findall(r"0x([0-9a-f]+)", txt)
                if hexes:
                    tgt = int(hexes[-1], 16)
                    if tgt <= addr: loops.append((tgt, addr))
        loops.sort()
        if auto:
            def own(lo, hi):
                inner = [(a, b) for (a, b) in loops if a >= lo and b <= hi and (a, b) != (lo, hi)]
                return collections.Counter(opcode(t).split(".")[0] for a, t in ins if lo <= a <= hi and not any(x <= a <= y for x, y in inner))
            counts = [own(lo, hi) for lo, hi in loops]
            fwd = max(range(len(loops)), key=lambda k: counts[k]["DFMA"])
            inv = max(range(len(loops)), key=lambda k: counts[k]["DADD"])
            parent = lambda k: min((j for j in range(len(loops)) if j != k and loops[j][0] <= loops[k][0] and loops[j][1] >= loops[k][1]),
                                   key=lambda j: loops[j][1] - loops[j][0], default=None)
            q = parent(fwd)
            step = parent(inv)
            weights = [0.0] * len(loops)
            weights[fwd] = 6.0; weights[inv] = 2.0
            if q is not None: weights[q] = 2.0
            if step is not None: weights[step] = 1.0
        total = 0.0
        for k, (lo, hi) in enumerate(loops):
            inner = [(a, b) for (a, b) in loops if a >= lo and b <= hi and (a, b) != (lo, hi)]
            body = [(a, t) for a, t in ins if lo <= a <= hi and not any(x <= a <= y for x, y in inner)]
            c = collections.Counter(opcode(t).split(".")[0] for a, t in body)
            half = sum(v for o, v in c.items() if o in HALF)
            cost = 2 * half + (len(body) - half) + 0.2 * sum(v for o, v in c.items() if o in ("DFMA", "DADD", "DMUL"))  # FP64: 2.2 (tools/dfma_probe.cu)
            w = weights[k] if weights and k < len(weights) else 0
            total += w * cost
            top = ", ".join(f"{o}:{v}" for o, v in c.most_common(14))
            print(f"  loop {k} [{lo:#x},{hi:#x}] own instr {len(body)} half-rate {half} cost {cost} weight {w}\n     {top}")
        if weights: print(f"  dispatch cycles per step {total:.0f} -> {total / 4:.0f} SM cycles per gate-step -> {148 * 1.965e9 / (total / 4) / 630 / 1e3:.1f} k gates/s at n = 630")

if __name__ == "__main__":
    main()
