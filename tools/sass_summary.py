"""Per-kernel SASS evidence from the built library: counts of the instructions that prove the hardware path
(DMMA = FP64 tensor core, UBLKCP = TMA 1-D bulk copy, LDGSTS = cp.async, UCGABAR = cluster barrier, DFMA = FP64 pipe).
    python tools/sass_summary.py llckbdm_b200/libllck.so profiles/r02_sass_summary.md"""
import collections
import re
import subprocess
import sys

lib, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "LDGSTS", "SYNCS", "UCGABAR", "BAR", "ATOM", "RED", "LDG", "STG", "LDS", "STS", "SHFL"]
kern, counts, order = None, collections.defaultdict(collections.Counter), []
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        order.append(kern)
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        op = m.group(1).split(".")[0]
        if op.startswith("UCGABAR"):
            op = "UCGABAR"
        counts[kern]["total"] += 1
        if op in WANT:
            counts[kern][op] += 1
arch = re.findall(r"arch = (sm_\w+)", txt)
rows = ["# SASS summary of " + lib + " (" + ", ".join(sorted(set(arch))) + ")", "",
        "Instruction counts per kernel (static, from `cuobjdump -sass`). DMMA = FP64 tensor-core MMA (`mma.sync.m8n8k4.f64`), "
        "UBLKCP = TMA bulk copy (`cp.async.bulk`), LDGSTS = `cp.async`, UCGABAR = thread-block-cluster barrier.", "",
        "| kernel | total | " + " | ".join(WANT) + " |", "|---|---|" + "---|" * len(WANT)]
tot = collections.Counter()
for k in order:
    c = counts[k]
    rows.append(f"| `{k}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in WANT) + " |")
    tot.update(c)
rows.append("| **all kernels** | %d | " % tot["total"] + " | ".join(str(tot[w]) for w in WANT) + " |")
open(out, "w").write("\n".join(rows) + "\n")
print("\n".join(rows[-3:]))
print(len(order), "kernels")
