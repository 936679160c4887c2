#!/usr/bin/env python
"""SASS evidence per kernel of libjmmgpu.so: opcode counts that show what the kernels are made of — fp64 pipe (DFMA, DMUL,
DADD, DSETP), MUFU.RCP64H (the fastlj.cuh reciprocal), TMA bulk copies (UBLKCP) and mbarrier traffic (SYNCS), shuffles,
votes, local-memory spills (LDL/STL).  Usage: sass_excerpt.py [lib] > profiles/<round>_sass_excerpt.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent.parent / "jmmonedmc_b200" / "libjmmgpu.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["DFMA", "DMUL", "DADD", "DSETP", "MUFU.RCP64H", "MUFU.EX2", "UBLKCP", "SYNCS", "SHFL", "VOTE", "MATCH", "LDS", "STS", "LDG", "STG",
        "LDL", "STL", "BAR", "WARPSYNC", "ATOM", "RED", "NANOSLEEP"]
cur, counts, total = None, {}, {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("void jmm::", "").replace("void ", "")
        counts[cur] = collections.Counter(); total[cur] = 0
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        total[cur] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {Path(lib).name}: static opcode counts per kernel (sm_100a)")
print(f"{'kernel':78s} {'instr':>7s} " + " ".join(f"{k.replace('MUFU.', ''):>7s}" for k in KEYS))
for k in sorted(counts):
    if total[k] < 50:
        continue
    print(f"{k[:78]:78s} {total[k]:7d} " + " ".join(f"{counts[k][q]:7d}" for q in KEYS))
