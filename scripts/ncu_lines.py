#!/usr/bin/env python
"""Per CUDA source line: executed warp instructions and stall samples (needs -lineinfo). Usage: ncu_lines.py rep units [topN]"""
import collections, csv, io, subprocess, sys
rep, units = sys.argv[1], float(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cnt, smp, text = collections.Counter(), collections.Counter(), {}
cur_file = None
for r in csv.reader(io.StringIO(out)):
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Line No", "Function Name"): continue
    try: line = int(r[0]); n = int(r[7]); s = int(r[6])
    except ValueError: continue
    key = (cur_file, line)
    cnt[key] += n; smp[key] += s; text.setdefault(key, r[1].strip())
tot, ts = sum(cnt.values()), sum(smp.values())
print(f"total {tot/units:.1f} instr/unit; samples {ts}")
for key, n in cnt.most_common(top):
    print(f"{n/units:7.2f} instr {100*smp[key]/max(ts,1):5.1f}% stall  {key[0]}:{key[1]:<4d} {text[key][:110]}")
