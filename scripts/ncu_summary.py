#!/usr/bin/env python
"""Condense an .ncu-rep into the numbers DESIGN.md / profiles/ quote: duration, occupancy, issue rate,
fp64 pipe, DRAM traffic, stall reasons and the SASS opcode mix.  Usage: ncu_summary.py rep [units-per-launch] [--traffic KEY]
--traffic KEY also records dram__bytes_read.sum + dram__bytes_write.sum of the captured launch under KEY in
profiles/traffic.json, which bench.py quotes as roofline.traffic."""
import collections
import csv
import io
import subprocess
import sys


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[-1]


def to_bytes(val, unit):
    x = float(val.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    argv = list(sys.argv)
    key = None
    if "--traffic" in argv:
        i = argv.index("--traffic"); key = argv[i + 1]; del argv[i:i + 2]
    rep = argv[1]
    units = float(argv[2]) if len(argv) > 2 else None
    hdr, unit, val = raw(rep)
    d = {h: (v, u) for h, u, v in zip(hdr, unit, val)}
    def g(k):
        return d.get(k, ("", ""))
    print(f"# {rep}\nkernel: {g('Kernel Name')[0]}")
    if key:
        import json
        from pathlib import Path
        import os
        tj = Path(os.environ.get("JMM_TRAFFIC_JSON", str(Path(__file__).resolve().parent.parent / "profiles" / "traffic.json")))
        t = json.loads(tj.read_text()) if tj.exists() else {}
        t[key] = {"dram_bytes": int(to_bytes(*g("dram__bytes_read.sum")) + to_bytes(*g("dram__bytes_write.sum"))),
                  "kernel": g("Kernel Name")[0], "duration_us_under_ncu": g("gpu__time_duration.sum")[0] + " " + g("gpu__time_duration.sum")[1],
                  "source": "ncu --set full --clock-control none, " + Path(rep).name}
        tj.write_text(json.dumps(t, indent=1, sort_keys=True) + "\n")
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum"]
    for k in keys:
        v, u = g(k)
        if v != "":
            print(f"  {k:70s} {v} {u}")
    stalls = []
    for h in hdr:
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h:
            try:
                stalls.append((float(d[h][0].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
            except ValueError:
                continue
    for x, name in sorted(stalls, reverse=True)[:8]:
        print(f"  stall {name:32s} {x:6.3f} warps per issue-active cycle")
    for k in ("sm__cycles_active.avg", "sm__cycles_active.max", "sm__cycles_elapsed.avg", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"):
        v, u = g(k)
        if v != "":
            print(f"  {k:70s} {v} {u}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h2 = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    ia, ie = h2.index("Source"), h2.index("Instructions Executed")
    ops, tot = collections.Counter(), 0
    for r in rows[rows.index(h2) + 1:]:
        try:
            n = int(r[ie])
        except (ValueError, IndexError):
            continue
        s = r[ia].strip().split()
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        ops[op] += n
        tot += n
    scale = units if units else 1.0
    print(f"  warp instructions: {tot:.4g}" + (f"  = {tot / units:.1f} per unit" if units else ""))
    print("  opcode mix: " + ", ".join(f"{op} {n / scale:.1f}" if units else f"{op} {100 * n / tot:.1f}%" for op, n in ops.most_common(22)))
    fp64 = sum(n for op, n in ops.items() if op in ("DADD", "DMUL", "DFMA", "DSETP", "MUFU"))
    print(f"  fp64-pipe instructions: {100 * fp64 / tot:.1f} % of all")


if __name__ == "__main__":
    main()
