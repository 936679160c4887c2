#!/usr/bin/env python
"""Launch-shape experiments for k_sweep_fast (C3 / C5 of BASELINE.json): one process, many JMM_SWEEP_* settings.
Usage: sweep_grid.py c3|c5 "K=1,WARPS=24" "K=2" ...   (an empty string = the library's own choice)"""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import jmmonedmc_b200 as J
from jmmonedmc_b200.capi import config
from bench import EXTRA

wl = sys.argv[1]
w = EXTRA[wl]
pot = {"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT}[w["pot"]]
KEYS = ("K", "WARPS", "G", "NSUB", "STAGGER")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
peak = J.lib().jmm_fp64_peak_tflops(0)
print(f"{wl}: fp64 peak {peak:.1f} TFLOP/s")
for spec in sys.argv[2:]:
    for k in KEYS:
        os.environ.pop("JMM_SWEEP_" + k, None)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        os.environ["JMM_SWEEP_" + k] = v
    cfg = config(N=w["N"], pot=pot, nbn=w["nbn"], cutoff=w["cutoff"], ensemble=J.ENS_NLT, L=1.12 * w["N"], T=w["T"],
                 maxStep=w["maxStep"], seed=w["seed"], nchains=w["nchains"], mode=J.MODE_CHECKERBOARD, device=0, arith=J.ARITH_FAST)
    try:
        with J.Handle(cfg) as h:
            stream = torch.cuda.current_stream()
            h.set_stream(stream.cuda_stream)
            h.start()
            for _ in range(3):
                h.sweep(64)
            torch.cuda.synchronize()
            ms, trials = 0.0, 0
            for _ in range(5):
                flush.fill_(1.0)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream); n = h.sweep(64); b.record(stream)
                torch.cuda.synchronize()
                ms += a.elapsed_time(b); trials += n
            v = trials / (ms * 1e-3)
            st = h.get_state(r=False)
            print(f"{spec or 'auto':28s} {v:.4g} trials/s  {v * w['flop'] / 1e12 / peak:.4f} of fp64 peak  "
                  f"E={st['totals'][0][0]:.10g}", flush=True)
    except Exception as e:
        print(f"{spec:28s} failed: {str(e)[:120]}", flush=True)
