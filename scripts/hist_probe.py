"""Timing probe for the histogram path: C4 deck, a few thousand chains, step counts growing (prints as it goes)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jmmonedmc_b200 as J
from jmmonedmc_b200.capi import config

C = int(os.environ.get("PROBE_CHAINS", "4096"))
cfg = config(N=80, pot=J.POT_LJ, nbn=-1, ensemble=J.ENS_NPT, relax=1, P=0.5, T=0.5, maxStep=0.1, maxdl=2.0, eci=10000,
             mdai=10 ** 6, mvai=10 ** 6, seed=92847, nchains=C, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE,
             adapt=J.ADAPT_DEVICE, device=0, arith=J.ARITH_FAST)
h = J.Handle(cfg)
g = np.linspace(0.1, 1.0, 256); ids = np.arange(C)
h.set_state(P=g[(ids // 256) % 256], T=g[ids % 256])
if os.environ.get("PROBE_HIST", "1") == "1":
    h.enable_histograms(1000, 0.1, 10, 1000, 200.0, 0.1)
h.start()
print("started", flush=True)
for n in (10, 50, 100, 150, 200, 250, 250):
    t0 = time.perf_counter(); h.step(n); s = h.get_state(r=False); dt = time.perf_counter() - t0
    cnt = s["counters"].sum(axis=0)
    print(f"chains {C} steps {n}: {dt*1e3:.1f} ms  ({dt*1e6/n:.0f} us/step)  counters {cnt.tolist()}", flush=True)
h.close()
