#!/bin/bash
# compute-sanitizer on small configurations of every kernel family (memcheck + racecheck + initcheck)
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, math
import jmmonedmc_b200 as J
from jmmonedmc_b200.capi import config
def run(cfg, steps=60, sweep=False):
    with J.Handle(cfg) as h:
        h.start()
        if sweep: h.sweep(steps)
        else: h.step(steps, accept_log=True)
        s = h.get_state(); h.energy()
        if not sweep: h.energy(exact_order=True)
    return s
std = dict(N=10, pot=J.POT_HARMONIC, nbn=1, P=0.7, T=0.4, maxStep=0.1, maxdl=1.0, eci=1, mdai=20, mvai=20, seed=125)
lj = dict(N=12, pot=J.POT_LJ, nbn=-1, P=1.0, T=0.9, maxStep=0.1, maxdl=0.1, eci=10, mdai=20, mvai=20, seed=9, relax=1)
ljc = dict(N=24, pot=J.POT_LJCUT, nbn=3, cutoff=2.5, P=0.5, T=0.7, maxStep=0.15, maxdl=0.4, eci=5, mdai=20, mvai=30, seed=7, relax=1)
which = os.environ.get("SAN_CASE", "all")
cases = {
 "bond": lambda: (os.environ.__setitem__("JMM_BOND", "1"), run(config(nchains=37, **std)), os.environ.pop("JMM_BOND")),
 "solo": lambda: (os.environ.__setitem__("JMM_BOND", "3"), run(config(nchains=37, **std), 150), os.environ.pop("JMM_BOND")),
 "trio": lambda: (os.environ.__setitem__("JMM_BOND", "4"), run(config(nchains=37, **std), 150), run(config(nchains=37, adapt=J.ADAPT_DEVICE, **std), 150), os.environ.pop("JMM_BOND")),
 "crew": lambda: (os.environ.__setitem__("JMM_BOND", "5"), run(config(nchains=37, **std), 150), run(config(nchains=37, adapt=J.ADAPT_DEVICE, **std), 150),
                  os.environ.__setitem__("JMM_SOLO_FORCE_REDO", "1"), run(config(nchains=37, **std), 70), os.environ.pop("JMM_SOLO_FORCE_REDO"), os.environ.pop("JMM_BOND")),
 "bond2": lambda: (os.environ.__setitem__("JMM_BOND", "2"), run(config(nchains=37, **std)), os.environ.__setitem__("JMM_BOND", "1")),
 "lanes": lambda: [(os.environ.__setitem__("JMM_LANES_G", g), run(config(nchains=21, arith=J.ARITH_FAST, adapt=J.ADAPT_DEVICE, **lj), 90),
                    run(config(nchains=21, arith=J.ARITH_FAST, **ljc), 90)) for g in ("8", "4", "32")] + [os.environ.pop("JMM_LANES_G")],
 "lanes80": lambda: (os.environ.__setitem__("JMM_FORCE_SLICE", "1"), os.environ.__setitem__("JMM_SLICE_CHUNK", "13"),
                     run(config(nchains=70, arith=J.ARITH_FAST, N=80, pot=J.POT_LJ, nbn=-1, P=0.5, T=0.5, maxStep=0.1, maxdl=2.0, eci=30, mdai=10**6, mvai=10**6, seed=92847, relax=1), 70),
                     os.environ.pop("JMM_FORCE_SLICE"), os.environ.pop("JMM_SLICE_CHUNK")),
 "coop": lambda: (os.environ.__setitem__("JMM_BOND", "0"), run(config(nchains=37, **std)), run(config(nchains=9, **lj)), run(config(nchains=9, **ljc))),
 "prod": lambda: (os.environ.__setitem__("JMM_COOP_G", "0"), run(config(nchains=70, **lj)), run(config(nchains=70, arith=J.ARITH_FAST, **ljc))),
 "sliced": lambda: (os.environ.__setitem__("JMM_COOP_G", "0"), os.environ.__setitem__("JMM_FORCE_SLICE", "1"), os.environ.__setitem__("JMM_SLICE_CHUNK", "7"), run(config(nchains=70, **lj))),
 "generic": lambda: (os.environ.__setitem__("JMM_COOP_G", "0"), os.environ.__setitem__("JMM_NO_PROD", "1"), run(config(nchains=40, **ljc)), run(config(nchains=3, mode=J.MODE_TABLE, rng_kind=J.RNG_TAUS2, adapt=J.ADAPT_HOST, **lj))),
 "sweep_shapes": lambda: (os.environ.__setitem__("JMM_SWEEP_G", "8"), os.environ.__setitem__("JMM_SWEEP_WARPS", "4"), os.environ.__setitem__("JMM_SWEEP_K", "1"),
                   run(config(N=30000, pot=J.POT_LJ, nbn=20, ensemble=J.ENS_NLT, L=33600.0, T=0.9, maxStep=0.12, seed=3, nchains=1, mode=J.MODE_CHECKERBOARD, arith=J.ARITH_FAST), 12, True),
                   os.environ.__setitem__("JMM_SWEEP_G", "1"), os.environ.__setitem__("JMM_SWEEP_WARPS", "3"),
                   run(config(N=60000, pot=J.POT_LJCUT, nbn=4, cutoff=5.0, ensemble=J.ENS_NLT, L=67200.0, T=0.9, maxStep=0.12, seed=3, nchains=1, mode=J.MODE_CHECKERBOARD, arith=J.ARITH_FAST), 12, True)),
 "sweep": lambda: (run(config(N=3000, pot=J.POT_LJCUT, nbn=4, cutoff=5.0, ensemble=J.ENS_NLT, L=3360.0, T=0.9, maxStep=0.12, seed=3, nchains=2, mode=J.MODE_CHECKERBOARD), 11, True),
                   run(config(N=3000, pot=J.POT_LJ, nbn=20, ensemble=J.ENS_NLT, L=3360.0, T=0.9, maxStep=0.12, seed=3, nchains=2, mode=J.MODE_CHECKERBOARD, arith=J.ARITH_FAST), 25, True)),
}
for k, f in cases.items():
    if which in ("all", k): f(); print("case", k, "ok")
PY
for tool in memcheck racecheck; do
  for c in ${SAN_CASES:-bond bond2 solo trio crew lanes lanes80 coop prod sliced generic sweep sweep_shapes}; do
    echo "== $tool $c"
    SAN_CASE=$c timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|case|Error|error|hazard|Invalid" | head -8
  done
done
