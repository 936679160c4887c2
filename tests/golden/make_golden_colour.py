#!/usr/bin/env python
"""Golden for the statistical validation of the colour (checkerboard) sampler: tests/golden/colour_ensemble.

The checkerboard half-sweep of jmm_sweep (csrc/sweep.cuh) is a different Markov chain from the reference's
one-random-particle-per-Step (src/jmmMCState.cpp:1758-1811); both must sample the same Boltzmann distribution.
This script runs the COMPILED REFERENCE (oracle/_ref/jmmOneDMC_ref, OMP_NUM_THREADS=1) on a test/INPUT-style deck
— NLT, LJcut 5.0, NBN 4, lattice spacing 1.12, T = 0.9: config C3's physics — at N = 64, 16 independent seeds x
40 000 sweeps (2 560 000 steps), one thermo row per 100 sweeps (TPI = 100 N), and stores the block means of E and
E^2 and the acceptance counters.  N is small on purpose: a 1-D chain at fixed L relaxes its long-wavelength density
modes in ~N^2 sweeps — at N = 2000 the energy of both samplers still drifts after 3000 sweeps (measured with the
oracle), and a 2 sigma test on a drifting quantity tests the transient, not the distribution; at N = 64 the block
means are flat from the first 4000 sweeps on.  ~25 s per run; needs /root/reference.

    python tests/golden/make_golden_colour.py
"""
import json
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
from oracle import oracle as O  # noqa: E402
from make_golden import deck_with, summarise  # noqa: E402

DECK = """ENSEMBLE   NLT
N          64
L          71.68
T          0.9
NUMSTEPS   2560000
POT        LJcut 5.0
NBN        4
MAXSTEP    0.12
MAXDV      2.0
CPI        100000000
TPI        6400
RBW        0.05
RHONB      1
RHOPI      50000000
GSW        5.0
GNS        1
GBW        0.05
GNB        1
GPI        50000000
SEED       774281
ENGCHECK   10000000
DADJ       10000000
VADJ       10000000
"""


def main(nseeds=16, seed0=774281, workers=8):
    O.build()
    out = HERE / "colour_ensemble"
    out.mkdir(exist_ok=True)

    def one(k):
        with tempfile.TemporaryDirectory() as tmp:
            R = O.run_reference(deck_with(DECK, SEED=seed0 + k), tmp)
            rows = [l.split("\t") for l in Path(R["thermo"]).read_text().splitlines()[1:]]
            return {"seed": seed0 + k, "counters": summarise(R["stdout"])["counters"],
                    "E_blocks": [float(r[1]) for r in rows], "E2_blocks": [float(r[2]) for r in rows]}

    with ThreadPoolExecutor(workers) as ex:
        runs = list(ex.map(one, range(nseeds)))
    (out / "INPUT").write_text(DECK)
    (out / "summary.json").write_text(json.dumps({"note": "thermo rows of the compiled reference: row k = mean over the k-th block of 100 sweeps (TPI = 100 N = 6400 steps); row 0 = step 0",
                                                  "runs": runs}) + "\n")
    print("colour_ensemble:", len(runs), "runs")


if __name__ == "__main__":
    main()
