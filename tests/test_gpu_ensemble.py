"""north_star, third criterion: production Philox runs must give ensemble averages (energy, length/density,
acceptance) within 2 sigma of the block-averaged reference error bars.

Reference side: tests/golden/smalltest_ensemble = 32 independent runs (seeds 92847+k) of the COMPILED
REFERENCE on test/INPUT_smalltest (RELAX line removed so that the deterministic relaxVolume calls of the
first 10^6 steps do not bias the sample), 1 000 000 steps each, TPI 200 000: its thermo.dat.mcs rows ARE block
averages; the first block (200 000 steps) is discarded as equilibration.  Error bar = standard error over the
32 independent runs (the 50k-step blocks of ONE chain are autocorrelated, which understates a single-chain
error bar: a 2 000 000-step single chain sat 2.5 of its own sigmas from the 64-chain mean).
GPU side: 512 independent chains, Philox stream, positions-only arithmetic, in-kernel step-size adaptation,
500 000 steps each, first 200 000 discarded; error bar = standard error over chains."""
import json

import numpy as np
import pytest

from helpers import jmm_config_from_deck

pytestmark = pytest.mark.gpu


def test_production_ensemble_matches_reference_within_2_sigma(J, O, gold):
    g = gold("smalltest_ensemble")
    runs = g["summary"]["runs"]
    cols = g["summary"]["columns"]
    assert len(runs) == 32 and all(len(r["blocks"]) == 6 for r in runs)          # step 0 + 5 blocks
    per_run = np.array([np.mean(r["blocks"][2:], axis=0) for r in runs])         # drop step-0 row and first block
    ref = {"E": per_run[:, cols.index("Econf")], "L": per_run[:, cols.index("L")], "rho": per_run[:, cols.index("rho")]}
    d = O.parse_deck(g["deck_text"])
    C, n_eq, n_run = 512, 200_000, 300_000
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_DEVICE, nchains=C, seed=20261017)
    with J.Handle(cfg) as h:
        h.start()
        h.step(n_eq)
        h.zero_accum()
        c0 = h.get_state(r=False, l=False, totals=False, accum=False)["counters"].astype(np.float64)
        h.step(n_run)
        s = h.get_state()
        checks, disc = h.echeck_stats()
    assert disc == 0 and checks == C * ((n_eq + n_run) // 1000)
    acc = s["accum"] / n_run
    gpu = {"E": acc[:, 4], "L": acc[:, 2], "rho": acc[:, 0]}
    report = {}
    for k in ("E", "L", "rho"):
        m_ref, se_ref = ref[k].mean(), ref[k].std(ddof=1) / np.sqrt(ref[k].size)
        m_gpu, se_gpu = gpu[k].mean(), gpu[k].std(ddof=1) / np.sqrt(C)
        sigma = np.hypot(se_ref, se_gpu)
        report[k] = (m_ref, se_ref, m_gpu, se_gpu, (m_gpu - m_ref) / sigma)
        assert abs(m_gpu - m_ref) < 2 * sigma, f"{k}: reference {m_ref:.5f}+-{se_ref:.5f}, GPU {m_gpu:.5f}+-{se_gpu:.5f}"
    # acceptance ratios (whole-run counters of the reference runs; sampled part of the GPU chains)
    cr = np.array([r["counters"] for r in runs], dtype=np.float64)
    d_ref = cr[:, 0] / (cr[:, 0] + cr[:, 1]); v_ref = cr[:, 2] / (cr[:, 2] + cr[:, 3])
    dc = s["counters"].astype(np.float64) - c0
    d_gpu = dc[:, 0] / (dc[:, 0] + dc[:, 1]); v_gpu = dc[:, 2] / (dc[:, 2] + dc[:, 3])
    for name, a, b in (("displacement", d_ref, d_gpu), ("volume", v_ref, v_gpu)):
        sigma = np.hypot(a.std(ddof=1) / np.sqrt(a.size), b.std(ddof=1) / np.sqrt(b.size))
        # the reference counters include its equilibration phase (first 20 % of the run): allow 0.5 % absolute for that
        assert abs(a.mean() - b.mean()) < 2 * sigma + 0.005, f"{name} acceptance: reference {a.mean():.4f}, GPU {b.mean():.4f}"
        report["acc_" + name] = (a.mean(), b.mean())
    print("ensemble check (ref mean, ref se, gpu mean, gpu se, z):", {k: tuple(round(float(x), 5) for x in v) for k, v in report.items()})
