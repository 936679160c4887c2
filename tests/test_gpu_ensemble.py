"""north_star, third criterion: production Philox runs must give ensemble averages (energy, length/density,
acceptance) within 2 sigma of the block-averaged reference error bars.

Reference side: tests/golden/smalltest_blocks = the COMPILED REFERENCE on test/INPUT_smalltest (RELAX line
removed so that the deterministic relaxVolume calls of the first 10^6 steps do not bias the sample),
2 000 000 steps, TPI 50 000: its thermo.dat.mcs rows ARE block averages.  The first 4 blocks are discarded.
GPU side: 512 independent chains, Philox stream, positions-only arithmetic, in-kernel step-size adaptation,
300 000 steps each, first 100 000 discarded; error bar = standard error over chains."""
import numpy as np
import pytest

from helpers import jmm_config_from_deck

pytestmark = pytest.mark.gpu


def test_production_ensemble_matches_reference_within_2_sigma(J, O, gold):
    g = gold("smalltest_blocks")
    rows = np.array([[float(x) for x in l.split("\t")] for l in (g["dir"] / "thermo.dat.mcs").read_text().splitlines()[1:]])
    blocks = rows[5:]                                   # row 0 = step 0, rows 1-4 = equilibration
    assert blocks.shape[0] == 36
    ref = {"E": blocks[:, 1], "L": blocks[:, 3], "rho": blocks[:, 6]}
    d = O.parse_deck(g["deck_text"])
    C, n_eq, n_run = 512, 100_000, 200_000
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
    # acceptance ratios: reference = whole-run counters (2e6 steps); GPU = counters of the sampled part
    cr = np.array(g["summary"]["counters"], dtype=np.float64)
    dc = s["counters"].astype(np.float64) - c0
    d_gpu = dc[:, 0] / (dc[:, 0] + dc[:, 1]); v_gpu = dc[:, 2] / (dc[:, 2] + dc[:, 3])
    d_ref, v_ref = cr[0] / (cr[0] + cr[1]), cr[2] / (cr[2] + cr[3])
    # the Swendsen rule drives both to its fixed point; the reference value is one long chain (no error bar):
    # allow 2 sigma of the GPU ensemble spread of a single chain plus 1 % absolute
    assert abs(d_gpu.mean() - d_ref) < 2 * d_gpu.std(ddof=1) / np.sqrt(C) + 0.01
    assert abs(v_gpu.mean() - v_ref) < 2 * v_gpu.std(ddof=1) / np.sqrt(C) + 0.02
    print("ensemble check (ref mean, ref se, gpu mean, gpu se, z):", {k: tuple(round(float(x), 5) for x in v) for k, v in report.items()},
          "acc d", round(float(d_gpu.mean()), 4), round(float(d_ref), 4), "v", round(float(v_gpu.mean()), 4), round(float(v_ref), 4))
