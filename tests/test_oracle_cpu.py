"""CPU suite: pins the oracle (the checker) and covers the host-side logic.  No GPU needed.

The reference has no golden vectors of its own: the goldens under tests/golden/ are
outputs of the reference compiled here (tests/golden/make_golden.py).
"""
import filecmp
import hashlib
import math
import re
from pathlib import Path

import numpy as np
import pytest

from helpers import bits_equal

PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_taus2_known_answers(O, gold):
    w = O.taus2_words(1, 10000)
    assert int(w[-1]) == 2733957125          # GSL rng/test.c: rng_test(gsl_rng_taus2, 1, 10000, 2733957125UL)
    g = gold("smalltest_2000")
    assert np.array_equal(O.taus2_words(92847, 8), g["summary"]["rng_first8"])
    # the stream the compiled reference actually consumed (recorded through the shim)
    assert np.array_equal(O.taus2_words(92847, g["rng"].size), g["rng"])
    assert np.array_equal(O.taus2_words(0, 4), O.taus2_words(1, 4))          # seed 0 -> 1


def test_philox_known_answers(O):
    for ctr, key, want in PHILOX_KAT:
        assert tuple(O.philox(ctr, key)) == want


def test_phi_components(O):
    p = O.phi(O.POT["LJ"], 1.1, math.inf, 1, 5.0)
    r6 = 1 / (1.1 * 1.1 * 1.1) ** 2
    assert p[4] == 4 * r6 and p[2] == 4 * r6 * r6 and p[0] == p[2] - p[4]
    assert p[5] == 24 * r6 and p[3] == 48 * r6 * r6 and p[8] == 144 * r6 and p[7] == 576 * r6 * r6
    assert np.all(O.phi(O.POT["LJcut"], 2.6, 2.5, 1, 5.0) == 0)               # beyond the cut-off
    assert O.phi(O.POT["LJcut"], 2.5, 2.5, 1, 5.0)[0] != 0                     # d <= cutOff is inside
    assert np.all(O.phi(O.POT["LJ"], 1.3, math.inf, 0, 5.0)[[1, 3, 5, 6, 7, 8]] == 0)   # params[0] = 0: no virial
    h = O.phi(O.POT["HARMONIC"], 1.25, math.inf, 1, 8.0)
    assert h[0] == 0.25 * 0.25 and h[1] == (2 / 8.0) * 1.25 * 0.25 and np.all(h[2:] == 0)
    assert O.phi(O.POT["HARMONIC"], -0.1, math.inf, 1, 8.0)[0] == 1e11         # overlap sentinel
    assert O.phi(O.POT["HARMONIC"], 2.0, 2.0, 1, 8.0)[0] == 0                  # d < cutOff is strict


@pytest.mark.parametrize("name", ["smalltest_12", "smalltest_2000", "smalltest_20000", "inputstd", "input_n2000_40"])
@pytest.mark.parametrize("rng", ["taus2", "recorded"])
def test_oracle_reproduces_reference_files(O, gold, tmp_path, name, rng):
    """TABLE mode = the reference's arithmetic: thermo.dat.mcs and config.dat.mcs byte for byte."""
    g = gold(name)
    if rng == "recorded" and "rng" not in g:
        pytest.skip("stream not committed for this deck")
    d = O.parse_deck(g["deck_text"])
    c = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_RECORDED if rng == "recorded" else O.RNG_TAUS2, mode=O.MODE_TABLE))
    if rng == "recorded":
        c.set_recorded(g["rng"])
    c.run_deck(d["NUMSTEPS"], d["TPI"], d["CPI"], tmp_path / "thermo", tmp_path / "config", tmp_path / "log")
    s = g["summary"]
    thermo, config = (tmp_path / "thermo").read_bytes(), (tmp_path / "config").read_bytes()
    if d["POT"] == "HARMONIC":
        # the reference prints uninitialised stack in the HV columns for HARMONIC (src/pot.cpp:116-131)
        strip = lambda b: [l.split(b"\t")[:11] for l in b.splitlines()[1:]]
        assert strip(thermo) == strip((g["dir"] / "thermo.dat.mcs").read_bytes())
    else:
        assert hashlib.md5(thermo).hexdigest() == s["thermo_md5"]
    assert hashlib.md5(config).hexdigest() == s["config_md5"]
    assert [int(x) for x in c.counters] == s["counters"]
    assert "%.8G" % c.totals[0] == s["final_E_printed"]
    assert c.echecks == (s["verified"], s["discrepancy"])
    if rng == "recorded":
        assert c.recorded_cursor() == s["rng_words"]           # consumed exactly the words the reference drew
    if s.get("not_understood"):
        assert d["unknown"] == s["not_understood"]


def test_oracle_full_smalltest_summary(O, gold):
    """5 000 000 steps of test/INPUT_smalltest (~10 s): the survey's headline goldens."""
    g = gold("smalltest_full")
    d = O.parse_deck(g["deck_text"])
    c = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_TAUS2, mode=O.MODE_TABLE))
    c.start()
    c.run(int(d["NUMSTEPS"]))
    s = g["summary"]
    assert [int(x) for x in c.counters] == s["counters"] == [2644332, 1900863, 116301, 338505]
    assert "%.8G" % c.totals[0] == s["final_E_printed"] == "-3.2554505"
    assert c.echecks == (5000, 0)
    # the last config frame is printed at step 5 000 000, before that step's relaxVolume is skipped (sn >= 1e6)
    assert np.allclose(c.r, s["last_frame_r"], rtol=0, atol=5e-7)
    assert abs(c.l - s["last_frame_box"]) < 5e-6
    assert c.step_sizes == (0.5, 5.0)


def test_recompute_mode_tracks_table_mode(O, gold):
    """RECOMPUTE (what the GPU computes) differs from TABLE only through ulp-level drift of the
    incrementally updated rij: same decisions for thousands of steps, energies equal to ~1e-12."""
    d = O.parse_deck(gold("smalltest_2000")["deck_text"])
    a = O.Chain(O.config_from_deck(d, mode=O.MODE_TABLE))
    b = O.Chain(O.config_from_deck(d, mode=O.MODE_RECOMPUTE))
    a.start(); b.start()
    a.run(2000); b.run(2000)
    assert np.array_equal(a.counters, b.counters)
    assert np.allclose(a.r, b.r, rtol=0, atol=1e-12)
    assert abs(a.totals[0] - b.totals[0]) < 1e-10


def test_incremental_totals_equal_fresh_recompute(O):
    from test_gpu_parity import DECKS
    for name, d in DECKS.items():
        c = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=11))
        c.start(); c.run(3000)
        fresh = c.config_totals(1.0, 1)
        assert abs(fresh[0] - c.totals[0]) < 1e-9 * max(1, abs(fresh[0])), name
        assert int(c.counters.sum()) == 3001
        assert np.all(np.diff(c.r) > 0) and np.all(np.abs(c.r) <= c.l / 2)


def test_checkerboard_oracle_is_a_valid_sampler(O):
    """Colour half-sweeps keep the incremental energy consistent and never reorder particles."""
    N, nbn, L, seed = 600, 3, 600 * 1.12, 5
    r = ((np.arange(N) + 0.5) / N - 0.5) * L
    tot = O.totals_of(r, nbn, O.POT["LJcut"], 5.0, 1.0, 1, L)
    cols = []
    for t in range(40):
        col = O.colour_of_step(seed, 0, t, nbn + 1)
        cols.append(col)
        a, dt = O.colour_halfsweep(r, L, nbn, O.POT["LJcut"], 5.0, 0.9, 0.12, seed, 0, t, nbn + 1, col)
        tot += dt
    assert set(cols) == set(range(nbn + 1))
    fresh = O.totals_of(r, nbn, O.POT["LJcut"], 5.0, 1.0, 1, L)
    assert np.allclose(tot, fresh, rtol=1e-11, atol=1e-9)
    assert np.all(np.diff(r) > 0)


@pytest.mark.parametrize("name", ["smalltest_hist", "inputstd_hist"])
def test_oracle_histograms_reproduce_reference_files(O, gold, tmp_path, name):
    """rho.dat.mcs and g<k>.dat.mcs (fgrho/qagrho/ugrho/printRho/printG) byte for byte, including the reference's
    quirks: fav never updates or accumulates the histograms, the old bin is re-derived from r[nm]-md."""
    g = gold(name)
    d = O.parse_deck(g["deck_text"])
    c = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_TAUS2, mode=O.MODE_TABLE))
    c.run_deck_hist(d, tmp_path)
    files = ["rho.dat.mcs", "config.dat.mcs"] + [f"g{k}.dat.mcs" for k in range(int(d["GNS"]))]
    if d["POT"] != "HARMONIC":
        files.append("thermo.dat.mcs")
    for f in files:
        assert (tmp_path / f).read_bytes() == (g["dir"] / f).read_bytes(), f
    assert [int(x) for x in c.counters] == g["summary"]["counters"]


def test_oracle_continues_at_a_step_number(O):
    """jmo_set_sn (the restart branch of setupMCS takes the step count from the last frame, src/jmmMCState.cpp:641):
    the Philox stream and the relaxVolume cadence (every 10 000 steps below 1 000 000, src/Main.cpp:173) follow it."""
    d = dict(N=10, POT="LJ", NBN=-1, CUTOFF=float("inf"), ENSEMBLE="NPT", P=1.0, T=0.9, MAXSTEP=0.1, MAXDV=0.1,
             ENGCHECK=5000, DADJ=4000, VADJ=6000, SEED=92847, RELAX=1)

    def run(sn0, n):
        c = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=3))
        c.start()
        if sn0 is not None:
            c.set_step_number(sn0)
        r0 = c.relax_calls
        for _ in range(n):
            c.step(); c.cadence()
        return c, c.relax_calls - r0

    a, _ = run(None, 300)
    b, _ = run(0, 300)
    assert np.array_equal(a.r.view(np.uint64), b.r.view(np.uint64)) and a.sn == b.sn == 300
    c, _ = run(5000, 300)
    assert c.sn == 5300 and not np.array_equal(a.r.view(np.uint64), c.r.view(np.uint64))     # other Philox blocks
    e, relaxes = run(989_500, 12_000)
    assert e.sn == 1_001_500 and relaxes == 1                                              # 990 000 yes, 1 000 000 no
    f, relaxes = run(1_000_000, 25_000)
    assert relaxes == 0
    assert int(f.counters.sum()) == 25_001
