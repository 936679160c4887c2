"""GPU parity tests: the CUDA path, called through the C ABI (include/jmm_gpu.h), against the oracle.

Bars: bit-exact for positions, accept/reject sequences, counters and (in the reference-order paths)
totals and running sums; 1e-12 relative for sums whose order of addition differs (north_star).
"""
import math

import numpy as np
import pytest

from helpers import bits_equal, exact_totals, jmm_config_from_deck, oracle_accept_log, rel_err, totals_close

pytestmark = pytest.mark.gpu

PHILOX_KAT = [  # Random123 kat_vectors, philox4x32-10
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_device_generators(J, O, gold):
    for ctr, key, want in PHILOX_KAT:
        got, _ = J.rng_selftest(ctr, key, 1, 0)
        assert tuple(got) == want
    _, words = J.rng_selftest((0,) * 4, (0, 0), 92847, 4096)
    assert np.array_equal(words, gold("smalltest_2000")["rng"][:4096])      # the reference's own stream
    _, w1 = J.rng_selftest((0,) * 4, (0, 0), 1, 10000)
    assert int(w1[9999]) == 2733957125                                       # GSL rng/test.c known answer


@pytest.mark.parametrize("pot,nbn,cutoff", [("LJ", -1, math.inf), ("LJ", 3, math.inf), ("LJcut", -1, 2.5),
                                             ("LJcut", 4, 5.0), ("HARMONIC", 1, math.inf), ("HARMONIC", 2, 1.6)])
@pytest.mark.parametrize("N,C", [(10, 7), (80, 33), (257, 3)])
def test_configuration_totals(J, O, pot, nbn, cutoff, N, C):
    rng = np.random.default_rng(N * 1000 + C + nbn)
    from jmmonedmc_b200.capi import config
    P = {"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT, "HARMONIC": J.POT_HARMONIC}[pot]
    l = rng.uniform(1.0, 1.3, C) * N
    # ordered, non-overlapping particles (LJ diverges at 0), jittered lattice
    r = np.stack([((np.arange(N) + 0.5) / N - 0.5) * l[c] + rng.uniform(-0.2, 0.2, N) for c in range(C)])
    with J.Handle(config(N=N, pot=P, nbn=nbn, cutoff=cutoff, nchains=C, T=1.0, P=1.0)) as h:
        h.set_state(r=r, l=l)
        exact = h.energy(exact_order=True)
        fast = h.energy(exact_order=False)
    want = np.stack([O.totals_of(r[c], nbn, O.POT[pot], cutoff, 1.0, 1, l[c]) for c in range(C)])
    assert bits_equal(exact, want), "reference-order totals must be bit-exact"
    truth = np.stack([exact_totals(r[c], nbn, pot, cutoff, l[c]) for c in range(C)])
    assert totals_close(want, truth, 1e-12)                      # the oracle's serial sums against exact sums
    assert totals_close(fast, truth, 1e-12), "parallel totals within 1e-12 relative (north_star)"


def _lockstep(J, O, gold, name, nsteps, rng_kind, chunk=None):
    g = gold(name)
    d = O.parse_deck(g["deck_text"])
    words = g.get("rng")
    if words is None or rng_kind == "taus2-as-recorded":
        words = O.taus2_words(int(d["SEED"]), 3 * nsteps + 64)
    # oracle: TABLE mode (bit-exact against the compiled reference, see test_oracle_vs_reference.py)
    oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_TAUS2, mode=O.MODE_TABLE))
    oc.start()
    kind = J.RNG_TAUS2 if rng_kind == "taus2" else J.RNG_RECORDED
    cfg = jmm_config_from_deck(J, d, rng_kind=kind, mode=J.MODE_TABLE, adapt=J.ADAPT_HOST)
    with J.Handle(cfg) as h:
        h.start()
        s0 = h.get_state()
        assert bits_equal(s0["r"][0], oc.r) and bits_equal(s0["totals"][0][:2], oc.totals[:2])
        assert bits_equal(s0["l"], [oc.l]) and bits_equal(s0["accum"][0][:10], oc.accum[:10])
        want_log = oracle_accept_log(oc, nsteps)
        logs = []
        if chunk is None:
            logs.append(h.step(nsteps, rng_stream=words if kind == J.RNG_RECORDED else None, accept_log=True))
        else:
            done = 0
            while done < nsteps:
                n = min(chunk, nsteps - done)
                logs.append(h.step(n, rng_stream=words if (kind == J.RNG_RECORDED and done == 0) else None, accept_log=True))
                done += n
        got_log = np.concatenate(logs)[:, 0]
        s = h.get_state()
        assert np.array_equal(got_log & 3, want_log), "accept/reject sequence differs from the reference's"
        assert bits_equal(s["r"][0], oc.r), "final positions must be bit-identical"
        assert bits_equal(s["l"], [oc.l])
        assert np.array_equal(s["counters"][0], oc.counters)
        nc = 2 if d["POT"] == "HARMONIC" else 9
        assert bits_equal(s["totals"][0][:nc], oc.totals[:nc])
        na = 10 if d["POT"] == "HARMONIC" else 12
        assert bits_equal(s["accum"][0][:na], oc.accum[:na])
        ms, mv = h.get_step_sizes()
        assert bits_equal([ms[0], mv[0]], list(oc.step_sizes))
        assert h.echeck_stats() == oc.echecks
        if kind == J.RNG_RECORDED and g.get("rng") is not None and nsteps == int(d["NUMSTEPS"]):
            assert h.stream_cursor == g["summary"]["rng_words"]
        return s, g


def test_lockstep_smalltest_recorded_stream(J, O, gold):
    """north_star: fed the stream recorded from the compiled reference, reproduce INPUT_smalltest exactly."""
    s, g = _lockstep(J, O, gold, "smalltest_2000", 2000, "recorded")
    assert [int(x) for x in s["counters"][0]] == g["summary"]["counters"]
    assert np.allclose(s["r"][0], g["summary"]["last_frame_r"], rtol=0, atol=5e-7)     # %.8G text of the reference
    assert float("%.8G" % s["totals"][0][0]) == float(g["summary"]["final_E_printed"])


def test_lockstep_smalltest_chunked_launches(J, O, gold):
    _lockstep(J, O, gold, "smalltest_2000", 2000, "recorded", chunk=333)


def test_lockstep_smalltest_20000_device_taus2(J, O, gold):
    """Same deck, taus2 generated on the device; 20 000 steps = 20 adjustments, 2 periodic relaxations."""
    s, g = _lockstep(J, O, gold, "smalltest_20000", 20000, "taus2")
    assert [int(x) for x in s["counters"][0]] == g["summary"]["counters"]
    assert float("%.8G" % s["totals"][0][0]) == float(g["summary"]["final_E_printed"])


def test_lockstep_inputstd(J, O, gold):
    s, g = _lockstep(J, O, gold, "inputstd", 10, "recorded")
    assert [int(x) for x in s["counters"][0]] == g["summary"]["counters"]
    assert float("%.8G" % s["l"][0]) == g["summary"]["last_frame_box"] or abs(s["l"][0] - g["summary"]["last_frame_box"]) < 5e-6


def test_lockstep_input_n2000(J, O, gold):
    s, g = _lockstep(J, O, gold, "input_n2000_40", 40, "recorded")
    assert [int(x) for x in s["counters"][0]] == g["summary"]["counters"]
    assert float("%.8G" % s["totals"][0][0]) == float(g["summary"]["final_E_printed"])


def test_recorded_stream_exhaustion_is_an_error(J, O, gold):
    g = gold("smalltest_2000")
    d = O.parse_deck(g["deck_text"])
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_RECORDED, mode=J.MODE_TABLE, adapt=J.ADAPT_HOST)
    with J.Handle(cfg) as h:
        h.start()
        with pytest.raises(J.JmmError) as e:
            h.step(100, rng_stream=g["rng"][:50])
        assert e.value.status == -4


DECKS = {
    "std": dict(N=10, POT="HARMONIC", NBN=1, CUTOFF=math.inf, ENSEMBLE="NPT", P=0.7, T=0.4, MAXSTEP=0.1, MAXDV=1.0,
                ENGCHECK=1, DADJ=100, VADJ=100, SEED=125, RELAX=0),
    "std_cut": dict(N=13, POT="HARMONIC", NBN=1, CUTOFF=1.35, ENSEMBLE="NPT", P=0.4, T=0.6, MAXSTEP=0.2, MAXDV=1.5,
                    ENGCHECK=7, DADJ=130, VADJ=90, SEED=31, RELAX=0),
    "small": dict(N=10, POT="LJ", NBN=-1, CUTOFF=math.inf, ENSEMBLE="NPT", P=1.0, T=0.9, MAXSTEP=0.1, MAXDV=0.1,
                  ENGCHECK=1000, DADJ=1000, VADJ=1000, SEED=92847, RELAX=1),
    "ljcut_nbn": dict(N=24, POT="LJcut", NBN=3, CUTOFF=2.5, ENSEMBLE="NPT", P=0.5, T=0.7, MAXSTEP=0.15, MAXDV=0.4,
                      ENGCHECK=50, DADJ=200, VADJ=300, SEED=7, RELAX=1),
    "nlt": dict(N=16, POT="LJ", NBN=2, CUTOFF=math.inf, ENSEMBLE="NLT", L=20.0, T=0.8, MAXSTEP=0.2, MAXDV=0.1,
                ENGCHECK=10, DADJ=0, VADJ=0, SEED=3, RELAX=0),
}


ENGINES = {   # which kernel serves a JMM_MODE_RECOMPUTE + Philox handle (jmm_gpu.cu: launch_step_table)
    "coop": {"JMM_COOP_G": "16", "JMM_BOND": "0"},     # coop.cuh, 16 lanes per chain
    "bond": {"JMM_COOP_G": "16", "JMM_BOND": "1"},      # bond.cuh, k_chains_step_bond (HARMONIC NBN 1 decks; others fall to coop.cuh)
    "bond2": {"JMM_COOP_G": "16", "JMM_BOND": "2"},     # bond.cuh, k_chains_step_bond2 (registers-only, deferred ECheck): the default
    "solo": {"JMM_COOP_G": "16", "JMM_BOND": "3"},      # solo.cuh, k_chains_step_solo (one chain per thread, a warp per SM; HARMONIC NBN 1 decks)
    "trio": {"JMM_COOP_G": "16", "JMM_BOND": "4"},      # solo.cuh, k_chains_step_trio (trials | Philox | ECheck + sums in three warps)
    "trioredo": {"JMM_COOP_G": "16", "JMM_BOND": "4", "JMM_SOLO_FORCE_REDO": "1"},   # ... every CTA reports a discrepancy: bond.cuh repeats the launch
    "crew": {"JMM_COOP_G": "16", "JMM_BOND": "5"},      # solo.cuh, k_chains_step_crew (displacement | volume | Philox | virial + sums | ECheck)
    "crewredo": {"JMM_COOP_G": "16", "JMM_BOND": "5", "JMM_SOLO_FORCE_REDO": "1"},
    "coop32": {"JMM_COOP_G": "32"},
    "coop8": {"JMM_COOP_G": "8"},
    "prod": {"JMM_COOP_G": "0"},                        # prod.cuh, one chain per thread, shared tile
    "sliced": {"JMM_COOP_G": "0", "JMM_FORCE_SLICE": "1", "JMM_SLICE_CHUNK": "7"},   # prod.cuh, persistent time-sliced launch
    "generic": {"JMM_COOP_G": "0", "JMM_NO_PROD": "1"}, # chains.cuh
}


@pytest.mark.parametrize("name", list(DECKS))
@pytest.mark.parametrize("mode", ["recompute", "table", "recompute-coop", "recompute-bond", "recompute-bond2", "recompute-solo", "recompute-trio", "recompute-trioredo", "recompute-crew", "recompute-crewredo", "recompute-coop32", "recompute-coop8", "recompute-prod", "recompute-sliced", "recompute-generic"])
def test_philox_many_chains_bit_exact(J, O, name, mode, monkeypatch):
    """Production stream, several chains per launch, host-side adaptation (glibc log on both sides):
    every chain must equal the oracle bit for bit, including after adjustments and relaxations —
    whichever kernel (cooperative, per-thread production, generic) serves the handle."""
    d = DECKS[name]
    C, nsteps, id0 = 37, 1500, 1000
    if "-" in mode:
        mode, engine = mode.split("-")
        for k, v in ENGINES[engine].items():
            monkeypatch.setenv(k, v)
    jm = J.MODE_RECOMPUTE if mode == "recompute" else J.MODE_TABLE
    om = O.MODE_RECOMPUTE if mode == "recompute" else O.MODE_TABLE
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=jm, adapt=J.ADAPT_HOST, nchains=C, chain_id0=id0)
    with J.Handle(cfg) as h:
        h.start()
        log = h.step(nsteps, accept_log=True)
        s = h.get_state()
        ms, mv = h.get_step_sizes()
    nc = 2 if d["POT"] == "HARMONIC" else 9
    na = 10 if d["POT"] == "HARMONIC" else 12
    for c in range(C):
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=om, chain_id=id0 + c))
        oc.start()
        want_log = oracle_accept_log(oc, nsteps)
        assert np.array_equal(log[:, c] & 3, want_log), f"chain {c}: accept sequence"
        assert bits_equal(s["r"][c], oc.r), f"chain {c}: positions"
        assert bits_equal(s["l"][c:c + 1], [oc.l])
        assert np.array_equal(s["counters"][c], oc.counters)
        assert bits_equal(s["totals"][c][:nc], oc.totals[:nc]), f"chain {c}: totals"
        assert bits_equal(s["accum"][c][:na], oc.accum[:na]), f"chain {c}: running sums"
        assert bits_equal([ms[c], mv[c]], list(oc.step_sizes))


@pytest.mark.parametrize("engine", ["coop", "prod", "sliced"])
def test_continue_at_a_step_number_matches_oracle(J, O, engine, monkeypatch):
    """jmm_set_step_number (restart at the step count of a frame, src/jmmMCState.cpp:641): the Philox blocks and the
    relaxVolume cadence (every 10 000 steps below 1 000 000, src/Main.cpp:173) follow the step number.  From step
    989 500 a run of 12 000 steps relaxes at 990 000 and must not at 1 000 000; every chain equals the oracle."""
    for k, v in ENGINES[engine].items():
        monkeypatch.setenv(k, v)
    d = dict(DECKS["small"], ENGCHECK=5000, DADJ=4000, VADJ=6000)
    C, sn0, nsteps, id0 = 33, 989_500, 12_000, 40
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_HOST, nchains=C, chain_id0=id0)
    with J.Handle(cfg) as h:
        h.start()
        h.set_step_number(sn0)
        assert h.step_number == sn0
        h.step(nsteps)
        assert h.step_number == sn0 + nsteps
        s = h.get_state()
        ms, mv = h.get_step_sizes()
    for c in (0, 7, 32):
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=id0 + c))
        oc.start()
        oc.set_step_number(sn0)
        relax0 = oc.relax_calls
        for _ in range(nsteps):
            oc.step(); oc.cadence()
        assert oc.relax_calls == relax0 + 1                       # 990 000 only
        assert bits_equal(s["r"][c], oc.r) and bits_equal(s["l"][c:c + 1], [oc.l])
        assert np.array_equal(s["counters"][c], oc.counters)
        assert bits_equal(s["totals"][c], oc.totals) and bits_equal(s["accum"][c], oc.accum)
        assert bits_equal([ms[c], mv[c]], list(oc.step_sizes))


@pytest.mark.parametrize("variant", ["unroll4", "unroll8", "sliced", "default", "lanes2", "lanes4", "lanes8", "lanes16", "lanes32",
                                     "lanes8-sliced", "lanes4-sliced"])
@pytest.mark.parametrize("name", ["small", "ljcut_nbn", "nlt"])
def test_fast_arithmetic_same_trajectory_totals_within_1e12(J, O, name, variant, monkeypatch):
    """JMM_ARITH_FAST (fastlj.cuh): one reciprocal per partner, r^-6/r^-12 differences only — in prod.cuh (one chain
    per thread; 4 or 8 partners in flight, plain and time-sliced launches; 40 chains = a ragged last tile) and in
    lanes.cuh (2 ... 32 lanes per chain, what a FAST handle with few chains gets by default; one chunk and
    time-sliced).  Positions and the accept/reject sequence must equal the oracle's exactly; the nine totals and
    twelve sums agree to 1e-12."""
    if variant.startswith("lanes"):
        monkeypatch.setenv("JMM_LANES_G", variant.split("-")[0][5:])
        if variant.endswith("sliced"):
            monkeypatch.setenv("JMM_FORCE_SLICE", "1")
            monkeypatch.setenv("JMM_SLICE_CHUNK", "7")
    elif variant != "default":
        monkeypatch.setenv("JMM_LANES_G", "0")
        monkeypatch.setenv("JMM_COOP_G", "0")
    if variant == "unroll4":
        monkeypatch.setenv("JMM_PROD_UNROLL", "4")
    if variant == "sliced":
        monkeypatch.setenv("JMM_FORCE_SLICE", "1")
        monkeypatch.setenv("JMM_SLICE_CHUNK", "7")
    d = DECKS[name]
    C, nsteps, id0 = 40, 1500, 7
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_HOST, nchains=C, chain_id0=id0,
                               arith=J.ARITH_FAST)
    with J.Handle(cfg) as h:
        h.start()
        log = h.step(nsteps, accept_log=True)
        s = h.get_state()
        fresh = h.energy(exact_order=True)
    for c in range(C):
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=id0 + c))
        oc.start()
        want_log = oracle_accept_log(oc, nsteps)
        assert np.array_equal(log[:, c] & 3, want_log), f"chain {c}: accept sequence"
        assert bits_equal(s["r"][c], oc.r), f"chain {c}: positions"
        assert bits_equal(s["l"][c:c + 1], [oc.l]) and np.array_equal(s["counters"][c], oc.counters)
        assert totals_close(s["totals"][c], oc.totals, 1e-12), f"chain {c}: totals"
        # against a fresh recompute only the energies are comparable: qavLJ leaves HV untouched and adds the
        # ideal term to Vir (src/jmmMCState.cpp:1675-1686), in the reference as here
        for k in (0, 2, 4):
            assert abs(s["totals"][c][k] - fresh[c][k]) <= 1e-11 * (abs(fresh[c][2]) + abs(fresh[c][4]))
        assert np.allclose(s["accum"][c], oc.accum, rtol=1e-11, atol=1e-9)
    with pytest.raises(J.JmmError):
        J.Handle(jmm_config_from_deck(J, DECKS["std"], arith=J.ARITH_FAST))          # HARMONIC has no fast arithmetic


def _sweep_shape_pt(C):
    """per-chain (P, T) of a RunJobs-style sweep (scripts/RunJobs.bash:16-27), a sqrt(C) x sqrt(C)-ish grid"""
    side = int(math.ceil(math.sqrt(C)))
    grid = np.linspace(0.1, 1.0, side)
    P = np.array([grid[c // side] for c in range(C)])
    T = np.array([grid[c % side] for c in range(C)])
    return P, T


SWEEP_DECK = dict(N=80, POT="LJ", NBN=-1, CUTOFF=math.inf, ENSEMBLE="NPT", P=0.5, T=0.5, MAXSTEP=0.1, MAXDV=2.0,
                  ENGCHECK=10000, DADJ=1000000, VADJ=1000000, SEED=92847, RELAX=1)


@pytest.mark.parametrize("engine", ["prod-reference", "sliced-reference", "coop8-reference", "prod-fast", "sliced-fast",
                                    "lanes8-fast", "lanes4-fast", "lanes8sliced-fast", "lanes16-fast", "lanes2-fast", "team-fast", "teamsliced-fast"])
@pytest.mark.parametrize("start", ["from0", "from1e6"])
def test_sweep_shape_matches_oracle(J, O, engine, start, monkeypatch):
    """The shape bench.py times as C4 (RunJobs deck: N = 80, LJ, NBN -1, NPT, RELAX, ENGCHECK 10000; per-chain P, T set
    through jmm_set_state; device-side cadence) on every kernel that can serve it: 427 chains (13 full tiles of one
    chain per thread + a ragged one; 107 warps of four chains), from step 0 across the in-kernel relaxVolume at step
    10 000 (src/Main.cpp:173-176) and the ECheck there, and from step 10^6 (the production phase, no relaxation).
    Sampled chains against the oracle: positions, box, counters, accept sequence bit-identical; totals and sums
    bit-identical in reference arithmetic, <= 1e-12 in fast arithmetic."""
    eng, arith = engine.split("-")
    env = {"prod": {"JMM_COOP_G": "0", "JMM_LANES_G": "0"},
           "sliced": {"JMM_COOP_G": "0", "JMM_LANES_G": "0", "JMM_FORCE_SLICE": "1", "JMM_SLICE_CHUNK": "97"},
           "coop8": {"JMM_COOP_G": "8"},
           "lanes8": {"JMM_LANES_G": "8"}, "lanes4": {"JMM_LANES_G": "4"}, "lanes16": {"JMM_LANES_G": "16"},
           "lanes2": {"JMM_LANES_G": "2"},
           "lanes8sliced": {"JMM_LANES_G": "8", "JMM_FORCE_SLICE": "1", "JMM_SLICE_CHUNK": "97"},
           "team": {"JMM_LANES_G": "8", "JMM_TEAM": "1"},                     # team.cuh: 7 loop warps + a bookkeeper warp per 28 chains
           "teamsliced": {"JMM_LANES_G": "8", "JMM_TEAM": "1", "JMM_FORCE_SLICE": "1", "JMM_SLICE_CHUNK": "97"}}[eng]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    d = SWEEP_DECK
    C, id0 = 427, 5000
    sn0, nsteps = (0, 10_400) if start == "from0" else (1_000_000, 2_000)
    # Fast arithmetic keeps positions and decisions bit-identical as long as no step feeds the RUNNING totals back into
    # the positions.  relaxVolume does (its Newton step uses the running E, src/jmmMCState.cpp:2452): the fast totals
    # differ from the reference-order ones by rounding (<= 1e-12), so from the first in-run relaxation (step 10 000) on
    # the box length and the positions carry that rounding.  Bit-exact up to step 9 999; after it 1e-9 absolute.
    exact_until = 9_999 if (arith == "fast" and start == "from0") else nsteps
    P, T = _sweep_shape_pt(C)
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_DEVICE, nchains=C, chain_id0=id0,
                               arith=J.ARITH_FAST if arith == "fast" else J.ARITH_REFERENCE)
    log_steps = 300
    with J.Handle(cfg) as h:
        h.set_state(P=P, T=T)
        h.start()
        if sn0:
            h.set_step_number(sn0)
        log = h.step(log_steps, accept_log=True)
        h.step(exact_until - log_steps)
        s_exact = h.get_state()
        if nsteps > exact_until:
            h.step(nsteps - exact_until)
        s = h.get_state()
        checks, disc = h.echeck_stats()
    assert disc == 0 and checks == C * ((sn0 + nsteps) // 10000 - sn0 // 10000)
    assert np.all(s["counters"].sum(axis=1) == nsteps + 1)
    sample = sorted(set([0, 1, 31, 32, 33, 3, 4, 7, 8, 127, 128, 351, 352, 415, 416, 417, 423, 424, 425, 426] +
                        list(range(200, 204))))
    for c in sample:
        dc = dict(d, P=float(P[c]), T=float(T[c]))
        oc = O.Chain(O.config_from_deck(dc, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=id0 + c))
        oc.start()
        if sn0:
            oc.set_step_number(sn0)
        relax0 = oc.relax_calls
        want_log = oracle_accept_log(oc, log_steps)
        oc.run(exact_until - log_steps)
        assert np.array_equal(log[:, c] & 3, want_log), f"chain {c}: accept sequence"
        assert bits_equal(s_exact["r"][c], oc.r), f"chain {c}: positions"
        assert bits_equal(s_exact["l"][c:c + 1], [oc.l]) and np.array_equal(s_exact["counters"][c], oc.counters)
        if arith == "reference":
            assert bits_equal(s_exact["totals"][c], oc.totals), f"chain {c}: totals"
            assert bits_equal(s_exact["accum"][c], oc.accum), f"chain {c}: running sums"
        else:
            assert totals_close(s_exact["totals"][c], oc.totals, 1e-12), f"chain {c}: totals"
            assert np.allclose(s_exact["accum"][c], oc.accum, rtol=1e-11, atol=1e-9)
        if nsteps > exact_until:
            oc.run(nsteps - exact_until)
            assert np.array_equal(s["counters"][c], oc.counters), f"chain {c}: counters after the relaxation"
            assert np.allclose(s["r"][c], oc.r, rtol=0, atol=1e-9) and abs(s["l"][c] - oc.l) < 1e-9
            assert totals_close(s["totals"][c], oc.totals, 1e-11)
        assert oc.relax_calls == relax0 + (1 if start == "from0" else 0)


@pytest.mark.parametrize("engine", ["default", "bond", "solo", "trio", "crew"])
def test_c2_bench_mode_matches_oracle_on_sampled_chains(J, O, engine, monkeypatch):
    """The C2 workload as bench.py runs it (INPUTstd x 4096 chains; the default kernel, bond.cuh and solo.cuh), with the
    step sizes adapted by the host's libm (JMM_ADAPT_HOST: glibc log on both sides, so adaptation cannot hide a
    difference): 72 chains spread over the launch, 2 500 steps = 25 adjustments of each kind, bit-identical to the oracle."""
    if engine != "default":
        for k, v in ENGINES[engine].items():
            monkeypatch.setenv(k, v)
    d = dict(DECKS["std"])
    C, nsteps = 4096, 2500
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_HOST, nchains=C)
    with J.Handle(cfg) as h:
        h.start()
        h.step(nsteps)
        s = h.get_state()
        ms, mv = h.get_step_sizes()
    sample = sorted(set(list(range(0, C, 61)) + [1, 2, 15, 16, 17, 4094, 4095]))
    assert len(sample) >= 64
    for c in sample:
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=c))
        oc.start(); oc.run(nsteps)
        assert bits_equal(s["r"][c], oc.r), f"chain {c}: positions"
        assert bits_equal(s["l"][c:c + 1], [oc.l]) and np.array_equal(s["counters"][c], oc.counters)
        assert bits_equal(s["totals"][c][:2], oc.totals[:2]) and bits_equal(s["accum"][c][:10], oc.accum[:10])
        assert bits_equal([ms[c], mv[c]], list(oc.step_sizes))


@pytest.mark.parametrize("engine", ["default", "trio"])
def test_c2_one_long_launch_matches_oracle(J, O, engine, monkeypatch):
    """One jmm_step of 1 000 000 steps (no adjustments in the deck, so nothing splits the launch): 31 250 hand-overs of each
    ring of solo.cuh's kernels, a million step barriers between the displacement and the volume warp, ECheck on every step.
    4096 chains on the GPU, 17 of them spread over the launch compared with the oracle bit for bit."""
    if engine != "default":
        for k, v in ENGINES[engine].items():
            monkeypatch.setenv(k, v)
    d = dict(DECKS["std"], DADJ=0, VADJ=0)
    C, nsteps = 4096, 1_000_000
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_HOST, nchains=C)
    with J.Handle(cfg) as h:
        if engine == "default":
            assert h.engine == "k_chains_step_crew"
        h.start()
        l0 = h.kernel_launches
        h.step(nsteps)
        assert h.kernel_launches - l0 == 2                         # the kernel + its (empty) repair launch
        s = h.get_state()
        checks, disc = h.echeck_stats()
    assert disc == 0
    for c in list(range(0, C, 293)) + [31, 32, 4095]:
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=c))
        oc.start(); oc.run(nsteps)
        assert bits_equal(s["r"][c], oc.r), f"chain {c}: positions"
        assert bits_equal(s["l"][c:c + 1], [oc.l]) and np.array_equal(s["counters"][c], oc.counters)
        assert bits_equal(s["totals"][c][:2], oc.totals[:2]) and bits_equal(s["accum"][c][:10], oc.accum[:10])


def test_c2_two_ctas_per_sm_match_oracle(J, O):
    """8192 chains of the INPUTstd shape = 256 CTAs of k_chains_step_crew on 148 SMs: two CTAs share an SM, its shared memory and
    its sixteen named barriers.  50 000 steps with host-side adjustments every 100; 20 sampled chains equal the oracle bit for bit."""
    d = dict(DECKS["std"])
    C, nsteps = 8192, 50_000
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_HOST, nchains=C)
    with J.Handle(cfg) as h:
        assert h.engine == "k_chains_step_crew"
        h.start()
        h.step(nsteps)
        s = h.get_state()
        ms, mv = h.get_step_sizes()
    for c in list(range(5, C, 431)) + [8191]:
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=c))
        oc.start(); oc.run(nsteps)
        assert bits_equal(s["r"][c], oc.r), f"chain {c}: positions"
        assert bits_equal(s["l"][c:c + 1], [oc.l]) and np.array_equal(s["counters"][c], oc.counters)
        assert bits_equal(s["totals"][c][:2], oc.totals[:2]) and bits_equal(s["accum"][c][:10], oc.accum[:10])
        assert bits_equal([ms[c], mv[c]], list(oc.step_sizes))


def test_device_adaptation_matches_until_first_adjust_then_statistically(J, O):
    d = dict(DECKS["std"])
    C = 64
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_DEVICE, nchains=C)
    with J.Handle(cfg) as h:
        h.start()
        h.step(99)                                   # DADJ = 100: no adjustment yet
        s = h.get_state()
        for c in (0, 17, 63):
            oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=c))
            oc.start(); oc.run(99)
            assert bits_equal(s["r"][c], oc.r) and np.array_equal(s["counters"][c], oc.counters)
        h.step(20000 - 99)
        ms, mv = h.get_step_sizes()
        s = h.get_state()
    ref = []
    for c in range(8):
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=c))
        oc.start(); oc.run(20000)
        ref.append(oc.step_sizes)
    ref = np.array(ref)
    # CUDA log vs glibc log differ by <= 1 ulp; the adapted step sizes stay statistically the same
    assert abs(np.mean(ms) - np.mean(ref[:, 0])) < 0.25 * np.mean(ref[:, 0])
    assert np.all(s["counters"].sum(axis=1) == 20001)
    assert h is not None


def test_energy_bookkeeping_stays_consistent(J, O):
    """Size-independent property at the bench size: after many steps the incrementally maintained
    totals equal a fresh recompute from the positions (the invariant ECheck guards, :1998)."""
    d = dict(DECKS["std"])
    C = 4096
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_DEVICE, nchains=C)
    with J.Handle(cfg) as h:
        h.start()
        h.step(5000)
        s = h.get_state()
        fresh = h.energy(exact_order=True)
        checks, disc = h.echeck_stats()
    assert checks == C * 5000 and disc == 0
    assert np.all(s["counters"].sum(axis=1) == 5001)                      # step 0 counts as a displacement (:968)
    assert np.max(np.abs(s["totals"][:, 0] - fresh[:, 0])) < 1e-9
    assert np.max(np.abs(s["totals"][:, 1] - fresh[:, 1])) < 1e-9
    assert np.all(np.diff(s["r"], axis=1) > 0), "particles never reorder (HARMONIC returns 1e11 for d<=0)"
    assert np.all(np.abs(s["r"]) <= s["l"][:, None] / 2)                   # hard walls at +-l/2 (:1188)
    acc = s["accum"]
    assert abs(np.mean(acc[:, 2] / 5001) - np.mean(s["l"])) < 0.2 * np.mean(s["l"])   # ensemble <L> ~ current L


def test_banded_acceptance_rules_decide_like_the_reference_expressions(J):
    """pot.cuh settles the Metropolis rule (:1367-1377) and the volume rule (:1666-1672, :2249-2255) by cheap
    approximations with an error band and evaluates the reference expression only inside the band.  40 million random
    cases each, half with the random number within 1e-4 .. 1e-16 (relative) of the exact acceptance probability: the
    decision must be the reference's every time, and the exact path must actually be exercised."""
    bad_m, bad_v, exact_m, exact_v = J.accept_selftest(40_000_000, seed=20261017)
    assert bad_m == 0 and bad_v == 0
    assert exact_m > 1000 and exact_v > 1000


@pytest.mark.parametrize("arith", ["reference", "fast", "fast-g1", "fast-g8", "fast-g32"])
@pytest.mark.parametrize("pot,nbn,cutoff,N,C", [("LJcut", 4, 5.0, 20000, 2), ("LJ", 2, math.inf, 5001, 1),
                                                 ("HARMONIC", 1, math.inf, 8192, 3), ("LJ", 24, math.inf, 6000, 2),
                                                 ("LJcut", 4, 5.0, 64, 7)])
def test_checkerboard_sweeps_match_oracle(J, O, pot, nbn, cutoff, N, C, arith, monkeypatch):
    from jmmonedmc_b200.capi import config
    if arith != "reference" and pot == "HARMONIC":
        pytest.skip("fast arithmetic is an LJ-family path")
    if "-g" in arith:                       # lanes per trial of k_sweep_fast (default: chosen from the trial count)
        arith, g = arith.split("-g")
        monkeypatch.setenv("JMM_SWEEP_G", g)
    P = {"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT, "HARMONIC": J.POT_HARMONIC}[pot]
    seed, id0, T, ms, nhs = 92847, 5, 0.9, 0.12, 3 * (nbn + 1) + 1
    L = N * 1.12
    cfg = config(N=N, pot=P, nbn=nbn, cutoff=cutoff, ensemble=J.ENS_NLT, L=L, T=T, maxStep=ms, seed=seed,
                 nchains=C, chain_id0=id0, mode=J.MODE_CHECKERBOARD,
                 arith=J.ARITH_FAST if arith == "fast" else J.ARITH_REFERENCE)
    with J.Handle(cfg) as h:
        h.start()
        s0 = h.get_state()
        trials = h.sweep(nhs)
        s = h.get_state()
        fresh = h.energy()
    ncol = nbn + 1
    for c in range(C):
        r = ((np.arange(N) + 0.5) / N - 0.5) * L
        assert bits_equal(s0["r"][c], r)
        t0 = O.totals_of(r, nbn, O.POT[pot], cutoff, 1.0, 1, L)
        assert totals_close(s0["totals"][c], exact_totals(r, nbn, pot, cutoff, L), 1e-12)
        tot, nacc, ntry = t0.copy(), 0, 0
        for t in range(nhs):
            col = O.colour_of_step(seed, id0 + c, t, ncol)
            a, dt = O.colour_halfsweep(r, L, nbn, O.POT[pot], cutoff, T, ms, seed, id0 + c, t, ncol, col)
            nacc += a; ntry += len(range(col, N, ncol)); tot += dt
        assert bits_equal(s["r"][c], r), f"chain {c}: positions after {nhs} half-sweeps"
        assert int(s["counters"][c][0]) == nacc and int(s["counters"][c].sum()) == ntry
        assert totals_close(fresh[c], exact_totals(r, nbn, pot, cutoff, L), 1e-12)
        # incrementally maintained totals: thousands of added deltas on both sides, in different orders
        assert totals_close(s["totals"][c], fresh[c], 1e-11) and totals_close(tot, fresh[c], 1e-11)
    assert trials == sum(int(x) for x in s["counters"].sum(axis=1))


@pytest.mark.parametrize("shape", ["K=1,WARPS=4,G=8", "K=1,WARPS=3,G=4", "K=2,WARPS=5,G=2", "K=4,WARPS=2,G=1", "K=1,WARPS=24,G=8",
                                   "K=2,WARPS=7,G=16,NSUB=32"])
@pytest.mark.parametrize("pot,nbn,cutoff,N,C", [("LJcut", 4, 5.0, 60000, 1), ("LJ", 24, math.inf, 40000, 2)])
def test_checkerboard_launch_shapes_match_oracle(J, O, pot, nbn, cutoff, N, C, shape, monkeypatch):
    """k_sweep_fast under forced launch shapes (CTAs per SM, warps per CTA, lanes per trial, half-sweeps per launch):
    many rounds per warp, so that the interior-first order and the deferred neighbour hand-shake (sweep.cuh) are what
    runs.  Positions must be bit-identical to the oracle whatever the shape; a race between warps would show here."""
    from jmmonedmc_b200.capi import config
    for kv in shape.split(","):
        k, v = kv.split("=")
        monkeypatch.setenv("JMM_SWEEP_" + k, v)
    seed, id0, T, ms, nhs = 4711, 3, 0.9, 0.12, 40
    L = N * 1.12
    cfg = config(N=N, pot={"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT}[pot], nbn=nbn, cutoff=cutoff, ensemble=J.ENS_NLT, L=L, T=T,
                 maxStep=ms, seed=seed, nchains=C, chain_id0=id0, mode=J.MODE_CHECKERBOARD, arith=J.ARITH_FAST)
    with J.Handle(cfg) as h:
        h.start()
        trials = h.sweep(nhs)
        s = h.get_state()
        fresh = h.energy()
    ncol = nbn + 1
    for c in range(C):
        r = ((np.arange(N) + 0.5) / N - 0.5) * L
        nacc = ntry = 0
        for t in range(nhs):
            col = O.colour_of_step(seed, id0 + c, t, ncol)
            a, _ = O.colour_halfsweep(r, L, nbn, O.POT[pot], cutoff, T, ms, seed, id0 + c, t, ncol, col)
            nacc += a; ntry += len(range(col, N, ncol))
        assert bits_equal(s["r"][c], r), f"chain {c}: positions after {nhs} half-sweeps with {shape}"
        assert int(s["counters"][c][0]) == nacc and int(s["counters"][c].sum()) == ntry
        assert totals_close(s["totals"][c], fresh[c], 1e-11)
    assert trials == sum(int(x) for x in s["counters"].sum(axis=1))


@pytest.mark.parametrize("pot,nbn,cutoff,N,C,nhs", [("LJcut", 4, 5.0, 1 << 20, 1, 150), ("LJ", 64, math.inf, 1 << 18, 8, 60)])
def test_checkerboard_at_bench_size_matches_oracle(J, O, pot, nbn, cutoff, N, C, nhs):
    """BASELINE.json configs C3 and C5 at their FULL size and with the launch shape bench.py runs (one 24-warp CTA
    per SM, neighbour hand-shakes instead of block barriers, in-kernel reductions, more than one launch): positions
    bit-identical to the oracle, counters equal, totals equal to a fresh evaluation.  compute-sanitizer's racecheck
    cannot follow the flag protocol of sweep.cuh (profiles/r02g_sanitizer.txt); this is the check that it holds."""
    from jmmonedmc_b200.capi import config
    seed, id0, T, ms = 92847, 0, 0.9, 0.12
    L = N * 1.12
    cfg = config(N=N, pot={"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT}[pot], nbn=nbn, cutoff=cutoff, ensemble=J.ENS_NLT, L=L, T=T,
                 maxStep=ms, seed=seed, nchains=C, chain_id0=id0, mode=J.MODE_CHECKERBOARD, arith=J.ARITH_FAST)
    with J.Handle(cfg) as h:
        h.start()
        trials = h.sweep(nhs)
        s = h.get_state()
        fresh = h.energy()
    ncol = nbn + 1
    for c in range(C):
        r = ((np.arange(N) + 0.5) / N - 0.5) * L
        nacc = ntry = 0
        for t in range(nhs):
            col = O.colour_of_step(seed, id0 + c, t, ncol)
            a, _ = O.colour_halfsweep(r, L, nbn, O.POT[pot], cutoff, T, ms, seed, id0 + c, t, ncol, col)
            nacc += a; ntry += len(range(col, N, ncol))
        assert bits_equal(s["r"][c], r), f"chain {c}: positions after {nhs} half-sweeps"
        assert int(s["counters"][c][0]) == nacc and int(s["counters"][c].sum()) == ntry
        assert totals_close(s["totals"][c], fresh[c], 1e-11)
    assert trials == sum(int(x) for x in s["counters"].sum(axis=1))


@pytest.mark.parametrize("name", ["small", "std", "ljcut_nbn"])
@pytest.mark.parametrize("engine", ["prod", "sliced", "generic", "table"])
def test_histograms_match_oracle_integers(J, O, name, engine, monkeypatch):
    """rho(x) and g(x) counts (fgrho/qagrho/ugrho) kept on the device as (count, sum of delta*u) by RED instructions,
    with warp-cooperative refills, must equal, integer for integer, the oracle's add-the-whole-histogram-every-step
    accumulation — at every read-out, for every chain (21 chains: owners and helper lanes share a warp)."""
    if engine == "generic":
        monkeypatch.setenv("JMM_NO_PROD", "1")
    if engine == "sliced":                  # persistent time-sliced launch: a chain changes SM between chunks
        monkeypatch.setenv("JMM_FORCE_SLICE", "1")
        monkeypatch.setenv("JMM_SLICE_CHUNK", "7")
    d = DECKS[name]
    C, id0 = 21, 300
    geo = dict(rhonb=48, rbw=0.5, gns=4, gnb=70, gsw=6.0, gbw=0.25)
    table = engine == "table"
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_TABLE if table else J.MODE_RECOMPUTE,
                               adapt=J.ADAPT_HOST, nchains=C, chain_id0=id0)
    chains = []
    for c in range(C):
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_TABLE if table else O.MODE_RECOMPUTE, chain_id=id0 + c))
        oc.enable_histograms(**geo)
        oc.start()
        chains.append(oc)
    with J.Handle(cfg) as h:
        h.enable_histograms(**geo)
        h.start()
        for nsteps, take_rho, take_g in ((0, True, True), (700, True, False), (500, False, True), (900, True, True)):
            if nsteps:
                h.step(nsteps)
                for oc in chains:
                    oc.run(nsteps)
            rho, g = h.take_histograms(rho=take_rho, g=take_g)
            for c, oc in enumerate(chains):
                wr, wg = oc.take_histograms() if (take_rho and take_g) else (None, None)
                if wr is None:       # take only one of the two on the oracle side as well
                    import ctypes as Ct
                    a = np.zeros(geo["rhonb"], dtype=np.int64); b = np.zeros((geo["gns"], geo["gnb"]), dtype=np.int64)
                    oc.L.jmo_take_histograms(oc.h, a.ctypes.data_as(Ct.POINTER(Ct.c_int64)) if take_rho else None,
                                             b.ctypes.data_as(Ct.POINTER(Ct.c_int64)) if take_g else None)
                    wr, wg = a, b
                if take_rho:
                    assert np.array_equal(rho[c], wr), f"chain {c}: rho counts after {nsteps} more steps"
                if take_g:
                    assert np.array_equal(g[c], wg), f"chain {c}: g counts after {nsteps} more steps"
        s = h.get_state()
    for c, oc in enumerate(chains):
        assert bits_equal(s["r"][c], oc.r)


@pytest.mark.parametrize("case", ["philox-prod", "philox-bond", "taus2-table", "philox-hist", "checkerboard"])
def test_checkpoint_restart_is_an_exact_continuation(J, O, case, tmp_path, monkeypatch):
    """jmm_checkpoint_save / jmm_checkpoint_load (N4): a run stopped after n1 steps and resumed in a NEW handle must
    end in the same bits as the uninterrupted run — positions, box, totals, sums, counters, step sizes, histogram
    counts.  (The reference's RESTART re-seeds and forgets maxStep/dAcc: src/jmmMCState.cpp:572-765.)"""
    from jmmonedmc_b200.capi import config
    ck = tmp_path / "state.jmmckpt"
    if case == "checkerboard":
        N = 6000
        mk = lambda: J.Handle(config(N=N, pot=J.POT_LJCUT, nbn=4, cutoff=5.0, ensemble=J.ENS_NLT, L=N * 1.12, T=0.9, maxStep=0.12,
                                     seed=92847, nchains=2, chain_id0=3, mode=J.MODE_CHECKERBOARD, arith=J.ARITH_FAST))
        run = lambda h, n: h.sweep(n)
        n1, n2 = 7, 9
    else:
        if case == "philox-prod":
            monkeypatch.setenv("JMM_COOP_G", "0")
        deck = DECKS["std"] if case == "philox-bond" else DECKS["small"]
        kw = dict(rng_kind=J.RNG_TAUS2, mode=J.MODE_TABLE) if case == "taus2-table" else dict(rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE)
        mk = lambda: J.Handle(jmm_config_from_deck(J, deck, adapt=J.ADAPT_DEVICE, nchains=45, chain_id0=11, **kw))
        run = lambda h, n: h.step(n)
        n1, n2 = 1234, 2345                      # crosses DADJ/VADJ 1000 and an ECheck on both sides of the cut
    geo = dict(rhonb=48, rbw=0.5, gns=4, gnb=70, gsw=6.0, gbw=0.25)
    hist = case == "philox-hist"

    def state(h):
        s = h.get_state()
        out = [s["r"], s["l"], s["totals"], s["accum"], s["counters"]]
        if case != "checkerboard":
            out += list(h.get_step_sizes())
        if hist:
            out += list(h.take_histograms())
        return out

    with mk() as h:                              # the uninterrupted run
        if hist:
            h.enable_histograms(**geo)
        h.start(); run(h, n1); run(h, n2)
        want = state(h)
    with mk() as h:                              # first leg, checkpoint
        if hist:
            h.enable_histograms(**geo)
        h.start(); run(h, n1)
        h.checkpoint_save(ck)
    with mk() as h:                              # second leg in a fresh handle: no start(), state comes from the file
        if hist:
            h.enable_histograms(**geo)
        h.checkpoint_load(ck)
        assert h.step_number == n1
        run(h, n2)
        got = state(h)
    for a, b in zip(want, got):
        a, b = np.asarray(a), np.asarray(b)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint64) if a.dtype == np.float64 else a,
                                                     b.view(np.uint64) if b.dtype == np.float64 else b)
    # a handle with another configuration refuses the file
    with J.Handle(jmm_config_from_deck(J, DECKS["nlt"], nchains=45)) as h:
        with pytest.raises(J.JmmError):
            h.checkpoint_load(ck)
    with mk() as h:
        with pytest.raises(J.JmmError):
            h.checkpoint_load(tmp_path / "missing.jmmckpt")


@pytest.mark.parametrize("engine", ["prod", "coop", "generic", "lanes8-fast"])
def test_consistent_virial_flag_changes_only_the_virial_bookkeeping(J, O, engine, monkeypatch):
    """JMM_FLAG_CONSISTENT_VIRIAL (include/jmm_gpu.h): the reference's accepted qavLJ move rescales Vir6/Vir12 by
    (l'/l)^-7/-13, adds N T / l and never touches HV (src/jmmMCState.cpp:1675-1686), so its running virial is
    path-dependent.  With the flag the running Vir and HV must equal the configuration sums (a fresh jmm_energy) at any
    time, while positions, box, E, counters and the accept sequence are those of the default (reference-exact) run."""
    env = {"prod": {"JMM_COOP_G": "0"}, "coop": {"JMM_COOP_G": "16"}, "generic": {"JMM_COOP_G": "0", "JMM_NO_PROD": "1"},
           "lanes8": {"JMM_LANES_G": "8"}}[engine.split("-")[0]]
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    fast = engine.endswith("fast")
    d = dict(DECKS["small"], N=24, MAXDV=0.4)
    C, nsteps = 45, 3000
    runs = {}
    for flags in (0, J.FLAG_CONSISTENT_VIRIAL):
        cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_HOST, nchains=C, chain_id0=9,
                                   flags=flags, arith=J.ARITH_FAST if fast else J.ARITH_REFERENCE)
        with J.Handle(cfg) as h:
            h.start()
            log = h.step(nsteps, accept_log=True)
            runs[flags] = (h.get_state(), log, h.energy(exact_order=True))
    (s0, log0, fresh0), (s1, log1, fresh1) = runs[0], runs[J.FLAG_CONSISTENT_VIRIAL]
    assert np.array_equal(log0, log1) and bits_equal(s0["r"], s1["r"]) and bits_equal(s0["l"], s1["l"])
    assert np.array_equal(s0["counters"], s1["counters"]) and np.all(s0["counters"][:, 2] > 10)      # volume moves were accepted
    for k in (0, 2, 4):                                                    # E, E12, E6: the same bookkeeping
        assert bits_equal(s0["totals"][:, k], s1["totals"][:, k])
    assert bits_equal(s0["accum"][:, :7], s1["accum"][:, :7])              # rho, rho^2, L, L^2, E, E^2, LE
    assert totals_close(s1["totals"], fresh1, 1e-11), "with the flag every running total is the configuration sum"
    # without it the reference's bookkeeping is reproduced: Vir carries N T / l and HV lags behind
    assert not totals_close(s0["totals"], fresh0, 1e-6)
    with pytest.raises(J.JmmError):
        J.Handle(jmm_config_from_deck(J, d, rng_kind=J.RNG_TAUS2, mode=J.MODE_TABLE, flags=J.FLAG_CONSISTENT_VIRIAL))


@pytest.mark.parametrize("g", ["2", "4", "8", "16", "team"])
@pytest.mark.parametrize("pot,N", [("LJcut", 78), ("LJ", 73)])
def test_lanes_with_padded_rows_match_oracle(J, O, pot, N, g, monkeypatch):
    """lanes.cuh / team.cuh with the unrolled partner loop (NBN -1, N in (G (NPL-1), G NPL]): rows padded with far-away
    slots (N = 78 and 73 against 80 slots), LJcut's `d <= cutOff` on the SIGNED distance in the reference's orientation
    (src/pot.cpp:53), volume trials through fav (LJcut) or qavLJ (LJ), ECheck every 50 steps, device-side adjustments off
    (DADJ/VADJ beyond the run).  Every chain against the oracle: accept sequence, positions, box and counters
    bit-identical; totals and sums to 1e-12."""
    monkeypatch.setenv("JMM_LANES_G", "8" if g == "team" else g)
    if g == "team":
        monkeypatch.setenv("JMM_TEAM", "1")
    d = dict(N=N, POT=pot, NBN=-1, CUTOFF=2.5 if pot == "LJcut" else math.inf, ENSEMBLE="NPT", P=0.6, T=0.8, MAXSTEP=0.12, MAXDV=1.5,
             ENGCHECK=50, DADJ=10 ** 6, VADJ=10 ** 6, SEED=4242, RELAX=0)
    C, nsteps, id0 = 11, 1200, 77
    cfg = jmm_config_from_deck(J, d, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_DEVICE, nchains=C, chain_id0=id0,
                               arith=J.ARITH_FAST)
    with J.Handle(cfg) as h:
        h.start()
        log = h.step(nsteps, accept_log=True)
        s = h.get_state()
        checks, disc = h.echeck_stats()
    assert disc == 0 and checks == C * (nsteps // 50)
    for c in range(C):
        oc = O.Chain(O.config_from_deck(d, rng_kind=O.RNG_PHILOX, mode=O.MODE_RECOMPUTE, chain_id=id0 + c))
        oc.start()
        want_log = oracle_accept_log(oc, nsteps)
        assert np.array_equal(log[:, c] & 3, want_log), f"chain {c}: accept sequence"
        assert bits_equal(s["r"][c], oc.r), f"chain {c}: positions"
        assert bits_equal(s["l"][c:c + 1], [oc.l]) and np.array_equal(s["counters"][c], oc.counters)
        assert totals_close(s["totals"][c], oc.totals, 1e-12), f"chain {c}: totals"
        assert np.allclose(s["accum"][c], oc.accum, rtol=1e-11, atol=1e-9)

def test_engine_names_follow_the_dispatch(J, monkeypatch):
    """jmm_engine(): the kernel family jmm_step launches for a handle (what bench.py prints beside its roofline)."""
    for k in ("JMM_BOND", "JMM_COOP_G", "JMM_LANES_G", "JMM_TEAM", "JMM_NO_PROD"):
        monkeypatch.delenv(k, raising=False)
    def engine(deck, **kw):
        cfg = jmm_config_from_deck(J, DECKS[deck], rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_HOST, **kw)
        with J.Handle(cfg) as h:
            return h.engine
    assert engine("std", nchains=64) == "k_chains_step_crew"              # HARMONIC, NBN 1, NPT: five warps per 32 chains
    assert engine("std", nchains=12000) == "k_chains_step_bond"           # more CTAs than two per SM hold
    assert engine("small", nchains=64) == "k_chains_step_coop"
    assert engine("small", nchains=64, arith=J.ARITH_FAST).startswith("k_chains_step_lanes")
    assert engine("small", nchains=70000) == "k_chains_step_prod"
    monkeypatch.setenv("JMM_BOND", "4")
    assert engine("std", nchains=64) == "k_chains_step_trio"
