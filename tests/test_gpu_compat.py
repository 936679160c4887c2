"""Drop-in proof: the reference's UNMODIFIED Main.cpp + readInput.cpp, linked against the
jmmMCState.h-compatible shim (csrc/host/jmm_mcstate_compat.cpp) and libjmmgpu.so, must write the same
thermo.dat.mcs and config.dat.mcs as the reference engine — byte for byte (lock-step mode)."""
import hashlib
import os
import subprocess
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "oracle" / "_ref" / "jmmOneDMC_gpu"


def _run(deck_text, tmp_path, env_extra=None):
    (tmp_path / "INPUT").write_text(deck_text)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    env.update(env_extra or {})
    out = subprocess.run([str(BIN)], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    return out.stdout


@pytest.mark.parametrize("name", ["smalltest_12", "smalltest_2000", "smalltest_20000", "input_n2000_40", "inputstd"])
def test_reference_main_on_gpu_engine_writes_identical_files(gold, tmp_path, name):
    if not BIN.exists():
        pytest.skip("oracle/_ref/jmmOneDMC_gpu not built (needs /root/reference at build time)")
    g = gold(name)
    stdout = _run(g["deck_text"], tmp_path)
    s = g["summary"]
    thermo = (tmp_path / "thermo.dat.mcs").read_bytes()
    config = (tmp_path / "config.dat.mcs").read_bytes()
    assert hashlib.md5(config).hexdigest() == s["config_md5"]
    if name == "inputstd":       # HV columns: uninitialised stack in the reference (src/pot.cpp:116-131)
        strip = lambda b: [l.split(b"\t")[:11] for l in b.splitlines()[1:]]
        assert strip(thermo) == strip((g["dir"] / "thermo.dat.mcs").read_bytes())
        assert "Property command YADA not understood." in stdout          # the reference's own parser ran
    else:
        assert hashlib.md5(thermo).hexdigest() == s["thermo_md5"]
    assert f"\nE = {s['final_E_printed']}\n" in stdout
    c = s["counters"]
    assert f"{c[0]}/{c[1]}          {c[2]}/{c[3]}" in stdout
    assert "PROGRAM COMPLETED SUCCESSFULLY!" in stdout


@pytest.mark.parametrize("name", ["smalltest_hist", "inputstd_hist"])
def test_reference_main_on_gpu_engine_histogram_files(gold, tmp_path, name):
    """rho.dat.mcs and g<k>.dat.mcs written by the reference's Main.cpp driving the GPU engine: byte-identical."""
    if not BIN.exists():
        pytest.skip("oracle/_ref/jmmOneDMC_gpu not built")
    g = gold(name)
    _run(g["deck_text"], tmp_path)
    files = ["rho.dat.mcs", "config.dat.mcs"] + sorted(p.name for p in g["dir"].glob("g*.dat.mcs"))
    if name != "inputstd_hist":
        files.append("thermo.dat.mcs")
    for f in files:
        assert (tmp_path / f).read_bytes() == (g["dir"] / f).read_bytes(), f


def test_reference_main_on_gpu_engine_production_mode_runs(gold, tmp_path):
    if not BIN.exists():
        pytest.skip("oracle/_ref/jmmOneDMC_gpu not built")
    stdout = _run(gold("smalltest_2000")["deck_text"], tmp_path, {"JMM_COMPAT_PRODUCTION": "1"})
    assert "PROGRAM COMPLETED SUCCESSFULLY!" in stdout
    rows = (tmp_path / "thermo.dat.mcs").read_text().splitlines()
    assert len(rows) == 1 + 21 and rows[1].split("\t")[0] == "0"
    # step-0 row is deterministic (no random numbers yet): identical to the reference's
    assert rows[1] == (gold("smalltest_2000")["dir"] / "thermo.dat.mcs").read_text().splitlines()[1]


RUN = ROOT / "jmmonedmc_b200" / "bin" / "jmm_run"


@pytest.mark.parametrize("name", ["smalltest_2000", "smalltest_20000", "input_n2000_40"])
def test_batch_driver_lockstep_files_identical(gold, tmp_path, name):
    """jmm_run --lockstep (our own Main: batched launches between print boundaries) on the reference decks."""
    g = gold(name)
    (tmp_path / "INPUT").write_text(g["deck_text"])
    out = subprocess.run([str(RUN), "INPUT", "--lockstep"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    s = g["summary"]
    assert hashlib.md5((tmp_path / "thermo.dat.mcs").read_bytes()).hexdigest() == s["thermo_md5"]
    assert hashlib.md5((tmp_path / "config.dat.mcs").read_bytes()).hexdigest() == s["config_md5"]
    assert f"\nE = {s['final_E_printed']}\n" in out.stdout
    c = s["counters"]
    assert f"{c[0]}/{c[1]}          {c[2]}/{c[3]}" in out.stdout


@pytest.mark.parametrize("name", ["smalltest_hist"])
def test_batch_driver_lockstep_histogram_files(gold, tmp_path, name):
    g = gold(name)
    (tmp_path / "INPUT").write_text(g["deck_text"])
    out = subprocess.run([str(RUN), "INPUT", "--lockstep"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    for f in ["rho.dat.mcs", "thermo.dat.mcs", "config.dat.mcs"] + sorted(p.name for p in g["dir"].glob("g*.dat.mcs")):
        assert (tmp_path / f).read_bytes() == (g["dir"] / f).read_bytes(), f


def test_batch_driver_state_point_sweep(gold, tmp_path):
    """A P x T grid in one process (what scripts/RunJobs.bash does with one LSF job per point)."""
    deck = gold("smalltest_2000")["deck_text"].replace("NUMSTEPS   2000", "NUMSTEPS   4000")
    (tmp_path / "INPUT").write_text(deck)
    out = subprocess.run([str(RUN), "INPUT", "--chains", "24", "--sweep-p", "0.5", "1.5", "3", "--sweep-t", "0.6", "1.2", "4"],
                         cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    rows = [l.split("\t") for l in (tmp_path / "Summary.dat").read_text().splitlines()]
    assert len(rows) == 1 + 24 and rows[0][:5] == ["chain", "P", "T", "N", "samples"]
    P = sorted({float(r[1]) for r in rows[1:]}); T = sorted({float(r[2]) for r in rows[1:]})
    assert P == [0.5, 1.0, 1.5] and len(T) == 4 and abs(T[-1] - 1.2) < 1e-12
    col = rows[0].index("L")
    L = {(float(r[1]), float(r[2])): [] for r in rows[1:]}
    for r in rows[1:]:
        L[(float(r[1]), float(r[2]))].append(float(r[col]))
    # physics sanity: at fixed T the mean box length shrinks as the pressure rises
    for t in T:
        assert sum(L[(0.5, t)]) / 2 > sum(L[(1.5, t)]) / 2
    thermo = (tmp_path / "thermo_chains.dat.mcs").read_text().splitlines()
    assert len(thermo) == 1 + 24 * (1 + 4000 // 100)            # TPI 100: step 0 + 40 block rows per chain
    assert "0 discrepancies" in out.stdout


def test_reference_main_on_gpu_engine_with_an_openmp_team(gold, tmp_path):
    """The reference's Main.cpp runs printCoords / printRho / updateThermo+printThermo in `omp sections` on different
    threads (src/Main.cpp:77-106).  With a team of 4 the shim must still write the reference's bytes: its GPU-touching
    functions are serialised by one critical section (jmm_mcstate_compat.cpp)."""
    if not BIN.exists():
        pytest.skip("oracle/_ref/jmmOneDMC_gpu not built")
    g = gold("smalltest_2000")
    stdout = _run(g["deck_text"], tmp_path, {"OMP_NUM_THREADS": "4"})
    s = g["summary"]
    assert hashlib.md5((tmp_path / "thermo.dat.mcs").read_bytes()).hexdigest() == s["thermo_md5"]
    assert hashlib.md5((tmp_path / "config.dat.mcs").read_bytes()).hexdigest() == s["config_md5"]
    assert f"\nE = {s['final_E_printed']}\n" in stdout


@pytest.mark.slow
def test_batch_driver_lockstep_full_smalltest_as_shipped(gold, tmp_path):
    """BASELINE.json configs[0]: test/INPUT_smalltest VERBATIM — 5 000 000 steps, 5000 adjustments of each kind, 100
    relaxations, 5000 energy checks — in lock-step on the GPU (device taus2 = the reference's own stream, rij-table
    arithmetic).  thermo.dat.mcs and config.dat.mcs must be the reference's bytes (md5 of the compiled reference's
    files, tests/golden/smalltest_full), the counters 2644332/1900863 116301/338505, final E -3.2554505, and the last
    frame the reference's final positions."""
    g = gold("smalltest_full")
    (tmp_path / "INPUT").write_text(g["deck_text"])
    out = subprocess.run([str(RUN), "INPUT", "--lockstep"], cwd=tmp_path, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    s = g["summary"]
    assert hashlib.md5((tmp_path / "thermo.dat.mcs").read_bytes()).hexdigest() == s["thermo_md5"] == "66eb0e4ba0e28398cb24a8ca11577339"
    assert hashlib.md5((tmp_path / "config.dat.mcs").read_bytes()).hexdigest() == s["config_md5"]
    assert f"\nE = {s['final_E_printed']}\n" in out.stdout
    assert "2644332/1900863          116301/338505" in out.stdout
    frame = (tmp_path / "config.dat.mcs").read_text().splitlines()[-10:]
    assert [float(l.split()[-1]) for l in frame] == s["last_frame_r"]
    assert "5000 checks, 0 discrepancies" in out.stdout
    tail = (g["dir"] / "thermo.tail.mcs").read_text()
    assert (tmp_path / "thermo.dat.mcs").read_text().endswith(tail)


def _sweep_deck(gold, numsteps):
    return gold("smalltest_2000")["deck_text"].replace("NUMSTEPS   2000", f"NUMSTEPS   {numsteps}")


def test_batch_driver_checkpoint_and_resume_append_exactly(gold, tmp_path):
    """jmm_run --checkpoint / --resume: 2000 steps + a resumed 2000 must leave the thermo file of the uninterrupted
    4000-step run (rows appended, nothing truncated); a RESTART deck is refused without touching the outputs."""
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(); b.mkdir()
    args = ["--chains", "12", "--sweep-p", "0.5", "1.5", "3", "--sweep-t", "0.6", "1.2", "4"]
    (a / "INPUT").write_text(_sweep_deck(gold, 4000))
    r = subprocess.run([str(RUN), "INPUT"] + args, cwd=a, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    (b / "INPUT").write_text(_sweep_deck(gold, 2000))
    r = subprocess.run([str(RUN), "INPUT", "--checkpoint", "state.ckpt"] + args, cwd=b, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and (b / "state.ckpt").exists(), r.stdout[-1500:] + r.stderr[-1500:]
    (b / "INPUT").write_text(_sweep_deck(gold, 4000))
    r = subprocess.run([str(RUN), "INPUT", "--resume", "state.ckpt"] + args, cwd=b, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Resumed from state.ckpt at step 2000" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
    assert (b / "thermo_chains.dat.mcs").read_bytes() == (a / "thermo_chains.dat.mcs").read_bytes()
    before = (b / "thermo_chains.dat.mcs").read_bytes()
    (b / "INPUT").write_text("RESTART\n" + _sweep_deck(gold, 4000))
    r = subprocess.run([str(RUN), "INPUT"] + args, cwd=b, capture_output=True, text=True, timeout=600)
    assert r.returncode == 1 and "RESTART" in r.stderr
    assert (b / "thermo_chains.dat.mcs").read_bytes() == before


def test_batch_driver_runjobs_directory_layout(gold, tmp_path):
    """--layout runjobs: the tree scripts/RunJobs.bash:27 makes with one LSF job per state point —
    data/<POT>/m<NBN>/N<N>/P<P>_T<T>/{INPUT, thermo.dat.mcs} — each thermo file in the reference's 13-column format
    with the rows thermo_chains.dat.mcs holds for that chain."""
    (tmp_path / "INPUT").write_text(_sweep_deck(gold, 1000))
    r = subprocess.run([str(RUN), "INPUT", "--chains", "6", "--sweep-p", "0.5", "1.0", "2", "--sweep-t", "0.6", "1.2", "3",
                        "--layout", "runjobs"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    base = tmp_path / "data" / "LJ" / "m-1" / "N10"
    dirs = sorted(p.name for p in base.iterdir())
    assert dirs == ["P0.5_T0.6", "P0.5_T0.9", "P0.5_T1.2", "P1_T0.6", "P1_T0.9", "P1_T1.2"]
    chains = (tmp_path / "thermo_chains.dat.mcs").read_text().splitlines()
    header = chains[0].split("\t", 1)[1]
    for c, d in enumerate(dirs):
        deck = (base / d / "INPUT").read_text()
        p, t = d[1:].split("_T")
        assert f"P          {p}\n" in deck and f"T          {t}\n" in deck and "NUMSTEPS   1000" in deck
        rows = (base / d / "thermo.dat.mcs").read_text().splitlines()
        assert rows[0] == header
        want = [l.split("\t", 1)[1] for l in chains[1:] if l.split("\t", 1)[0] == str(c)]
        assert rows[1:] == want and len(want) == 11


def test_summary_records_and_single_rank_allgather(J):
    """jmm_summaries packs the per-chain records on the device; jmm_allgather_summaries with a one-rank NCCL
    communicator returns the same table (the N > 1 case: tests/test_gpu_multi.py, needs two GPUs)."""
    import numpy as np
    from jmmonedmc_b200.capi import config
    C, id0 = 37, 100
    cfg = config(N=10, pot=J.POT_LJ, nbn=-1, ensemble=J.ENS_NPT, P=1.0, T=0.9, maxStep=0.1, maxdl=0.1, eci=1000, mdai=1000,
                 mvai=1000, seed=92847, nchains=C, chain_id0=id0)
    with J.Handle(cfg) as h:
        h.set_state(P=np.linspace(0.5, 1.5, C), T=np.linspace(0.6, 1.2, C))
        h.start(); h.step(777)
        s = h.get_state()
        rec = h.summaries()
        assert np.array_equal(rec[:, 0], id0 + np.arange(C)) and np.all(rec[:, 3] == 10) and np.all(rec[:, 4] == 778)
        assert np.array_equal(rec[:, 1], np.linspace(0.5, 1.5, C)) and np.array_equal(rec[:, 2], np.linspace(0.6, 1.2, C))
        assert np.array_equal(rec[:, 5:17], s["accum"]) and np.array_equal(rec[:, 17:21], s["counters"].astype(float))
        assert np.array_equal(rec[:, 21], s["l"]) and np.array_equal(rec[:, 22:24], s["totals"][:, :2])
        if J.lib().jmm_nccl_version() == 0:
            pytest.skip("no NCCL library on this box")
        with pytest.raises(J.JmmError):          # chain ids 100..136 do not fit a table of 37 chains starting at 0
            with J.Comm(J.comm_unique_id(), 0, 1, 0) as comm:
                h.allgather_summaries(comm, C)
    cfg0 = cfg.copy(chain_id0=0)
    with J.Handle(cfg0) as h:
        h.start(); h.step(100)
        with J.Comm(J.comm_unique_id(), 0, 1, 0) as comm:
            full = h.allgather_summaries(comm, C)
        assert np.array_equal(full, h.summaries())
