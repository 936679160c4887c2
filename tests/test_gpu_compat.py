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
