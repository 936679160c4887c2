"""Host-side checks of the C ABI that need no GPU: the library builds, loads, exports every symbol
include/jmm_gpu.h declares, parses INPUT decks like readInput, and refuses to compute without CUDA."""
import ctypes as C
import math
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(J):
    L = J.lib()
    names = J.declared_symbols()
    assert len(names) >= 24
    for n in names:
        assert hasattr(L, n), n
    out = subprocess.run(["nm", "-D", "--defined-only", str(J.lib_path())], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(names) <= exported
    assert b"sm_100a" in L.jmm_version()


def test_library_is_built_for_sm_100a(J):
    out = subprocess.run(["cuobjdump", "-lelf", str(J.lib_path())], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.parametrize("deck,want", [
    ("INPUT_smalltest", dict(N=10, pot=0, nbn=-1, ensemble=0, relax=1, P=1.0, T=0.9, maxStep=0.1, maxdl=0.1, eci=1000,
                             mdai=1000, mvai=1000, seed=92847, numsteps=5000000, tpi=1000, cpi=500000)),
    ("INPUTstd", dict(N=10, pot=2, nbn=1, ensemble=0, relax=0, P=0.7, T=0.4, maxStep=0.1, maxdl=1.0, eci=1, mdai=100,
                      mvai=100, seed=125, numsteps=10, tpi=2, cpi=1, n_unknown=1)),
    ("INPUT", dict(N=2000, pot=1, nbn=-1, ensemble=1, L=4000.0, T=0.5, cutoff=5.0, seed=774281, numsteps=1000, eci=1)),
])
def test_read_input_matches_golden_decks(J, deck, want, capfd):
    src = {"INPUT_smalltest": "smalltest_full", "INPUTstd": "inputstd", "INPUT": "input_n2000_40"}[deck]
    path = ROOT / "tests" / "golden" / src / "INPUT"
    cfg, dk = J.read_input(path)
    for k, v in want.items():
        if k in ("numsteps", "tpi", "cpi", "n_unknown"):
            if src == "input_n2000_40" and k in ("numsteps",):
                continue                       # that golden deck shortens NUMSTEPS
            assert getattr(dk, k) == v, k
        else:
            assert getattr(cfg, k) == v, k
    if deck == "INPUT_smalltest":
        assert math.isinf(cfg.cutoff)           # POT LJ: cut-off forced to infinity (src/jmmMCState.cpp:294)
    if deck == "INPUTstd":
        assert "Property command YADA not understood." in capfd.readouterr().out      # src/readInput.cpp:254-256


def test_read_input_errors(J, tmp_path):
    with pytest.raises(J.JmmError) as e:
        J.read_input(tmp_path / "nope")
    assert e.value.status == -3
    p = tmp_path / "INPUT"
    p.write_text("N 5\n\nPOT WEIRD\nENSEMBLE XYZ\nT 1\n")      # blank line + unknown names
    cfg, dk = J.read_input(p)
    assert cfg.pot == J.POT_LJ and cfg.ensemble == J.ENS_NPT and dk.pot_str == b"LJ"


def test_create_validates_and_never_falls_back(J):
    from jmmonedmc_b200.capi import config
    import torch
    bad = [config(N=1, pot=0), config(N=10, pot=7), config(N=10, pot=0, ensemble=5),
           config(N=10, pot=0, rng_kind=J.RNG_RECORDED, nchains=2),
           config(N=100, pot=0, nbn=-1, mode=J.MODE_CHECKERBOARD, ensemble=J.ENS_NLT, L=100.0)]
    want = [-1, -5, -6, -1, -1]
    for c, w in zip(bad, want):
        with pytest.raises(J.JmmError) as e:
            J.Handle(c)
        assert e.value.status == w
    if not torch.cuda.is_available():
        with pytest.raises(J.JmmError) as e:
            J.Handle(config(N=10, pot=0))
        assert e.value.status == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle(J):
    """The oracle is a checker: nothing under jmmonedmc_b200/ may import, include, link or call it."""
    import re
    for p in (ROOT / "jmmonedmc_b200").rglob("*"):
        if p.suffix == ".py":
            text = p.read_text()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), p
            assert "libjmm_oracle" not in text and "jmo_" not in text, p
        elif p.suffix in (".cu", ".cuh", ".cpp", ".h", ".hpp"):
            text = p.read_text()
            assert not re.search(r"#\s*include\s*[<\"][^>\"]*(oracle|jmm_oracle)", text), p
            assert "jmo_" not in text, p
    deps = subprocess.run(["ldd", str(J.lib_path())], capture_output=True, text=True).stdout
    assert "oracle" not in deps
    syms = subprocess.run(["nm", "-D", str(J.lib_path())], capture_output=True, text=True).stdout
    assert "jmo_" not in syms


def test_bench_cpu_decks_are_the_gpu_workloads(J, tmp_path):
    """bench.py times the reference binary on INPUT decks it writes itself (C2 for the headline line, one chain of
    the C4 sweep for --workload c4): parsed by jmm_read_input they must give the configuration the GPU arm runs."""
    import sys
    sys.path.insert(0, str(ROOT))
    import bench
    p = tmp_path / "INPUT"
    p.write_text(bench.deck_text(12345, 125))
    cfg, dk = J.read_input(p)
    c2 = bench.C2
    assert (cfg.N, cfg.pot, cfg.nbn, cfg.ensemble, cfg.relax) == (c2["N"], J.POT_HARMONIC, c2["nbn"], J.ENS_NPT, 0)
    assert (cfg.P, cfg.T, cfg.maxStep, cfg.maxdl, cfg.eci, cfg.mdai, cfg.mvai) == (c2["P"], c2["T"], c2["maxStep"], c2["maxdl"],
                                                                                c2["eci"], c2["mdai"], c2["mvai"])
    assert math.isinf(cfg.cutoff) and dk.numsteps == 12345
    p.write_text(bench.deck_text_c4(777, 9))
    cfg, dk = J.read_input(p)
    c4 = bench.EXTRA["c4"]
    assert (cfg.N, cfg.pot, cfg.nbn, cfg.ensemble, cfg.relax) == (c4["N"], J.POT_LJ, c4["nbn"], J.ENS_NPT, c4["relax"])
    assert (cfg.maxStep, cfg.maxdl, cfg.eci, cfg.mdai, cfg.mvai) == (c4["maxStep"], c4["maxdl"], c4["eci"], c4["mdai"], c4["mvai"])
    assert (cfg.P, cfg.T, cfg.seed, dk.numsteps) == (0.5, 0.5, 9, 777)


def test_bench_reference_arm_prints_the_contract_line(tmp_path):
    """bench.py --impl reference: the compiled reference on the host cores, one JSON line with the keys the driver
    reads (a short sample here; the default is 400 000 steps per process and bench step)."""
    import json, os, sys
    env = dict(os.environ, JMM_BENCH_REF_STEPS="20000")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "MC trial moves/sec" and line["unit"] == "trial moves/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["config"]["workload"].startswith("C2:")


def test_read_input_defaults_and_nbn_zero(J, tmp_path):
    """A deck without an NBN line means "no neighbour limit" (the reference leaves nbn uninitialised; 0 would silently
    exclude every pair), and NBN 0 itself is refused by jmm_create."""
    from jmmonedmc_b200.capi import config
    p = tmp_path / "INPUT"
    p.write_text("N 12\nPOT LJ\nT 0.9\nP 1.0\nNUMSTEPS 10\n")
    cfg, _ = J.read_input(p)
    assert cfg.nbn == -1
    with pytest.raises(J.JmmError) as e:
        J.Handle(config(N=10, pot=0, nbn=0))
    assert e.value.status == -1 and "NBN 0" in str(e.value)


def test_comm_api_without_a_gpu(J):
    """The NCCL side of the ABI: the library loads without NCCL being linked (dlopen at first use); the unique id comes
    from ncclGetUniqueId (no device needed); a communicator needs a CUDA device and says so."""
    import torch
    deps = subprocess.run(["ldd", str(J.lib_path())], capture_output=True, text=True).stdout
    assert "nccl" not in deps
    if J.lib().jmm_nccl_version() == 0:
        with pytest.raises(J.JmmError) as e:
            J.comm_unique_id()
        assert e.value.status == -7
        return
    uid = J.comm_unique_id()
    assert len(uid) == 128 and any(uid) and uid != J.comm_unique_id()
    if not torch.cuda.is_available():
        with pytest.raises(J.JmmError) as e:
            J.Comm(uid, 0, 1, 0)
        assert e.value.status == -2
    with pytest.raises(J.JmmError) as e:
        J.Comm(uid, 3, 2, 0)
    assert e.value.status == -1


def test_batch_driver_refuses_a_restart_deck_before_touching_outputs(J, tmp_path):
    """jmm_run on a deck with a RESTART line: exit 1 with an explanation, existing outputs untouched (the reference
    would reopen config.dat.mcs; silently starting from the lattice and truncating the old files would be worse)."""
    run = ROOT / "jmmonedmc_b200" / "bin" / "jmm_run"
    (tmp_path / "INPUT").write_text("RESTART\nN 10\nPOT LJ\nNBN -1\nT 0.9\nP 1.0\nNUMSTEPS 10\nTPI 1\nCPI 1\n")
    (tmp_path / "thermo.dat.mcs").write_text("previous run\n")
    (tmp_path / "config.dat.mcs").write_text("previous frames\n")
    out = subprocess.run([str(run), "INPUT"], cwd=tmp_path, capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "RESTART" in out.stderr and "--resume" in out.stderr
    assert (tmp_path / "thermo.dat.mcs").read_text() == "previous run\n"
    assert (tmp_path / "config.dat.mcs").read_text() == "previous frames\n"


def test_bench_reference_arm_at_several_gpus_times_the_sweep_deck():
    """bench.py --impl reference --gpus N (N > 1): the GPU arm's headline there is BASELINE config 4 (the 65 536-chain sweep,
    strong scaling), so the reference arm times one chain of that sweep per host core and says so."""
    import json, os, sys
    env = dict(os.environ, JMM_BENCH_REF_STEPS="4000")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "4", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 4 and line["scaling"] == "strong"
    assert line["config"]["workload"].startswith("C4:") and "C4 (one chain)" in line["cpu_baseline"]["sample"]
    assert line["value"] > 0 and line["gpu_launches"] == 0
