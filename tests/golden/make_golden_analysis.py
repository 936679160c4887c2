#!/usr/bin/env python
"""Golden outputs of the reference's OWN analysis scripts (scripts/Analyze_Mean.py, Analyze_A_SD.py,
Plot_AutoCorrelation.py), for pinning jmmonedmc_b200/analysis.py: tests/golden/analysis_scripts.

The scripts are python-2 programs with matplotlib plots, hard-coded data directories and thermo column names the
current writer no longer emits.  They are run here UNMODIFIED IN THEIR ARITHMETIC, read from /root/reference/scripts at
run time (nothing of them is copied into the repository), through a small compatibility shim:
  * tabs expanded to 8 columns (python 2's rule for their mixed indentation), `np.float` = float, `xrange` = range;
  * python 2's integer `/` restored where the scripts divide step counts to get slice indices
    (`/printInterval`, `)/blockSize`, `B.shape[1]/...`), as regex patches listed next to each run below;
  * matplotlib replaced by a stub whose scatter()/plot() record the plotted arrays (the response functions of
    Analyze_Mean.py are only ever plotted, never written);
  * run in a scratch tree shaped like the hard-coded paths (`../data/LJ/m-1/Longest/P0.1_T0.8_<job>/thermo.dat`, ...),
    on a synthetic thermo file written twice from the same numbers: with the scripts' column names
    (Step Energy Energy2 l l2 Virial Virial2 lE) for them, with the writer's 13-column header for analysis.py.

    python tests/golden/make_golden_analysis.py        # needs /root/reference
"""
import contextlib
import io
import json
import os
import re
import shutil
import sys
import tempfile
import types
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/scripts")
OUT = HERE / "analysis_scripts"


def run_py2(name, cwd, subs=(), run=None):
    """exec() a reference script under python 3 with the shim; returns (globals, recorded plot calls, stdout)."""
    src = (REF / name).read_text().expandtabs(8)
    for pat, rep in subs:
        assert re.search(pat, src), (name, pat)
        src = re.sub(pat, rep, src)
    calls = []

    class _Any:
        def __getattr__(self, k): return lambda *a, **k2: _Any()
        def __call__(self, *a, **k): return _Any()

    def rec(kind):
        def f(*a, **k):
            calls.append((kind, [np.asarray(x, dtype=float).ravel().tolist() for x in a[:2]]))
            return _Any()
        return f
    plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = lambda k: (lambda *a, **k2: _Any())
    plt.scatter, plt.plot = rec("scatter"), rec("plot")
    mpl = types.ModuleType("matplotlib"); mpl.pyplot = plt; mpl.use = lambda *a, **k: None
    saved = {k: sys.modules.get(k) for k in ("matplotlib", "matplotlib.pyplot")}
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if not hasattr(np, "float"):
        np.float = float
    g = {"__name__": "__main__" if run is None else "script", "xrange": range}
    old, out = os.getcwd(), io.StringIO()
    os.chdir(cwd)
    try:
        with contextlib.redirect_stdout(out):
            exec(compile(src, name, "exec"), g)
            ret = run(g) if run else None
    finally:
        os.chdir(old)
        for k, v in saved.items():
            if v is None: sys.modules.pop(k, None)
            else: sys.modules[k] = v
    return g, calls, out.getvalue(), ret


def synthetic_thermo(rows, interval, seed):
    """A correlated random walk in (E, L) with the 13 thermo columns derived from it (block means of a run would look
    like this); the values only have to be the same numbers in both file formats."""
    rng = np.random.default_rng(seed)
    step = np.arange(rows) * interval
    E = -1500 + 0.3 * np.cumsum(rng.normal(0, 1, rows)); L = 2300 + 0.2 * np.cumsum(rng.normal(0, 1, rows))
    E2 = E * E + 30 + rng.normal(0, 1, rows); L2 = L * L + 9 + rng.normal(0, 0.5, rows); LE = L * E - 5 + rng.normal(0, 2, rows)
    rho = 2000 / L; vir = 0.01 * L + rng.normal(0, 0.01, rows)
    new = np.stack([step, E, E2, L, L2, LE, rho, rho * rho, vir, vir * vir, E * vir, 7 * vir, 49 * vir * vir], axis=1)
    return new


def write_both(new, path_old, path_new):
    fmt = lambda x: "%.10G" % x
    with open(path_new, "w") as f:                                    # the writer's header, src/jmmMCState.cpp:566-568
        f.write("Step    Econf           Econf2          L       L2      LEconf          rho             rho2            Virial         Virial2         EconfVir        HV              HV2 \n")
        for r in new:
            f.write("%d\t" % r[0] + "\t".join(fmt(x) for x in r[1:]) + "\n")
    back = np.loadtxt(path_new, skiprows=1)                            # the numbers as the text holds them
    with open(path_old, "w") as f:                                    # the names the scripts read
        f.write("Step Energy Energy2 l l2 Virial Virial2 lE\n")
        for r in back:
            f.write("%d %s %s %s %s %s %s %s\n" % (r[0], fmt(r[1]), fmt(r[2]), fmt(r[3]), fmt(r[4]), fmt(r[8]), fmt(r[9]), fmt(r[5])))


def main():
    OUT.mkdir(exist_ok=True)
    gold = {}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        (tmp / "scripts").mkdir()
        # ---- Analyze_Mean.py: P0.1*T0.8* under ../data/LJ/m-1/Longest/, blockSize 1e6, printInterval 10000 (hard-coded)
        d1 = tmp / "data/LJ/m-1/Longest/P0.1_T0.8_77"; d1.mkdir(parents=True)
        new = synthetic_thermo(801, 10000, 7)
        write_both(new, d1 / "thermo.dat", OUT / "mean_thermo.dat.mcs")
        (d1.parent / "EqSteps.dat").write_text("0.1 0.8 1000000 50000 2000000\n1.0 1.0 0 50000 0\n")
        g, calls, _, _ = run_py2("Analyze_Mean.py", tmp / "scripts", [(r"/printInterval", "//printInterval"), (r"\)/blockSize", ")//blockSize")])
        g["summaryFile"].close()
        gold["Analyze_Mean"] = {"P": 0.1, "T": 0.8, "eq_steps": 1000000, "le_start_steps": 2000000, "block_size": 1000000, "interval": 10000,
                                "LJ_Means.dat": (tmp / "scripts/LJ_Means.dat").read_text(),
                                "plots": dict(zip(["E_times_epsilon", "L", "E2", "L2", "LE", "cp", "betaT", "betaS", "alphaP", "gammaV", "muJT"],
                                                  [{"x": c[1][0], "y": c[1][1]} for c in calls])),
                                "LEstartBlock": int(g["LEstartBlock"]), "epsilon": float(g["epsilon"])}
        assert len(calls) == 11
        # ---- Analyze_A_SD.py: P1.0*T1.0* under ../data/LJ/m-1/N2000/Longest/, blockSizeMin 1e5; once with an uncorrelated
        #      block size named in EqSteps.dat (first row = the Summary_SD_tmp.txt row), once without
        for tag, eq_line in (("with_uncorrelated_block", "1.0 1.0 300000 250000 800000\n"), ("scan_only", "1.0 1.0 300000 0 0\n")):
            shutil.rmtree(tmp / "data/LJ/m-1/N2000", ignore_errors=True)
            d2 = tmp / "data/LJ/m-1/N2000/Longest/P1.0_T1.0_5"; d2.mkdir(parents=True)
            new = synthetic_thermo(400, 5000, 11)
            write_both(new, d2 / "thermo.dat", OUT / "sd_thermo.dat.mcs")
            (d2.parent / "EqSteps.dat").write_text(eq_line + "0.5 0.5 0 0 0\n")      # (two rows: genfromtxt must return a 2-D table)
            (tmp / "scripts/Summary_SD_tmp.txt").unlink(missing_ok=True)
            g, _, _, _ = run_py2("Analyze_A_SD.py", tmp / "scripts", [(r"\(blockSize / printInterval\)", "(blockSize // printInterval)")])
            g["summaryFile"].close()
            f = eq_line.split()
            gold["Analyze_A_SD:" + tag] = {"eq_steps": int(f[2]), "uncorrelated_block_size": int(f[3]), "le_start_steps": int(f[4]),
                                           "DataBlockingResults.dat": (d2 / "DataBlockingResults.dat").read_text(),
                                           "Summary_SD_tmp.txt": (tmp / "scripts/Summary_SD_tmp.txt").read_text()}
        # ---- Plot_AutoCorrelation.py: the two functions on a fixed series
        x = np.cumsum(np.random.default_rng(3).normal(0, 1, 240)) * 0.1 + np.random.default_rng(4).normal(0, 1, 240)
        # (the debug print at :28 indexes c[jj] with jj = numBlocks-1, out of bounds at the first block size for any input: as
        #  shipped the function raises IndexError; the line is dropped, the arithmetic is untouched)
        subs = [(r"B\.shape\[1\]/2\+1", "B.shape[1]//2+1"), (r"B\.shape\[1\]/blockSize", "B.shape[1]//blockSize"),
                (r"\n +print\('c\[' \+ str\(jj\)[^\n]*", "")]
        _, _, _, ret = run_py2("Plot_AutoCorrelation.py", tmp / "scripts", subs,
                               run=lambda g: (g["Plot_AutoCorrelation"](x.reshape(1, -1))[0], g["unMeaned"](x.reshape(1, -1))[0]))
        gold["Plot_AutoCorrelation"] = {"x": x.tolist(), "blocked": np.asarray(ret[0]).tolist(), "lag": np.asarray(ret[1]).tolist()}
    (OUT / "golden.json").write_text(json.dumps(gold) + "\n")
    print("analysis_scripts:", sorted(gold), "->", OUT)


if __name__ == "__main__":
    main()
