"""Post-processing of thermo.dat.mcs: the tables the reference's scripts/ build.

The reference scripts are python-2 programs with hard-coded paths that read column names the current writer no
longer emits (`Energy`, `l`, `lE` vs the header `Econf`, `L`, `LEconf` of src/jmmMCState.cpp:566-568), so they
cannot run on today's output.  This module restates their arithmetic on the columns as written today:

  run_means            scripts/Analyze_Mean.py:119-131     whole-run means after equilibration -> LJ_Means.dat row
  block_means          scripts/Analyze_Mean.py:143-160     means over consecutive blocks of `block_size` MC steps
  response_functions   scripts/Analyze_Mean.py:162-176     cp, betaT, alphaP, gammaV, betaS, muJT from block means
  data_blocking        scripts/Analyze_A_SD.py:86-213      standard error of the block means vs block size
                                                           -> DataBlockingResults.dat, Summary_SD_tmp.txt
  block_stdevs         scripts/Analyze_Fluctuations.py     std of block means, block size growing by 1.3x
  blocked_autocorrelation / lag_autocorrelation            scripts/Plot_AutoCorrelation.py:15-48
  sort_summary         scripts/Sort_Summary.bash           sort -k1,1g -k2,2g | uniq

Pinned against the scripts themselves: tests/golden/make_golden_analysis.py runs Analyze_Mean.py, Analyze_A_SD.py and
Plot_AutoCorrelation.py from /root/reference/scripts under a python-2 compatibility shim on a synthetic thermo file, and
tests/test_analysis_cpu.py compares this module with what they computed.

Host-side numpy only: none of this is on the Monte-Carlo hot path and nothing here touches the GPU.
`python -m jmmonedmc_b200.analysis DIR` analyses every `P<P>_T<T>*/thermo.dat.mcs` under DIR (the layout of
scripts/RunJobs.bash:27) and writes the tables into DIR.
"""
from __future__ import annotations

import glob
import os
import re
import sys
from pathlib import Path

import numpy as np

# header of thermo.dat.mcs (src/jmmMCState.cpp:566-568) and the names the scripts used for the same columns
COLUMNS = ("Step", "Econf", "Econf2", "L", "L2", "LEconf", "rho", "rho2", "Virial", "Virial2", "EconfVir", "HV", "HV2")
ALIASES = {"Energy": "Econf", "Energy2": "Econf2", "l": "L", "l2": "L2", "lE": "LEconf", "lEnergy": "LEconf"}


def read_thermo(path) -> np.ndarray:
    """thermo.dat.mcs -> structured array with the writer's column names (np.genfromtxt(names=True) of the scripts)."""
    with open(path) as f:
        names = f.readline().split()
    if tuple(names) != COLUMNS:
        raise ValueError(f"{path}: unexpected thermo header {names}")
    data = np.loadtxt(path, skiprows=1, ndmin=2)
    out = np.empty(data.shape[0], dtype=[(n, np.float64) for n in COLUMNS])
    for k, n in enumerate(COLUMNS):
        out[n] = data[:, k]
    return out


def col(thermo: np.ndarray, name: str) -> np.ndarray:
    return thermo[ALIASES.get(name, name)]


def print_interval(thermo: np.ndarray) -> int:
    """Analyze_A_SD.py: printInterval = Step[1] - Step[0]."""
    return int(thermo["Step"][1] - thermo["Step"][0])


def check_spacing(thermo: np.ndarray, interval: int) -> None:
    """The scripts abort on "INCONSISTENT STEP SPACING IN THERMO FILE" (Analyze_Mean.py:134-139)."""
    d = np.diff(thermo["Step"])
    bad = np.nonzero(d != interval)[0]
    if bad.size:
        raise ValueError(f"inconsistent step spacing in thermo file at row {int(bad[0])}: {d[bad[0]]} != {interval}")


def after_equilibration(thermo: np.ndarray, eq_steps: int, interval: int) -> np.ndarray:
    """thermo_data_raw[eqSteps/printInterval+1:] (Analyze_Mean.py:114): drops the step-0 row and the equilibration rows."""
    return thermo[int(eq_steps // interval) + 1:]


def run_means(thermo: np.ndarray, eq_steps: int = 0, le_start_steps: int | None = None, interval: int | None = None) -> dict:
    """Analyze_Mean.py:119-131: means of the per-interval averages; <LE> may start later (column 5 of EqSteps.dat)."""
    interval = interval or print_interval(thermo)
    t = after_equilibration(thermo, eq_steps, interval)
    le0 = 0 if le_start_steps is None else max(0, int((le_start_steps - eq_steps) // interval))
    return {"Energy": float(np.mean(t["Econf"])), "Energy2": float(np.mean(t["Econf2"])), "Length": float(np.mean(t["L"])),
            "Length2": float(np.mean(t["L2"])), "LengthEnergy": float(np.mean(t["LEconf"][le0:])), "rows": int(t.size)}


BLOCK_COLUMNS = ("Step", "Econf", "Econf2", "L", "L2", "Virial", "Virial2", "LEconf")


def block_means(thermo: np.ndarray, block_size: int, interval: int, columns=BLOCK_COLUMNS) -> np.ndarray:
    """[numBlocks][len(columns)] means over rows [b*block_size/interval, (b+1)*block_size/interval)
    (Analyze_Mean.py:143-160; integer division exactly as the script's python-2 `/`)."""
    nb = int(np.floor(thermo.size / (block_size / interval)))
    out = np.empty((nb, len(columns)))
    for b in range(nb):
        lo, hi = int(b * block_size / interval), int((b + 1) * block_size / interval)
        for k, n in enumerate(columns):
            out[b, k] = np.mean(thermo[n][lo:hi])
    return out


def response_functions(P: float, T: float, eaE, eaESq, eaL, eaLSq, eaLE) -> dict:
    """Configurational response functions from (block) averages, Analyze_Mean.py:162-176, term for term."""
    eaE, eaESq, eaL, eaLSq, eaLE = (np.asarray(x, dtype=np.float64) for x in (eaE, eaESq, eaL, eaLSq, eaLE))
    cp = eaESq - eaE * eaE + 2.0 * P * (eaLE - eaL * eaE) + P * P * (eaLSq - eaL * eaL)
    cp = cp / T / T
    betaT = 1.0 / eaL / T * (eaLSq - eaL * eaL)
    alphaP = 1 / T / T / eaL * ((eaLE - eaL * eaE) + P * (eaLSq - eaL * eaL))
    gammaV = alphaP / betaT
    betaS = betaT - alphaP * alphaP * T * eaL / cp
    muJT = eaL / cp * (alphaP * T - 1.0)
    return {"cp": cp, "betaT": betaT, "alphaP": alphaP, "gammaV": gammaV, "betaS": betaS, "muJT": muJT}


def data_blocking(thermo: np.ndarray, eq_steps: int = 0, uncorrelated_block_size: int = 0, le_start_steps: int = 0,
                  block_size_min: int = 100000, interval: int | None = None):
    """Analyze_A_SD.py:86-200.  Returns (stderrs, summary_row): stderrs[k] = blockSize, numBlocks, SEM of the block
    means of E, E^2, L, L^2, LE (np.std / sqrt(numBlocks)) for blockSize = block_size_min, 2*block_size_min, ...
    up to half the run; when EqSteps.dat names an uncorrelated block size that size comes first and is also
    the Summary_SD_tmp.txt row."""
    interval = interval or print_interval(thermo)
    le_start_steps = max(le_start_steps, eq_steps)
    le = thermo[int(le_start_steps / interval) + 1:]
    t = thermo[int(eq_steps / interval) + 1:]
    check_spacing(t, interval)
    rows, summary = [], None
    counter, bs = (1, block_size_min) if uncorrelated_block_size < 1 else (0, uncorrelated_block_size)
    bs_max = int(max(t.size, le.size) * interval / 2)
    while bs <= bs_max:
        bs = int(bs)
        nb = int(np.floor(t.size / (bs / interval)))
        nble = int(np.floor(le.size / (bs / interval)))
        bm = block_means(t, bs, interval, ("Step", "Econf", "Econf2", "L", "L2", "Virial", "Virial2"))
        ble = block_means(le, bs, interval, ("Step", "LEconf"))
        se = [np.std(bm[:, k]) / np.sqrt(nb) for k in (1, 2, 3, 4)] + [np.std(ble[:, 1]) / np.sqrt(nble)]
        rows.append([bs, nb] + [float(x) for x in se])
        if counter > 0:
            bs = bs + block_size_min
        else:
            summary = [bs, nb] + [float(x) for x in se]
            bs = block_size_min
        counter += 1
    return np.array(rows, dtype=np.float64).reshape(-1, 7), summary


def write_data_blocking(path, stderrs: np.ndarray) -> None:
    """DataBlockingResults.dat with the header and formats of Analyze_A_SD.py:200."""
    np.savetxt(path, stderrs, header="blockSize\tnumBlocks\tStdErrE\tStdErrE2\tStdErrL\tStdErrL2\tStdErrLE", comments="",
               fmt=["%.8g", "%.12g", "%.10g", "%.10g", "%.10g", "%.10g", "%.10g"], delimiter="\t")


def block_stdevs(thermo: np.ndarray, eq_steps: int = 0, block_size_min: int = 100000, interval: int | None = None) -> np.ndarray:
    """Analyze_Fluctuations-style scan: [blockSize, std(block <E>), std(block <L>)], block size growing by
    interval*ceil(1.3*blockSize/interval) up to half the run."""
    interval = interval or print_interval(thermo)
    t = after_equilibration(thermo, eq_steps, interval)
    check_spacing(t, interval)
    rows, bs = [], block_size_min
    bs_max = t.size * interval / 2
    while bs <= bs_max:
        bm = block_means(t, bs, interval, ("Step", "Econf", "L"))
        rows.append([bs, float(np.std(bm[:, 1])), float(np.std(bm[:, 2]))])
        bs = interval * int(np.ceil(1.3 * bs / interval))
    return np.array(rows, dtype=np.float64).reshape(-1, 3)


def blocked_autocorrelation(x) -> np.ndarray:
    """Plot_AutoCorrelation.py:15-29: c[b] = corrcoef(block means m[:-1], m[1:]) for block size b = 1 .. len/2; c[0] = 1."""
    x = np.asarray(x, dtype=np.float64).ravel()
    n = x.size
    c = 1000.0 * np.ones(n // 2 + 1)
    c[0] = 1
    for b in range(1, n // 2 + 1):
        nb = n // b
        m = x[:nb * b].reshape(nb, b).mean(axis=1)
        c[b] = np.corrcoef(m[:-1], m[1:])[0, 1] if nb > 2 else np.nan
    return c


def lag_autocorrelation(x, max_lag: int | None = None) -> np.ndarray:
    """Plot_AutoCorrelation.py:37-43 (`unMeaned`): c[d] = corrcoef(x[:-d], x[d:]) for d = 1 .. len-3; c[0] = 1."""
    x = np.asarray(x, dtype=np.float64).ravel()
    length = x.size - 2 if max_lag is None else min(x.size - 2, max_lag + 1)
    c = 1000.0 * np.ones(max(length, 1))
    c[0] = 1
    for d in range(1, length):
        c[d] = np.corrcoef(x[:-d], x[d:])[0, 1]
    return c


def sort_summary(lines) -> list:
    """scripts/Sort_Summary.bash: sort -k1,1g -k2,2g | uniq on whitespace-separated rows (header lines stay first)."""
    def key(s):
        f = s.split()
        try:
            return (0, float(f[0]), float(f[1]))
        except (ValueError, IndexError):
            return (-1, 0.0, 0.0)
    out, last = [], None
    for s in sorted((s for s in lines if s.strip()), key=key):
        if s != last:
            out.append(s)
        last = s
    return out


def state_point_of(path: str):
    """P and T from a RunJobs-style directory name P<P>_T<T>[_...] (Analyze_Mean.py:100-101)."""
    m = re.search(r"P([0-9.eE+-]+)_T([0-9.eE+-]+)", path)
    if not m:
        raise ValueError(f"no P<P>_T<T> in {path}")
    return float(m.group(1)), float(m.group(2).rstrip("_/"))


def read_eq_steps(dirpath) -> np.ndarray:
    """EqSteps.dat rows: P T EquilSteps [UncorrelatedBlockSize LEstartSteps]; absent file = no rows."""
    f = Path(dirpath) / "EqSteps.dat"
    if not f.exists():
        return np.empty((0, 5))
    d = np.atleast_2d(np.genfromtxt(f))
    if d.shape[1] < 5:
        d = np.hstack([d, np.zeros((d.shape[0], 5 - d.shape[1]))])
    return d


def analyse_directory(dirpath, potential: str = "LJ", N: int | None = None, block_size: int = 1000000,
                      block_size_min: int = 100000) -> dict:
    """Everything the scripts produce as tables, for every state-point directory under `dirpath`:
    <potential>_Means.dat (Analyze_Mean.py), Summary_sorted.txt, Summary_SD_tmp.txt and per-run
    DataBlockingResults.dat (Analyze_A_SD.py), Response.dat (block-averaged response functions)."""
    dirpath = Path(dirpath)
    eq = read_eq_steps(dirpath)
    files = sorted(glob.glob(str(dirpath / "P*_T*" / "thermo.dat.mcs")))
    means_lines, sd_lines, resp_lines = [], [], []
    for tf in files:
        P, T = state_point_of(tf)
        thermo = read_thermo(tf)
        if thermo.size < 3:
            continue
        interval = print_interval(thermo)
        row = eq[(eq[:, 0] == P) & (eq[:, 1] == T)] if eq.size else np.empty((0, 5))
        eq_steps, unc, le_start = (int(row[0, 2]), int(row[0, 3]), int(row[0, 4])) if row.shape[0] else (0, 0, 0)
        m = run_means(thermo, eq_steps, le_start if le_start > 0 else None, interval)
        means_lines.append("\t".join(str(x) for x in (P, T, m["Energy"], m["Energy2"], m["Length"], m["Length2"], m["LengthEnergy"])))
        t = after_equilibration(thermo, eq_steps, interval)
        bs = min(block_size, max(interval, (t.size // 4) * interval))
        bm = block_means(t, bs, interval)
        if bm.shape[0]:
            rf = response_functions(P, T, bm[:, 1], bm[:, 2], bm[:, 3], bm[:, 4], bm[:, 7])
            resp_lines.append("\t".join(str(x) for x in [P, T, bs, bm.shape[0]] +
                                        [float(np.mean(rf[k])) for k in ("cp", "betaT", "alphaP", "gammaV", "betaS", "muJT")]))
        stderrs, summary = data_blocking(thermo, eq_steps, unc, le_start, min(block_size_min, max(interval, (t.size // 8) * interval)), interval)
        if stderrs.shape[0]:
            write_data_blocking(Path(tf).parent / "DataBlockingResults.dat", stderrs)
            pick = summary if summary is not None else list(stderrs[-1])
            sd_lines.append("\t".join([str(N if N is not None else "")] + [str(P), str(T), str(int(pick[0])), str(int(pick[1]))] +
                                      [str(x) for x in pick[2:]]))
    (dirpath / f"{potential}_Means.dat").write_text(
        "P\tT\tEnergy\tEnergy2\tLength\tLength2\tLengthEnergy\n" + "".join(s + "\n" for s in means_lines))
    (dirpath / "Summary_sorted.txt").write_text("".join(s + "\n" for s in sort_summary(means_lines)))
    (dirpath / "Summary_SD_tmp.txt").write_text(
        "N\tP\tT\tblockSize\tnumBlocks\tSEM_Energy\tSEM_EnergySq\tSEM_Length\tSEM_LengthSq\tSEM_LE\n" + "".join(s + "\n" for s in sd_lines))
    (dirpath / "Response.dat").write_text(
        "P\tT\tblockSize\tnumBlocks\tcp\tbetaT\talphaP\tgammaV\tbetaS\tmuJT\n" + "".join(s + "\n" for s in resp_lines))
    return {"runs": len(files), "means": means_lines, "sd": sd_lines, "response": resp_lines}


def read_thermo_chains(path) -> dict:
    """jmm_run's many-chain file thermo_chains.dat.mcs (chain id + the 13 thermo columns per row,
    csrc/host/jmm_main.cpp print_thermo) -> {chain id: structured array like read_thermo}."""
    with open(path) as f:
        first = f.readline().split()
    skip = 1 if first and not first[0].lstrip("-").isdigit() else 0
    data = np.loadtxt(path, skiprows=skip, ndmin=2)
    out = {}
    ids = data[:, 0].astype(np.int64)
    for c in np.unique(ids):
        rows = data[ids == c]
        a = np.empty(rows.shape[0], dtype=[(n, np.float64) for n in COLUMNS])
        for k, n in enumerate(COLUMNS):
            a[n] = rows[:, k + 1]
        out[int(c)] = a
    return out


def read_summary(path) -> np.ndarray:
    """jmm_run's Summary.dat (one row per chain: P, T, N, samples, the 12 run means, acceptance ratios, final E and L)."""
    return np.atleast_1d(np.genfromtxt(path, names=True, delimiter="\t"))


def analyse_sweep(outdir, potential: str = "LJ", eq_steps: int = 0, block_size: int | None = None,
                  block_size_min: int | None = None) -> dict:
    """The tables of analyse_directory for a jmm_run state-point sweep (one process, many chains): reads
    Summary.dat + thermo_chains.dat.mcs from `outdir`, writes <potential>_Means.dat, Summary_sorted.txt,
    Summary_SD_tmp.txt, Response.dat and DataBlockingResults.chain<id>.dat there."""
    outdir = Path(outdir)
    summ = read_summary(outdir / "Summary.dat")
    chains = read_thermo_chains(outdir / "thermo_chains.dat.mcs")
    means_lines, sd_lines, resp_lines = [], [], []
    for row in summ:
        c, P, T, N = int(row["chain"]), float(row["P"]), float(row["T"]), int(row["N"])
        thermo = chains.get(c)
        if thermo is None or thermo.size < 3:
            continue
        interval = print_interval(thermo)
        m = run_means(thermo, eq_steps, None, interval)
        means_lines.append("\t".join(str(x) for x in (P, T, m["Energy"], m["Energy2"], m["Length"], m["Length2"], m["LengthEnergy"])))
        t = after_equilibration(thermo, eq_steps, interval)
        bs = block_size or max(interval, (t.size // 4) * interval)
        bm = block_means(t, bs, interval)
        if bm.shape[0]:
            rf = response_functions(P, T, bm[:, 1], bm[:, 2], bm[:, 3], bm[:, 4], bm[:, 7])
            resp_lines.append("\t".join(str(x) for x in [P, T, bs, bm.shape[0]] +
                                        [float(np.mean(rf[k])) for k in ("cp", "betaT", "alphaP", "gammaV", "betaS", "muJT")]))
        stderrs, _ = data_blocking(thermo, eq_steps, 0, 0, block_size_min or max(interval, (t.size // 8) * interval), interval)
        if stderrs.shape[0]:
            write_data_blocking(outdir / f"DataBlockingResults.chain{c}.dat", stderrs)
            pick = list(stderrs[-1])
            sd_lines.append("\t".join([str(N), str(P), str(T), str(int(pick[0])), str(int(pick[1]))] + [str(x) for x in pick[2:]]))
    (outdir / f"{potential}_Means.dat").write_text(
        "P\tT\tEnergy\tEnergy2\tLength\tLength2\tLengthEnergy\n" + "".join(s + "\n" for s in means_lines))
    (outdir / "Summary_sorted.txt").write_text("".join(s + "\n" for s in sort_summary(means_lines)))
    (outdir / "Summary_SD_tmp.txt").write_text(
        "N\tP\tT\tblockSize\tnumBlocks\tSEM_Energy\tSEM_EnergySq\tSEM_Length\tSEM_LengthSq\tSEM_LE\n" + "".join(s + "\n" for s in sd_lines))
    (outdir / "Response.dat").write_text(
        "P\tT\tblockSize\tnumBlocks\tcp\tbetaT\talphaP\tgammaV\tbetaS\tmuJT\n" + "".join(s + "\n" for s in resp_lines))
    return {"runs": len(means_lines), "means": means_lines, "sd": sd_lines, "response": resp_lines}


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        print(__doc__)
        return 1
    pot = argv[1] if len(argv) > 1 else "LJ"
    if (Path(argv[0]) / "thermo_chains.dat.mcs").exists():
        res = analyse_sweep(argv[0], potential=pot)
    else:
        res = analyse_directory(argv[0], potential=pot)
    print(f"analysed {res['runs']} run(s) under {os.path.abspath(argv[0])}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
