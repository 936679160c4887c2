"""jmmonedmc_b200.analysis against independent restatements of scripts/Analyze_Mean.py, Analyze_A_SD.py,
Analyze_Fluctuations.py and Plot_AutoCorrelation.py, on the reference's own thermo golden."""
import math
from pathlib import Path

import numpy as np
import pytest

from jmmonedmc_b200 import analysis as A

GOLD = Path(__file__).parent / "golden"


@pytest.fixture(scope="module")
def thermo():
    return A.read_thermo(GOLD / "smalltest_20000" / "thermo.dat.mcs")       # reference output, TPI 1000, 21 rows


def test_read_thermo_columns_and_aliases(thermo):
    assert thermo.dtype.names == A.COLUMNS and thermo.size == 21
    assert thermo["Step"][0] == 0 and A.print_interval(thermo) == 1000
    assert np.array_equal(A.col(thermo, "Energy"), thermo["Econf"]) and np.array_equal(A.col(thermo, "lE"), thermo["LEconf"])
    # the rows SURVEY §8c quotes from the compiled reference
    assert thermo["Econf"][1] == pytest.approx(-6.8997892, abs=1e-7) and thermo["L"][2] == pytest.approx(11.001627, abs=1e-6)
    A.check_spacing(thermo, 1000)
    bad = thermo.copy(); bad["Step"][5] += 1
    with pytest.raises(ValueError):
        A.check_spacing(bad, 1000)


def test_run_means_follow_analyze_mean(thermo):
    m = A.run_means(thermo, eq_steps=3000, le_start_steps=8000)
    raw = np.loadtxt(GOLD / "smalltest_20000" / "thermo.dat.mcs", skiprows=1)
    cut = raw[3000 // 1000 + 1:]                                             # Analyze_Mean.py:114
    assert m["Energy"] == np.mean(cut[:, 1]) and m["Energy2"] == np.mean(cut[:, 2])
    assert m["Length"] == np.mean(cut[:, 3]) and m["Length2"] == np.mean(cut[:, 4])
    assert m["LengthEnergy"] == np.mean(cut[(8000 - 3000) // 1000:, 5])      # Analyze_Mean.py:123
    assert m["rows"] == 17


def test_block_means_and_response_functions(thermo):
    t = A.after_equilibration(thermo, 0, 1000)
    bm = A.block_means(t, 5000, 1000)
    assert bm.shape == (4, 8)
    raw = np.loadtxt(GOLD / "smalltest_20000" / "thermo.dat.mcs", skiprows=1)[1:]
    for b in range(4):
        blk = raw[5 * b:5 * b + 5]
        assert np.allclose(bm[b], [blk[:, 0].mean(), blk[:, 1].mean(), blk[:, 2].mean(), blk[:, 3].mean(), blk[:, 4].mean(),
                                   blk[:, 8].mean(), blk[:, 9].mean(), blk[:, 5].mean()], rtol=0, atol=0)
    P, T = 1.0, 0.9
    rf = A.response_functions(P, T, bm[:, 1], bm[:, 2], bm[:, 3], bm[:, 4], bm[:, 7])
    # scalar restatement of Analyze_Mean.py:162-176 for one block
    E, E2, L, L2, LE = bm[2, 1], bm[2, 2], bm[2, 3], bm[2, 4], bm[2, 7]
    cp = (E2 - E * E + 2.0 * P * (LE - L * E) + P * P * (L2 - L * L)) / T / T
    betaT = 1.0 / L / T * (L2 - L * L)
    alphaP = 1 / T / T / L * ((LE - L * E) + P * (L2 - L * L))
    assert rf["cp"][2] == cp and rf["betaT"][2] == betaT and rf["alphaP"][2] == alphaP
    assert rf["gammaV"][2] == alphaP / betaT and rf["betaS"][2] == betaT - alphaP * alphaP * T * L / cp
    assert rf["muJT"][2] == L / cp * (alphaP * T - 1.0)
    # an ideal-gas-like check of the formulas: no energy, L exponential-distributed moments -> betaT = <L>/T * 1
    rf0 = A.response_functions(2.0, 0.5, 0.0, 0.0, 3.0, 18.0, 0.0)
    assert float(rf0["betaT"]) == pytest.approx((18.0 - 9.0) / 3.0 / 0.5)


def test_data_blocking_progression_and_sem(thermo, tmp_path):
    se, summary = A.data_blocking(thermo, eq_steps=0, block_size_min=2000)
    assert summary is None                                                    # no uncorrelated block size given
    assert list(se[:, 0]) == [2000, 4000, 6000, 8000, 10000]                  # blockSizeMin steps up to half the run
    assert list(se[:, 1]) == [10, 5, 3, 2, 2]
    raw = np.loadtxt(GOLD / "smalltest_20000" / "thermo.dat.mcs", skiprows=1)[1:]
    means = raw[:20, 1].reshape(5, 4).mean(axis=1)                            # block size 4000 = 4 rows
    assert se[1, 2] == pytest.approx(np.std(means) / math.sqrt(5), rel=1e-14)
    se2, summary2 = A.data_blocking(thermo, eq_steps=2000, uncorrelated_block_size=3000, le_start_steps=5000, block_size_min=2000)
    assert se2[0, 0] == 3000 and summary2[0] == 3000 and se2[1, 0] == 2000    # Analyze_A_SD.py:146-151,186-197
    le = raw[5:, 5]
    assert summary2[6] == pytest.approx(np.std(le[:15].reshape(5, 3).mean(axis=1)) / math.sqrt(5), rel=1e-14)
    A.write_data_blocking(tmp_path / "DataBlockingResults.dat", se)
    lines = (tmp_path / "DataBlockingResults.dat").read_text().splitlines()
    assert lines[0] == "blockSize\tnumBlocks\tStdErrE\tStdErrE2\tStdErrL\tStdErrL2\tStdErrLE" and len(lines) == 6
    back = np.genfromtxt(tmp_path / "DataBlockingResults.dat", names=True)
    assert np.allclose(back["StdErrE"], se[:, 2], rtol=1e-9)


def test_block_stdevs_growth(thermo):
    sd = A.block_stdevs(thermo, block_size_min=1000)
    assert list(sd[:, 0]) == [1000, 2000, 3000, 4000, 6000, 8000]             # interval*ceil(1.3*b/interval)
    raw = np.loadtxt(GOLD / "smalltest_20000" / "thermo.dat.mcs", skiprows=1)[1:]
    assert sd[0, 1] == pytest.approx(np.std(raw[:, 1])) and sd[0, 2] == pytest.approx(np.std(raw[:, 3]))


def test_autocorrelations():
    rng = np.random.default_rng(7)
    white = rng.standard_normal(4000)
    c = A.lag_autocorrelation(white, max_lag=20)
    assert c[0] == 1 and np.all(np.abs(c[1:]) < 0.06)
    phi, x = 0.8, np.zeros(20000)
    e = rng.standard_normal(x.size)
    for i in range(1, x.size):
        x[i] = phi * x[i - 1] + e[i]
    c = A.lag_autocorrelation(x, max_lag=5)
    assert np.allclose(c[1:6], phi ** np.arange(1, 6), atol=0.03)
    for d in (1, 3):
        assert c[d] == np.corrcoef(x[:-d], x[d:])[0, 1]                       # Plot_AutoCorrelation.py:42
    cb = A.blocked_autocorrelation(x[:400])
    assert cb[0] == 1 and cb.size == 201
    m = x[:400].reshape(100, 4).mean(axis=1)
    assert cb[4] == pytest.approx(np.corrcoef(m[:-1], m[1:])[0, 1], rel=1e-12)  # :21-27
    assert cb[1] > cb[40]                                                     # block means decorrelate as blocks grow


def test_sort_summary_and_directory_layout(tmp_path, thermo):
    assert A.sort_summary(["1.0\t0.5\ta", "0.1\t0.9\tb", "0.1\t0.2\tc", "0.1\t0.9\tb", ""]) == ["0.1\t0.2\tc", "0.1\t0.9\tb", "1.0\t0.5\ta"]
    assert A.state_point_of("data/LJ/m-1/N80/P0.5_T0.7/thermo.dat.mcs") == (0.5, 0.7)
    assert A.state_point_of("x/P1.0_T0.9_123456/thermo.dat.mcs") == (1.0, 0.9)
    src = (GOLD / "smalltest_20000" / "thermo.dat.mcs").read_text()
    for name in ("P1.0_T0.9", "P0.5_T0.9"):                                   # scripts/RunJobs.bash:27 layout
        (tmp_path / name).mkdir()
        (tmp_path / name / "thermo.dat.mcs").write_text(src)
    (tmp_path / "EqSteps.dat").write_text("1.0 0.9 2000 0 4000\n0.5 0.9 0 0 0\n")
    res = A.analyse_directory(tmp_path, potential="LJ", N=10)
    assert res["runs"] == 2
    means = (tmp_path / "LJ_Means.dat").read_text().splitlines()
    assert means[0] == "P\tT\tEnergy\tEnergy2\tLength\tLength2\tLengthEnergy" and len(means) == 3
    want = A.run_means(thermo, 2000, 4000)
    row = [float(x) for x in means[2].split("\t")]
    assert row[:2] == [1.0, 0.9] and row[2] == want["Energy"] and row[6] == want["LengthEnergy"]
    assert (tmp_path / "Summary_sorted.txt").read_text().splitlines()[0].startswith("0.5\t0.9")
    assert (tmp_path / "P1.0_T0.9" / "DataBlockingResults.dat").exists()
    assert len((tmp_path / "Summary_SD_tmp.txt").read_text().splitlines()) == 3
    assert len((tmp_path / "Response.dat").read_text().splitlines()) == 3


def test_sweep_tables_from_batch_driver_files(tmp_path, thermo):
    """Summary.dat + thermo_chains.dat.mcs as csrc/host/jmm_main.cpp writes them for a many-chain run."""
    raw = np.loadtxt(GOLD / "smalltest_20000" / "thermo.dat.mcs", skiprows=1)
    with open(tmp_path / "thermo_chains.dat.mcs", "w") as f:
        f.write("chain\t" + "\t".join(A.COLUMNS) + "\n")
        for row in raw:
            for c in (4, 5):
                f.write(f"{c}\t{int(row[0])}\t" + "\t".join(f"{(1 + 0.01 * (c - 4)) * v:.8G}" for v in row[1:]) + "\n")
    hdr = ("chain\tP\tT\tN\tsamples\tEconf\tEconf2\tL\tL2\tLEconf\trho\trho2\tVirial\tVirial2\tEconfVir\tHV\tHV2\t"
           "dAccRatio\tvAccRatio\tEfinal\tLfinal\n")
    with open(tmp_path / "Summary.dat", "w") as f:
        f.write(hdr)
        for c, P in ((4, 0.7), (5, 0.3)):
            f.write("\t".join([str(c), str(P), "0.9", "10", "20001"] + ["0"] * 16) + "\n")
    ch = A.read_thermo_chains(tmp_path / "thermo_chains.dat.mcs")
    assert sorted(ch) == [4, 5] and ch[4].size == 21 and np.allclose(ch[4]["Econf"], thermo["Econf"], rtol=1e-7)
    res = A.analyse_sweep(tmp_path)
    assert res["runs"] == 2
    srt = (tmp_path / "Summary_sorted.txt").read_text().splitlines()
    assert srt[0].startswith("0.3\t0.9") and srt[1].startswith("0.7\t0.9")
    assert (tmp_path / "DataBlockingResults.chain5.dat").exists()


# ---------------------------------------------------------------------------------------------------------------------
# Pinned against the reference's OWN scripts: tests/golden/analysis_scripts holds what scripts/Analyze_Mean.py,
# Analyze_A_SD.py and Plot_AutoCorrelation.py (read from /root/reference/scripts, run under a python-2 compatibility shim,
# tests/golden/make_golden_analysis.py) computed on a synthetic thermo file; analysis.py gets the same numbers in the
# writer's 13-column format.
SCRIPTS = GOLD / "analysis_scripts"


@pytest.fixture(scope="module")
def script_gold():
    import json
    return json.loads((SCRIPTS / "golden.json").read_text())


def test_means_and_response_functions_equal_analyze_mean_py(script_gold):
    g = script_gold["Analyze_Mean"]
    thermo = A.read_thermo(SCRIPTS / "mean_thermo.dat.mcs")
    m = A.run_means(thermo, eq_steps=g["eq_steps"], le_start_steps=g["le_start_steps"], interval=g["interval"])
    header, row = g["LJ_Means.dat"].strip().splitlines()
    assert header.split("\t") == ["P", "T", "Energy", "Energy2", "Length", "Length2", "LengthEnergy"]          # Analyze_Mean.py:64
    want = [float(x) for x in row.split("\t")]
    got = [g["P"], g["T"], m["Energy"], m["Energy2"], m["Length"], m["Length2"], m["LengthEnergy"]]
    assert np.allclose(got, want, rtol=1e-14, atol=0)
    t = A.after_equilibration(thermo, g["eq_steps"], g["interval"])
    bm = A.block_means(t, g["block_size"], g["interval"])
    plots = g["plots"]
    assert bm.shape[0] == len(plots["L"]["y"]) == 7
    lb = g["LEstartBlock"]                                                   # the LE-dependent plots start at this block (:105)
    assert np.allclose(bm[:, 0], plots["L"]["x"], rtol=1e-14) and np.allclose(bm[:, 3], plots["L"]["y"], rtol=1e-14)
    assert np.allclose(bm[lb:, 1] * g["epsilon"], plots["E_times_epsilon"]["y"], rtol=1e-13)
    assert np.allclose(bm[lb:, 2], plots["E2"]["y"], rtol=1e-14) and np.allclose(bm[:, 4], plots["L2"]["y"], rtol=1e-14)
    assert np.allclose(bm[lb:, 7], plots["LE"]["y"], rtol=1e-14)
    rf = A.response_functions(g["P"], g["T"], bm[:, 1], bm[:, 2], bm[:, 3], bm[:, 4], bm[:, 7])
    # differences of large numbers (E2 - E E ~ 30 against 2.3e6): compare at the accuracy the cancellation leaves
    for name, sl in (("cp", slice(lb, None)), ("betaT", slice(None)), ("betaS", slice(lb, None)), ("alphaP", slice(lb, None)),
                     ("gammaV", slice(lb, None)), ("muJT", slice(lb, None))):
        assert np.allclose(rf[name][sl], plots[name]["y"], rtol=1e-9, atol=0), name


@pytest.mark.parametrize("case", ["with_uncorrelated_block", "scan_only"])
def test_data_blocking_equals_analyze_a_sd_py(script_gold, case, tmp_path):
    g = script_gold["Analyze_A_SD:" + case]
    thermo = A.read_thermo(SCRIPTS / "sd_thermo.dat.mcs")
    stderrs, summary = A.data_blocking(thermo, eq_steps=g["eq_steps"], uncorrelated_block_size=g["uncorrelated_block_size"],
                                       le_start_steps=g["le_start_steps"])
    A.write_data_blocking(tmp_path / "DataBlockingResults.dat", stderrs)
    got = (tmp_path / "DataBlockingResults.dat").read_text().splitlines()
    want = g["DataBlockingResults.dat"].splitlines()
    assert got[0] == want[0] and len(got) == len(want)                                   # header and block-size ladder
    G = np.array([[float(x) for x in l.split("\t")] for l in got[1:]]); W = np.array([[float(x) for x in l.split("\t")] for l in want[1:]])
    assert np.array_equal(G[:, :2], W[:, :2]) and np.allclose(G, W, rtol=1e-9)           # (files hold 10 significant digits)
    rows = g["Summary_SD_tmp.txt"].strip().splitlines()
    assert rows[0].split("\t") == ["N", "P", "T", "blockSize", "numBlocks", "SEM_Energy", "SEM_EnergySq", "SEM_Length", "SEM_LengthSq", "SEM_LE"]
    if case == "with_uncorrelated_block":
        f = rows[1].split("\t")
        assert [int(f[3]), int(f[4])] == [int(summary[0]), int(summary[1])]
        assert np.allclose([float(x) for x in f[5:]], summary[2:], rtol=1e-12)
    else:
        assert len(rows) == 1 and summary is None


def test_autocorrelations_equal_plot_autocorrelation_py(script_gold):
    g = script_gold["Plot_AutoCorrelation"]
    x = np.array(g["x"])
    want_b, want_l = np.array(g["blocked"], dtype=float), np.array(g["lag"], dtype=float)
    got_b, got_l = A.blocked_autocorrelation(x), A.lag_autocorrelation(x)
    assert got_b.shape == want_b.shape and got_l.shape == want_l.shape
    ok = np.isfinite(want_b)                                                  # two blocks: corrcoef of single points is nan in both
    assert np.array_equal(np.isfinite(got_b), ok) and np.allclose(got_b[ok], want_b[ok], rtol=1e-12, atol=1e-14)
    assert np.allclose(got_l, want_l, rtol=1e-12, atol=1e-14)
