#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the COMPILED REFERENCE.

The reference ships no expected outputs, so the goldens are outputs of the reference
itself: oracle/Makefile (`make ref`) compiles /root/reference/src unchanged except for the one-token
phiij[6]->[9] fix, with the gsl_rng_taus2 shim, and this script runs that binary single-threaded
(the only mode in which it is deterministic) on the reference's own decks.  Run it in the build
container (it needs /root/reference); the GPU box only reads the committed outputs.

    python tests/golden/make_golden.py [--full]     # --full also runs the 5 000 000-step smalltest
"""
import hashlib
import json
import re
import shutil
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

REF_TEST = Path("/root/reference/test")


def deck_with(text, **kw):
    for k, v in kw.items():
        text = re.sub(rf"^{k}\s+.*$", f"{k:<10} {v}", text, flags=re.M)
    return text


def summarise(stdout: str) -> dict:
    s = {}
    m = re.search(r"\nE = (\S+)", stdout)
    s["final_E_printed"] = m.group(1)
    m = re.search(r"(\d+)/(\d+)\s+(\d+)/(\d+)\s*$", stdout.strip())
    s["counters"] = [int(x) for x in m.groups()]
    s["verified"] = stdout.count("Energy was just verified")
    s["discrepancy"] = stdout.count("Energy discrepancy")
    s["relax_calls"] = stdout.count("Relaxing the Volume")
    s["not_understood"] = re.findall(r"Property command (\S+) not understood", stdout)
    m = re.search(r"Relaxation converged\. Length,energy: (\S+),(\S+)", stdout)
    if m:
        s["first_relax"] = [m.group(1), m.group(2)]
    s["maxStep_updates"] = re.findall(r"new maxStep: (\S+)", stdout)[-4:]
    return s


def last_frame(config_text: str):
    lines = config_text.strip().splitlines()
    n = int(lines[0])
    frame = lines[-(n + 2):]
    box = float(frame[1].split("Box length:")[1])
    return box, [float(x.split()[3]) for x in frame[2:]]


def run(name, deck, keep_rng=True, keep_files=True, keep_hist=False):
    out = HERE / name
    out.mkdir(exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        R = O.run_reference(deck, tmp, record_rng=True)
        (out / "INPUT").write_text(deck)
        s = summarise(R["stdout"])
        thermo = Path(R["thermo"]).read_bytes()
        config = Path(R["config"]).read_bytes()
        s["thermo_md5"] = hashlib.md5(thermo).hexdigest()
        s["config_md5"] = hashlib.md5(config).hexdigest()
        s["rng_words"] = int(R["rng"].size)
        s["rng_first8"] = [int(x) for x in R["rng"][:8]]
        s["rng_md5"] = hashlib.md5(R["rng"].tobytes()).hexdigest()
        box, pos = last_frame(config.decode())
        s["last_frame_box"] = box
        s["last_frame_r"] = pos
        s["stdout_thermo_lines"] = re.findall(r"^\d+  \S+  \S+  \S+  \S+$", R["stdout"], flags=re.M)[:40]
        if keep_files:
            (out / "thermo.dat.mcs").write_bytes(thermo)
            (out / "config.dat.mcs").write_bytes(config)
        else:
            (out / "thermo.tail.mcs").write_bytes(b"\n".join(thermo.splitlines()[-5:]) + b"\n")
        if keep_hist:
            for f in sorted(Path(tmp).glob("rho.dat.mcs")) + sorted(Path(tmp).glob("g*.dat.mcs")):
                (out / f.name).write_bytes(f.read_bytes())
                s[f.name + "_md5"] = hashlib.md5(f.read_bytes()).hexdigest()
        if keep_rng:
            R["rng"].tofile(out / "rng.u32")
        (out / "summary.json").write_text(json.dumps(s, indent=1) + "\n")
        print(name, s["final_E_printed"], s["counters"], "words", s["rng_words"])


def reference_ensemble(name, deck, nseeds=32, seed0=92847):
    """`nseeds` independent runs of the COMPILED REFERENCE (seeds seed0+k), run in parallel; stores the
    per-run block means (thermo rows) so that the ensemble error bar is a plain standard error over
    independent chains (a single chain's 50k-step blocks are visibly autocorrelated)."""
    from concurrent.futures import ThreadPoolExecutor
    out = HERE / name
    out.mkdir(exist_ok=True)

    def one(k):
        with tempfile.TemporaryDirectory() as tmp:
            R = O.run_reference(deck_with(deck, SEED=seed0 + k), tmp)
            rows = [[float(x) for x in l.split("\t")] for l in Path(R["thermo"]).read_text().splitlines()[1:]]
            return {"seed": seed0 + k, "counters": summarise(R["stdout"])["counters"], "blocks": rows}

    with ThreadPoolExecutor(8) as ex:
        runs = list(ex.map(one, range(nseeds)))
    (out / "INPUT").write_text(deck)
    cols = ["Step", "Econf", "Econf2", "L", "L2", "LEconf", "rho", "rho2", "Virial", "Virial2", "EconfVir", "HV", "HV2"]
    (out / "summary.json").write_text(json.dumps({"columns": cols, "runs": runs}) + "\n")
    print(name, len(runs), "runs")


def main():
    O.build()
    small = (REF_TEST / "INPUT_smalltest").read_text()
    run("smalltest_12", deck_with(small, NUMSTEPS=12, TPI=1, CPI=1))
    run("smalltest_2000", deck_with(small, NUMSTEPS=2000, TPI=100, CPI=500))
    run("smalltest_20000", deck_with(small, NUMSTEPS=20000), keep_rng=False)
    run("inputstd", (REF_TEST / "INPUTstd").read_text())
    big = (REF_TEST / "INPUT").read_text()
    run("input_n2000_40", deck_with(big, NUMSTEPS=40, CPI=40, TPI=5), keep_rng=True)
    # density / two-particle-density histograms (rho.dat.mcs, g<k>.dat.mcs): SURVEY §8f N2
    run("smalltest_hist", deck_with(small, NUMSTEPS=3000, TPI=500, CPI=1500, RBW=0.5, RHONB=40, RHOPI=500, GSW=5.0, GNS=4,
                                    GBW=0.2, GNB=60, GPI=1000), keep_rng=False, keep_hist=True)
    run("inputstd_hist", deck_with((REF_TEST / "INPUTstd").read_text(), NUMSTEPS=2000, TPI=500, CPI=1000, RBW=0.5, RHONB=40,
                                   RHOPI=250, GSW=10, GNS=2, GBW=0.1, GNB=50, GPI=500), keep_rng=False, keep_hist=True)
    # block averages for the 2-sigma ensemble test (north_star): 32 reference runs x 5 blocks of 200 000 steps, no RELAX
    blocks = deck_with(small.replace("RELAX\n", ""), NUMSTEPS=1000000, TPI=200000, CPI=1000000)
    reference_ensemble("smalltest_ensemble", blocks)
    if "--full" in sys.argv:
        run("smalltest_full", small, keep_rng=False, keep_files=False)


if __name__ == "__main__":
    main()
