"""TEST INFRASTRUCTURE — ctypes binding of the CPU restatement (oracle/jmm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (jmmonedmc_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "libjmm_oracle.so"
REF_DIR = HERE / "_ref"
REF_BIN = REF_DIR / "jmmOneDMC_ref"
REF_SERIAL_BIN = REF_DIR / "jmmOneDMC_serial"
COMPAT_BIN = REF_DIR / "jmmOneDMC_gpu"       # reference Main.cpp + readInput.cpp linked against libjmmgpu.so

POT = {"LJ": 0, "LJcut": 1, "HARMONIC": 2}
ENS = {"NPT": 0, "NLT": 1}
RNG_TAUS2, RNG_PHILOX, RNG_RECORDED = 0, 1, 2
MODE_TABLE, MODE_RECOMPUTE = 0, 1
TOT_NAMES = ("E", "Vir", "E12", "Vir12", "E6", "Vir6", "HV", "HV12", "HV6")
ACC_NAMES = ("rho", "rho2", "L", "L2", "E", "E2", "LE", "Vir", "Vir2", "EVir", "HV", "HV2")


class Config(C.Structure):
    _fields_ = [
        ("N", C.c_uint64), ("nbn", C.c_int32), ("pot", C.c_int32), ("cutoff", C.c_double),
        ("ensemble", C.c_int32), ("relax", C.c_int32),
        ("P", C.c_double), ("T", C.c_double), ("L", C.c_double),
        ("maxStep", C.c_double), ("maxdl", C.c_double),
        ("eci", C.c_uint64), ("mdai", C.c_uint64), ("mvai", C.c_uint64),
        ("seed", C.c_uint64), ("chain_id", C.c_uint64),
        ("rng_kind", C.c_int32), ("mode", C.c_int32),
    ]


def build(force: bool = False) -> None:
    """Compile the restatement (and the reference itself when /root/reference is present)."""
    make = ["make", "-s", "-C", str(HERE)]
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < (HERE / "jmm_oracle.c").stat().st_mtime:
        subprocess.run(make + ["oracle"], check=True)
    if Path("/root/reference/src").is_dir():
        if force or not REF_BIN.exists():
            subprocess.run(make + ["ref"], check=True)
        gpu_lib = HERE.parent / "jmmonedmc_b200" / "libjmmgpu.so"
        shim = HERE.parent / "jmmonedmc_b200" / "csrc" / "host" / "jmm_mcstate_compat.cpp"
        if gpu_lib.exists() and (force or not COMPAT_BIN.exists() or COMPAT_BIN.stat().st_mtime < shim.stat().st_mtime):
            subprocess.run(make + ["compat"], check=True)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        dp, u64p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
        L.jmo_create.restype = C.c_void_p
        L.jmo_create.argtypes = [C.POINTER(Config)]
        L.jmo_destroy.argtypes = [C.c_void_p]
        L.jmo_set_recorded.argtypes = [C.c_void_p, u32p, C.c_uint64]
        L.jmo_recorded_cursor.restype = C.c_uint64
        L.jmo_recorded_cursor.argtypes = [C.c_void_p]
        L.jmo_phi.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, dp]
        for name in ("jmo_step0", "jmo_update_thermo", "jmo_cadence", "jmo_zero_accum"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.jmo_relax_volume.argtypes = [C.c_void_p]
        L.jmo_step.argtypes = [C.c_void_p]
        L.jmo_run.argtypes = [C.c_void_p, C.c_uint64]
        L.jmo_config_totals.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_double, dp]
        L.jmo_totals_of.argtypes = [dp, C.c_uint64, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, dp]
        for name in ("jmo_N", "jmo_sn", "jmo_relax_calls"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = C.c_uint64
        L.jmo_set_sn.argtypes = [C.c_void_p, C.c_uint64]
        L.jmo_set_sn.restype = None
        L.jmo_l.argtypes = [C.c_void_p]
        L.jmo_l.restype = C.c_double
        L.jmo_get_r.argtypes = [C.c_void_p, dp]
        L.jmo_set_r.argtypes = [C.c_void_p, dp, C.c_double]
        L.jmo_get_totals.argtypes = [C.c_void_p, dp]
        L.jmo_get_accum.argtypes = [C.c_void_p, dp]
        L.jmo_get_counters.argtypes = [C.c_void_p, u64p]
        L.jmo_get_step_sizes.argtypes = [C.c_void_p, dp, dp]
        L.jmo_set_step_sizes.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.jmo_echeck_count.argtypes = [C.c_void_p, u64p]
        L.jmo_echeck_count.restype = C.c_uint64
        L.jmo_taus2_seed.argtypes = [u32p, C.c_uint64]
        L.jmo_taus2_next.argtypes = [u32p]
        L.jmo_taus2_next.restype = C.c_uint32
        L.jmo_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.jmo_colour_halfsweep.argtypes = [dp, C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                           C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, dp]
        L.jmo_colour_halfsweep.restype = C.c_uint64
        L.jmo_colour_of_step.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.jmo_colour_of_step.restype = C.c_int
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---------------------------------------------------------------------------- decks

def parse_deck(text: str) -> dict:
    """KEY value parser with the reference's keyword set (src/readInput.cpp:69-253); test-side only."""
    d = {"ENSEMBLE": "NPT", "RELAX": 0, "CUTOFF": math.inf, "unknown": []}
    num = {"N", "P", "L", "T", "NBN", "NUMSTEPS", "MAXSTEP", "MAXDV", "CPI", "TPI", "GPI", "RHOPI", "RBW",
           "RHONB", "GBW", "GNB", "GSW", "GNS", "SEED", "ENGCHECK", "DADJ", "VADJ"}
    for line in text.splitlines():
        tok = line.split()
        if not tok:
            continue
        k = tok[0]
        if k in num:
            d[k] = float(tok[1])
        elif k == "POT":
            d["POT"] = tok[1] if tok[1] in POT else "LJ"
            d["CUTOFF"] = float(tok[2]) if len(tok) > 2 else math.inf
        elif k == "RELAX":
            d["RELAX"] = 1
        elif k == "ENSEMBLE":
            d["ENSEMBLE"] = tok[1] if tok[1] in ENS else "NPT"
        else:
            d["unknown"].append(k)
    return d


def config_from_deck(d: dict, rng_kind=RNG_TAUS2, mode=MODE_TABLE, chain_id=0, seed=None) -> Config:
    c = Config()
    c.N = int(d["N"]); c.nbn = int(d["NBN"]); c.pot = POT[d["POT"]]
    c.cutoff = d["CUTOFF"] if d["POT"] != "LJ" else math.inf
    c.ensemble = ENS[d["ENSEMBLE"]]; c.relax = int(d.get("RELAX", 0))
    c.P = d.get("P", 0.0); c.T = d["T"]; c.L = d.get("L", 0.0)
    c.maxStep = d["MAXSTEP"]; c.maxdl = d["MAXDV"]
    c.eci = int(d.get("ENGCHECK", 0)); c.mdai = int(d.get("DADJ", 0)); c.mvai = int(d.get("VADJ", 0))
    c.seed = int(d["SEED"]) if seed is None else int(seed)
    c.chain_id = chain_id; c.rng_kind = rng_kind; c.mode = mode
    return c


class Chain:
    """One chain of the restatement."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.L = lib()
        self.h = self.L.jmo_create(C.byref(cfg))
        self._rec = None

    def __del__(self):
        if getattr(self, "h", None):
            self.L.jmo_destroy(self.h)
            self.h = None

    def set_recorded(self, words: np.ndarray):
        self._rec = np.ascontiguousarray(words, dtype=np.uint32)
        self.L.jmo_set_recorded(self.h, self._rec.ctypes.data_as(C.POINTER(C.c_uint32)), self._rec.size)

    def recorded_cursor(self): return int(self.L.jmo_recorded_cursor(self.h))
    def step0(self): self.L.jmo_step0(self.h)
    def relax_volume(self): return self.L.jmo_relax_volume(self.h)
    def update_thermo(self): self.L.jmo_update_thermo(self.h)
    def step(self): return self.L.jmo_step(self.h)
    def cadence(self): self.L.jmo_cadence(self.h)
    def run(self, n): self.L.jmo_run(self.h, int(n))
    def zero_accum(self): self.L.jmo_zero_accum(self.h)

    def start(self):
        """What Main.cpp does before the loop: fad step 0, optional relax, first updateThermo."""
        self.step0()
        if self.cfg.relax and self.cfg.ensemble == ENS["NPT"]:
            self.relax_volume()
        self.update_thermo()

    @property
    def N(self): return int(self.L.jmo_N(self.h))
    @property
    def sn(self): return int(self.L.jmo_sn(self.h))
    def set_step_number(self, sn): self.L.jmo_set_sn(self.h, int(sn))
    @property
    def relax_calls(self): return int(self.L.jmo_relax_calls(self.h))
    @property
    def l(self): return float(self.L.jmo_l(self.h))

    @property
    def r(self):
        a = np.empty(self.N); self.L.jmo_get_r(self.h, _dp(a)); return a

    def set_r(self, r, l):
        a = np.ascontiguousarray(r, dtype=np.float64); self.L.jmo_set_r(self.h, _dp(a), float(l))

    @property
    def totals(self):
        a = np.empty(9); self.L.jmo_get_totals(self.h, _dp(a)); return a

    @property
    def accum(self):
        a = np.empty(12); self.L.jmo_get_accum(self.h, _dp(a)); return a

    @property
    def counters(self):
        a = np.empty(4, dtype=np.uint64); self.L.jmo_get_counters(self.h, a.ctypes.data_as(C.POINTER(C.c_uint64))); return a

    @property
    def step_sizes(self):
        a, b = C.c_double(), C.c_double(); self.L.jmo_get_step_sizes(self.h, C.byref(a), C.byref(b)); return a.value, b.value

    def set_step_sizes(self, a, b): self.L.jmo_set_step_sizes(self.h, a, b)

    @property
    def echecks(self):
        d = C.c_uint64(); n = self.L.jmo_echeck_count(self.h, C.byref(d)); return int(n), int(d.value)

    def config_totals(self, scale=1.0, virflag=1, l=None):
        a = np.empty(9); self.L.jmo_config_totals(self.h, scale, virflag, self.l if l is None else l, _dp(a)); return a

    def enable_histograms(self, rhonb, rbw, gns, gnb, gsw, gbw):
        self.L.jmo_enable_histograms.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_int, C.c_uint64, C.c_double, C.c_double]
        self.L.jmo_enable_histograms(self.h, int(rhonb), float(rbw), int(gns), int(gnb), float(gsw), float(gbw))
        self._hist = (int(rhonb), int(gns), int(gnb))

    def take_histograms(self):
        rhonb, gns, gnb = self._hist
        a = np.zeros(rhonb, dtype=np.int64); g = np.zeros((gns, gnb), dtype=np.int64)
        self.L.jmo_take_histograms.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        self.L.jmo_take_histograms(self.h, a.ctypes.data_as(C.POINTER(C.c_int64)), g.ctypes.data_as(C.POINTER(C.c_int64)))
        return a, g

    def run_deck_hist(self, d, outdir):
        """Full run of deck dict `d` with thermo/config/rho/g<k> files written into outdir."""
        outdir = Path(outdir)
        self.enable_histograms(d["RHONB"], d["RBW"], d["GNS"], d["GNB"], d["GSW"], d["GBW"])
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        op = lambda n: libc.fopen(str(outdir / n).encode(), b"w")
        tf, cf, rf = op("thermo.dat.mcs"), op("config.dat.mcs"), op("rho.dat.mcs")
        gns = int(d["GNS"])
        gfs = (C.c_void_p * gns)(*[op(f"g{k}.dat.mcs") for k in range(gns)])
        self.L.jmo_run_deck_hist.argtypes = [C.c_void_p] + [C.c_uint64] * 5 + [C.c_void_p] * 5
        self.L.jmo_run_deck_hist(self.h, int(d["NUMSTEPS"]), int(d["TPI"]), int(d["CPI"]), int(d["RHOPI"]), int(d["GPI"]),
                                 tf, cf, rf, gfs, None)
        for f in [tf, cf, rf] + list(gfs):
            libc.fclose(f)

    def run_deck(self, numsteps, tpi, cpi, thermo_path=None, config_path=None, log_path=None):
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p; libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        fs = [libc.fopen(str(p).encode(), b"w") if p else None for p in (thermo_path, config_path, log_path)]
        self.L.jmo_run_deck.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        self.L.jmo_run_deck(self.h, int(numsteps), int(tpi), int(cpi), *fs)
        for f in fs:
            if f: libc.fclose(f)


def totals_of(r, nbn, pot, cutoff, scale=1.0, virflag=1, l=1.0):
    a = np.ascontiguousarray(r, dtype=np.float64); out = np.empty(9)
    lib().jmo_totals_of(_dp(a), a.size, int(nbn), int(pot), float(cutoff), float(scale), int(virflag), float(l), _dp(out))
    return out


def phi(pot, d, cutoff=math.inf, virflag=1, l=1.0):
    out = np.empty(9); lib().jmo_phi(int(pot), float(d), float(cutoff), int(virflag), float(l), _dp(out)); return out


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().jmo_philox4x32_10(c, k, o); return list(o)


def taus2_words(seed, n):
    st = (C.c_uint32 * 3)(); L = lib(); L.jmo_taus2_seed(st, int(seed))
    return np.array([L.jmo_taus2_next(st) for _ in range(n)], dtype=np.uint32)


def colour_of_step(seed, chain_id, sweep_step, ncolours):
    return int(lib().jmo_colour_of_step(int(seed), int(chain_id), int(sweep_step), int(ncolours)))


def colour_halfsweep(r, l, nbn, pot, cutoff, T, maxStep, seed, chain_id, sweep_step, ncolours, colour):
    """In place on r (float64, contiguous). Returns (accepted, dtot[9])."""
    assert r.dtype == np.float64 and r.flags.c_contiguous
    d = np.empty(9)
    n = lib().jmo_colour_halfsweep(_dp(r), r.size, float(l), int(nbn), int(pot), float(cutoff), float(T), float(maxStep),
                                   int(seed), int(chain_id), int(sweep_step), int(ncolours), int(colour), _dp(d))
    return int(n), d


# ---------------------------------------------------------------------------- the compiled reference

def run_reference(deck_text: str, workdir, record_rng: bool = False, threads: int = 1, timeout=None) -> dict:
    """Run oracle/_ref/jmmOneDMC_ref on a deck in `workdir`; returns paths + stdout text."""
    workdir = Path(workdir); workdir.mkdir(parents=True, exist_ok=True)
    (workdir / "INPUT").write_text(deck_text)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    if record_rng:
        env["JMM_RNG_LOG"] = str(workdir / "rng.bin")
    else:
        env.pop("JMM_RNG_LOG", None)
    with open(workdir / "out.out", "w") as f:
        subprocess.run([str(REF_BIN)], cwd=workdir, env=env, stdout=f, stderr=subprocess.STDOUT, check=True, timeout=timeout)
    out = {"dir": workdir, "stdout": (workdir / "out.out").read_text(),
           "thermo": workdir / "thermo.dat.mcs", "config": workdir / "config.dat.mcs"}
    if record_rng:
        out["rng"] = np.fromfile(workdir / "rng.bin", dtype=np.uint32)
    return out
