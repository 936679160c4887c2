"""ctypes view of include/jmm_gpu.h.  No arithmetic here; every call goes to libjmmgpu.so."""
from __future__ import annotations

import ctypes as C
import math
import re
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "jmm_gpu.h"

POT_LJ, POT_LJCUT, POT_HARMONIC = 0, 1, 2
ENS_NPT, ENS_NLT = 0, 1
RNG_TAUS2, RNG_PHILOX, RNG_RECORDED = 0, 1, 2
MODE_TABLE, MODE_RECOMPUTE, MODE_CHECKERBOARD = 0, 1, 2
ADAPT_HOST, ADAPT_DEVICE, ADAPT_CALLER = 0, 1, 2
ARITH_REFERENCE, ARITH_FAST = 0, 1
FLAG_CONSISTENT_VIRIAL = 1
LOG_ACCEPTED, LOG_VOLUME, LOG_WALL = 1, 2, 4


class JmmError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"jmm status {status}: {msg}")
        self.status = status


class Config(C.Structure):
    _fields_ = [
        ("N", C.c_uint64), ("nbn", C.c_int32), ("pot", C.c_int32), ("cutoff", C.c_double),
        ("ensemble", C.c_int32), ("relax", C.c_int32),
        ("P", C.c_double), ("T", C.c_double), ("L", C.c_double),
        ("maxStep", C.c_double), ("maxdl", C.c_double),
        ("eci", C.c_uint64), ("mdai", C.c_uint64), ("mvai", C.c_uint64),
        ("seed", C.c_uint64), ("nchains", C.c_uint64), ("chain_id0", C.c_uint64),
        ("rng_kind", C.c_int32), ("mode", C.c_int32), ("adapt", C.c_int32), ("device", C.c_int32),
        ("arith", C.c_int32), ("flags", C.c_int32),
    ]

    def copy(self, **kw):
        c = Config.from_buffer_copy(bytes(self))
        for k, v in kw.items():
            setattr(c, k, v)
        return c


class Deck(C.Structure):
    _fields_ = [
        ("numsteps", C.c_uint64), ("cpi", C.c_uint64), ("tpi", C.c_uint64), ("gpi", C.c_uint64),
        ("rhopi", C.c_uint64), ("gnb", C.c_uint64), ("rhonb", C.c_uint64),
        ("rbw", C.c_double), ("gsw", C.c_double), ("gbw", C.c_double),
        ("gns", C.c_int32), ("is_restart", C.c_int32), ("n_unknown", C.c_int32),
        ("pot_str", C.c_char * 80), ("ensemble_str", C.c_char * 80),
    ]


def lib_path() -> Path:
    # JMM_LIBJMMGPU: load another build of the same library (kernel tuning experiments)
    import os
    return Path(os.environ["JMM_LIBJMMGPU"]) if os.environ.get("JMM_LIBJMMGPU") else PKG / "libjmmgpu.so"


def declared_symbols() -> list[str]:
    """Every function include/jmm_gpu.h declares."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(jmm_[a-z_0-9]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    """Load libjmmgpu.so; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    p = lib_path()
    if not p.exists():
        raise JmmError(-2, f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the CUDA library is the only implementation)")
    L = C.CDLL(str(p))
    dp, u64p, u32p, u8p = (C.POINTER(t) for t in (C.c_double, C.c_uint64, C.c_uint32, C.c_uint8))
    H = C.c_void_p
    sig = {
        "jmm_read_input": (C.c_int32, [C.c_char_p, C.POINTER(Config), C.POINTER(Deck)]),
        "jmm_create": (C.c_int32, [C.POINTER(Config), C.POINTER(H)]),
        "jmm_destroy": (C.c_int32, [H]),
        "jmm_set_state": (C.c_int32, [H, dp, dp, dp, dp]),
        "jmm_set_step_sizes": (C.c_int32, [H, dp, dp]),
        "jmm_get_step_sizes": (C.c_int32, [H, dp, dp]),
        "jmm_start": (C.c_int32, [H]),
        "jmm_start_parts": (C.c_int32, [H, C.c_int32, C.c_int32, C.c_int32]),
        "jmm_energy": (C.c_int32, [H, dp, C.c_int32]),
        "jmm_step": (C.c_int32, [H, C.c_uint64, u32p, C.c_uint64, u8p]),
        "jmm_relax_volume": (C.c_int32, [H]),
        "jmm_adjust_step_sizes": (C.c_int32, [H, C.c_int32, C.c_int32]),
        "jmm_get_state": (C.c_int32, [H, dp, dp, dp, dp, u64p]),
        "jmm_zero_accum": (C.c_int32, [H]),
        "jmm_set_accum": (C.c_int32, [H, dp, C.c_uint64]),
        "jmm_step_number": (C.c_uint64, [H]),
        "jmm_set_step_number": (C.c_int32, [H, C.c_uint64]),
        "jmm_echeck_stats": (C.c_int32, [H, u64p, u64p]),
        "jmm_stream_cursor": (C.c_uint64, [H]),
        "jmm_sweep": (C.c_int32, [H, C.c_uint64, u64p]),
        "jmm_enable_histograms": (C.c_int32, [H, C.c_uint64, C.c_double, C.c_int32, C.c_uint64, C.c_double, C.c_double]),
        "jmm_take_histograms": (C.c_int32, [H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "jmm_checkpoint_save": (C.c_int32, [H, C.c_char_p]),
        "jmm_checkpoint_load": (C.c_int32, [H, C.c_char_p]),
        "jmm_kernel_launches": (C.c_uint64, [H]),
        "jmm_last_kernel_ms": (C.c_double, [H]),
        "jmm_engine": (C.c_char_p, [H]),
        "jmm_set_stream": (C.c_int32, [H, C.c_void_p]),
        "jmm_host_alloc": (C.c_void_p, [C.c_uint64]),
        "jmm_host_free": (None, [C.c_void_p]),
        "jmm_fp64_peak_tflops": (C.c_double, [C.c_int32]),
        "jmm_last_error": (C.c_char_p, []),
        "jmm_version": (C.c_char_p, []),
        "jmm_rng_selftest": (C.c_int32, [u32p, u32p, u32p, C.c_uint64, u32p, C.c_uint32, C.c_int32]),
        "jmm_accept_selftest": (C.c_int32, [C.c_uint64, C.c_uint64, u64p, C.c_int32]),
        "jmm_comm_unique_id": (C.c_int32, [u8p]),
        "jmm_comm_create": (C.c_int32, [u8p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(H)]),
        "jmm_comm_destroy": (C.c_int32, [H]),
        "jmm_summaries": (C.c_int32, [H, dp]),
        "jmm_allgather_summaries": (C.c_int32, [H, H, C.c_uint64, dp]),
        "jmm_nccl_version": (C.c_int32, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(st):
    if st != 0:
        raise JmmError(st, lib().jmm_last_error().decode())


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = np.ascontiguousarray(np.broadcast_to(a, shape))
    return a


def read_input(path) -> tuple[Config, Deck]:
    cfg, deck = Config(), Deck()
    _check(lib().jmm_read_input(str(path).encode(), C.byref(cfg), C.byref(deck)))
    return cfg, deck


def accept_selftest(n, seed=1, device=0):
    """(Metropolis mismatches, volume mismatches, Metropolis cases decided exactly, volume cases decided exactly)."""
    out = np.zeros(4, dtype=np.uint64)
    _check(lib().jmm_accept_selftest(int(n), int(seed), out.ctypes.data_as(C.POINTER(C.c_uint64)), int(device)))
    return tuple(int(x) for x in out)


def rng_selftest(ctr, key, seed, n, device=0):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    t = np.zeros(max(n, 1), dtype=np.uint32)
    _check(lib().jmm_rng_selftest(c, k, o, int(seed), t.ctypes.data_as(C.POINTER(C.c_uint32)), int(n), int(device)))
    return list(o), t[:n]


SUMMARY_DOUBLES, COMM_ID_BYTES = 24, 128
SUMMARY_FIELDS = ("chain", "P", "T", "N", "samples", "rho", "rho2", "L", "L2", "E", "E2", "LE", "Vir", "Vir2", "EVir", "HV", "HV2",
                  "dAcc", "dRej", "vAcc", "vRej", "L_final", "E_final", "Vir_final")


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (rank 0); hand the 128 bytes to the other ranks."""
    buf = (C.c_uint8 * COMM_ID_BYTES)()
    _check(lib().jmm_comm_unique_id(buf))
    return bytes(buf)


class Comm:
    """RAII wrapper of jmm_comm (one NCCL communicator; collective constructor)."""

    def __init__(self, uid: bytes, rank: int, world: int, device: int = 0):
        self.L = lib()
        self.c = C.c_void_p()
        buf = (C.c_uint8 * COMM_ID_BYTES).from_buffer_copy(uid)
        _check(self.L.jmm_comm_create(buf, int(rank), int(world), int(device), C.byref(self.c)))
        self.rank, self.world = int(rank), int(world)

    def close(self):
        if getattr(self, "c", None) is not None and self.c:
            self.L.jmm_comm_destroy(self.c)
            self.c = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Handle:
    """RAII wrapper of jmm_handle; array arguments/results are chain-major numpy float64."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self.L = lib()
        self.h = C.c_void_p()
        _check(self.L.jmm_create(C.byref(cfg), C.byref(self.h)))
        self.C = int(cfg.nchains)
        self.N = int(cfg.N)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.jmm_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_state(self, r=None, l=None, P=None, T=None):
        r = _f64(r, (self.C, self.N)); l = _f64(l, (self.C,)); P = _f64(P, (self.C,)); T = _f64(T, (self.C,))
        _check(self.L.jmm_set_state(self.h, _dp(r), _dp(l), _dp(P), _dp(T)))

    def set_step_sizes(self, maxStep=None, maxdl=None):
        a = _f64(maxStep, (self.C,)); b = _f64(maxdl, (self.C,))
        _check(self.L.jmm_set_step_sizes(self.h, _dp(a), _dp(b)))

    def get_step_sizes(self):
        a = np.empty(self.C); b = np.empty(self.C)
        _check(self.L.jmm_get_step_sizes(self.h, _dp(a), _dp(b)))
        return a, b

    def start(self):
        _check(self.L.jmm_start(self.h))

    def start_parts(self, fad=True, relax=True, thermo=True):
        _check(self.L.jmm_start_parts(self.h, int(fad), int(relax), int(thermo)))

    def energy(self, exact_order=False):
        t = np.empty((self.C, 9))
        _check(self.L.jmm_energy(self.h, _dp(t), 1 if exact_order else 0))
        return t

    def step(self, nsteps, rng_stream=None, accept_log=False):
        log = np.zeros((int(nsteps), self.C), dtype=np.uint8) if accept_log else None
        sp, sn = None, 0
        if rng_stream is not None:
            self._stream = np.ascontiguousarray(rng_stream, dtype=np.uint32)
            sp, sn = self._stream.ctypes.data_as(C.POINTER(C.c_uint32)), self._stream.size
        lp = log.ctypes.data_as(C.POINTER(C.c_uint8)) if log is not None else None
        _check(self.L.jmm_step(self.h, int(nsteps), sp, sn, lp))
        return log

    def relax_volume(self):
        _check(self.L.jmm_relax_volume(self.h))

    def adjust_step_sizes(self, dis=True, vol=True):
        _check(self.L.jmm_adjust_step_sizes(self.h, int(dis), int(vol)))

    def get_state(self, r=True, l=True, totals=True, accum=True, counters=True, out=None):
        """out: optional dict of preallocated arrays (e.g. pinned, see pinned_empty) to fill instead of new ones."""
        pre = out or {}
        out = {}
        R = pre.get("r", np.empty((self.C, self.N)) if r else None) if r else None
        Lb = pre.get("l", np.empty(self.C) if l else None) if l else None
        T = pre.get("totals", np.empty((self.C, 9)) if totals else None) if totals else None
        A = pre.get("accum", np.empty((self.C, 12)) if accum else None) if accum else None
        Cn = pre.get("counters", np.empty((self.C, 4), dtype=np.uint64) if counters else None) if counters else None
        _check(self.L.jmm_get_state(self.h, _dp(R), _dp(Lb), _dp(T), _dp(A),
                                    Cn.ctypes.data_as(C.POINTER(C.c_uint64)) if counters else None))
        out.update(r=R, l=Lb, totals=T, accum=A, counters=Cn)
        return out

    def zero_accum(self):
        _check(self.L.jmm_zero_accum(self.h))

    def set_accum(self, accum, samples):
        a = _f64(accum, (self.C, 12))
        _check(self.L.jmm_set_accum(self.h, _dp(a), int(samples)))

    @property
    def step_number(self):
        return int(self.L.jmm_step_number(self.h))

    def set_step_number(self, sn):
        _check(self.L.jmm_set_step_number(self.h, int(sn)))

    @property
    def stream_cursor(self):
        return int(self.L.jmm_stream_cursor(self.h))

    def echeck_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        _check(self.L.jmm_echeck_stats(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def enable_histograms(self, rhonb, rbw, gns, gnb, gsw, gbw):
        _check(self.L.jmm_enable_histograms(self.h, int(rhonb), float(rbw), int(gns), int(gnb), float(gsw), float(gbw)))
        self._hist = (int(rhonb), int(gns), int(gnb))

    def take_histograms(self, rho=True, g=True):
        rhonb, gns, gnb = self._hist
        a = np.zeros((self.C, rhonb), dtype=np.int64) if rho else None
        b = np.zeros((self.C, gns, gnb), dtype=np.int64) if g else None
        p = lambda x: x.ctypes.data_as(C.POINTER(C.c_int64)) if x is not None else None
        _check(self.L.jmm_take_histograms(self.h, p(a), p(b)))
        return a, b

    def checkpoint_save(self, path):
        _check(self.L.jmm_checkpoint_save(self.h, str(path).encode()))

    def checkpoint_load(self, path):
        _check(self.L.jmm_checkpoint_load(self.h, str(path).encode()))

    def sweep(self, n_halfsweeps):
        t = C.c_uint64()
        _check(self.L.jmm_sweep(self.h, int(n_halfsweeps), C.byref(t)))
        return int(t.value)

    @property
    def kernel_launches(self):
        return int(self.L.jmm_kernel_launches(self.h))

    @property
    def last_kernel_ms(self):
        return float(self.L.jmm_last_kernel_ms(self.h))

    @property
    def engine(self) -> str:
        """Name of the kernel family jmm_step / jmm_sweep launches for this handle (jmm_engine)."""
        return self.L.jmm_engine(self.h).decode()

    def set_stream(self, cuda_stream_ptr):
        _check(self.L.jmm_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def summaries(self):
        """[nchains][SUMMARY_DOUBLES] per-chain records (SUMMARY_FIELDS), packed on the device."""
        out = np.empty((self.C, SUMMARY_DOUBLES))
        _check(self.L.jmm_summaries(self.h, _dp(out)))
        return out

    def allgather_summaries(self, comm: "Comm", total_chains: int):
        """Collective: ONE ncclAllGather of every rank's records; [total_chains][SUMMARY_DOUBLES] in global chain order."""
        out = np.empty((int(total_chains), SUMMARY_DOUBLES))
        _check(self.L.jmm_allgather_summaries(self.h, comm.c, int(total_chains), _dp(out)))
        return out


class PinnedBuffer:
    """cudaHostAlloc'd memory (jmm_host_alloc) viewed as a numpy array; freed with the object."""

    def __init__(self, shape, dtype=np.float64):
        self.L = lib()
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = self.L.jmm_host_alloc(max(n, 1))
        if not self.ptr:
            raise JmmError(-2, "jmm_host_alloc failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        if getattr(self, "ptr", None):
            self.array = None
            self.L.jmm_host_free(self.ptr)
            self.ptr = None


def config(N, pot, nbn=-1, cutoff=math.inf, ensemble=ENS_NPT, relax=0, P=0.0, T=1.0, L=0.0, maxStep=0.1,
           maxdl=0.1, eci=0, mdai=0, mvai=0, seed=1, nchains=1, chain_id0=0, rng_kind=RNG_PHILOX,
           mode=MODE_RECOMPUTE, adapt=ADAPT_DEVICE, device=0, arith=ARITH_REFERENCE, flags=0) -> Config:
    c = Config()
    c.N, c.nbn, c.pot, c.cutoff, c.ensemble, c.relax = N, nbn, pot, cutoff, ensemble, relax
    c.P, c.T, c.L, c.maxStep, c.maxdl = P, T, L, maxStep, maxdl
    c.eci, c.mdai, c.mvai, c.seed, c.nchains, c.chain_id0 = eci, mdai, mvai, seed, nchains, chain_id0
    c.rng_kind, c.mode, c.adapt, c.device = rng_kind, mode, adapt, device
    c.arith = arith
    c.flags = flags
    return c
