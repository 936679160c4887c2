"""Shared by the CPU and GPU tests: deck -> configs for both sides of a parity check."""
import math

import numpy as np


def jmm_config_from_deck(J, d, **kw):
    """d: oracle.parse_deck() dict -> jmmonedmc_b200 Config (the C-ABI struct)."""
    from jmmonedmc_b200.capi import config
    pot = {"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT, "HARMONIC": J.POT_HARMONIC}[d["POT"]]
    base = dict(N=int(d["N"]), pot=pot, nbn=int(d["NBN"]), cutoff=d["CUTOFF"] if d["POT"] != "LJ" else math.inf,
                ensemble={"NPT": J.ENS_NPT, "NLT": J.ENS_NLT}[d["ENSEMBLE"]], relax=int(d.get("RELAX", 0)),
                P=d.get("P", 0.0), T=d["T"], L=d.get("L", 0.0), maxStep=d["MAXSTEP"], maxdl=d["MAXDV"],
                eci=int(d.get("ENGCHECK", 0)), mdai=int(d.get("DADJ", 0)), mvai=int(d.get("VADJ", 0)),
                seed=int(d["SEED"]))
    base.update(kw)
    return config(**base)


def oracle_accept_log(chain, nsteps, with_cadence=True):
    """Run `nsteps` of the oracle one by one and rebuild the accept_log byte the GPU writes."""
    log = np.zeros(nsteps, dtype=np.uint8)
    N = chain.N
    for s in range(nsteps):
        c0 = chain.counters.copy()
        chain.step()
        c1 = chain.counters
        dv = (c1[2] + c1[3]) - (c0[2] + c0[3])
        acc = (c1[0] - c0[0]) + (c1[2] - c0[2])
        log[s] = (1 if acc else 0) | (2 if dv else 0)
        if with_cadence:
            chain.cadence()
    return log


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
