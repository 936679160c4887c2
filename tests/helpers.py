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


def totals_close(a, b, rtol=1e-12):
    """|a-b| <= rtol * gross magnitude of the summed terms.  E, Vir and HV are differences of the large
    positive 12- and 6-sums, so their error is measured against E12+E6 etc. (north_star's 1e-12)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    g = np.abs(b).copy()
    g[..., 0] = np.abs(b[..., 2]) + np.abs(b[..., 4]) + np.abs(b[..., 0])
    g[..., 1] = np.abs(b[..., 3]) + np.abs(b[..., 5]) + np.abs(b[..., 1])
    g[..., 6] = np.abs(b[..., 7]) + np.abs(b[..., 8]) + np.abs(b[..., 6])
    return bool(np.all(np.abs(a - b) <= rtol * g + 1e-300))


def exact_totals(r, nbn, pot_name, cutoff, l):
    """The nine configuration totals with every pair term computed as the reference computes it
    (same fp64 expression per term, numpy elementwise) but summed EXACTLY (math.fsum): the yardstick for
    sums whose order of addition is free.  (A serial left-to-right sum over ~1e5 lattice terms is itself
    ~1e-12 away from this value.)"""
    import math
    r = np.asarray(r, dtype=np.float64)
    N = r.size
    out = np.zeros(9)
    terms = [[] for _ in range(9)]
    kmax = N - 1 if nbn < 0 else min(nbn, N - 1)
    for k in range(1, kmax + 1):
        d = r[k:] - r[:-k]
        if pot_name == "HARMONIC":
            e = np.where(d <= 0, 1e11, np.where(d < cutoff, (d - 1.0) * (d - 1.0), 0.0))
            v = np.where(d <= 0, 1e11, np.where(d < cutoff, (2 / l) * d * (d - 1.0), 0.0))
            terms[0].append(e); terms[1].append(v)
            continue
        r3 = d * d * d
        r6 = 1 / (r3 * r3)
        r12 = r6 * r6
        inside = np.ones_like(d, dtype=bool) if pot_name == "LJ" else (d <= cutoff)
        z = lambda x: np.where(inside, x, 0.0)
        p6, p12, v6, v12, h6, h12 = 4 * r6, 4 * r12, 24 * r6, 48 * r12, 144 * r6, 576 * r12
        for idx, val in ((0, p12 - p6), (1, v12 - v6), (2, p12), (3, v12), (4, p6), (5, v6), (6, h12 - h6), (7, h12), (8, h6)):
            terms[idx].append(z(val))
    for k in range(9):
        if terms[k]:
            out[k] = math.fsum(np.concatenate(terms[k]).tolist())
    return out
