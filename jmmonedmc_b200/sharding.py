"""Multi-GPU plumbing of the hot path: chains and state points are independent, so the
data path has NO collective.  Rank g owns the chain range [g*C/G, (g+1)*C/G); every random number is keyed
by the GLOBAL chain id, so results do not depend on G.  At the end of a run each rank contributes a
fixed-size summary record per chain and one all_gather (NCCL on GPUs, gloo in the CPU tests) assembles the
table that scripts/Analyze_Mean.py builds from one thermo file per LSF job (scripts/RunJobs.bash:16-27).

No arithmetic of the path lives here (that is libjmmgpu.so); this is bookkeeping."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

RECORD_FIELDS = ("chain", "P", "T", "samples",
                 "rho", "rho2", "L", "L2", "E", "E2", "LE", "Vir", "Vir2", "EVir", "HV", "HV2",   # the 12 sums / samples
                 "E_final", "L_final", "dAcc", "dRej", "vAcc", "vRej")
RECORD_WIDTH = len(RECORD_FIELDS)


def chain_range(rank: int, world: int, total: int) -> tuple[int, int]:
    """Block partition of `total` chains; sizes differ by at most one and cover [0, total) exactly."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad rank/world/total")
    return rank * total // world, (rank + 1) * total // world


def weak_range(rank: int, per_gpu: int) -> tuple[int, int]:
    """Weak scaling (bench.py): every rank holds `per_gpu` chains."""
    return rank * per_gpu, (rank + 1) * per_gpu


def grid_point(chain: int, n_p: int, n_t: int, replicas: int) -> tuple[int, int]:
    """State point (iP, iT) of a global chain id: replicas of a point are consecutive chains (jmm_run)."""
    point = chain // replicas
    return point // n_t, point % n_t


def summary_records(chain_id0: int, P, T, samples: int, accum, totals, l, counters) -> torch.Tensor:
    """[C, RECORD_WIDTH] float64 records from jmm_get_state() arrays (accum = sums over `samples` steps)."""
    accum = np.asarray(accum, dtype=np.float64)
    C = accum.shape[0]
    rec = np.empty((C, RECORD_WIDTH), dtype=np.float64)
    rec[:, 0] = chain_id0 + np.arange(C)
    rec[:, 1] = np.broadcast_to(np.asarray(P, dtype=np.float64), (C,))
    rec[:, 2] = np.broadcast_to(np.asarray(T, dtype=np.float64), (C,))
    rec[:, 3] = samples
    rec[:, 4:16] = accum / max(samples, 1)
    rec[:, 16] = np.asarray(totals)[:, 0]
    rec[:, 17] = np.asarray(l)
    rec[:, 18:22] = np.asarray(counters, dtype=np.float64)
    return torch.from_numpy(rec)


def allgather_summaries(rec: torch.Tensor, sizes: list[int] | None = None) -> torch.Tensor:
    """One all_gather of the per-chain records; returns all chains ordered by global chain id.
    `sizes`: chains per rank when they differ (strong partition); None = equal on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rec
    world = dist.get_world_size()
    if sizes is None:
        out = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(out, rec)
    else:
        m = max(sizes)
        pad = torch.zeros((m, rec.shape[1]), dtype=rec.dtype, device=rec.device)
        pad[: rec.shape[0]] = rec
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad)
        out = [o[:n] for o, n in zip(out, sizes)]
    full = torch.cat(out)
    return full[torch.argsort(full[:, 0])]


def state_point_means(full: torch.Tensor) -> dict:
    """Average the replicas of each (P, T): the per-state-point table of scripts/Analyze_Mean.py."""
    a = full.cpu().numpy()
    out = {}
    for row in a:
        out.setdefault((row[1], row[2]), []).append(row[4:16])
    return {k: np.mean(v, axis=0) for k, v in out.items()}
