"""world_size-2 gloo tests (CPU) of the N>1 plumbing: chain ranges, global chain ids, summary allgather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_chain_ranges_cover_exactly():
    from jmmonedmc_b200.sharding import chain_range, grid_point, weak_range
    for total in (1, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            r = [chain_range(g, world, total) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
    assert weak_range(3, 4096) == (12288, 16384)
    assert grid_point(0, 3, 4, 2) == (0, 0) and grid_point(7, 3, 4, 2) == (0, 3) and grid_point(23, 3, 4, 2) == (2, 3)
    with pytest.raises(ValueError):
        chain_range(2, 2, 10)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from jmmonedmc_b200.sharding import allgather_summaries, chain_range, summary_records, RECORD_WIDTH
    c0, c1 = chain_range(rank, world, total)
    C = c1 - c0
    ids = np.arange(c0, c1)
    # synthetic per-chain state that is a pure function of the GLOBAL chain id (as Philox keying guarantees)
    accum = np.stack([ids * 10.0 + k for k in range(12)], axis=1) * 100
    totals = np.stack([ids * 0.5 + k for k in range(9)], axis=1)
    rec = summary_records(c0, 0.7, 0.4 + 0.001 * ids, 100, accum, totals, ids + 5.0, np.stack([ids, ids + 1, ids + 2, ids + 3], axis=1))
    sizes = [chain_range(g, world, total)[1] - chain_range(g, world, total)[0] for g in range(world)]
    full = allgather_summaries(rec, sizes)
    ok = full.shape == (total, RECORD_WIDTH) and bool(torch.all(full[:, 0] == torch.arange(total, dtype=torch.float64)))
    ok = ok and bool(torch.allclose(full[:, 4], torch.arange(total, dtype=torch.float64) * 10.0))     # accum/samples
    ok = ok and bool(torch.allclose(full[:, 17], torch.arange(total, dtype=torch.float64) + 5.0))
    ok = ok and bool(torch.allclose(full[:, 2], 0.4 + 0.001 * torch.arange(total, dtype=torch.float64)))
    q.put((rank, ok, C))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_summary_allgather_world2_gloo(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sum(c for _, _, c in res) == total


def test_state_point_means():
    from jmmonedmc_b200.sharding import state_point_means, RECORD_WIDTH
    full = torch.zeros((4, RECORD_WIDTH), dtype=torch.float64)
    full[:, 0] = torch.arange(4)
    full[:, 1] = torch.tensor([0.5, 0.5, 1.0, 1.0]); full[:, 2] = 0.9
    full[:, 4] = torch.tensor([1.0, 3.0, 10.0, 20.0])
    m = state_point_means(full)
    assert m[(0.5, 0.9)][0] == 2.0 and m[(1.0, 0.9)][0] == 15.0
