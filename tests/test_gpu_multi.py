"""Two ranks on two GPUs through the C ABI / jmm_run: the chain partition and the one NCCL collective of a job.
Skipped on a one-GPU box (NCCL refuses two ranks on one device); run with `gpurun --gpus 2`."""
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
RUN = ROOT / "jmmonedmc_b200" / "bin" / "jmm_run"


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_rank_sweep_merges_to_the_single_rank_summary(gold, tmp_path):
    """jmm_run --world 2 (one process per GPU, block partition of 24 chains, Philox keyed by the global chain id) must
    produce, through ONE ncclAllGather, the Summary.dat a single process produces for the same sweep — byte for byte."""
    deck = gold("smalltest_2000")["deck_text"].replace("NUMSTEPS   2000", "NUMSTEPS   3000")
    args = ["--chains", "24", "--sweep-p", "0.5", "1.5", "3", "--sweep-t", "0.6", "1.2", "4"]
    one, two = tmp_path / "one", tmp_path / "two"
    one.mkdir(); two.mkdir()
    (one / "INPUT").write_text(deck); (two / "INPUT").write_text(deck)
    r = subprocess.run([str(RUN), "INPUT"] + args, cwd=one, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    procs = []
    for rank in range(2):
        procs.append(subprocess.Popen([str(RUN), "INPUT", "--rank", str(rank), "--world", "2", "--device", str(rank)] + args, cwd=two,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-1500:] for o in outs)
    assert "one ncclAllGather" in outs[0]
    assert (two / "Summary.dat").read_bytes() == (one / "Summary.dat").read_bytes()
    rows = [(two / f"thermo_chains.rank{k}.dat.mcs").read_text().splitlines() for k in range(2)]
    whole = (one / "thermo_chains.dat.mcs").read_text().splitlines()
    by_chain = lambda lines: sorted(lines[1:], key=lambda l: (int(l.split("\t")[1]), int(l.split("\t")[0])))
    assert by_chain(rows[0]) + by_chain(rows[1]) == by_chain(whole) or sorted(rows[0][1:] + rows[1][1:]) == sorted(whole[1:])


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_rank_allgather_through_the_c_abi(tmp_path):
    """jmm_comm_create + jmm_allgather_summaries from two processes with a ragged partition (37 chains: 18 + 19)."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import jmmonedmc_b200 as J
from jmmonedmc_b200.capi import config
rank, world, total = int(sys.argv[1]), 2, 37
idf = sys.argv[2]
import os, time
if rank == 0:
    uid = J.comm_unique_id(); open(idf + ".tmp", "wb").write(uid); os.rename(idf + ".tmp", idf)
else:
    while not os.path.exists(idf): time.sleep(0.05)
    uid = open(idf, "rb").read()
c0, c1 = rank * total // world, (rank + 1) * total // world
cfg = config(N=10, pot=J.POT_LJ, nbn=-1, ensemble=J.ENS_NPT, P=1.0, T=0.9, maxStep=0.1, maxdl=0.1, eci=1000, mdai=1000, mvai=1000,
             seed=92847, nchains=c1 - c0, chain_id0=c0, device=rank)
with J.Handle(cfg) as h, J.Comm(uid, rank, world, rank) as comm:
    h.start(); h.step(500)
    full = h.allgather_summaries(comm, total)
    mine = h.summaries()
assert np.array_equal(full[c0:c1], mine)
np.save(sys.argv[3], full)
''' % str(ROOT)
    idf = str(tmp_path / "id")
    procs = [subprocess.Popen([os.sys.executable, "-c", code, str(r), idf, str(tmp_path / f"full{r}.npy")],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(o[-2000:] for o in outs)
    a, b = np.load(tmp_path / "full0.npy"), np.load(tmp_path / "full1.npy")
    assert np.array_equal(a, b) and np.array_equal(a[:, 0], np.arange(37))
    # the same 37 chains in one process: identical records (results do not depend on the sharding)
    import jmmonedmc_b200 as J
    from jmmonedmc_b200.capi import config
    cfg = config(N=10, pot=J.POT_LJ, nbn=-1, ensemble=J.ENS_NPT, P=1.0, T=0.9, maxStep=0.1, maxdl=0.1, eci=1000, mdai=1000,
                 mvai=1000, seed=92847, nchains=37, chain_id0=0)
    with J.Handle(cfg) as h:
        h.start(); h.step(500)
        assert np.array_equal(h.summaries(), a)
