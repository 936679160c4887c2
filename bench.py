#!/usr/bin/env python
"""bench.py — trial moves/s of the jmmOneDMC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4|c5]

N = 1  headline = BASELINE.json configs[1] (C2): the test/INPUTstd deck (N=10, HARMONIC, NBN 1, NPT, P=0.7, T=0.4,
       MAXSTEP 0.1, MAXDV 1.0, ENGCHECK 1, DADJ/VADJ 100) replicated as 4096 independent chains on one B200, Philox
       stream keyed by the global chain id.  The other configurations (C3, C4, C5) ride in the same JSON
       line as `other_workloads`, each with its own e2e and CPU side-by-side; `strong` holds C4 on one GPU, the N = 1
       point of the strong-scaling curve.
N > 1  headline = BASELINE.json configs[3] (C4) AS WRITTEN: the RunJobs-style sweep of 65 536 chains (256 x 256 P,T grid,
       N = 80, LJ, NBN -1, NPT, RELAX) block-partitioned over the N GPUs ("scaling": "strong"), no data-path collective,
       and ONE ncclAllGather of the per-chain summary records (jmm_allgather_summaries, inside the e2e timed region).
       The weak C2 number (4096 chains per GPU) is kept as the secondary field `weak_c2`.

One bench "step" = one jmm_step() / jmm_sweep() call advancing every chain by `mc_steps_per_chain_per_step`.

  value     trial moves/s, all ranks, state resident in HBM, CUDA-event timed on the launching stream (max over ranks)
  e2e       the same through the C ABI with (pinned) HOST buffers: jmm_set_state (H2D) + jmm_step + jmm_get_state (D2H)
            (+ the NCCL allgather at N > 1) inside the timed region
  roofline  the dominant kernel against the fp64 pipe — the binding roof of these layouts (bytes/trial -> 0)
            — with the HBM view beside it
  cpu_baseline  the reference's own CPU program (oracle/_ref, compiled from /root/reference) on the host cores, one
            single-threaded process per core (its OpenMP path is racy); `main_serial` = the legacy
            Main.serial.c (LJ only) timed the same way

--impl reference times that CPU program alone on the same config and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

C2 = dict(N=10, P=0.7, T=0.4, pot="HARMONIC", nbn=1, maxStep=0.1, maxdl=1.0, eci=1, mdai=100, mvai=100, seed=125,
          nchains=4096)
MC_PER_STEP = 400000          # ~145 ms per bench step on the crew kernel: ten timed steps give the clock sampler >= 1 s
METRIC = "MC trial moves/sec"
UNIT = "trial moves/s"
WORKLOAD = "C2: test/INPUTstd deck (N=10, HARMONIC, NBN 1, NPT, ENGCHECK 1, DADJ/VADJ 100) x 4096 chains per GPU"
C4_TOTAL = 65536
WORKLOAD_C4 = ("C4: RunJobs-style sweep (scripts/RunJobs.bash deck: N=80, LJ, NBN -1, NPT, RELAX, ENGCHECK 10000), 65,536 chains = "
               "256x256 P,T grid in [0.1,1], block-partitioned over the GPUs, fast arithmetic (JMM_ARITH_FAST)")

# algorithmic work per trial, SURVEY.md §8(d): displacement 10*p+37 flop with p in {1,2} (mean 1.8 over the
# 10 particles), volume trial (fav) 15*9 = 135 flop at 1/11 of the trials, ECheck every step 4*9 = 36 flop
FLOP_PER_TRIAL = (10.0 / 11.0) * (10 * 1.8 + 37) + (1.0 / 11.0) * 135 + 36
BYTES_PER_CHAIN_PER_LAUNCH = 2 * (8 * C2["N"] + 256)          # §8(d): state load + store

# The other BASELINE.json configurations; SURVEY.md §8(d) table.  e2e_mult: a user call that moves the whole state
# over PCIe both ways advances a long chain by more than 64 half-sweeps; the e2e step of C3/C5 is e2e_mult x per_step.
EXTRA = {
    "c3": dict(desc="C3: one chain, N=1,048,576, LJcut 5.0, NBN 4, NLT (L=1.12N), T=0.9, checkerboard half-sweeps",
               kind="sweep", N=1 << 20, nchains=1, pot="LJcut", nbn=4, cutoff=5.0, T=0.9, maxStep=0.12, seed=92847,
               per_step=64, e2e_mult=16, flop=33 * 8 + 37, bytes_per_trial=16.0),
    "c4": dict(desc="C4: RunJobs-style sweep, 65,536 chains (256x256 P,T grid in [0.1,1]) x N=80, LJ, NBN -1, NPT, RELAX",
               kind="chains", N=80, nchains=C4_TOTAL, pot="LJ", nbn=-1, cutoff=math.inf, maxStep=0.1, maxdl=2.0, eci=10000,
               mdai=10 ** 6, mvai=10 ** 6, seed=92847, relax=1, per_step=20000, e2e_mult=1, flop=33 * 79 + 37,
               bytes_per_trial=None),
    "c5": dict(desc="C5: 8 chains x N=262,144, LJ, NBN 64 (128 partners), NLT (L=1.12N), T=0.9, checkerboard half-sweeps",
               kind="sweep", N=1 << 18, nchains=8, pot="LJ", nbn=64, cutoff=math.inf, T=0.9, maxStep=0.12, seed=92847,
               per_step=64, e2e_mult=16, flop=33 * 128 + 37, bytes_per_trial=16.0),
}


def deck_text(numsteps: int, seed: int) -> str:
    """The C2 deck as an INPUT file for the reference binary: print intervals pushed out and the
    (out-of-scope) histograms reduced to one bin so they do not burden the reference."""
    big = 10 ** 12
    return (f"N          10\nP          0.7\nT          0.4\nNUMSTEPS   {numsteps}\nPOT        HARMONIC\nNBN        1\n"
            f"MAXSTEP    0.1\nMAXDV      1.0\nCPI        {big}\nTPI        {big}\nRBW        0.01\nRHONB      1\n"
            f"RHOPI      {big}\nGSW        100\nGNS        1\nGBW        0.01\nGNB        1\nGPI        {big}\n"
            f"SEED       {seed}\nENGCHECK   1\nDADJ       100\nVADJ       100\n")


def deck_text_c4(numsteps: int, seed: int) -> str:
    """One chain of the C4 sweep (scripts/RunJobs.bash:40-62 deck, P = T = 0.5) for the reference binary: prints pushed
    out, histograms reduced to one bin."""
    big = 10 ** 12
    return (f"N          80\nP          0.5\nT          0.5\nNUMSTEPS   {numsteps}\nPOT        LJ\nNBN        -1\nRELAX\n"
            f"MAXSTEP    0.1\nMAXDV      2.0\nCPI        {big}\nTPI        {big}\nRBW        0.1\nRHONB      1\n"
            f"RHOPI      {big}\nGSW        200\nGNS        1\nGBW        0.1\nGNB        1\nGPI        {big}\n"
            f"SEED       {seed}\nENGCHECK   10000\nDADJ       1000000\nVADJ       1000000\n")


# ------------------------------------------------------------------------------------------ CPU arm

def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_sample(numsteps: int, nproc: int, deck=None, what: str = "C2") -> dict:
    """nproc independent single-threaded copies of the compiled reference, one deck each."""
    deck = deck or deck_text
    from oracle import oracle as O
    if not O.REF_BIN.exists():
        O.build()
    if not O.REF_BIN.exists():
        if what != "C2":
            raise RuntimeError("oracle/_ref/jmmOneDMC_ref is missing")
        return run_port_sample(numsteps, nproc)
    with tempfile.TemporaryDirectory() as tmp:
        procs = []
        env = dict(os.environ, OMP_NUM_THREADS="1")
        env.pop("JMM_RNG_LOG", None)
        for p in range(nproc):
            d = Path(tmp) / f"p{p}"
            d.mkdir()
            (d / "INPUT").write_text(deck(numsteps, 125 + p))
        t0 = time.perf_counter()
        for p in range(nproc):
            procs.append(subprocess.Popen([str(O.REF_BIN)], cwd=Path(tmp) / f"p{p}", env=env,
                                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError(f"reference binary failed: {rcs}")
    return {"seconds": dt, "trials": numsteps * nproc, "kind": "reference", "cores": nproc,
            "sample": f"{nproc} independent single-thread processes of oracle/_ref/jmmOneDMC_ref (reference Main.cpp "
                      f"-O3, OMP_NUM_THREADS=1) x {numsteps} steps of the {what} deck, stdout to /dev/null, set-up included"}


def run_serial_sample(numsteps: int, nproc: int, N: int, P: float, T: float, maxdl: float, what: str) -> dict:
    """north_star's second CPU program: the legacy src/Main.serial.c (v0.0.1, LJ only, 17 positional arguments,
    src/Main.serial.c:541-594; time(NULL) seed, so timing only), one process per core.  Histogram geometry = the
    RunJobs one (scripts/RunJobs.bash:46-54); it aborts in malloc with one-bin histograms."""
    exe = ROOT / "oracle" / "_ref" / "jmmOneDMC_serial"
    if not exe.exists():
        raise RuntimeError("oracle/_ref/jmmOneDMC_serial is missing")
    big = "1000000000000"
    argv = [str(exe), str(N), str(P), str(T), str(numsteps), "lj", "0.1", str(maxdl), big, big, "0.1", "1000", big, "200", "10", "0.1",
            "1000", big]
    with tempfile.TemporaryDirectory() as tmp:
        for p in range(nproc):
            (Path(tmp) / f"p{p}").mkdir()
        t0 = time.perf_counter()
        procs = [subprocess.Popen(argv, cwd=Path(tmp) / f"p{p}", stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) for p in range(nproc)]
        rcs = [p.wait() for p in procs]
        dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError(f"Main.serial failed: {rcs}")
    return {"value": numsteps * nproc / dt, "unit": UNIT, "cores": nproc, "kind": "reference",
            "sample": f"{nproc} processes of oracle/_ref/jmmOneDMC_serial (reference src/Main.serial.c -O3) x {numsteps} steps, "
                      f"{what}, histograms RBW 0.1 x 1000, GNS 10 x GNB 1000 (its argv has no way to switch them off)"}


def run_port_sample(numsteps: int, nproc: int) -> dict:
    """Fallback when the compiled reference did not travel: the C restatement, one process per core."""
    code = ("import sys; sys.path.insert(0, %r)\nfrom oracle import oracle as O\n"
            "d=dict(N=10,POT='HARMONIC',NBN=1,CUTOFF=float('inf'),ENSEMBLE='NPT',P=0.7,T=0.4,MAXSTEP=0.1,MAXDV=1.0,"
            "ENGCHECK=1,DADJ=100,VADJ=100,SEED=int(sys.argv[1]),RELAX=0)\n"
            "c=O.Chain(O.config_from_deck(d,mode=O.MODE_TABLE)); c.start(); c.run(int(sys.argv[2]))\n") % str(ROOT)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, "-c", code, str(125 + p), str(numsteps)]) for p in range(nproc)]
    rcs = [p.wait() for p in procs]
    dt = time.perf_counter() - t0
    if any(rcs):
        raise RuntimeError("oracle port failed")
    return {"seconds": dt, "trials": numsteps * nproc, "kind": "port", "cores": nproc,
            "sample": f"{nproc} processes of the C restatement (oracle/jmm_oracle.c, table mode) x {numsteps} steps, "
                      "interpreter start-up included"}


def run_sweep_port_sample(w: dict, nhs: int, nproc: int) -> dict:
    """C3/C5 on the CPU: the reference cannot allocate these sizes (232 B x N^2/2 of pair tables), so the
    O(N x neighbours) C restatement of the colour half-sweep (oracle/jmm_oracle.c) is timed, one chain per process."""
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\nfrom oracle import oracle as O\n"
            "N, nbn, nhs, cid = %d, %d, %d, int(sys.argv[1]); L = 1.12 * N; ncol = nbn + 1\n"
            "r = ((np.arange(N) + 0.5) / N - 0.5) * L; n = 0\n"
            "for t in range(nhs):\n"
            "    col = O.colour_of_step(%d, cid, t, ncol)\n"
            "    O.colour_halfsweep(r, L, nbn, O.POT[%r], %r, %r, %r, %d, cid, t, ncol, col)\n"
            "    n += len(range(col, N, ncol))\n"
            "print(n)\n") % (str(ROOT), w["N"], w["nbn"], nhs, w["seed"], w["pot"],
                             float(w["cutoff"]) if math.isfinite(w["cutoff"]) else 1e300, w["T"], w["maxStep"], w["seed"])
    from oracle import oracle as O
    O.build()
    t0 = time.perf_counter()
    procs = [subprocess.Popen([sys.executable, "-c", code, str(p)], stdout=subprocess.PIPE, text=True) for p in range(nproc)]
    outs = [p.communicate()[0] for p in procs]
    dt = time.perf_counter() - t0
    if any(p.returncode for p in procs):
        raise RuntimeError("oracle port failed")
    return {"seconds": dt, "trials": sum(int(o.strip().splitlines()[-1]) for o in outs), "kind": "port", "cores": nproc,
            "sample": f"{nproc} process(es) of the C restatement of the colour half-sweep (oracle/jmm_oracle.c, reference "
                      f"arithmetic) x {nhs} half-sweeps of one N={w['N']} chain each, interpreter start-up included; the "
                      "reference itself cannot allocate this N"}


def cpu_baseline_extra(workload: str, w: dict) -> dict:
    """Bounded CPU sample of a secondary workload on the box's host cores (SURVEY §8d side-by-side)."""
    cores = host_cores()
    try:
        if w["kind"] == "chains":
            smp = run_reference_sample(int(os.environ.get("JMM_BENCH_CPU_STEPS", "400000")), cores, deck_text_c4, "C4 (one chain)")
        else:
            nhs = {"c3": 60, "c5": 200}.get(workload, 40)
            smp = run_sweep_port_sample(w, nhs, cores)
            one = run_sweep_port_sample(w, max(1, nhs // 4), 1)
            smp["single_core_value"] = one["trials"] / one["seconds"]
        out = {"value": smp["trials"] / smp["seconds"], "unit": UNIT, "cores": smp["cores"], "kind": smp["kind"], "sample": smp["sample"]}
        if "single_core_value" in smp:
            out["single_core_value"] = smp["single_core_value"]
        if w["kind"] == "chains":
            try:
                out["main_serial"] = run_serial_sample(int(os.environ.get("JMM_BENCH_SERIAL_STEPS", "400000")), cores, 80, 0.5, 0.5, 2.0,
                                                       "one chain of the C4 sweep (N=80, P=T=0.5)")
            except Exception as e:
                out["main_serial"] = {"value": None, "sample": f"failed: {e}"}
        return out
    except Exception as e:                      # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}


def reference_arm(args) -> None:
    """The reference's own CPU program on the host cores, on the config the GPU arm runs at this N (C2 at N = 1, one
    chain of the C4 sweep per core at N > 1); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    strong = args.gpus > 1
    deck, what = (deck_text_c4, "C4 (one chain)") if strong else (deck_text, "C2")
    per_step = int(os.environ.get("JMM_BENCH_REF_STEPS", "100000" if strong else "400000"))     # a few s of CPU work per process per bench step
    for _ in range(args.warmup if args.warmup < 2 else 1):
        run_reference_sample(per_step // 8, cores, deck, what)
    t, trials, last = 0.0, 0, None
    for _ in range(args.steps):
        last = run_reference_sample(per_step, cores, deck, what)
        t += last["seconds"]; trials += last["trials"]
    v = trials / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_C4 if strong else WORKLOAD, "chains_per_step": cores, "mc_steps_per_chain_per_step": per_step},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"], "sample": last["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm

class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 8 and r[2].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) > 8 for n, v in zip(names, r[5:9]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def _peaks():
    try:
        return json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        return {}


def _traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed ncu --set full capture
    of the same bench command (profiles/traffic.json, written by scripts/ncu_summary.py); None when there is none."""
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        e = t.get(kernel)
        return (int(e["dram_bytes"]), e["source"]) if e else (None, None)
    except Exception:
        return None, None


class Env:
    """Process-wide state of one bench rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        import jmmonedmc_b200 as J
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device — the CUDA path is the only implementation (use --impl reference for the CPU arm)")
        self.torch, self.dist, self.J = torch, dist, J
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        J.build()
        self.stream = torch.cuda.current_stream()
        self.flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2
        self._fp64 = None
        self.comm = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def fp64_peak(self) -> float:
        if self._fp64 is None:
            self._fp64 = self.J.lib().jmm_fp64_peak_tflops(self.local)
        return self._fp64

    def nccl_comm(self):
        """The library's own NCCL communicator (jmm_comm_create); the 128-byte id travels over torch.distributed."""
        if self.comm is None and self.world > 1:
            box = [self.J.comm_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(box, src=0)
            self.comm = self.J.Comm(box[0], self.rank, self.world, self.local)
        return self.comm

    def close(self):
        if self.comm is not None:
            self.comm.close()
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(E: Env, one, steps: int, warmup: int, min_seconds: float = 0.0, max_steps: int = 20000):
    """W untimed calls, then `steps` calls (more if needed to reach min_seconds of device time), each between CUDA
    events on the launching stream with an L2 flush before it; barrier + synchronize on both sides.
    Returns (device ms max over ranks, steps taken, trials on this rank, per-step ms, clocks)."""
    torch = E.torch
    for _ in range(max(warmup, 3)):
        one()
    torch.cuda.synchronize()
    t_w = time.perf_counter(); one(); torch.cuda.synchronize()
    t_one = max(time.perf_counter() - t_w, 1e-5)
    if t_one < 0.01:                                   # sub-10-ms steps: keep the device busy for ~50 ms first (clocks)
        for _ in range(min(2000, int(0.05 / t_one))):
            one()
        torch.cuda.synchronize()
    if min_seconds > 0:
        steps = int(E.allmax(float(min(max_steps, max(steps, math.ceil(min_seconds / t_one))))))
    E.barrier()
    sampler = ClockSampler(E.local); sampler.start()
    ev, trials = [], 0
    for _ in range(steps):
        E.flush.fill_(1.0)                             # L2 flush between timed iterations, outside the event pair
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(E.stream); trials += one(); b.record(E.stream)
        ev.append((a, b))
    E.barrier()
    clocks = sampler.stop()
    per = [a.elapsed_time(b) for a, b in ev]
    return E.allmax(sum(per)), steps, trials, per, clocks


def make_chains_handle(E: Env, w: dict, C: int, id0: int, arith: str, adapt=None):
    J = E.J
    from jmmonedmc_b200.capi import config
    pot = {"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT, "HARMONIC": J.POT_HARMONIC}[w["pot"]]
    cfg = config(N=w["N"], pot=pot, nbn=w["nbn"], cutoff=w.get("cutoff", math.inf), ensemble=J.ENS_NPT, relax=w.get("relax", 0),
                 P=w.get("P", 0.5), T=w.get("T", 0.5), maxStep=w["maxStep"], maxdl=w["maxdl"], eci=w["eci"], mdai=w["mdai"], mvai=w["mvai"],
                 seed=w["seed"], nchains=C, chain_id0=id0, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE,
                 adapt=J.ADAPT_DEVICE if adapt is None else adapt, device=E.local,
                 arith=J.ARITH_FAST if arith == "fast" else J.ARITH_REFERENCE)
    h = J.Handle(cfg)
    h.set_stream(E.stream.cuda_stream)          # the library launches on torch's stream so torch events time it
    return h


def e2e_chains(E: Env, h, per_step: int, steps: int, total_chains: int | None = None):
    """`steps` x { jmm_set_state(r, l) from pinned host memory ; jmm_step ; jmm_get_state into pinned host memory
    [; jmm_allgather_summaries] }, wall clock, max over ranks.  Returns (seconds, h2d bytes, d2h bytes per step)."""
    from jmmonedmc_b200.capi import PinnedBuffer
    C, N = h.C, h.N
    import numpy as np
    bufs = {"r": PinnedBuffer((C, N)), "l": PinnedBuffer((C,)), "totals": PinnedBuffer((C, 9)), "accum": PinnedBuffer((C, 12)),
            "counters": PinnedBuffer((C, 4), np.uint64)}
    out = {k: b.array for k, b in bufs.items()}
    h.get_state(out=out)
    comm = E.nccl_comm() if total_chains else None
    h2d = out["r"].nbytes + out["l"].nbytes
    d2h = sum(a.nbytes for a in out.values()) + (total_chains * 24 * 8 if comm else 0)

    def one():
        h.set_state(r=out["r"], l=out["l"])
        h.step(per_step)
        h.get_state(out=out)
        if comm:
            h.allgather_summaries(comm, total_chains)
    for _ in range(2):
        one()
    E.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    E.torch.cuda.synchronize()
    sec = E.allmax(time.perf_counter() - t0)
    return sec, int(h2d), int(d2h)


def measure_c2(E: Env, steps: int, warmup: int, with_e2e: bool = True) -> dict:
    """C2, 4096 chains on this rank (weak: chain ids rank*4096...)."""
    J = E.J
    C = C2["nchains"]
    w = dict(C2, pot="HARMONIC", cutoff=math.inf, relax=0)
    h = make_chains_handle(E, w, C, E.rank * C, "reference")
    h.set_state(P=C2["P"], T=C2["T"])
    h.start()
    l0 = h.kernel_launches
    dev_ms, steps, trials, per, clocks = timed_steps(E, lambda: (h.step(MC_PER_STEP) or C * MC_PER_STEP), steps, warmup)
    launches = h.kernel_launches - l0
    h.step(MC_PER_STEP)
    k_ms = h.last_kernel_ms
    out = {"dev_ms": dev_ms, "steps": steps, "trials_rank": trials, "value": E.allsum(trials) / (dev_ms * 1e-3), "per_step_ms": per,
           "clocks": clocks, "launches": int(launches), "kernel_ms": k_ms, "chains": C, "engine": h.engine}
    if with_e2e:
        sec, h2d, d2h = e2e_chains(E, h, MC_PER_STEP, steps)
        out["e2e"] = {"value": E.world * C * MC_PER_STEP * steps / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h}
    out["disc"] = h.echeck_stats()[1]
    h.close()
    return out


def measure_c4(E: Env, arith: str, steps: int, warmup: int, total: int = C4_TOTAL, strong: bool = False, from_zero: bool = False,
               min_seconds: float = 1.0, with_e2e: bool = True, hist: bool = False, per_step: int | None = None) -> dict:
    """C4: `total` chains of the RunJobs deck; strong = block-partitioned over the ranks, else `total` per rank."""
    import numpy as np
    from jmmonedmc_b200.sharding import chain_range
    w = dict(EXTRA["c4"])
    per_step = per_step or int(os.environ.get("JMM_BENCH_PER_STEP", w["per_step"]))
    if strong:
        c0, c1 = chain_range(E.rank, E.world, total)
    else:
        c0, c1 = E.rank * total, (E.rank + 1) * total
    C = c1 - c0
    h = make_chains_handle(E, w, C, c0, arith)
    g = np.linspace(0.1, 1.0, 256)
    ids = c0 + np.arange(C)
    h.set_state(P=g[(ids // 256) % 256], T=g[ids % 256])
    if hist:      # scripts/RunJobs.bash:46-54 histogram geometry: RBW 0.1 x 1000, GSW 200 x 10, GBW 0.1 x 1000
        h.enable_histograms(1000, 0.1, 10, 1000, 200.0, 0.1)
    h.start()
    if not from_zero:
        # the deck runs 1e7 steps per chain; relaxVolume fires every 10 000 steps during the first 1e6 only
        # (src/Main.cpp:173).  The production phase is the other 90 %.
        h.set_step_number(1_000_000)
    l0 = h.kernel_launches
    dev_ms, steps, trials, per, clocks = timed_steps(E, lambda: (h.step(per_step) or C * per_step), steps, warmup, min_seconds)
    launches = h.kernel_launches - l0
    h.step(per_step)
    k_ms = h.last_kernel_ms
    value = E.allsum(trials) / (dev_ms * 1e-3)
    out = {"dev_ms": dev_ms, "steps": steps, "value": value, "per_step_ms": per, "clocks": clocks, "launches": int(launches),
           "kernel_ms": k_ms, "chains_rank": C, "chains_total": int(E.allsum(C)), "per_step": per_step, "first_step": 0 if from_zero else 1_000_000}
    if with_e2e:
        n_e2e = max(3, min(steps, 10))
        sec, h2d, d2h = e2e_chains(E, h, per_step, n_e2e, total_chains=total if (strong and E.world > 1) else None)
        out["e2e"] = {"value": E.allsum(C) * per_step * n_e2e / sec, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "steps": n_e2e, "allgather_in_timed_region": bool(strong and E.world > 1)}
    st = h.get_state(r=False)
    out["acceptance"] = float(st["counters"][:, 0].sum() / max(1, st["counters"][:, :2].sum()))
    out["disc"] = h.echeck_stats()[1]
    h.close()
    return out


def measure_sweep(E: Env, workload: str, arith: str, steps: int, warmup: int, min_seconds: float = 1.0, with_e2e: bool = True) -> dict:
    """C3 / C5: checkerboard half-sweeps of long chains (weak: the same chains on every rank)."""
    J = E.J
    from jmmonedmc_b200.capi import config, PinnedBuffer
    w = dict(EXTRA[workload])
    if os.environ.get("JMM_BENCH_PER_STEP"):
        w["per_step"] = int(os.environ["JMM_BENCH_PER_STEP"])
    pot = {"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT}[w["pot"]]
    C, N = w["nchains"], w["N"]
    cfg = config(N=N, pot=pot, nbn=w["nbn"], cutoff=w["cutoff"], ensemble=J.ENS_NLT, L=1.12 * N, T=w["T"], maxStep=w["maxStep"],
                 seed=w["seed"], nchains=C, chain_id0=E.rank * C, mode=J.MODE_CHECKERBOARD, device=E.local,
                 arith=J.ARITH_FAST if arith == "fast" else J.ARITH_REFERENCE)
    h = J.Handle(cfg)
    h.set_stream(E.stream.cuda_stream)
    h.start()
    l0 = h.kernel_launches
    dev_ms, steps, trials, per, clocks = timed_steps(E, lambda: h.sweep(w["per_step"]), steps, warmup, min_seconds)
    launches = h.kernel_launches - l0
    h.sweep(w["per_step"])
    k_ms = h.last_kernel_ms
    out = {"dev_ms": dev_ms, "steps": steps, "value": E.allsum(trials) / (dev_ms * 1e-3), "per_step_ms": per, "clocks": clocks,
           "launches": int(launches), "kernel_ms": k_ms, "per_step": w["per_step"]}
    if with_e2e:
        nhs = w["per_step"] * w["e2e_mult"]
        rb = PinnedBuffer((C, N)); tb = PinnedBuffer((C, 9)); ab = PinnedBuffer((C, 12))
        pre = {"r": rb.array, "totals": tb.array, "accum": ab.array}
        h.get_state(l=False, counters=False, out=pre)

        def one():
            h.set_state(r=pre["r"])
            n = h.sweep(nhs)
            h.get_state(l=False, counters=False, out=pre)
            return n
        one()
        n_e2e = max(3, int(math.ceil(0.5 / max(1e-4, dev_ms * 1e-3 / steps * w["e2e_mult"]))))
        n_e2e = min(n_e2e, 200)
        E.barrier()
        t0 = time.perf_counter()
        tr = 0
        for _ in range(n_e2e):
            tr += one()
        E.torch.cuda.synchronize()
        sec = E.allmax(time.perf_counter() - t0)
        out["e2e"] = {"value": E.allsum(tr) / sec, "unit": UNIT, "h2d_bytes_per_step": int(pre["r"].nbytes),
                      "d2h_bytes_per_step": int(sum(a.nbytes for a in pre.values())), "steps": n_e2e, "half_sweeps_per_step": nhs}
    st = h.get_state(r=False)
    out["acceptance"] = float(st["counters"][:, 0].sum() / max(1, st["counters"][:, :2].sum()))
    h.close()
    return out


def roofline_fp64(E: Env, per_gpu_value: float, flop: float, kernel: str, k_ms, bytes_per_trial=None, extra=None, traffic_key=None) -> dict:
    peak = E.fp64_peak()
    tf = per_gpu_value * flop / 1e12
    traffic, tsrc = _traffic(traffic_key or kernel)
    peaks = _peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    r = {"bound": "fp64", "kernel": kernel, "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak if peak and peak > 0 else None,
         "peak_source": "DFMA microbenchmark in libjmmgpu (jmm_fp64_peak_tflops), measured in this run; MEASURED_PEAKS.json has no fp64 "
                        "entry (148 SM x 64 FMA/clk x 1.965 GHz = 37.2 TFLOP/s)",
         "flop_per_trial": flop, "kernel_ms": k_ms, "traffic": traffic, "traffic_source": tsrc}
    if bytes_per_trial:
        gbs = per_gpu_value * bytes_per_trial / 1e9
        r["hbm"] = {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                    "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"}
    if extra:
        r.update(extra)
    return r


def others_block(E: Env) -> dict:
    """C3, C4, C5 in both arithmetic modes, >= 1 s timed each; the fast lines carry e2e and a CPU side-by-side."""
    others = {}
    for wl in ("c3", "c4", "c5"):
        for ar in ("fast", "reference"):
            key = f"{wl}_{ar}"
            try:
                full = ar == "fast"
                if wl == "c4":
                    m = measure_c4(E, ar, 3, 3, with_e2e=full)
                else:
                    m = measure_sweep(E, wl, ar, 10, 3, with_e2e=full)
                w = EXTRA[wl]
                per_gpu = m["value"] / E.world
                peak = E.fp64_peak()
                o = {"workload": w["desc"], "arith": ar, "value": m["value"], "unit": UNIT, "steps": m["steps"],
                     "ms_per_step": m["dev_ms"] / m["steps"], "timed_s": m["dev_ms"] * 1e-3, "gpu_launches": m["launches"],
                     "fp64_tflops": per_gpu * w["flop"] / 1e12, "fp64_frac": per_gpu * w["flop"] / 1e12 / peak if peak > 0 else None,
                     "flop_per_trial": w["flop"], "acceptance": m["acceptance"], "clocks": m["clocks"], "e2e": m.get("e2e")}
                if wl == "c4":
                    o["first_step"] = m["first_step"]
                    if full:
                        z = measure_c4(E, ar, 2, 1, from_zero=True, min_seconds=0.0, with_e2e=False, per_step=10000)
                        o["from_step_zero"] = {"value": z["value"], "fp64_frac": z["value"] / E.world * w["flop"] / 1e12 / peak,
                                               "note": "the first 10^6 steps of the deck: relaxVolume every 10 000 steps (src/Main.cpp:173), "
                                                       "whose O(N^2) Newton iterations are not in the 2644 flop/trial", "steps": z["steps"],
                                               "mc_steps_per_chain_per_step": 10000}
                if full and E.rank == 0:
                    o["cpu_baseline"] = cpu_baseline_extra(wl, w)
                others[key] = o
            except Exception as e:                      # a secondary workload never fails the headline line
                others[key] = {"error": str(e)[:300]}
    return others


def gpu_arm(args) -> None:
    E = Env()
    J = E.J
    if E.world > 1:
        return gpu_arm_strong(E, args)
    m = measure_c2(E, args.steps, args.warmup)
    C = m["chains"]
    others = {} if args.no_extras else others_block(E)
    fp64_peak = E.fp64_peak()
    peaks = _peaks()
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    k_s = m["kernel_ms"] * 1e-3
    achieved_gbs = C * BYTES_PER_CHAIN_PER_LAUNCH / k_s / 1e9
    cpu, serial = None, None
    try:
        smp = run_reference_sample(int(os.environ.get("JMM_BENCH_CPU_STEPS", "400000")), host_cores())
        cpu = {"value": smp["trials"] / smp["seconds"], "unit": UNIT, "cores": smp["cores"], "kind": smp["kind"], "sample": smp["sample"]}
    except Exception as e:                      # the baseline is reported, never required
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    try:                                        # Main.serial.c is LJ only: it cannot run the HARMONIC C2 deck; the nearest deck it can
        serial = run_serial_sample(int(os.environ.get("JMM_BENCH_SERIAL_STEPS", "400000")), host_cores(), 10, 1.0, 0.9, 0.1,
                                   "the INPUT_smalltest shape (N=10, LJ, P=1.0, T=0.9): Main.serial.c is LJ only and cannot run the HARMONIC C2 deck")
    except Exception as e:
        serial = {"value": None, "sample": f"failed: {e}"}
    roof = roofline_fp64(E, m["value"], FLOP_PER_TRIAL, m["engine"], m["kernel_ms"], extra={
        "kernel": m["engine"] + " (HARMONIC NBN 1: solo.cuh / bond.cuh)",
        "note": "serial Markov chains: bound by the latency of one step, see DESIGN.md §3.1",
        "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                "peak_source": "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)",
                "bytes_per_launch": C * BYTES_PER_CHAIN_PER_LAUNCH}})
    strong = None
    if "c4_fast" in others and "value" in others["c4_fast"]:
        o = others["c4_fast"]
        strong = {"workload": WORKLOAD_C4, "n_gpus": 1, "value": o["value"], "e2e": o.get("e2e"), "fp64_frac": o["fp64_frac"],
                  "note": "the N = 1 point of the strong-scaling curve that `bench.py --gpus N` (N > 1) reports as its headline"}
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": 1, "steps": m["steps"], "warmup": max(args.warmup, 3),
        "ms_per_step": m["dev_ms"] / m["steps"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "chains_per_gpu": C, "chains_total": C, "mc_steps_per_chain_per_step": MC_PER_STEP,
                   "rng": "philox4x32-10", "adapt": "device", "l2": "flushed between timed iterations (256 MiB fill)",
                   "echeck_discrepancies": m["disc"]},
        "e2e": m["e2e"], "gpu_launches": m["launches"], "clocks": m["clocks"], "timed_s": m["dev_ms"] * 1e-3,
        "roofline": roof, "cpu_baseline": cpu, "main_serial": serial, "strong": strong, "other_workloads": others,
    }
    print(json.dumps(line), flush=True)
    E.close()


def gpu_arm_strong(E: Env, args) -> None:
    """N > 1: BASELINE.json configs[3] as written — 65 536 chains sharded over the GPUs, strong scaling."""
    # 50 000 MC steps per chain and bench step: 0.07 s (8 GPUs) ... 0.35 s (1 GPU) per step, so that the K timed steps the driver asks
    # for give the clock sampler a region of ~1 s and the host copies of the e2e leg are those of a realistic call
    m = measure_c4(E, "fast", args.steps, args.warmup, strong=True, min_seconds=0.0, per_step=int(os.environ.get("JMM_BENCH_PER_STEP", 50000)))
    z = measure_c4(E, "fast", 2, 1, strong=True, from_zero=True, min_seconds=0.0, with_e2e=False, per_step=10000)
    weak = measure_c2(E, 5, 3, with_e2e=False)
    w = EXTRA["c4"]
    per_gpu = m["value"] / E.world
    if E.rank == 0:
        cpu = cpu_baseline_extra("c4", w)
        roof = roofline_fp64(E, per_gpu, w["flop"], "k_chains_step_lanes", m["kernel_ms"], extra={
            "kernel": f"k_chains_step_lanes (lanes.cuh: {m['chains_rank']} chains per GPU, G lanes per chain, fast arithmetic)",
            "per": "GPU (value / n_gpus x flop_per_trial against one GPU's fp64 peak)"})
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": E.world, "steps": m["steps"], "warmup": max(args.warmup, 3),
            "ms_per_step": m["dev_ms"] / m["steps"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_C4, "chains_total": m["chains_total"], "chains_per_gpu": m["chains_rank"],
                       "mc_steps_per_chain_per_step": m["per_step"], "first_step": m["first_step"], "rng": "philox4x32-10 keyed by the global chain id",
                       "adapt": "device", "arith": "fast", "l2": "flushed between timed iterations (256 MiB fill)",
                       "collective": "one ncclAllGather of 24-double per-chain records per e2e step (jmm_allgather_summaries); none in the data path",
                       "echeck_discrepancies": m["disc"], "acceptance": m["acceptance"]},
            "e2e": m["e2e"], "gpu_launches": m["launches"], "clocks": m["clocks"], "timed_s": m["dev_ms"] * 1e-3,
            "roofline": roof, "cpu_baseline": cpu,
            "from_step_zero": {"value": z["value"], "fp64_frac": z["value"] / E.world * w["flop"] / 1e12 / E.fp64_peak(),
                               "mc_steps_per_chain_per_step": 10000, "steps": z["steps"],
                               "note": "the first 10^6 steps of the deck relax the volume every 10 000 steps (src/Main.cpp:173)"},
            "weak_c2": {"workload": WORKLOAD, "value": weak["value"], "chains_per_gpu": weak["chains"], "scaling": "weak",
                        "ms_per_step": weak["dev_ms"] / weak["steps"], "steps": weak["steps"]},
        }
        print(json.dumps(line), flush=True)
    E.close()


def extra_arm(args) -> None:
    """One secondary workload (C3/C5: the same chains on every rank; C4: 65 536 chains per rank, or sharded with --strong)."""
    E = Env()
    wl, ar = args.workload, args.arith
    w = dict(EXTRA[wl])
    total = int(os.environ.get("JMM_BENCH_CHAINS", w["nchains"]))
    if wl == "c4":
        m = measure_c4(E, ar, args.steps, args.warmup, total=total, strong=args.strong, from_zero=bool(os.environ.get("JMM_BENCH_FROM_ZERO")),
                       min_seconds=args.min_seconds, hist=args.hist, with_e2e=not args.no_e2e)
        kernel = "k_chains_step_lanes" if (ar == "fast" and m["chains_rank"] < 65536) else "k_chains_step_prod_sliced"
    else:
        m = measure_sweep(E, wl, ar, args.steps, args.warmup, min_seconds=args.min_seconds, with_e2e=not args.no_e2e)
        kernel = "k_sweep_fast" if ar == "fast" else "k_sweep"
    if E.rank == 0:
        per_gpu = m["value"] / E.world
        desc = w["desc"] + (f" [chains overridden: {total}]" if total != w["nchains"] else "")
        line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": E.world, "steps": m["steps"], "warmup": max(args.warmup, 3),
                "ms_per_step": m["dev_ms"] / m["steps"], "higher_is_better": True, "scaling": "strong" if (wl == "c4" and args.strong) else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": desc, "per_step": m["per_step"], "arith": ar, "histograms": bool(args.hist),
                           "first_step": m.get("first_step", 0), "l2": "flushed between timed iterations (256 MiB fill)"},
                "gpu_launches": m["launches"], "clocks": m["clocks"], "timed_s": m["dev_ms"] * 1e-3,
                "ms_steps": [round(x, 4) for x in m["per_step_ms"][:64]],
                "roofline": roofline_fp64(E, per_gpu, w["flop"], kernel, m["kernel_ms"], w["bytes_per_trial"],
                                          traffic_key=(kernel + "_" + wl) if kernel == "k_sweep_fast" else None),
                "acceptance": m["acceptance"], "e2e": m.get("e2e"),
                "cpu_baseline": cpu_baseline_extra(wl, w) if not args.no_cpu else None}
        print(json.dumps(line), flush=True)
    E.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--arith", default="reference", choices=["reference", "fast"], help="c3/c4/c5: JMM_ARITH_*")
    ap.add_argument("--hist", action="store_true", help="c4 only: rho(x)/g(x) histograms with the RunJobs geometry")
    ap.add_argument("--strong", action="store_true", help="c4 only: shard the 65 536 chains over the ranks (the N > 1 headline)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="c3/c4/c5: take more steps until this much device time is timed")
    ap.add_argument("--no-extras", action="store_true", help="c2: skip the C3/C4/C5 runs reported as other_workloads")
    ap.add_argument("--no-cpu", action="store_true", help="c3/c4/c5: skip the CPU side-by-side sample (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="c3/c4/c5: skip the host-buffer leg (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload != "c2":
        extra_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
