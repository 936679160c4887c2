#!/usr/bin/env python
"""bench.py — trial moves/s of the jmmOneDMC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3|c4|c5]

Workload at N=1 (and per GPU at N>1, weak scaling): BASELINE.json configs[1] = the test/INPUTstd deck
(N=10, HARMONIC, NBN 1, NPT, P=0.7, T=0.4, MAXSTEP 0.1, MAXDV 1.0, ENGCHECK 1, DADJ/VADJ 100) replicated
as 4096 independent chains per B200, Philox stream keyed by global chain id.
One bench "step" = one jmm_step() launch advancing every chain by MC_PER_STEP Monte-Carlo steps.

  value     trial moves/s, all ranks, state resident in HBM, CUDA-event timed (max over ranks)
  e2e       the same through the C ABI with HOST buffers: jmm_set_state (H2D) + jmm_step + jmm_get_state
            (D2H) inside the timed region
  roofline  the dominant kernel (k_chains_step) against the fp64 pipe — the binding roof of this
            layout (bytes/trial -> 0) — with the HBM view beside it
  cpu_baseline  the reference's own CPU program (oracle/_ref, compiled from /root/reference) on the
            host cores, one single-threaded process per core (its OpenMP path is racy)

--impl reference times that CPU program alone on the same config and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
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
MC_PER_STEP = 20000
METRIC = "MC trial moves/sec"
UNIT = "trial moves/s"
WORKLOAD = "C2: test/INPUTstd deck (N=10, HARMONIC, NBN 1, NPT, ENGCHECK 1, DADJ/VADJ 100) x 4096 chains per GPU"

# algorithmic work per trial, SURVEY.md §8(d): displacement 10*p+37 flop with p in {1,2} (mean 1.8 over the
# 10 particles), volume trial (fav) 15*9 = 135 flop at 1/11 of the trials, ECheck every step 4*9 = 36 flop
FLOP_PER_TRIAL = (10.0 / 11.0) * (10 * 1.8 + 37) + (1.0 / 11.0) * 135 + 36
BYTES_PER_CHAIN_PER_LAUNCH = 2 * (8 * C2["N"] + 256)          # §8(d): state load + store

# The other BASELINE.json configurations, as secondary workloads (--workload c3|c4|c5); SURVEY.md §8(d) table.
EXTRA = {
    "c3": dict(desc="C3: one chain, N=1,048,576, LJcut 5.0, NBN 4, NLT (L=1.12N), T=0.9, checkerboard half-sweeps",
               kind="sweep", N=1 << 20, nchains=1, pot="LJcut", nbn=4, cutoff=5.0, T=0.9, maxStep=0.12, seed=92847,
               per_step=64, flop=33 * 8 + 37, bytes_per_trial=16.0),
    "c4": dict(desc="C4: RunJobs-style sweep, 65,536 chains (256x256 P,T grid in [0.1,1]) x N=80, LJ, NBN -1, NPT, RELAX",
               kind="chains", N=80, nchains=65536, pot="LJ", nbn=-1, cutoff=math.inf, maxStep=0.1, maxdl=2.0, eci=10000,
               mdai=10 ** 6, mvai=10 ** 6, seed=92847, relax=1, per_step=2000, flop=33 * 79 + 37,
               bytes_per_trial=None),
    "c5": dict(desc="C5: 8 chains x N=262,144, LJ, NBN 64 (128 partners), NLT (L=1.12N), T=0.9, checkerboard half-sweeps",
               kind="sweep", N=1 << 18, nchains=8, pot="LJ", nbn=64, cutoff=math.inf, T=0.9, maxStep=0.12, seed=92847,
               per_step=64, flop=33 * 128 + 37, bytes_per_trial=16.0),
}


def deck_text(numsteps: int, seed: int) -> str:
    """The C2 deck as an INPUT file for the reference binary: print intervals pushed out and the
    (out-of-scope) histograms reduced to one bin so they do not burden the reference."""
    big = 10 ** 12
    return (f"N          10\nP          0.7\nT          0.4\nNUMSTEPS   {numsteps}\nPOT        HARMONIC\nNBN        1\n"
            f"MAXSTEP    0.1\nMAXDV      1.0\nCPI        {big}\nTPI        {big}\nRBW        0.01\nRHONB      1\n"
            f"RHOPI      {big}\nGSW        100\nGNS        1\nGBW        0.01\nGNB        1\nGPI        {big}\n"
            f"SEED       {seed}\nENGCHECK   1\nDADJ       100\nVADJ       100\n")


# ------------------------------------------------------------------------------------------ CPU arm

def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def deck_text_c4(numsteps: int, seed: int) -> str:
    """One chain of the C4 sweep (scripts/RunJobs.bash:40-62 deck, P = T = 0.5) for the reference binary: prints pushed
    out, histograms reduced to one bin."""
    big = 10 ** 12
    return (f"N          80\nP          0.5\nT          0.5\nNUMSTEPS   {numsteps}\nPOT        LJ\nNBN        -1\nRELAX\n"
            f"MAXSTEP    0.1\nMAXDV      2.0\nCPI        {big}\nTPI        {big}\nRBW        0.1\nRHONB      1\n"
            f"RHOPI      {big}\nGSW        200\nGNS        1\nGBW        0.1\nGNB        1\nGPI        {big}\n"
            f"SEED       {seed}\nENGCHECK   10000\nDADJ       1000000\nVADJ       1000000\n")


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
            nhs = {"c3": 80, "c5": 300}.get(workload, 40)
            smp = run_sweep_port_sample(w, nhs, cores)
            one = run_sweep_port_sample(w, max(1, nhs // 4), 1)
            smp["single_core_value"] = one["trials"] / one["seconds"]
        out = {"value": smp["trials"] / smp["seconds"], "unit": UNIT, "cores": smp["cores"], "kind": smp["kind"], "sample": smp["sample"]}
        if "single_core_value" in smp:
            out["single_core_value"] = smp["single_core_value"]
        return out
    except Exception as e:                      # the baseline is reported, never required
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    per_step = int(os.environ.get("JMM_BENCH_REF_STEPS", "400000"))     # ~4 s of CPU work per process per bench step
    for _ in range(args.warmup if args.warmup < 2 else 1):
        run_reference_sample(50_000, cores)
    t, trials, last = 0.0, 0, None
    for _ in range(args.steps):
        last = run_reference_sample(per_step, cores)
        t += last["seconds"]; trials += last["trials"]
    v = trials / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_step": cores, "mc_steps_per_chain_per_step": per_step},
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
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
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


def gpu_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist
    import jmmonedmc_b200 as J
    from jmmonedmc_b200.capi import config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path is the only implementation (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    J.build()

    C = C2["nchains"]
    cfg = config(N=C2["N"], pot=J.POT_HARMONIC, nbn=C2["nbn"], ensemble=J.ENS_NPT, P=C2["P"], T=C2["T"],
                 maxStep=C2["maxStep"], maxdl=C2["maxdl"], eci=C2["eci"], mdai=C2["mdai"], mvai=C2["mvai"], seed=C2["seed"],
                 nchains=C, chain_id0=rank * C, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_DEVICE,
                 device=local)
    h = J.Handle(cfg)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)          # the library launches on torch's stream so torch events time it
    h.start()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        h.step(MC_PER_STEP)
    launches0 = h.kernel_launches
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(1.0)                      # L2 flush between timed iterations, outside the event pair
        a.record(stream)
        h.step(MC_PER_STEP)
        b.record(stream)
        kernel_ms.append(None)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = h.kernel_launches - launches0
    t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    trials_per_rank = C * MC_PER_STEP * args.steps
    value = world * trials_per_rank / (dev_ms * 1e-3)

    # dominant kernel alone: library-side CUDA events around k_chains_step (same stream)
    h.step(MC_PER_STEP)
    k_ms = h.last_kernel_ms
    trials_per_launch = C * MC_PER_STEP

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region
    s = h.get_state()
    r_host, l_host = s["r"], s["l"]
    h2d = r_host.nbytes + l_host.nbytes
    d2h = r_host.nbytes + l_host.nbytes + s["totals"].nbytes + s["accum"].nbytes + s["counters"].nbytes
    for _ in range(2):
        h.set_state(r=r_host, l=l_host); h.step(MC_PER_STEP); s = h.get_state(); r_host, l_host = s["r"], s["l"]
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h.set_state(r=r_host, l=l_host)
        h.step(MC_PER_STEP)
        s = h.get_state()
        r_host, l_host = s["r"], s["l"]
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * trials_per_rank / float(te.item())

    # ---- final summary reduction: one NCCL allgather of the per-chain records
    from jmmonedmc_b200.sharding import allgather_summaries, summary_records
    rec = summary_records(rank * C, C2["P"], C2["T"], h.step_number + 1, s["accum"], s["totals"], s["l"], s["counters"]).cuda()
    rec = allgather_summaries(rec)                     # the only collective of the job (NCCL), SURVEY §8e
    nchains_total = int(rec.shape[0])
    disc = h.echeck_stats()[1]

    # ---- the throughput-bound BASELINE.json configurations (C3, C4, C5), short runs, both arithmetic modes:
    #      carried in the same JSON line as "other_workloads" so that one bench run shows every kernel family
    others = {}
    if not args.no_extras:
        for wl in ("c3", "c4", "c5"):
            for ar in ("fast", "reference"):
                try:
                    ln = measure_extra(wl, ar, False, 3 if wl == "c4" else 10, 3, rank, world, local)
                    if ln:
                        others[f"{wl}_{ar}"] = {"workload": ln["config"]["workload"], "arith": ar, "value": ln["value"], "unit": UNIT,
                                                "ms_per_step": ln["ms_per_step"], "gpu_launches": ln["gpu_launches"],
                                                "fp64_tflops": ln["roofline"]["achieved"], "fp64_frac": ln["roofline"]["frac"],
                                                "flop_per_trial": ln["roofline"]["flop_per_trial"], "acceptance": ln["acceptance"]}
                except Exception as e:                      # a secondary workload never fails the headline line
                    others[f"{wl}_{ar}"] = {"error": str(e)[:200]}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = J.lib().jmm_fp64_peak_tflops(local)
        k_s = k_ms * 1e-3
        achieved_tf = trials_per_launch * FLOP_PER_TRIAL / k_s / 1e12
        achieved_gbs = C * BYTES_PER_CHAIN_PER_LAUNCH / k_s / 1e9
        cpu = None
        try:
            smp = run_reference_sample(int(os.environ.get("JMM_BENCH_CPU_STEPS", "400000")), host_cores())
            cpu = {"value": smp["trials"] / smp["seconds"], "unit": UNIT, "cores": smp["cores"], "kind": smp["kind"],
                   "sample": smp["sample"]}
        except Exception as e:                      # the baseline is reported, never required
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "chains_per_gpu": C, "chains_total": nchains_total,
                       "mc_steps_per_chain_per_step": MC_PER_STEP, "rng": "philox4x32-10", "adapt": "device",
                       "l2": "flushed between timed iterations (256 MiB fill)", "echeck_discrepancies": disc},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "fp64", "kernel": "k_chains_step_bond (bond.cuh: HARMONIC NBN 1, 16 lanes per chain)",
                         "achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved_tf / fp64_peak if fp64_peak and fp64_peak > 0 else None,
                         "peak_source": "DFMA microbenchmark in libjmmgpu (jmm_fp64_peak_tflops), measured in this run; "
                                        "MEASURED_PEAKS.json has no fp64 entry",
                         "flop_per_trial": FLOP_PER_TRIAL, "kernel_ms": k_ms,
                         # dram__bytes_read.sum + dram__bytes_write.sum of this kernel, per launch, from the committed
                         # ncu --set full capture (profiles/r02p_c2_k_chains_step_bond.txt); the 2.75 MB of state
                         # (algorithmic bytes) mostly stay in the 126 MB L2 between launches
                         "traffic": 1240320,
                         "note": "serial Markov chains: latency-bound, see DESIGN.md §roofline",
                         "hbm": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": achieved_gbs / hbm_peak, "peak_source": hbm_src,
                                 "bytes_per_launch": C * BYTES_PER_CHAIN_PER_LAUNCH}},
            "cpu_baseline": cpu,
            "other_workloads": others,
        }
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


def measure_extra(workload: str, arith: str, hist: bool, steps: int, warmup: int, rank: int, world: int, local: int, cpu: bool = False):
    """One secondary workload on an initialised device/process group; returns the JSON line on rank 0 (None elsewhere)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import jmmonedmc_b200 as J
    from jmmonedmc_b200.capi import config

    w = dict(EXTRA[workload])
    if os.environ.get("JMM_BENCH_CHAINS"):          # experiments only: a smaller/larger chain count than the named config
        w["nchains"] = int(os.environ["JMM_BENCH_CHAINS"]); w["desc"] += f" [chains overridden: {w['nchains']}]"
    if os.environ.get("JMM_BENCH_PER_STEP"):
        w["per_step"] = int(os.environ["JMM_BENCH_PER_STEP"])
    pot = {"LJ": J.POT_LJ, "LJcut": J.POT_LJCUT, "HARMONIC": J.POT_HARMONIC}[w["pot"]]
    C, N = w["nchains"], w["N"]
    if w["kind"] == "sweep":
        cfg = config(N=N, pot=pot, nbn=w["nbn"], cutoff=w["cutoff"], ensemble=J.ENS_NLT, L=1.12 * N, T=w["T"],
                     maxStep=w["maxStep"], seed=w["seed"], nchains=C, chain_id0=rank * C, mode=J.MODE_CHECKERBOARD, device=local,
                     arith=J.ARITH_FAST if arith == "fast" else J.ARITH_REFERENCE)
    else:
        cfg = config(N=N, pot=pot, nbn=w["nbn"], cutoff=w["cutoff"], ensemble=J.ENS_NPT, relax=w["relax"], P=0.5, T=0.5,
                     maxStep=w["maxStep"], maxdl=w["maxdl"], eci=w["eci"], mdai=w["mdai"], mvai=w["mvai"], seed=w["seed"],
                     nchains=C, chain_id0=rank * C, rng_kind=J.RNG_PHILOX, mode=J.MODE_RECOMPUTE, adapt=J.ADAPT_DEVICE,
                     device=local, arith=J.ARITH_FAST if arith == "fast" else J.ARITH_REFERENCE)
    h = J.Handle(cfg)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    if w["kind"] == "chains":
        g = np.linspace(0.1, 1.0, 256)
        ids = rank * C + np.arange(C)
        h.set_state(P=g[(ids // 256) % 256], T=g[ids % 256])
        if hist:      # scripts/RunJobs.bash:46-54 histogram geometry: RBW 0.1 x 1000, GSW 200 x 10, GBW 0.1 x 1000
            h.enable_histograms(1000, 0.1, 10, 1000, 200.0, 0.1)
    h.start()
    if w["kind"] == "chains" and not os.environ.get("JMM_BENCH_FROM_ZERO"):
        # the deck runs 1e7 steps per chain; relaxVolume fires every 10 000 steps during the first 1e6 only
        # (src/Main.cpp:173).  The timed steps are taken from the other 90 %: the production phase.
        h.set_step_number(w.get("start_step", 1_000_000))
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")

    def one():
        return h.sweep(w["per_step"]) if w["kind"] == "sweep" else (h.step(w["per_step"]) or C * w["per_step"])

    for _ in range(max(warmup, 3)):
        one()
    torch.cuda.synchronize()
    # a step of the sweep workloads is a fraction of a millisecond: keep warming up (untimed) until the device has
    # been busy for ~50 ms, so that the timed steps do not run on clocks that are still ramping
    t_w = time.perf_counter()
    one()
    torch.cuda.synchronize()
    t_one = max(time.perf_counter() - t_w, 1e-5)
    for _ in range(min(2000, int(0.05 / t_one))):          # back to back, one synchronisation at the end
        one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local); sampler.start()
    l0 = h.kernel_launches
    ev, trials = [], 0
    for _ in range(steps):
        flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); n = one(); b.record(stream)
        trials += n if isinstance(n, int) and n else C * w["per_step"]
        ev.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms, float(trials)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX); ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, trials = float(tm[0]), float(ts[1])
    launches = h.kernel_launches - l0
    one(); k_ms = h.last_kernel_ms
    st = h.get_state(r=False)
    line = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        fp64_peak = J.lib().jmm_fp64_peak_tflops(local)
        value = trials / (ms * 1e-3)
        per_gpu = value / world
        tf = per_gpu * w["flop"] / 1e12
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": max(warmup, 3),
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": {"workload": w["desc"], "per_step": w["per_step"], "arith": arith, "histograms": bool(hist),
                                                "first_step": (0 if (w["kind"] != "chains" or os.environ.get("JMM_BENCH_FROM_ZERO")) else 1_000_000),
                                                "l2": "flushed between timed iterations (256 MiB fill)"},
                "gpu_launches": int(launches), "clocks": clocks, "ms_steps": [round(a.elapsed_time(b), 4) for a, b in ev],
                "roofline": {"bound": "fp64", "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s",
                             "frac": tf / fp64_peak if fp64_peak > 0 else None, "flop_per_trial": w["flop"],
                             "kernel_ms_last_call": k_ms, "traffic": None,
                             "peak_source": "DFMA microbenchmark in libjmmgpu, this run",
                             "hbm": None if not w["bytes_per_trial"] else {
                                 "bound": "hbm", "achieved": per_gpu * w["bytes_per_trial"] / 1e9, "unit": "GB/s",
                                 "peak": float(peaks.get("hbm_gbs", 6650.0)),
                                 "frac": per_gpu * w["bytes_per_trial"] / 1e9 / float(peaks.get("hbm_gbs", 6650.0))}},
                "acceptance": float(st["counters"][:, 0].sum() / max(1, st["counters"][:, :2].sum())),
                "e2e": None, "cpu_baseline": cpu_baseline_extra(workload, w) if cpu else None}
    h.close()
    del flush
    return line


def extra_arm(args) -> None:
    """Secondary workloads (one GPU per rank, weak scaling): same JSON shape, no e2e/cpu legs beyond a note."""
    import torch
    import torch.distributed as dist
    import jmmonedmc_b200 as J

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    J.build()
    line = measure_extra(args.workload, args.arith, args.hist, args.steps, args.warmup, rank, world, local, cpu=not args.no_cpu)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--arith", default="reference", choices=["reference", "fast"], help="c3/c4/c5: JMM_ARITH_*")
    ap.add_argument("--hist", action="store_true", help="c4 only: rho(x)/g(x) histograms with the RunJobs geometry")
    ap.add_argument("--no-extras", action="store_true", help="c2: skip the short C3/C4/C5 runs reported as other_workloads")
    ap.add_argument("--no-cpu", action="store_true", help="c3/c4/c5: skip the CPU side-by-side sample (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload != "c2":
        extra_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
