#!/usr/bin/env python
"""bench.py -- headline benchmark of the ABCdeZ particle hot path on B200.

Metric (BASELINE.json): particle simulate-and-score evaluations per second, and wall time for
abcdesmc! to reach the target eps, beside the host-CPU baseline.

Workload at N=1 (BASELINE.json configs[1]): 10-d correlated Gaussian model, abcdesmc! with 10^6
particles on one B200.  A *step* is one complete abcdesmc! run (prior draws, abcde_init!, and the
SMC loop down to eps_target) over that batch of synthetic input; every step uses a fresh Philox
seed.  `value` = sum(nsims) of the K timed runs / device time, with nothing but scalars leaving
the GPU; `e2e` = the same run through the reference-facing call (abcdez_smc_run with HOST result
buffers: P, Wns, C come back over PCIe inside the timed region); ms_per_step is the
time-to-target-eps.  With --gpus N (torchrun, one process per GPU) the run is ONE sharded population of
N x 10^6 particles (weak scaling: 10^6 particles per GPU): per-iteration exchanges inside the kernels over
NVLink peer memory, global stratified resampling through peer loads.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--particles P]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# ---- config 2 of BASELINE.json, written out exactly (DESIGN.md "Models", gauss_corr10) ---------
D = 10
SIGMA0 = 2.0                                   # prior theta_k ~ N(0, SIGMA0)
RHO = 0.5                                      # Sigma_ij = RHO^|i-j|
Y_OBS = [0.5 * math.sin(1.0 + k) for k in range(D)]
EPS_TARGET = 1.0
PHILOX_KEY = 0xABCDE2 + 1                      # key = 0xABCDE2 + config index
BYTES_PER_EVAL = 32 * D + 33                   # algorithmic bytes per alive particle per MCMC step
METRIC = "particle simulate-and-score evaluations per second (abcdesmc!, time-to-target-eps in ms_per_step)"
UNIT = "evals/s"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def sweep_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the sweep kernel, from the committed ncu --set full
    capture of this same command (profiles/sweep_traffic.json); None when no capture is on record."""
    try:
        with open(os.path.join(ROOT, "profiles", "sweep_traffic.json")) as f:
            return float(json.load(f)["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def workload_spec():
    prior = [("normal", 0.0, SIGMA0)] * D
    data = Y_OBS + [RHO]
    return prior, data


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C + OpenMP restatement of the reference; Julia is not installable here)
# ---------------------------------------------------------------------------------------------
def cpu_sample(nparticles: int, max_iters: int, seed: int, faithful: bool = False):
    from oracle import oracle as O
    O.build()
    prior, data = workload_spec()
    t0 = time.perf_counter()
    r = O.smc_run(prior, "gauss_corr10", data, EPS_TARGET, nparticles=nparticles, nsims_max=10**12, seed=seed,
                  faithful=faithful, max_iters=max_iters, hist_cap=16)
    dt = time.perf_counter() - t0
    return r.nsims, dt, r.sweep_seconds, O.num_threads()


def run_reference(args, rank: int):
    """--impl reference: the reference algorithm's CPU path on the box's host cores.  Each step is a
    bounded sample of the workload: the first `ref_iters` SMC iterations of abcdesmc! at `ref_particles`
    particles (O(1)-partner port: kinder to the CPU than the reference's O(N) wsample scans)."""
    if rank != 0:
        return
    nsims, secs = 0, 0.0
    for s in range(args.warmup):
        cpu_sample(args.ref_particles, 2, PHILOX_KEY + 1000 + s)
    cores = 1
    for s in range(args.steps):
        n, dt, _, cores = cpu_sample(args.ref_particles, args.ref_iters, PHILOX_KEY + s)
        nsims += n; secs += dt
    value = nsims / secs
    sample = (f"{args.steps} x first {args.ref_iters} SMC iterations of abcdesmc! at {args.ref_particles} particles "
              f"(O(1)-partner C/OpenMP port of the reference, incl. prior draws + abcde_init!)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, args.particles),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(args, particles, world=1):
    return {"workload": "BASELINE.json configs[1]: 10-d correlated Gaussian model (gauss_corr10), abcdesmc!, "
                        f"{particles} particles per GPU, eps_target={EPS_TARGET}, defaults alpha=0.95 delta_ess=0.5 Kmcmc=3",
            "particles_per_gpu": particles, "particles": particles * world,
            "parallelism": "single GPU" if world == 1 else f"one population sharded over {world} GPUs (contiguous blocks, "
                           "in-kernel NVLink exchanges, rank-local DE partners)", "d": D, "eps_target": EPS_TARGET, "rho": RHO, "sigma0": SIGMA0,
            "kernel": "IndicatorStrict0to-eps", "step": "one complete abcdesmc! run (init + SMC loop to eps_target), fresh seed per step",
            "l2": "working set ~230 MB of particle state per run exceeds the 126 MB L2; no explicit flush"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--particles", type=int, default=1_000_000)
    ap.add_argument("--ref-particles", type=int, default=0, help="particles of the CPU samples (0: the workload's own --particles)")
    ap.add_argument("--ref-iters", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.ref_particles <= 0:
        args.ref_particles = args.particles
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import abcdez_b200 as A
    stream = torch.cuda.current_stream().cuda_stream
    ctx = A.Context(local_rank, stream=stream)
    if world > 1:
        A.dist.init_sharded(ctx)               # collective: NCCL bootstrap + IPC-mapped mailboxes
    spec, data = workload_spec()
    prior = A.Factored(*[A.host.Normal(0.0, SIGMA0)] * D)
    model = A.Model("gauss_corr10", data)
    N = args.particles                         # per GPU
    Nglobal = N * world
    lo, hi = A.shard_range(Nglobal, rank, world)
    Nloc = hi - lo
    L = A.lib()
    import ctypes as C

    def run(seed, host_out=None, profile=False):
        o = A.host._SmcOpts(); L.abcdez_smc_opts_default(C.byref(o))
        o.nparticles = Nglobal; o.nsims_max = 10**15; o.seed = seed; o.verboseout = 0; o.profile = int(profile); o.sync_every = 4
        r = A.host._SmcResult()
        if host_out is not None:
            r.P, r.Wns, r.C = host_out
        rc = L.abcdez_smc_run(ctx._h, prior.handle(ctx), model.handle(ctx), C.c_double(EPS_TARGET), C.byref(o), C.byref(r))
        if rc:
            raise RuntimeError(L.abcdez_last_error().decode())
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    base = PHILOX_KEY                          # one population: the same seed on every rank
    for s in range(args.warmup):
        run(base + 100000 + s, profile=True)
    # ---- timed region: K complete runs, device-resident results -------------------------------
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nsims = iters = launches = sweeps = 0; sweep_ms = 0.0; logZ = []
    for s in range(args.steps):
        r = run(base + s, profile=True)
        nsims += r.nsims; iters += r.iters; launches += r.n_launches; sweeps += r.n_sweeps; sweep_ms += r.sweep_ms
        logZ.append(r.logZ)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # ---- e2e: the reference-facing call with HOST result buffers ------------------------------
    Ph = torch.empty((Nloc, D), dtype=torch.float64).pin_memory(); Wh = torch.empty(Nloc, dtype=torch.float64).pin_memory()
    Ch = torch.empty(Nloc, dtype=torch.float64).pin_memory()
    host_out = (Ph.data_ptr(), Wh.data_ptr(), Ch.data_ptr())
    run(base + 200000, host_out=host_out)
    barrier()
    t0 = time.perf_counter(); e2e_nsims = 0
    ke = max(1, min(args.steps, 5))
    for s in range(ke):
        r = run(base + 300000 + s, host_out=host_out)
        e2e_nsims += r.nsims
        assert math.isfinite(float(Wh.sum()))           # the host reads the result
    barrier()
    e2e_s = time.perf_counter() - t0
    d2h = Nglobal * (D + 2) * 8                 # all ranks together
    h2d = (len(data) + 4 * D) * 8 + 4 * D               # bound data + prior parameters; the state is born on the device

    # ---- max over ranks / sums ---------------------------------------------------------------
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    # nsims / iters / sweeps are properties of the one sharded population (identical on every rank); launches add up
    c = torch.tensor([launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms_all, e2e_ms_all = t.tolist(); launches_all = c.tolist()[0]
    nsims_all, e2e_nsims_all, iters_all, sweeps_all = nsims, e2e_nsims, iters * world, sweeps * world
    if rank == 0:
        peak, peak_src = peaks()
        value = nsims_all / (ms_all * 1e-3)
        # rank 0's sweep kernel: it simulates its own block, 1/world of the population's evaluations
        ach = BYTES_PER_EVAL * (nsims / world) / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else None
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config_dict(args, N, world),
                "time_to_target_eps_ms": ms_all / args.steps,
                "iters_per_run": iters_all / (args.steps * world), "sweeps_per_run": sweeps_all / (args.steps * world),
                "logZ_mean": float(np.mean(logZ)), "logZ_sd": float(np.std(logZ)),
                "gpu_launches": int(launches_all),
                "clocks": clocks,
                "e2e": {"value": e2e_nsims_all / (e2e_ms_all * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "steps": ke, "ms_per_step": e2e_ms_all / ke},
                "roofline": {"bound": "hbm", "kernel": "smc_sweep_kernel<GaussCorr10>", "achieved": ach, "peak": peak,
                             "unit": "GB/s", "frac": (ach / peak) if ach else None, "traffic": sweep_traffic(),
                             "peak_source": peak_src,
                             "algorithmic_bytes_per_eval": BYTES_PER_EVAL,
                             "algorithmic_bytes_per_launch": BYTES_PER_EVAL * (nsims / world) / max(sweeps, 1),
                             "avg_launch_ms": sweep_ms / max(sweeps, 1),
                             "sweep_share_of_step": sweep_ms / ms if ms > 0 else None}}
        if world == 1 and not args.no_cpu_baseline:
            n, dt, sw, cores = cpu_sample(args.ref_particles, args.ref_iters, PHILOX_KEY)
            line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"first {args.ref_iters} SMC iterations of abcdesmc! at {args.ref_particles} particles, "
                                              "O(1)-partner C/OpenMP port of the reference (oracle/abcdez_oracle.c)"}
            nf, dtf, _, _ = cpu_sample(20000, 2, PHILOX_KEY, faithful=True)
            line["cpu_baseline_faithful"] = {"value": nf / dtf, "unit": UNIT, "cores": cores, "kind": "port",
                                             "sample": "first 2 SMC iterations at 20000 particles with the reference's O(N) "
                                                       "StatsBase.wsample partner scans (O(N^2) per sweep)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
