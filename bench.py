#!/usr/bin/env python
"""bench.py -- headline benchmark of the ABCdeZ particle hot path on B200.

Metric (BASELINE.json): particle simulate-and-score evaluations per second, and wall time for
abcdesmc! to reach the target eps, beside the host-CPU baseline.

Default workload (--config 2 = BASELINE.json configs[1]): 10-d correlated Gaussian model, abcdesmc! with 10^6
particles on one B200.  A *step* is one complete abcdesmc! run (prior draws, abcde_init!, and the SMC loop down
to eps_target) over that batch of synthetic input; every step uses a fresh Philox seed.  `value` = sum(nsims) of
the K timed runs / device time, with nothing but scalars leaving the GPU; `e2e` = the same run through the
reference-facing call (abcdez_smc_run with HOST result buffers: P, Wns, C come back over PCIe inside the timed
region); ms_per_step is the time-to-target-eps.  With --gpus N (torchrun, one process per GPU) the run is ONE
sharded population of N x 10^6 particles (weak scaling) or, with --particles-total T, of T particles split over
the N GPUs (strong scaling): per-iteration exchanges inside the kernels over NVLink peer memory, global stratified
resampling through peer loads.  Before anything is timed at N > 1 a small sharded run is
compared with the CPU oracle's island variant (fail loudly).

The other BASELINE.json configurations run with --config 3 | 4 | 5 (g-and-k, Lotka-Volterra, birth-death); configs
3 and 5 are bounded samples of their workload by default (--max-iters), see CONFIGS below.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C] [--particles P]
                  [--particles-total T] [--max-iters I] [--mc]
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

PHILOX_KEY0 = 0xABCDE2                          # key = 0xABCDE2 + config index
METRIC = "particle simulate-and-score evaluations per second (abcdesmc!, time-to-target-eps in ms_per_step)"
UNIT = "evals/s"
FP64_PEAK_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12   # nominal: 148 SMs x 64 FP64 FMA lanes x 2 x 1.965 GHz = 37.2 (no measured figure on record)
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12

# ---- the BASELINE.json configurations, written out exactly (DESIGN.md "Models") ------------------------------
D2 = 10
GK_OCTILES = [2.39384, 2.569082, 2.748052, 3.0, 3.4169, 4.196232, 5.900654]      # exact octiles of g-and-k(A=3, B=1, g=2, k=0.5)
LV_OBS = [1.4385, 0.5655, 1.9586, 0.7589, 2.3239, 1.2022, 2.0925, 1.9724, 1.3201, 2.6696, 0.6896, 2.7822, 0.3821, 2.4679,
          0.2525, 2.0363]                                                       # RK4 trajectory of theta* = (1.2, 0.9, 0.7, 0.6), 8 observations
CONFIGS = {
    2: dict(name="BASELINE.json configs[1]: 10-d correlated Gaussian model (gauss_corr10), abcdesmc!",
            model="gauss_corr10", prior=[("normal", 0.0, 2.0)] * D2, data=[0.5 * math.sin(1.0 + k) for k in range(D2)] + [0.5],
            eps_target=1.0, particles=1_000_000, max_iters=0, bound="hbm", bytes_per_eval=32 * D2 + 33, dtype="f64"),
    3: dict(name="BASELINE.json configs[2]: g-and-k distribution (4 params, 10^4-draw octile summaries, FP64), abcdesmc!; "
                 "1.25e6 particles per GPU = the per-GPU share of 10^7 on 8 GPUs",
            model="gk", prior=[("uniform", 0.0, 10.0)] * 4, data=[10000.0] + GK_OCTILES,
            eps_target=0.05, particles=1_250_000, max_iters=3, bound="fp64", flops_per_eval=53.0 * 10000, definition_flops_per_eval=151.0 * 10000, dtype="f64"),
    4: dict(name="BASELINE.json configs[3]: Lotka-Volterra ODE, fixed-step RK4 (400 steps, 8 noisy observations), 4 params, "
                 "abcdesmc! (and abcdemc! with --mc); model comparison against lotka_volterra_lin in tests/ and profiles/",
            model="lotka_volterra", prior=[("uniform", 0.0, 2.0)] * 4, data=[1.0, 0.5, 0.01, 50, 8, 0.05] + LV_OBS,
            eps_target=0.12, particles=1_000_000, max_iters=0, bound="fp64", flops_per_eval=27100.0, dtype="f64"),
    5: dict(name="BASELINE.json configs[4]: linear birth-death Gillespie SSA (divergent trajectory lengths), 2 params, abcdesmc!; "
                 "1.25e7 particles per GPU = the per-GPU share of 10^8 on 8 GPUs",
            model="birth_death", prior=[("uniform", 0.0, 2.0)] * 2, data=[20.0, 8, 0.5, 5000.0, 22, 25, 24, 30, 33, 31, 36, 40],
            eps_target=1.5, particles=12_500_000, max_iters=3, bound="fp64", flops_per_event=37.0, dtype="f64"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def sweep_traffic():
    """dram__bytes_read + dram__bytes_write per launch of the config-2 sweep kernel, from the committed ncu --set full
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


def config_dict(cfg, args, particles_per_gpu, total, world, scaling):
    return {"workload": f"{cfg['name']}, {particles_per_gpu} particles per GPU, eps_target={cfg['eps_target']}, "
                        "defaults alpha=0.95 delta_ess=0.5 Kmcmc=3",
            "config": args.config, "particles_per_gpu": particles_per_gpu, "particles": total,
            "parallelism": "single GPU" if world == 1 else f"one population sharded over {world} GPUs (contiguous blocks, "
                           f"in-kernel NVLink exchanges, rank-local DE partners), {scaling} scaling",
            "d": len(cfg["prior"]), "eps_target": cfg["eps_target"], "model": cfg["model"],
            "kernel": "IndicatorStrict0to-eps",
            "relaxed_modes": [m for m, on in (("partner_segments", args.segments), ("systematic_resampling", args.systematic), ("fp32_state", args.fp32_state)) if on],
            "step": ("one complete abcdesmc! run (init + SMC loop to eps_target), fresh seed per step" if not args.max_iters else
                     f"init + the first {args.max_iters} SMC iterations of abcdesmc! (bounded sample of the run), fresh seed per step"),
            "l2": "particle state per run exceeds the 126 MB L2 (>= 218 B per particle); no explicit flush"}


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (C + OpenMP restatement of the reference; Julia is not installable here)
# ---------------------------------------------------------------------------------------------
def cpu_threads():
    """All host cores: torchrun exports OMP_NUM_THREADS=1, which must not throttle the CPU arm."""
    from oracle import oracle as O
    O.build()
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    O.set_threads(n)
    return O, n


def cpu_sample(cfg, nparticles: int, max_iters: int, seed: int, faithful: bool = False):
    O, _ = cpu_threads()
    t0 = time.perf_counter()
    r = O.smc_run(cfg["prior"], cfg["model"], cfg["data"], cfg["eps_target"], nparticles=nparticles, nsims_max=10**12, seed=seed,
                  faithful=faithful, max_iters=max_iters, hist_cap=4096)
    dt = time.perf_counter() - t0
    return r, dt, O.num_threads()


def run_reference(args, cfg, rank: int, world: int):
    """--impl reference: the reference algorithm's CPU path on the box's host cores, all of them, on the SAME
    configuration (world x particles-per-GPU particles).  Each step is a bounded sample of the workload: the first
    `ref_iters` SMC iterations of abcdesmc! (O(1)-partner port: kinder to the CPU than the reference's O(N) wsample
    scans), fewer iterations at larger populations so that a step stays a few seconds."""
    if rank != 0:
        return
    total = args.particles_total if args.particles_total else args.particles * world
    iters = max(1, args.ref_iters // world) if not args.particles_total else args.ref_iters
    nsims, secs, sweeps_s = 0, 0.0, 0.0
    for s in range(min(args.warmup, 1)):
        cpu_sample(cfg, min(total, 200000), 2, PHILOX_KEY0 + 1000 + s)
    cores = 1
    for s in range(args.steps):
        r, dt, cores = cpu_sample(cfg, total, iters, PHILOX_KEY0 + args.config - 1 + s)
        nsims += r.nsims; secs += dt; sweeps_s += r.sweep_seconds
    value = nsims / secs
    sample = (f"{args.steps} x first {iters} SMC iterations of abcdesmc! at {total} particles (O(1)-partner C/OpenMP port of the "
              f"reference, incl. prior draws + abcde_init!; {cores} threads; sweeps alone: {nsims / max(sweeps_s, 1e-9):.3e} evals/s)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(args.steps, 1), "higher_is_better": True,
            "scaling": "strong" if args.particles_total else "weak", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": config_dict(cfg, args, total // world, total, world, "strong" if args.particles_total else "weak"),
            "sample_iters": iters, "sweeps_only_value": nsims / max(sweeps_s, 1e-9),
            "note": "ms_per_step of this arm is the time of the bounded sample, not a time-to-target; compare the evals/s rates",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def sharded_selfcheck(A, ctx, rank, world):
    """One small sharded run against the oracle's island variant (rank-local DE partners, everything else global):
    identical iteration / simulation counts and eps history, this rank's rows within 1e-9.  Raises on mismatch."""
    O, _ = cpu_threads()
    n = 3000 * world
    spec, data = [("normal", 0.0, math.sqrt(10.0))], [3.0, 1.0]
    want = O.smc_run(spec, "gauss1d", data, 0.3, nparticles=n, seed=97, islands=world)
    got = A.abcdesmc(A.host.Normal(0.0, math.sqrt(10.0)), A.Model("gauss1d", data), 0.3, None, nparticles=n, rng=97, verbose=False, ctx=ctx)
    lo, hi = A.shard_range(n, rank, world)
    ok = (got.iters, got.nsims) == (want.iters, want.nsims) and np.array_equal(got.eps_hist, want.hist["eps"]) and \
        np.allclose(got.P, want.P[lo:hi, 0], rtol=1e-9, atol=1e-12) and np.array_equal(got.Wns > 0, want.Wns[lo:hi] > 0) and \
        abs(got.logZ - want.logZ) <= 1e-9 * abs(want.logZ)
    if not ok:
        raise SystemExit(f"bench.py: sharded self-check FAILED on rank {rank}/{world}: iters {got.iters} vs {want.iters}, nsims {got.nsims} vs "
                         f"{want.nsims}, logZ {got.logZ} vs {want.logZ}")
    return dict(particles=n, iters=int(got.iters), nsims=int(got.nsims), equals="oracle.smc_run(islands=%d)" % world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 10 for config 2, 3 otherwise)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the configuration's)")
    ap.add_argument("--particles-total", type=int, default=0, help="strong scaling: total population, split over the GPUs")
    ap.add_argument("--max-iters", type=int, default=-1, help="bound each step to this many SMC iterations (0: run to eps_target; default: the configuration's)")
    ap.add_argument("--model", default="", help="override the configuration's model (e.g. gk_f32, lotka_volterra_lin)")
    ap.add_argument("--mc", action="store_true", help="time abcdemc! (--generations) instead of abcdesmc!")
    ap.add_argument("--generations", type=int, default=50)
    ap.add_argument("--ref-particles", type=int, default=0, help="particles of the cpu_baseline sample (0: the workload's own)")
    ap.add_argument("--ref-iters", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--segments", action="store_true", help="relaxed-parity mode: warp-coherent partner segments (NOT the headline)")
    ap.add_argument("--systematic", action="store_true", help="relaxed-parity mode: systematic resampling (NOT the headline)")
    ap.add_argument("--fp32-state", dest="fp32_state", action="store_true", help="relaxed-parity mode: FP32 particle state (NOT the headline)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.model:
        cfg["model"] = args.model
        if args.model == "gk_f32":
            cfg.update(bound="fp32", flops_per_eval=30.0 * 10000, definition_flops_per_eval=95.0 * 10000, dtype="f32 draws (relaxed-precision mode), f64 state")
    if args.particles <= 0:
        args.particles = cfg["particles"]
    if args.max_iters < 0:
        args.max_iters = cfg["max_iters"]
    if args.steps <= 0:
        args.steps = 10 if args.config == 2 else 3
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import abcdez_b200 as A
    import ctypes as C
    stream = torch.cuda.current_stream().cuda_stream
    ctx = A.Context(local_rank, stream=stream)
    selfcheck = None
    if world > 1:
        A.dist.init_sharded(ctx)               # collective: NCCL bootstrap + IPC-mapped mailboxes
        selfcheck = sharded_selfcheck(A, ctx, rank, world)
    cls = {"normal": A.host.Normal, "uniform": A.host.Uniform}
    prior = A.Factored(*[cls[p[0]](*p[1:]) for p in cfg["prior"]])
    model = A.Model(cfg["model"], cfg["data"])
    D = len(cfg["prior"])
    scaling = "strong" if args.particles_total else "weak"
    Nglobal = args.particles_total if args.particles_total else args.particles * world
    lo, hi = A.shard_range(Nglobal, rank, world)
    Nloc = hi - lo
    L = A.lib()
    key = PHILOX_KEY0 + args.config - 1
    eps_target = cfg["eps_target"]

    def run(seed, host_out=None, profile=False):
        if args.mc:
            o = A.host._McOpts(Nglobal, args.generations, seed)
            r = A.host._McResult()
            if host_out is not None:
                r.P, r.C = host_out[0], host_out[2]
            rc = L.abcdez_mc_run(ctx._h, prior.handle(ctx), model.handle(ctx), C.c_double(eps_target), C.byref(o), C.byref(r))
            if rc:
                raise RuntimeError(L.abcdez_last_error().decode())
            r.iters = args.generations; r.n_sweeps = args.generations; r.logZ = float("nan")
            r.head_ms = r.resample_ms = r.init_ms = 0.0
            return r
        o = A.host._SmcOpts(); L.abcdez_smc_opts_default(C.byref(o))
        o.nparticles = Nglobal; o.nsims_max = 10**15; o.seed = seed; o.verboseout = 0; o.profile = int(profile); o.sync_every = 4
        o.max_iters = args.max_iters
        o.partner_segments = int(args.segments); o.systematic_resampling = int(args.systematic); o.fp32_state = int(args.fp32_state)
        r = A.host._SmcResult()
        if host_out is not None:
            r.P, r.Wns, r.C = host_out[:3]
            if len(host_out) > 3:
                r.blobs = host_out[3]
        rc = L.abcdez_smc_run(ctx._h, prior.handle(ctx), model.handle(ctx), C.c_double(eps_target), C.byref(o), C.byref(r))
        if rc:
            raise RuntimeError(L.abcdez_last_error().decode())
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    base = key                                 # one population: the same seed on every rank
    for s in range(args.warmup):
        run(base + 100000 + s)
    # ---- timed region: K complete runs, device-resident results, no per-iteration profiling events -------
    sampler = ClockSampler(local_rank); sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nsims = iters = launches = sweeps = 0; logZ = []
    for s in range(args.steps):
        r = run(base + s)
        nsims += r.nsims; iters += r.iters; launches += r.n_launches; sweeps += r.n_sweeps
        logZ.append(r.logZ)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # ---- per-kernel times: the same runs again with the library's profile events (outside the headline) ----
    kp = dict(nsims=0, sweeps=0, iters=0, resamples=0, sweep_ms=0.0, head_ms=0.0, resample_ms=0.0, init_ms=0.0, total_ms=0.0)
    for s in range(min(args.steps, 3)):
        r = run(base + s, profile=True)
        kp["nsims"] += r.nsims; kp["sweeps"] += r.n_sweeps; kp["iters"] += r.iters; kp["sweep_ms"] += r.sweep_ms
        kp["head_ms"] += r.head_ms; kp["resample_ms"] += r.resample_ms; kp["init_ms"] += r.init_ms
        if not args.mc:
            kp["resamples"] += r.n_resamples; kp["total_ms"] += r.total_ms + r.init_ms
        else:
            kp["total_ms"] += r.total_ms
    kp["runs"] = min(args.steps, 3)
    # ---- e2e: the reference-facing call with HOST result buffers ------------------------------
    Ph = torch.empty((Nloc, D), dtype=torch.float64).pin_memory(); Wh = torch.empty(Nloc, dtype=torch.float64).pin_memory()
    Ch = torch.empty(Nloc, dtype=torch.float64).pin_memory()
    host_out = (Ph.data_ptr(), Wh.data_ptr(), Ch.data_ptr())
    Bh = None
    if model.blob_bytes:
        Bh = torch.empty((Nloc, model.blob_bytes // 8), dtype=torch.float64).pin_memory()
        host_out = host_out + (Bh.data_ptr(),)
    run(base + 200000, host_out=host_out)
    barrier()
    t0 = time.perf_counter(); e2e_nsims = 0
    ke = max(1, min(args.steps, 5))
    trace = []                                           # (ABCDEZ_TRACE: where the host time of an e2e step goes)
    for s in range(ke):
        ta = time.perf_counter()
        r = run(base + 300000 + s, host_out=host_out)
        tb = time.perf_counter()
        e2e_nsims += r.nsims
        # the host reads the result: a strided sample of the distances and the evidence.  (Summing all of Ch with torch took
        # 15-30 ms per step under torchrun at 4 ranks -- one intra-op thread competing with the other ranks' polling threads --
        # and the uneven finish times then stalled the next run's first collective: profiles/README.md, "e2e of sharded runs".)
        assert math.isfinite(float(Ch[::4099].sum())) and (args.mc or math.isfinite(r.logZ))
        trace.append((1e3 * (tb - ta), 1e3 * (time.perf_counter() - tb)))
    tc = time.perf_counter()
    barrier()
    e2e_s = time.perf_counter() - t0
    if os.environ.get("ABCDEZ_TRACE"):
        print(f"[bench] rank {rank}: e2e steps (library call ms, host read ms) {[(round(a, 2), round(b, 2)) for a, b in trace]}, "
              f"final barrier {1e3 * (time.perf_counter() - tc):.2f} ms, total {1e3 * e2e_s:.2f} ms", file=sys.stderr, flush=True)
    d2h = Nglobal * ((D + 2) * 8 + model.blob_bytes)    # all ranks together
    mean_events = float(Bh[:, 1].mean()) if (Bh is not None and cfg["model"] == "birth_death") else None
    h2d = (len(cfg["data"]) + 4 * D) * 8 + 4 * D        # bound data + prior parameters; the state is born on the device

    # ---- max over ranks / sums ---------------------------------------------------------------
    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    # nsims / iters / sweeps are properties of the one sharded population (identical on every rank); launches add up
    c = torch.tensor([launches], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(c, op=dist.ReduceOp.SUM)
    ms_all, e2e_ms_all = t.tolist(); launches_all = c.tolist()[0]
    if rank == 0:
        peak, peak_src = peaks()
        value = nsims / (ms_all * 1e-3)
        line = {"metric": METRIC if not args.mc else METRIC.replace("abcdesmc!, time-to-target-eps", f"abcdemc!, {args.generations} generations"),
                "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": cfg["dtype"], "data": "synthetic", "config": config_dict(cfg, args, Nglobal // world, Nglobal, world, scaling),
                "time_to_target_eps_ms": (ms_all / args.steps) if not args.max_iters and not args.mc else None,
                "iters_per_run": iters / args.steps, "sweeps_per_run": sweeps / args.steps,
                "logZ_mean": float(np.mean(logZ)), "logZ_sd": float(np.std(logZ)),
                "gpu_launches": int(launches_all),
                "clocks": clocks,
                "e2e": {"value": e2e_nsims / (e2e_ms_all * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": d2h, "steps": ke, "ms_per_step": e2e_ms_all / ke}}
        if selfcheck:
            line["sharded_selfcheck"] = selfcheck
        # ---- roofline of the dominant kernel (the sweep), from the profiled runs: rank 0 simulates 1/world of the evaluations
        sw_ms, sw_n = kp["sweep_ms"], max(kp["sweeps"], 1)
        ev_rank = kp["nsims"] / world
        if args.mc:
            sw_ms = kp["total_ms"]
        if cfg["bound"] == "hbm":
            ach = cfg["bytes_per_eval"] * ev_rank / (sw_ms * 1e-3) / 1e9 if sw_ms > 0 else None
            line["roofline"] = {"bound": "hbm", "kernel": f"smc_sweep_kernel<{cfg['model']}>", "achieved": ach, "peak": peak,
                                "unit": "GB/s", "frac": (ach / peak) if ach else None, "traffic": sweep_traffic() if args.config == 2 and args.particles == 1_000_000 else None,
                                "peak_source": peak_src, "algorithmic_bytes_per_eval": cfg["bytes_per_eval"],
                                "algorithmic_bytes_per_launch": cfg["bytes_per_eval"] * ev_rank / sw_n,
                                "avg_launch_ms": sw_ms / sw_n, "sweep_share_of_step": sw_ms / kp["total_ms"] if kp["total_ms"] > 0 else None}
        else:
            fl = cfg.get("flops_per_eval")
            note = "FMA = 2 flops, counted from the model definition (DESIGN.md section 5)"
            if fl is None:                      # birth-death: flops follow the event count, estimated from the returned blobs
                fl = cfg["flops_per_event"] * float(mean_events or 60.0)
                note += f"; {mean_events:.1f} events per evaluation (mean over the returned particles' blobs)"
            if cfg.get("definition_flops_per_eval"):
                note = ("FMA = 2 flops; counted for what the evaluation needs: the n normal draws (Box-Muller in portable math) -- the octiles are "
                        "selected in z space and only ~6 candidates per octile go through the quantile function (DESIGN.md section 5); pushing all "
                        f"n draws through it, as the definition reads and the CPU port does, is {cfg['definition_flops_per_eval']:.0f} flops")
            pk = FP32_PEAK_TFLOPS if cfg["bound"] == "fp32" else FP64_PEAK_TFLOPS
            ach = fl * ev_rank / (sw_ms * 1e-3) / 1e12 if sw_ms > 0 else None
            line["roofline"] = {"bound": cfg["bound"], "kernel": f"sweep kernel of {cfg['model']}", "achieved": ach, "peak": pk,
                                "unit": "TFLOP/s", "frac": (ach / pk) if ach else None, "traffic": None,
                                "peak_source": "nominal vector-pipe peak (148 SMs x 64 FP64 / 128 FP32 FMA lanes x 2 x 1.965 GHz); no measured figure on record; tensor cores are not used (no dense contraction on this path)",
                                "algorithmic_flops_per_eval": fl, "definition_flops_per_eval": cfg.get("definition_flops_per_eval", fl),
                                "note": note, "avg_launch_ms": sw_ms / sw_n,
                                "sweep_share_of_step": sw_ms / kp["total_ms"] if kp["total_ms"] > 0 else None}
        if not args.mc:
            # every other kernel of the step with its own roofline fraction (SURVEY.md 8(d) algorithmic bytes per particle)
            Nl = Nglobal / world; Bb = 16 if cfg["model"] == "birth_death" else 0
            kern = {}
            if kp["iters"]:
                hb = 40.0 * Nl; ht = kp["head_ms"] / kp["iters"]
                kern["head_kernel"] = {"bound": "hbm (L2-resident below ~5e6 particles; latency-bound: 3 grid barriers)", "avg_launch_ms": ht,
                                       "algorithmic_bytes_per_launch": hb, "achieved_gbs": hb / (ht * 1e-3) / 1e9 if ht > 0 else None,
                                       "frac": hb / (ht * 1e-3) / 1e9 / peak if ht > 0 else None, "share_of_step": kp["head_ms"] / kp["total_ms"]}
            if kp["resamples"]:
                rb = (16.0 * D + 57 + 2 * Bb) * Nl; rt = kp["resample_ms"] / kp["resamples"]
                kern["resample_uniform_kernel"] = {"bound": "hbm (random gather)", "avg_launch_ms": rt, "algorithmic_bytes_per_launch": rb,
                                                   "achieved_gbs": rb / (rt * 1e-3) / 1e9 if rt > 0 else None,
                                                   "frac": rb / (rt * 1e-3) / 1e9 / peak if rt > 0 else None,
                                                   "share_of_step": kp["resample_ms"] / kp["total_ms"], "launches_that_resampled": kp["resamples"] / kp["runs"]}
            ib = (8.0 * D + 16 + Bb) * Nl; it_ = kp["init_ms"] / kp["runs"]
            kern["init_kernel"] = {"bound": "simulator (prior draws + one simulation per particle); bytes written vs HBM for reference",
                                   "avg_launch_ms": it_, "algorithmic_bytes_per_launch": ib, "achieved_gbs": ib / (it_ * 1e-3) / 1e9 if it_ > 0 else None,
                                   "frac": ib / (it_ * 1e-3) / 1e9 / peak if it_ > 0 else None, "share_of_step": kp["init_ms"] / kp["total_ms"]}
            line["kernels"] = kern
        if world == 1 and not args.no_cpu_baseline and not args.mc:
            rp = args.ref_particles if args.ref_particles > 0 else args.particles
            ri = args.ref_iters if args.config in (2, 4) else 2
            if args.config == 3:
                rp = min(rp, 20000)             # 10^4 draws + a sort per evaluation: ~0.15 ms per evaluation and core
            if args.config == 5:
                rp = min(rp, 2_000_000)
            r, dt, cores = cpu_sample(cfg, rp, ri, key)
            line["cpu_baseline"] = {"value": r.nsims / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sweeps_only_value": r.nsims / max(r.sweep_seconds, 1e-9),
                                    "sample": f"first {ri} SMC iterations of abcdesmc! at {rp} particles, O(1)-partner C/OpenMP port of the "
                                              "reference (oracle/abcdez_oracle.c), incl. prior draws + abcde_init!"}
            if args.config == 2:
                rf, dtf, _ = cpu_sample(cfg, 20000, 2, key, faithful=True)
                line["cpu_baseline_faithful"] = {"value": rf.nsims / dtf, "unit": UNIT, "cores": cores, "kind": "port",
                                                 "sample": "first 2 SMC iterations at 20000 particles with the reference's O(N) "
                                                           "StatsBase.wsample partner scans (O(N^2) per sweep)"}
                # time-to-target on a size the CPU port can finish: both arms run the SAME complete abcdesmc! run
                nt = 200_000
                rt, dtt, _ = cpu_sample(cfg, nt, 0, key)
                o = A.host._SmcOpts(); L.abcdez_smc_opts_default(C.byref(o))
                o.nparticles = nt; o.nsims_max = 10**15; o.seed = key; o.verboseout = 0; o.sync_every = 4
                rr = A.host._SmcResult()
                for _ in range(2):
                    tg0 = time.perf_counter()
                    L.abcdez_smc_run(ctx._h, prior.handle(ctx), model.handle(ctx), C.c_double(eps_target), C.byref(o), C.byref(rr))
                    tg = time.perf_counter() - tg0
                line["time_to_target_same_run"] = {"particles": nt, "eps_target": eps_target, "cpu_port_s": dtt, "cpu_cores": cores, "gpu_s": tg,
                                                   "iters": (int(rt.iters), int(rr.iters)), "nsims": (int(rt.nsims), int(rr.nsims)),
                                                   "note": "same seed: the two runs are the same run decision by decision"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
