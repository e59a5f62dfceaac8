"""One complete abcdesmc! run of a BASELINE.json configuration at its named size, on 1..8 GPUs (torchrun: one process per
GPU, one sharded population), with what the north star asks to see: iterations, simulations, time-to-target-eps, logZ and
the posterior against the generating parameters.  Not a bench.py value; the output goes under profiles/.

  python -m torch.distributed.run --nproc-per-node 8 ... scripts/run_full.py --config 3 --particles-total 10000000 --eps 0.25
"""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, required=True)
ap.add_argument("--particles-total", type=int, required=True)
ap.add_argument("--eps", type=float, default=-1.0)
ap.add_argument("--seed", type=int, default=12345)
ap.add_argument("--model", default="")
ap.add_argument("--fp32-state", dest="fp32_state", action="store_true", help="relaxed-parity mode: FP32 particle state")
ap.add_argument("--max-iters", type=int, default=0, help="safety bound on the SMC iterations (0: none); the reached eps is reported")
args = ap.parse_args()

import torch
import torch.distributed as dist
import bench
import abcdez_b200 as A

cfg = dict(bench.CONFIGS[args.config])
if args.model:
    cfg["model"] = args.model
eps_target = args.eps if args.eps >= 0 else cfg["eps_target"]
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = A.Context(local, stream=torch.cuda.current_stream().cuda_stream)
if world > 1:
    A.dist.init_sharded(ctx)
cls = {"normal": A.host.Normal, "uniform": A.host.Uniform}
prior = A.Factored(*[cls[p[0]](*p[1:]) for p in cfg["prior"]])
model = A.Model(cfg["model"], cfg["data"])
N = args.particles_total
A.abcdesmc(prior, model, eps_target, None, nparticles=min(N, 20000 * world), rng=1, verbose=False, ctx=ctx, nsims_max=10**15, max_iters=3)   # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
r = A.abcdesmc(prior, model, eps_target, None, nparticles=N, rng=args.seed, verbose=False, ctx=ctx, nsims_max=10**15, sync_every=4, profile=True, max_iters=args.max_iters, fp32_state=args.fp32_state)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
# posterior moments over the whole population: weighted sums of this rank's block, reduced over the ranks
w = r.Wns; P = r.P.reshape(len(w), -1)
acc = np.concatenate([[w.sum()], (P * w[:, None]).sum(0), (P * P * w[:, None]).sum(0), [float((w > 0).sum())]])
t = torch.tensor(acc, dtype=torch.float64, device="cuda")
tm = torch.tensor([wall], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
acc = t.cpu().numpy(); d = P.shape[1]
mean = acc[1:1 + d] / acc[0]; sd = np.sqrt(np.maximum(acc[1 + d:1 + 2 * d] / acc[0] - mean ** 2, 0.0))
if rank == 0:
    truth = {3: [3.0, 1.0, 2.0, 0.5], 4: [1.2, 0.9, 0.7, 0.6]}.get(args.config)
    out = {"config": args.config, "workload": cfg["name"], "model": cfg["model"], "n_gpus": world, "particles": N, "eps_target": eps_target,
           "eps_reached": r.eps, "iters": int(r.iters), "nsims": int(r.nsims), "logZ": r.logZ, "time_to_target_eps_s": float(tm.item()),
           "device_loop_ms": r.stats["total_ms"], "sweep_ms": r.stats["sweep_ms"], "head_ms": r.stats["head_ms"], "resample_ms": r.stats["resample_ms"],
           "evals_per_s": r.nsims / float(tm.item()), "n_resamples": int(r.stats["n_resamples"]), "alive_at_end": int(acc[-1]),
           "posterior_mean": [float(x) for x in mean], "posterior_sd": [float(x) for x in sd], "generating_parameters": truth}
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
