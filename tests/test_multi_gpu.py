"""The N > 1 path: host-side sharding logic on CPU (gloo, world 2 and 3) and the sharded runs
themselves on a box with >= 2 GPUs (torchrun + NCCL bootstrap, in-kernel NVLink exchanges)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "multi_gpu_worker.py")


def _torchrun(nproc, extra, timeout, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), WORKER, *extra]
    env = dict(os.environ)
    env.setdefault("OMP_NUM_THREADS", "4")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


@pytest.mark.parametrize("world", [2, 3])
def test_sharding_host_logic_gloo(world):
    r = _torchrun(world, ["--cpu"], 300, 29511 + world)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("cpu sharding logic ok") == world


def test_shard_range_is_a_partition(A):
    for N in (0, 1, 7, 1000, 2**31 - 2):
        for world in (1, 2, 3, 8):
            b = [A.shard_range(N, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == N
            assert all(b[i][1] == b[i + 1][0] and b[i][0] <= b[i][1] for i in range(world - 1))


def test_oracle_islands_change_partners_only(oracle):
    """islands = R keeps the init / quantile / reweight / resampling global: the first iteration's eps
    (before any DE move) is identical, the runs differ afterwards only through the partner pools."""
    import math
    spec, data = [("normal", 0.0, math.sqrt(10.0))], [3.0, 1.0]
    a = oracle.smc_run(spec, "gauss1d", data, 0.3, nparticles=2000, seed=9)
    b = oracle.smc_run(spec, "gauss1d", data, 0.3, nparticles=2000, seed=9, islands=4)
    assert a.hist["eps"][1] == b.hist["eps"][1]
    assert abs(a.logZ - b.logZ) < 0.5 and (a.nsims != b.nsims or a.iters != b.iters or a.logZ != b.logZ)


@pytest.mark.gpu
def test_sharded_runs_match_island_oracle():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    r = _torchrun(world, [], 1500, 29533)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert r.stdout.count("multi-gpu parity ok") == world


@pytest.mark.gpu
def test_single_process_multi_gpu_context_matches_island_oracle(A, oracle):
    """abcdez_init_multi (SURVEY.md 8b "Threading"): ONE host call drives every GPU -- the same sharded kernels and
    in-kernel exchanges as the one-process-per-GPU runs, peers mapped with cudaDeviceEnablePeerAccess.  The whole
    population comes back, identical to oracle(islands = n_gpus); the Python `parallel=True` maps to it."""
    import math
    import numpy as np
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    R = 2 if n < 4 else 4
    ctx = A.Context.multi(R)
    assert ctx.n_gpus == R
    cases = [([("normal", 0.0, math.sqrt(10.0))], "gauss1d", [3.0, 1.0], 0.3, 20011, {}),
             ([("normal", 0.0, 2.0)] * 10, "gauss_corr10", list(np.linspace(-1, 1, 10)) + [0.5], 2.5, 40000, {}),
             ([("uniform", 0.0, 2.0)] * 2, "birth_death", [20.0, 8, 0.5, 5000.0, 22, 25, 24, 30, 33, 31, 36, 40], 3.0, 8000, {}),
             ([("normal", 0.0, math.sqrt(10.0))], "gauss1d", [3.0, 1.0], 0.3, 3001, dict(kind="epa"))]
    cls = {"normal": A.host.Normal, "uniform": A.host.Uniform}
    for spec, name, data, eps, N, kw in cases:
        prior = A.Factored(*[cls[s_[0]](*s_[1:]) for s_ in spec])
        kind = kw.get("kind", "indicator_strict")
        want = oracle.smc_run(spec, name, data, eps, nparticles=N, seed=31, islands=R, nsims_max=10**9, kind=kind)
        got = A.abcdesmc(prior, A.Model(name, data), eps, None, nparticles=N, rng=31, nsims_max=10**9, verbose=False, ctx=ctx,
                         ABCk=kind, exact_scan=True)
        assert (got.iters, got.nsims) == (want.iters, want.nsims), name
        assert np.array_equal(got.eps_hist, want.hist["eps"])
        assert got.P.shape[0] == N and np.array_equal(got.Wns > 0, want.Wns > 0)
        np.testing.assert_allclose(got.P.reshape(N, -1), want.P, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got.C, want.C, rtol=1e-9, atol=1e-10)
        assert abs(got.logZ - want.logZ) <= 1e-9 * abs(want.logZ)
    wm = oracle.mc_run(cases[0][0], "gauss1d", [3.0, 1.0], 0.3, nparticles=4003, generations=30, seed=32, islands=R)
    gm = A.abcdemc(A.host.Normal(0.0, math.sqrt(10.0)), A.Model("gauss1d", [3.0, 1.0]), 0.3, None, nparticles=4003, generations=30, rng=32,
                   verbose=False, ctx=ctx)
    assert gm.nsims == wm.nsims and gm.P.shape[0] == 4003
    np.testing.assert_allclose(gm.C, wm.C, rtol=1e-9, atol=1e-10)
    # stage-level calls on the multi context run on its first GPU
    assert np.isfinite(A.Factored(A.host.Normal(0.0, 1.0)).logpdf([0.3], ctx=ctx))
    ctx.close()
