"""Worker of tests/test_multi_gpu.py: run under torchrun with one process per GPU (NCCL) -- or, with
--cpu, as a gloo world that exercises only the host-side sharding logic (no CUDA calls).

GPU mode checks, on every rank:
  1. the in-kernel exchange (abcdez_comm_selftest) against its closed-form checksum;
  2. sharded abcdesmc! runs == the CPU oracle with `islands = world` on the same Philox seed: identical
     (iters, nsims), eps history, logZ <= 1e-9 relative, and the gathered (P, C, Wns) row by row.
"""
import argparse
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def selftest_checksum(world: int, rounds: int, bins: int = 2048) -> int:
    """What xchg_selftest_kernel (comm.cu) accumulates, for either ring."""
    tot = 0
    for it in range(rounds):
        for r in range(world):
            tot += (r + 1) * 1000 + it + (r + 1)        # header word 0 + the high half of header word 1
            tot += sum(r * 7 + b + it for b in range(bins))
            tot += 3 * r + it
    return tot & 0xFFFFFFFFFFFFFFFF


def cpu_main():
    import torch.distributed as dist
    import abcdez_b200 as A
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # the NCCL id travels from rank 0 to everyone (here a fake one: no GPU, no communicator)
    fake = bytes((7 * i + 1) % 256 for i in range(128))
    got = A.dist.broadcast_bytes(fake if rank == 0 else None, 128)
    assert got == fake
    # shard ranges tile [0, N) in rank order
    for N in (8, 1000, 1001, 10**6 + 3):
        b = A.dist.shard_bounds(N, world)
        assert b[0][0] == 0 and b[-1][1] == N and all(b[i][1] == b[i + 1][0] for i in range(world - 1))
        assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1
    # gather_result reassembles the global particle order from ragged blocks
    N = 1001
    lo, hi = A.shard_range(N, rank, world)
    full_P = np.arange(N * 3, dtype=np.float64).reshape(N, 3)
    res = A.host.SMCResult(P=full_P[lo:hi].copy(), Wns=np.full(hi - lo, 1.0 / N), C=np.arange(lo, hi, dtype=np.float64),
                           eps=0.5, logZ=-1.0, blobs=np.empty((hi - lo, 0), dtype=np.uint8),
                           stats={"nparticles": N})
    g = A.dist.gather_result(res)
    assert np.array_equal(g.P, full_P) and np.array_equal(g.C, np.arange(N, dtype=np.float64)) and g.Wns.shape == (N,)
    g0 = A.dist.gather_result(res, root=0)
    assert (g0 is not None) == (rank == 0)
    # scalar-prior results (1-d P) and blobs
    res1 = A.host.SMCResult(P=full_P[lo:hi, 0].copy(), Wns=res.Wns, C=res.C, eps=0.5, logZ=-1.0,
                            blobs=np.full((hi - lo, 8), rank, dtype=np.uint8), stats={"nparticles": N})
    g1 = A.dist.gather_result(res1)
    assert g1.P.shape == (N,) and np.array_equal(g1.P, full_P[:, 0]) and g1.blobs.shape == (N, 8)
    assert g1.blobs[0, 0] == 0 and g1.blobs[-1, 0] == world - 1
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: cpu sharding logic ok")


CASES = [
    # name, prior spec, data, eps_target, N, seed, kwargs
    ("gauss1d", [("normal", 0.0, math.sqrt(10.0))], [3.0, 1.0], 0.3, 1000, 2024, {}),
    ("gauss1d", [("normal", 0.0, math.sqrt(10.0))], [3.0, 1.0], 0.05, 20011, 5, {}),
    ("gauss_corr10", [("normal", 0.0, 2.0)] * 10, list(np.linspace(-1, 1, 10)) + [0.5], 3.0, 40000, 7, {}),
    ("twod", [("normal", 0.0, 5.0)] * 2, [], 0.05, 30001, 11, {"kind": "indicator"}),
    ("birth_death", [("uniform", 0.0, 2.0)] * 2, [20.0, 4.0, 0.5, 400.0, 24.0, 30.0, 33.0, 41.0], 6.0, 8000, 3, {}),
]


def gpu_main():
    import torch
    import torch.distributed as dist
    import abcdez_b200 as A
    from oracle import oracle as O
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = A.Context(local)
    A.dist.init_sharded(ctx)
    assert (ctx.rank, ctx.world) == (rank, world)
    for mode, label in ((0, "fenced ring"), (1, "low-latency ring")):
        chk, us = ctx.comm_selftest(64, mode)
        assert chk == selftest_checksum(world, 64), (mode, chk, selftest_checksum(world, 64))
        if rank == 0:
            print(f"in-kernel exchange ok, {label}: {us:.2f} us per round (2048-bin histogram + scalar record), world {world}")
    O.build()
    fams = {"normal": A.host.Normal, "uniform": A.host.Uniform}
    for name, spec, data, eps_t, N, seed, kw in CASES:
        prior = A.Factored(*[fams[s[0]](*s[1:]) for s in spec])
        kind = kw.get("kind", "indicator_strict")
        for sync_every in (1, 4):
            got = A.abcdesmc(prior, A.Model(name, data), eps_t, None, nparticles=N, rng=seed, verbose=False, ctx=ctx,
                             ABCk=kind, nsims_max=10**9, sync_every=sync_every)
            full = A.dist.gather_result(got)
            want = O.smc_run(spec, name, data, eps_t, nparticles=N, seed=seed, kind=kind, nsims_max=10**9, islands=world)
            assert (got.iters, got.nsims) == (want.iters, want.nsims), (name, got.iters, got.nsims, want.iters, want.nsims)
            assert got.eps == want.eps
            assert abs(got.logZ - want.logZ) <= 1e-9 * max(1.0, abs(want.logZ)), (got.logZ, want.logZ)
            n = len(got.eps_hist)
            assert np.array_equal(got.eps_hist, want.hist["eps"][:n])
            np.testing.assert_allclose(got.esss, want.hist["ess"][:n], rtol=1e-9)
            np.testing.assert_array_equal(got.ranges_eps[:, 0], want.hist["dmin"][:n])
            np.testing.assert_array_equal(got.ranges_eps[:, 1], want.hist["dmax"][:n])
            assert np.array_equal(full.Wns > 0, want.Wns > 0)
            np.testing.assert_allclose(full.P.reshape(N, -1), want.P, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(full.C, want.C, rtol=1e-9, atol=1e-12)
            np.testing.assert_allclose(full.Wns, want.Wns, rtol=1e-9)
        if rank == 0:
            print(f"sharded {name} N={N} world={world}: iters {got.iters} nsims {got.nsims} logZ {got.logZ:.6f} == oracle(islands={world})")
    # sharded abcdemc!: global extrema / counts, rank-local base particle and partners
    for name, spec, data, eps_t, N, seed, gens in [("gauss1d", [("normal", 0.0, math.sqrt(10.0))], [3.0, 1.0], 0.5, 4003, 13, 12),
                                                   ("twod", [("normal", 0.0, 5.0)] * 2, [], 1.0, 3000, 17, 8),
                                                   # a stepped simulator: init and sweeps through the queue on every rank
                                                   ("birth_death", [("uniform", 0.0, 2.0)] * 2, [20.0, 4.0, 0.5, 400.0, 24.0, 30.0, 33.0, 41.0], 6.0, 2500, 19, 6)]:
        prior = A.Factored(*[fams[s_[0]](*s_[1:]) for s_ in spec])
        got = A.abcdemc(prior, A.Model(name, data), eps_t, None, nparticles=N, generations=gens, rng=seed, verbose=False, ctx=ctx)
        full = A.dist.gather_result(got)
        want = O.mc_run(spec, name, data, eps_t, nparticles=N, generations=gens, seed=seed, islands=world)
        assert got.nsims == want.nsims and got.reached_eps == want.reached_eps, (name, got.nsims, want.nsims)
        np.testing.assert_allclose(full.P.reshape(N, -1), want.P, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(full.C, want.C, rtol=1e-9, atol=1e-12)
        if rank == 0:
            print(f"sharded abcdemc! {name} N={N} world={world}: nsims {got.nsims} reached {got.reached_eps} == oracle(islands={world})")
    # Epanechnikov kernels (general weights): the cumulative weights continue from rank to rank.  exact_scan: the
    # reference's sequential sum as a chain over the ranks == oracle(islands); default: parallel scans + one exchange
    for name, spec, data, eps_t, N, seed, kind in [("gauss1d", [("normal", 0.0, math.sqrt(10.0))], [3.0, 1.0], 0.3, 3001, 21, "epa"),
                                                   ("twod", [("normal", 0.0, 5.0)] * 2, [], 0.1, 6000, 23, "epa_strict")]:
        prior = A.Factored(*[fams[s_[0]](*s_[1:]) for s_ in spec])
        got = A.abcdesmc(prior, A.Model(name, data), eps_t, None, nparticles=N, rng=seed, verbose=False, ctx=ctx, ABCk=kind,
                         nsims_max=10**9, exact_scan=True)
        full = A.dist.gather_result(got)
        want = O.smc_run(spec, name, data, eps_t, nparticles=N, seed=seed, kind=kind, nsims_max=10**9, islands=world)
        assert (got.iters, got.nsims) == (want.iters, want.nsims), (name, kind, got.iters, got.nsims, want.iters, want.nsims)
        assert got.eps == want.eps and got.stats["n_resamples"] >= 1
        assert abs(got.logZ - want.logZ) <= 1e-9 * max(1.0, abs(want.logZ)), (got.logZ, want.logZ)
        np.testing.assert_allclose(full.P.reshape(N, -1), want.P, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(full.C, want.C, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(full.Wns, want.Wns, rtol=1e-9)
        fast = A.abcdesmc(prior, A.Model(name, data), eps_t, None, nparticles=N, rng=seed, verbose=False, ctx=ctx, ABCk=kind,
                          nsims_max=10**9)
        assert fast.eps == want.eps and abs(fast.logZ - want.logZ) < 0.5 and fast.stats["n_resamples"] >= 1
        if rank == 0:
            print(f"sharded {name} {kind} N={N} world={world}: iters {got.iters} nsims {got.nsims} logZ {got.logZ:.6f} == oracle(islands={world}); "
                  f"parallel scan: iters {fast.iters} logZ {fast.logZ:.6f}")
    # run-state snapshots of a sharded run: every rank keeps its own block; the resumed run is the uninterrupted one
    name, spec, data, eps_t, N, seed, kw = CASES[2]
    prior = A.Factored(*[fams[s[0]](*s[1:]) for s in spec])
    kws = dict(nparticles=N, rng=seed, verbose=False, ctx=ctx, nsims_max=10**9)
    whole = A.abcdesmc(prior, A.Model(name, data), eps_t, None, **kws)
    part = A.abcdesmc(prior, A.Model(name, data), eps_t, None, max_iters=11, return_state=True, **kws)
    assert part.iters == 11 and part.state is not None
    rest = A.abcdesmc(prior, A.Model(name, data), eps_t, None, state=part.state, **kws)
    assert (rest.iters, rest.nsims, rest.logZ, rest.eps) == (whole.iters, whole.nsims, whole.logZ, whole.eps)
    assert np.array_equal(rest.P, whole.P) and np.array_equal(rest.Wns, whole.Wns) and np.array_equal(rest.eps_hist, whole.eps_hist)
    if rank == 0:
        print(f"sharded run-state snapshot: cut at 11 of {whole.iters} iterations, resumed run == uninterrupted run, world {world}")
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    print(f"rank {rank}: multi-gpu parity ok")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    a = ap.parse_args()
    cpu_main() if a.cpu else gpu_main()
