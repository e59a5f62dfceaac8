"""Sharded runs from Python: one process per GPU (torchrun), one population across all of them.

torch.distributed is plumbing here: it carries the 128-byte NCCL id from rank 0 to the others (any
backend: nccl on the GPU box, gloo in the CPU tests) and gathers the per-rank result blocks on request.
The run itself is the library's: `Context.comm_init` maps the peers' mailboxes and population slabs, after
which `abcdesmc(..., ctx=ctx)` is collective -- eps-quantile histograms, weight sums, ESS and counters are
exchanged inside the kernels over NVLink and resampling gathers particles straight from peer HBM
(include/abcdez_cuda.h, "sharded runs").
"""
from __future__ import annotations

import dataclasses
import os
from typing import Callable, Optional

import numpy as np

from . import host


def _dist():
    import torch.distributed as dist
    if not dist.is_initialized():
        raise host.ABCdeZError(host.ERR_BAD_ARG, "torch.distributed is not initialised (launch with torchrun and call "
                                                 "init_process_group first)")
    return dist


def broadcast_bytes(payload: Optional[bytes], nbytes: int, src: int = 0) -> bytes:
    """Rank `src` passes `payload`; every rank returns it."""
    import torch
    dist = _dist()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def init_sharded(ctx: Optional[host.Context] = None, make_id: Callable[[], bytes] = host.Context.nccl_unique_id) -> host.Context:
    """Collective: attach this process' context to the sharded communicator of the torch.distributed world."""
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = ctx or host.default_context()
    uid = broadcast_bytes(make_id() if rank == 0 else None, 128) if world > 1 else None
    ctx.comm_init(rank, world, uid)
    return ctx


def shard_bounds(N: int, world: int):
    return [host.shard_range(N, r, world) for r in range(world)]


def gather_result(res, root: Optional[int] = None):
    """Assemble the whole population's (P, Wns, C, blobs) of an SMCResult / (P, C, blobs) of an MCResult from the
    per-rank blocks (global particle order).
    root=None: every rank gets it (all_gather); else only `root` (others return None)."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(), dist.get_world_size()
    if world == 1:
        return res
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    N = int(res.stats.get("nparticles", 0))
    bounds = shard_bounds(N, world)
    nmax = max(hi - lo for lo, hi in bounds)

    def gather(a: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a)
        flat = a.reshape(a.shape[0], -1)
        pad = np.zeros((nmax, flat.shape[1]), dtype=flat.dtype)
        pad[:flat.shape[0]] = flat
        t = torch.from_numpy(pad).to(dev)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        parts = [o.cpu().numpy()[:hi - lo] for o, (lo, hi) in zip(outs, bounds)]
        full = np.concatenate(parts, axis=0)
        return full.reshape((N,) + a.shape[1:])

    blobs = gather(res.blobs) if res.blobs.ndim == 2 and res.blobs.shape[1] > 0 else np.empty((N, 0), dtype=np.uint8)
    if isinstance(res, host.MCResult):
        full = dataclasses.replace(res, P=gather(res.P), C=gather(res.C), blobs=blobs)
    else:
        full = dataclasses.replace(res, P=gather(res.P), Wns=gather(res.Wns), C=gather(res.C), blobs=blobs)
    if root is not None and rank != root:
        return None
    return full
