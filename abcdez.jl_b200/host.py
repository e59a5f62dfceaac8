"""Host-side mirror of the ABCdeZ.jl user interface on top of libabcdez_cuda.so.

The reference's entry points (`abcdesmc!`, `abcdemc!`, `Factored`, the four ABC kernels,
`src/ABCdeZ.jl:1-20`) keep their names, positional arguments, keyword names (ASCII and the
reference's Greek spellings are both accepted), defaults and error messages; `dist!` is a
:class:`Model` -- a registered CUDA device functor bound to the observed data.  Everything
below is a thin ctypes binding: the particle work happens in the sm_100a kernels of the
library.  There is NO CPU fallback: if the library or a GPU is missing the calls raise.

The same C symbols are bound by the Julia shim in ``julia/ABCdeZCUDA.jl``.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABCDEZ_LIB", os.path.join(_HERE, "libabcdez_cuda.so"))

# status codes, include/abcdez_cuda.h
OK, ERR_BAD_ARG, ERR_CUDA, ERR_NAN_DISTANCE, ERR_NO_ALIVE, ERR_INIT_RETRY, ERR_PARTNER_RETRY, ERR_NCCL, \
    ERR_UNSUPPORTED = range(9)
FLAG_SIM, FLAG_ACC = 1, 2
TAG_MODEL, TAG_INIT_MODEL = 4, 7


class ABCdeZError(RuntimeError):
    """Raised for every non-zero status; mirrors Julia's ErrorException from `error(...)`."""

    def __init__(self, code: int, msg: str):
        super().__init__(msg)
        self.code = code


class _SmcOpts(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("alpha", C.c_double), ("delta_ess", C.c_double),
                ("nsims_max", C.c_int64), ("Kmcmc", C.c_int32), ("Kmcmc_min", C.c_double),
                ("kernel", C.c_int32), ("facc_stop", C.c_double), ("facc_min", C.c_double),
                ("facc_tune", C.c_double), ("seed", C.c_uint64), ("verboseout", C.c_int32),
                ("max_iters", C.c_int32), ("exact_scan", C.c_int32), ("profile", C.c_int32),
                ("sync_every", C.c_int32), ("fused_head", C.c_int32), ("systematic_resampling", C.c_int32),
                ("partner_segments", C.c_int32), ("fp32_state", C.c_int32), ("reserved1", C.c_int32)]


class _SmcResult(C.Structure):
    _fields_ = [("P", C.c_void_p), ("Wns", C.c_void_p), ("C", C.c_void_p), ("blobs", C.c_void_p),
                ("hist_cap", C.c_int32),
                ("h_eps", C.c_void_p), ("h_dmin", C.c_void_p), ("h_dmax", C.c_void_p), ("h_logZ", C.c_void_p),
                ("h_ess", C.c_void_p), ("h_facc", C.c_void_p), ("h_gamma0", C.c_void_p), ("h_Kmcmc", C.c_void_p),
                ("eps", C.c_double), ("logZ", C.c_double), ("iters", C.c_int64), ("nsims", C.c_int64),
                ("hist_len", C.c_int32), ("status", C.c_int32),
                ("n_resamples", C.c_int64), ("n_sweeps", C.c_int64), ("n_launches", C.c_int64),
                ("sweep_ms", C.c_double), ("total_ms", C.c_double), ("init_ms", C.c_double),
                ("hist_dropped", C.c_int32), ("reserved0", C.c_int32), ("head_ms", C.c_double), ("resample_ms", C.c_double)]


class _McOpts(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("generations", C.c_int32), ("seed", C.c_uint64)]


class _McResult(C.Structure):
    _fields_ = [("P", C.c_void_p), ("C", C.c_void_p), ("blobs", C.c_void_p), ("reached_eps", C.c_int32),
                ("nsims", C.c_int64), ("dmin", C.c_double), ("dmax", C.c_double), ("sweep_ms", C.c_double),
                ("total_ms", C.c_double), ("n_launches", C.c_int64)]


# every symbol include/abcdez_cuda.h declares (tests check the library exports all of them)
EXPORTS = [
    "abcdez_init", "abcdez_init_multi", "abcdez_ctx_gpus", "abcdez_destroy", "abcdez_version", "abcdez_last_error", "abcdez_sync",
    "abcdez_nccl_unique_id", "abcdez_comm_init", "abcdez_shard_range", "abcdez_comm_selftest",
    "abcdez_prior_create", "abcdez_prior_destroy", "abcdez_prior_sample", "abcdez_prior_logpdf", "abcdez_prior_push",
    "abcdez_model_count", "abcdez_model_name", "abcdez_model_lookup", "abcdez_model_info", "abcdez_model_bind",
    "abcdez_model_destroy", "abcdez_model_compile", "abcdez_simulate", "abcdez_kernel_pdf", "abcdez_kernel_logpdf",
    "abcdez_smc_opts_default", "abcdez_smc_run", "abcdez_smc_state_bytes", "abcdez_smc_run_state", "abcdez_smc_run_batch",
    "abcdez_posterior_sample",
    "abcdez_mc_opts_default", "abcdez_mc_run",
    "abcdez_pop_create", "abcdez_pop_destroy", "abcdez_pop_upload", "abcdez_pop_download", "abcdez_pop_set",
    "abcdez_pop_init", "abcdez_pop_smc_sweep", "abcdez_pop_mc_sweep", "abcdez_pop_eps_quantile",
    "abcdez_pop_reweight", "abcdez_pop_head", "abcdez_pop_resample", "abcdez_wsample_stratified", "abcdez_pop_last_timing",
    "abcdez_pop_bench_sweeps",
]

_lib = None


def lib():
    """Load libabcdez_cuda.so; fails loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ABCdeZError(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python abcdez.jl_b200/build.py` "
                                        "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.abcdez_last_error.restype = C.c_char_p
        L.abcdez_model_name.restype = C.c_char_p
        L.abcdez_kernel_pdf.restype = C.c_double
        L.abcdez_kernel_pdf.argtypes = [C.c_int, C.c_double, C.c_double]
        L.abcdez_kernel_logpdf.restype = C.c_double
        L.abcdez_kernel_logpdf.argtypes = [C.c_int, C.c_double, C.c_double]
        _lib = L
    return _lib


def _check(rc: int):
    if rc != OK:
        raise ABCdeZError(rc, lib().abcdez_last_error().decode("utf-8", "replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# univariate marginals (the subset of Distributions.jl the reference's tests and configs use)
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class _Marginal:
    family = -1
    discrete = False

    def params(self):
        raise NotImplementedError


@dataclass(frozen=True)
class Normal(_Marginal):
    mu: float = 0.0
    sigma: float = 1.0
    family = 0

    def params(self):
        return (self.mu, self.sigma)


@dataclass(frozen=True)
class Uniform(_Marginal):
    a: float = 0.0
    b: float = 1.0
    family = 1

    def params(self):
        return (self.a, self.b)


@dataclass(frozen=True)
class DiscreteUniform(_Marginal):
    a: int = 0
    b: int = 1
    family = 2
    discrete = True

    def params(self):
        return (float(self.a), float(self.b))


@dataclass(frozen=True)
class LogNormal(_Marginal):
    mu: float = 0.0
    sigma: float = 1.0
    family = 3

    def params(self):
        return (self.mu, self.sigma)


@dataclass(frozen=True)
class Exponential(_Marginal):
    scale: float = 1.0
    family = 4

    def params(self):
        return (self.scale,)


@dataclass(frozen=True)
class Gamma(_Marginal):
    shape: float = 1.0
    scale: float = 1.0
    family = 5

    def params(self):
        return (self.shape, self.scale)


@dataclass(frozen=True)
class Beta(_Marginal):
    alpha: float = 1.0
    beta: float = 1.0
    family = 6

    def params(self):
        return (self.alpha, self.beta)


@dataclass(frozen=True)
class NegativeBinomial(_Marginal):
    r: float = 1.0
    p: float = 0.5
    family = 7
    discrete = True

    def params(self):
        return (self.r, self.p)


# ---------------------------------------------------------------------------------------------
# context (one per process and GPU)
# ---------------------------------------------------------------------------------------------
class Context:
    def __init__(self, device: int = 0, stream: Optional[int] = None):
        self._h = C.c_void_p()
        _check(lib().abcdez_init(int(device), C.c_void_p(stream) if stream else None, C.byref(self._h)))
        self.device = device
        self.rank, self.world = 0, 1
        self.n_gpus = 1

    @classmethod
    def multi(cls, n_gpus: int = 0, device_ids: Optional[Sequence[int]] = None) -> "Context":
        """Single-process multi-GPU context (abcdez_init_multi): `abcdesmc` / `abcdemc` on it run ONE population
        sharded over the GPUs and return the whole population.  n_gpus = 0: every visible GPU."""
        self = cls.__new__(cls)
        self._h = C.c_void_p()
        ids = None if device_ids is None else np.ascontiguousarray(device_ids, dtype=np.int32)
        if ids is not None:
            n_gpus = ids.size
        _check(lib().abcdez_init_multi(int(n_gpus), _p(ids), C.byref(self._h)))
        self.device = int(ids[0]) if ids is not None else 0
        self.rank, self.world = 0, 1                     # one caller: the result is the whole population
        self.n_gpus = int(lib().abcdez_ctx_gpus(self._h))
        return self

    # ---- sharded runs (one process per GPU), include/abcdez_cuda.h "sharded runs" -----------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().abcdez_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: Optional[bytes]):
        """Collective.  After it abcdesmc/abcdemc on this context run ONE population sharded over `world` GPUs."""
        if world > 1 and (unique_id is None or len(unique_id) != 128):
            raise ABCdeZError(ERR_BAD_ARG, "comm_init: unique_id must be the 128 bytes of Context.nccl_unique_id()")
        _check(lib().abcdez_comm_init(self._h, int(rank), int(world), unique_id))
        self.rank, self.world = int(rank), int(world)

    def comm_selftest(self, rounds: int = 16, mode: int = 1):
        chk = C.c_uint64(); us = C.c_double()
        _check(lib().abcdez_comm_selftest(self._h, int(rounds), int(mode), C.byref(chk), C.byref(us)))
        return chk.value, us.value

    def sync(self):
        _check(lib().abcdez_sync(self._h))

    def close(self):
        if self._h:
            lib().abcdez_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def shard_range(N: int, rank: int, world: int):
    """The contiguous block [lo, hi) of global particle indices owned by `rank` (abcdez_shard_range)."""
    lo, hi = C.c_int64(), C.c_int64()
    _check(lib().abcdez_shard_range(C.c_int64(N), int(rank), int(world), C.byref(lo), C.byref(hi)))
    return lo.value, hi.value


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        _default_ctx = Context(dev)
    return _default_ctx


_multi_ctx: Optional[Context] = None


def default_multi_context() -> Context:
    """`parallel=true` (src/abcdez_smc.jl:237): every GPU visible to this process, one population across them."""
    global _multi_ctx
    if _multi_ctx is None:
        _multi_ctx = Context.multi(0)
    return _multi_ctx


# ---------------------------------------------------------------------------------------------
# Factored, src/abcdez_priors.jl:18-61
# ---------------------------------------------------------------------------------------------
class Factored:
    """`Factored(dists...)`: independent product of univariate marginals."""

    def __init__(self, *dists: _Marginal):
        if len(dists) == 1 and isinstance(dists[0], (list, tuple)):
            dists = tuple(dists[0])
        for d in dists:
            if not isinstance(d, _Marginal):
                raise ABCdeZError(ERR_UNSUPPORTED, f"unsupported marginal {d!r}")
        self.p = tuple(dists)
        self._h = None
        self._ctx = None

    def __len__(self):                      # length(p::Factored) = N, src/abcdez_priors.jl:61
        return len(self.p)

    @property
    def discrete(self):
        return np.array([d.discrete for d in self.p], dtype=bool)

    def _arrays(self):
        d = len(self.p)
        fam = np.array([m.family for m in self.p], dtype=np.int32)
        par = np.zeros((d, 4))
        for k, m in enumerate(self.p):
            q = m.params()
            par[k, :len(q)] = q
        return d, fam, par

    def handle(self, ctx: Context):
        if self._h is None or self._ctx is not ctx:
            d, fam, par = self._arrays()
            h = C.c_void_p()
            _check(lib().abcdez_prior_create(ctx._h, d, _p(fam), _p(par), C.byref(h)))
            self._h, self._ctx = h, ctx
        return self._h

    # the Distributions-style interface the reference extends (src/abcdez_priors.jl:27-54)
    def logpdf(self, x, ctx: Optional[Context] = None):
        """logpdf(prior, push_p(prior, x)); x is one point or an (N, d) array."""
        ctx = ctx or default_context()
        th = _f64(x).reshape(-1, len(self))
        out = np.empty(th.shape[0])
        _check(lib().abcdez_prior_logpdf(ctx._h, self.handle(ctx), C.c_int64(th.shape[0]), _p(th), _p(out)))
        return out if np.ndim(x) > 1 else float(out[0])

    def pdf(self, x, ctx: Optional[Context] = None):
        return np.exp(self.logpdf(x, ctx))

    def rand(self, n: Optional[int] = None, seed: int = 0, epoch: int = 0, id0: int = 0, ctx: Optional[Context] = None):
        ctx = ctx or default_context()
        N = 1 if n is None else int(n)
        th = np.empty((N, len(self)))
        _check(lib().abcdez_prior_sample(ctx._h, self.handle(ctx), C.c_int64(N), C.c_uint64(seed), C.c_uint32(epoch),
                                         C.c_int64(id0), _p(th)))
        return th[0] if n is None else th

    def push_p(self, x, ctx: Optional[Context] = None):
        """push_p(prior, x), src/abcdez_types.jl:20-23."""
        ctx = ctx or default_context()
        th = _f64(x).reshape(-1, len(self))
        out = np.empty_like(th)
        _check(lib().abcdez_prior_push(ctx._h, self.handle(ctx), C.c_int64(th.shape[0]), _p(th), _p(out)))
        return out if np.ndim(x) > 1 else out[0]


def _as_prior(prior) -> tuple["Factored", bool]:
    """The reference also accepts a bare univariate Distribution (d = 1, scalar theta)."""
    if isinstance(prior, Factored):
        return prior, False
    if isinstance(prior, _Marginal):
        return Factored(prior), True
    raise ABCdeZError(ERR_UNSUPPORTED, f"unsupported prior {prior!r}: use Factored(...) of the supported marginals")


# ---------------------------------------------------------------------------------------------
# ABC kernels, src/abcdez_types.jl:26-73
# ---------------------------------------------------------------------------------------------
class _ABCKernel:
    kind = -1

    def __init__(self, eps: float):
        if not (eps >= 0.0):
            raise ABCdeZError(ERR_BAD_ARG, "Expected ϵ ≥ 0.0")       # src/abcdez_types.jl:30
        self.eps = float(eps)
        self.ϵ = self.eps

    def pdf(self, x: float) -> float:
        return lib().abcdez_kernel_pdf(self.kind, self.eps, float(x))

    def logpdf(self, x: float) -> float:
        return lib().abcdez_kernel_logpdf(self.kind, self.eps, float(x))


class Indicator0toEps(_ABCKernel):
    kind = 0


class IndicatorStrict0toEps(_ABCKernel):
    kind = 1


class Epa0toEps(_ABCKernel):
    kind = 2


class EpaStrict0toEps(_ABCKernel):
    kind = 3


# the reference's spellings
Indicator0toϵ = Indicator0toEps
IndicatorStrict0toϵ = IndicatorStrict0toEps
Epa0toϵ = Epa0toEps
EpaStrict0toϵ = EpaStrict0toEps
_KERNELS = {"indicator": 0, "indicator_strict": 1, "epa": 2, "epa_strict": 3}


def _kernel_kind(k) -> int:
    if isinstance(k, str):
        return _KERNELS[k]
    if isinstance(k, int):
        return k
    if isinstance(k, type) and issubclass(k, _ABCKernel):
        return k.kind
    raise ABCdeZError(ERR_BAD_ARG, f"unknown ABC kernel {k!r}")


# ---------------------------------------------------------------------------------------------
# models: dist!(theta, ve) -> (d, blob) as registered device functors
# ---------------------------------------------------------------------------------------------
def model_names():
    L = lib()
    return [L.abcdez_model_name(i).decode() for i in range(L.abcdez_model_count())]


def compile_model(name: str, struct_name: str, cuda_src: str, d: int, blob_bytes: int = 0,
                  ctx: Optional["Context"] = None, load: bool = True) -> str:
    """Runtime-supplied `dist!`: compile the CUDA source of one model struct (include/abcdez_cuda.h,
    abcdez_model_compile) and register it under `name`; afterwards `Model(name, data)` works like a built-in.
    load=False: compile only (no GPU needed).  Returns the NVRTC log."""
    log = C.create_string_buffer(1 << 16)
    mid = C.c_int(-1)
    h = None
    if load:
        ctx = ctx or default_context()
        h = ctx._h
    _check(lib().abcdez_model_compile(h, name.encode(), struct_name.encode(), cuda_src.encode(), int(d), int(blob_bytes),
                                      C.byref(mid), log, C.c_size_t(len(log))))
    return log.value.decode("utf-8", "replace")


class Model:
    """A registered CUDA simulator bound to observed data (the closure `dist!` captures `data`)."""

    def __init__(self, name: str, data: Sequence[float] = ()):
        self.name = name
        self.data = _f64(np.asarray(data, dtype=np.float64).ravel())
        mid = C.c_int()
        _check(lib().abcdez_model_lookup(name.encode(), C.byref(mid)))
        self.id = mid.value
        d = C.c_int(); b = C.c_int()
        _check(lib().abcdez_model_info(self.id, C.byref(d), C.byref(b)))
        self.d, self.blob_bytes = d.value, b.value
        self._h = None
        self._ctx = None

    def handle(self, ctx: Context):
        if self._h is None or self._ctx is not ctx:
            h = C.c_void_p()
            _check(lib().abcdez_model_bind(ctx._h, self.id, _p(self.data), C.c_size_t(self.data.size), C.byref(h)))
            self._h, self._ctx = h, ctx
        return self._h

    def simulate(self, theta_pushed, seed=0, epoch=0, tag=TAG_MODEL, id0=0, ctx: Optional[Context] = None):
        ctx = ctx or default_context()
        th = _f64(theta_pushed).reshape(-1, self.d)
        N = th.shape[0]
        dist = np.empty(N)
        blobs = np.zeros((N, max(self.blob_bytes, 1)), dtype=np.uint8)
        _check(lib().abcdez_simulate(ctx._h, self.handle(ctx), C.c_int64(N), _p(th), C.c_uint64(seed), C.c_uint32(epoch),
                                     C.c_uint32(tag), C.c_int64(id0), _p(dist), _p(blobs)))
        return dist, blobs[:, :self.blob_bytes]


# ---------------------------------------------------------------------------------------------
# results
# ---------------------------------------------------------------------------------------------
@dataclass
class SMCResult:
    """NamedTuple of src/abcdez_smc.jl:388-393 (Greek field names available as aliases)."""
    P: np.ndarray
    Wns: np.ndarray
    C: np.ndarray
    eps: float
    logZ: float
    blobs: np.ndarray
    eps_hist: Optional[np.ndarray] = None
    ranges_eps: Optional[np.ndarray] = None
    logZs: Optional[np.ndarray] = None
    esss: Optional[np.ndarray] = None
    faccs: Optional[np.ndarray] = None
    gamma0s: Optional[np.ndarray] = None
    Kmcmcs: Optional[np.ndarray] = None
    iters: int = 0
    nsims: int = 0
    status: int = 0
    stats: dict = field(default_factory=dict)
    state: Optional[bytes] = None          # run-state snapshot (return_state=True), see abcdesmc

    ϵ = property(lambda s: s.eps)
    ϵs = property(lambda s: s.eps_hist)
    ranges_ϵ = property(lambda s: s.ranges_eps)
    γ0s = property(lambda s: s.gamma0s)


@dataclass
class MCResult:
    """NamedTuple of src/abcdez_mc.jl:171."""
    P: np.ndarray
    C: np.ndarray
    reached_eps: bool
    blobs: np.ndarray
    nsims: int = 0
    stats: dict = field(default_factory=dict)

    reached_ϵ = property(lambda s: s.reached_eps)


def _greek(kw: dict, mapping: dict):
    for g, a in mapping.items():
        if g in kw:
            kw[a] = kw.pop(g)
    return kw


def _seed_from(rng) -> int:
    """`rng` -> a 64-bit Philox key."""
    if rng is None:
        return int.from_bytes(os.urandom(8), "little")
    if isinstance(rng, (int, np.integer)):
        return int(rng) & 0xFFFFFFFFFFFFFFFF
    if isinstance(rng, np.random.Generator):
        return int(rng.integers(0, 2**63 - 1))
    raise ABCdeZError(ERR_BAD_ARG, "rng must be None, an int seed or a numpy Generator")


def _blob_view(raw: np.ndarray, B: int):
    return raw[:, :B] if B else np.empty((raw.shape[0], 0), dtype=np.uint8)


def abcdesmc(prior, dist, eps_target, varexternal=None, *, nparticles: int = 100, alpha=0.95, delta_ess=0.5,
             nsims_max: int = 10**7, Kmcmc: int = 3, Kmcmc_min=1.0, ABCk=IndicatorStrict0toEps, facc_stop=0.0,
             facc_min=0.0, facc_tune=0.975, verbose: bool = True, verboseout: bool = True, rng=None,
             parallel: bool = False, ctx: Optional[Context] = None, max_iters: int = 0, exact_scan: bool = False,
             profile: bool = False, sync_every: int = 1, hist_cap: int = 8192, fused_head: bool = True,
             state=None, return_state: bool = False, systematic_resampling: bool = False, partner_segments: bool = False, fp32_state: bool = False,
             **greek) -> SMCResult:
    """`abcdesmc!(prior, dist!, ϵ_target, varexternal; kwargs...)`, src/abcdez_smc.jl:215-394.

    `dist` is a :class:`Model`; `varexternal` is accepted for signature compatibility (the device
    functors keep their scratch in registers); `parallel=True` without an explicit `ctx` runs the population
    sharded over every GPU visible to the process (`Context.multi`; one GPU: the plain single-GPU run).

    Run-state snapshots (not in the reference): `return_state=True` attaches the state the run ended in
    (`result.state`, bytes; typically after `max_iters` iterations) and `state=<bytes>` continues from one --
    with this call's `eps_target`, `nsims_max`, `facc_stop`, `max_iters` and the snapshot's seed.
    """
    kw = _greek(dict(greek), {"α": "alpha", "δess": "delta_ess"})
    alpha = kw.pop("alpha", alpha); delta_ess = kw.pop("delta_ess", delta_ess)
    if kw:
        raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
    if not isinstance(dist, Model):
        raise ABCdeZError(ERR_UNSUPPORTED, "dist! must be a registered device functor: abcdez Model(name, data)")
    ctx = ctx or (default_multi_context() if parallel else default_context())
    fprior, scalar = _as_prior(prior)
    if Kmcmc_min <= facc_min and verbose:
        print("Warning: Kmcmc_min should be larger than facc_min")            # src/abcdez_smc.jl:232
    N, d, B = int(nparticles), len(fprior), dist.blob_bytes
    o = _SmcOpts()
    lib().abcdez_smc_opts_default(C.byref(o))
    Nglobal = N
    if ctx.world > 1:                         # sharded: this rank returns the rows of its own block
        lo, hi = shard_range(Nglobal, ctx.rank, ctx.world)
        N = hi - lo
    o.nparticles = Nglobal; o.alpha = alpha; o.delta_ess = delta_ess; o.nsims_max = int(nsims_max); o.Kmcmc = int(Kmcmc)
    o.Kmcmc_min = float(Kmcmc_min); o.kernel = _kernel_kind(ABCk); o.facc_stop = facc_stop; o.facc_min = facc_min
    o.facc_tune = facc_tune; o.seed = _seed_from(rng); o.verboseout = int(verboseout); o.max_iters = int(max_iters)
    o.exact_scan = int(exact_scan); o.profile = int(profile); o.sync_every = int(sync_every)
    o.fused_head = int(fused_head)
    o.systematic_resampling = int(systematic_resampling); o.partner_segments = int(partner_segments)   # relaxed-parity modes
    o.fp32_state = int(fp32_state)
    Np = max(N, 1)
    P = np.empty((Np, d)); W = np.empty(Np); Cc = np.empty(Np); bl = np.zeros((Np, max(B, 1)), dtype=np.uint8)
    h = {k: np.zeros(hist_cap) for k in ("eps", "dmin", "dmax", "logZ", "ess", "facc", "gamma0")}
    hK = np.zeros(hist_cap, dtype=np.int32)
    r = _SmcResult()
    r.P, r.Wns, r.C, r.blobs = _p(P), _p(W), _p(Cc), _p(bl)
    r.hist_cap = hist_cap if verboseout else 0
    r.h_eps, r.h_dmin, r.h_dmax, r.h_logZ = _p(h["eps"]), _p(h["dmin"]), _p(h["dmax"]), _p(h["logZ"])
    r.h_ess, r.h_facc, r.h_gamma0, r.h_Kmcmc = _p(h["ess"]), _p(h["facc"]), _p(h["gamma0"]), _p(hK)
    state_out = None
    if state is None and not return_state:
        _check(lib().abcdez_smc_run(ctx._h, fprior.handle(ctx), dist.handle(ctx), C.c_double(eps_target), C.byref(o),
                                    C.byref(r)))
    else:
        L = lib()
        L.abcdez_smc_state_bytes.restype = C.c_int64
        need = L.abcdez_smc_state_bytes(fprior.handle(ctx), dist.handle(ctx), C.c_int64(N), C.c_int32(r.hist_cap))
        sin = np.frombuffer(state, dtype=np.uint8) if state is not None else None
        sout = np.empty(need, dtype=np.uint8) if return_state else None
        nout = C.c_int64(0)
        _check(L.abcdez_smc_run_state(ctx._h, fprior.handle(ctx), dist.handle(ctx), C.c_double(eps_target), C.byref(o),
                                      C.byref(r), _p(sin) if sin is not None else None,
                                      C.c_int64(sin.size if sin is not None else 0),
                                      _p(sout) if sout is not None else None, C.c_int64(need if return_state else 0),
                                      C.byref(nout)))
        if return_state:
            state_out = sout[:nout.value].tobytes()
    if r.status == ERR_NO_ALIVE and verbose:
        print("Warning: No alive particles")                                 # src/abcdez_smc.jl:375
    stats = dict(n_resamples=r.n_resamples, n_sweeps=r.n_sweeps, n_launches=r.n_launches, sweep_ms=r.sweep_ms,
                 total_ms=r.total_ms, init_ms=r.init_ms, head_ms=r.head_ms, resample_ms=r.resample_ms, seed=o.seed, rank=ctx.rank, world=ctx.world,
                 nparticles=Nglobal, hist_dropped=r.hist_dropped)
    if r.hist_dropped and verbose:
        print(f"Warning: {r.hist_dropped} history records did not fit hist_cap={hist_cap}; the histories are truncated")
    Pout = P[:, 0] if scalar else P
    out = SMCResult(Pout, W, Cc, r.eps, r.logZ, _blob_view(bl, B), iters=r.iters, nsims=r.nsims, status=r.status,
                    stats=stats)
    out.state = state_out
    if verboseout:
        n = r.hist_len
        out.eps_hist = h["eps"][:n]; out.ranges_eps = np.stack([h["dmin"][:n], h["dmax"][:n]], axis=1)
        out.logZs = h["logZ"][:n]; out.esss = h["ess"][:n]; out.faccs = h["facc"][:n]
        out.gamma0s = h["gamma0"][:n]; out.Kmcmcs = hK[:n]
    if verbose:
        print(f"Final run: iteration = {r.iters} nsim = {r.nsims} ϵ = {r.eps} logZ = {r.logZ}")
    return out


def abcdemc(prior, dist, eps_target, varexternal=None, *, nparticles: int = 50, generations: int = 20,
            verbose: bool = True, rng=None, parallel: bool = False, ctx: Optional[Context] = None) -> MCResult:
    """`abcdemc!(prior, dist!, ϵ_target, varexternal; kwargs...)`, src/abcdez_mc.jl:102-172."""
    if not isinstance(dist, Model):
        raise ABCdeZError(ERR_UNSUPPORTED, "dist! must be a registered device functor: abcdez Model(name, data)")
    ctx = ctx or (default_multi_context() if parallel else default_context())
    fprior, scalar = _as_prior(prior)
    N, d, B = int(nparticles), len(fprior), dist.blob_bytes
    o = _McOpts(N, int(generations), _seed_from(rng))
    Nglobal = N
    if ctx.world > 1:                         # sharded: this rank returns the rows of its own block
        lo, hi = shard_range(Nglobal, ctx.rank, ctx.world)
        N = hi - lo
    Np = max(N, 1)
    P = np.empty((Np, d)); Cc = np.empty(Np); bl = np.zeros((Np, max(B, 1)), dtype=np.uint8)
    r = _McResult()
    r.P, r.C, r.blobs = _p(P), _p(Cc), _p(bl)
    _check(lib().abcdez_mc_run(ctx._h, fprior.handle(ctx), dist.handle(ctx), C.c_double(eps_target), C.byref(o),
                               C.byref(r)))
    if verbose:
        print(f"End: converged = {bool(r.reached_eps)} nsim = {r.nsims} range_ϵ = ({r.dmin}, {r.dmax})")
    stats = dict(total_ms=r.total_ms, n_launches=r.n_launches, seed=o.seed, rank=ctx.rank, world=ctx.world,
                 nparticles=Nglobal)
    return MCResult(P[:, 0] if scalar else P, Cc, bool(r.reached_eps), _blob_view(bl, B), nsims=r.nsims, stats=stats)


# the reference's names with the bang are not valid Python identifiers; expose them via getattr
globals()["abcdesmc!"] = abcdesmc
globals()["abcdemc!"] = abcdemc


def wsample_stratified(weights, uniforms, mode: int = 1, ctx: Optional[Context] = None):
    """`ABCdeZ.wsample_stratified!(rng, weights, inds)` (src/abcdez_smc.jl:15-56) with the uniforms
    supplied by the caller; returns 1-based indices.  mode 2 = sequential (bit-exact) cumsum."""
    ctx = ctx or default_context()
    w = _f64(weights); u = _f64(uniforms)
    inds = np.empty(w.size, dtype=np.int64)
    _check(lib().abcdez_wsample_stratified(ctx._h, C.c_int64(w.size), _p(w), _p(u), int(mode), _p(inds)))
    return inds


# ---------------------------------------------------------------------------------------------
# after the run: what the reference's tests, example and docs do with a result
# ---------------------------------------------------------------------------------------------
def weightinds(weights, rng=None, ctx: Optional[Context] = None):
    """`weightinds` of test/runtests.jl:13-19: stratified resampling indices (0-based here) of normalised weights,
    through the library's `wsample_stratified!` kernel; `rng` seeds the per-stratum uniforms."""
    w = _f64(weights)
    if not abs(w.sum() - 1.0) < 1e-8:
        raise ABCdeZError(ERR_BAD_ARG, "Sum of weights expected to be 1.0 (approximately)")
    u = np.random.default_rng(_seed_from(rng)).random(w.size)
    return np.clip(wsample_stratified(w, u, ctx=ctx), 1, w.size) - 1


def posterior_sample(result: "SMCResult", rng=None, ctx: Optional[Context] = None):
    """Equally weighted posterior sample of an `abcdesmc!` result: `P[weightinds(Wns)]` (test/runtests.jl:287-291),
    resampled and gathered on the device (abcdez_posterior_sample)."""
    ctx = ctx or default_context()
    P = _f64(result.P); P2 = P.reshape(P.shape[0], -1)
    W = _f64(result.Wns)
    out = np.empty_like(P2)
    _check(lib().abcdez_posterior_sample(ctx._h, C.c_int64(P2.shape[0]), int(P2.shape[1]), _p(P2), _p(W), C.c_uint64(_seed_from(rng)),
                                         _p(out), None))
    return out.reshape(P.shape)


def evidence_at(result: "SMCResult", eps: float) -> float:
    """log evidence of a `verboseout` run at tolerance `eps` from its ladder (`logZs`, `ϵs`): the record with the
    smallest ϵ >= eps, i.e. the docs' workaround 1 for comparing models whose runs ended at different ϵ
    (docs/src/index.md:282-284)."""
    if result.eps_hist is None:
        raise ABCdeZError(ERR_BAD_ARG, "evidence_at needs a run with verboseout=true")
    e = np.asarray(result.eps_hist); ok = np.nonzero(e >= eps)[0]
    if ok.size == 0:
        raise ABCdeZError(ERR_BAD_ARG, "eps is above the whole ladder of this run")
    return float(np.asarray(result.logZs)[ok[-1]])


def model_probabilities(results_or_logZs, prior_probs=None, eps: Optional[float] = None):
    """Posterior model probabilities (examples/minimal_example.jl:58-65) from log evidences or `abcdesmc!` results.
    Results are compared at a common tolerance: `eps`, default the largest final ϵ among them (evidences at
    different ϵ carry different kernel normalisations, docs/src/index.md:202-212)."""
    items = list(results_or_logZs)
    if items and isinstance(items[0], SMCResult):
        at = max(r.eps for r in items) if eps is None else eps
        lz = np.array([r.logZ if (r.eps == at and eps is None) else evidence_at(r, at) for r in items])
    else:
        lz = np.asarray(items, dtype=float)
    pri = np.full(lz.size, 1.0 / lz.size) if prior_probs is None else np.asarray(prior_probs, dtype=float)
    w = np.exp(lz - lz.max()) * pri
    return w / w.sum()


def abcdesmc_batch(runs, *, ctx: Optional[Context] = None, verboseout: bool = False, hist_cap: int = 8192):
    """Several independent `abcdesmc!` runs in flight together on one GPU (abcdez_smc_run_batch): replicates of one model
    for the evidence uncertainty, several models for a comparison.  `runs` is a list of dicts with the arguments of
    :func:`abcdesmc` (`prior`, `dist`, `eps_target` and keyword options); returns the list of results -- each exactly
    what its own `abcdesmc` call returns."""
    ctx = ctx or default_context()
    n = len(runs)
    opts = (_SmcOpts * n)(); res = (_SmcResult * n)()
    ph = (C.c_void_p * n)(); mh = (C.c_void_p * n)(); eps = (C.c_double * n)(); st = (C.c_int * n)()
    keep = []
    for i, r in enumerate(runs):
        r = dict(r)
        fprior, scalar = _as_prior(r.pop("prior")); dist = r.pop("dist"); eps[i] = float(r.pop("eps_target"))
        N = int(r.get("nparticles", 100)); d, B = len(fprior), dist.blob_bytes
        o = opts[i]
        lib().abcdez_smc_opts_default(C.byref(o))
        o.nparticles = N; o.alpha = r.get("alpha", 0.95); o.delta_ess = r.get("delta_ess", 0.5); o.nsims_max = int(r.get("nsims_max", 10**7))
        o.Kmcmc = int(r.get("Kmcmc", 3)); o.Kmcmc_min = float(r.get("Kmcmc_min", 1.0)); o.kernel = _kernel_kind(r.get("ABCk", IndicatorStrict0toEps))
        o.facc_stop = r.get("facc_stop", 0.0); o.facc_min = r.get("facc_min", 0.0); o.facc_tune = r.get("facc_tune", 0.975)
        o.seed = _seed_from(r.get("rng")); o.verboseout = int(verboseout)
        o.systematic_resampling = int(r.get("systematic_resampling", False)); o.partner_segments = int(r.get("partner_segments", False))
        o.fp32_state = int(r.get("fp32_state", False))                        # relaxed-parity modes
        P = np.empty((N, d)); W = np.empty(N); Cc = np.empty(N); bl = np.zeros((N, max(B, 1)), dtype=np.uint8)
        h = {k: np.zeros(hist_cap) for k in ("eps", "dmin", "dmax", "logZ", "ess", "facc", "gamma0")}; hK = np.zeros(hist_cap, dtype=np.int32)
        q = res[i]
        q.P, q.Wns, q.C, q.blobs = _p(P), _p(W), _p(Cc), _p(bl)
        q.hist_cap = hist_cap if verboseout else 0
        q.h_eps, q.h_dmin, q.h_dmax, q.h_logZ = _p(h["eps"]), _p(h["dmin"]), _p(h["dmax"]), _p(h["logZ"])
        q.h_ess, q.h_facc, q.h_gamma0, q.h_Kmcmc = _p(h["ess"]), _p(h["facc"]), _p(h["gamma0"]), _p(hK)
        ph[i] = fprior.handle(ctx); mh[i] = dist.handle(ctx)
        keep.append((P, W, Cc, bl, h, hK, scalar, B, fprior, dist))
    _check(lib().abcdez_smc_run_batch(ctx._h, n, ph, mh, eps, opts, res, st))
    out = []
    for i, (P, W, Cc, bl, h, hK, scalar, B, _, _) in enumerate(keep):
        q = res[i]
        r = SMCResult(P[:, 0] if scalar else P, W, Cc, q.eps, q.logZ, _blob_view(bl, B), iters=q.iters, nsims=q.nsims, status=q.status,
                      stats=dict(n_resamples=q.n_resamples, n_sweeps=q.n_sweeps, n_launches=q.n_launches, total_ms=q.total_ms, seed=opts[i].seed))
        if verboseout:
            m = q.hist_len
            r.eps_hist = h["eps"][:m]; r.ranges_eps = np.stack([h["dmin"][:m], h["dmax"][:m]], axis=1); r.logZs = h["logZ"][:m]
            r.esss = h["ess"][:m]; r.faccs = h["facc"][:m]; r.gamma0s = h["gamma0"][:m]; r.Kmcmcs = hK[:m]
        out.append(r)
    return out


def evidence_uncertainty(prior, dist, eps_target, repeats: int = 8, rng=None, ctx: Optional[Context] = None, **kw):
    """The docs' advice (docs/src/index.md:214-220): repeat `abcdesmc!` and summarise the log evidences.
    Returns (mean, std, list of logZ); each repeat uses its own Philox key.  The repeats run as ONE batch
    (abcdez_smc_run_batch): their kernels overlap on the GPU instead of queueing run after run."""
    base = _seed_from(rng)
    runs = [dict(prior=prior, dist=dist, eps_target=eps_target, rng=(base + 0x9E3779B97F4A7C15 * (k + 1)) % 2**64, **kw)
            for k in range(int(repeats))]
    lz = [r.logZ for r in abcdesmc_batch(runs, ctx=ctx)]
    return float(np.mean(lz)), float(np.std(lz, ddof=1)) if len(lz) > 1 else 0.0, lz


# ---------------------------------------------------------------------------------------------
# device-resident population: the stage-level entry points
# ---------------------------------------------------------------------------------------------
class Population:
    def __init__(self, prior, model: Model, N: int, id0: int = 0, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.prior, _ = _as_prior(prior)
        self.model = model
        self.N, self.d, self.B = int(N), len(self.prior), model.blob_bytes
        self._h = C.c_void_p()
        _check(lib().abcdez_pop_create(self.ctx._h, self.prior.handle(self.ctx), model.handle(self.ctx),
                                       C.c_int64(N), C.c_int64(id0), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().abcdez_pop_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, theta=None, logpi=None, delta=None, blobs=None, W=None, alive=None):
        th = None if theta is None else _f64(theta).reshape(self.N, self.d)
        lp = None if logpi is None else _f64(logpi)
        dl = None if delta is None else _f64(delta)
        bl = None if blobs is None or not self.B else np.ascontiguousarray(blobs, dtype=np.uint8)
        w = None if W is None else _f64(W)
        al = None if alive is None else np.ascontiguousarray(alive, dtype=np.uint8)
        _check(lib().abcdez_pop_upload(self._h, _p(th), _p(lp), _p(dl), _p(bl), _p(w), _p(al)))

    def download(self):
        th = np.empty((self.N, self.d)); lp = np.empty(self.N); dl = np.empty(self.N)
        bl = np.zeros((self.N, max(self.B, 1)), dtype=np.uint8); w = np.empty(self.N)
        al = np.empty(self.N, dtype=np.uint8)
        _check(lib().abcdez_pop_download(self._h, _p(th), _p(lp), _p(dl), _p(bl) if self.B else None, _p(w), _p(al)))
        return dict(theta=th, logpi=lp, delta=dl, blobs=_blob_view(bl, self.B), W=w, alive=al)

    def set(self, eps, eps_prev=math.inf, kernel="indicator_strict", gamma0=None, gsig=1e-5, seed=0, epoch=0):
        if gamma0 is None:
            gamma0 = 2.38 / math.sqrt(2 * self.d)
        _check(lib().abcdez_pop_set(self._h, C.c_double(eps), C.c_double(eps_prev), _kernel_kind(kernel),
                                    C.c_double(gamma0), C.c_double(gsig), C.c_uint64(seed), C.c_uint32(epoch)))

    def init(self, seed, draw_prior=True):
        n = C.c_int64()
        _check(lib().abcdez_pop_init(self._h, C.c_uint64(seed), int(draw_prior), C.byref(n)))
        return n.value

    def smc_sweep(self, a=None, b=None, z=None, u=None, want_flags=True):
        i32 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.int32)
        ia, ib = i32(a), i32(b)
        iz = None if z is None else _f64(z)
        iu = None if u is None else _f64(u)
        flags = np.zeros(self.N, dtype=np.uint8) if want_flags else None
        ns = C.c_int64(); na = C.c_int64()
        _check(lib().abcdez_pop_smc_sweep(self._h, _p(ia), _p(ib), _p(iz), _p(iu), _p(flags), C.byref(ns), C.byref(na)))
        return dict(flags=flags, nsims=ns.value, naccs=na.value)

    def mc_sweep(self, eps_pop, eps_target, s=None, a=None, b=None, z=None, u=None, want_flags=True):
        i32 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.int32)
        is_, ia, ib = i32(s), i32(a), i32(b)
        iz = None if z is None else _f64(z)
        iu = None if u is None else _f64(u)
        flags = np.zeros(self.N, dtype=np.uint8) if want_flags else None
        ns = C.c_int64()
        _check(lib().abcdez_pop_mc_sweep(self._h, C.c_double(eps_pop), C.c_double(eps_target), _p(is_), _p(ia), _p(ib),
                                         _p(iz), _p(iu), _p(flags), C.byref(ns)))
        return dict(flags=flags, nsims=ns.value)

    def eps_quantile(self, alpha):
        q = C.c_double(); lo = C.c_double(); hi = C.c_double()
        _check(lib().abcdez_pop_eps_quantile(self._h, C.c_double(alpha), C.byref(q), C.byref(lo), C.byref(hi)))
        return q.value, lo.value, hi.value

    def reweight(self, eps_new):
        wn = C.c_double(); ess = C.c_double(); na = C.c_int64()
        _check(lib().abcdez_pop_reweight(self._h, C.c_double(eps_new), C.byref(wn), C.byref(ess), C.byref(na)))
        return wn.value, ess.value, na.value

    def head(self, alpha, eps_target=0.0):
        """One fused iteration head (head.cu): returns (q, eps, wnorm, ess, n_alive)."""
        q = C.c_double(); eps = C.c_double(); wn = C.c_double(); ess = C.c_double(); na = C.c_int64()
        _check(lib().abcdez_pop_head(self._h, C.c_double(alpha), C.c_double(eps_target), C.byref(q), C.byref(eps),
                                     C.byref(wn), C.byref(ess), C.byref(na)))
        return q.value, eps.value, wn.value, ess.value, na.value

    def resample(self, uniforms=None, epoch=0, mode=0):
        u = None if uniforms is None else _f64(uniforms)
        inds = np.empty(self.N, dtype=np.int32)
        _check(lib().abcdez_pop_resample(self._h, _p(u), C.c_uint32(epoch), int(mode), _p(inds)))
        return inds

    def bench_sweeps(self, sweeps: int):
        ns = C.c_int64(); na = C.c_int64(); ms = C.c_double()
        _check(lib().abcdez_pop_bench_sweeps(self._h, int(sweeps), C.byref(ns), C.byref(na), C.byref(ms)))
        return ns.value, na.value, ms.value

    def last_timing(self):
        ms = C.c_double(); nl = C.c_int64()
        _check(lib().abcdez_pop_last_timing(self._h, C.byref(ms), C.byref(nl)))
        return ms.value, nl.value
