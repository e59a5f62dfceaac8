"""ctypes binding of the CPU oracle (oracle/abcdez_oracle.c -> liborc.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
the cpu_baseline / --impl reference legs of bench.py.  The product package
(abcdez.jl_b200/) never imports this module.

The functions mirror the reference's internal functions one to one
(reference file:line in the C source); indices are 0-based here except where
stated.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAMILIES = {"normal": 0, "uniform": 1, "discrete_uniform": 2, "lognormal": 3,
            "exponential": 4, "gamma": 5, "beta": 6, "negbin": 7}
KERNELS = {"indicator": 0, "indicator_strict": 1, "epa": 2, "epa_strict": 3}
TAG_MODEL = 4
TAG_INIT_MODEL = 7
FLAG_SIM, FLAG_ACC = 1, 2


def build(force: bool = False) -> str:
    """Compile liborc.so with the committed Makefile (gcc only, a few seconds)."""
    so = os.path.join(_HERE, "liborc.so")
    src = os.path.join(_HERE, "abcdez_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s", "liborc.so"], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


class _SmcOpts(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("alpha", C.c_double), ("delta_ess", C.c_double),
                ("nsims_max", C.c_int64), ("Kmcmc", C.c_int32), ("Kmcmc_min", C.c_double),
                ("kind", C.c_int32), ("facc_stop", C.c_double), ("facc_min", C.c_double),
                ("facc_tune", C.c_double), ("seed", C.c_uint64), ("faithful_wsample", C.c_int32),
                ("max_iters", C.c_int32), ("islands", C.c_int32)]


class _SmcResult(C.Structure):
    _fields_ = [("eps", C.c_double), ("logZ", C.c_double), ("iters", C.c_int64), ("nsims", C.c_int64),
                ("status", C.c_int32), ("hist_len", C.c_int32), ("sweep_seconds", C.c_double),
                ("nsweeps", C.c_int64)]


class _McOpts(C.Structure):
    _fields_ = [("nparticles", C.c_int64), ("generations", C.c_int32), ("seed", C.c_uint64), ("islands", C.c_int32)]


class _McResult(C.Structure):
    _fields_ = [("reached_eps", C.c_int32), ("nsims", C.c_int64), ("dmin", C.c_double),
                ("dmax", C.c_double), ("sweep_seconds", C.c_double)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_kernel_pdf.restype = C.c_double
        _LIB.orc_kernel_pdf.argtypes = [C.c_int, C.c_double, C.c_double]
        _LIB.orc_kernel_logpdf.restype = C.c_double
        _LIB.orc_kernel_logpdf.argtypes = [C.c_int, C.c_double, C.c_double]
        _LIB.orc_model_name.restype = C.c_char_p
    return _LIB


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _prior_args(prior):
    """prior: list of (family_name, p0, p1, ...)"""
    d = len(prior)
    fam = np.array([FAMILIES[p[0]] for p in prior], dtype=np.int32)
    par = np.zeros((d, 4), dtype=np.float64)
    for k, p in enumerate(prior):
        par[k, :len(p) - 1] = p[1:]
    return d, fam, par


def model_id(name: str) -> int:
    L = lib()
    for i in range(L.orc_model_count()):
        if L.orc_model_name(i).decode() == name:
            return i
    raise KeyError(name)


def model_dim(name: str) -> int:
    return lib().orc_model_dim(model_id(name))


def model_blob(name: str) -> int:
    return lib().orc_model_blob(model_id(name))


def _data(data):
    buf = np.zeros(64, dtype=np.float64)
    data = np.asarray(data, dtype=np.float64).ravel()
    buf[:data.size] = data
    return buf


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox(_p(c), _p(k), _p(out))
    return out


def kernel_pdf(kind, eps, x):
    return lib().orc_kernel_pdf(KERNELS[kind], float(eps), float(x))


def kernel_logpdf(kind, eps, x):
    return lib().orc_kernel_logpdf(KERNELS[kind], float(eps), float(x))


def push(prior, theta):
    d, fam, _ = _prior_args(prior)
    th = _f64(theta).reshape(-1, d)
    out = np.empty_like(th)
    lib().orc_push(d, _p(fam), C.c_int64(th.shape[0]), _p(th), _p(out))
    return out


def prior_logpdf(prior, theta, raw=False):
    d, fam, par = _prior_args(prior)
    th = _f64(theta).reshape(-1, d)
    out = np.empty(th.shape[0])
    f = lib().orc_prior_logpdf_raw if raw else lib().orc_prior_logpdf
    f(d, _p(fam), _p(par), C.c_int64(th.shape[0]), _p(th), _p(out))
    return out


def prior_sample(prior, N, seed, epoch=0, id0=0):
    d, fam, par = _prior_args(prior)
    th = np.empty((N, d))
    lib().orc_prior_sample(d, _p(fam), _p(par), C.c_int64(N), C.c_uint64(seed), C.c_uint32(epoch),
                           C.c_int64(id0), _p(th))
    return th


def simulate(model, data, theta_pushed, seed, epoch=0, tag=TAG_MODEL, id0=0):
    mid = model_id(model); d = model_dim(model); B = model_blob(model)
    th = _f64(theta_pushed).reshape(-1, d); N = th.shape[0]
    dist = np.empty(N); blobs = np.zeros((N, max(B, 1)), dtype=np.uint8)
    lib().orc_simulate(mid, _p(_data(data)), d, C.c_int64(N), _p(th), C.c_uint64(seed), C.c_uint32(epoch),
                       C.c_uint32(tag), C.c_int64(id0), _p(dist), _p(blobs))
    return dist, blobs[:, :B]


def init(prior, model, data, N, seed, theta=None, logpi=None, id0=0):
    d, fam, par = _prior_args(prior)
    mid = model_id(model); B = model_blob(model)
    draw = theta is None
    th = np.empty((N, d)) if draw else _f64(theta).reshape(N, d).copy()
    lp = np.empty(N) if draw else _f64(logpi).copy()
    dl = np.empty(N); bl = np.zeros((N, max(B, 1)), dtype=np.uint8)
    nred = C.c_int64(0)
    rc = lib().orc_init(d, _p(fam), _p(par), mid, _p(_data(data)), C.c_int64(N), C.c_uint64(seed),
                        C.c_int64(id0), int(draw), _p(th), _p(lp), _p(dl), _p(bl), C.byref(nred))
    if rc:
        raise RuntimeError(f"orc_init rc={rc}")
    return th, lp, dl, bl[:, :B], nred.value


def smc_sweep(prior, model, data, theta, logpi, delta, alive, eps, kind, gamma0, gsig=1e-5,
              seed=0, epoch=0, id0=0, a=None, b=None, z=None, u=None, blobs=None, faithful=False):
    d, fam, par = _prior_args(prior)
    mid = model_id(model); B = model_blob(model)
    th = _f64(theta).reshape(-1, d); N = th.shape[0]
    lp = _f64(logpi); dl = _f64(delta)
    al = np.ascontiguousarray(alive, dtype=np.uint8)
    bl = np.zeros((N, max(B, 1)), dtype=np.uint8)
    if blobs is not None and B:
        bl[:, :B] = blobs
    nth = np.empty_like(th); nlp = np.empty(N); ndl = np.empty(N); nbl = np.zeros_like(bl)
    flags = np.zeros(N, dtype=np.uint8)
    ns = C.c_int64(0); na = C.c_int64(0)
    ia = None if a is None else np.ascontiguousarray(a, dtype=np.int32)
    ib = None if b is None else np.ascontiguousarray(b, dtype=np.int32)
    iz = None if z is None else _f64(z)
    iu = None if u is None else _f64(u)
    rc = lib().orc_smc_sweep(d, _p(fam), _p(par), mid, _p(_data(data)), C.c_int64(N), _p(th), _p(lp), _p(dl),
                             _p(bl), _p(al), C.c_double(eps), KERNELS[kind], C.c_double(gamma0),
                             C.c_double(gsig), C.c_uint64(seed), C.c_uint32(epoch), C.c_int64(id0),
                             _p(ia), _p(ib), _p(iz), _p(iu), int(faithful),
                             _p(nth), _p(nlp), _p(ndl), _p(nbl), _p(flags), C.byref(ns), C.byref(na))
    if rc:
        raise RuntimeError(f"orc_smc_sweep rc={rc}")
    return dict(theta=nth, logpi=nlp, delta=ndl, blobs=nbl[:, :B], flags=flags, nsims=ns.value, naccs=na.value)


def mc_sweep(prior, model, data, theta, logpi, delta, eps_pop, eps_target, gamma0, gsig=1e-5,
             seed=0, epoch=0, id0=0, s=None, a=None, b=None, z=None, u=None, blobs=None):
    d, fam, par = _prior_args(prior)
    mid = model_id(model); B = model_blob(model)
    th = _f64(theta).reshape(-1, d); N = th.shape[0]
    lp = _f64(logpi); dl = _f64(delta)
    bl = np.zeros((N, max(B, 1)), dtype=np.uint8)
    if blobs is not None and B:
        bl[:, :B] = blobs
    nth = np.empty_like(th); nlp = np.empty(N); ndl = np.empty(N); nbl = np.zeros_like(bl)
    flags = np.zeros(N, dtype=np.uint8)
    ns = C.c_int64(0)
    i32 = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.int32)
    is_, ia, ib = i32(s), i32(a), i32(b)
    iz = None if z is None else _f64(z)
    iu = None if u is None else _f64(u)
    rc = lib().orc_mc_sweep(d, _p(fam), _p(par), mid, _p(_data(data)), C.c_int64(N), _p(th), _p(lp), _p(dl),
                            _p(bl), C.c_double(eps_pop), C.c_double(eps_target), C.c_double(gamma0),
                            C.c_double(gsig), C.c_uint64(seed), C.c_uint32(epoch), C.c_int64(id0),
                            _p(is_), _p(ia), _p(ib), _p(iz), _p(iu),
                            _p(nth), _p(nlp), _p(ndl), _p(nbl), _p(flags), C.byref(ns))
    if rc:
        raise RuntimeError(f"orc_mc_sweep rc={rc}")
    return dict(theta=nth, logpi=nlp, delta=ndl, blobs=nbl[:, :B], flags=flags, nsims=ns.value)


def quantile_alive(delta, alive, p):
    dl = _f64(delta); al = np.ascontiguousarray(alive, dtype=np.uint8)
    q = C.c_double(); a = C.c_double(); b = C.c_double(); j = C.c_int64()
    rc = lib().orc_quantile_alive(C.c_int64(dl.size), _p(dl), _p(al), C.c_double(p),
                                  C.byref(q), C.byref(a), C.byref(b), C.byref(j))
    if rc:
        raise RuntimeError(f"orc_quantile_alive rc={rc}")
    return q.value, a.value, b.value, j.value


def reweight(delta, W, alive, eps_old, eps_new, kind):
    dl = _f64(delta); Wn = _f64(W).copy(); al = np.ascontiguousarray(alive, dtype=np.uint8).copy()
    wn = C.c_double(); ess = C.c_double(); na = C.c_int64()
    lib().orc_reweight(C.c_int64(dl.size), _p(dl), _p(Wn), _p(al), C.c_double(eps_old), C.c_double(eps_new),
                       KERNELS[kind], C.byref(wn), C.byref(ess), C.byref(na))
    return Wn, al, wn.value, ess.value, na.value


def wsample_stratified(weights, uniforms):
    """1-based indices exactly as the reference's wsample_stratified! returns them."""
    w = _f64(weights); u = _f64(uniforms)
    inds = np.empty(w.size, dtype=np.int64)
    lib().orc_wsample_stratified(C.c_int64(w.size), _p(w), _p(u), _p(inds))
    return inds


def resample_uniforms(N, seed, epoch, id0=0):
    u = np.empty(N)
    lib().orc_resample_uniforms(C.c_int64(N), C.c_uint64(seed), C.c_uint32(epoch), C.c_int64(id0), _p(u))
    return u


@dataclass
class SmcOut:
    P: np.ndarray
    Wns: np.ndarray
    C: np.ndarray
    blobs: np.ndarray
    eps: float
    logZ: float
    iters: int
    nsims: int
    status: int
    hist: dict
    sweep_seconds: float
    nsweeps: int


def smc_run(prior, model, data, eps_target, nparticles=100, alpha=0.95, delta_ess=0.5, nsims_max=10**7,
            Kmcmc=3, Kmcmc_min=1.0, kind="indicator_strict", facc_stop=0.0, facc_min=0.0, facc_tune=0.975,
            seed=1, faithful=False, max_iters=0, hist_cap=4096, islands=1):
    """islands = R > 1: DE partners are drawn inside R contiguous blocks of the population (the sharded
    runs of the CUDA library); everything else (quantile, weights, resampling) stays global."""
    d, fam, par = _prior_args(prior)
    mid = model_id(model); B = model_blob(model); N = nparticles
    o = _SmcOpts(N, alpha, delta_ess, nsims_max, Kmcmc, Kmcmc_min, KERNELS[kind], facc_stop, facc_min,
                 facc_tune, seed, int(faithful), max_iters, int(islands))
    r = _SmcResult()
    P = np.empty((N, d)); W = np.empty(N); Cc = np.empty(N); bl = np.zeros((N, max(B, 1)), dtype=np.uint8)
    h = {k: np.zeros(hist_cap) for k in ("eps", "dmin", "dmax", "logZ", "ess", "facc", "gamma0")}
    hK = np.zeros(hist_cap, dtype=np.int32)
    rc = lib().orc_smc_run(d, _p(fam), _p(par), mid, _p(_data(data)), C.c_double(eps_target), C.byref(o),
                           C.byref(r), _p(P), _p(W), _p(Cc), _p(bl), hist_cap, _p(h["eps"]), _p(h["dmin"]),
                           _p(h["dmax"]), _p(h["logZ"]), _p(h["ess"]), _p(h["facc"]), _p(h["gamma0"]), _p(hK))
    if rc:
        raise RuntimeError(f"orc_smc_run rc={rc}")
    n = r.hist_len
    hist = {k: v[:n] for k, v in h.items()}; hist["Kmcmc"] = hK[:n]
    return SmcOut(P, W, Cc, bl[:, :B], r.eps, r.logZ, r.iters, r.nsims, r.status, hist, r.sweep_seconds, r.nsweeps)


@dataclass
class McOut:
    P: np.ndarray
    C: np.ndarray
    blobs: np.ndarray
    reached_eps: bool
    nsims: int
    sweep_seconds: float


def mc_run(prior, model, data, eps_target, nparticles=50, generations=20, seed=1, islands=1):
    d, fam, par = _prior_args(prior)
    mid = model_id(model); B = model_blob(model); N = nparticles
    o = _McOpts(N, generations, seed, int(islands)); r = _McResult()
    P = np.empty((N, d)); Cc = np.empty(N); bl = np.zeros((N, max(B, 1)), dtype=np.uint8)
    rc = lib().orc_mc_run(d, _p(fam), _p(par), mid, _p(_data(data)), C.c_double(eps_target), C.byref(o),
                          C.byref(r), _p(P), _p(Cc), _p(bl))
    if rc:
        raise RuntimeError(f"orc_mc_run rc={rc}")
    return McOut(P, Cc, bl[:, :B], bool(r.reached_eps), r.nsims, r.sweep_seconds)


def num_threads():
    return lib().orc_num_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))
