"""Small-N pass over every kernel family of the hot path, meant to run under
    compute-sanitizer --tool memcheck | racecheck | synccheck python scripts/sanitize_cases.py
(scripts/gpu_sanitize.sh).  Sizes are small because the tools slow kernels down 10-100x; every case still checks its
result against the CPU oracle, so a hazard that changes a value is caught twice."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abcdez_b200 as A
from oracle import oracle as O

O.build(); O.lib()
ctx = A.default_context()
done = []


ONLY = os.environ.get("SANITIZE_ONLY", "")          # substring of a case name: run just those


def case(name):
    def deco(f):
        if ONLY and ONLY not in name:
            return f
        f(); done.append(name); print("ok:", name, flush=True)
        return f
    return deco


G1 = ([("normal", 0.0, math.sqrt(10.0))], "gauss1d", [3.0, 1.0])
C10 = ([("normal", 0.0, 2.0)] * 10, "gauss_corr10", list(np.linspace(-1, 1, 10)) + [0.5])


def prior_of(spec):
    cls = {"normal": A.host.Normal, "uniform": A.host.Uniform}
    return A.Factored(*[cls[s[0]](*s[1:]) for s in spec])


@case("init + injected / Philox smc sweep (lazy ping-pong, tickets)")
def _():
    spec, name, data = C10
    N = 3000
    th, lp, dl, _, _ = O.init(spec, name, data, N, seed=5)
    alive = (np.random.default_rng(1).random(N) > 0.3).astype(np.uint8)
    pop = A.Population(prior_of(spec), A.Model(name, data), N)
    assert pop.init(seed=5) == 0
    pop.upload(alive=alive, W=np.where(alive > 0, 1.0 / alive.sum(), 0.0))
    eps = float(np.quantile(dl, 0.7))
    for epoch in (1, 2, 3):
        pop.set(eps=eps, kernel="indicator_strict", seed=9, epoch=epoch)
        got = pop.smc_sweep()
        want = O.smc_sweep(spec, name, data, th, lp, dl, alive, eps, "indicator_strict", 2.38 / math.sqrt(20), seed=9, epoch=epoch)
        assert np.array_equal(got["flags"], want["flags"])
        th, lp, dl = want["theta"], want["logpi"], want["delta"]
    assert np.array_equal(pop.download()["theta"], th)
    pop.close()


@case("fused head x4 (window, generic passes, candidates, reweight, alive list) + stage kernels")
def _():
    spec, name, data = G1
    for N in (5000, 40000):
        rng = np.random.default_rng(N)
        dl = np.abs(rng.normal(size=N)) ** 1.5 + 0.01
        pf, ps = A.Population(prior_of(spec), A.Model(name, data), N), A.Population(prior_of(spec), A.Model(name, data), N)
        for p in (pf, ps):
            p.upload(delta=dl); p.set(eps=math.inf, eps_prev=math.inf)
        eps_prev = math.inf
        for it in range(4):
            q, eps, wn, ess, na = pf.head(0.9, 0.0)
            sq = ps.eps_quantile(0.9)[0]
            ps.set(eps=eps_prev, eps_prev=eps_prev)
            swn, sess, sna = ps.reweight(max(min(sq, eps_prev), 0.0))
            assert (q, wn, ess, na) == (sq, swn, sess, sna), (it, q, sq)
            eps_prev = eps
        assert np.array_equal(pf.download()["alive"], ps.download()["alive"])
        pf.close(); ps.close()


@case("stratified resampling: closed form (indicator) and scans (Epanechnikov)")
def _():
    spec, name, data = G1
    N = 6000
    rng = np.random.default_rng(3)
    alive = (rng.random(N) > 0.5).astype(np.uint8)
    u = rng.random(N)
    for kind, mode in (("indicator_strict", 0), ("epa", 1), ("epa", 2)):
        W = np.where(alive > 0, 1.0 if mode == 0 else rng.random(N), 0.0); W /= W.sum()
        pop = A.Population(prior_of(spec), A.Model(name, data), N)
        pop.init(seed=2)
        pop.upload(W=W, alive=alive)
        pop.set(eps=1.0, kernel=kind)
        inds = pop.resample(uniforms=u, mode=mode)
        want = np.clip(O.wsample_stratified(W, u), 1, N) - 1
        assert (inds == want).mean() > 0.999
        pop.close()


@case("whole abcdesmc! runs (device-side control, resampling) incl. blobs and Epanechnikov")
def _():
    for (spec, name, data), eps, kw in ((G1, 0.3, {}), (C10, 3.0, {}), (([("uniform", 0.0, 2.0)] * 2, "birth_death", [20.0, 8, 0.5, 5000.0, 22, 25, 24, 30, 33, 31, 36, 40]), 4.0, {}),
                                        (G1, 0.3, dict(ABCk="epa", exact_scan=True))):
        okw = dict(kind=kw.get("ABCk", "indicator_strict"))
        want = O.smc_run(spec, name, data, eps, nparticles=1500, seed=4, nsims_max=10**8, **okw)
        got = A.abcdesmc(prior_of(spec), A.Model(name, data), eps, None, nparticles=1500, rng=4, nsims_max=10**8, verbose=False, **kw)
        assert (got.iters, got.nsims) == (want.iters, want.nsims), name
        assert abs(got.logZ - want.logZ) <= 1e-9 * abs(want.logZ)


@case("abcdemc! (bitonic and radix sort paths, device-side generation loop)")
def _():
    spec, name, data = G1
    for N, gens in ((600, 15), (6000, 6)):
        want = O.mc_run(spec, name, data, 0.3, nparticles=N, generations=gens, seed=6)
        got = A.abcdemc(prior_of(spec), A.Model(name, data), 0.3, None, nparticles=N, generations=gens, rng=6, verbose=False)
        assert got.nsims == want.nsims and np.allclose(got.C, want.C, rtol=1e-12)


@case("queue-driven sweeps of heavy simulators: init / abcdesmc! / abcdemc! of birth-death (stepped) and Lotka-Volterra")
def _():
    bd = ([("uniform", 0.0, 2.0)] * 2, "birth_death", [20.0, 8, 0.5, 5000.0, 22, 25, 24, 30, 33, 31, 36, 40])
    lv = ([("uniform", 0.0, 2.0)] * 4, "lotka_volterra", [1.0, 0.5, 0.01, 50, 8, 0.05] + list(np.tile([1.2, 0.6], 8)))
    for (spec, name, data), eps in ((bd, 5.0), (lv, 0.6)):
        want = O.smc_run(spec, name, data, eps, nparticles=700, seed=9, nsims_max=20000)
        got = A.abcdesmc(prior_of(spec), A.Model(name, data), eps, None, nparticles=700, rng=9, nsims_max=20000, verbose=False)
        assert (got.iters, got.nsims) == (want.iters, want.nsims), name
        wm = O.mc_run(spec, name, data, eps, nparticles=400, generations=5, seed=10)
        gm = A.abcdemc(prior_of(spec), A.Model(name, data), eps, None, nparticles=400, generations=5, rng=10, verbose=False)
        assert gm.nsims == wm.nsims and np.array_equal(gm.C, wm.C), name


@case("relaxed-parity modes: FP32 particle state, partner segments, systematic resampling (fused sweep, resampling, result assembly)")
def _():
    spec, name, data = C10
    r = A.abcdesmc(prior_of(spec), A.Model(name, data), 3.0, None, nparticles=1500, rng=4, nsims_max=10**8, verbose=False,
                   fp32_state=True, partner_segments=True, systematic_resampling=True)
    assert r.iters > 5 and np.isfinite(r.logZ) and np.array_equal(r.P, r.P.astype(np.float32).astype(np.float64))
    r1 = A.abcdesmc(prior_of(G1[0]), A.Model(G1[1], G1[2]), 0.3, None, nparticles=1500, rng=4, verbose=False, fp32_state=True)
    assert abs(math.exp(r1.logZ) / 0.047940112540007955 - 1.0) < 0.3


@case("g-and-k CTA-cooperative simulator (FP64 and FP32), multi-select")
def _():
    data = [1000.0, 2.39384, 2.569082, 2.748052, 3.0, 3.4169, 4.196232, 5.900654]
    th = O.prior_sample([("uniform", 0.0, 10.0)] * 4, 64, seed=3)
    for m in ("gk", "gk_f32"):
        assert np.array_equal(A.Model(m, data).simulate(th, seed=1)[0], O.simulate(m, data, th, seed=1)[0])
    data4k = [4096.0] + data[1:]                     # n >= 4096: the z-space fast path (rows 0, 1: its fallback)
    th4k = th[:24].copy(); th4k[0] = [3.0, -1.0, 2.0, 0.5]; th4k[1] = [3.0, 1.0, 2.0, -0.3]
    for m in ("gk", "gk_f32"):
        assert np.array_equal(A.Model(m, data4k).simulate(th4k, seed=1)[0], O.simulate(m, data4k, th4k, seed=1)[0])
    spec = [("uniform", 0.0, 10.0)] * 3 + [("uniform", 0.0, 2.0)]
    got = A.abcdesmc(prior_of(spec), A.Model("gk", data), 1.0, None, nparticles=300, rng=2, nsims_max=3000, verbose=False)
    want = O.smc_run(spec, "gk", data, 1.0, nparticles=300, seed=2, nsims_max=3000)
    assert (got.iters, got.nsims) == (want.iters, want.nsims)


print(f"sanitize_cases: {len(done)} cases passed")
