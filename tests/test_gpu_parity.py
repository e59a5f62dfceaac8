"""GPU parity tests: every stage of the hot path, called through the C ABI of
libabcdez_cuda.so, against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): with identical injected uniforms/proposals the kernels
reproduce resampling indices and accept decisions bit-exactly; distances / weights / logZ
agree within 1e-6 relative (observed: ~1e-15, CUDA vs glibc libm ulps).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-6          # the north-star tolerance for floating-point quantities
TIGHT = 1e-11        # what we actually observe and guard (libm ulp differences only)

NORMAL10 = [("normal", 0.0, math.sqrt(10))]


def to_prior(A, spec):
    cls = {"normal": A.host.Normal, "uniform": A.host.Uniform, "discrete_uniform": A.host.DiscreteUniform,
           "lognormal": A.host.LogNormal, "exponential": A.host.Exponential, "gamma": A.host.Gamma,
           "beta": A.host.Beta, "negbin": A.host.NegativeBinomial}
    return A.Factored(*[cls[s[0]](*s[1:]) for s in spec])


MODEL_CASES = {
    # name: (prior spec, data)
    "gauss1d": (NORMAL10, [3.0, 1.0]),
    "gauss1d_blob": (NORMAL10, [3.0, 1.0]),
    "gauss_corr10": ([("normal", 0.0, 2.0)] * 10, list(np.linspace(-1, 1, 10)) + [0.5]),
    "dirac": ([("normal", 1.0, 0.2)], [1.5]),
    "normdu": ([("normal", 1.0, 0.5), ("discrete_uniform", 1, 10)], [5.5]),
    "twod": ([("normal", 0.0, 5.0)] * 2, []),
    "twod_inf": ([("normal", 0.0, 5.0)] * 2, []),
    "mixture": ([("uniform", -10.0, 10.0)], [0.0]),
    "wiener": ([("uniform", 0.0, 1.0), ("uniform", 0.0, 4.0)], list(np.sqrt(0.25 * np.arange(31.0) ** 2 + 4.0 * np.arange(31.0)))),
    "lotka_volterra": ([("uniform", 0.0, 2.0)] * 4,
                       [1.0, 0.5, 0.01, 50, 8, 0.05] + list(np.tile([1.2, 0.6], 8))),
    "lotka_volterra_lin": ([("uniform", 0.0, 2.0)] * 4,
                           [1.0, 0.5, 0.01, 50, 8, 0.05] + list(np.tile([1.2, 0.6], 8))),
    "birth_death": ([("uniform", 0.0, 2.0), ("uniform", 0.0, 2.0)], [20.0, 8, 0.5, 5000.0] + [22, 25, 24, 30, 33, 31, 36, 40]),
    "socks": ([("negbin", 4.5, 0.13), ("beta", 15.0, 2.0)], [0.0, 11.0]),
}


# ---------------------------------------------------------------------------------------------
# Philox + priors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("spec,exact", [
    ([("uniform", -1.0, 3.0), ("discrete_uniform", 1, 10)], True),
    ([("normal", 1.0, 2.5), ("lognormal", 0.3, 0.8), ("exponential", 2.0)], False),
    ([("gamma", 2.5, 1.5), ("gamma", 0.6, 2.0), ("beta", 15.0, 2.0), ("negbin", 4.0, 0.1)], False),
])
def test_prior_sample_parity(A, oracle, gpu_ctx, spec, exact):
    """rand(rng, prior) (src/abcdez_priors.jl:53-54): same Philox stream -> same draws (this also
    pins the device Philox4x32-10 against the KAT-checked oracle implementation)."""
    N = 5000
    pr = to_prior(A, spec)
    got = pr.rand(N, seed=0xABCDE2, epoch=3, id0=17)
    want = oracle.prior_sample(spec, N, seed=0xABCDE2, epoch=3, id0=17)
    if exact:
        assert np.array_equal(got, want)
    else:
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)


def test_prior_logpdf_parity_bit_exact_families(A, oracle, gpu_ctx):
    """logpdf(::Factored) (src/abcdez_priors.jl:40-46): Normal/Uniform/DiscreteUniform are bit-exact."""
    spec = [("normal", 1.0, 2.5), ("uniform", -1.0, 3.0), ("discrete_uniform", 1, 10), ("normal", -4.0, 0.3)]
    rng = np.random.default_rng(1)
    th = np.column_stack([rng.normal(1, 3, 4000), rng.uniform(-1.5, 3.5, 4000), rng.uniform(0, 11, 4000), rng.normal(-4, 1, 4000)])
    got = to_prior(A, spec).logpdf(th)
    want = oracle.prior_logpdf(spec, th)
    assert np.array_equal(got, want)
    assert np.isneginf(got).any() and np.isfinite(got).any()
    assert np.array_equal(to_prior(A, spec).push_p(th), oracle.push(spec, th))


def test_prior_logpdf_parity_other_families(A, oracle, gpu_ctx):
    spec = [("lognormal", 0.3, 0.8), ("exponential", 2.0), ("gamma", 2.5, 1.5), ("beta", 15.0, 2.0), ("negbin", 4.0, 0.1)]
    rng = np.random.default_rng(2)
    th = np.column_stack([rng.lognormal(0.3, 0.8, 3000), rng.exponential(2.0, 3000), rng.gamma(2.5, 1.5, 3000),
                          rng.beta(15, 2, 3000), rng.integers(0, 90, 3000)])
    th[::97, 0] = -1.0
    got = to_prior(A, spec).logpdf(th)
    want = oracle.prior_logpdf(spec, th)
    assert np.array_equal(np.isneginf(got), np.isneginf(want))
    f = np.isfinite(want)
    np.testing.assert_allclose(got[f], want[f], rtol=1e-12, atol=1e-12)


def test_factored_reference_testset(A, gpu_ctx):
    """test/runtests.jl:21-36 through the GPU path."""
    d = A.Factored(A.host.Uniform(0, 1), A.host.Uniform(100, 101))
    s = d.rand(50, seed=1)
    assert np.all((s >= [0, 100]) & (s <= [1, 101]))
    assert d.pdf((0.0, 0.0)) == 0.0 and d.pdf((0.5, 100.5)) == 1.0
    assert d.logpdf((0.5, 100.5)) == 0.0 and d.logpdf((0.0, 0.0)) == -math.inf
    assert len(d) == 2
    m = A.Factored(A.host.Uniform(0.0, 1.0), A.host.DiscreteUniform(1, 2))
    smp = m.rand(seed=2)
    assert 0 < smp[0] < 1 and smp[1] in (1.0, 2.0)
    assert math.isclose(m.pdf(smp), 0.5, rel_tol=1e-15) and math.isclose(m.logpdf(smp), math.log(0.5))
    # "Push" testset, test/runtests.jl:38-46
    assert list(A.Factored(A.host.Normal(), A.host.DiscreteUniform()).push_p((2, 1.0))) == [2.0, 1.0]
    assert list(A.Factored(*[A.host.DiscreteUniform(0, 9)] * 4).push_p((0.5, 1.5, 2.5, 3.5))) == [0.0, 2.0, 2.0, 4.0]


# ---------------------------------------------------------------------------------------------
# simulators (dist!)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(MODEL_CASES))
def test_simulate_parity(A, oracle, gpu_ctx, name):
    spec, data = MODEL_CASES[name]
    N = 3000
    th = oracle.push(spec, oracle.prior_sample(spec, N, seed=5))
    if name == "socks":
        th[:, 0] = np.minimum(th[:, 0], 200)
    want, wblob = oracle.simulate(name, data, th, seed=77, epoch=4)
    got, gblob = A.Model(name, data).simulate(th, seed=77, epoch=4)
    assert np.array_equal(np.isfinite(got), np.isfinite(want))
    f = np.isfinite(want)
    np.testing.assert_allclose(got[f], want[f], rtol=1e-9, atol=1e-10)
    if wblob.shape[1]:
        np.testing.assert_allclose(gblob.view(np.float64), wblob.view(np.float64), rtol=1e-9, atol=1e-10)


# ---------------------------------------------------------------------------------------------
# abcde_init!
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gauss1d", "twod_inf", "gauss_corr10", "normdu"])
def test_init_parity(A, oracle, gpu_ctx, name):
    """src/abcdez_init.jl:2-22 incl. the redraw loop (twod_inf returns Inf half the time)."""
    spec, data = MODEL_CASES[name]
    N = 4096 + 37
    wth, wlp, wdl, wbl, wred = oracle.init(spec, name, data, N, seed=99)
    pop = A.Population(to_prior(A, spec), A.Model(name, data), N)
    red = pop.init(seed=99)
    g = pop.download()
    assert red == wred
    if name == "twod_inf":
        assert red > N // 4
    np.testing.assert_allclose(g["theta"], wth, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g["logpi"], wlp, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(g["delta"], wdl, rtol=1e-9, atol=1e-10)
    assert np.all(np.isfinite(g["delta"]))
    pop.close()


def test_init_given_theta(A, oracle, gpu_ctx):
    """Injected prior draws theta0 (the serial draws of src/abcdez_smc.jl:242-243)."""
    spec, data = MODEL_CASES["gauss1d"]
    N = 1000
    th0 = np.random.default_rng(4).normal(0, 3, (N, 1)); lp0 = oracle.prior_logpdf(spec, th0)
    wth, wlp, wdl, _, _ = oracle.init(spec, "gauss1d", data, N, seed=5, theta=th0, logpi=lp0)
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
    pop.upload(theta=th0, logpi=lp0)
    pop.init(seed=5, draw_prior=False)
    g = pop.download()
    assert np.array_equal(g["theta"], wth) and np.array_equal(g["logpi"], wlp)
    np.testing.assert_allclose(g["delta"], wdl, rtol=1e-9, atol=1e-12)
    pop.close()


# ---------------------------------------------------------------------------------------------
# abcdesmc_swarm!
# ---------------------------------------------------------------------------------------------
def _population_state(oracle, spec, name, data, N, seed, dead_frac):
    th, lp, dl, bl, _ = oracle.init(spec, name, data, N, seed=seed)
    rng = np.random.default_rng(seed)
    alive = (rng.random(N) >= dead_frac).astype(np.uint8)
    return th, lp, dl, bl, alive


def _inject_partners(rng, alive):
    N = alive.size
    idx = np.flatnonzero(alive)
    a = np.zeros(N, dtype=np.int32); b = np.zeros(N, dtype=np.int32)
    for i in range(N):
        if not alive[i]:
            continue
        ai = i
        while ai == i:
            ai = rng.choice(idx)
        bi = ai
        while bi == ai or bi == i:
            bi = rng.choice(idx)
        a[i], b[i] = ai, bi
    return a, b


@pytest.mark.parametrize("name,kind,dead", [
    ("gauss1d", "indicator_strict", 0.3), ("gauss1d", "epa", 0.0), ("gauss1d_blob", "indicator", 0.5),
    ("gauss_corr10", "indicator_strict", 0.2), ("normdu", "indicator_strict", 0.1), ("twod_inf", "epa_strict", 0.4),
    ("mixture", "indicator_strict", 0.6), ("wiener", "indicator", 0.0), ("lotka_volterra", "indicator_strict", 0.3),
    ("birth_death", "indicator_strict", 0.3), ("socks", "indicator_strict", 0.2),
])
def test_smc_sweep_injected_parity(A, oracle, gpu_ctx, name, kind, dead):
    """src/abcdez_smc.jl:106-153 with injected (a, b, z, u): accept decisions bit-exact; accepted
    theta/logpi bit-exact (same unfused sub, mul, add as :128); distances within tolerance."""
    spec, data = MODEL_CASES[name]
    N = 3000 + 13
    th, lp, dl, bl, alive = _population_state(oracle, spec, name, data, N, 21, dead)
    rng = np.random.default_rng(7)
    a, b = _inject_partners(rng, alive)
    z = rng.normal(size=N); u = rng.random(N)
    eps = float(np.quantile(dl[alive > 0], 0.8))
    g0 = 2.38 / math.sqrt(2 * len(spec))
    want = oracle.smc_sweep(spec, name, data, th, lp, dl, alive, eps, kind, g0, seed=31, epoch=6, a=a, b=b, z=z, u=u, blobs=bl)
    pop = A.Population(to_prior(A, spec), A.Model(name, data), N)
    pop.upload(theta=th, logpi=lp, delta=dl, blobs=bl, alive=alive)
    pop.set(eps=eps, kernel=kind, gamma0=g0, seed=31, epoch=6)
    r = pop.smc_sweep(a=a, b=b, z=z, u=u)
    g = pop.download()
    assert np.array_equal(r["flags"], want["flags"]), "accept / simulate decisions differ"
    assert (r["nsims"], r["naccs"]) == (want["nsims"], want["naccs"])
    assert want["naccs"] > 0 and want["naccs"] < want["nsims"] <= int(alive.sum())
    assert np.array_equal(g["theta"], want["theta"])
    np.testing.assert_allclose(g["logpi"], want["logpi"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g["delta"], want["delta"], rtol=1e-9, atol=1e-10)
    if bl.shape[1]:
        np.testing.assert_allclose(g["blobs"].view(np.float64), want["blobs"].view(np.float64), rtol=1e-9, atol=1e-10)
    # dead particles keep their state (src/abcdez_smc.jl:114)
    d = alive == 0
    assert np.array_equal(g["theta"][d], th[d]) and np.array_equal(g["delta"][d], dl[d])
    pop.close()


@pytest.mark.parametrize("name,dead", [("gauss1d", 0.0), ("gauss1d", 0.7), ("gauss_corr10", 0.3), ("twod", 0.5)])
def test_smc_sweep_philox_parity(A, oracle, gpu_ctx, name, dead):
    """Same sweep with the Philox contract: the O(1) alive-list partner draw must pick the same
    partners as StatsBase's O(N) scan (oracle run in faithful mode)."""
    spec, data = MODEL_CASES[name]
    N = 2500
    th, lp, dl, bl, alive = _population_state(oracle, spec, name, data, N, 23, dead)
    eps = float(np.quantile(dl[alive > 0], 0.7))
    g0 = 2.38 / math.sqrt(2 * len(spec))
    want = oracle.smc_sweep(spec, name, data, th, lp, dl, alive, eps, "indicator_strict", g0, seed=41, epoch=9, faithful=True)
    pop = A.Population(to_prior(A, spec), A.Model(name, data), N)
    pop.upload(theta=th, logpi=lp, delta=dl, alive=alive)
    pop.set(eps=eps, kernel="indicator_strict", gamma0=g0, seed=41, epoch=9)
    r = pop.smc_sweep()
    g = pop.download()
    assert np.array_equal(r["flags"], want["flags"])
    np.testing.assert_allclose(g["theta"], want["theta"], rtol=1e-12, atol=0)
    np.testing.assert_allclose(g["delta"], want["delta"], rtol=1e-9, atol=1e-10)
    pop.close()


# ---------------------------------------------------------------------------------------------
# abcdemc_swarm!
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["gauss1d", "twod", "normdu"])
def test_mc_sweep_parity(A, oracle, gpu_ctx, name):
    """src/abcdez_mc.jl:5-61, Philox contract and injected (s, a, b, z, u)."""
    spec, data = MODEL_CASES[name]
    N = 2000 + 3
    th, lp, dl, bl, _ = oracle.init(spec, name, data, N, seed=51)
    eps_target = float(np.quantile(dl, 0.3)); eps_pop = max(eps_target, float(dl.min()))
    g0 = 2.38 / math.sqrt(2 * len(spec))
    rng = np.random.default_rng(3)
    inj = {}
    for mode in ("philox", "injected"):
        if mode == "injected":
            s = np.array([rng.choice(np.flatnonzero(dl <= dl[i])) for i in range(N)], dtype=np.int32)
            a = np.empty(N, dtype=np.int32); b = np.empty(N, dtype=np.int32)
            for i in range(N):
                si = s[i] if dl[i] > (eps_target if dl[i] <= eps_target else eps_pop) else i
                ai = si
                while ai == si:
                    ai = rng.integers(N)
                bi = ai
                while bi == ai or bi == si:
                    bi = rng.integers(N)
                a[i], b[i] = ai, bi
            inj = dict(s=s, a=a, b=b, z=rng.normal(size=N), u=rng.random(N))
        want = oracle.mc_sweep(spec, name, data, th, lp, dl, eps_pop, eps_target, g0, seed=61, epoch=2, **inj)
        pop = A.Population(to_prior(A, spec), A.Model(name, data), N)
        pop.upload(theta=th, logpi=lp, delta=dl)
        pop.set(eps=1.0, gamma0=g0, seed=61, epoch=2)
        r = pop.mc_sweep(eps_pop, eps_target, **inj)
        g = pop.download()
        assert np.array_equal(r["flags"], want["flags"]), mode
        assert r["nsims"] == want["nsims"]
        np.testing.assert_allclose(g["theta"], want["theta"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(g["delta"], want["delta"], rtol=1e-9, atol=1e-10)
        pop.close()


@pytest.mark.parametrize("N,ties", [(4096, False), (4097, False), (30011, True), (300007, False)])
def test_mc_sweep_sorted_order_at_scale(A, oracle, gpu_ctx, N, ties):
    """The library's own sort behind src/abcdez_mc.jl:23 (bitonic in shared memory up to 4096 particles, LSD radix
    above; ties in index order): one Philox-driven abcdemc_swarm! sweep equals the oracle's, whose base-particle draw
    walks the qsort-ed (delta, index) order."""
    name = "gauss1d"
    spec, data = MODEL_CASES[name]
    th, lp, dl, bl, _ = oracle.init(spec, name, data, N, seed=52)
    if ties:
        dl = np.round(dl, 1)                                      # heavy ties: the order inside a tie is the index order
    eps_target = float(np.quantile(dl, 0.2)); eps_pop = max(eps_target, float(dl.min()))
    g0 = 2.38 / math.sqrt(2)
    want = oracle.mc_sweep(spec, name, data, th, lp, dl, eps_pop, eps_target, g0, seed=62, epoch=3)
    pop = A.Population(to_prior(A, spec), A.Model(name, data), N)
    pop.upload(theta=th, logpi=lp, delta=dl)
    pop.set(eps=1.0, gamma0=g0, seed=62, epoch=3)
    r = pop.mc_sweep(eps_pop, eps_target)
    g = pop.download()
    assert np.array_equal(r["flags"], want["flags"]) and r["nsims"] == want["nsims"]
    np.testing.assert_array_equal(g["theta"], want["theta"])
    np.testing.assert_allclose(g["delta"], want["delta"], rtol=1e-12, atol=0)
    pop.close()


# ---------------------------------------------------------------------------------------------
# eps schedule: quantile(delta[alive], alpha)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,dead,alpha", [(1, 0.0, 0.95), (2, 0.0, 0.95), (5, 0.0, 0.5), (1000, 0.3, 0.95),
                                          (4099, 0.9, 0.95), (100003, 0.5, 0.95), (100003, 0.0, 0.0), (50000, 0.2, 0.999)])
def test_eps_quantile_bit_exact(A, oracle, gpu_ctx, N, dead, alpha):
    """src/abcdez_smc.jl:301 -- exact select + type-7 interpolation == oracle (sort based), bit for bit."""
    rng = np.random.default_rng(N)
    dl = rng.exponential(size=N) ** 3
    if N > 10:
        dl[rng.choice(N, N // 10, replace=False)] = dl[0]        # ties
        dl[1] = 0.0; dl[2] = 1e300; dl[3] = 5e-324
    alive = (rng.random(N) >= dead).astype(np.uint8)
    alive[0] = 1
    spec, data = MODEL_CASES["gauss1d"]
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
    pop.upload(delta=dl, alive=alive)
    q, lo, hi = pop.eps_quantile(alpha)
    wq, wa, wb, _ = oracle.quantile_alive(dl, alive, alpha)
    assert (q, lo, hi) == (wq, wa, wb)
    pop.close()


def test_eps_quantile_inf_and_nan(A, oracle, gpu_ctx):
    spec, data = MODEL_CASES["gauss1d"]
    N = 2000
    dl = np.random.default_rng(1).exponential(size=N)
    dl[::3] = math.inf
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
    pop.upload(delta=dl, alive=np.ones(N, dtype=np.uint8))
    assert pop.eps_quantile(0.95)[0] == oracle.quantile_alive(dl, np.ones(N, dtype=np.uint8), 0.95)[0] == math.inf
    dl[5] = math.nan
    pop.upload(delta=dl)
    with pytest.raises(A.ABCdeZError) as e:
        pop.eps_quantile(0.95)
    assert e.value.code == A.host.ERR_NAN_DISTANCE
    pop.close()


# ---------------------------------------------------------------------------------------------
# reweighting / ESS / alive
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["indicator", "indicator_strict", "epa", "epa_strict"])
@pytest.mark.parametrize("N", [1000, 65537])
def test_reweight_parity(A, oracle, gpu_ctx, kind, N):
    """src/abcdez_smc.jl:59-83,305-311,323: alive mask bit-exact, weights/wnorm/ESS to 1e-12."""
    rng = np.random.default_rng(N + len(kind))
    dl = rng.exponential(size=N)
    eps_old = float(np.quantile(dl, 0.9)); eps_new = float(np.quantile(dl, 0.8))
    dl[7] = eps_new                     # boundary inclusion differs between strict / non-strict kernels
    alive = np.array([oracle.kernel_pdf(kind, eps_old, x) > 0 for x in dl], dtype=np.uint8)
    W = np.where(alive > 0, rng.random(N) if kind.startswith("epa") else 1.0, 0.0); W /= W.sum()
    wW, wal, wnorm, wess, wna = oracle.reweight(dl, W, alive, eps_old, eps_new, kind)
    spec, data = MODEL_CASES["gauss1d"]
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
    pop.upload(delta=dl, W=W, alive=alive)
    pop.set(eps=eps_old, eps_prev=eps_old, kernel=kind)
    gn, gess, gna = pop.reweight(eps_new)
    g = pop.download()
    assert np.array_equal(g["alive"], wal) and gna == wna
    assert bool(g["alive"][7]) == (kind == "indicator")
    np.testing.assert_allclose(g["W"], wW, rtol=1e-12, atol=0)
    assert math.isclose(gn, wnorm, rel_tol=1e-12) and math.isclose(gess, wess, rel_tol=1e-12)
    if not kind.startswith("epa"):
        assert np.unique(g["W"][g["alive"] > 0]).size == 1
    pop.close()


# ---------------------------------------------------------------------------------------------
# fused iteration head (head.cu): quantile + clamp + reweight + ESS + alive list in one launch
# ---------------------------------------------------------------------------------------------
def _head_case(rng, N, dead, ties):
    dl = rng.exponential(size=N) ** 2
    if ties == "discrete":
        dl = np.floor(dl * 3.0)                                  # few distinct values: huge ties (socks-like)
    elif ties == "some" and N > 10:
        dl[rng.choice(N, N // 10, replace=False)] = dl[0]
        dl[1] = 0.0; dl[2] = 1e300; dl[3] = 5e-324
    elif ties == "narrow":
        dl = 1.0 + dl * 1e-9                                     # all keys share > 22 leading bits: long candidate list
    alive = (rng.random(N) >= dead).astype(np.uint8)
    alive[0] = 1
    return dl, alive


@pytest.mark.parametrize("kind", ["indicator_strict", "epa"])
@pytest.mark.parametrize("N,dead,alpha,ties", [
    (1, 0.0, 0.95, "none"), (2, 0.0, 0.95, "none"), (7, 0.3, 0.5, "none"), (1000, 0.3, 0.95, "some"),
    (4099, 0.9, 0.95, "some"), (100003, 0.5, 0.95, "some"), (100003, 0.0, 0.0, "none"), (50000, 0.2, 0.999, "discrete"),
    (300000, 0.1, 0.95, "narrow"), (1 << 20, 0.4, 0.95, "none"), (3000001, 0.0, 0.9, "narrow")])
def test_fused_head_equals_stage_calls_and_oracle(A, oracle, gpu_ctx, kind, N, dead, alpha, ties):
    """head_kernel == eps_quantile + clamp + reweight + compact, bit for bit, and == the oracle."""
    rng = np.random.default_rng(N + int(alpha * 1000))
    dl, alive = _head_case(rng, N, dead, ties)
    eps_prev = float(np.max(dl[alive > 0])) * 2 + 1.0
    Wv = np.where(alive > 0, rng.random(N) if kind == "epa" else 1.0, 0.0); Wv /= Wv.sum()
    spec, data = MODEL_CASES["gauss1d"]
    res = {}
    for mode in ("fused", "stage"):
        pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
        pop.upload(delta=dl, W=Wv, alive=alive)
        pop.set(eps=eps_prev, eps_prev=eps_prev, kernel=kind)
        if mode == "fused":
            q, eps, wn, ess, na = pop.head(alpha, 0.0)
        else:
            q = pop.eps_quantile(alpha)[0]
            eps = max(min(q, eps_prev), 0.0)
            wn, ess, na = pop.reweight(eps)
        st = pop.download()
        res[mode] = (q, eps, wn, ess, na, st["W"].copy(), st["alive"].copy())
        pop.close()
    f, s_ = res["fused"], res["stage"]
    assert np.array_equal(np.array(f[:5], float), np.array(s_[:5], float), equal_nan=True)   # wnorm = 0 -> ess NaN in both
    assert np.array_equal(f[5], s_[5], equal_nan=True) and np.array_equal(f[6], s_[6])
    wq = oracle.quantile_alive(dl, alive, alpha)[0]
    assert f[0] == wq
    wW, wal, wnorm, wess, wna = oracle.reweight(dl, Wv, alive, eps_prev, f[1], kind)
    assert np.array_equal(f[6], wal) and f[4] == wna
    np.testing.assert_allclose(f[5], wW, rtol=1e-12, atol=0, equal_nan=True)
    assert math.isclose(f[2], wnorm, rel_tol=1e-12)
    assert math.isclose(f[3], wess, rel_tol=1e-12) or (math.isnan(f[3]) and math.isnan(wess)) or (wnorm == 0.0)


@pytest.mark.parametrize("kind", ["indicator_strict", "epa"])
@pytest.mark.parametrize("N,alpha", [(50000, 0.95), (1 << 20, 0.95), (1300001, 0.9), (20000, 0.3)])
def test_fused_head_sequence_adapts_its_window(A, oracle, gpu_ctx, kind, N, alpha):
    """Successive heads on one population (each starts from the previous eps and the window the previous head
    derived from the eps gap): every call equals the stage kernels bit for bit and the oracle's quantile."""
    rng = np.random.default_rng(N + 7)
    dl = np.abs(rng.normal(size=N)) ** 1.5 + 0.01
    alive = np.ones(N, dtype=np.uint8)
    Wv = np.where(alive > 0, rng.random(N) if kind == "epa" else 1.0, 0.0); Wv /= Wv.sum()
    spec, data = MODEL_CASES["gauss1d"]
    pops = {}
    for mode in ("fused", "stage"):
        pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
        pop.upload(delta=dl, W=Wv, alive=alive)
        pop.set(eps=math.inf, eps_prev=math.inf, kernel=kind)
        pops[mode] = pop
    eps_prev = math.inf
    for it in range(6):
        q, eps, wn, ess, na = pops["fused"].head(alpha, 0.0)
        sq = pops["stage"].eps_quantile(alpha)[0]
        seps = max(min(sq, eps_prev), 0.0)
        pops["stage"].set(eps=eps_prev, eps_prev=eps_prev, kernel=kind)
        swn, sess, sna = pops["stage"].reweight(seps)
        f, s_ = pops["fused"].download(), pops["stage"].download()
        wq = oracle.quantile_alive(dl, alive if it == 0 else prev_alive, alpha)[0]
        assert q == sq == wq, (it, q, sq, wq)
        assert (eps, wn, ess, na) == (seps, swn, sess, sna), it
        assert np.array_equal(f["W"], s_["W"]) and np.array_equal(f["alive"], s_["alive"])
        prev_alive = s_["alive"].copy()
        eps_prev = eps
        assert 0 < na < N
    for pop in pops.values():
        pop.close()


def test_fused_head_alive_list_drives_partner_draws(A, oracle, gpu_ctx):
    """After the fused head, a Philox sweep (partners through the compacted alive list) equals the oracle's."""
    name = "gauss_corr10"
    spec, data = MODEL_CASES[name]
    N = 20000
    th, lp, dl, _, alive = _population_state(oracle, spec, name, data, N, seed=5, dead_frac=0.0)
    pop = A.Population(to_prior(A, spec), A.Model(name, data), N)
    pop.upload(theta=th, logpi=lp, delta=dl, alive=alive)
    pop.set(eps=math.inf, eps_prev=math.inf, kernel="indicator_strict", seed=99, epoch=3)
    q, eps, wn, ess, na = pop.head(0.6, 0.0)
    st = pop.download()
    assert na == int(st["alive"].sum()) and 0.55 * N < na < 0.65 * N
    pop.set(eps=eps, eps_prev=eps, kernel="indicator_strict", seed=99, epoch=3)
    got = pop.smc_sweep()          # NB: pop.set leaves n_alive / alive_list of the head untouched
    want = oracle.smc_sweep(spec, name, data, th, lp, dl, st["alive"], eps, "indicator_strict",
                            2.38 / math.sqrt(20), seed=99, epoch=3)
    assert np.array_equal(got["flags"], want["flags"])
    g = pop.download()
    np.testing.assert_array_equal(g["theta"], want["theta"])
    pop.close()


def test_fused_head_nan_and_empty(A, gpu_ctx):
    spec, data = MODEL_CASES["gauss1d"]
    N = 5000
    dl = np.random.default_rng(2).exponential(size=N)
    dl[17] = math.nan
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
    pop.upload(delta=dl, alive=np.ones(N, dtype=np.uint8))
    pop.set(eps=math.inf, eps_prev=math.inf)
    with pytest.raises(A.ABCdeZError) as e:
        pop.head(0.95)
    assert e.value.code == A.host.ERR_NAN_DISTANCE
    pop.close()


# ---------------------------------------------------------------------------------------------
# stratified resampling
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("N,alive_frac", [(3, 1.0), (100, 0.5), (1000, 0.49), (4097, 0.3), (65536, 0.5), (100003, 0.37),
                                          (1000003, 0.45)])
def test_resample_uniform_weights_bit_exact(A, oracle, gpu_ctx, N, alive_frac):
    """wsample_stratified! (src/abcdez_smc.jl:15-56) for indicator-kernel weights: the closed form
    of the sequential FP64 running sums reproduces the reference's indices bit for bit, at any N."""
    rng = np.random.default_rng(N)
    alive = (rng.random(N) < alive_frac).astype(np.uint8); alive[rng.integers(N)] = 1
    na = int(alive.sum())
    c = (1.0 / N) / (na / N)                      # Wns after normalisation: (1/N * 1) / wnorm
    W = np.where(alive > 0, c, 0.0)
    u = rng.random(N)
    u[0] = 0.999999; u[-1] = 0.9999     # (u == 0 in stratum 1 / r above the total are the reference's
                                        # out-of-bounds corners, see DESIGN.md 'Quirks')
    want = np.clip(oracle.wsample_stratified(W, u), 1, N) - 1
    spec, data = MODEL_CASES["gauss1d"]
    th = rng.normal(size=(N, 1)); lp = rng.normal(size=N); dl = rng.exponential(size=N)
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
    pop.upload(theta=th, logpi=lp, delta=dl, W=W, alive=alive)
    pop.set(eps=1.0, kernel="indicator_strict")
    inds = pop.resample(uniforms=u, mode=0)
    assert np.array_equal(inds, want.astype(np.int32))
    g = pop.download()
    assert np.array_equal(g["theta"], th[want]) and np.array_equal(g["logpi"], lp[want]) and np.array_equal(g["delta"], dl[want])
    assert np.all(g["alive"] == 1) and np.all(g["W"] == 1.0 / N)       # :102-103
    pop.close()


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("N", [1000, 40000])
def test_resample_general_weights(A, oracle, gpu_ctx, N, mode):
    """Continuous (Epanechnikov) weights.  mode 2 = sequential cumsum, bit-exact; mode 1 = parallel scan,
    whose rounding differs from the sequential sum: any mismatch must be a boundary tie (adjacent
    alive particle) and rare."""
    rng = np.random.default_rng(N + mode)
    W = rng.random(N) * (rng.random(N) < 0.6); W /= W.sum()
    alive = (W > 0).astype(np.uint8)
    u = rng.random(N)
    want = np.clip(oracle.wsample_stratified(W, u), 1, N) - 1
    spec, data = MODEL_CASES["gauss1d_blob"]
    th = rng.normal(size=(N, 1)); dl = rng.exponential(size=N); bl = rng.normal(size=N).view(np.uint8).reshape(N, 8)
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d_blob", data), N)
    pop.upload(theta=th, logpi=dl, delta=dl, blobs=bl, W=W, alive=alive)
    pop.set(eps=1.0, kernel="epa")
    inds = pop.resample(uniforms=u, mode=mode)
    if mode == 2:
        assert np.array_equal(inds, want)
    else:
        bad = np.flatnonzero(inds != want)
        assert bad.size <= max(1, N // 10000)
        rank = np.cumsum(alive) - 1
        assert np.all(np.abs(rank[inds[bad]] - rank[want[bad]]) <= 1)
    g = pop.download()
    assert np.array_equal(g["theta"], th[inds]) and np.array_equal(g["blobs"], bl[inds])
    pop.close()


def test_weightinds_entry_point(A, oracle, gpu_ctx):
    """`weightinds` of test/runtests.jl:13-19 (ABCdeZ.wsample_stratified! on user weights)."""
    rng = np.random.default_rng(12)
    N = 30000
    k = np.maximum(0.0, 1.0 - rng.normal(0, 2, N) ** 2); w = k / k.sum()
    u = rng.random(N)
    want = oracle.wsample_stratified(w, u)
    assert np.array_equal(A.wsample_stratified(w, u, mode=2), want)
    got = A.wsample_stratified(w, u, mode=1)
    assert np.mean(got != want) < 1e-3


def test_resample_philox_uniforms(A, oracle, gpu_ctx):
    """Resampling uniforms from the Philox contract (TAG_RESAMPLE stream)."""
    N = 5000
    rng = np.random.default_rng(2)
    alive = (rng.random(N) < 0.4).astype(np.uint8)
    W = np.where(alive > 0, 1.0 / alive.sum(), 0.0)
    u = oracle.resample_uniforms(N, seed=123, epoch=14)
    want = np.clip(oracle.wsample_stratified(W, u), 1, N) - 1
    spec, data = MODEL_CASES["gauss1d"]
    pop = A.Population(to_prior(A, spec), A.Model("gauss1d", data), N)
    pop.upload(W=W, alive=alive)
    pop.set(eps=1.0, kernel="indicator_strict", seed=123)
    assert np.array_equal(pop.resample(uniforms=None, epoch=14, mode=0), want)
    pop.close()
