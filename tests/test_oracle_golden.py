"""Pins the CPU oracle against every known answer the reference's own tests hold for the hot
path (test/runtests.jl of ABCdeZ.jl v0.6.0), and cross-checks the restated third-party
semantics (Statistics.quantile type 7, StatsBase.wsample, Distributions logpdfs) against
numpy / scipy.  CPU only."""
import math

import numpy as np
import pytest
import scipy.stats as st

INF = math.inf


# ---- Philox4x32-10: Random123 known-answer vectors ---------------------------
def test_philox_kat(oracle):
    kat = [(([0, 0, 0, 0], [0, 0]), [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
           (([0xffffffff] * 4, [0xffffffff] * 2), [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
           (([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]),
            [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for (ctr, key), want in kat:
        assert list(oracle.philox(ctr, key)) == want


# ---- "Factored" testset, test/runtests.jl:21-36 ---------------------------------------------
def test_factored_known_answers(oracle):
    d = [("uniform", 0.0, 1.0), ("uniform", 100.0, 101.0)]
    s = oracle.prior_sample(d, 64, seed=3)
    assert np.all((s >= [0, 100]) & (s <= [1, 101]))                      # :23
    assert math.exp(oracle.prior_logpdf(d, [0.0, 0.0], raw=True)[0]) == 0.0   # :24
    assert math.exp(oracle.prior_logpdf(d, [0.5, 100.5], raw=True)[0]) == 1.0  # :25
    assert oracle.prior_logpdf(d, [0.5, 100.5], raw=True)[0] == 0.0       # :26
    assert oracle.prior_logpdf(d, [0.0, 0.0], raw=True)[0] == -INF        # :27
    m = [("uniform", 0.0, 1.0), ("discrete_uniform", 1, 2)]
    smp = oracle.prior_sample(m, 200, seed=5)
    assert np.all((smp[:, 0] > 0) & (smp[:, 0] < 1))                      # :32
    assert set(np.unique(smp[:, 1])) <= {1.0, 2.0}                        # :33
    lp = oracle.prior_logpdf(m, smp, raw=True)
    assert np.all(np.exp(lp) == 0.5)                                      # :34
    assert np.allclose(lp, math.log(0.5), rtol=0, atol=1e-15)             # :35


# ---- "Push" testset, test/runtests.jl:38-46 -------------------------------------------------
def test_push_known_answers(oracle):
    assert oracle.push([("normal", 0, 1)], [1])[0, 0] == 1.0                           # :42
    assert oracle.push([("discrete_uniform", 0, 1)], [1.0])[0, 0] == 1                 # :43
    assert list(oracle.push([("normal", 0, 1), ("discrete_uniform", 0, 1)], [2, 1.0])[0]) == [2.0, 1.0]   # :44
    # round(Int, x) is ties-to-even
    assert list(oracle.push([("discrete_uniform", 0, 9)] * 4, [0.5, 1.5, 2.5, -0.5])[0]) == [0.0, 2.0, 2.0, -0.0]


# ---- the four kernel testsets, test/runtests.jl:48-108 --------------------------------------
KERNEL_CASES = {
    "indicator": {0.1: [(0.1, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 0.0, -INF)],                       # :51-53
                  INF: [(0.1, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 1.0, 0.0)]},                      # :57-59
    "indicator_strict": {0.1: [(0.1, 0.0, -INF), (0.01, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 0.0, -INF)],   # :65-68
                         INF: [(0.1, 1.0, 0.0), (0.01, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 1.0, 0.0)]},  # :72-75
    "epa": {0.1: [(0.1, 0.0, -INF), (0.0, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 0.0, -INF)],          # :81-84
            INF: [(0.1, 1.0, 0.0), (0.0, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 1.0, 0.0)]},           # :88-91
    "epa_strict": {0.1: [(0.1, 0.0, -INF), (0.0, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 0.0, -INF)],   # :97-100
                   INF: [(0.1, 1.0, 0.0), (0.0, 1.0, 0.0), (-0.1, 0.0, -INF), (0.2, 1.0, 0.0)]},    # :104-107
}


@pytest.mark.parametrize("kind", list(KERNEL_CASES))
def test_kernel_known_answers(oracle, kind):
    for eps, cases in KERNEL_CASES[kind].items():
        for x, pdf, logpdf in cases:
            assert oracle.kernel_pdf(kind, eps, x) == pdf, (kind, eps, x)
            assert oracle.kernel_logpdf(kind, eps, x) == logpdf, (kind, eps, x)


def test_epa_interior_value(oracle):
    assert oracle.kernel_pdf("epa", 0.3, 0.15) == 1.0 - (0.15 / 0.3) ** 2
    assert oracle.kernel_logpdf("epa_strict", 0.3, 0.15) == math.log(1.0 - (0.15 / 0.3) ** 2)


# ---- marginal logpdfs vs scipy (Distributions.jl semantics restated) ------------------------
def test_marginal_logpdfs_vs_scipy(oracle):
    x = np.linspace(-3, 12, 61)
    cases = [
        (("normal", 1.0, 2.5), st.norm(1.0, 2.5).logpdf(x)),
        (("uniform", -1.0, 3.0), st.uniform(-1.0, 4.0).logpdf(x)),
        (("lognormal", 0.3, 0.8), st.lognorm(0.8, scale=math.exp(0.3)).logpdf(x)),
        (("exponential", 2.0), st.expon(scale=2.0).logpdf(x)),
        (("gamma", 2.5, 1.5), st.gamma(2.5, scale=1.5).logpdf(x)),
        (("beta", 15.0, 2.0), st.beta(15.0, 2.0).logpdf(x / 12.0)),
    ]
    for fam, want in cases:
        xx = x / 12.0 if fam[0] == "beta" else x
        got = oracle.prior_logpdf([fam], xx.reshape(-1, 1), raw=True)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12, err_msg=str(fam))
    k = np.arange(0, 60, dtype=float)
    got = oracle.prior_logpdf([("negbin", 4.0, 0.1)], k.reshape(-1, 1), raw=True)
    np.testing.assert_allclose(got, st.nbinom(4.0, 0.1).logpmf(k), rtol=1e-12)
    got = oracle.prior_logpdf([("discrete_uniform", 1, 10)], k.reshape(-1, 1), raw=True)
    np.testing.assert_allclose(got, st.randint(1, 11).logpmf(k), rtol=1e-12)


def test_prior_samplers_distribution(oracle):
    """Samplers are not reference arithmetic; check them distributionally (KS / moments)."""
    N = 20000
    for fam, dist in [(("normal", 1.0, 2.0), st.norm(1.0, 2.0)), (("uniform", -2.0, 5.0), st.uniform(-2.0, 7.0)),
                      (("lognormal", 0.2, 0.5), st.lognorm(0.5, scale=math.exp(0.2))),
                      (("exponential", 3.0), st.expon(scale=3.0)), (("gamma", 2.5, 1.5), st.gamma(2.5, scale=1.5)),
                      (("gamma", 0.6, 2.0), st.gamma(0.6, scale=2.0)), (("beta", 15.0, 2.0), st.beta(15.0, 2.0))]:
        s = oracle.prior_sample([fam], N, seed=11)[:, 0]
        assert st.kstest(s, dist.cdf).pvalue > 1e-3, fam
    s = oracle.prior_sample([("negbin", 4.0, 0.1)], N, seed=12)[:, 0]
    assert abs(s.mean() - st.nbinom(4.0, 0.1).mean()) < 4 * st.nbinom(4.0, 0.1).std() / math.sqrt(N)
    s = oracle.prior_sample([("discrete_uniform", 1, 10)], N, seed=13)[:, 0]
    assert set(np.unique(s)) == set(float(v) for v in range(1, 11))


# ---- Statistics.quantile type 7 == numpy 'linear' -------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 10, 257, 5000])
@pytest.mark.parametrize("p", [0.0, 0.5, 0.95, 0.999])
def test_quantile_type7_vs_numpy(oracle, n, p):
    rng = np.random.default_rng(n * 1000 + int(p * 1000))
    d = rng.exponential(size=n + 7)
    alive = np.ones(n + 7, dtype=np.uint8)
    alive[rng.choice(n + 7, 7, replace=False)] = 0
    q, a, b, j = oracle.quantile_alive(d, alive, p)
    want = np.quantile(d[alive > 0], p, method="linear")
    assert a <= q <= b
    np.testing.assert_allclose(q, want, rtol=1e-14)


def test_quantile_with_inf(oracle):
    d = np.array([0.1, 0.2, INF, INF]); al = np.ones(4, dtype=np.uint8)
    q, a, b, j = oracle.quantile_alive(d, al, 0.95)
    assert q == INF


# ---- StatsBase.wsample: O(N) scan == O(1) alive-list form, for the same uniform ---------------
def test_wsample_faithful_equals_list(oracle):
    rng = np.random.default_rng(5)
    N = 300
    prior = [("normal", 0.0, 1.0)]
    th = rng.normal(size=(N, 1)); lp = oracle.prior_logpdf(prior, th); dl = rng.exponential(size=N)
    alive = (rng.random(N) < 0.6).astype(np.uint8)
    r1 = oracle.smc_sweep(prior, "gauss1d", [0.0, 1.0], th, lp, dl, alive, 1.0, "indicator_strict", 1.68, seed=9, epoch=2, faithful=True)
    r2 = oracle.smc_sweep(prior, "gauss1d", [0.0, 1.0], th, lp, dl, alive, 1.0, "indicator_strict", 1.68, seed=9, epoch=2, faithful=False)
    for k in ("theta", "logpi", "delta", "flags"):
        assert np.array_equal(r1[k], r2[k]), k


# ---- wsample_stratified!: indices are a valid stratified draw ---------------------------------
def test_stratified_properties(oracle):
    rng = np.random.default_rng(1)
    N = 4000
    w = rng.random(N) * (rng.random(N) < 0.5); w /= w.sum()
    u = rng.random(N)
    inds = oracle.wsample_stratified(w, u)
    assert np.all(np.diff(inds) >= 0) and inds.min() >= 1 and inds.max() <= N
    assert np.all(w[inds - 1] > 0)
    # closed form: first index whose cumulative weight reaches r
    cs = np.cumsum(w)
    edges = np.concatenate([[0.0], np.cumsum(np.full(N, 1.0 / N))])
    r = edges[:-1] + (edges[1:] - edges[:-1]) * u
    assert np.array_equal(inds, np.searchsorted(cs, r, side="left") + 1)
    # expected copies ~ N * w
    counts = np.bincount(inds - 1, minlength=N)
    assert np.all(np.abs(counts - N * w) < 2.0)      # stratified (one uniform per stratum), not systematic


# ---- analytic evidences / posterior means, test/runtests.jl:110-423 ---------------------------
def _isaround(x, val, f=1.0):
    return x.mean() - f * x.std(ddof=1) <= val <= x.mean() + f * x.std(ddof=1)


def test_analytic_constants():
    assert math.isclose(st.norm(0, math.sqrt(11)).pdf(3) * 0.6, 0.047940112540007955, rel_tol=1e-12)    # :121
    assert math.isclose(st.norm(0, math.sqrt(11)).pdf(7) * 0.6, 0.007781668367620676, rel_tol=1e-12)    # :176
    assert math.isclose(st.norm(0, math.sqrt(11)).pdf(3) * (4 / 3) * 0.3, 0.03196007502667197, rel_tol=1e-12)   # :336


@pytest.mark.parametrize("xdata,Z,tol", [(3.0, 0.047940112540007955, 0.10), (7.0, 0.007781668367620676, 0.20)])
def test_smc_evidence_1d_normal(oracle, xdata, Z, tol):
    """test/runtests.jl:110-163 and :165-218 (abcdesmc! part)."""
    r = oracle.smc_run([("normal", 0.0, math.sqrt(10))], "gauss1d", [xdata, 1.0], 0.3, nparticles=5000, seed=101)
    assert Z * (1 - tol) <= math.exp(r.logZ) <= Z * (1 + tol)
    assert _isaround(r.P[r.Wns > 0, 0], 10 / 11 * xdata)
    assert r.eps == 0.3


@pytest.mark.parametrize("kind,Z", [("indicator", 0.047940112540007955), ("epa", 0.03196007502667197),
                                    ("epa_strict", 0.03196007502667197)])
def test_smc_evidence_kernels(oracle, kind, Z):
    """test/runtests.jl:268-423."""
    r = oracle.smc_run([("normal", 0.0, math.sqrt(10))], "gauss1d", [3.0, 1.0], 0.3, nparticles=5000, kind=kind, seed=202)
    assert Z * 0.9 <= math.exp(r.logZ) <= Z * 1.1
    u = np.random.default_rng(3).random(5000)
    inds = np.clip(oracle.wsample_stratified(r.Wns, u), 1, 5000) - 1      # `weightinds`, test/runtests.jl:13-19
    assert _isaround(r.P[inds, 0], 10 / 11 * 3.0)


def test_smc_bayes_factor_uniform_priors(oracle):
    """test/runtests.jl:220-266."""
    r1 = oracle.smc_run([("uniform", -10.0, 10.0)], "gauss1d", [3.0, 1.0], 0.3, nparticles=5000, seed=303)
    r2 = oracle.smc_run([("uniform", -20.0, 20.0)], "gauss1d", [3.0, 1.0], 0.3, nparticles=5000, seed=304)
    Z1, Z2 = math.exp(r1.logZ), math.exp(r2.logZ)
    assert 0.02998511 * 0.8 <= Z1 <= 0.02998511 * 1.2
    assert 0.01500489 * 0.8 <= Z2 <= 0.01500489 * 1.2
    assert 2.0 * 0.8 <= Z1 / Z2 <= 2.0 * 1.2
    assert _isaround(r1.P[r1.Wns > 0, 0], 3.0) and _isaround(r2.P[r2.Wns > 0, 0], 3.0)


def test_mc_posterior_1d_normal(oracle):
    """test/runtests.jl:149-161 (abcdemc! part)."""
    r = oracle.mc_run([("normal", 0.0, math.sqrt(10))], "gauss1d", [3.0, 1.0], 0.3, nparticles=5000, generations=500, seed=404)
    assert r.reached_eps
    assert _isaround(r.P[:, 0], 10 / 11 * 3.0)


def test_dirac_and_2d_and_normdu(oracle):
    """test/runtests.jl:493-535, :600-624."""
    r = oracle.smc_run([("normal", 1.0, 0.2)], "dirac", [1.5], 0.1, nparticles=100, seed=5)
    assert _isaround(r.P[r.Wns > 0, 0], 0.707)
    r = oracle.mc_run([("normal", 1.0, 0.2)], "dirac", [1.5], 0.1, nparticles=50, generations=20, seed=5)
    assert _isaround(r.P[:, 0], 0.707)
    pr = [("normal", 0.0, 5.0), ("normal", 0.0, 5.0)]
    for model in ("twod", "twod_inf"):
        r = oracle.smc_run(pr, model, [], 0.01, nparticles=500, seed=6)
        al = r.Wns > 0
        assert _isaround(r.P[al, 0], 1) and _isaround(r.P[al, 1], 1)
    r = oracle.smc_run([("normal", 1.0, 0.5), ("discrete_uniform", 1, 10)], "normdu", [5.5], 0.01, nparticles=100, seed=8)
    al = r.Wns > 0
    assert _isaround(r.P[al, 0], 1) and _isaround(r.P[al, 1], 5)
    assert np.all(r.P[:, 1] == np.rint(r.P[:, 1]))          # push_p rounds the discrete coordinate (:382)


def test_wiener_and_socks_and_mixture(oracle):
    """test/runtests.jl:425-491, :537-598."""
    t = np.arange(31.0)
    tdata = np.sqrt(0.25 * t * t + 4.0 * t) * 1.0             # brownianrms((0.5, 2.0)) with the mean factor
    r = oracle.smc_run([("uniform", 0.0, 1.0), ("uniform", 0.0, 4.0)], "wiener", tdata, 0.05, nparticles=1000, seed=9)
    al = r.Wns > 0
    assert _isaround(r.P[al, 0], 0.5, 2.0) and _isaround(r.P[al, 1], 2.0, 2.0)
    size = -30.0 ** 2 / (30.0 - 15.0 ** 2)
    pr = [("negbin", size, size / (30.0 + size)), ("beta", 15.0, 2.0)]
    r = oracle.smc_run(pr, "socks", [0.0, 11.0], 0.01, nparticles=5000, seed=10)
    al = r.Wns > 0
    assert _isaround(r.P[al, 0], 46.2) and _isaround(r.P[al, 1], 0.866)
    st_n = np.array([0.0, 0.04680825481526908, 0.1057221226763449, 0.2682111969397526, 0.8309228020477986])
    r = oracle.smc_run([("uniform", -10.0, 10.0)], "mixture", [0.0], 0.01, nparticles=2000, seed=11)
    x = r.P[r.Wns > 0, 0]
    qs = np.quantile(x, np.arange(0.1, 0.95, 0.1))
    stv = ((qs - qs[::-1]) / 2)[4:]
    assert np.mean(np.abs(stv - st_n)) < 0.1


# ---- portable math contract (DESIGN.md "Numerics"): accuracy vs libm ---------------------------
def test_portable_math_accuracy(oracle):
    import ctypes as C
    from scipy.special import gammaln
    L = oracle.lib()
    for f in ("orc_plog", "orc_pexp", "orc_plgamma"):
        getattr(L, f).restype = C.c_double; getattr(L, f).argtypes = [C.c_double]
    rng = np.random.default_rng(0)

    def ulps(got, want):
        return np.max(np.abs(got - want) / np.spacing(np.abs(want)))

    xs = np.concatenate([rng.random(20000), 10 ** rng.uniform(-300, 300, 20000), 1 - 2.0 ** -np.arange(1, 54), [2.0 ** -53, 0.5, 2.0, 1e-310]])
    got = np.array([L.orc_plog(float(x)) for x in xs])
    assert ulps(got, np.log(xs)) <= 2.0
    assert L.orc_plog(1.0) == 0.0 and L.orc_plog(0.0) == -INF and math.isnan(L.orc_plog(-1.0))
    xs = np.concatenate([rng.uniform(-700, 700, 20000), rng.uniform(-2, 2, 20000), [0.0, -745.0, 709.7]])
    got = np.array([L.orc_pexp(float(x)) for x in xs])
    assert ulps(got, np.exp(xs)) <= 2.0
    assert L.orc_pexp(-INF) == 0.0 and L.orc_pexp(0.0) == 1.0
    xs = np.concatenate([rng.uniform(0.01, 300, 20000), 10 ** rng.uniform(-5, 8, 5000)])
    got = np.array([L.orc_plgamma(float(x)) for x in xs])
    assert np.max(np.abs(got - gammaln(xs)) / np.maximum(1.0, np.abs(gammaln(xs)))) < 5e-14
    s, c = C.c_double(), C.c_double()
    us = np.concatenate([rng.random(20000), [0.0, 0.25, 0.5, 0.75, 0.125, 1 - 2.0 ** -53]])
    err = 0.0
    for u in us:
        L.orc_psincos2pi(C.c_double(u), C.byref(s), C.byref(c))
        err = max(err, abs(s.value - math.sin(2 * math.pi * u)), abs(c.value - math.cos(2 * math.pi * u)))
        assert abs(s.value ** 2 + c.value ** 2 - 1.0) < 1e-15
    assert err < 2e-15


# ---- the library's division shortcut for Normal z-scores (csrc/common.cuh pdiv_r) equals a / sigma --------------
def test_reciprocal_division_shortcut_is_exact(oracle):
    import ctypes as C
    L = oracle.lib()
    L.orc_pdiv_mismatches.restype = C.c_int64
    L.orc_pdiv_mismatches.argtypes = [C.c_int64, C.c_uint64, C.c_double]
    sigmas = [1.0, 2.0, 3.0, 10.0, math.sqrt(10.0), 0.1, 0.01, 1.0 / 3.0, 5.0, 1e-3, 123.456, 2.0 ** 0.5, 7.0, 1e5]
    rng = np.random.default_rng(7)
    sigmas += list(10 ** rng.uniform(-6, 6, 26))
    for k, s in enumerate(sigmas):
        assert L.orc_pdiv_mismatches(500_000, 1234 + k, float(s)) == 0, s
