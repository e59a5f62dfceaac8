"""GPU end-to-end tests of abcdesmc! / abcdemc! through the C ABI.

Tier 1: same Philox seed -> the GPU run follows the oracle's run decision by decision
(identical iteration count, simulation count, resampling events, Kmcmc history; eps / logZ /
ESS histories within 1e-9).  Tier 2: the reference's own statistical tests
(test/runtests.jl:110-624) with the reference's tolerances.
"""
import math

import numpy as np
import pytest

from test_gpu_parity import MODEL_CASES, to_prior

pytestmark = pytest.mark.gpu

SQ10 = math.sqrt(10)


def isaround(x, val, f=1.0):
    """test/runtests.jl:9"""
    x = np.asarray(x, dtype=float)
    return x.mean() - f * x.std(ddof=1) <= val <= x.mean() + f * x.std(ddof=1)


# ---------------------------------------------------------------------------------------------
# tier 1: decision-level parity of whole runs
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,eps_target,N,kw", [
    ("gauss1d", 0.3, 1000, {}),
    ("gauss1d", 0.3, 3001, dict(kernel="indicator")),
    ("gauss1d", 0.3, 2000, dict(kernel="epa", exact_scan=True)),
    ("gauss1d_blob", 0.3, 1500, dict(Kmcmc=5, Kmcmc_min=2.0)),
    ("gauss1d", 0.05, 2000, dict(nsims_max=20000)),
    ("gauss1d", 0.3, 2000, dict(facc_min=0.3, facc_tune=0.9, alpha=0.8, delta_ess=0.3)),
    ("gauss_corr10", 2.5, 4000, {}),
    ("twod_inf", 0.05, 1000, {}),
    ("normdu", 0.05, 500, {}),
    ("lotka_volterra", 0.3, 1200, dict(nsims_max=40000)),
    # round 2: the paths the verdict asked for
    ("gauss1d", 0.3, 3000, dict(alpha=0.2)),              # small alpha: the eps select leaves its window (generic passes; ranges_eps)
    ("gauss1d", 0.02, 2500, dict(alpha=0.5, delta_ess=0.9)),
    ("gauss_corr10", 3.0, 120000, dict(nsims_max=10**9)),  # config 2 at >= 10^5 particles, whole run
    ("birth_death", 3.0, 3000, dict(nsims_max=10**8)),     # config 5, whole run on one GPU
    ("lotka_volterra", 0.25, 1500, dict(nsims_max=10**8)), # config 4, uncapped
])
@pytest.mark.parametrize("fused", [True, False])
def test_smc_run_follows_oracle(A, oracle, gpu_ctx, name, eps_target, N, kw, fused):
    kw = dict(kw)
    spec, data = MODEL_CASES[name]
    kind = kw.pop("kernel", "indicator_strict")
    exact = kw.pop("exact_scan", False)
    okw = dict(nparticles=N, seed=4242, kind=kind, **kw)
    want = oracle.smc_run(spec, name, data, eps_target, **okw)
    gkw = {k: v for k, v in kw.items()}
    got = A.abcdesmc(to_prior(A, spec), A.Model(name, data), eps_target, None, nparticles=N, rng=4242, ABCk=kind,
                     exact_scan=exact, verbose=False, fused_head=fused, **gkw)
    assert (got.iters, got.nsims, got.status) == (want.iters, want.nsims, want.status)
    assert np.array_equal(got.Kmcmcs, want.hist["Kmcmc"])
    np.testing.assert_allclose(got.eps_hist, want.hist["eps"], rtol=1e-9)
    np.testing.assert_allclose(got.logZs, want.hist["logZ"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got.esss, want.hist["ess"], rtol=1e-9)
    np.testing.assert_allclose(got.faccs, want.hist["facc"], rtol=1e-12)
    np.testing.assert_allclose(got.gamma0s, want.hist["gamma0"], rtol=1e-15)
    np.testing.assert_allclose(got.ranges_eps[:, 0], want.hist["dmin"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got.ranges_eps[:, 1], want.hist["dmax"], rtol=1e-9, atol=1e-12)
    assert math.isclose(got.logZ, want.logZ, rel_tol=1e-9) and math.isclose(got.eps, want.eps, rel_tol=1e-9)
    assert np.array_equal(got.Wns > 0, want.Wns > 0)
    np.testing.assert_allclose(got.Wns, want.Wns, rtol=1e-9)
    np.testing.assert_allclose(got.P, want.P, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got.C, want.C, rtol=1e-9, atol=1e-10)
    if want.blobs.shape[1]:
        np.testing.assert_allclose(got.blobs.view(np.float64), want.blobs.view(np.float64), rtol=1e-9, atol=1e-10)
    if want.iters > 30:
        assert got.stats["n_resamples"] >= 1


@pytest.mark.parametrize("name,eps_target,N,gens", [("gauss1d", 0.3, 1000, 60), ("twod", 0.05, 500, 80), ("dirac", 0.1, 50, 20),
                                                    ("gauss1d", 0.3, 20000, 40),          # radix-sort path, device-side generation loop
                                                    ("lotka_volterra", 0.3, 800, 30), ("lotka_volterra_lin", 0.5, 800, 30),
                                                    ("birth_death", 3.0, 1000, 30)])
def test_mc_run_follows_oracle(A, oracle, gpu_ctx, name, eps_target, N, gens):
    spec, data = MODEL_CASES[name]
    want = oracle.mc_run(spec, name, data, eps_target, nparticles=N, generations=gens, seed=777)
    got = A.abcdemc(to_prior(A, spec), A.Model(name, data), eps_target, None, nparticles=N, generations=gens, rng=777, verbose=False)
    assert (got.nsims, got.reached_eps) == (want.nsims, want.reached_eps)
    np.testing.assert_allclose(got.P, want.P, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got.C, want.C, rtol=1e-9, atol=1e-10)


def test_sync_every_does_not_change_results(A, gpu_ctx):
    """Running ahead of the host (several iterations enqueued per stop-flag poll) is invisible."""
    spec, data = MODEL_CASES["gauss1d"]
    r1 = A.abcdesmc(to_prior(A, spec), A.Model("gauss1d", data), 0.3, None, nparticles=2000, rng=5, verbose=False)
    r2 = A.abcdesmc(to_prior(A, spec), A.Model("gauss1d", data), 0.3, None, nparticles=2000, rng=5, verbose=False, sync_every=8)
    assert r1.iters == r2.iters and r1.logZ == r2.logZ and np.array_equal(r1.P, r2.P) and np.array_equal(r1.Wns, r2.Wns)


def test_run_is_deterministic(A, gpu_ctx):
    spec, data = MODEL_CASES["gauss_corr10"]
    rs = [A.abcdesmc(to_prior(A, spec), A.Model("gauss_corr10", data), 2.0, None, nparticles=20000, rng=9, verbose=False) for _ in range(2)]
    assert rs[0].logZ == rs[1].logZ and np.array_equal(rs[0].P, rs[1].P) and np.array_equal(rs[0].esss, rs[1].esss)


# ---------------------------------------------------------------------------------------------
# argument validation, src/abcdez_smc.jl:223-235, src/abcdez_mc.jl:108-110
# ---------------------------------------------------------------------------------------------
def test_argument_errors(A, gpu_ctx):
    pr, m = A.host.Normal(0, SQ10), A.Model("gauss1d", [3.0, 1.0])
    bad = [(dict(alpha=1.0), "α must be in 0 <= α < 1"), (dict(delta_ess=1.5), "δess must be in"),
           (dict(facc_stop=-0.1), "facc_stop must be"), (dict(facc_min=2.0), "facc_min must be"),
           (dict(facc_tune=1.5), "facc_tune must be"), (dict(Kmcmc=0), "Kmcmc must be at least 1"),
           (dict(Kmcmc_min=-1.0), "Kmcmc_min must be"), (dict(nsims_max=0), "nsims_max must be at least 1"),
           (dict(nparticles=5), "nparticles must be at least 6")]
    for kw, msg in bad:
        with pytest.raises(A.ABCdeZError, match=msg):
            A.abcdesmc(pr, m, 0.3, None, verbose=False, **kw)
    with pytest.raises(A.ABCdeZError, match="ϵ_target must be non-negative"):
        A.abcdesmc(pr, m, -0.3, None, verbose=False)
    with pytest.raises(A.ABCdeZError, match="nparticles must be at least 5"):
        A.abcdemc(pr, m, 0.3, None, nparticles=4, verbose=False)
    with pytest.raises(A.ABCdeZError, match="generations must be at least 1"):
        A.abcdemc(pr, m, 0.3, None, generations=0, verbose=False)
    # Greek keyword spellings of the reference
    r = A.abcdesmc(pr, m, 0.3, None, nparticles=200, α=0.9, δess=0.4, rng=1, verbose=False)
    assert r.ϵ == 0.3 and r.P.shape == (200,)


# ---------------------------------------------------------------------------------------------
# tier 2: the reference's statistical tests (independent Philox streams)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("xdata,Z,tol", [(3.0, 0.047940112540007955, 0.10), (7.0, 0.007781668367620676, 0.20)])
def test_1d_normal_evidence(A, gpu_ctx, xdata, Z, tol):
    """test/runtests.jl:110-218: abcdesmc! and abcdemc!, N = 5000."""
    pr, m = A.host.Normal(0, SQ10), A.Model("gauss1d", [xdata, 1.0])
    r = A.abcdesmc(pr, m, 0.3, None, nparticles=5000, verbose=False, rng=1001)
    assert Z * (1 - tol) <= math.exp(r.logZ) <= Z * (1 + tol)
    assert isaround(r.P[r.Wns > 0], 10 / 11 * xdata)
    rmc = A.abcdemc(pr, m, 0.3, None, nparticles=5000, generations=500, verbose=False, rng=1002)
    assert isaround(rmc.P, 10 / 11 * xdata)
    assert rmc.reached_eps


def test_evidence_tight_at_large_N(A, gpu_ctx):
    """What the GPU buys: N = 10^6 brings the evidence estimate within 1 % of the analytic value."""
    r = A.abcdesmc(A.host.Normal(0, SQ10), A.Model("gauss1d", [3.0, 1.0]), 0.3, None, nparticles=1_000_000,
                   nsims_max=10**10, verbose=False, rng=31337)
    assert abs(math.exp(r.logZ) / 0.047940112540007955 - 1.0) < 0.01
    x = r.P[r.Wns > 0]
    assert abs(x.mean() - 30 / 11) < 0.01 and abs(x.var() - 10 / 11) < 0.05


def test_config2_logZ_against_analytic(A, gpu_ctx):
    """BASELINE.json configs[1]: 10-d correlated Gaussian model, 10^6 particles, logZ vs the analytic evidence.
    Z = P(||Y - y_obs|| < eps), Y ~ N(0, Sigma + sigma0^2 I), evaluated by integrating the Gaussian density over the
    eps-ball (4*10^6 uniform points in the ball, relative standard error 3e-5): log Z = -16.45296.  Run-to-run
    sd(logZ) at N = 10^6 is 0.04 (bench.py logZ_sd), so 0.2 is a five-sigma band.  Posterior mean for eps -> 0:
    sigma0^2 (Sigma + sigma0^2 I)^-1 y_obs; after ~320 iterations of 5 % selection with three DE-MCMC sweeps each the
    particles are strongly correlated (that is the algorithm, the oracle does the same), so the mean of one run is
    only good to a few tenths of a posterior standard deviation (~0.9)."""
    d, rho, s0 = 10, 0.5, 2.0
    y = np.array([0.5 * math.sin(1.0 + k) for k in range(d)])
    pr = A.Factored(*[A.host.Normal(0.0, s0)] * d)
    r = A.abcdesmc(pr, A.Model("gauss_corr10", list(y) + [rho]), 1.0, None, nparticles=1_000_000, nsims_max=10**11,
                   verbose=False, rng=20261017)
    assert r.eps == 1.0
    assert abs(r.logZ - (-16.45296)) < 0.2
    S = np.array([[rho ** abs(i - j) for j in range(d)] for i in range(d)]) + s0 ** 2 * np.eye(d)
    want = s0 ** 2 * np.linalg.solve(S, y)
    P = np.asarray(r.P)[np.asarray(r.Wns) > 0]
    assert np.max(np.abs(P.mean(axis=0) - want)) < 0.4
    assert 0.5 < P.std(axis=0).mean() < 1.2


def test_bayes_factor_uniform_priors(A, gpu_ctx):
    """test/runtests.jl:220-266."""
    m = A.Model("gauss1d", [3.0, 1.0])
    r1 = A.abcdesmc(A.host.Uniform(-10, 10), m, 0.3, None, nparticles=5000, verbose=False, rng=11)
    r2 = A.abcdesmc(A.host.Uniform(-20, 20), m, 0.3, None, nparticles=5000, verbose=False, rng=12)
    Z1, Z2 = math.exp(r1.logZ), math.exp(r2.logZ)
    assert 0.02998511 * 0.8 <= Z1 <= 0.02998511 * 1.2 and 0.01500489 * 0.8 <= Z2 <= 0.01500489 * 1.2
    assert 1.6 <= Z1 / Z2 <= 2.4
    assert isaround(r1.P[r1.Wns > 0], 3.0) and isaround(r2.P[r2.Wns > 0], 3.0)


@pytest.mark.parametrize("kind,Z", [("indicator", 0.047940112540007955), ("epa", 0.03196007502667197),
                                    ("epa_strict", 0.03196007502667197)])
def test_kernel_variants(A, gpu_ctx, kind, Z):
    """test/runtests.jl:268-423 incl. `weightinds` through the library's wsample_stratified."""
    r = A.abcdesmc(A.host.Normal(0, SQ10), A.Model("gauss1d", [3.0, 1.0]), 0.3, None, ABCk=kind, nparticles=5000,
                   verbose=False, rng=21)
    assert Z * 0.9 <= math.exp(r.logZ) <= Z * 1.1
    if kind == "indicator":
        x = r.P[r.Wns > 0]
    else:
        assert math.isclose(r.Wns.sum(), 1.0, rel_tol=1e-9)
        inds = A.wsample_stratified(r.Wns, np.random.default_rng(1).random(5000)) - 1
        x = r.P[inds]
    assert isaround(x, 10 / 11 * 3.0)


def test_socks(A, gpu_ctx):
    """test/runtests.jl:425-491."""
    size = -30.0 ** 2 / (30.0 - 15.0 ** 2)
    pr = A.Factored(A.host.NegativeBinomial(size, size / (30.0 + size)), A.host.Beta(15, 2))
    m = A.Model("socks", [0.0, 11.0])
    r = A.abcdesmc(pr, m, 0.01, None, nparticles=5000, verbose=False, rng=31)
    al = r.Wns > 0
    assert isaround(r.P[al, 0], 46.2) and isaround(r.P[al, 1], 0.866)
    assert np.all(r.P[:, 0] == np.rint(r.P[:, 0]))
    rmc = A.abcdemc(pr, m, 0.01, None, nparticles=5000, generations=500, verbose=False, rng=32)
    assert isaround(rmc.P[:, 0], 46.2) and isaround(rmc.P[:, 1], 0.866)


def test_dirac_normdu_wiener_mixture_2d(A, gpu_ctx):
    """test/runtests.jl:493-624 (default nparticles where the reference uses defaults)."""
    pr, m = A.host.Normal(1, 0.2), A.Model("dirac", [1.5])
    assert isaround(A.abcdemc(pr, m, 0.1, None, verbose=False, rng=41).P, 0.707)
    r = A.abcdesmc(pr, m, 0.1, None, verbose=False, rng=42)
    assert isaround(r.P[r.Wns > 0], 0.707)

    pr = A.Factored(A.host.Normal(1, 0.5), A.host.DiscreteUniform(1, 10)); m = A.Model("normdu", [5.5])
    rm = A.abcdemc(pr, m, 0.01, None, nparticles=100, generations=1000, verbose=False, rng=43).P
    assert isaround(rm[:, 0], 1) and isaround(rm[:, 1], 5)
    r = A.abcdesmc(pr, m, 0.01, None, nparticles=100, verbose=False, rng=44)
    al = r.Wns > 0
    assert isaround(r.P[al, 0], 1) and isaround(r.P[al, 1], 5)

    t = np.arange(31.0); tdata = np.sqrt(0.25 * t * t + 4.0 * t) * 1.003
    pr = A.Factored(A.host.Uniform(0, 1), A.host.Uniform(0, 4)); m = A.Model("wiener", tdata)
    rm = A.abcdemc(pr, m, 0.05, None, nparticles=1000, generations=300, verbose=False, rng=45).P
    assert isaround(rm[:, 0], 0.5, 2.0) and isaround(rm[:, 1], 2.0, 2.0)
    r = A.abcdesmc(pr, m, 0.05, None, nparticles=1000, verbose=False, rng=46)
    al = r.Wns > 0
    assert isaround(r.P[al, 0], 0.5, 2.0) and isaround(r.P[al, 1], 2.0, 2.0)

    st_n = np.array([0.0, 0.04680825481526908, 0.1057221226763449, 0.2682111969397526, 0.8309228020477986])
    def stv(x):
        qs = np.quantile(x, np.arange(0.1, 0.95, 0.1)); return ((qs - qs[::-1]) / 2)[4:]
    pr, m = A.host.Uniform(-10, 10), A.Model("mixture", [0.0])
    rm = A.abcdemc(pr, m, 0.01, None, nparticles=2000, generations=1000, verbose=False, rng=47).P
    r = A.abcdesmc(pr, m, 0.01, None, nparticles=2000, verbose=False, rng=48)
    assert np.mean(np.abs(stv(rm) - st_n)) < 0.1 and np.mean(np.abs(stv(r.P[r.Wns > 0]) - st_n)) < 0.1

    pr = A.Factored(A.host.Normal(0, 5), A.host.Normal(0, 5))
    for name in ("twod", "twod_inf"):
        m = A.Model(name, [])
        rm = A.abcdemc(pr, m, 0.01, None, nparticles=500, generations=500, verbose=False, rng=49).P
        assert isaround(rm[:, 0], 1) and isaround(rm[:, 1], 1)
        r = A.abcdesmc(pr, m, 0.01, None, nparticles=500, verbose=False, rng=50)
        al = r.Wns > 0
        assert isaround(r.P[al, 0], 1) and isaround(r.P[al, 1], 1)


def test_minimal_example_model_probabilities(A, gpu_ctx):
    """examples/minimal_example.jl (config 1): two models differing in the prior, N = 1000 in the
    reference; here averaged over 8 seeds to make the 0.678 / 0.322 check sharp."""
    m = A.Model("gauss1d", [3.0, 1.0])
    Z1 = np.mean([math.exp(A.abcdesmc(A.host.Normal(0, SQ10), m, 0.3, None, nparticles=1000, verbose=False, rng=s).logZ) for s in range(8)])
    Z2 = np.mean([math.exp(A.abcdesmc(A.host.Normal(0, 10.0), m, 0.3, None, nparticles=1000, verbose=False, rng=100 + s).logZ) for s in range(8)])
    assert abs(Z1 / 0.047940 - 1) < 0.1 and abs(Z2 / 0.022780 - 1) < 0.12
    assert abs(Z1 / (Z1 + Z2) - 0.678) < 0.04


def test_blobs_travel_with_particles(A, gpu_ctx):
    """docs/src/index.md:298-324: the blob returned with a particle is the simulation that produced its distance."""
    r = A.abcdesmc(A.host.Normal(0, SQ10), A.Model("gauss1d_blob", [3.0, 1.0]), 0.3, None, nparticles=3000, verbose=False, rng=77)
    y = r.blobs.view(np.float64)[:, 0]
    np.testing.assert_allclose(np.abs(y - 3.0), r.C, rtol=1e-12, atol=1e-12)
    assert np.all(r.C[r.Wns > 0] < 0.3)


# ---------------------------------------------------------------------------------------------
# config 4: Lotka-Volterra model comparison (the pattern of test/runtests.jl:220-266: evidences against ABC-rejection
# ground truth, and their ratio)
# ---------------------------------------------------------------------------------------------
LV_OBS = [1.4385, 0.5655, 1.9586, 0.7589, 2.3239, 1.2022, 2.0925, 1.9724, 1.3201, 2.6696, 0.6896, 2.7822, 0.3821, 2.4679, 0.2525, 2.0363]
LV_DATA = [1.0, 0.5, 0.01, 50, 8, 0.05] + LV_OBS        # trajectory of the classical model at theta* = (1.2, 0.9, 0.7, 0.6)


def test_lotka_volterra_bayes_factor(A, gpu_ctx):
    """Two competing ODE models for the same observations, abcdesmc! evidences vs ABC rejection from the prior
    (2 * 10^6 prior draws per model, simulated with the same device functors), and abcdemc! posteriors for both."""
    spec = [("uniform", 0.0, 2.0)] * 4
    prior = to_prior(A, spec)
    eps = 0.7
    Z, Zrej = {}, {}
    for name in ("lotka_volterra", "lotka_volterra_lin"):
        m = A.Model(name, LV_DATA)
        M = 2_000_000
        th = prior.rand(M, seed=101)
        d, _ = m.simulate(th, seed=202)
        Zrej[name] = float(np.mean(d < eps))                  # indicator kernel: Z(eps) = P_prior(dist < eps)
        assert Zrej[name] * M > 150, (name, Zrej[name])       # enough accepted draws for a 10 % ground truth
        lz = [A.abcdesmc(prior, m, eps, None, nparticles=20000, rng=7 + k, verbose=False, nsims_max=10**9).logZ for k in range(3)]
        Z[name] = float(np.exp(np.mean(lz)))
        assert abs(Z[name] / Zrej[name] - 1.0) < 0.25, (name, Z[name], Zrej[name])   # reference tolerance: 20 % (+ the ground truth's own error)
    bf, bf_rej = Z["lotka_volterra"] / Z["lotka_volterra_lin"], Zrej["lotka_volterra"] / Zrej["lotka_volterra_lin"]
    assert abs(bf / bf_rej - 1.0) < 0.35, (bf, bf_rej)
    p = A.host.model_probabilities(np.log([Z["lotka_volterra"], Z["lotka_volterra_lin"]]))
    assert abs(p.sum() - 1.0) < 1e-12 and (p[0] > p[1]) == (bf_rej > 1.0)
    # abcdemc! on both models: the generating model reaches a tighter tolerance, and its posterior sits on theta*
    r = A.abcdemc(prior, A.Model("lotka_volterra", LV_DATA), 0.25, None, nparticles=2000, generations=150, rng=3, verbose=False)
    assert np.median(r.C) < 0.35
    assert np.all(np.abs(np.median(r.P, axis=0) - np.array([1.2, 0.9, 0.7, 0.6])) < 0.35)
    rl = A.abcdemc(prior, A.Model("lotka_volterra_lin", LV_DATA), 0.25, None, nparticles=2000, generations=150, rng=3, verbose=False)
    assert np.median(rl.C) > np.median(r.C)


# ---------------------------------------------------------------------------------------------
# config 3: g-and-k (CTA-cooperative simulator, FP32; csrc/gk.cu)
# ---------------------------------------------------------------------------------------------
GK_TRUE = (3.0, 1.0, 2.0, 0.5)
GK_PRIOR = [("uniform", 0.0, 10.0)] * 4


def _gk_octiles(theta, n=200000, seed=1):
    A_, B_, g, k = theta
    z = np.random.default_rng(seed).standard_normal(n)
    x = A_ + B_ * (1 + 0.8 * np.tanh(0.5 * g * z)) * (1 + z * z) ** k * z
    xs = np.sort(x)
    return [float(xs[(n * j + 7) // 8 - 1]) for j in range(1, 8)]


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gk", "gk_f32"])
@pytest.mark.parametrize("n", [10000, 16384, 4096, 4095, 1001, 8])
def test_gk_simulate_parity(A, oracle, gpu_ctx, model, n):
    """One dist! evaluation per row: same Philox blocks, the same portable arithmetic operation by operation (FP64 for
    "gk" -- the precision of a Julia dist! --, FP32 for the relaxed mode "gk_f32"), a key-space multi-select vs qsort
    for the octiles: the distances are BIT-IDENTICAL (north star: <= 1e-6 relative).  From n = 4096 on the kernel
    selects in z space and transforms only the candidates (gk.cu, "the fast path"); rows 5-12 are the parameter corners
    of that path and of its fallback."""
    data = [float(n)] + _gk_octiles(GK_TRUE)
    N = 1200 if n >= 4096 else 300
    th = oracle.prior_sample(GK_PRIOR, N, seed=3)
    th[5] = [3.0, 0.0, 1.0, 0.5]                    # B = 0: every draw equals A (one key fills every bucket)
    th[6] = [0.0, 1.0, 0.0, 0.0]                    # plain normal draws around 0: keys on both sides of zero
    th[7] = [3.0, -1.0, 2.0, 0.5]                   # B < 0: Q decreasing -> generic path at every n
    th[8] = [3.0, 1.0, 2.0, -0.3]                   # k < 0 -> generic path
    th[9] = [10.0, 1e-13, 2.0, 0.5]                 # x is a staircase of ulp(10) steps: plateaus wider than a bucket
    th[10] = [0.0, 1.0, 50.0, 9.9]                  # steep tanh, huge (1 + z^2)^k
    th[11] = [1e-3, 1e3, -7.0, 0.0]                 # negative g
    th[12] = [5.0, 1.0, 0.0, 1e-300]
    want, _ = oracle.simulate(model, data, th, seed=11, epoch=2)
    got, _ = A.Model(model, data).simulate(th, seed=11, epoch=2)
    np.testing.assert_array_equal(got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["gk", "gk_f32"])
def test_gk_init_and_sweep_follow_oracle(A, oracle, gpu_ctx, model):
    """abcde_init! and one injected-randomness abcdesmc_swarm! sweep of the CTA-cooperative kernels: theta, distances
    and accept decisions bit-identical to the oracle."""
    data = [2000.0] + _gk_octiles(GK_TRUE)
    spec = [("uniform", 0.0, 10.0)] * 3 + [("uniform", 0.0, 1.0)]
    N = 1500
    wth, wlp, wdl, _, wred = oracle.init(spec, model, data, N, seed=21)
    fam = {"uniform": A.host.Uniform}
    prior = A.Factored(*[fam[s[0]](*s[1:]) for s in spec])
    pop = A.Population(prior, A.Model(model, data), N)
    red = pop.init(seed=21)
    g = pop.download()
    assert red == wred
    np.testing.assert_array_equal(g["theta"], wth)
    np.testing.assert_array_equal(g["delta"], wdl)
    rng = np.random.default_rng(5)
    a = rng.integers(0, N, N).astype(np.int32); b = rng.integers(0, N, N).astype(np.int32)
    idx = np.arange(N)
    a = np.where(a == idx, (a + 1) % N, a).astype(np.int32)
    b = np.where((b == idx) | (b == a), (b + 2) % N, b).astype(np.int32)
    b = np.where((b == idx) | (b == a), (b + 1) % N, b).astype(np.int32)
    z = rng.standard_normal(N); u = rng.random(N)
    eps = float(np.quantile(wdl, 0.6))
    alive = np.ones(N, dtype=np.uint8)
    w = oracle.smc_sweep(spec, model, data, wth, wlp, wdl, alive, eps, "indicator_strict", gamma0=2.38 / math.sqrt(8), gsig=1e-5,
                         seed=4, epoch=1, a=a, b=b, z=z, u=u)
    pop.set(eps=eps, kernel="indicator_strict", gamma0=2.38 / math.sqrt(8), seed=4, epoch=1)
    flags = pop.smc_sweep(a=a, b=b, z=z, u=u, want_flags=True)["flags"]
    got = pop.download()
    assert np.array_equal(flags, w["flags"])                           # simulated AND accepted flags, bit for bit
    np.testing.assert_array_equal(got["delta"], w["delta"])
    np.testing.assert_array_equal(got["theta"], w["theta"])
    pop.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model,N", [("gk", 600), ("gk_f32", 600)])
def test_gk_whole_runs_follow_oracle(A, oracle, gpu_ctx, model, N):
    """config 3's simulator through complete abcdesmc! and abcdemc! runs: decision by decision like the oracle."""
    data = [1000.0] + _gk_octiles(GK_TRUE)
    spec = [("uniform", 0.0, 10.0)] * 3 + [("uniform", 0.0, 2.0)]
    prior = A.Factored(*[A.host.Uniform(*s[1:]) for s in spec])
    want = oracle.smc_run(spec, model, data, 0.5, nparticles=N, seed=12, nsims_max=60000)
    got = A.abcdesmc(prior, A.Model(model, data), 0.5, None, nparticles=N, rng=12, verbose=False, nsims_max=60000)
    assert (got.iters, got.nsims) == (want.iters, want.nsims)
    np.testing.assert_array_equal(got.eps_hist, want.hist["eps"])
    np.testing.assert_allclose(got.logZs, want.hist["logZ"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(got.P, want.P, rtol=1e-12)
    np.testing.assert_array_equal(got.C, want.C)
    wm = oracle.mc_run(spec, model, data, 0.6, nparticles=300, generations=25, seed=13)
    gm = A.abcdemc(prior, A.Model(model, data), 0.6, None, nparticles=300, generations=25, rng=13, verbose=False)
    assert (gm.nsims, gm.reached_eps) == (wm.nsims, wm.reached_eps)
    np.testing.assert_allclose(gm.P, wm.P, rtol=1e-12)
    np.testing.assert_array_equal(gm.C, wm.C)


@pytest.mark.gpu
def test_gk_posterior_recovers_parameters(A, gpu_ctx):
    """Statistical tier: abcdesmc! on g-and-k octiles concentrates around the generating parameters -- in FP64 and
    in the relaxed FP32 mode, which must agree with each other within Monte-Carlo error."""
    data = [10000.0] + _gk_octiles(GK_TRUE)
    prior = A.Factored(*[A.host.Uniform(0.0, 10.0)] * 3, A.host.Uniform(0.0, 2.0))
    means = {}
    for model in ("gk", "gk_f32"):
        r = A.abcdesmc(prior, A.Model(model, data), 0.15, None, nparticles=2000, rng=8, verbose=False, nsims_max=400000)
        assert r.eps <= 0.6
        w = r.Wns / r.Wns.sum()
        mean = (r.P * w[:, None]).sum(0)
        assert abs(mean[0] - GK_TRUE[0]) < 0.25 and abs(mean[1] - GK_TRUE[1]) < 0.5
        assert abs(mean[3] - GK_TRUE[3]) < 0.3
        means[model] = mean
    assert np.all(np.abs(means["gk"] - means["gk_f32"]) < [0.2, 0.4, 1.5, 0.25])


# ---------------------------------------------------------------------------------------------
# run-state snapshots (abcdez_smc_run_state): stop between two iterations, continue later
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,eps_target,N,cut", [("gauss1d", 0.3, 2000, 7), ("gauss_corr10", 2.5, 20000, 25),
                                                  ("birth_death", 3.0, 1500, 5), ("gauss1d_blob", 0.3, 1000, 1)])
def test_run_state_resume_is_the_uninterrupted_run(A, gpu_ctx, name, eps_target, N, cut):
    spec, data = MODEL_CASES[name]
    pr, m = to_prior(A, spec), A.Model(name, data)
    kw = dict(nparticles=N, nsims_max=10**9, verbose=False, rng=77)
    full = A.abcdesmc(pr, m, eps_target, None, **kw)
    assert full.iters >= 5
    cut = max(1, min(cut, full.iters - 3))
    part = A.abcdesmc(pr, m, eps_target, None, max_iters=cut, return_state=True, **kw)
    assert part.iters == cut and part.state is not None and len(part.state) > 8 * N
    assert np.array_equal(part.eps_hist, full.eps_hist[:cut + 1])
    mid = A.abcdesmc(pr, m, eps_target, None, state=part.state, max_iters=2, return_state=True, **kw)
    assert mid.iters == cut + 2
    rest = A.abcdesmc(pr, m, eps_target, None, state=mid.state, **kw)
    assert (rest.iters, rest.nsims, rest.eps, rest.logZ) == (full.iters, full.nsims, full.eps, full.logZ)
    for a, b in ((rest.P, full.P), (rest.Wns, full.Wns), (rest.C, full.C), (rest.eps_hist, full.eps_hist),
                 (rest.logZs, full.logZs), (rest.esss, full.esss), (rest.faccs, full.faccs), (rest.Kmcmcs, full.Kmcmcs),
                 (rest.ranges_eps, full.ranges_eps)):
        assert np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
    if m.blob_bytes:
        assert np.array_equal(rest.blobs, full.blobs)


def test_run_state_continue_to_a_smaller_target_and_errors(A, gpu_ctx):
    pr, m = A.host.Normal(0, SQ10), A.Model("gauss1d", [3.0, 1.0])
    kw = dict(nparticles=3000, verbose=False, rng=5)
    direct = A.abcdesmc(pr, m, 0.3, None, **kw)
    first = A.abcdesmc(pr, m, 1.0, None, return_state=True, **kw)
    assert first.eps == 1.0 and first.iters < direct.iters
    more = A.abcdesmc(pr, m, 0.3, None, state=first.state, **kw)          # same schedule below eps = 1, clamped later
    assert more.eps == 0.3 and more.iters >= first.iters
    assert abs(more.logZ - direct.logZ) < 0.25
    with pytest.raises(A.ABCdeZError):                                     # not a snapshot of this population size
        A.abcdesmc(pr, m, 0.3, None, state=first.state, nparticles=2999, verbose=False, rng=5)
    with pytest.raises(A.ABCdeZError):                                     # options other than the targets must match
        A.abcdesmc(pr, m, 0.3, None, state=first.state, alpha=0.9, **kw)
    with pytest.raises(A.ABCdeZError):
        A.abcdesmc(pr, m, 0.3, None, state=first.state[:100], **kw)


# ---------------------------------------------------------------------------------------------
# after the run: weightinds / posterior sample / model probabilities at a matched tolerance / repeated evidences
# ---------------------------------------------------------------------------------------------
def test_post_run_helpers(A, gpu_ctx):
    """examples/minimal_example.jl:10-65 end to end: two priors on the same data, posterior model probabilities
    0.678 / 0.322 analytically; the resampled posterior has the weighted posterior's moments."""
    m = A.Model("gauss1d", [3.0, 1.0])
    r1 = A.abcdesmc(A.host.Normal(0, SQ10), m, 0.3, None, nparticles=20000, verbose=False, rng=101)
    r2 = A.abcdesmc(A.host.Normal(0, 10.0), m, 0.3, None, nparticles=20000, verbose=False, rng=102)
    p = A.host.model_probabilities([r1, r2])
    assert abs(p[0] - 0.678) < 0.03 and abs(p.sum() - 1.0) < 1e-12
    assert np.allclose(A.host.model_probabilities([r1.logZ, r2.logZ]), p)
    # a run stopped at a larger tolerance is compared on the ladders (docs/src/index.md:282-284)
    r2b = A.abcdesmc(A.host.Normal(0, 10.0), m, 0.5, None, nparticles=20000, verbose=False, rng=103)
    pb = A.host.model_probabilities([r1, r2b])
    assert A.host.evidence_at(r1, r2b.eps) > r1.logZ and 0.55 < pb[0] < 0.8
    post = A.host.posterior_sample(r1, rng=7)
    w = r1.Wns > 0
    assert post.shape == r1.P.shape and abs(post.mean() - r1.P[w].mean()) < 0.02 and abs(post.std() - r1.P[w].std()) < 0.02
    mean, sd, lz = A.host.evidence_uncertainty(A.host.Normal(0, SQ10), m, 0.3, repeats=6, rng=9, nparticles=5000)
    assert len(lz) == 6 and len(set(lz)) == 6 and abs(mean - math.log(0.047940112540007955)) < 0.1 and 0.0 < sd < 0.15


# ---------------------------------------------------------------------------------------------
# batched runs and the device-side posterior sample
# ---------------------------------------------------------------------------------------------
def test_batched_runs_equal_their_own_calls(A, gpu_ctx):
    """abcdez_smc_run_batch: replicates and different models in flight together; every run is bit for bit the run its
    own abcdesmc! call gives (examples/minimal_example.jl:27-65: two models on the same data)."""
    m = A.Model("gauss1d", [3.0, 1.0])
    runs = [dict(prior=A.host.Normal(0, SQ10), dist=m, eps_target=0.3, nparticles=1000, rng=500 + k) for k in range(20)]
    runs += [dict(prior=A.host.Normal(0, 10.0), dist=m, eps_target=0.3, nparticles=1000, rng=600 + k) for k in range(20)]
    spec, data = MODEL_CASES["gauss_corr10"]
    runs.append(dict(prior=to_prior(A, spec), dist=A.Model("gauss_corr10", data), eps_target=3.0, nparticles=8000, rng=3, nsims_max=10**8))
    got = A.abcdesmc_batch(runs, verboseout=True)
    assert len(got) == 41
    for k in (0, 7, 19, 20, 33, 40):
        r = dict(runs[k]); pr = r.pop("prior"); d = r.pop("dist"); e = r.pop("eps_target")
        one = A.abcdesmc(pr, d, e, None, verbose=False, **r)
        assert (one.iters, one.nsims, one.logZ, one.eps) == (got[k].iters, got[k].nsims, got[k].logZ, got[k].eps), k
        assert np.array_equal(one.P, got[k].P) and np.array_equal(one.Wns, got[k].Wns) and np.array_equal(one.eps_hist, got[k].eps_hist)
    z1 = np.exp(np.mean([r.logZ for r in got[:20]])); z2 = np.exp(np.mean([r.logZ for r in got[20:40]]))
    assert abs(z1 / 0.047940112540007955 - 1) < 0.1 and abs(z1 / (z1 + z2) - 0.678) < 0.03


def test_posterior_sample_on_device(A, oracle, gpu_ctx):
    """abcdez_posterior_sample == P[weightinds(Wns)] (test/runtests.jl:13-19,287-291): the returned indices are the oracle's
    wsample_stratified! indices for the same Philox uniforms, and the rows are gathered accordingly."""
    import ctypes as C
    r = A.abcdesmc(to_prior(A, MODEL_CASES["twod"][0]), A.Model("twod", []), 0.05, None, nparticles=5000, verbose=False, rng=11, ABCk="epa")
    N, d = r.P.shape
    out = np.empty_like(r.P); inds = np.empty(N, dtype=np.int64)
    rc = A.lib().abcdez_posterior_sample(gpu_ctx._h, C.c_int64(N), d, r.P.ctypes.data_as(C.c_void_p), r.Wns.ctypes.data_as(C.c_void_p),
                                         C.c_uint64(77), out.ctypes.data_as(C.c_void_p), inds.ctypes.data_as(C.c_void_p))
    assert rc == 0
    u = oracle.resample_uniforms(N, 77, 0)
    want = oracle.wsample_stratified(r.Wns, u)
    assert np.array_equal(inds, want)
    assert np.array_equal(out, r.P[np.clip(want, 1, N) - 1])
    assert np.all(r.Wns[np.clip(inds, 1, N) - 1] > 0)


# ---------------------------------------------------------------------------------------------
# relaxed-parity performance modes -- statistical tests with the reference's tolerances
# (test/runtests.jl:147,159: evidence within 10 %, posterior mean within one sample sd)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [dict(systematic_resampling=True), dict(partner_segments=True),
                                  dict(systematic_resampling=True, partner_segments=True), dict(fp32_state=True),
                                  dict(fp32_state=True, partner_segments=True, systematic_resampling=True)])
def test_relaxed_modes_keep_evidence_and_posterior(A, gpu_ctx, mode):
    m = A.Model("gauss1d", [3.0, 1.0])
    lz, means = [], []
    for k in range(4):
        r = A.abcdesmc(A.host.Normal(0, SQ10), m, 0.3, None, nparticles=5000, verbose=False, rng=900 + k, **mode)
        lz.append(r.logZ); means.append(A.host.posterior_sample(r, rng=k))
        if k == 0:
            base = A.abcdesmc(A.host.Normal(0, SQ10), m, 0.3, None, nparticles=5000, verbose=False, rng=900)
            assert base.logZ != r.logZ or base.nsims != r.nsims            # the mode does change the run
    Z = float(np.exp(np.mean(lz)))
    assert abs(Z / 0.047940112540007955 - 1.0) < 0.1
    assert isaround(np.concatenate(means), 2.7272727272727275)
    # config 2 (d = 10): logZ against the analytic value, as for the parity mode
    spec, data = MODEL_CASES["gauss_corr10"]
    r = A.abcdesmc(to_prior(A, spec), A.Model("gauss_corr10", data), 2.0, None, nparticles=100000, verbose=False, rng=5, nsims_max=10**10, **mode)
    r0 = A.abcdesmc(to_prior(A, spec), A.Model("gauss_corr10", data), 2.0, None, nparticles=100000, verbose=False, rng=5, nsims_max=10**10)
    assert abs(r.logZ - r0.logZ) < 0.1 and abs(r.iters - r0.iters) <= 3
    w = r.Wns / r.Wns.sum(); w0 = r0.Wns / r0.Wns.sum()
    # (two independent Monte-Carlo runs: the posterior sd of a coordinate is ~0.9, the runs' effective sample sizes a few thousand)
    assert np.all(np.abs((r.P * w[:, None]).sum(0) - (r0.P * w0[:, None]).sum(0)) < 0.15)
    if mode.get("fp32_state"):
        assert np.array_equal(r.P, r.P.astype(np.float32).astype(np.float64))   # the particles ARE float rows


def test_fp32_state_details(A, gpu_ctx):
    """FP32 particle state (opts.fp32_state): discrete marginals and blobs survive the float rows, odd row strides (d = 1) and
    resampling work, and the mode is refused where the kernels keep FP64 rows (g-and-k, snapshots)."""
    spec, data = MODEL_CASES["normdu"]
    r = A.abcdesmc(to_prior(A, spec), A.Model("normdu", data), 0.05, None, nparticles=3000, verbose=False, rng=3, fp32_state=True)
    assert r.iters > 3 and np.isfinite(r.logZ) and np.array_equal(r.P, r.P.astype(np.float32).astype(np.float64))
    assert np.array_equal(r.P[:, 1], np.rint(r.P[:, 1])) and r.P[:, 1].min() >= 1 and r.P[:, 1].max() <= 10
    rb = A.abcdesmc(A.host.Normal(0, SQ10), A.Model("gauss1d_blob", [3.0, 1.0]), 0.3, None, nparticles=2000, verbose=False, rng=4, fp32_state=True)
    y = rb.blobs.view(np.float64)[:, 0]
    assert rb.stats["n_resamples"] >= 1 and np.allclose(np.abs(y - 3.0), rb.C, rtol=1e-12, atol=1e-12)   # blob = y, d = abs(y - 3): the blobs travelled with their particles
    with pytest.raises(A.ABCdeZError) as e:
        A.abcdesmc(A.Factored(*[A.host.Uniform(0.0, 10.0)] * 4), A.Model("gk", [1000.0] + _gk_octiles(GK_TRUE)), 1.0, None, nparticles=300,
                   verbose=False, rng=1, fp32_state=True)
    assert e.value.code == A.host.ERR_UNSUPPORTED
    with pytest.raises(A.ABCdeZError):
        A.abcdesmc(A.host.Normal(0, SQ10), A.Model("gauss1d", [3.0, 1.0]), 0.3, None, nparticles=1000, verbose=False, rng=1, fp32_state=True,
                   max_iters=3, return_state=True)
