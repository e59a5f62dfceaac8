"""Runtime-supplied models (abcdez_model_compile, csrc/rtc.cu): the `dist!` plugin as CUDA source compiled with
NVRTC against the library's own kernel templates.

CPU: sources compile for sm_100a without a GPU; NVRTC diagnostics (with the caller's line numbers) and the D / BLOB
checks come back as errors.  GPU: a runtime-compiled copy of a built-in model reproduces the built-in (and hence
the oracle) bit for bit through init, sweeps and whole runs; a model that exists nowhere else recovers its
parameters."""
import math

import numpy as np
import pytest

GAUSS_BLOB = r'''
struct UserGaussBlob {                     // == gauss1d_blob (models.cuh): y ~ N(theta, sigma), d = |y - data|, blob = y
    static constexpr int D = 1, BLOB = 8, NOISE = 0;
    static constexpr const char* name = "user_gauss_blob";
    __device__ static __forceinline__ double run(const double* th, const double* data, SimRng& r, double* blob)
    {
        double y = th[0] + data[1] * r.n();
        blob[0] = y;
        return fabs(y - data[0]);
    }
};
'''

TWOD = r'''
struct UserTwoD {                          // == twod (test/runtests.jl:603)
    static constexpr int D = 2, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "user_twod";
    __device__ static __forceinline__ double run(const double* th, const double*, SimRng& r, double*)
    {
        double z1, z2;
        r.n2(z1, z2);
        double t1 = th[0] + z1 * 0.01 - th[1] * th[1];
        double t2 = th[1] - 1.0 + z2 * 0.01;
        return 50.0 * (t1 * t1) + t2 * t2;
    }
};
'''

DECAY = r'''
// a model that exists nowhere else: noisy exponential decay y_t = a exp(-b t) + 0.05 z_t at t = 0..9, RMS distance
double decay_mean(double a, double b, int t) { return a * pexp(-b * (double)t); }      // helper: -default-device
struct UserDecay {
    static constexpr int D = 2, BLOB = 0, NOISE = 0;
    static constexpr const char* name = "user_decay";
    __device__ static double run(const double* th, const double* data, SimRng& r, double*)
    {
        double acc = 0.0;
        for (int t = 0; t < 10; t += 2) {
            double z1, z2;
            r.n2(z1, z2);
            double e1 = decay_mean(th[0], th[1], t) + 0.05 * z1 - data[t];
            double e2 = decay_mean(th[0], th[1], t + 1) + 0.05 * z2 - data[t + 1];
            acc += e1 * e1 + e2 * e2;
        }
        return sqrt(acc / 10.0);
    }
};
'''


def test_compile_only_without_gpu(A):
    for name, struct, src, d, blob in [("user_gauss_blob", "UserGaussBlob", GAUSS_BLOB, 1, 8), ("user_twod", "UserTwoD", TWOD, 2, 0),
                                       ("user_decay", "UserDecay", DECAY, 2, 0)]:
        A.compile_model(name, struct, src, d, blob, load=False)


def test_compile_errors_are_reported(A):
    with pytest.raises(A.ABCdeZError) as e:
        A.compile_model("x", "UserTwoD", TWOD.replace("r.n2(z1, z2);", "r.n2(z1, z3);"), 2, 0, load=False)
    assert "model.cu(" in str(e.value) and "z3" in str(e.value)
    with pytest.raises(A.ABCdeZError) as e:
        A.compile_model("x", "UserTwoD", TWOD, 3, 0, load=False)             # D mismatch
    assert "D differs" in str(e.value)
    with pytest.raises(A.ABCdeZError) as e:
        A.compile_model("x", "UserGaussBlob", GAUSS_BLOB, 1, 0, load=False)  # BLOB mismatch
    assert "BLOB differs" in str(e.value)
    with pytest.raises(A.ABCdeZError):
        A.compile_model("x", "NoSuchStruct", TWOD, 2, 0, load=False)
    with pytest.raises(A.ABCdeZError):
        A.compile_model("x", "UserTwoD", TWOD, 0, 0, load=False)             # argument check before NVRTC


@pytest.fixture(scope="module")
def user_models(A, gpu_ctx):
    have = set(A.model_names())
    for name, struct, src, d, blob in [("user_gauss_blob", "UserGaussBlob", GAUSS_BLOB, 1, 8), ("user_twod", "UserTwoD", TWOD, 2, 0),
                                       ("user_decay", "UserDecay", DECAY, 2, 0)]:
        if name not in have:
            A.compile_model(name, struct, src, d, blob, ctx=gpu_ctx)
    return True


@pytest.mark.gpu
def test_runtime_model_is_registered(A, gpu_ctx, user_models):
    assert {"user_gauss_blob", "user_twod", "user_decay"} <= set(A.model_names())
    m = A.Model("user_gauss_blob", [3.0, 1.0])
    assert (m.d, m.blob_bytes) == (1, 8)
    with pytest.raises(A.ABCdeZError):
        A.compile_model("user_twod", "UserTwoD", TWOD, 2, 0, ctx=gpu_ctx)     # duplicate name


@pytest.mark.gpu
@pytest.mark.parametrize("user,builtin,spec,data,eps", [
    ("user_gauss_blob", "gauss1d_blob", [("normal", 0.0, math.sqrt(10.0))], [3.0, 1.0], 0.3),
    ("user_twod", "twod", [("normal", 0.0, 5.0)] * 2, [], 0.05)])
def test_runtime_copy_equals_builtin_and_oracle(A, oracle, gpu_ctx, user_models, user, builtin, spec, data, eps):
    prior = A.Factored(*[A.host.Normal(*s[1:]) for s in spec])
    th = oracle.push(spec, oracle.prior_sample(spec, 2000, seed=5))
    d_user, b_user = A.Model(user, data).simulate(th, seed=77, epoch=4)
    d_built, b_built = A.Model(builtin, data).simulate(th, seed=77, epoch=4)
    assert np.array_equal(d_user, d_built) and np.array_equal(b_user, b_built)
    kw = dict(nparticles=3000, rng=11, verbose=False, nsims_max=10**8)
    ru = A.abcdesmc(prior, A.Model(user, data), eps, None, **kw)
    rb = A.abcdesmc(prior, A.Model(builtin, data), eps, None, **kw)
    want = oracle.smc_run(spec, builtin, data, eps, nparticles=3000, seed=11, nsims_max=10**8)
    assert (ru.iters, ru.nsims) == (rb.iters, rb.nsims) == (want.iters, want.nsims)
    assert ru.logZ == rb.logZ and np.array_equal(ru.P, rb.P) and np.array_equal(ru.C, rb.C) and np.array_equal(ru.blobs, rb.blobs)
    assert abs(ru.logZ - want.logZ) <= 1e-9 * abs(want.logZ)
    mu = A.abcdemc(prior, A.Model(user, data), eps * 3, None, nparticles=500, generations=6, rng=3, verbose=False)
    mb = A.abcdemc(prior, A.Model(builtin, data), eps * 3, None, nparticles=500, generations=6, rng=3, verbose=False)
    assert mu.nsims == mb.nsims and np.array_equal(mu.P, mb.P) and np.array_equal(mu.C, mb.C)


@pytest.mark.gpu
def test_new_runtime_model_recovers_parameters(A, gpu_ctx, user_models):
    a, b = 2.0, 0.3
    data = [a * math.exp(-b * t) for t in range(10)]
    prior = A.Factored(A.host.Uniform(0.0, 5.0), A.host.Uniform(0.0, 2.0))
    r = A.abcdesmc(prior, A.Model("user_decay", data), 0.08, None, nparticles=5000, rng=4, verbose=False, nsims_max=10**7)
    assert r.eps <= 0.1
    w = r.Wns / r.Wns.sum()
    mean = (r.P * w[:, None]).sum(0)
    assert abs(mean[0] - a) < 0.1 and abs(mean[1] - b) < 0.05
