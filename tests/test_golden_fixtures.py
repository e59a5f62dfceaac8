"""Committed fixtures of tests/golden/ (made by tests/golden/make_golden.py from the CPU oracle).

CPU: the oracle must still reproduce every vector bit for bit (it cannot drift silently).
GPU: the CUDA path through the C ABI must reproduce the same files: integers and decisions exactly, eps history
exactly, floating-point state within 1e-9 relative (observed: identical)."""
import glob
import json
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = sorted(glob.glob(os.path.join(HERE, "*.json")))
RUNS = [f for f in FILES if json.load(open(f))["kind"] in ("abcdesmc", "abcdemc")]


def unhex(v):
    return np.array([float.fromhex(x) for x in v])


def _spec(rec):
    return [tuple(p) for p in rec["prior"]]


def test_fixture_inventory():
    assert len(RUNS) >= 6 and os.path.exists(os.path.join(HERE, "stages.json"))


@pytest.mark.parametrize("path", RUNS, ids=[os.path.basename(f)[:-5] for f in RUNS])
def test_oracle_reproduces_fixture(oracle, path):
    rec = json.load(open(path))
    if rec["kind"] == "abcdesmc":
        r = oracle.smc_run(_spec(rec), rec["model"], rec["data"], rec["eps_target"], **rec["kwargs"])
        n = len(rec["eps_hist"])
        assert (r.iters, r.nsims) == (rec["iters"], rec["nsims"])
        assert r.eps == float.fromhex(rec["eps"]) and r.logZ == float.fromhex(rec["logZ"])
        assert np.array_equal(r.hist["eps"][:n], unhex(rec["eps_hist"]))
        assert [int(k) for k in r.hist["Kmcmc"][:n]] == rec["Kmcmc_hist"]
        assert int((r.Wns > 0).sum()) == rec["n_alive"]
        assert np.array_equal(r.P[:8].ravel(), unhex(rec["P_head"])) and np.array_equal(r.C[:8], unhex(rec["C_head"]))
    else:
        r = oracle.mc_run(_spec(rec), rec["model"], rec["data"], rec["eps_target"], **rec["kwargs"])
        assert r.nsims == rec["nsims"] and bool(r.reached_eps) == rec["reached_eps"]
        assert np.array_equal(r.P[:8].ravel(), unhex(rec["P_head"])) and np.array_equal(r.C[:8], unhex(rec["C_head"]))


def test_oracle_reproduces_stage_vectors(oracle):
    rec = json.load(open(os.path.join(HERE, "stages.json")))
    w, u = unhex(rec["weights"]), unhex(rec["uniforms"])
    assert [int(i) for i in oracle.wsample_stratified(w, u)] == rec["inds"]
    # independent restatement of wsample_stratified! (src/abcdez_smc.jl:15-56) in numpy: cumulative weights + search
    n = w.size
    edges = np.cumsum(np.full(n, 1.0 / n)) - 1.0 / n
    cs = np.cumsum(w)
    ref = np.searchsorted(cs, edges + u / n, side="left") + 1
    assert np.mean(np.minimum(ref, n) == np.array(rec["inds"])) > 0.99      # ties at rounding boundaries may differ
    dl, al = unhex(rec["delta"]), np.array(rec["alive"], dtype=np.uint8)
    for p, q in zip(rec["quantile_p"], unhex(rec["quantile"])):
        assert oracle.quantile_alive(dl, al, p)[0] == q
        assert np.isclose(q, np.quantile(dl[al > 0], p, method="linear"), rtol=1e-14)


FAM = {"normal": "Normal", "uniform": "Uniform", "discrete_uniform": "DiscreteUniform"}


@pytest.mark.gpu
@pytest.mark.parametrize("path", RUNS, ids=[os.path.basename(f)[:-5] for f in RUNS])
def test_cuda_reproduces_fixture(A, gpu_ctx, path):
    rec = json.load(open(path))
    kw = dict(rec["kwargs"])
    if kw.pop("islands", 1) != 1:
        pytest.skip("island fixtures are checked by the sharded runs (tests/multi_gpu_worker.py)")
    prior = A.Factored(*[getattr(A.host, FAM[p[0]])(*p[1:]) for p in _spec(rec)])
    model = A.Model(rec["model"], rec["data"])
    N = kw.pop("nparticles"); seed = kw.pop("seed")
    if rec["kind"] == "abcdesmc":
        r = A.abcdesmc(prior, model, rec["eps_target"], None, nparticles=N, rng=seed, verbose=False,
                       ABCk=kw.pop("kind", "indicator_strict"), **kw)
        n = len(rec["eps_hist"])
        assert (r.iters, r.nsims) == (rec["iters"], rec["nsims"])
        assert np.array_equal(r.eps_hist[:n], unhex(rec["eps_hist"])) and list(r.Kmcmcs[:n]) == rec["Kmcmc_hist"]
        assert int((r.Wns > 0).sum()) == rec["n_alive"]
        assert abs(r.logZ - float.fromhex(rec["logZ"])) <= 1e-9 * abs(float.fromhex(rec["logZ"]))
        np.testing.assert_allclose(np.asarray(r.P).reshape(N, -1)[:8].ravel(), unhex(rec["P_head"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r.C[:8], unhex(rec["C_head"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(float(np.sum(r.C)), float.fromhex(rec["C_sum"]), rtol=1e-9)
    else:
        r = A.abcdemc(prior, model, rec["eps_target"], None, nparticles=N, generations=kw.pop("generations"), rng=seed,
                      verbose=False)
        assert r.nsims == rec["nsims"] and bool(r.reached_eps) == rec["reached_eps"]
        np.testing.assert_allclose(np.asarray(r.P).reshape(N, -1)[:8].ravel(), unhex(rec["P_head"]), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(r.C[:8], unhex(rec["C_head"]), rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
def test_cuda_reproduces_stage_vectors(A, gpu_ctx):
    rec = json.load(open(os.path.join(HERE, "stages.json")))
    w, u = unhex(rec["weights"]), unhex(rec["uniforms"])
    got = A.wsample_stratified(w, u, mode=2)
    assert [int(i) for i in got] == rec["inds"]
