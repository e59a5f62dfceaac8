"""Generates the committed fixtures of tests/golden/ from the CPU oracle (oracle/abcdez_oracle.c).

The reference is Julia and cannot run in this image, so these vectors do NOT come from
ABCdeZ.jl itself: they freeze the oracle's outputs (which are pinned to the reference's own known answers by
tests/test_oracle_golden.py) so that (a) the oracle cannot drift silently -- test_golden_fixtures.py re-derives
every vector on CPU -- and (b) the CUDA path is compared with committed files as well as with the live oracle.

    python tests/golden/make_golden.py          # rewrites tests/golden/*.json
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402

NORMAL10 = [("normal", 0.0, math.sqrt(10.0))]

RUNS = {
    # name: (prior spec, model, data, eps_target, kwargs)
    "smc_gauss1d_n1000_seed2024": (NORMAL10, "gauss1d", [3.0, 1.0], 0.3, dict(nparticles=1000, seed=2024)),
    "smc_gauss1d_islands4_n2000_seed9": (NORMAL10, "gauss1d", [3.0, 1.0], 0.3, dict(nparticles=2000, seed=9, islands=4)),
    "smc_gauss_corr10_n4096_seed7": ([("normal", 0.0, 2.0)] * 10, "gauss_corr10", list(np.linspace(-1, 1, 10)) + [0.5], 3.0,
                                     dict(nparticles=4096, seed=7)),
    "smc_twod_indicator_n3000_seed11": ([("normal", 0.0, 5.0)] * 2, "twod", [], 0.1, dict(nparticles=3000, seed=11, kind="indicator")),
    "smc_normdu_n2000_seed5": ([("normal", 1.0, 0.5), ("discrete_uniform", 1, 10)], "normdu", [5.5], 0.05, dict(nparticles=2000, seed=5)),
}
MC_RUNS = {
    "mc_gauss1d_n500_seed3": (NORMAL10, "gauss1d", [3.0, 1.0], 0.5, dict(nparticles=500, generations=10, seed=3)),
}


def hexf(a):
    """bit-exact, JSON-safe encoding of float64 arrays"""
    return [float(x).hex() for x in np.asarray(a, dtype=np.float64).ravel()]


def main():
    O.build()
    for name, (spec, model, data, eps_t, kw) in RUNS.items():
        r = O.smc_run(spec, model, data, eps_t, **kw)
        n = r.hist_len if hasattr(r, "hist_len") else len(r.hist["eps"])
        rec = dict(kind="abcdesmc", prior=spec, model=model, data=list(map(float, data)), eps_target=eps_t, kwargs=kw,
                   iters=int(r.iters), nsims=int(r.nsims), eps=float(r.eps).hex(), logZ=float(r.logZ).hex(),
                   eps_hist=hexf(r.hist["eps"][:n]), Kmcmc_hist=[int(k) for k in r.hist["Kmcmc"][:n]],
                   n_alive=int((r.Wns > 0).sum()),
                   P_head=hexf(r.P[:8]), C_head=hexf(r.C[:8]), P_sum=float(np.sum(r.P)).hex(), C_sum=float(np.sum(r.C)).hex())
        json.dump(rec, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
    for name, (spec, model, data, eps_t, kw) in MC_RUNS.items():
        r = O.mc_run(spec, model, data, eps_t, **kw)
        rec = dict(kind="abcdemc", prior=spec, model=model, data=list(map(float, data)), eps_target=eps_t, kwargs=kw,
                   nsims=int(r.nsims), reached_eps=bool(r.reached_eps), P_head=hexf(r.P[:8]), C_head=hexf(r.C[:8]),
                   C_sum=float(np.sum(r.C)).hex())
        json.dump(rec, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
    # stage-level vectors: stratified resampling indices and the type-7 quantile on fixed inputs
    rng = np.random.default_rng(12345)
    w = rng.random(257); w[rng.random(257) < 0.3] = 0.0; w /= w.sum()
    u = rng.random(257)
    inds = O.wsample_stratified(w, u)
    dl = rng.gamma(2.0, 1.0, 1001); al = (rng.random(1001) < 0.8).astype(np.uint8)
    q = [O.quantile_alive(dl, al, p)[0] for p in (0.0, 0.5, 0.95, 0.999)]
    json.dump(dict(kind="stages", weights=hexf(w), uniforms=hexf(u), inds=[int(i) for i in inds],
                   delta=hexf(dl), alive=[int(a) for a in al], quantile_p=[0.0, 0.5, 0.95, 0.999], quantile=hexf(q)),
              open(os.path.join(HERE, "stages.json"), "w"), indent=1)
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".json")))


if __name__ == "__main__":
    main()
