"""Micro-benchmark of the fused sweep kernel on a resident population (not a bench.py value)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abcdez_b200 as A

model_name = sys.argv[1] if len(sys.argv) > 1 else "gauss_corr10"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
dead = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
cases = {
    "gauss_corr10": (A.Factored(*[A.host.Normal(0.0, 2.0)] * 10), [0.5 * math.sin(1.0 + k) for k in range(10)] + [0.5], 353),
    "gauss1d": (A.Factored(A.host.Normal(0.0, math.sqrt(10))), [3.0, 1.0], 65),
    "twod": (A.Factored(A.host.Normal(0, 5), A.host.Normal(0, 5)), [], 97),
    "lotka_volterra": (A.Factored(*[A.host.Uniform(0.0, 2.0)] * 4), [1.0, 0.5, 0.01, 50, 8, 0.05] + list(np.tile([1.2, 0.6], 8)), 161),
    "birth_death": (A.Factored(*[A.host.Uniform(0.0, 2.0)] * 2), [20.0, 8, 0.5, 5000.0] + [22, 25, 24, 30, 33, 31, 36, 40], 113),
    "gk": (A.Factored(*[A.host.Uniform(0.0, 10.0)] * 3, A.host.Uniform(0.0, 2.0)),
           [10000.0, 1.86, 2.39, 2.73, 3.0, 3.67, 4.86, 7.9], 161),
    "gk_f32": (A.Factored(*[A.host.Uniform(0.0, 10.0)] * 3, A.host.Uniform(0.0, 2.0)),
               [10000.0, 1.86, 2.39, 2.73, 3.0, 3.67, 4.86, 7.9], 161),
}
prior, data, bytes_per = cases[model_name]
pop = A.Population(prior, A.Model(model_name, data), N)
pop.init(seed=1)
st = pop.download()
eps = float(np.quantile(st["delta"], 0.5))
if dead > 0:
    alive = (np.random.default_rng(0).random(N) >= dead).astype(np.uint8)
    pop.upload(alive=alive, W=np.where(alive > 0, 1.0 / alive.sum(), 0.0))
pop.set(eps=eps, kernel="indicator_strict", seed=3)
pop.bench_sweeps(3)
ns, na, ms = pop.bench_sweeps(20)
per = ms / 20
print(f"{os.environ.get('ABCDEZ_LIB', 'default')}: {model_name} N={N} dead={dead}: {per*1e3:.1f} us/sweep, {ns/20/per*1e-6:.1f} M evals/ms-> "
      f"{ns/(ms*1e-3):.3e} evals/s, acc {na/max(ns,1):.3f}, algorithmic {bytes_per*ns/(ms*1e-3)/1e9:.0f} GB/s")
