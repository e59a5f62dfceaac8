"""Micro-benchmark of the iteration head (eps select + reweight + ESS + alive list) on a resident population:
the fused cooperative kernel vs the stage kernels, and the resampling gather (not a bench.py value)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abcdez_b200 as A

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
prior = A.Factored(*[A.host.Normal(0.0, 2.0)] * 10)
data = [0.5 * math.sin(1.0 + k) for k in range(10)] + [0.5]
pop = A.Population(prior, A.Model("gauss_corr10", data), N)
pop.init(seed=1)
for rep in range(3):
    # alpha < 1 kills 5 % per call: successive calls see 100 %, 95 %, 90 % ... alive
    t_head = []
    for it in range(6):
        pop.head(0.95, 0.0)
        t_head.append(pop.last_timing()[0] * 1e3)
    pop.resample(epoch=rep + 1)
    t_res = pop.last_timing()[0] * 1e3
    print(f"N={N} rep {rep}: fused head us/call {[round(t, 1) for t in t_head]}  resample {t_res:.1f} us")
t_q = []; t_r = []
for it in range(4):
    q = pop.eps_quantile(0.95)[0]; t_q.append(pop.last_timing()[0] * 1e3)
    pop.reweight(q); t_r.append(pop.last_timing()[0] * 1e3)
print(f"N={N}: stage kernels: eps_quantile (7 launches) {[round(t, 1) for t in t_q]} us, reweight+compact (3 launches) {[round(t, 1) for t in t_r]} us")
