import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abcdez_b200 as A
from oracle import oracle as O
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
spec = [("normal", 0.0, math.sqrt(10))]
want = O.smc_run(spec, "gauss1d", [3.0, 1.0], 0.3, nparticles=N, seed=4242)
got = A.abcdesmc(A.host.Normal(0, math.sqrt(10)), A.Model("gauss1d", [3.0, 1.0]), 0.3, None, nparticles=N, rng=4242, verbose=False)
print("iters", got.iters, want.iters, "nsims", got.nsims, want.nsims)
n = min(len(got.eps_hist), len(want.hist["eps"]))
for i in range(n):
    row = (got.eps_hist[i], want.hist["eps"][i], got.esss[i], want.hist["ess"][i], got.faccs[i], want.hist["facc"][i],
           got.Kmcmcs[i], want.hist["Kmcmc"][i], got.logZs[i], want.hist["logZ"][i], got.ranges_eps[i, 0], want.hist["dmin"][i])
    bad = not (math.isclose(row[0], row[1], rel_tol=1e-9) and math.isclose(row[2], row[3], rel_tol=1e-9) and math.isclose(row[4], row[5], rel_tol=1e-9))
    if bad or i < 3:
        print(i, "BAD" if bad else "ok", row)
    if bad:
        break
