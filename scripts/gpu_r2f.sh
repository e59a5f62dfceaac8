#!/bin/bash
# round 2, call f (1 GPU): full parity suite with the new features, A/B of the relaxed-parity modes, batched-run timing
set -u
mkdir -p gpurun_out
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -30 ) > gpurun_out/r2f_pytest.log
tail -n 30 gpurun_out/r2f_pytest.log
for flags in "" "--segments" "--systematic" "--segments --systematic"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline $flags 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench [$flags]', round(d['value']/1e9,3), 'G evals/s', round(d['ms_per_step'],2), 'ms  sweep frac', round(d['roofline']['frac'],3), 'sweep us', round(1e3*d['roofline']['avg_launch_ms'],1), 'logZ', round(d['logZ_mean'],3))"
done > gpurun_out/r2f_modes.log 2>&1
cat gpurun_out/r2f_modes.log
timeout 300 python - > gpurun_out/r2f_batch.log 2>&1 <<'PY'
import time, math, numpy as np, sys
sys.path.insert(0, '.')
import abcdez_b200 as A
m = A.Model("gauss1d", [3.0, 1.0]); pr = A.host.Normal(0, math.sqrt(10))
for n in (1000,):
    A.abcdesmc(pr, m, 0.3, None, nparticles=n, verbose=False, rng=1)
    t0 = time.perf_counter(); lz = [A.abcdesmc(pr, m, 0.3, None, nparticles=n, verbose=False, rng=10 + k, verboseout=False).logZ for k in range(64)]; t1 = time.perf_counter() - t0
    runs = [dict(prior=pr, dist=m, eps_target=0.3, nparticles=n, rng=10 + k) for k in range(64)]
    A.abcdesmc_batch(runs[:16])
    t0 = time.perf_counter(); lb = [r.logZ for r in A.abcdesmc_batch(runs)]; t2 = time.perf_counter() - t0
    print(f"64 replicates of config 1 (N={n}): one after the other {t1*1e3:.1f} ms, batched {t2*1e3:.1f} ms, identical logZ: {lz == lb}, Z = {np.exp(np.mean(lb)):.5f} +- {np.std(np.exp(lb)):.5f}")
PY
cat gpurun_out/r2f_batch.log
