"""Where does a full abcdesmc! run spend its time? (development aid)"""
import ctypes as C, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import abcdez_b200 as A
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = A.default_context(); L = A.lib()
prior = A.Factored(*[A.host.Normal(0.0, 2.0)] * 10)
model = A.Model("gauss_corr10", [0.5 * math.sin(1.0 + k) for k in range(10)] + [0.5])
def run(seed, profile, sync_every, eps=1.0):
    o = A.host._SmcOpts(); L.abcdez_smc_opts_default(C.byref(o))
    o.nparticles = N; o.nsims_max = 10**15; o.seed = seed; o.verboseout = 0; o.profile = profile; o.sync_every = sync_every
    r = A.host._SmcResult()
    t0 = time.perf_counter()
    rc = L.abcdez_smc_run(ctx._h, prior.handle(ctx), model.handle(ctx), C.c_double(eps), C.byref(o), C.byref(r))
    dt = (time.perf_counter() - t0) * 1e3
    assert rc == 0, L.abcdez_last_error()
    return r, dt
run(1, 0, 4)
for profile, se in [(0, 1), (0, 4), (0, 16), (0, 64), (1, 4)]:
    r, dt = run(2, profile, se)
    print(f"profile={profile} sync_every={se}: wall {dt:.1f} ms, loop {r.total_ms:.1f} ms, init {r.init_ms:.2f} ms, iters {r.iters}, "
          f"per-iter {r.total_ms/r.iters*1e3:.0f} us, sweeps {r.n_sweeps}, sweep_ms {r.sweep_ms:.1f}, launches {r.n_launches}, resamples {r.n_resamples}, nsims {r.nsims}")
