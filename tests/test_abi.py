"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/abcdez_cuda.h declares, and fails loudly (no CPU fallback) when no GPU is present.
No compute call is made here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "abcdez_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abcdez_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(A):
    L = A.lib()
    declared = _declared_symbols()
    assert len(declared) >= 35
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/abcdez_cuda.h but not exported"
    assert sorted(A.EXPORTS) == declared, "host.EXPORTS out of sync with the header"


def test_version_and_registry(A):
    L = A.lib()
    assert L.abcdez_version() == 100
    names = A.model_names()
    for n in ("gauss1d", "gauss_corr10", "lotka_volterra", "birth_death", "wiener", "socks"):
        assert n in names
    m = A.Model("gauss_corr10", [0.0] * 11)
    assert (m.d, m.blob_bytes) == (10, 0)
    assert A.Model("birth_death", []).blob_bytes == 16
    with pytest.raises(A.ABCdeZError):
        A.Model("no_such_model")


def test_opts_defaults_match_reference(A):
    """kwargs defaults of src/abcdez_smc.jl:215-220 and src/abcdez_mc.jl:102-104."""
    o = A.host._SmcOpts()
    A.lib().abcdez_smc_opts_default(ctypes.byref(o))
    assert (o.nparticles, o.alpha, o.delta_ess, o.nsims_max, o.Kmcmc, o.Kmcmc_min) == (100, 0.95, 0.5, 10**7, 3, 1.0)
    assert (o.kernel, o.facc_stop, o.facc_min, o.facc_tune) == (1, 0.0, 0.0, 0.975)
    m = A.host._McOpts()
    A.lib().abcdez_mc_opts_default(ctypes.byref(m))
    assert (m.nparticles, m.generations) == (50, 20)


def test_kernel_host_functions(A):
    """The kernel known answers of test/runtests.jl:48-108 through the C ABI (pure host code)."""
    from test_oracle_golden import KERNEL_CASES
    cls = {"indicator": A.host.Indicator0toEps, "indicator_strict": A.host.IndicatorStrict0toEps,
           "epa": A.host.Epa0toEps, "epa_strict": A.host.EpaStrict0toEps}
    for kind, per_eps in KERNEL_CASES.items():
        for eps, cases in per_eps.items():
            k = cls[kind](eps)
            assert k.ϵ == eps
            for x, pdf, logpdf in cases:
                assert (k.pdf(x), k.logpdf(x)) == (pdf, logpdf), (kind, eps, x)
    with pytest.raises(A.ABCdeZError, match="Expected ϵ ≥ 0.0"):       # src/abcdez_types.jl:30
        A.host.Indicator0toEps(-1.0)


def test_no_cpu_fallback(A):
    """Without a GPU the product must fail loudly; with one this test is vacuous."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(A.ABCdeZError) as e:
        A.Context(0)
    assert e.value.code == A.host.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "abcdez.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liborc" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_post_run_helpers_host_side(A):
    """model probabilities (examples/minimal_example.jl:58-65) and the evidence-ladder lookup
    (docs/src/index.md:282-284) are plain host arithmetic: no GPU needed."""
    import math
    import numpy as np
    p = A.host.model_probabilities([math.log(0.047940112540007955), math.log(0.022780)])
    assert abs(p[0] - 0.678) < 1e-3 and abs(p.sum() - 1.0) < 1e-15
    p = A.host.model_probabilities([-3.0, -3.0, -3.0], prior_probs=[0.5, 0.25, 0.25])
    assert np.allclose(p, [0.5, 0.25, 0.25])
    mk = lambda eps, lz, eh, lzs: A.host.SMCResult(P=np.zeros(1), Wns=np.ones(1), C=np.zeros(1), eps=eps, logZ=lz, blobs=np.zeros((1, 0)),
                                                   eps_hist=np.array(eh), logZs=np.array(lzs))
    r1 = mk(0.3, -3.0, [np.inf, 1.0, 0.6, 0.45, 0.3], [0.0, -1.0, -2.0, -2.5, -3.0])
    r2 = mk(0.5, -2.8, [np.inf, 2.0, 0.9, 0.5], [0.0, -0.5, -1.7, -2.8])
    assert A.host.evidence_at(r1, 0.5) == -2.0 and A.host.evidence_at(r1, 0.3) == -3.0 and A.host.evidence_at(r1, 5.0) == 0.0
    p = A.host.model_probabilities([r1, r2])            # compared at eps = 0.5: -2.0 against -2.8
    assert abs(p[0] - 1.0 / (1.0 + math.exp(-0.8))) < 1e-12
    with pytest.raises(A.ABCdeZError):
        A.host.weightinds([0.5, 0.2])                   # weights must sum to 1 (test/runtests.jl:14)
