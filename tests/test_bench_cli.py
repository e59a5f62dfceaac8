"""bench.py on a machine without a GPU: the reference arm (the CPU port of the oracle) prints the contract's JSON line,
and the product arm refuses to run -- there is no CPU fallback behind the GPU numbers."""
import json, os, subprocess, sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    p = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--particles", "20000")
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "evals/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("BASELINE.json configs[1]") and line["dtype"] == "f64"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a machine WITHOUT a GPU")
def test_product_arm_has_no_cpu_fallback():
    p = _bench("--steps", "1", "--warmup", "0", timeout=300)
    assert p.returncode != 0
    assert "no CUDA device" in (p.stdout + p.stderr)
    assert not any(l.startswith("{") for l in p.stdout.splitlines())      # no number without the GPU


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    """N > 1: the driver launches the reference arm like the product arm; rank 0 alone runs it on the SAME total
    population with every host thread (torchrun's OMP_NUM_THREADS=1 is overridden), the other ranks exit 0."""
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29657", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--particles", "10000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["config"]["particles"] == 20000
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1) or line["cpu_baseline"]["cores"] > 1 or (os.cpu_count() or 1) == 1
