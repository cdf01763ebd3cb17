"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the contract's keys (it times the
oracle port on the host cores), and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_contract_line():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                                   "--warmup", "0"], timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    line = json.loads(out.decode().strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["unit"] == "grad-evals/s"
    assert line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    assert line["config"]["workload"] == "case3c_65536"
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "chains" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                                   "--steps", "1", "--warmup", "0"], timeout=120,
                                  env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert out.decode().strip() == ""


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu"],
                       capture_output=True, timeout=300)
    assert r.returncode != 0 and b"CUDA" in r.stderr
