"""bench.py on a box without a GPU: the reference arm (CPU oracle port) prints the contract's JSON line, the product arm
refuses to run (no CPU fallback), and under torchrun only rank 0 of the reference arm works."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=300):
    e = dict(os.environ, CS_BENCH_CPU_MIN_S="0.3")
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          cwd=ROOT, timeout=timeout)


@pytest.mark.parametrize("workload,metric,unit", [("cfg2", "scan-point map lookups/sec", "lookups/s"),
                                                  ("cfg3", "HoleMap cell visits/sec", "visits/s")])
def test_reference_arm_line(workload, metric, unit):
    r = _run(["--impl", "reference", "--workload", workload, "--steps", "2", "--warmup", "3"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == unit
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["config"]["workload"].startswith(workload)
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a CUDA device")
@pytest.mark.parametrize("workload", ["cfg2", "cfg3", "cfg5"])
def test_product_arm_refuses_to_run_without_a_gpu(workload):
    r = _run(["--workload", workload, "--steps", "2", "--warmup", "3"])
    assert r.returncode != 0
    assert "CUDA device" in (r.stderr + r.stdout)
    assert not any(l.startswith("{") for l in r.stdout.splitlines())  # no JSON line: nothing was measured
