"""The driver's contract for `bench.py --impl reference`, checked on CPU with the small configuration (C1).

The reference arm is one of the two places outside tests/ that may execute oracle/ (the other is the cpu_baseline leg);
it needs no GPU, so its JSON line and its behaviour under torchrun (rank 0 alone works and prints) are checked here.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _check_line(line, n_gpus, steps, warmup):
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference"
    assert line["metric"] == "rlap_views_per_sec" and line["unit"] == "views/s"
    assert (line["n_gpus"], line["steps"], line["warmup"]) == (n_gpus, steps, warmup)
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("C1")
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert line["gpu_launches"] == 0


def _json_lines(text):
    return [json.loads(l) for l in text.splitlines() if l.startswith("{")]


@pytest.mark.timeout(300)
def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--config", "C1", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1, r.stdout
    _check_line(lines[0], 1, 2, 1)


@pytest.mark.timeout(400)
def test_reference_arm_under_torchrun_only_rank0_works():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", BENCH, "--impl", "reference",
                        "--config", "C1", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=380)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = _json_lines(r.stdout)
    assert len(lines) == 1, r.stdout
    _check_line(lines[0], 2, 1, 1)


def test_default_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm runs")
    r = subprocess.run([sys.executable, BENCH, "--config", "C1", "--steps", "1", "--warmup", "1", "--no-cpu-baseline"],
                       capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode != 0
    assert not _json_lines(r.stdout), "no bench line may be printed without the CUDA path"
