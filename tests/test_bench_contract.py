"""bench.py's reference arm runs without a GPU: check the JSON-line contract the driver parses (keys, units, the
cpu_baseline / e2e objects of the reference arm) on a tiny sample, and that the GPU arm refuses to run on the CPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600,
                          env={**os.environ, **(env or {})})


def test_reference_arm_prints_one_contract_line():
    res = _run("--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--cpu-sample-agents", "20000")
    assert res.returncode == 0, res.stderr
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "agent-steps/s" and line["higher_is_better"] is True
    assert line["metric"] == "agent-steps/s fwd+bwd, 8.4M-agent SEIRM" and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["vs_baseline"] is None
    # config.agents states the N the CPU arm really ran (a bounded sample), agents_gpu_arm the GPU arm's workload
    assert line["config"]["agents"] == 20_000 and line["config"]["agents_gpu_arm"] == 8_400_000
    assert "workload" in line["config"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    # "reference": the unmodified reference package is installed (baseline/_ref, or /root/reference in the build container)
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and "N=20000" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    res = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert res.returncode == 0 and res.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_gpu_arm_fails_loudly_without_a_gpu():
    res = _run("--steps", "1", "--warmup", "0", "--agents", "1000")
    assert res.returncode != 0 and "no CPU path" in res.stderr
