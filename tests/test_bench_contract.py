"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`, the reference's own CPU encoder on the host cores) prints exactly one JSON
line with the keys the driver reads; the GPU arm refuses to run without a CUDA device (no CPU
fallback) and leaves stdout empty."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=env,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
             {"FLAKE_BENCH_REF_SAMPLES_PER_THREAD": str(4096 * 40), "FLAKE_BENCH_SKIP_SINGLE_THREAD": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    assert d["metric"].startswith("MSamples/s encoded, flake -8") and d["unit"] == "MSamples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("C2:")
    cb = d["cpu_baseline"]
    # every host core, one corpus track each: the shape the GPU arm's e2e leg is measured on
    assert cb["kind"] in ("reference", "port") and cb["cores"] == (os.cpu_count() or 1)
    assert cb["value"] == d["value"] and cb["sample"] and "e2e_workload" in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], {"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a CUDA device")
    r = _run(["--steps", "1", "--warmup", "1"], {})
    assert r.returncode != 0
    assert r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr
