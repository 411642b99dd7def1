"""bench.py's contract on a GPU-less host: the reference arm prints ONE JSON line with the keys the driver reads, on the
same metric / unit / config as the GPU arm, and a non-zero rank under torchrun prints nothing and exits 0."""

import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(env_extra, *args):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)


def test_reference_arm_prints_one_contract_line():
    res = _run({}, "--impl", "reference", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout[-2000:]
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "streams_x_frames_per_sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "streams" in cb["sample"]
    assert cb["cores"] == cb["host"]["threads_used"] <= cb["host"]["affinity"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "configs[1]" in d["config"]["workload"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    res = _run({"RANK": "3", "WORLD_SIZE": "8", "LOCAL_RANK": "3"}, "--impl", "reference", "--gpus", "8", "--steps", "1", "--warmup", "0")
    assert res.returncode == 0 and res.stdout.strip() == ""
