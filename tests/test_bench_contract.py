"""CPU: the JSON line of `bench.py --impl reference` (the only arm that runs without a GPU) carries every key of the
measurement contract, for the default workload and for the others; and the B200 arm refuses to run without a device
instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "cpu_baseline", "e2e", "impl"}


def run(args, timeout=600):
    return subprocess.run([sys.executable, "bench.py"] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("workload", [None, "cylinder_bb", "poiseuille", "kbc_shear", "csf_rt"])
def test_reference_arm_line(workload):
    args = ["--impl", "reference", "--steps", "2", "--warmup", "1", "--cpu-sample", "64"]
    if workload:
        args += ["--workload", workload]
    r = run(args)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "MLUPS" and d["unit"] == "MLUPS" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_b200_arm_needs_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = run(["--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0
    assert "no CUDA device" in (r.stdout + r.stderr)
