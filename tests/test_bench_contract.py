"""bench.py contract pieces that run without a GPU: the reference arm (CPU restatement of the reference path) prints
one JSON line with the keys the driver reads; the default arm refuses to run without a device."""
import json
import os
import subprocess
import sys

import helpers as H


def _run(*argv, env=None):
    return subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), *argv], capture_output=True, text=True,
                          cwd=H.ROOT, timeout=600, env=env)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--ref-sample", "300", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("reads/sec") and d["unit"] == "reads/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "ont037" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    # the unmodified reference CLI when baseline/_ref is present (build container, GPU box), else the oracle port
    ref_copy = os.path.isdir(os.path.join(H.ROOT, "baseline", "_ref", "src", "specimux"))
    assert cb["kind"] == ("reference" if ref_copy else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and "300" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = _run("--impl", "reference", "--gpus", "2", "--ref-sample", "300", "--steps", "1", "--warmup", "0", env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_default_arm_needs_a_gpu():
    from specimux_b200 import _lib
    if _lib.load().smx_device_count() > 0:
        import pytest
        pytest.skip("box has a GPU")
    r = _run("--steps", "1", "--warmup", "0", "--reads", "256")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "NO_DEVICE" in (r.stderr + r.stdout) or "error 3" in (r.stderr + r.stdout)
