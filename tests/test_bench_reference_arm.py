"""The reference arm of bench.py on the CPU (no GPU needed): `bench.py --impl reference` must print ONE JSON line with the
contract's keys, time the UNMODIFIED reference from baseline/_ref (`cpu_baseline.kind == "reference"`) and carry the same
`config` object as the GPU arm would for that workload."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT

sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_env  # noqa: E402


@pytest.mark.skipif(not ref_env.installed(), reason="baseline/_ref not built (no reference tree at build time)")
def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "0", "--steps", "1",
                        "--warmup", "0", "--gpus", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines  # stdout is reserved for the one JSON line
    j = json.loads(lines[0])
    import bench

    w, name = bench.workload(0)
    assert j["impl"] == "reference" and j["metric"] == "images/sec" and j["unit"] == "images/s" and j["higher_is_better"]
    assert j["value"] > 0 and j["steps"] == 1 and j["warmup"] == 0 and j["n_gpus"] == 1
    assert j["config"] == bench.config_obj(name, w, 1)
    cb = j["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == j["value"] and cb["cores"] >= 1 and "images per step" in cb["sample"]
    assert cb["as_shipped"]["threads"] <= 8  # the reference's own cap: min(8, ncpu - 1)
    assert j["e2e"] == {"value": j["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""
