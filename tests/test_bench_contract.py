"""bench.py contract pieces that run without a GPU: the reference arm (the unmodified reference modules from oracle/_ref,
or the op-for-op torch port where that copy is absent) prints one JSON line with the keys the driver reads; ranks other
than 0 print nothing and exit 0; both arms name the workload identically."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--height", "32", "--width", "32", "--hist", "3"], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_line():
    r = _run({})
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "cells*steps/s"
    assert d["metric"].startswith("grid-cells*steps/s")
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "src", "lib", "model", "networks", "model.py"))
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"]["workload"] == bench.workload_name(32, 32, 9)       # the string our arm emits for the same grid


def test_reference_arm_other_ranks_are_silent():
    r = _run({"WORLD_SIZE": "2", "RANK": "1", "LOCAL_RANK": "1"})
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ""
