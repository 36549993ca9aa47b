"""Host logic of the f16x3 path, no GPU needed: tools/v2_plan_check.cu builds the 18 GEMM launches of one encoder-decoder
step the way csrc/urnn_v2.cu does, runs v2::plan_gemm (csrc/v2_host.cuh) on them and verifies ring depth, TMEM layout,
shared-memory budget and every precomputed MMA descriptor (start addresses inside the ring / weight image, one overwriting
MMA per accumulator, K covered exactly once)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    path = str(tmp_path_factory.mktemp("v2plan") / "v2_plan_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-o", path,
                    os.path.join(ROOT, "tools", "v2_plan_check.cu")], check=True, capture_output=True)
    return path


def run(exe, *args):
    r = subprocess.run([exe, *[str(a) for a in args]], capture_output=True, text=True)
    rows = {}
    for line in r.stdout.strip().splitlines():
        kv = {k: int(v) for k, v in re.findall(r"(\w+)=(-?\d+)", line)}
        if kv:
            rows[" ".join(line.split()[:2]) if line.startswith("chain") else line.split()[0]] = kv
    return r, rows


@pytest.mark.parametrize("grid", [(500, 500), (32, 32), (128, 128), (4096, 4096), (512, 4096)])
def test_step_launch_plans(exe, grid):
    r, rows = run(exe, *grid)
    assert r.returncode == 0 and "PLAN CHECK PASSED" in r.stdout, r.stdout + r.stderr
    launches = {k: v for k, v in rows.items() if not k.startswith("chain")}
    assert len(launches) == 18
    for name, p in launches.items():
        assert p["smem"] <= p["cap"] <= 227 * 1024, name
        assert p["stages"] * p["nacc"] * p["stride"] <= p["tmem"] <= 512, name
    tiles = 16 * (((grid[0] // 4) * (grid[1] // 4) + 127) // 128)      # 16 phase blocks of n4p pixels, 128 pixels per tile
    assert launches["dec1.A"]["grid"] == min(148, tiles)


def test_location1_plan_details(exe):
    _, rows = run(exe, 500, 500)
    assert rows["dec2.A1"]["N"] == 96 and "dec2.A2" in rows          # 2F = 192 rows x K = 288 hi+lo do not fit: two launches of F rows
    assert rows["enc1.A"]["stages"] == 4 and rows["enc3.A"]["stages"] == 2      # N = 128 -> four accumulator stages, N = 192 -> two
    assert rows["stem3"]["nacc"] == 4 and rows["stem3"]["stages"] == 1           # four phase accumulators of 96 columns fill TMEM
    assert rows["enc1.B"]["gdepth"] == 2 and rows["dec2.B"]["gdepth"] == 1        # double-buffered gate operands only beside a deep ring
    assert rows["enc3.A"]["grid"] == 123                                          # quarter-resolution maps: fewer tiles than SMs
    # recompute schedule: only the full-resolution encoder cell fits [W1 ; W2] hi+lo beside its buffers
    assert rows["chain enc1"]["fits"] == 1
    assert all(rows[f"chain {c}"]["fits"] == 0 for c in ("enc2", "enc3", "dec3", "dec2", "dec1"))
