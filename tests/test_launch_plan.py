"""Host logic of the tcgen05 path, no GPU needed: tools/plan_check.cu runs tc::plan_launch (csrc/tc_pixgemm.cuh) over the
18 GEMM shapes of the location1 encoder-decoder step and prints the producer mode, ring depths and shared-memory size."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def plan(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("plan") / "plan_check")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-o", exe,
                    os.path.join(ROOT, "tools", "plan_check.cu")], check=True, capture_output=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    rows = []
    for line in out.strip().splitlines():
        name = line.split()[0]
        rows.append((name, {k: int(v) for k, v in re.findall(r"(\w+)=(\d+)", line)}))
    return rows


def test_every_step_gemm_gets_the_staged_pipeline(plan):
    assert len(plan) == 18
    for name, p in plan:
        assert p["bulk"] == 1, name
        assert p["smem"] <= p["cap"] <= 227 * 1024, name


def test_staging_depth_is_even(plan):
    """tools/sim_pipeline.py / tests/test_pipeline_model.py: an odd depth breaks the two-loader / two-group protocol."""
    for name, p in plan:
        assert p["nraw"] >= 2 and p["nraw"] % 2 == 0, (name, p["nraw"])
        assert 2 <= p["na"] <= 4, (name, p["na"])


def test_bf16_row_outputs_use_the_staged_epilogue(plan):
    for name, p in plan:
        if p["epi"] == 0 or name == "stem1":            # GroupNorm sweeps and the bf16 stage-1 stem
            assert p["out_vec"] == 1, name
