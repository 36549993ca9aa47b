"""The mbarrier protocol of the bulk-mode pixel-GEMM (2 loader warps, 2 converter groups, in-order MMA issuer, double
buffered accumulator) checked on the randomised model in tools/sim_pipeline.py: an even staging depth never violates
slot ownership or deadlocks; an odd depth does (the hang found on the GPU, tc_pixgemm.cuh: gemm_smem_bytes_bulk)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load():
    spec = importlib.util.spec_from_file_location("sim_pipeline", os.path.join(ROOT, "tools", "sim_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    mod.__name__ = "sim_pipeline"          # not "__main__": only the definitions are executed
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("nraw", [2, 4, 6])
def test_even_staging_depth_is_safe(nraw):
    sim = _load()
    for na in (2, 3, 4):
        for nu in (1, 2, 3, 7, 10):
            for ntile in (1, 3):
                for seed in range(3):
                    assert sim.run(nraw, na, nu, ntile, seed) == "ok", (nraw, na, nu, ntile, seed)


def test_odd_staging_depth_breaks_the_protocol():
    sim = _load()
    bad = 0
    for seed in range(40):
        try:
            if sim.run(3, 3, 7, 3, seed) != "ok":
                bad += 1
        except AssertionError:
            bad += 1
    assert bad > 0
