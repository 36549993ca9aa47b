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


def _load_sharded():
    spec = importlib.util.spec_from_file_location("sim_sharded_pipeline", os.path.join(ROOT, "tools", "sim_sharded_pipeline.py"))
    mod = importlib.util.module_from_spec(spec)
    mod.__name__ = "sim_sharded_pipeline"
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_step_on_one_stream_never_deadlocks(world):
    """What ships: sharded sequence calls keep every launch on one stream (urnn_v2.cu: pipe_wanted)."""
    sim = _load_sharded()
    for seed in range(60):
        assert sim.run(world=world, streams=1, pdl=True, seed=seed) == "ok", seed


@pytest.mark.parametrize("world", [2, 4])
def test_two_streams_deadlock_only_with_dependent_launch_residency(world):
    """The encoder / decoder pipeline in a sharded run: a launch that is already resident under programmatic dependent
    launch holds the SMs the other stream needs on the peer a spinning last CTA waits for (the 4-GPU hang).
    Without that residency the two streams always finish -- the order a future sharded pipeline has to use."""
    sim = _load_sharded()
    with_pdl = [sim.run(world=world, streams=2, pdl=True, seed=s) for s in range(60)]
    without = [sim.run(world=world, streams=2, pdl=False, seed=s) for s in range(60)]
    assert "deadlock" in with_pdl
    assert set(without) == {"ok"}
