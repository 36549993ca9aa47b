"""Device-side post-processing: the numpy restatement against the reference's own compute_metrics
(golden from tests/golden/make_golden_metrics.py), and the CUDA streaming reductions (urnn_metrics_*) against both."""
import os

import numpy as np
import pytest

from oracle import metrics_oracle as MO

KEYS = ("R2", "MSE", "RMSE", "MAE", "PeakR2", "CSI")
RTOL = 1e-5            # fp64 reductions here vs numpy's float32 pairwise sums in the reference


def _golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "metrics.npz"))
    return z, {k: float(z["m." + k]) for k in KEYS}


def test_oracle_matches_reference_metrics(golden_dir):
    z, ref = _golden(golden_dir)
    got = MO.compute_metrics(MO.denormalise(z["pred_norm"], float(z["flood_max"])), z["gt_mm"], float(z["flood_thres"]))
    for k in KEYS:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, err_msg=k)


def test_oracle_edge_cases():
    T, H, W = 3, 4, 4
    z = np.zeros((T, H, W), np.float32)
    m = MO.compute_metrics(z, z)                       # all dry, perfect: the 1e-10 guards decide (test.py:642,663)
    assert m["R2"] == 1.0 and m["MSE"] == 0.0 and m["CSI"] == 0.0 and m["t_peak"] == 0
    g = z.copy(); g[1] = 200.0
    m = MO.compute_metrics(z, g)                       # everything missed: fn only
    assert (m["tp"], m["fp"], m["fn"]) == (0, 0, H * W) and m["t_peak"] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("chunks", [(9,), (4, 5), (1, 1, 7), (2, 2, 2, 2, 1)])
def test_cuda_metrics_match_reference_golden(golden_dir, chunks):
    import torch
    from urnn_b200.metrics import StreamingMetrics
    z, ref = _golden(golden_dir)
    T, H, W = z["pred_norm"].shape
    dev = "cuda:0"
    pred, gt = torch.from_numpy(z["pred_norm"]).to(dev), torch.from_numpy(z["gt_mm"]).to(dev)
    m = StreamingMetrics(H, W, T, float(z["flood_max"]), float(z["flood_thres"]), device=dev)
    t0 = 0
    for n in chunks:
        m.update(pred[t0:t0 + n].contiguous(), gt[t0:t0 + n].contiguous())
        t0 += n
    got = m.result()
    for k in KEYS:
        np.testing.assert_allclose(got[k], ref[k], rtol=RTOL, err_msg=k)
    exact = MO.compute_metrics(MO.denormalise(z["pred_norm"], float(z["flood_max"])), z["gt_mm"], float(z["flood_thres"]))
    assert (m.detail["tp"], m.detail["fp"], m.detail["fn"], m.detail["t_peak"]) == (exact["tp"], exact["fp"], exact["fn"], exact["t_peak"])
    assert m.detail["elements"] == T * H * W


@pytest.mark.gpu
def test_cuda_metrics_large_ragged_event():
    """A grid that is no multiple of the block size, more steps than one launch takes (64), negative predictions, against
    the oracle; wet / dry counts bit-exact."""
    import torch
    from urnn_b200.metrics import StreamingMetrics
    rng = np.random.RandomState(0)
    T, H, W = 70, 61, 45
    gt = (rng.rand(T, H, W).astype(np.float32) ** 4 * 1200.0 * np.sin(np.linspace(0.2, 3.0, T))[:, None, None]).astype(np.float32)
    pred = (gt / 5000.0 + rng.randn(T, H, W).astype(np.float32) * 0.01).astype(np.float32)
    exact = MO.compute_metrics(MO.denormalise(pred, 5000.0), gt, 150.0)
    dev = "cuda:0"
    m = StreamingMetrics(H, W, T, device=dev)
    m.update(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev))
    got = m.result()
    for k in KEYS:
        np.testing.assert_allclose(got[k], exact[k], rtol=RTOL, err_msg=k)
    assert (m.detail["tp"], m.detail["fp"], m.detail["fn"], m.detail["t_peak"]) == (exact["tp"], exact["fp"], exact["fn"], exact["t_peak"])
    m.reset()                                           # a second event through the same object
    m.update(torch.from_numpy(pred).to(dev), torch.from_numpy(gt).to(dev))
    again = m.result()                                  # atomics: the summation order differs run to run in the last bits
    for k in KEYS:
        np.testing.assert_allclose(again[k], got[k], rtol=1e-12, err_msg=k)


@pytest.mark.gpu
def test_cuda_metrics_argument_errors():
    import torch
    from urnn_b200.metrics import StreamingMetrics
    m = StreamingMetrics(8, 8, 4, device="cuda:0")
    x = torch.zeros(2, 8, 8, device="cuda:0")
    with pytest.raises(ValueError):
        m.update(x.double(), x)
    with pytest.raises(ValueError):
        m.update(torch.zeros(2, 8, 9, device="cuda:0"), x)
    m.update(x, x)
    with pytest.raises(RuntimeError):
        m.result()                                      # only 2 of 4 steps seen
    with pytest.raises(ValueError):
        m.update(torch.zeros(3, 8, 8, device="cuda:0"), torch.zeros(3, 8, 8, device="cuda:0"))


def test_metrics_abi_argument_checks_without_a_gpu():
    """The C entry points validate before touching the device: sizes, NULL workspace, step ranges (no GPU needed)."""
    import ctypes as C
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "u-rnn_b200"))
    from urnn_b200 import _capi
    lib = _capi.load()
    assert lib.urnn_metrics_workspace_bytes(0, 8, 4) == 0 and lib.urnn_metrics_workspace_bytes(8, 8, 0) == 0
    need = lib.urnn_metrics_workspace_bytes(500, 500, 180)
    assert need >= 2 * 500 * 500 * 4 + 180 * 32 and need % 256 == 0          # two fp32 maxima maps + per-step sums
    assert lib.urnn_metrics_reset(8, 8, 4, None, 0, None) != 0
    assert b"workspace" in lib.urnn_last_error()
    fake = C.c_void_p(0x1000)                                                  # never dereferenced: the size check comes first
    assert lib.urnn_metrics_reset(8, 8, 4, fake, 16, None) != 0
    assert lib.urnn_metrics_accumulate(8, 8, 4, 3, 2, fake, fake, 5000.0, fake, 1 << 20, None) != 0      # steps [3, 5) of 4
    assert b"outside" in lib.urnn_last_error()
    assert lib.urnn_metrics_finalize(8, 8, 4, 150.0, fake, 1 << 20, None, None) != 0
