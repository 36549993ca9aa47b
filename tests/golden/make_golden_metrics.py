"""
make_golden_metrics.py -- golden vector for the device-side metrics, produced by the reference's own
function (build container only: needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_metrics.py

  metrics.npz   test.py:468 r_MinMaxScaler + test.py:607-675 compute_metrics (extracted verbatim into oracle/_ref/callers.py)
                on a seeded synthetic event: normalised prediction, ground truth in mm, the six metrics
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
from oracle import make_ref  # noqa: E402

REF = make_ref.make()
sys.path.insert(0, REF)
import callers  # noqa: E402
from src.lib.dataset.Dynamic2DFlood import r_MinMaxScaler  # noqa: E402


def synthetic(T, H, W, seed):
    """A flood that rises and recedes: ground truth in mm (float32), a normalised prediction = truth + noise, some negatives."""
    rng = np.random.RandomState(seed)
    base = rng.rand(H, W).astype(np.float32) ** 3
    hydro = np.sin(np.linspace(0.1, 2.8, T)).astype(np.float32)
    gt_mm = (base[None] * hydro[:, None, None] * 900.0).astype(np.float32)
    gt_mm[gt_mm < 40.0] = 0.0
    pred = gt_mm / 5000.0 + rng.randn(T, H, W).astype(np.float32) * 0.004
    pred[rng.rand(T, H, W) < 0.3] *= 0.0                              # dry cells of the classification mask
    return pred.astype(np.float32), gt_mm


def main():
    T, H, W = 9, 20, 28
    pred, gt_mm = synthetic(T, H, W, 3)
    out_mm = r_MinMaxScaler(pred, max=5000.0, min=0)
    m = callers.compute_metrics(out_mm, gt_mm, flood_thres=150.0)
    np.savez(os.path.join(HERE, "metrics.npz"), pred_norm=pred, gt_mm=gt_mm, flood_max=np.float32(5000.0), flood_thres=np.float32(150.0),
             **{"m." + k: np.float64(v) for k, v in m.items()})
    print({k: float(v) for k, v in m.items()})


if __name__ == "__main__":
    main()
