"""metrics_oracle.py -- TEST INFRASTRUCTURE (only tests/ may import it): numpy restatement of the reference's inference
post-processing, pinned against the reference's own function by tests/golden/metrics.npz
(tests/golden/make_golden_metrics.py runs test.py:607-675 compute_metrics, extracted verbatim by oracle/make_ref.py).

  denormalise      Dynamic2DFlood.py:379-385 r_MinMaxScaler as test.py:468 calls it (max = flood_max, min = 0)
  compute_metrics  test.py:607-675: R2, MSE, RMSE, MAE over all space x time in metres; PeakR2 at the step of maximum
                   spatial-mean ground truth; CSI of the temporal-maximum wet masks (> flood_thres mm)
Element-wise arithmetic in the arrays' own dtype (float32 like the reference's arrays), reductions in float64 (the
reference reduces in float32 with numpy's pairwise summation: agreement to ~1e-6 relative, tolerance in the tests 1e-5)."""
import numpy as np


def denormalise(pred_norm, flood_max):
    return pred_norm * (flood_max - 0) + 0                      # Dynamic2DFlood.py:385


def compute_metrics(pred_mm, gt_mm, flood_thres=150.0):
    pred_m = pred_mm / 1000.0                                    # test.py:636-637
    gt_m = gt_mm / 1000.0
    d = pred_m - gt_m
    n = d.size
    ss_res = float(np.sum((d * d).astype(np.float64)))          # test.py:640
    g64 = gt_m.astype(np.float64)
    ss_tot = float(np.sum((g64 - g64.mean()) ** 2))              # test.py:641
    r2 = 1.0 - ss_res / (ss_tot + 1e-10)
    mse = ss_res / n                                             # test.py:645-647
    mae = float(np.sum(np.abs(d).astype(np.float64))) / n
    t_peak = int(np.argmax(g64.mean(axis=(1, 2))))               # test.py:650-651
    dp = d[t_peak].astype(np.float64)
    gp = g64[t_peak]
    peak_r2 = 1.0 - float(np.sum(dp * dp)) / (float(np.sum((gp - gp.mean()) ** 2)) + 1e-10)
    pf = pred_mm.max(axis=0) > flood_thres                       # test.py:658-663
    gf = gt_mm.max(axis=0) > flood_thres
    tp = int(np.sum(pf & gf)); fp = int(np.sum(pf & ~gf)); fn = int(np.sum(~pf & gf))
    return {"R2": r2, "MSE": mse, "RMSE": float(np.sqrt(mse)), "MAE": mae, "PeakR2": peak_r2,
            "CSI": tp / (tp + fp + fn + 1e-10), "tp": tp, "fp": fp, "fn": fn, "t_peak": t_peak}
