"""CPU: the oracle's per-step input recipe against the reference's own Dynamic2DFlood._prepare_input + preprocess_inputs
(golden produced by tests/golden/make_golden_callers.py from the unmodified reference)."""
import os

import numpy as np

from oracle import urnn_oracle as O


def test_oracle_event_inputs_match_reference_preprocess(golden_dir):
    z = np.load(os.path.join(golden_dir, "preprocess_inputs.npz"))
    H, W, T, hist = [int(v) for v in z["meta"]]
    xs = O.synthetic_event_inputs(H, W, T, hist, seed=42, rain_scale=6.0, rain_max=6.0)
    assert xs.shape == (T, 2 * hist + 3, H, W)
    # (the reference converts the DEM to mm in fp32 before scaling it to [0,1]: rounding differences of ~1e-7)
    np.testing.assert_allclose(xs[z["steps"]], z["dense"], atol=1e-6, rtol=0)
    # zero padding before the event start, newest sample in the last rainfall channel (get_past_rainfall)
    assert np.all(z["dense"][0][:hist - 1] == 0) and np.all(z["dense"][0][hist - 1] > 0)


def test_caller_goldens_are_self_consistent(golden_dir):
    z = np.load(os.path.join(golden_dir, "window_loop.npz"))
    H, W, T, hist, ind, seq = [int(v) for v in z["meta"]]
    assert z["reg"].shape == (seq, H, W) and z["cls"].shape == (seq, H, W)
    assert set(np.unique(z["cls"])) <= {0, 1}
    assert len([k for k in z.files if k.startswith("g.")]) == 79
    assert all(float(np.abs(z[k]).max()) == 0.0 for k in z.files if k.startswith("g.head.cls_"))
