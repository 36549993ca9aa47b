"""Pins the CPU oracle (oracle/urnn_oracle.py) against the golden vectors generated from the
unmodified reference by tests/golden/make_golden.py.  Tolerance: the reference's own fp32-vs-fp64
noise floor is ~2e-6; the oracle runs in fp64 here so its distance to the fp32
reference is that floor: atol 1e-5, rtol 1e-4."""
import os

import numpy as np
import pytest
import torch

from oracle import urnn_oracle as O

ATOL, RTOL = 1e-5, 1e-4


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    w = {k[2:]: z[k].astype(np.float64) for k in z.files if k.startswith("w.")}
    return z, w


def cell_inputs(meta, module):
    k, cin, F, H, W, S, with_x = [int(v) for v in meta]
    torch.manual_seed(1)                       # same draw order as make_golden.cell_case
    x = torch.rand(S, 1, cin, H, W).numpy()[:, 0].astype(np.float64) if with_x else None
    if module == "encoder":
        hidden = np.zeros((F, H, W))
    else:
        hidden = torch.rand(1, 2 * F, H, W).numpy()[0].astype(np.float64)
    return x, hidden, F, S


@pytest.mark.parametrize("name,module", [
    ("cell_enc_k1", "encoder"), ("cell_dec_k1", "decoder"), ("cell_dec_k1_nox", "decoder"),
    ("cell_enc_k3", "encoder"), ("cell_dec_k3", "decoder")])
def test_cell_matches_reference(golden_dir, name, module):
    z, w = load(golden_dir, name)
    x, hidden, F, S = cell_inputs(z["meta"], module)
    out = O.cgru_cell_forward(w, "", x, hidden, module, F, S)
    if "out" in z.files:
        np.testing.assert_allclose(out, z["out"], atol=ATOL, rtol=RTOL)
        h2 = z["hidden2"].astype(np.float64)
        out2 = O.cgru_cell_forward(w, "", None if x is None else x[:1], h2, module, F, 1)[0]
        np.testing.assert_allclose(out2, z["out2"], atol=ATOL, rtol=RTOL)
    else:
        np.testing.assert_allclose(out[-1], z["out_last"], atol=ATOL, rtol=RTOL)
        np.testing.assert_allclose(out.mean(axis=(2, 3)), z["out_chan_mean"], atol=ATOL, rtol=RTOL)


def test_ed_steps_match_reference(golden_dir):
    z, w = load(golden_dir, "ed_32x32_c9")
    H, W, hist, T, every = [int(v) for v in z["meta"]]
    xs = O.synthetic_event_inputs(H, W, T, hist).astype(np.float64)
    states = O.zero_states(H, W, np.float64)
    for t in range(T):
        res = O.ed_step(w, xs[t], states)
        states = res["states"]
        np.testing.assert_allclose(res["prob"], z["prob"][t], atol=ATOL, rtol=RTOL)
        np.testing.assert_allclose(res["depth_raw"], z["depth_raw"][t], atol=ATOL, rtol=RTOL)
        safe = np.abs(res["prob"] - 0.5) > 1e-5        # mask is exact outside the eps band
        np.testing.assert_allclose(res["out"][safe], z["out"][t][safe], atol=ATOL, rtol=RTOL)
    for i in range(6):
        np.testing.assert_allclose(states[i], z[f"state{i}"], atol=ATOL, rtol=RTOL)


def test_oracle_float32_close_to_float64(golden_dir):
    """Error budget: the oracle in fp32 stays within the same tolerance of the fp32 reference."""
    z, w = load(golden_dir, "ed_32x32_c9")
    w32 = {k: v.astype(np.float32) for k, v in w.items()}
    H, W, hist, T, _ = [int(v) for v in z["meta"]]
    xs = O.synthetic_event_inputs(H, W, T, hist)
    out, states = O.run_sequence(w32, xs)
    assert out.dtype == np.float32
    for i in range(6):
        np.testing.assert_allclose(states[i], z[f"state{i}"], atol=ATOL, rtol=RTOL)


def test_synthetic_inputs_shape():
    xs = O.synthetic_event_inputs(8, 12, 5, 3)
    assert xs.shape == (5, 9, 8, 12) and xs.dtype == np.float32
    assert np.all(xs[0, :2] == 0) and np.all(xs[0, 3:5] == 0)     # zero-padded history at t=0


def test_torch_port_matches_reference(golden_dir):
    """The functional-torch port (bench.py's CPU baseline) reproduces the reference's outputs."""
    from oracle import torch_port as TP
    z, w = load(golden_dir, "ed_32x32_c9")
    p = {k: torch.from_numpy(v.astype(np.float32)) for k, v in w.items()}
    H, W, hist, T, _ = [int(v) for v in z["meta"]]
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist))
    st = [torch.from_numpy(s)[None] for s in O.zero_states(H, W)]
    with torch.no_grad():
        for t in range(T):
            depth, prob, st = TP.ed_step(p, xs[t][None], st)
            np.testing.assert_allclose(prob.numpy()[0, 0], z["prob"][t], atol=ATOL, rtol=RTOL)
            safe = np.abs(z["prob"][t] - 0.5) > 1e-5
            np.testing.assert_allclose(depth.numpy()[0, 0][safe], z["out"][t][safe], atol=ATOL, rtol=RTOL)
    for i in range(6):
        np.testing.assert_allclose(st[i].numpy()[0], z[f"state{i}"], atol=ATOL, rtol=RTOL)
