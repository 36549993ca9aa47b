"""GPU: the reference's OWN loops, unmodified, on top of the drop-in modules.

oracle/_ref/callers.py holds test.py:326-377 `Inference` and main.py:598-692 `process_window` (+ `prewarming`,
`accumulate_predictions`), extracted verbatim by oracle/make_ref.py; the dataset / state helpers and the loss they use are
the reference's files as well.  Only `src.lib.model.networks.{model, net_params, ...}` resolve to this repository (same
sys.path overlay as tools/run_with_urnn_b200.py).  Outputs are compared with what the same loops produced on the
unmodified reference network (tests/golden/make_golden_callers.py)."""
import argparse
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ref_callers():
    if not os.path.exists(os.path.join(REF, "callers.py")):
        pytest.skip("oracle/_ref is not built (python oracle/make_ref.py in the build container)")
    if REF not in sys.path:
        sys.path.append(REF)                    # after the drop-in package: the overlay of INTEGRATION.md
    # earlier tests may have imported the drop-in packages before the overlay path existed: recompute their search paths
    # (what their __init__ does at first import, pkgutil.extend_path)
    import pkgutil
    import src.lib.model as pkg_model
    import src.lib.model.networks as pkg_networks
    for pkg in (pkg_model, pkg_networks):
        pkg.__path__ = pkgutil.extend_path(pkg.__path__, pkg.__name__)
    import callers
    import src.lib.model.networks.model as m
    assert "u-rnn_b200" in m.__file__           # the network under the reference's loops is ours
    import src.lib.dataset.Dynamic2DFlood as d
    assert os.path.join("oracle", "_ref") in d.__file__
    return callers


def batch_of(z, T, device):
    from src.lib.dataset.Dynamic2DFlood import Dynamic2DFlood
    ev = {k[3:]: z[k] for k in z.files if k.startswith("ev.")}
    d = Dynamic2DFlood._prepare_input(None, ev, None, duration=T)
    return {k: v.unsqueeze(0).to(device) for k, v in d.items()}       # DataLoader batch of 1 + to_device (test.py:447)


def build(H, W, C, math):
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    from src.lib.utils.net_config import load_net_config
    cfg = load_net_config(None)
    torch.manual_seed(0)
    p = get_network_params(False, H, W, input_channels=C, net_cfg=cfg, math=math)
    return ED(False, p[0], p[1], 0.5, False, input_height=H, input_width=W), cfg


@pytest.mark.parametrize("math,atol", [("fp32", 1e-5), ("f16x3", 2e-5)])
def test_reference_inference_loop_runs_on_the_dropin(golden_dir, ref_callers, math, atol):
    z = np.load(os.path.join(golden_dir, "inference_loop.npz"))
    H, W, T, hist = [int(v) for v in z["meta"]]
    net, cfg = build(H, W, 2 * hist + 3, math)
    np.testing.assert_allclose(np.array([float(v.double().sum()) for v in net.state_dict().values()]), z["w_fingerprint"], rtol=1e-12)
    net = net.to(DEV)
    out = ref_callers.Inference(net, batch_of(z, T, DEV), DEV, historical_nums=hist, rain_max=60.0, cumsum_rain_max=250.0,
                                input_height=H, input_width=W, net_cfg=cfg)
    assert out.shape == (T, H, W)
    ref = z["out"]
    wet_same = (out != 0) == (ref != 0)
    assert wet_same.mean() > 0.999                                   # mask flips only inside the |p - 0.5| band
    np.testing.assert_allclose(out[wet_same], ref[wet_same], atol=atol, rtol=1e-4)


def test_reference_window_loop_and_loss_run_on_the_dropin(golden_dir, ref_callers):
    """main.py:598-692 process_window with pre-warming, the reference's FocalBCE_and_WMSE, backward: prediction, loss and
    all 79 gradients against the reference's (fp32 mode: the backward kernels are fp32)."""
    from src.lib.model.networks.losses import select_loss_function
    z = np.load(os.path.join(golden_dir, "window_loop.npz"))
    H, W, T, hist, ind, seq = [int(v) for v in z["meta"]]
    net, cfg = build(H, W, 2 * hist + 3, "fp32")
    np.testing.assert_allclose(np.array([float(v.double().sum()) for v in net.state_dict().values()]), z["w_fingerprint"], rtol=1e-12)
    net = net.to(DEV).train()
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    args = argparse.Namespace(input_height=H, input_width=W, net_cfg=cfg, prewarming=True, historical_nums=hist,
                              rain_max=60.0, cumsum_rain_max=250.0, seq_num=seq)
    pred, final_states = ref_callers.process_window(ind, args, net, batch_of(z, T, DEV), DEV, opt)
    np.testing.assert_allclose(pred["reg"].detach().cpu().numpy()[0], z["reg"], atol=1e-5, rtol=1e-4)
    np.testing.assert_array_equal(pred["cls"].cpu().numpy()[0], z["cls"])
    for i, s in enumerate(final_states[0] + final_states[1]):
        np.testing.assert_allclose(s.cpu().numpy()[0], z[f"final{i}"], atol=1e-5, rtol=1e-4)
    label = torch.from_numpy(z["label"])[None].to(DEV)
    losses = select_loss_function("FocalBCE_and_WMSE", "mean")(pred, label[:, ind:ind + seq], 0)
    np.testing.assert_allclose(float(losses["loss"]), float(z["loss"]), rtol=1e-5)
    losses["loss"].backward()
    n = 0
    for k, v in net.named_parameters():
        ref = z["g." + k]
        g = v.grad.cpu().numpy() if v.grad is not None else np.zeros_like(ref)
        scale = max(float(np.abs(ref).max()), 1e-6)
        assert float(np.abs(g - ref).max()) <= 1e-3 * scale + 1e-7, (k, float(np.abs(g - ref).max()), scale)
        n += 1
    assert n == 79
