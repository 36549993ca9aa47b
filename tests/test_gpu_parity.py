"""GPU parity: the CUDA path, called through the C ABI (ctypes) by the drop-in modules, against
(a) the golden vectors produced by the unmodified reference and (b) the numpy oracle on seeded inputs.

Tolerance (fp32 mode): atol 1e-5, rtol 1e-4 -- the reference's own fp32-vs-fp64 noise floor over 36 steps is
~2e-6.  The wet/dry mask is compared exactly outside |p - 0.5| < 1e-5.
"""
import os

import numpy as np
import pytest
import torch

from oracle import urnn_oracle as O

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-5, 1e-4
DEV = "cuda:0"


def t2n(t):
    return t.detach().cpu().numpy()


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return z, {k[2:]: z[k] for k in z.files if k.startswith("w.")}


def make_cell(meta, module, w):
    from src.lib.model.networks.ConvRNN import CGRU_cell
    k, cin, F, H, W, S, with_x = [int(v) for v in meta]
    cell = CGRU_cell(False, (H, W), cin, k, F, module)
    cell.load_state_dict({kk: torch.from_numpy(v) for kk, v in w.items()}, strict=False)
    return cell.to(DEV).eval(), (k, cin, F, H, W, S, with_x)


def cell_inputs(meta, module):
    k, cin, F, H, W, S, with_x = meta
    torch.manual_seed(1)
    x = torch.rand(S, 1, cin, H, W) if with_x else None
    hidden = torch.zeros(1, F, H, W) if module == "encoder" else torch.rand(1, 2 * F, H, W)
    return x, hidden


@pytest.mark.parametrize("name,module", [
    ("cell_enc_k1", "encoder"), ("cell_dec_k1", "decoder"), ("cell_dec_k1_nox", "decoder"),
    ("cell_enc_k3", "encoder"), ("cell_dec_k3", "decoder")])
def test_cell_vs_reference_golden(golden_dir, name, module):
    z, w = load(golden_dir, name)
    cell, meta = make_cell(z["meta"], module, w)
    x, hidden = cell_inputs(meta, module)
    S = meta[5]
    with torch.no_grad():
        out = cell(None if x is None else x.to(DEV), hidden.to(DEV), S)
    assert out.shape == (S, 1, meta[2], meta[3], meta[4])
    if "out" in z.files:
        np.testing.assert_allclose(t2n(out)[:, 0], z["out"], atol=ATOL, rtol=RTOL)
        with torch.no_grad():
            out2 = cell(None if x is None else x[:1].to(DEV), torch.from_numpy(z["hidden2"])[None].to(DEV), 1)
        np.testing.assert_allclose(t2n(out2)[0, 0], z["out2"], atol=ATOL, rtol=RTOL)
    else:
        np.testing.assert_allclose(t2n(out)[-1, 0], z["out_last"], atol=ATOL, rtol=RTOL)
        np.testing.assert_allclose(t2n(out)[:, 0].mean(axis=(2, 3)), z["out_chan_mean"], atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("H,W,cin,F,module,k", [
    (20, 36, 16, 64, "encoder", 1), (12, 200, 96, 96, "decoder", 1), (8, 8, 5, 32, "encoder", 1),
    (4, 4, 3, 32, "decoder", 1), (16, 16, 7, 32, "encoder", 5), (125, 128, 96, 96, "encoder", 1),
    (125, 125, 96, 96, "decoder", 1), (5, 7, 3, 32, "encoder", 1), (9, 11, 4, 32, "decoder", 3)])
def test_cell_vs_oracle_ragged_shapes(H, W, cin, F, module, k):
    """Shapes that do not fill a 128-pixel tile, tiny grids, odd channel counts, a 5x5 filter."""
    from src.lib.model.networks.ConvRNN import CGRU_cell
    torch.manual_seed(H * 1000 + W)
    cell = CGRU_cell(False, (H, W), cin, k, F, module).to(DEV).eval()
    x = torch.rand(1, 1, cin, H, W) * 2 - 1
    hid = torch.rand(1, F * (2 if module == "decoder" else 1), H, W) * 2 - 1
    with torch.no_grad():
        out = cell(x.to(DEV), hid.to(DEV), 1)
    w = {kk: t2n(v).astype(np.float64) for kk, v in cell.state_dict().items()}
    ref = O.cgru_cell_forward(w, "", x.numpy()[:, 0].astype(np.float64), hid.numpy()[0].astype(np.float64), module, F, 1)
    np.testing.assert_allclose(t2n(out)[:, 0], ref, atol=ATOL, rtol=RTOL)


def test_stems_vs_oracle():
    from urnn_b200 import ops
    rng = np.random.RandomState(0)
    for cin, cout, H, W, pool in [(63, 16, 20, 24, 1), (9, 16, 8, 132, 1), (64, 64, 12, 20, 2), (96, 96, 10, 8, 2),
                                  (64, 16, 4, 4, 1), (7, 40, 6, 8, 2), (96, 96, 250, 250, 2), (5, 16, 7, 9, 1)]:
        x = rng.randn(cin, H, W).astype(np.float32)
        w = (rng.randn(cout, cin, 1, 1) / np.sqrt(cin)).astype(np.float32)
        b = rng.randn(cout).astype(np.float32)
        y = ops.conv1x1_lrelu_fwd(torch.from_numpy(x).to(DEV), torch.from_numpy(w).to(DEV), torch.from_numpy(b).to(DEV), pool)
        ref = O.leaky_relu(O.conv2d_same(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64)))
        if pool == 2:
            ref = O.avg_pool2(ref)
        np.testing.assert_allclose(t2n(y), ref, atol=ATOL, rtol=RTOL)
    for cin, cout, H, W in [(96, 96, 8, 8), (96, 96, 5, 12), (32, 8, 3, 4), (96, 96, 125, 125), (8, 8, 3, 5)]:
        x = rng.randn(cin, H, W).astype(np.float32)
        w = (rng.randn(cin, cout, 2, 2) / np.sqrt(cin)).astype(np.float32)
        b = rng.randn(cout).astype(np.float32)
        y = ops.deconv2x2_lrelu_fwd(torch.from_numpy(x).to(DEV), torch.from_numpy(w).to(DEV), torch.from_numpy(b).to(DEV))
        ref = O.leaky_relu(O.conv_transpose2x2(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64)))
        np.testing.assert_allclose(t2n(y), ref, atol=ATOL, rtol=RTOL)


def build_ed(H, W, C, weights=None, math=None):
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    torch.manual_seed(0)
    enc, dec = get_network_params(False, H, W, input_channels=C, net_cfg=None, math=math)
    net = ED(False, enc, dec, 0.5, False, input_height=H, input_width=W)
    if weights is not None:
        net.load_state_dict({k: torch.from_numpy(v) for k, v in weights.items()}, strict=False)
    return net.to(DEV).eval()


def zero_states(H, W):
    return [torch.zeros(1, *s.shape, device=DEV) for s in O.zero_states(H, W)]


def test_head_vs_oracle(golden_dir):
    z, w = load(golden_dir, "ed_32x32_c9")
    net = build_ed(32, 32, 9, w)
    rng = np.random.RandomState(5)
    feat = rng.randn(16, 32, 32).astype(np.float32)
    with torch.no_grad():
        out = net.head(torch.from_numpy(feat)[None, None].to(DEV))
    w64 = {k: v.astype(np.float64) for k, v in w.items()}
    masked, prob, raw = O.head_forward(w64, feat.astype(np.float64))
    np.testing.assert_allclose(t2n(out)[0, 0, 1], prob, atol=ATOL, rtol=RTOL)
    safe = np.abs(prob - 0.5) > 1e-5
    np.testing.assert_allclose(t2n(out)[0, 0, 0][safe], masked[safe], atol=ATOL, rtol=RTOL)


def test_ed_unaligned_quarter_resolution_vs_oracle():
    """20x28: the quarter-resolution planes have 35 cells (odd), like 500x500 -> 125x125."""
    H, W, hist = 20, 28, 3
    net = build_ed(H, W, 9)
    w = {k: t2n(v).astype(np.float64) for k, v in net.state_dict().items()}
    xs = O.synthetic_event_inputs(H, W, 3, hist)
    ref_out, ref_st = O.run_sequence(w, xs.astype(np.float64))
    st = zero_states(H, W)
    with torch.no_grad():
        for t in range(3):
            out, *st = net(torch.from_numpy(xs[t])[None, None].to(DEV), *st)
    for i in range(6):
        np.testing.assert_allclose(t2n(st[i])[0], ref_st[i], atol=ATOL, rtol=RTOL)


@pytest.mark.parametrize("name,rain", [("ed_32x32_c9", (30.0, 60.0)), ("ed_24x40_c63", (6.0, 6.0))])
def test_ed_sequence_vs_reference_golden(golden_dir, name, rain):
    z, w = load(golden_dir, name)
    H, W, hist, T, every = [int(v) for v in z["meta"]]
    net = build_ed(H, W, 2 * hist + 3, w if w else None)
    if not w:
        fp = np.array([float(v.double().sum()) for v in net.state_dict().values()])
        np.testing.assert_allclose(fp, z["w_fingerprint"], rtol=1e-12)
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist, rain_scale=rain[0], rain_max=rain[1])).to(DEV)
    st = zero_states(H, W)
    with torch.no_grad():
        for t in range(T):
            out, *st = net(xs[t][None, None], *st)
            assert out.shape == (1, 1, H, W)
            safe = np.abs(z["prob"][t] - 0.5) > 1e-5
            np.testing.assert_allclose(t2n(out)[0, 0][safe], z["out"][t][safe], atol=ATOL, rtol=RTOL)
    for i in range(6):
        np.testing.assert_allclose(t2n(st[i])[0], z[f"state{i}"], atol=ATOL, rtol=RTOL)


def test_ed_lite_config2_36_steps(golden_dir):
    """BASELINE config 2: lite 128x128, C_in=9, T=36, fp32 vs the reference's output."""
    z, _ = load(golden_dir, "ed_lite128")
    H, W, hist, T, every = [int(v) for v in z["meta"]]
    net = build_ed(H, W, 2 * hist + 3)
    fp = np.array([float(v.double().sum()) for v in net.state_dict().values()])
    np.testing.assert_allclose(fp, z["w_fingerprint"], rtol=1e-12)
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist)).to(DEV)
    st = zero_states(H, W)
    flips = 0
    with torch.no_grad():
        for t in range(T):
            out, *st = net(xs[t][None, None], *st)
            if t % every == 0:
                i = t // every
                safe = np.abs(z["prob"][i] - 0.5) > 1e-5
                np.testing.assert_allclose(t2n(out)[0, 0][safe], z["out"][i][safe], atol=ATOL, rtol=RTOL)
                flips += int(((t2n(out)[0, 0] != 0) != (z["out"][i] != 0))[safe].sum())
    assert flips == 0
    for i in range(6):
        a = t2n(st[i])[0]
        np.testing.assert_allclose(a[:, ::4, ::4], z[f"state{i}_s4"], atol=ATOL, rtol=RTOL)
        np.testing.assert_allclose(a.mean(axis=(1, 2)), z[f"state{i}_mean"], atol=ATOL, rtol=RTOL)


def test_module_route_equals_fused_route(golden_dir):
    """ED.forward's per-module route (autograd-capable) and the single-call route enqueue the same kernels."""
    z, w = load(golden_dir, "ed_32x32_c9")
    net = build_ed(32, 32, 9, w)
    xs = torch.from_numpy(O.synthetic_event_inputs(32, 32, 2, 3)).to(DEV)
    st = [torch.rand_like(s) for s in zero_states(32, 32)]
    with torch.no_grad():
        a = net(xs[1][None, None], *st)
        enc = net.encoder(xs[1][None, None].permute(1, 0, 2, 3, 4), st[:3])
        feat, dec = net.decoder(enc, st[3:])
        out = net.head(feat)[:, :, 0]
    for u, v in zip(a, (out, *enc, *dec)):
        assert torch.equal(u, v)


def test_run_to_run_determinism():
    net = build_ed(64, 64, 9)
    x = torch.rand(1, 1, 9, 64, 64, device=DEV)
    st = [torch.rand_like(s) for s in zero_states(64, 64)]
    with torch.no_grad():
        a = net(x, *st)
        b = net(x, *st)
    for u, v in zip(a, b):
        assert torch.equal(u, v)


def test_launches_are_counted():
    from urnn_b200 import _capi
    lib = _capi.load()
    net = build_ed(16, 16, 9)
    st = zero_states(16, 16)
    n0 = lib.urnn_launch_count()
    with torch.no_grad():
        net(torch.rand(1, 1, 9, 16, 16, device=DEV), *st)
    torch.cuda.synchronize()
    assert lib.urnn_launch_count() - n0 == 6 * 3 + 6 + 4      # 6 cells x 3 kernels, 6 stems, 4 head sweeps


@pytest.mark.parametrize("use_graph", [False, True])
def test_sequence_runner_equals_step_loop(use_graph):
    """SequenceRunner (ping-pong states, optional CUDA graph) == the reference-style per-step loop, bit for bit."""
    from urnn_b200.runner import SequenceRunner
    H, W, T = 32, 48, 5
    net = build_ed(H, W, 9)
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, 3)).to(DEV)
    st = zero_states(H, W)
    outs = []
    with torch.no_grad():
        for t in range(T):
            out, *st = net(xs[t][None, None], *st)
            outs.append(out[0, 0])
    runner = SequenceRunner(net, H, W, 9, use_graph=use_graph)
    depth, prob, final = runner.run(xs)
    assert torch.equal(depth, torch.stack(outs))
    for a, b in zip(final, st):
        assert torch.equal(a, b[0])
    depth2, _, _ = runner.run(xs)           # second run re-initialises the states
    assert torch.equal(depth2, depth)


def test_sequence_host_entry_point():
    """urnn_ed_sequence_host (host buffers, overlapped copies) == the device-resident runner."""
    from urnn_b200.runner import SequenceRunner
    H, W = 24, 40
    net = build_ed(H, W, 9)
    for T in (1, 4, 7):
        xs_host = torch.from_numpy(O.synthetic_event_inputs(H, W, T, 3)).pin_memory()
        runner = SequenceRunner(net, H, W, 9, use_graph=False)
        depth, _, final = runner.run(xs_host.to(DEV))
        out_host, final_h = runner.run_host(xs_host)
        assert torch.equal(out_host, depth.cpu())
        for a, b in zip(final_h, final):
            assert torch.equal(a, b)


@pytest.mark.parametrize("math", ["fp32", "f16x3"])
def test_event_entry_point_equals_dense_inputs(math):
    """urnn_ed_event_host (raw maps + scalar rainfall series, rainfall folded into a per-step stage-1 bias) reproduces
    the dense-input loop built with the reference's preprocess_inputs recipe (summation order only).  The recipe
    (oracle.synthetic_event_inputs) is pinned to the reference's Dynamic2DFlood.preprocess_inputs by
    tests/test_callers_golden.py."""
    from urnn_b200.runner import SequenceRunner
    H, W, hist, T = 24, 40, 3, 7
    rng = np.random.RandomState(42)                  # same draws as O.synthetic_event_inputs
    dem = rng.rand(H, W) * 10.0
    imperv = rng.rand(H, W)
    manhole = (rng.rand(H, W) > 0.95).astype(np.float64)
    rain = rng.rand(T) * 30.0
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist)).to(DEV)
    net = build_ed(H, W, 2 * hist + 3, math=math)
    runner = SequenceRunner(net, H, W, 2 * hist + 3, use_graph=False)
    depth, _, final = runner.run_dev(xs)
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    rain32 = f32(rain)
    out_host, final_e = runner.run_event_host(f32(dem), f32(imperv), f32(manhole), rain32, torch.cumsum(rain32, 0), hist, 60.0, 250.0)
    ref = depth.cpu().numpy()
    got = out_host.numpy()
    same_mask = (ref != 0) == (got != 0)
    assert same_mask.mean() > 0.999
    np.testing.assert_allclose(got[same_mask], ref[same_mask], atol=2e-5, rtol=1e-4)
    for a, b in zip(final_e, final):
        np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), atol=2e-5, rtol=1e-4)
