"""Backward parity (BASELINE config 4 building blocks): the CUDA backward of every op against torch autograd of the
op-for-op CPU port (oracle/torch_port.py, fp64), and the whole encoder-decoder against the gradients the unmodified
reference produced for a 3-step window (tests/golden/ed_grad_16x16.npz).  Tolerance rtol 1e-3
on the gradient scale: |d| <= 1e-3 * max|ref| + 1e-6 elementwise."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import torch_port as TP
from oracle import urnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def close(a, ref, what, rtol=1e-3):
    a = a.detach().cpu().double().numpy(); ref = ref.detach().cpu().double().numpy()
    tol = rtol * np.abs(ref).max() + 1e-6
    err = np.abs(a - ref).max()
    assert err <= tol, f"{what}: max err {err:.3e} > tol {tol:.3e}"


@pytest.mark.parametrize("cin,cout,H,W,pool", [(9, 16, 8, 12, 1), (64, 64, 12, 20, 2), (96, 96, 10, 8, 2), (64, 16, 7, 9, 1)])
def test_conv_stem_backward(cin, cout, H, W, pool):
    from urnn_b200 import ops
    torch.manual_seed(cin + H)
    x = torch.randn(cin, H, W, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(cout, cin, 1, 1, dtype=torch.float64) / cin ** 0.5).requires_grad_(True)
    b = torch.randn(cout, dtype=torch.float64, requires_grad=True)
    y = F.leaky_relu(F.conv2d(x[None], w, b), 0.2)
    if pool == 2:
        y = F.avg_pool2d(y, 2, 2)
    g = torch.randn_like(y)
    y.backward(g)
    xc, wc, bc = [t.detach().float().to(DEV).requires_grad_(True) for t in (x, w, b)]
    yc = ops.conv1x1_lrelu(xc, wc, bc, pool=pool, math="fp32")
    yc.backward(g[0].float().to(DEV))
    close(xc.grad, x.grad, "dx"); close(wc.grad, w.grad, "dw"); close(bc.grad, b.grad, "db")


@pytest.mark.parametrize("cin,cout,H,W", [(96, 96, 8, 8), (32, 8, 5, 12), (96, 96, 5, 7)])
def test_deconv_stem_backward(cin, cout, H, W):
    from urnn_b200 import ops
    torch.manual_seed(cin + H)
    x = torch.randn(cin, H, W, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(cin, cout, 2, 2, dtype=torch.float64) / cin ** 0.5).requires_grad_(True)
    b = torch.randn(cout, dtype=torch.float64, requires_grad=True)
    y = F.leaky_relu(F.conv_transpose2d(x[None], w, b, stride=2), 0.2)
    g = torch.randn_like(y)
    y.backward(g)
    xc, wc, bc = [t.detach().float().to(DEV).requires_grad_(True) for t in (x, w, b)]
    yc = ops.deconv2x2_lrelu(xc, wc, bc, math="fp32")
    yc.backward(g[0].float().to(DEV))
    close(xc.grad, x.grad, "dx"); close(wc.grad, w.grad, "dw"); close(bc.grad, b.grad, "db")


@pytest.mark.parametrize("H,W,cin,nf,module,with_x", [
    (16, 16, 16, 64, "encoder", True), (12, 20, 96, 64, "decoder", True), (8, 8, 96, 96, "decoder", False),
    (10, 14, 64, 96, "encoder", True), (5, 7, 3, 32, "decoder", True)])
def test_cell_backward(H, W, cin, nf, module, with_x):
    from src.lib.model.networks.ConvRNN import CGRU_cell
    torch.manual_seed(H * 7 + W)
    cell = CGRU_cell(False, (H, W), cin, 1, nf, module, math="fp32")
    with torch.no_grad():                                   # non-trivial GroupNorm affines
        for m in (cell.conv1[1], cell.conv2[1]):
            m.weight.uniform_(0.5, 1.5); m.bias.uniform_(-0.5, 0.5)
    p64 = {k: v.detach().double().requires_grad_(True) for k, v in cell.state_dict().items() if "wrapper" not in k}
    x = (torch.rand(1, cin, H, W, dtype=torch.float64) * 2 - 1).requires_grad_(True) if with_x else None
    hid = (torch.rand(1, nf * (2 if module == "decoder" else 1), H, W, dtype=torch.float64) * 2 - 1).requires_grad_(True)
    out = TP.cgru_cell(p64, "", x, hid, module, nf)
    g = torch.randn_like(out)
    out.backward(g)
    cellc = cell.to(DEV)
    xc = x.detach().float().to(DEV).requires_grad_(True) if with_x else None
    hc = hid.detach().float().to(DEV).requires_grad_(True)
    outc = cellc(None if xc is None else xc[None], hc, 1)
    outc.backward(g.float().to(DEV)[None])
    if with_x:
        close(xc.grad, x.grad, "dx")
    close(hc.grad, hid.grad, "dhidden")
    names = {"conv1.0.weight": cellc.conv1[0].weight, "conv1.0.bias": cellc.conv1[0].bias,
             "conv1.1.weight": cellc.conv1[1].weight, "conv1.1.bias": cellc.conv1[1].bias,
             "conv2.0.weight": cellc.conv2[0].weight, "conv2.0.bias": cellc.conv2[0].bias,
             "conv2.1.weight": cellc.conv2[1].weight, "conv2.1.bias": cellc.conv2[1].bias}
    for k, t in names.items():
        ref = p64[k].grad
        if not with_x and k.endswith("0.weight"):
            # the zero-input columns get exactly zero gradient (ConvRNN.py:143-146: x is a zeros tensor)
            assert float(t.grad[:, :cin].abs().max()) == 0.0
        close(t.grad, ref, k)


def test_head_backward(golden_dir):
    """All 19 head tensors + the feature gradient; a non-zero gradient on the probability channel exercises the cls
    branch, which main.py never does."""
    from src.lib.model.networks.head.flood_head import YOLOXHead
    H, W = 12, 20
    torch.manual_seed(2)
    head = YOLOXHead(0.5, use_checkpoint=False, input_height=H, input_width=W)
    with torch.no_grad():
        for blk in (head.stems, head.cls_convs[0], head.cls_convs[1], head.reg_convs[0], head.reg_convs[1]):
            blk.ln.weight.uniform_(0.5, 1.5); blk.ln.bias.uniform_(-0.5, 0.5)
    feat = torch.randn(1, 16, H, W, dtype=torch.float64, requires_grad=True)
    p = {("head." + k): v.detach().double().requires_grad_(True) for k, v in head.state_dict().items() if "wrapper" not in k}

    def ref_head(x):
        bc = lambda pre, t: F.silu(F.layer_norm(F.conv2d(t, p[pre + ".conv.weight"]), t.shape[1:], p[pre + ".ln.weight"], p[pre + ".ln.bias"], 1e-5))
        s = bc("head.stems", x)
        c = bc("head.cls_convs.1", bc("head.cls_convs.0", s)); r = bc("head.reg_convs.1", bc("head.reg_convs.0", s))
        prob = torch.sigmoid(F.conv2d(c, p["head.cls_preds.conv.weight"], p["head.cls_preds.conv.bias"]))
        depth = F.leaky_relu(F.conv2d(r, p["head.reg_preds.conv.weight"], p["head.reg_preds.conv.bias"]), 0.2)
        return torch.cat([depth * (prob >= 0.5).double(), prob], 1)

    out = ref_head(feat)
    g = torch.randn_like(out)
    out.backward(g)
    hc = head.to(DEV)
    fc = feat.detach().float().to(DEV).requires_grad_(True)
    oc = hc(fc[None])                      # (S=1,B=1,16,H,W) -> (1,1,2,H,W)
    oc.backward(g.float().to(DEV)[None])
    close(fc.grad, feat.grad, "dfeat")
    for k, t in hc.state_dict(keep_vars=True).items():
        if "wrapper" in k:
            continue
        close(t.grad, p["head." + k].grad, k)


def test_ed_window_gradients_match_reference(golden_dir):
    """3-step window, BPTT through the states, MSE on the depth output: all 79 parameter gradients and the gradients
    w.r.t. the 6 incoming states against what the unmodified reference produced (tests/golden/make_golden.py)."""
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    z = np.load(os.path.join(golden_dir, "ed_grad_16x16.npz"))
    H, W, hist, T = [int(v) for v in z["meta"]]
    torch.manual_seed(0)
    enc, dec = get_network_params(False, H, W, input_channels=2 * hist + 3, math="fp32")
    net = ED(False, enc, dec, 0.5, False, input_height=H, input_width=W)
    fp = np.array([float(v.double().sum()) for v in net.state_dict().values()])
    np.testing.assert_allclose(fp, z["w_fingerprint"], rtol=1e-12)
    net = net.to(DEV).train()
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist)).to(DEV)
    st0 = [torch.from_numpy(z[f"state{i}"])[None].to(DEV).requires_grad_(True) for i in range(6)]
    label = torch.from_numpy(z["label"])[None].to(DEV)
    cur, regs = st0, []
    for t in range(T):
        out, *cur = net(xs[t][None, None], *cur)
        regs.append(out)
    reg = torch.cat(regs, dim=1)
    np.testing.assert_allclose(reg.detach().cpu().numpy()[0], z["reg"], atol=1e-5, rtol=1e-4)
    loss = ((reg - label) ** 2).mean()
    np.testing.assert_allclose(float(loss), float(z["loss"]), rtol=1e-5)
    loss.backward()
    for i in range(6):
        close(st0[i].grad[0], torch.from_numpy(z[f"gstate{i}"]), f"d state{i}")
    n = 0
    for k, v in net.named_parameters():
        ref = torch.from_numpy(z["g." + k])
        if "cls_" in k:
            assert float(v.grad.abs().max()) == 0.0, k        # the classification branch never trains
            assert float(ref.abs().max()) == 0.0
        close(v.grad, ref, k)
        n += 1
    assert n == 79
