"""URNN_MATH_BF16 (tcgen05) mode on the GPU.

Two checks per case:
  * against the oracle's model of the mode (oracle quant="bf16": GEMM operands rounded to bf16, products exact,
    fp32 accumulation, pre-GroupNorm maps stored as bf16, statistics from the fp32 values): the kernel must
    implement exactly that dataflow, so the only differences are rare one-ulp bf16 rounding flips where the
    fp32 and fp64 pre-rounding values straddle a rounding boundary: mean |err| <= 5e-6, max |err| <= 8e-3;
  * against the exact fp32-semantics oracle: the stated bf16-mode tolerance, max |err| <= 3e-2 and
    mean |err| <= 2e-3 per cell step on O(1) states (SURVEY.md 8d config 3 budget: max |dstate| <= 1e-1).
"""
import numpy as np
import pytest
import torch

from oracle import urnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("H,W,cin,F,module,with_x", [
    (16, 16, 16, 64, "encoder", True), (20, 36, 16, 64, "encoder", True), (64, 64, 96, 64, "decoder", True),
    (32, 32, 96, 96, "decoder", False), (125, 125, 96, 96, "decoder", True), (9, 11, 5, 32, "encoder", True),
    (60, 100, 64, 96, "encoder", True), (128, 256, 96, 64, "decoder", True)])
def test_bf16_cell_vs_quantised_oracle(H, W, cin, F, module, with_x):
    from src.lib.model.networks.ConvRNN import CGRU_cell
    torch.manual_seed(H * 31 + W)
    cell = CGRU_cell(False, (H, W), cin, 1, F, module, math="bf16").to(DEV).eval()
    x = (torch.rand(1, 1, cin, H, W) * 2 - 1) if with_x else None
    hid = torch.rand(1, F * (2 if module == "decoder" else 1), H, W) * 2 - 1
    with torch.no_grad():
        out = cell(None if x is None else x.to(DEV), hid.to(DEV), 1).cpu().numpy()[0, 0]
    w = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in cell.state_dict().items()}
    xn = None if x is None else x.numpy()[:, 0].astype(np.float64)
    hn = hid.numpy()[0].astype(np.float64)
    refq = O.cgru_cell_forward(w, "", xn, hn, module, F, 1, quant="bf16")[0]
    ref = O.cgru_cell_forward(w, "", xn, hn, module, F, 1)[0]
    assert np.abs(out - refq).max() <= 8e-3
    assert np.abs(out - refq).mean() <= 5e-6
    assert np.abs(out - ref).max() <= 3e-2
    assert np.abs(out - ref).mean() <= 2e-3


def test_bf16_ed_sequence_tracks_fp32(golden_dir):
    """Whole encoder-decoder, 6 steps at 64x64: bf16 gate contractions vs the fp32 path of the same library.
    Stated tolerance (BASELINE.md bf16 probe): rms(d depth) <= 2e-2 * rms(depth) + 1e-3, max |d state| <= 1e-1,
    mask flip rate <= 1 %."""
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params

    def build(math):
        torch.manual_seed(0)
        enc, dec = get_network_params(False, 64, 64, input_channels=9, math=math)
        return ED(False, enc, dec, 0.5, False, input_height=64, input_width=64).to(DEV).eval()

    nets = {m: build(m) for m in ("fp32", "bf16")}
    xs = torch.from_numpy(O.synthetic_event_inputs(64, 64, 6, 3)).to(DEV)
    res = {}
    for m, net in nets.items():
        st = [torch.zeros(1, *s.shape, device=DEV) for s in O.zero_states(64, 64)]
        outs = []
        with torch.no_grad():
            for t in range(6):
                out, *st = net(xs[t][None, None], *st)
                outs.append(out)
        res[m] = (torch.cat(outs).cpu().numpy(), [s.cpu().numpy() for s in st])
    d32, s32 = res["fp32"]
    d16, s16 = res["bf16"]
    for a, b in zip(s32, s16):
        assert np.abs(a - b).max() <= 1e-1
    both = (d32 != 0) & (d16 != 0)
    rms = np.sqrt(np.mean((d32 - d16)[both] ** 2))
    assert rms <= 2e-2 * np.sqrt(np.mean(d32[both] ** 2)) + 1e-3
    assert np.mean((d32 != 0) != (d16 != 0)) <= 1e-2


def test_bf16_stems_vs_quantised_oracle():
    """tcgen05 stems: operands rounded to bf16, fp32 accumulate, LeakyReLU / AvgPool2 / 2x2 scatter in fp32."""
    from urnn_b200 import ops
    rng = np.random.RandomState(0)
    t = lambda a: torch.from_numpy(a).to(DEV)
    for cin, cout, H, W, pool in [(63, 16, 20, 24, 1), (9, 16, 8, 132, 1), (64, 64, 12, 20, 2), (96, 96, 10, 8, 2),
                                  (64, 16, 4, 4, 1), (7, 40, 6, 8, 2), (96, 96, 250, 250, 2), (5, 16, 7, 9, 1)]:
        x = rng.randn(cin, H, W).astype(np.float32)
        w = (rng.randn(cout, cin, 1, 1) / np.sqrt(cin)).astype(np.float32)
        b = rng.randn(cout).astype(np.float32)
        y = ops.conv1x1_lrelu_fwd(t(x), t(w), t(b), pool, math="bf16").cpu().numpy()
        ref = O.leaky_relu(O.conv2d_same(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), "bf16"))
        if pool == 2:
            ref = O.avg_pool2(ref)
        np.testing.assert_allclose(y, ref, atol=2e-5, rtol=1e-4)
    for cin, cout, H, W in [(96, 96, 8, 8), (96, 96, 5, 12), (32, 8, 3, 4), (96, 96, 125, 125), (8, 8, 3, 5), (16, 100, 6, 6)]:
        x = rng.randn(cin, H, W).astype(np.float32)
        w = (rng.randn(cin, cout, 2, 2) / np.sqrt(cin)).astype(np.float32)
        b = rng.randn(cout).astype(np.float32)
        y = ops.deconv2x2_lrelu_fwd(t(x), t(w), t(b), math="bf16").cpu().numpy()
        ref = O.leaky_relu(O.conv_transpose2x2(x.astype(np.float64), w.astype(np.float64), b.astype(np.float64), "bf16"))
        np.testing.assert_allclose(y, ref, atol=2e-5, rtol=1e-4)


def test_bf16_ed_fused_equals_module_route_and_tracks_model():
    """bf16 mode: the single-call route (bf16 stem maps) and the per-module route (fp32 stem maps, rounded by the
    consumer) are numerically the same computation; both follow the oracle's model of the mode."""
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    H, W = 32, 48
    torch.manual_seed(0)
    enc, dec = get_network_params(False, H, W, input_channels=9, math="bf16")
    net = ED(False, enc, dec, 0.5, False, input_height=H, input_width=W).to(DEV).eval()
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, 2, 3)).to(DEV)
    torch.manual_seed(5)
    st = [torch.rand(1, *s.shape, device=DEV) for s in O.zero_states(H, W)]
    with torch.no_grad():
        a = net(xs[1][None, None], *st)
        enc_s = net.encoder(xs[1][None, None].permute(1, 0, 2, 3, 4), st[:3])
        feat, dec_s = net.decoder(enc_s, st[3:])
        out = net.head(feat)[:, :, 0]
    for u, v in zip(a, (out, *enc_s, *dec_s)):
        assert torch.equal(u, v)
    w = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in net.state_dict().items()}
    res = O.ed_step(w, xs[1].cpu().numpy().astype(np.float64), [s.cpu().numpy()[0].astype(np.float64) for s in st], quant="bf16")
    for i in range(6):
        d = np.abs(a[1 + i].cpu().numpy()[0] - res["states"][i])
        assert d.max() <= 2e-2 and d.mean() <= 1e-4, (i, d.max(), d.mean())


def test_bf16_full_size_short_horizon_tracks_fp32():
    """BASELINE config 3 shapes (location1 grid 500 x 500, C_in = 63) for 12 steps: bf16/tcgen05 mode vs the fp32 path.
    SURVEY.md 8d tolerance: rms(d depth) <= 2e-2 * rms(depth) (+ margin for the run-to-run spread of rounding ties),
    mask flip rate <= 0.5 %.  Measured on B200: 1.6e-2 and 0 flips; the drift over the full T = 180 horizon (random-init,
    non-contractive weights: 7 % at T = 36, 13 % at T = 180, R^2 = 0.973) is recorded in profiles/r1_config3_drift.json
    (tests/tools/config3_drift.py) and discussed in DESIGN.md section 2."""
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    H = W = 500
    hist, T = 30, 12
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist, seed=42, rain_scale=6.0, rain_max=6.0))
    res = {}
    for m in ("fp32", "bf16"):
        torch.manual_seed(0)
        enc, dec = get_network_params(False, H, W, input_channels=2 * hist + 3, math=m)
        net = ED(False, enc, dec, 0.5, False, input_height=H, input_width=W).to(DEV).eval()
        st = [torch.zeros(1, *s.shape, device=DEV) for s in O.zero_states(H, W)]
        outs = []
        with torch.no_grad():
            for t in range(T):
                out, *st = net(xs[t][None, None].to(DEV), *st)
                outs.append(out[0, 0].cpu())
        res[m] = torch.stack(outs).numpy()
        del net
        torch.cuda.empty_cache()
    d32, d16 = res["fp32"], res["bf16"]
    both = (d32 != 0) & (d16 != 0)
    ratio = np.sqrt(np.mean((d32 - d16)[both] ** 2)) / np.sqrt(np.mean(d32[both] ** 2))
    flips = np.mean((d32 != 0) != (d16 != 0))
    assert ratio <= 3e-2, ratio
    assert flips <= 5e-3, flips
