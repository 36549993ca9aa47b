"""
urnn_oracle.py -- CPU restatement (numpy) of the U-RNN hot path.  TEST INFRASTRUCTURE ONLY.

This file is the *checker*, never the product: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product path
(u-rnn_b200/) never imports anything from oracle/ and fails loudly without its CUDA library.

Parity status: the reference ships no golden vectors or known-answer tests for this path
(SURVEY.md section 8c, F10).  The oracle is therefore pinned against outputs of the
reference itself, generated in the build container by tests/golden/make_golden.py (which
imports the unmodified reference modules from /root/reference/code) and committed under
tests/golden/*.npz.  tests/test_oracle_golden.py checks every function below against them.

Each function cites the reference lines it restates (paths relative to
/root/reference/code/src/lib/model/networks/).  Parameters are addressed by the
reference's own state_dict key names so that a reference checkpoint feeds the oracle
unchanged.  All arrays are (C, H, W) -- the reference's batch (B=1) and sequence (S=1)
dimensions are dropped.

dtype: every function computes in the dtype of its inputs; pass float64 arrays for an
error-budget run, float32 to mimic the reference's arithmetic width.
"""
import numpy as np

# Spatial-sharding model (tests only): when set, a callable that sums an array of local (sum, sumsq, count) triples
# over all ranks (e.g. torch.distributed.all_reduce on gloo).  Statistics are the only cross-band quantity for 1x1
# filters; everything else in this file is pixel-local or aligned to 4-row bands.
STATS_ALLREDUCE = None

GN_EPS = 1e-5      # torch.nn.GroupNorm default, ConvRNN.py:97,103
LN_EPS = 1e-5      # torch.nn.LayerNorm default, head/network_blocks.py:94
LRELU_SLOPE = 0.2  # utils.py:63, head/network_blocks.py:39


# --------------------------------------------------------------------------- primitives
def round_bf16(a):
    """Round to the nearest bfloat16 (ties to even), returned in the input dtype.  Used only to model the
    URNN_MATH_BF16 mode of the CUDA path (operands of the gate contractions are rounded to bf16, products
    and sums stay exact/fp32): the reference itself has no reduced-precision mode."""
    f = np.ascontiguousarray(a, dtype=np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).astype(a.dtype).reshape(a.shape)


def conv2d_same(x, w, b=None, quant=None):
    """nn.Conv2d(stride 1, padding (k-1)//2) for odd k.  ConvRNN.py:95-96,101-102 (cells),
    utils.py:110-114 (stems, k=1), head/network_blocks.py:84-92,151-152.
    x (Cin,H,W), w (Cout,Cin,k,k), b (Cout,) or None -> (Cout,H,W)."""
    cout, cin, kh, kw = w.shape
    c, h, wd = x.shape
    assert c == cin and kh == kw and kh % 2 == 1
    if quant == "bf16":
        x, w = round_bf16(x), round_bf16(w)
    if kh == 1:
        y = (w[:, :, 0, 0] @ x.reshape(cin, h * wd)).reshape(cout, h, wd)
    else:
        p = (kh - 1) // 2
        xp = np.zeros((cin, h + 2 * p, wd + 2 * p), dtype=x.dtype)
        xp[:, p:p + h, p:p + wd] = x
        y = np.zeros((cout, h * wd), dtype=x.dtype)
        for dy in range(kh):
            for dx in range(kw):
                patch = xp[:, dy:dy + h, dx:dx + wd].reshape(cin, h * wd)
                y += w[:, :, dy, dx] @ patch
        y = y.reshape(cout, h, wd)
    if b is not None:
        y = y + b[:, None, None]
    return y


def group_norm(x, num_groups, gamma, beta, eps=GN_EPS, stats_from=None):
    """nn.GroupNorm over (C/G channels x H x W), biased variance.  ConvRNN.py:97,103.
    stats_from (bf16 storage model only): take mean/variance from this array instead of x."""
    c, h, w = x.shape
    g = x.reshape(num_groups, -1)
    gs = g if stats_from is None else stats_from.reshape(num_groups, -1)
    if STATS_ALLREDUCE is not None:
        tri = np.stack([gs.sum(axis=1), (gs ** 2).sum(axis=1), np.full(num_groups, gs.shape[1], dtype=gs.dtype)])
        tri = STATS_ALLREDUCE(tri.astype(np.float64))
        mean = (tri[0] / tri[2])[:, None].astype(x.dtype)
        var = (tri[1] / tri[2] - (tri[0] / tri[2]) ** 2)[:, None].astype(x.dtype)
    else:
        mean = gs.mean(axis=1, keepdims=True)
        var = ((gs - mean) ** 2).mean(axis=1, keepdims=True)
    y = ((g - mean) / np.sqrt(var + x.dtype.type(eps))).reshape(c, h, w)
    return y * gamma[:, None, None] + beta[:, None, None]


def layer_norm_chw(x, weight, bias, eps=LN_EPS):
    """nn.LayerNorm([C,H,W]) with per-element affine.  head/network_blocks.py:93-94."""
    if STATS_ALLREDUCE is not None:
        tri = STATS_ALLREDUCE(np.array([[x.sum()], [(x ** 2).sum()], [x.size]], dtype=np.float64))
        mean = x.dtype.type(tri[0, 0] / tri[2, 0])
        var = x.dtype.type(tri[1, 0] / tri[2, 0] - (tri[0, 0] / tri[2, 0]) ** 2)
    else:
        mean = x.mean()
        var = ((x - mean) ** 2).mean()
    return (x - mean) / np.sqrt(var + x.dtype.type(eps)) * weight + bias


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def silu(x):
    return x * sigmoid(x)


def leaky_relu(x, slope=LRELU_SLOPE):
    return np.where(x >= 0, x, x * x.dtype.type(slope))


def avg_pool2(x):
    """nn.AvgPool2d(2, 2, 0): floor semantics.  utils.py:92-94."""
    c, h, w = x.shape
    h2, w2 = h // 2, w // 2
    v = x[:, :h2 * 2, :w2 * 2].reshape(c, h2, 2, w2, 2)
    return v.mean(axis=(2, 4)).astype(x.dtype)


def conv_transpose2x2(x, w, b, quant=None):
    """nn.ConvTranspose2d(k=2, s=2, p=0): non-overlapping 2x upsample.  utils.py:95-100.
    x (Cin,H,W), w (Cin,Cout,2,2) -> (Cout,2H,2W)."""
    cin, cout, kh, kw = w.shape
    c, h, wd = x.shape
    assert c == cin and kh == 2 and kw == 2
    if quant == "bf16":
        x, w = round_bf16(x), round_bf16(w)
    y = np.zeros((cout, 2 * h, 2 * wd), dtype=x.dtype)
    xf = x.reshape(cin, h * wd)
    for dy in range(2):
        for dx in range(2):
            y[:, dy::2, dx::2] = (w[:, :, dy, dx].T @ xf).reshape(cout, h, wd)
    return y + b[:, None, None]


# --------------------------------------------------------------------------- ConvGRU cell
def cgru_cell_step(p, prefix, x, hidden, module, num_features, quant=None):
    """One time step of CGRU_cell.forward, ConvRNN.py:140-190.

    p[prefix + 'conv1.0.weight'] etc. are the cell's 8 tensors (ConvRNN.py:94-104).
    encoder: hidden = h (F,H,W); decoder: hidden = cat(e, d) (2F,H,W); x may be None
    (decoder stage 3, ConvRNN.py:143-146 -> zeros).  Returns the new state (F,H,W).
    Also returns the intermediates (pre-GN gate map, pre-GN candidate map) for per-op tests.
    """
    F = num_features
    w1, b1 = p[prefix + "conv1.0.weight"], p[prefix + "conv1.0.bias"]
    g1w, g1b = p[prefix + "conv1.1.weight"], p[prefix + "conv1.1.bias"]
    w2, b2 = p[prefix + "conv2.0.weight"], p[prefix + "conv2.0.bias"]
    g2w, g2b = p[prefix + "conv2.1.weight"], p[prefix + "conv2.1.bias"]
    x_was_none = x is None
    if x is None:
        cin = w1.shape[1] - hidden.shape[0]
        x = np.zeros((cin,) + hidden.shape[1:], dtype=hidden.dtype)
    combined_1 = np.concatenate([x, hidden], axis=0)                 # ConvRNN.py:153
    gates_pre = conv2d_same(combined_1, w1, b1, quant)
    if quant == "bf16":
        # storage model of URNN_MATH_BF16: the pre-GN maps are kept as bf16, statistics come from the fp32 values
        gates = group_norm(round_bf16(gates_pre), (2 * F) // 32, g1w, g1b, stats_from=gates_pre)
    else:
        gates = group_norm(gates_pre, (2 * F) // 32, g1w, g1b)       # ConvRNN.py:97
    z = sigmoid(gates[:F])                                           # ConvRNN.py:160-162
    r = sigmoid(gates[F:])
    if module == "encoder":
        combined_2 = np.concatenate([x, r * hidden], axis=0)         # ConvRNN.py:167
        prev = hidden
    else:
        e, d = hidden[:F], hidden[F:]                                # ConvRNN.py:172
        combined_2 = np.concatenate([x, e, r * d], axis=0)           # ConvRNN.py:173
        prev = d
    kxe = combined_2.shape[0] - F                                    # channels not touched by the reset gate
    kskip = x.shape[0] if x_was_none else 0                          # zero-input columns are skipped by the kernel
    if quant == "bf16" and 3 * F <= 256 and kxe - kskip > 0:
        # the r-independent part of the candidate is produced by the first sweep and stored as bf16
        part = round_bf16(conv2d_same(combined_2[:kxe], w2[:, :kxe], None, quant))
        cand_pre = part + conv2d_same(combined_2[kxe:], w2[:, kxe:], b2, quant)
    else:
        cand_pre = conv2d_same(combined_2, w2, b2, quant)
    if quant == "bf16":
        cand = np.tanh(group_norm(round_bf16(cand_pre), F // 32, g2w, g2b, stats_from=cand_pre))
    else:
        cand = np.tanh(group_norm(cand_pre, F // 32, g2w, g2b))      # ConvRNN.py:103,180
    h_next = (1 - z) * prev + z * cand                               # ConvRNN.py:185,189
    return h_next.astype(hidden.dtype), gates_pre, cand_pre


def cgru_cell_forward(p, prefix, inputs, hidden, module, num_features, seq_len=1, quant=None):
    """CGRU_cell.forward over seq_len steps, ConvRNN.py:111-194: the encoder feeds h_t
    back as the next hidden state (ConvRNN.py:192); returns the stack (S,F,H,W).
    NB for the decoder variant the reference assigns the F-channel output to htprev, so
    seq_len>1 fails there in the reference as well; the oracle only accepts seq_len==1."""
    outs = []
    if module == "decoder":
        assert seq_len == 1
    for t in range(seq_len):
        x = None if inputs is None else inputs[t]
        hidden, _, _ = cgru_cell_step(p, prefix, x, hidden, module, num_features, quant)
        outs.append(hidden)
    return np.stack(outs)


# --------------------------------------------------------------------------- encoder / decoder
ENC_STEM_KEYS = ["encoder.stage1.conv1_leaky_1", "encoder.stage2.conv2_leaky_1",
                 "encoder.stage3.conv3_leaky_1"]
DEC_STEM_KEYS = {3: "decoder.stage3.deconv1_leaky_1", 2: "decoder.stage2.deconv2_leaky_1",
                 1: "decoder.stage1.conv3_leaky_1"}


def encoder_step(p, x, enc_states, down_factors=(1, 2, 2), quant=None):
    """Encoder.forward / forward_by_stage for S=B=1, encoder.py:119-215.
    Stage k: 1x1 conv + LeakyReLU(0.2) [+ AvgPool2 AFTER the conv] -> ConvGRU
    (utils.py:85-121 layer order; net_params.py:82-88)."""
    new_states = []
    cur = x
    for k in range(3):
        key = ENC_STEM_KEYS[k]
        cur = leaky_relu(conv2d_same(cur, p[key + ".weight"], p[key + ".bias"], quant))
        if down_factors[k] > 1:
            cur = avg_pool2(cur)
        F = enc_states[k].shape[0]
        h, _, _ = cgru_cell_step(p, f"encoder.rnn{k + 1}.", cur, enc_states[k], "encoder", F, quant)
        new_states.append(h)
        cur = h
    return new_states


def decoder_step(p, enc_states, dec_states, quant=None):
    """Decoder.forward / forward_by_stage, decoder.py:102-217.  dec_states are ordered
    deepest-first (stage 3, 2, 1; decoder.py:194-212).  Returns (features (16,H,W), new states)."""
    new_states = []
    cur = None
    for idx, stage in enumerate((3, 2, 1)):
        e = enc_states[stage - 1]
        d = dec_states[idx]
        F = d.shape[0]
        hidden = np.concatenate([e, d], axis=0)                      # decoder.py:135
        h, _, _ = cgru_cell_step(p, f"decoder.rnn{stage}.", cur, hidden, "decoder", F, quant)
        new_states.append(h)
        key = DEC_STEM_KEYS[stage]
        w, b = p[key + ".weight"], p[key + ".bias"]
        if "deconv" in key:
            cur = leaky_relu(conv_transpose2x2(h, w, b, quant))      # utils.py:95-107
        else:
            cur = leaky_relu(conv2d_same(h, w, b, quant))
    return cur, new_states


# --------------------------------------------------------------------------- head
def _base_conv(p, prefix, x):
    """BaseConv: 1x1 conv (no bias) -> LayerNorm([C,H,W]) -> SiLU.  network_blocks.py:74-101."""
    y = conv2d_same(x, p[prefix + ".conv.weight"], None)
    return silu(layer_norm_chw(y, p[prefix + ".ln.weight"], p[prefix + ".ln.bias"]))


def head_forward(p, feat, cls_thred=0.5):
    """YOLOXHead.forward + correction_depth, head/flood_head.py:131-202.
    Returns (masked depth (H,W), wet probability (H,W), unmasked depth (H,W))."""
    s = _base_conv(p, "head.stems", feat)
    c = _base_conv(p, "head.cls_convs.1", _base_conv(p, "head.cls_convs.0", s))
    r = _base_conv(p, "head.reg_convs.1", _base_conv(p, "head.reg_convs.0", s))
    prob = sigmoid(conv2d_same(c, p["head.cls_preds.conv.weight"], p["head.cls_preds.conv.bias"]))[0]
    depth = leaky_relu(conv2d_same(r, p["head.reg_preds.conv.weight"], p["head.reg_preds.conv.bias"]))[0]
    mask = (prob >= prob.dtype.type(cls_thred)).astype(depth.dtype)  # flood_head.py:201
    return depth * mask, prob, depth


# --------------------------------------------------------------------------- whole step
def ed_step(p, x, states, cls_thred=0.5, quant=None):
    """ED.forward for one time step, model.py:65-121.
    states = [e1,e2,e3,d(1/4),d(1/2),d(1x)] (general.py:50-95 order).
    Returns dict(out=(H,W) masked depth, prob, depth_raw, states=[6 new states])."""
    enc = encoder_step(p, x, states[:3], quant=quant)
    feat, dec = decoder_step(p, enc, states[3:], quant=quant)
    out, prob, raw = head_forward(p, feat, cls_thred)
    return {"out": out, "prob": prob, "depth_raw": raw, "states": list(enc) + list(dec), "feat": feat}


def zero_states(H, W, dtype=np.float32, enc_ch=(64, 96, 96), dec_ch=(96, 96, 64)):
    """initialize_states, utils/general.py:50-95 / utils/net_config.py:59-116."""
    sc = (1, 2, 4)
    enc = [np.zeros((enc_ch[k], H // sc[k], W // sc[k]), dtype=dtype) for k in range(3)]
    dec = [np.zeros((dec_ch[k], H // sc[2 - k], W // sc[2 - k]), dtype=dtype) for k in range(3)]
    return enc + dec


def run_sequence(p, inputs, states=None, cls_thred=0.5, quant=None):
    """The inference loop of test.py:356-367 without I/O: inputs (T,C,H,W) -> (T,H,W)."""
    T, _, H, W = inputs.shape
    if states is None:
        states = zero_states(H, W, inputs.dtype)
    outs = []
    for t in range(T):
        res = ed_step(p, inputs[t], states, cls_thred, quant)
        states = res["states"]
        outs.append(res["out"])
    return np.stack(outs), states


def synthetic_event_inputs(H, W, T, hist, seed=42, rain_scale=30.0, rain_max=60.0,
                           cumsum_rain_max=250.0):
    """Synthetic event + per-step input assembly: the notebook cell-13 recipe followed by
    dataset/Dynamic2DFlood.py:265-366 (preprocess_inputs, get_past_rainfall, MinMaxScaler)
    for scalar rainfall.  Returns (T, 2*hist+3, H, W) float32."""
    rng = np.random.RandomState(seed)
    dem = rng.rand(H, W) * 10.0
    imperv = rng.rand(H, W)
    manhole = (rng.rand(H, W) > 0.95).astype(np.float64)
    rain = rng.rand(T) * rain_scale
    cum = np.cumsum(rain)
    nd = (dem - dem.min()) / (dem.max() - dem.min())
    ni = (imperv - 0.05) / (0.95 - 0.05)
    nm = manhole
    out = np.zeros((T, 2 * hist + 3, H, W), dtype=np.float32)
    for t in range(T):
        s0 = max(0, t - hist + 1)
        n = t + 1 - s0
        out[t, hist - n:hist] = (rain[s0:t + 1] / rain_max)[:, None, None]
        out[t, 2 * hist - n:2 * hist] = (cum[s0:t + 1] / cumsum_rain_max)[:, None, None]
        out[t, 2 * hist] = nd
        out[t, 2 * hist + 1] = ni
        out[t, 2 * hist + 2] = nm
    return out
