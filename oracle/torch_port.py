"""
torch_port.py -- op-for-op CPU port of the reference's forward path on torch.nn.functional.
TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/urnn_oracle.py for the rules): it is the thing bench.py
times as `cpu_baseline` (kind "port") and as `--impl reference` on the GPU box, where the Python reference
itself is not present.  It issues the same ATen CPU kernels the reference issues (conv2d, group_norm,
layer_norm, conv_transpose2d, avg_pool2d, cat, sigmoid/tanh, elementwise), in the same order, with the same
intermediate tensors, so its timing is representative of the reference's CPU PyTorch path.  Pinned against the
golden vectors by tests/test_oracle_golden.py.  Parameters are addressed by the reference's state_dict keys.
"""
import torch
import torch.nn.functional as F


def cgru_cell(p, prefix, x, hidden, module, nf):
    """CGRU_cell.forward, one step (reference ConvRNN.py:140-190).  x (B,C,H,W) | None, hidden (B,F|2F,H,W)."""
    if x is None:
        cin = p[prefix + "conv1.0.weight"].shape[1] - hidden.shape[1]
        x = torch.zeros(hidden.size(0), cin, hidden.size(2), hidden.size(3), dtype=hidden.dtype)
    pad = (p[prefix + "conv1.0.weight"].shape[-1] - 1) // 2
    combined_1 = torch.cat((x, hidden), 1)
    gates = F.conv2d(combined_1, p[prefix + "conv1.0.weight"], p[prefix + "conv1.0.bias"], 1, pad)
    gates = F.group_norm(gates, 2 * nf // 32, p[prefix + "conv1.1.weight"], p[prefix + "conv1.1.bias"], 1e-5)
    zgate, rgate = torch.split(gates, nf, dim=1)
    z, r = torch.sigmoid(zgate), torch.sigmoid(rgate)
    if module == "encoder":
        combined_2 = torch.cat((x, r * hidden), 1)
        prev = hidden
    else:
        e, d = torch.split(hidden, nf, dim=1)
        combined_2 = torch.cat((x, e, r * d), 1)
        prev = d
    ht = F.conv2d(combined_2, p[prefix + "conv2.0.weight"], p[prefix + "conv2.0.bias"], 1, pad)
    ht = torch.tanh(F.group_norm(ht, nf // 32, p[prefix + "conv2.1.weight"], p[prefix + "conv2.1.bias"], 1e-5))
    return (1 - z) * prev + z * ht


def _base_conv(p, prefix, x):
    y = F.conv2d(x, p[prefix + ".conv.weight"])
    y = F.layer_norm(y, y.shape[1:], p[prefix + ".ln.weight"], p[prefix + ".ln.bias"], 1e-5)
    return F.silu(y)


def ed_step(p, x, states, cls_thred=0.5):
    """ED.forward (reference model.py:65-121) for B=S=1.  x (1,C,H,W); states: 6 tensors (1,C,h,w)."""
    enc = []
    cur = x
    stems = ["encoder.stage1.conv1_leaky_1", "encoder.stage2.conv2_leaky_1", "encoder.stage3.conv3_leaky_1"]
    for k in range(3):
        cur = F.leaky_relu(F.conv2d(cur, p[stems[k] + ".weight"], p[stems[k] + ".bias"]), 0.2)
        if k > 0:
            cur = F.avg_pool2d(cur, 2, 2)
        cur = cgru_cell(p, f"encoder.rnn{k + 1}.", cur, states[k], "encoder", states[k].shape[1])
        enc.append(cur)
    dec = []
    cur = None
    dstems = {3: "decoder.stage3.deconv1_leaky_1", 2: "decoder.stage2.deconv2_leaky_1", 1: "decoder.stage1.conv3_leaky_1"}
    for idx, stage in enumerate((3, 2, 1)):
        hidden = torch.cat((enc[stage - 1], states[3 + idx]), dim=1)
        h = cgru_cell(p, f"decoder.rnn{stage}.", cur, hidden, "decoder", states[3 + idx].shape[1])
        dec.append(h)
        w, b = p[dstems[stage] + ".weight"], p[dstems[stage] + ".bias"]
        cur = F.leaky_relu(F.conv_transpose2d(h, w, b, stride=2) if stage > 1 else F.conv2d(h, w, b), 0.2)
    s = _base_conv(p, "head.stems", cur)
    c = _base_conv(p, "head.cls_convs.1", _base_conv(p, "head.cls_convs.0", s))
    r = _base_conv(p, "head.reg_convs.1", _base_conv(p, "head.reg_convs.0", s))
    prob = torch.sigmoid(F.conv2d(c, p["head.cls_preds.conv.weight"], p["head.cls_preds.conv.bias"]))
    depth = F.leaky_relu(F.conv2d(r, p["head.reg_preds.conv.weight"], p["head.reg_preds.conv.bias"]), 0.2)
    depth = depth * (prob >= cls_thred).float()
    return depth, prob, enc + dec
