"""
make_golden.py -- generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference, CPU only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference ships no fixtures of its own, so these files are what pins the
oracle (oracle/urnn_oracle.py) and, through it, the CUDA path.  The reference modules are
imported from /root/reference/code with one shim: torch.Tensor.cuda is made a no-op because the
reference hard-codes .cuda() (ConvRNN.py:136,146; decoder.py:132) and this box has no GPU.
Nothing here is used at test time on the GPU box; only the .npz outputs travel.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference/code"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.dont_write_bytecode = True


def import_reference():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    from src.lib.model.networks.ConvRNN import CGRU_cell
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    from src.lib.utils.net_config import get_state_shapes, load_net_config
    from src.lib.model.networks.losses import select_loss_function
    return CGRU_cell, ED, get_network_params, get_state_shapes, load_net_config, select_loss_function


def unique_state(sd):
    """Drop the aliased keys: keep the first key of each storage."""
    seen, out = {}, {}
    for k, v in sd.items():
        key = v.data_ptr()
        if key not in seen:
            seen[key] = k
            out[k] = v.detach().numpy().copy()
    return out


def main():
    CGRU_cell, ED, get_network_params, get_state_shapes, load_net_config, select_loss_function = import_reference()
    torch.set_num_threads(8)

    # ---- config 1: single cells ------------------------------------------
    def cell_case(name, k, module, cin, F, hw, seq_len, with_x=True):
        torch.manual_seed(0)
        cell = CGRU_cell(False, hw, cin, k, F, module).eval()
        torch.manual_seed(1)
        x = torch.rand(seq_len, 1, cin, *hw) if with_x else None
        if module == "encoder":
            hidden = torch.zeros(1, F, *hw)
        else:
            hidden = torch.rand(1, 2 * F, *hw)
        with torch.no_grad():
            out = cell(x, hidden, seq_len)
            # second call from a non-zero state exercises r*h with h != 0
            hidden2 = torch.rand(1, F, *hw) if module == "encoder" else hidden
            out2 = cell(x[:1] if with_x else None, hidden2, 1)
        d = {"w." + k_: v for k_, v in unique_state(cell.state_dict()).items()}
        o = out.numpy()[:, 0]
        if hw[0] * hw[1] > 2048:     # keep the big config-1 fixture small: last step only
            d.update(out_last=o[-1], out_chan_mean=o.mean(axis=(2, 3)))
        else:
            d.update(out=o, out2=out2.numpy()[0, 0], hidden2=hidden2.numpy()[0])
        d["meta"] = np.array([k, cin, F, hw[0], hw[1], seq_len, int(with_x)])
        np.savez(os.path.join(HERE, name + ".npz"), **d)
        print(name, out.shape, float(out.abs().mean()))


    cell_case("cell_enc_k1", 1, "encoder", 16, 64, (64, 64), 4)          # config 1 proper
    cell_case("cell_dec_k1_nox", 1, "decoder", 96, 64, (32, 32), 1, with_x=False)
    cell_case("cell_dec_k1", 1, "decoder", 96, 64, (32, 32), 1)
    cell_case("cell_enc_k3", 3, "encoder", 16, 64, (32, 32), 2)
    cell_case("cell_dec_k3", 3, "decoder", 32, 32, (24, 20), 1)

    # ---- ED: state_dict key list + small full-step fixtures ------------------------------
    cfg = load_net_config(None)
    from oracle.urnn_oracle import synthetic_event_inputs

    def ed_case(name, H, W, hist, T, keep_weights, rain_scale=30.0, rain_max=60.0, every=1):
        C = 2 * hist + 3
        torch.manual_seed(0)
        p = get_network_params(False, H, W, input_channels=C, net_cfg=cfg)
        net = ED(False, p[0], p[1], 0.5, False, input_height=H, input_width=W).eval()
        st = [torch.zeros(s) for s in get_state_shapes(cfg, H, W)]
        xs = synthetic_event_inputs(H, W, T, hist, seed=42, rain_scale=rain_scale, rain_max=rain_max)
        taps = {}
        net.head.reg_preds.register_forward_hook(lambda m, i, o: taps.__setitem__("q", o.detach().numpy()[0, 0].copy()))
        net.head.cls_preds.register_forward_hook(lambda m, i, o: taps.__setitem__("p", o.detach().numpy()[0, 0].copy()))
        outs, probs, raws = [], [], []
        with torch.no_grad():
            for t in range(T):
                out, *st = net(torch.from_numpy(xs[t])[None, None], *st)
                outs.append(out.numpy()[0, 0].copy()); probs.append(taps["p"]); raws.append(taps["q"])
        d = dict(out=np.stack(outs)[::every], prob=np.stack(probs)[::every], depth_raw=np.stack(raws)[::every],
                 meta=np.array([H, W, hist, T, every]))
        for i, s in enumerate(st):
            a = s.numpy()[0]
            if H * W > 4096:         # big grids: strided sample + per-channel means
                d[f"state{i}_s4"] = a[:, ::4, ::4].copy()
                d[f"state{i}_mean"] = a.mean(axis=(1, 2))
            else:
                d[f"state{i}"] = a
        if keep_weights:
            d.update({"w." + k: v for k, v in unique_state(net.state_dict()).items()})
        else:
            # weights are reproducible from torch.manual_seed(0) + the module construction order;
            # keep a fingerprint so the test can prove it rebuilt the same ones
            d["w_fingerprint"] = np.array([float(v.double().sum()) for v in net.state_dict().values()])
        np.savez(os.path.join(HERE, name + ".npz"), **d)
        print(name, d["out"].shape, float(np.abs(d["out"]).mean()), float((d["prob"] >= 0.5).mean()))
        return net

    net = ed_case("ed_32x32_c9", 32, 32, 3, 4, keep_weights=True)
    with open(os.path.join(HERE, "state_dict_keys.txt"), "w") as f:
        for k, v in net.state_dict().items():
            f.write(f"{k} {tuple(v.shape)}\n")
    ed_case("ed_24x40_c63", 24, 40, 30, 3, keep_weights=False, rain_scale=6.0, rain_max=6.0)
    # config 2 (lite, 128x128, C_in=9, T=36): outputs every 6th step; weights by seed
    ed_case("ed_lite128", 128, 128, 3, 36, keep_weights=False, every=6)

    # ---- config 4 pin: loss + gradients of a 3-step window at 16x16 ----------------------
    H = W = 16; hist = 3; C = 9; T = 3
    torch.manual_seed(0)
    p = get_network_params(False, H, W, input_channels=C, net_cfg=cfg)
    net = ED(False, p[0], p[1], 0.5, False, input_height=H, input_width=W).train()
    xs = synthetic_event_inputs(H, W, T, hist, seed=42)
    rng = np.random.RandomState(7)
    flood = rng.rand(T, H, W) * 0.3
    flood[flood < 0.25] = 0
    label = torch.from_numpy((flood * 1000.0 / 5000.0).astype(np.float32))[None]       # (1,T,H,W)
    torch.manual_seed(3)
    st = [torch.rand(s).mul_(0.5).requires_grad_(True) for s in get_state_shapes(cfg, H, W)]
    st0 = [s for s in st]
    regs = []
    cur = st
    for t in range(T):
        out, *cur = net(torch.from_numpy(xs[t])[None, None], *cur)
        regs.append(out)
    reg = torch.cat(regs, dim=1)                                                        # (1,T,H,W)
    # the loss itself (losses.py) is outside the hot path; a plain MSE on the depth output pins
    # d(depth)/d(parameters, incoming states) through 3 steps of BPTT (main.py:674-684,756-757)
    loss_plain = ((reg - label) ** 2).mean()
    loss_plain.backward()
    d = {"loss": np.array(float(loss_plain)), "reg": reg.detach().numpy()[0], "label": label.numpy()[0],
         "meta": np.array([H, W, hist, T])}
    for i, s in enumerate(st0):
        d[f"state{i}"] = s.detach().numpy()[0]
        d[f"gstate{i}"] = s.grad.numpy()[0]
    d["w_fingerprint"] = np.array([float(v.double().sum()) for v in net.state_dict().values()])
    for k, v in net.named_parameters():
        d["g." + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    np.savez(os.path.join(HERE, "ed_grad_16x16.npz"), **d)
    print("grad", float(loss_plain), len([k for k in d if k.startswith("g.")]))


if __name__ == "__main__":
    main()
