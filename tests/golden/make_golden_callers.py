"""
make_golden_callers.py -- golden vectors for the CALLERS of the hot path, produced by the unmodified reference:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_callers.py      (build container only: needs /root/reference)

  preprocess_inputs.npz  Dynamic2DFlood._prepare_input + preprocess_inputs (Dynamic2DFlood.py:179-320) on a synthetic event:
                         the dense (1,1,C,H,W) inputs of a few time steps -- pins oracle.synthetic_event_inputs and the
                         fused event path (urnn_ed_event_host)
  inference_loop.npz     test.py:326-377 Inference (extracted verbatim by oracle/make_ref.py) over the reference's ED on the CPU
  window_loop.npz        main.py:598-692 process_window (pre-warming + 3 gradient steps) + FocalBCE_and_WMSE (losses.py:74-110)
                         + backward: prediction, loss value, gradients
The GPU tests (tests/test_gpu_callers.py) run the SAME extracted loops on top of the drop-in modules.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.join(HERE, "..", "..")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
from oracle import make_ref  # noqa: E402

REF = make_ref.make()
sys.path.insert(0, REF)                       # oracle/_ref: the unmodified reference modules
torch.Tensor.cuda = lambda self, *a, **k: self
import callers  # noqa: E402
from src.lib.dataset.Dynamic2DFlood import Dynamic2DFlood, preprocess_inputs  # noqa: E402
from src.lib.model.networks.losses import select_loss_function  # noqa: E402
from src.lib.model.networks.model import ED  # noqa: E402
from src.lib.model.networks.net_params import get_network_params  # noqa: E402
from src.lib.utils.net_config import load_net_config  # noqa: E402


def synthetic_event(H, W, T, seed=42, rain_scale=6.0):
    """notebook cell-13 recipe (raw arrays as they sit on disk: DEM in metres, rainfall (T,) in mm/step)"""
    rng = np.random.RandomState(seed)
    return {"absolute_DEM": (rng.rand(H, W) * 10.0 / 1000.0).astype(np.float32),     # metres (x1000 -> mm in _prepare_input)
            "impervious": rng.rand(H, W).astype(np.float32),
            "manhole": (rng.rand(H, W) > 0.95).astype(np.float32),
            "rainfall": (rng.rand(T) * rain_scale).astype(np.float32)}


def batch_of(event, T):
    """What the DataLoader (batch 1) hands to the loops: _prepare_input's dict with a leading batch dimension."""
    d = Dynamic2DFlood._prepare_input(None, event, None, duration=T)
    return {k: v.unsqueeze(0) for k, v in d.items()}


def main():
    torch.set_num_threads(8)
    cfg = load_net_config(None)
    # ---- preprocess_inputs
    H, W, T, hist = 16, 12, 8, 3
    ev = synthetic_event(H, W, T)
    inputs = batch_of(ev, T)
    steps = [0, 1, 2, 5, 7]
    dense = np.stack([preprocess_inputs(t, inputs, "cpu", nums=hist, rain_max=6.0, cumsum_rain_max=250.0).numpy()[0, 0] for t in steps])
    np.savez(os.path.join(HERE, "preprocess_inputs.npz"), steps=np.array(steps), dense=dense, meta=np.array([H, W, T, hist]),
             **{"ev." + k: v for k, v in ev.items()})
    print("preprocess_inputs", dense.shape)

    # ---- Inference loop
    H = W = 32; T = 6; hist = 3; C = 2 * hist + 3
    ev = synthetic_event(H, W, T, seed=11, rain_scale=30.0)
    inputs = batch_of(ev, T)
    torch.manual_seed(0)
    p = get_network_params(False, H, W, input_channels=C, net_cfg=cfg)
    net = ED(False, p[0], p[1], 0.5, False, input_height=H, input_width=W)
    out = callers.Inference(net, inputs, "cpu", historical_nums=hist, rain_max=60.0, cumsum_rain_max=250.0,
                            input_height=H, input_width=W, net_cfg=cfg)
    np.savez(os.path.join(HERE, "inference_loop.npz"), out=out, meta=np.array([H, W, T, hist]),
             w_fingerprint=np.array([float(v.double().sum()) for v in net.state_dict().values()]),
             **{"ev." + k: v for k, v in ev.items()})
    print("inference_loop", out.shape, float(np.abs(out).mean()))

    # ---- SWP window: pre-warming to ind, seq_num gradient steps, the reference's loss, backward
    H = W = 16; T = 6; hist = 3; C = 2 * hist + 3
    ev = synthetic_event(H, W, T, seed=5, rain_scale=30.0)
    inputs = batch_of(ev, T)
    rng = np.random.RandomState(7)
    flood = rng.rand(T, H, W) * 0.3
    flood[flood < 0.25] = 0
    label = torch.from_numpy((flood * 1000.0 / 5000.0).astype(np.float32))[None]          # (1,T,H,W), MinMaxScaler(mm, 5000, 0)
    torch.manual_seed(0)
    p = get_network_params(False, H, W, input_channels=C, net_cfg=cfg)
    net = ED(False, p[0], p[1], 0.5, False, input_height=H, input_width=W).train()
    opt = torch.optim.Adam(net.parameters(), lr=0.01)
    args = argparse.Namespace(input_height=H, input_width=W, net_cfg=cfg, prewarming=True, historical_nums=hist,
                              rain_max=60.0, cumsum_rain_max=250.0, seq_num=3)
    ind = 2
    pred, final_states = callers.process_window(ind, args, net, inputs, "cpu", opt)
    lossf = select_loss_function("FocalBCE_and_WMSE", "mean")
    losses = lossf(pred, label[:, ind:ind + args.seq_num], 0)
    losses["loss"].backward()
    d = {"reg": pred["reg"].detach().numpy()[0], "cls": pred["cls"].numpy()[0], "loss": np.array(float(losses["loss"])),
         "label": label.numpy()[0], "meta": np.array([H, W, T, hist, ind, args.seq_num]),
         "w_fingerprint": np.array([float(v.double().sum()) for v in net.state_dict().values()])}
    for i, s in enumerate(final_states[0] + final_states[1]):
        d[f"final{i}"] = s.numpy()[0]
    for k, v in net.named_parameters():
        d["g." + k] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    d.update({"ev." + k: v for k, v in ev.items()})
    np.savez(os.path.join(HERE, "window_loop.npz"), **d)
    print("window_loop loss", float(losses["loss"]), "grads", len([k for k in d if k.startswith("g.")]))


if __name__ == "__main__":
    main()
