"""Bring-up helper: one cell step in bf16 (tcgen05) mode vs the bf16-quantised oracle; prints error stats."""
import os, sys
os.environ["CUDA_MODULE_LOADING"] = "EAGER"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
import numpy as np, torch, faulthandler
faulthandler.dump_traceback_later(200, exit=True)
from oracle import urnn_oracle as O
from src.lib.model.networks.ConvRNN import CGRU_cell

def run(H, W, cin, F, module, with_x=True, math="bf16"):
    torch.manual_seed(H * 31 + W)
    cell = CGRU_cell(False, (H, W), cin, 1, F, module, math=math).cuda().eval()
    x = (torch.rand(1, 1, cin, H, W) * 2 - 1) if with_x else None
    hid = torch.rand(1, F * (2 if module == "decoder" else 1), H, W) * 2 - 1
    with torch.no_grad():
        out = cell(None if x is None else x.cuda(), hid.cuda(), 1)
    torch.cuda.synchronize()
    w = {k: v.detach().cpu().numpy().astype(np.float64) for k, v in cell.state_dict().items()}
    xn = None if x is None else x.numpy()[:, 0].astype(np.float64)
    refq = O.cgru_cell_forward(w, "", xn, hid.numpy()[0].astype(np.float64), module, F, 1, quant="bf16")[0]
    ref = O.cgru_cell_forward(w, "", xn, hid.numpy()[0].astype(np.float64), module, F, 1)[0]
    o = out.cpu().numpy()[0, 0]
    print(f"{module} {H}x{W} cin={cin} F={F} x={with_x}: vs bf16-oracle max {np.abs(o - refq).max():.2e} mean {np.abs(o - refq).mean():.2e}"
          f" | vs exact max {np.abs(o - ref).max():.2e} mean {np.abs(o - ref).mean():.2e}", flush=True)

if __name__ == "__main__":
    run(16, 16, 16, 64, "encoder")
    run(8, 16, 16, 64, "encoder")
    run(20, 36, 16, 64, "encoder")
    run(64, 64, 96, 64, "decoder")
    run(32, 32, 96, 96, "decoder", with_x=False)
    run(125, 125, 96, 96, "decoder")
    run(250, 250, 64, 96, "encoder")
    run(9, 11, 5, 32, "encoder")
