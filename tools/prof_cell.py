"""The bench's roofline kernel in isolation, for ncu: the full-resolution decoder Skip-ConvGRU cell step (3 launches)
at 500 x 500, run 3 times.  Usage under ncu:
  ncu --set full --import-source on --clock-control none -k "regex:gemm_gn_kernel|cgru_blend" -s 6 -c 3 -f -o gpurun_out/cell python tools/prof_cell.py
(of the matching kernels, launches 0-5 warm up; 6-8 are sweep A, sweep B and the blend of the third step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
import torch
from urnn_b200 import ops
from src.lib.model.networks.ConvRNN import CGRU_cell

H = W = int(os.environ.get("URNN_PROF_HW", "500"))
ops.set_default_math("bf16")
torch.manual_seed(0)
cell = CGRU_cell(False, (H, W), 96, 1, 64, "decoder", math="bf16").cuda().eval()
x = torch.rand(96, H, W, device="cuda"); e = torch.rand(64, H, W, device="cuda")
hs = [torch.rand(64, H, W, device="cuda") for _ in range(3)]
with torch.no_grad():
    for i in range(3):
        cell.step(x, e, hs[i])
torch.cuda.synchronize()
print("done")
