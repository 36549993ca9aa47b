"""Per-launch time table of one f16x3 encoder-decoder step (urnn_ed_profile_dev: CUDA events around every launch)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
import torch
from bench import build_net
from urnn_b200.runner import SequenceRunner
H = int(os.environ.get("URNN_H", "500")); W = int(os.environ.get("URNN_W", "500")); hist = 30; C = 2 * hist + 3
T = int(os.environ.get("URNN_T", "12"))
dev = torch.device("cuda:0")
net = build_net(H, W, C, "f16x3", dev)
run = SequenceRunner(net, H, W, C)
torch.manual_seed(1)
xs = torch.rand(T, C, H, W, device=dev)
run.run_dev(xs[:4], want_prob=False)
tab = run.profile_dev(xs)
tot = sum(ms for _, ms in tab)
for name, ms in tab:
    print(f"{name:14s} {ms * 1e3:8.1f} us")
print(f"{'sum':14s} {tot * 1e3:8.1f} us")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record(); run.run_dev(xs, want_prob=False); e1.record(); torch.cuda.synchronize()
print(f"run_dev: {e0.elapsed_time(e1) / T * 1e3:.1f} us/step (incl. state conversion at both ends) -> {H * W * T / (e0.elapsed_time(e1) * 1e-3) / 1e6:.1f} M cells*steps/s")
if len(sys.argv) > 1:
    json.dump({"grid": [H, W], "T": T, "ops": tab}, open(sys.argv[1], "w"))
