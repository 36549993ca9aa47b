import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from oracle import urnn_oracle as O
from test_gpu_x3 import build_ed, zero_states, DEV
from urnn_b200.runner import SequenceRunner
H, W, hist, T = 40, 24, 3, 5
net = build_ed(H, W, 2 * hist + 3)
xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist)).to(DEV)
torch.manual_seed(3)
st0 = [torch.rand_like(s) - 0.5 for s in zero_states(H, W)]
for trial in range(2):
    st = list(st0); outs = []
    with torch.no_grad():
        for t in range(T):
            out, *st = net(xs[t][None, None], *st)
            outs.append(out[0, 0].clone())
    if trial == 0: outs0 = outs
    else: print("step loop repeat max diff", [float((a - b).abs().max()) for a, b in zip(outs, outs0)])
run = SequenceRunner(net, H, W, 2 * hist + 3)
for T2 in (1, 2, 3, 5):
    depth, prob, fin = run.run_dev(xs[:T2], states=[s[0] for s in st0])
    print("T", T2, "per-step max diff seq vs loop:", [float((depth[t] - outs0[t]).abs().max()) for t in range(T2)])
d1, _, _ = run.run_dev(xs, states=[s[0] for s in st0]); d2, _, _ = run.run_dev(xs, states=[s[0] for s in st0])
print("seq repeat max diff", float((d1 - d2).abs().max()))
