"""BASELINE config 3 at full size on the GPU: location1 grid 500 x 500, C_in = 63, T = 180, bf16 (tcgen05) mode vs the
fp32 parity path of the same library (which is pinned to the reference's outputs at smaller sizes).  Prints the drift
statistics SURVEY.md 8d asks for: rms(d depth)/rms(depth), max |d state|, wet/dry mask flip rate, R^2."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
import numpy as np, torch
from oracle import urnn_oracle as O
from src.lib.model.networks.model import ED
from src.lib.model.networks.net_params import get_network_params

H = W = int(os.environ.get("URNN_HW", "500")); T = int(os.environ.get("URNN_T", "180")); hist = 30
C = 2 * hist + 3
dev = "cuda:0"
def build(math):
    torch.manual_seed(0)
    enc, dec = get_network_params(False, H, W, input_channels=C, math=math)
    return ED(False, enc, dec, 0.5, False, input_height=H, input_width=W).to(dev).eval()
xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, hist, seed=42, rain_scale=6.0, rain_max=6.0))
res = {}
for m in ("fp32", "bf16"):
    net = build(m)
    st = [torch.zeros(1, *s.shape, device=dev) for s in O.zero_states(H, W)]
    outs = []
    with torch.no_grad():
        for t in range(T):
            out, *st = net(xs[t][None, None].to(dev), *st)
            outs.append(out[0, 0].cpu())
    res[m] = (torch.stack(outs).numpy(), [s.cpu().numpy() for s in st])
    del net
d32, s32 = res["fp32"]; d16, s16 = res["bf16"]
both = (d32 != 0) & (d16 != 0)
rms_d = float(np.sqrt(np.mean((d32 - d16)[both] ** 2))); rms = float(np.sqrt(np.mean(d32[both] ** 2)))
ss_res = float(np.sum((d32 - d16) ** 2)); ss_tot = float(np.sum((d32 - d32.mean()) ** 2))
out = {"grid": [H, W], "T": T, "rms_ddepth": rms_d, "rms_depth": rms, "ratio": rms_d / max(rms, 1e-30),
       "max_dstate": [float(np.abs(a - b).max()) for a, b in zip(s32, s16)],
       "mask_flip_rate": float(np.mean((d32 != 0) != (d16 != 0))), "wet_fraction_fp32": float(np.mean(d32 != 0)),
       "r2_bf16_vs_fp32": 1.0 - ss_res / max(ss_tot, 1e-30),
       "per_step_ratio_last": float(np.sqrt(np.mean((d32[-1] - d16[-1]) ** 2)) / max(np.sqrt(np.mean(d32[-1] ** 2)), 1e-30))}
def horizon(t):
    a, b = d32[:t], d16[:t]
    m = (a != 0) & (b != 0)
    return {"T": t, "ratio": float(np.sqrt(np.mean((a - b)[m] ** 2)) / max(np.sqrt(np.mean(a[m] ** 2)), 1e-30)),
            "mask_flip_rate": float(np.mean((a != 0) != (b != 0))),
            "r2": 1.0 - float(np.sum((a - b) ** 2)) / max(float(np.sum((a - a.mean()) ** 2)), 1e-30),
            "ratio_at_step": float(np.sqrt(np.mean((a[-1] - b[-1]) ** 2)) / max(np.sqrt(np.mean(a[-1] ** 2)), 1e-30))}
out["horizons"] = [horizon(t) for t in (1, 4, 12, 36, 90, 180) if t <= T]
print(json.dumps(out))
