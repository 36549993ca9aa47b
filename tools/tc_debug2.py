import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
import torch, faulthandler
faulthandler.dump_traceback_later(30, exit=True)
from urnn_b200 import _capi
lib = _capi.load()
def p(t): return C.c_void_p(t.data_ptr())
F, Cx, H, W = 64, 16, 16, 16
dev = "cuda"
x = torch.rand(Cx, H, W, device=dev); h = torch.rand(F, H, W, device=dev); out = torch.empty_like(h)
w1 = torch.randn(2 * F, Cx + F, 1, 1, device=dev) * 0.1; b1 = torch.zeros(2 * F, device=dev)
g1w = torch.ones(2 * F, device=dev); g1b = torch.zeros(2 * F, device=dev)
w2 = torch.randn(F, Cx + F, 1, 1, device=dev) * 0.1; b2 = torch.zeros(F, device=dev)
g2w = torch.ones(F, device=dev); g2b = torch.zeros(F, device=dev)
cp = _capi.CellParams(*[p(t) for t in (w1, b1, g1w, g1b, w2, b2, g2w, g2b)])
for math in (0, 2):
    d = _capi.CellDesc(H, W, Cx, F, 1, 0, math, 1e-5)
    n = lib.urnn_cgru_fwd_workspace_bytes(C.byref(d))
    ws = torch.empty(n, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    print("calling math", math, "ws", n, flush=True)
    rc = lib.urnn_cgru_fwd(C.byref(d), C.byref(cp), p(x), None, p(h), p(out), p(ws), n, None)
    print("returned", rc, lib.urnn_last_error(), flush=True)
    torch.cuda.synchronize()
    print("synced", float(out.sum()), flush=True)
