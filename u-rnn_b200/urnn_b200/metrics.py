"""Device-side evaluation metrics of an inference run.

Host-side mirror of the reference's post-processing, test.py:468 (`r_MinMaxScaler(output, max=flood_max, min=0)`) followed
by test.py:607-675 `compute_metrics(pred_mm, gt_mm, flood_thres)`, for predictions that are still on the GPU: the (T, H, W)
result never has to be copied to the host and de-normalised there (24 GB at 4096 x 4096 x 360).  Same keys, same units:

    m = StreamingMetrics(H, W, T, flood_max=5000.0, flood_thres=150.0, device="cuda:0")
    for t0 in range(0, T, chunk):
        m.update(pred_norm[t0:t0 + chunk], gt_mm[t0:t0 + chunk])        # CUDA tensors, fp32, (n, H, W)
    m.result()      # {"R2", "MSE", "RMSE", "MAE", "PeakR2", "CSI"}  (+ "tp", "fp", "fn", "t_peak" under .detail)

The arithmetic lives in csrc/metrics.cu (urnn_metrics_* in include/urnn_b200.h); there is no PyTorch fallback."""
import torch

from . import _capi

KEYS = ("R2", "MSE", "RMSE", "MAE", "PeakR2", "CSI")


class StreamingMetrics:
    def __init__(self, H, W, T, flood_max=5000.0, flood_thres=150.0, device="cuda:0"):
        self.H, self.W, self.T = int(H), int(W), int(T)
        self.flood_max, self.flood_thres = float(flood_max), float(flood_thres)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("StreamingMetrics runs on a CUDA device only (urnn_b200 has no CPU path)")
        self.lib = _capi.load()
        n = self.lib.urnn_metrics_workspace_bytes(self.H, self.W, self.T)
        if n == 0:
            raise ValueError(f"invalid metric shape H={H} W={W} T={T}")
        self.ws = torch.empty(n, dtype=torch.uint8, device=self.device)
        self.out = torch.zeros(12, dtype=torch.float64, device=self.device)
        self.next_t = 0
        self.detail = {}
        self.reset()

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def reset(self):
        with torch.cuda.device(self.device):
            _capi.check(self.lib.urnn_metrics_reset(self.H, self.W, self.T, self.ws.data_ptr(), self.ws.numel(), self._stream()),
                        "urnn_metrics_reset")
        self.next_t = 0

    def update(self, pred_norm, gt_mm):
        """pred_norm: (n, H, W) normalised model output; gt_mm: (n, H, W) ground truth in mm; consecutive chunks of the event."""
        for name, t in (("pred_norm", pred_norm), ("gt_mm", gt_mm)):
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f"{name} must be a contiguous fp32 tensor on {self.device}")
            if t.dim() != 3 or tuple(t.shape[1:]) != (self.H, self.W):
                raise ValueError(f"{name} has shape {tuple(t.shape)}, expected (n, {self.H}, {self.W})")
        n = pred_norm.shape[0]
        if gt_mm.shape[0] != n or self.next_t + n > self.T:
            raise ValueError(f"chunk of {n} steps at t={self.next_t} does not fit an event of {self.T} steps")
        with torch.cuda.device(self.device):
            _capi.check(self.lib.urnn_metrics_accumulate(self.H, self.W, self.T, self.next_t, n, pred_norm.data_ptr(), gt_mm.data_ptr(),
                                                         self.flood_max, self.ws.data_ptr(), self.ws.numel(), self._stream()),
                        "urnn_metrics_accumulate")
        self.next_t += n

    def result(self):
        """The reference's metric dictionary (test.py:664-671); one 96-byte device-to-host copy."""
        if self.next_t != self.T:
            raise RuntimeError(f"only {self.next_t} of {self.T} time steps were accumulated")
        with torch.cuda.device(self.device):
            _capi.check(self.lib.urnn_metrics_finalize(self.H, self.W, self.T, self.flood_thres, self.ws.data_ptr(), self.ws.numel(),
                                                       self.out.data_ptr(), self._stream()), "urnn_metrics_finalize")
        o = self.out.cpu().tolist()
        self.detail = {"tp": int(o[6]), "fp": int(o[7]), "fn": int(o[8]), "t_peak": int(o[9]), "elements": int(o[10])}
        return dict(zip(KEYS, o[:6]))
