"""T-step runner: the inference / pre-warming loop of the reference (test.py:356-367, main.py:585-592) without
per-step Python, allocation or synchronisation.

`SequenceRunner.run` keeps the six recurrent states resident in two ping-pong sets, writes every step's depth map
into one (T,H,W) device tensor and replays one captured CUDA graph per step parity (the whole encoder-decoder
step is ~30 kernels; at small grids launch overhead dominates without the graph).  `run_host` is the same loop
through the C ABI entry point `urnn_ed_sequence_host`, which takes HOST buffers and overlaps the per-step H2D /
D2H copies with compute.  Both are additive: `ED.forward` keeps the reference's one-step contract.
"""
import ctypes as C

import torch

from . import _capi, ops


def state_shapes(H, W, enc_ch=(64, 96, 96), dec_ch=(96, 96, 64)):
    """Shapes of the 6 states in the reference's order (utils/general.py:50-95): e1,e2,e3,d(1/4),d(1/2),d(1x)."""
    return [(enc_ch[0], H, W), (enc_ch[1], H // 2, W // 2), (enc_ch[2], H // 4, W // 4),
            (dec_ch[0], H // 4, W // 4), (dec_ch[1], H // 2, W // 2), (dec_ch[2], H, W)]


class SequenceRunner:
    def __init__(self, net, H, W, Cin, math=None, use_graph=True, device=None):
        self.net, self.H, self.W, self.Cin = net, H, W, Cin
        self.device = torch.device(device) if device is not None else next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("SequenceRunner needs the model on a CUDA device (no CPU path)")
        self.desc = net.ed_desc(H, W, Cin, math)
        enc_cells, dec_cells = net._cells()
        self.shapes = state_shapes(H, W, [c.num_features for c in enc_cells], [c.num_features for c in dec_cells])
        dev = self.device
        self.states = [[torch.zeros(s, device=dev) for s in self.shapes] for _ in range(2)]
        self._ws = None                      # step workspace of the per-step route, allocated on first use
        self.static_in = torch.zeros((Cin, H, W), device=dev)
        self.static_out = torch.zeros((2, H, W), device=dev)
        self.use_graph = use_graph
        self.graphs = None
        self._params = None

    @property
    def ws(self):
        if self._ws is None:
            self._ws = torch.empty(ops.ed_workspace_bytes(self.desc), dtype=torch.uint8, device=self.device)
        return self._ws

    def _step(self, parity, x, out):
        ops.ed_step_fwd(self.desc, self._params, x, self.states[parity], self.states[parity ^ 1], out, self.ws)

    def _capture(self):
        s = torch.cuda.Stream(device=self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):            # warm-up outside capture (function attributes, lazy module load)
            saved = [[t.clone() for t in grp] for grp in self.states]
            self._step(0, self.static_in, self.static_out)
            self._step(1, self.static_in, self.static_out)
            for grp, sv in zip(self.states, saved):
                for t, v in zip(grp, sv):
                    t.copy_(v)
        torch.cuda.current_stream(self.device).wait_stream(s)
        self.graphs = []
        for parity in (0, 1):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step(parity, self.static_in, self.static_out)
            self.graphs.append(g)

    @torch.no_grad()
    def run(self, inputs, states=None):
        """inputs (T,Cin,H,W) float32 on the device; states: optional list of 6 (C,h,w) or (1,C,h,w) tensors.
        Returns (depth (T,H,W), prob (T,H,W), [6 final states])."""
        T = inputs.shape[0]
        inputs = ops._chk(inputs, "inputs", (T, self.Cin, self.H, self.W))
        self._params = self.net.ed_params()
        for i, dst in enumerate(self.states[0]):
            if states is None:
                dst.zero_()
            else:
                dst.copy_(states[i].reshape(dst.shape))
        depth = torch.empty((T, self.H, self.W), device=self.device)
        prob = torch.empty((T, self.H, self.W), device=self.device)
        if self.use_graph and self.graphs is None:
            self._capture()
        for t in range(T):
            if self.use_graph:
                self.static_in.copy_(inputs[t])
                self.graphs[t & 1].replay()
                depth[t].copy_(self.static_out[0]); prob[t].copy_(self.static_out[1])
            else:
                self._step(t & 1, inputs[t], self.static_out)
                depth[t].copy_(self.static_out[0]); prob[t].copy_(self.static_out[1])
        final = [s.clone() for s in self.states[T & 1]]
        return depth, prob, final

    @torch.no_grad()
    def run_dev(self, inputs, states=None, want_prob=True):
        """The whole loop in ONE library call (urnn_ed_sequence_dev): inputs (T,Cin,H,W) float32 on the device ->
        (depth (T,H,W), prob (T,H,W) or None, [6 final states]); stream-ordered, no host synchronisation."""
        lib = _capi.load()
        T = inputs.shape[0]
        inputs = ops._chk(inputs, "inputs", (T, self.Cin, self.H, self.W))
        st = self.states[0]
        for i, dst in enumerate(st):
            if states is None:
                dst.zero_()
            else:
                dst.copy_(states[i].reshape(dst.shape))
        need = lib.urnn_ed_sequence_dev_workspace_bytes(C.byref(self.desc))
        if need == 0:
            raise RuntimeError("urnn_ed_sequence_dev_workspace_bytes: " + lib.urnn_last_error().decode())
        if getattr(self, "_dev_ws", None) is None or self._dev_ws.numel() < need:
            self._dev_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        depth = torch.empty((T, self.H, self.W), device=self.device)
        prob = torch.empty((T, self.H, self.W), device=self.device) if want_prob else None
        params = self.net.ed_params()
        sp = (C.c_void_p * 6)(*[s.data_ptr() for s in st])
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(lib.urnn_ed_sequence_dev(C.byref(self.desc), C.byref(params), T, C.c_void_p(inputs.data_ptr()),
                                             C.c_void_p(depth.data_ptr()), C.c_void_p(prob.data_ptr()) if want_prob else None, sp,
                                             C.c_void_p(self._dev_ws.data_ptr()), self._dev_ws.numel(), stream), "urnn_ed_sequence_dev")
        return depth, prob, [s.clone() for s in st]

    @torch.no_grad()
    def profile_dev(self, inputs, states=None, max_ops=64):
        """Per-launch mean device time over the T steps of `inputs` (urnn_ed_profile_dev): [(name, milliseconds), ...]."""
        lib = _capi.load()
        T = inputs.shape[0]
        inputs = ops._chk(inputs, "inputs", (T, self.Cin, self.H, self.W))
        st = self.states[0]
        for i, dst in enumerate(st):
            if states is None:
                dst.zero_()
            else:
                dst.copy_(states[i].reshape(dst.shape))
        need = lib.urnn_ed_sequence_dev_workspace_bytes(C.byref(self.desc))
        if getattr(self, "_dev_ws", None) is None or self._dev_ws.numel() < need:
            self._dev_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        ms = (C.c_float * max_ops)()
        names = C.create_string_buffer(24 * max_ops)
        n = C.c_int32(0)
        params = self.net.ed_params()
        sp = (C.c_void_p * 6)(*[s.data_ptr() for s in st])
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(lib.urnn_ed_profile_dev(C.byref(self.desc), C.byref(params), T, C.c_void_p(inputs.data_ptr()), sp,
                                            C.c_void_p(self._dev_ws.data_ptr()), self._dev_ws.numel(), stream,
                                            C.cast(ms, C.c_void_p), C.cast(names, C.c_void_p), max_ops, C.cast(C.pointer(n), C.c_void_p)),
                    "urnn_ed_profile_dev")
        return [(names.raw[24 * i:24 * i + 24].split(b"\0")[0].decode(), float(ms[i])) for i in range(n.value)]

    @torch.no_grad()
    def run_host(self, inputs_host, out_host=None, states=None):
        """inputs_host (T,Cin,H,W) float32 HOST tensor (pin it for overlap) -> out_host (T,H,W) host tensor with the
        masked depth maps; one call into the C ABI (urnn_ed_sequence_host), which blocks until the last copy landed."""
        lib = _capi.load()
        T = inputs_host.shape[0]
        if inputs_host.is_cuda or inputs_host.dtype != torch.float32 or not inputs_host.is_contiguous():
            raise ValueError("run_host: inputs_host must be a contiguous float32 host tensor")
        if out_host is None:
            out_host = torch.empty((T, self.H, self.W), dtype=torch.float32).pin_memory()
        st = self.states[0]
        for i, dst in enumerate(st):
            if states is None:
                dst.zero_()
            else:
                dst.copy_(states[i].reshape(dst.shape))
        need = lib.urnn_ed_sequence_host_workspace_bytes(C.byref(self.desc))
        if getattr(self, "_host_ws", None) is None or self._host_ws.numel() < need:
            self._host_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        params = self.net.ed_params()
        sp = (C.c_void_p * 6)(*[s.data_ptr() for s in st])
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(lib.urnn_ed_sequence_host(C.byref(self.desc), C.byref(params), T, C.c_void_p(inputs_host.data_ptr()),
                                              C.c_void_p(out_host.data_ptr()), sp, C.c_void_p(self._host_ws.data_ptr()),
                                              self._host_ws.numel(), stream), "urnn_ed_sequence_host")
        return out_host, [s.clone() for s in st]

    @torch.no_grad()
    def run_event_host(self, dem, impervious, manhole, rainfall, cumsum_rainfall, hist, rain_max, cumsum_rain_max,
                       dem_min=None, dem_max=None, out_host=None, states=None):
        """One rainfall event as the reference's dataset provides it (raw (H,W) maps and scalar (T,) series on the HOST)
        -> (T,H,W) masked depth on the host, through `urnn_ed_event_host`: no dense per-step input tensor is ever built
        (the 2*hist rainfall channels are folded into a per-step stage-1 bias)."""
        lib = _capi.load()
        T = int(rainfall.shape[0])
        host = lambda t: t.detach().to(torch.float32).contiguous().cpu()
        dem, impervious, manhole, rainfall, cumsum_rainfall = [host(t) for t in (dem, impervious, manhole, rainfall, cumsum_rainfall)]
        if tuple(dem.shape[-2:]) != (self.H, self.W):
            raise ValueError(f"run_event_host: maps must be {self.H}x{self.W}")
        if self.Cin != 2 * hist + 3:
            raise ValueError(f"run_event_host: the model expects C_in={self.Cin}, historical_nums={hist} gives {2 * hist + 3}")
        ev = _capi.EventDesc(T, hist, float(rain_max), float(cumsum_rain_max),
                             float(dem.min() if dem_min is None else dem_min), float(dem.max() if dem_max is None else dem_max))
        if out_host is None:
            out_host = torch.empty((T, self.H, self.W), dtype=torch.float32).pin_memory()
        st = self.states[0]
        for i, dst in enumerate(st):
            if states is None:
                dst.zero_()
            else:
                dst.copy_(states[i].reshape(dst.shape))
        need = lib.urnn_ed_event_host_workspace_bytes(C.byref(self.desc), C.byref(ev))
        if need == 0:
            raise RuntimeError("urnn_ed_event_host_workspace_bytes: " + lib.urnn_last_error().decode())
        if getattr(self, "_event_ws", None) is None or self._event_ws.numel() < need:
            self._event_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        params = self.net.ed_params()
        sp = (C.c_void_p * 6)(*[s.data_ptr() for s in st])
        vp = lambda t: C.c_void_p(t.data_ptr())
        stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _capi.check(lib.urnn_ed_event_host(C.byref(self.desc), C.byref(params), C.byref(ev), vp(dem), vp(impervious), vp(manhole),
                                           vp(rainfall), vp(cumsum_rainfall), vp(out_host), sp, vp(self._event_ws),
                                           self._event_ws.numel(), stream), "urnn_ed_event_host")
        return out_host, [s.clone() for s in st]
