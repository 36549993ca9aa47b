"""Torch-facing operators over the C ABI (include/urnn_b200.h).

Each op checks its tensors (CUDA, fp32, contiguous), allocates outputs/workspace through torch's
caching allocator and enqueues the library call on torch's current stream.  Autograd support is
provided by the *Fn classes: their backward calls the library's backward entry points.
CPU tensors are rejected: there is no fallback path.
"""
import ctypes as C

import torch

from . import _capi
from ._capi import CellDesc, CellParams, CellGrads, HeadParams, HeadGrads, EdDesc, EdParams

LRELU_SLOPE = 0.2   # reference utils.py:63
GN_EPS = 1e-5       # nn.GroupNorm default (reference ConvRNN.py:97,103)
LN_EPS = 1e-5       # nn.LayerNorm default (reference head/network_blocks.py:94)

_default_math = "fp32"


def set_default_math(name):
    """'fp32' (FFMA parity mode), 'f16x3' (tcgen05, fp16 hi+lo split operands: the fast mode) or 'bf16' (tcgen05,
    single-pass bf16: short horizons only) for modules that do not pin one."""
    global _default_math
    if name not in _capi.MATH_BY_NAME:
        raise ValueError(f"unknown math mode {name!r}")
    _default_math = name


def get_default_math():
    return _default_math


def _math_code(math):
    return _capi.MATH_BY_NAME[math or _default_math]


def _chk(t, name, shape=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: urnn_b200 runs on CUDA tensors only (got {t.device}); there is no CPU fallback")
    if t.dtype != torch.float32:
        raise TypeError(f"{name}: expected float32, got {t.dtype}")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name}: expected shape {tuple(shape)}, got {tuple(t.shape)}")
    return t if t.is_contiguous() else t.contiguous()


def _empty_like_c(t):
    """A CONTIGUOUS tensor of t's shape (empty_like preserves dense non-contiguous strides, the library expects NCHW)."""
    return torch.empty(t.shape, dtype=t.dtype, device=t.device)


def _c(t):
    return None if t is None else (t if t.is_contiguous() else t.contiguous())


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def cell_params_struct(w1, b1, g1w, g1b, w2, b2, g2w, g2b):
    return CellParams(*[_p(t) for t in (w1, b1, g1w, g1b, w2, b2, g2w, g2b)])


# ---------------------------------------------------------------------------------------------- ConvGRU cell
def cgru_cell_fwd(x, e, h, params, ksize, variant, math=None, eps=GN_EPS):
    """One (Skip-)ConvGRU step (reference ConvRNN.py:140-190).  x (Cx,H,W) or None; e (F,H,W) or None;
    h (F,H,W); params = the cell's 8 tensors.  Returns the new state (F,H,W)."""
    lib = _capi.load()
    F, H, W = h.shape
    w1 = params[0]
    Ch = 2 * F if variant == _capi.URNN_CELL_DECODER else F
    Cx = w1.shape[1] - Ch
    h = _chk(h, "h")
    x = _chk(x, "x", (Cx, H, W)) if x is not None else None
    e = _chk(e, "e", (F, H, W)) if e is not None else None
    shapes = [(2 * F, Cx + Ch, ksize, ksize), (2 * F,), (2 * F,), (2 * F,), (F, Cx + Ch, ksize, ksize), (F,), (F,), (F,)]
    params = [_chk(t, f"cell parameter {i}", s) for i, (t, s) in enumerate(zip(params, shapes))]
    desc = CellDesc(H, W, Cx, F, ksize, variant, _math_code(math), eps)
    out = torch.empty_like(h)
    ws = _ws(lib.urnn_cgru_fwd_workspace_bytes(C.byref(desc)), h.device)
    cp = cell_params_struct(*params)
    _capi.check(lib.urnn_cgru_fwd(C.byref(desc), C.byref(cp), _p(x), _p(e), _p(h), _p(out),
                                  _p(ws), ws.numel(), _stream()), "urnn_cgru_fwd")
    return out


def cgru_cell_bwd(x, e, h, dh_out, params, grads, ksize, variant, need, math=None, eps=GN_EPS):
    """Backward of one cell step.  grads: 8 fp32 tensors that are accumulated into.  need = (dx?, de?, dh?)."""
    lib = _capi.load()
    F, H, W = h.shape
    Ch = 2 * F if variant == _capi.URNN_CELL_DECODER else F
    Cx = params[0].shape[1] - Ch
    desc = CellDesc(H, W, Cx, F, ksize, variant, _math_code(math), eps)
    dh_out = _chk(dh_out, "dh_out", (F, H, W))
    dx = _empty_like_c(x) if (x is not None and need[0]) else None
    de = _empty_like_c(e) if (e is not None and need[1]) else None
    dh = _empty_like_c(h) if need[2] else None
    ws = _ws(lib.urnn_cgru_bwd_workspace_bytes(C.byref(desc)), h.device)
    cp = cell_params_struct(*params)
    cg = CellGrads(*[_p(t) for t in grads])
    _capi.check(lib.urnn_cgru_bwd(C.byref(desc), C.byref(cp), _p(x), _p(e), _p(h), _p(dh_out), _p(dx), _p(de), _p(dh),
                                  C.byref(cg), _p(ws), ws.numel(), _stream()), "urnn_cgru_bwd")
    return dx, de, dh


class CgruCellFn(torch.autograd.Function):
    """autograd wrapper: forward saves only its inputs (recompute-in-backward, like the reference's
    reentrant checkpointing, ConvRNN.py:154-158)."""

    @staticmethod
    def forward(ctx, x, e, h, ksize, variant, math, *params):
        ctx.cfg = (ksize, variant, math)
        x, e, h = _c(x), _c(e), _c(h)            # the backward hands raw pointers of the saved tensors to the library
        params = tuple(_c(p) for p in params)
        ctx.save_for_backward(*[t for t in (x, e, h) if t is not None], *params)
        ctx.has = (x is not None, e is not None)
        with torch.no_grad():
            return cgru_cell_fwd(x, e, h, [p.detach() for p in params], ksize, variant, math)

    @staticmethod
    def backward(ctx, dh_out):
        saved = list(ctx.saved_tensors)
        x = saved.pop(0) if ctx.has[0] else None
        e = saved.pop(0) if ctx.has[1] else None
        h = saved.pop(0)
        params = saved
        ksize, variant, math = ctx.cfg
        grads = [torch.zeros_like(p) for p in params]
        need = (ctx.has[0] and ctx.needs_input_grad[0], ctx.has[1] and ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        dx, de, dh = cgru_cell_bwd(x, e, h, dh_out.contiguous(), params, grads, ksize, variant, need, math)
        return (dx, de, dh, None, None, None, *grads)


def cgru_cell(x, e, h, params, ksize, variant, math=None):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (x, e, h, *params)):
        return CgruCellFn.apply(x, e, h, ksize, variant, math, *params)
    return cgru_cell_fwd(x, e, h, [p.detach() for p in params], ksize, variant, math)


# ---------------------------------------------------------------------------------------------- stems
def conv1x1_lrelu_fwd(x, w, b, pool=1, slope=LRELU_SLOPE, math=None):
    """[AvgPool2](LeakyReLU(conv1x1(x))) (reference utils.py:85-121).  x (Cin,H,W) -> (Cout,H/pool,W/pool)."""
    lib = _capi.load()
    x = _chk(x, "x")
    Cin, H, W = x.shape
    Cout = w.shape[0]
    w = _chk(w, "w", (Cout, Cin, 1, 1)); b = _chk(b, "b", (Cout,))
    y = torch.empty((Cout, H // pool, W // pool), dtype=torch.float32, device=x.device)
    _capi.check(lib.urnn_conv1x1_lrelu_fwd(Cin, Cout, H, W, pool, slope, _math_code(math), _p(x), _p(w), _p(b), _p(y), _stream()),
                "urnn_conv1x1_lrelu_fwd")
    return y


class Conv1x1LreluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, pool, slope, math):
        x, w, b = _c(x), _c(w), _c(b)
        ctx.cfg = (pool, slope)
        ctx.save_for_backward(x, w, b)
        with torch.no_grad():
            return conv1x1_lrelu_fwd(x, w.detach(), b.detach(), pool, slope, math)

    @staticmethod
    def backward(ctx, dy):
        lib = _capi.load()
        x, w, b = ctx.saved_tensors
        pool, slope = ctx.cfg
        Cin, H, W = x.shape
        Cout = w.shape[0]
        dy = _chk(dy, "dy", (Cout, H // pool, W // pool))
        dx = _empty_like_c(x) if ctx.needs_input_grad[0] else None
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        ws = _ws(lib.urnn_conv1x1_lrelu_bwd_workspace_bytes(Cin, Cout, H, W, pool), x.device)
        _capi.check(lib.urnn_conv1x1_lrelu_bwd(Cin, Cout, H, W, pool, slope, _p(x), _p(w), _p(b), _p(dy), _p(dx),
                                               _p(dw), _p(db), _p(ws), ws.numel(), _stream()), "urnn_conv1x1_lrelu_bwd")
        return dx, dw, db, None, None, None


def conv1x1_lrelu(x, w, b, pool=1, slope=LRELU_SLOPE, math=None):
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or b.requires_grad):
        return Conv1x1LreluFn.apply(x, w, b, pool, slope, math)
    return conv1x1_lrelu_fwd(x, w.detach(), b.detach(), pool, slope, math)


def deconv2x2_lrelu_fwd(x, w, b, slope=LRELU_SLOPE, math=None):
    """LeakyReLU(ConvTranspose2d(k=2,s=2)(x)) (reference utils.py:95-107).  x (Cin,H,W), w (Cin,Cout,2,2)."""
    lib = _capi.load()
    x = _chk(x, "x")
    Cin, H, W = x.shape
    Cout = w.shape[1]
    w = _chk(w, "w", (Cin, Cout, 2, 2)); b = _chk(b, "b", (Cout,))
    y = torch.empty((Cout, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
    _capi.check(lib.urnn_deconv2x2_lrelu_fwd(Cin, Cout, H, W, slope, _math_code(math), _p(x), _p(w), _p(b), _p(y), _stream()),
                "urnn_deconv2x2_lrelu_fwd")
    return y


class Deconv2x2LreluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, slope, math):
        x, w, b = _c(x), _c(w), _c(b)
        ctx.slope = slope
        ctx.save_for_backward(x, w, b)
        with torch.no_grad():
            return deconv2x2_lrelu_fwd(x, w.detach(), b.detach(), slope, math)

    @staticmethod
    def backward(ctx, dy):
        lib = _capi.load()
        x, w, b = ctx.saved_tensors
        Cin, H, W = x.shape
        Cout = w.shape[1]
        dy = _chk(dy, "dy", (Cout, 2 * H, 2 * W))
        dx = _empty_like_c(x) if ctx.needs_input_grad[0] else None
        dw, db = torch.zeros_like(w), torch.zeros_like(b)
        ws = _ws(lib.urnn_deconv2x2_lrelu_bwd_workspace_bytes(Cin, Cout, H, W), x.device)
        _capi.check(lib.urnn_deconv2x2_lrelu_bwd(Cin, Cout, H, W, ctx.slope, _p(x), _p(w), _p(b), _p(dy), _p(dx),
                                                 _p(dw), _p(db), _p(ws), ws.numel(), _stream()), "urnn_deconv2x2_lrelu_bwd")
        return dx, dw, db, None, None


def deconv2x2_lrelu(x, w, b, slope=LRELU_SLOPE, math=None):
    if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad or b.requires_grad):
        return Deconv2x2LreluFn.apply(x, w, b, slope, math)
    return deconv2x2_lrelu_fwd(x, w.detach(), b.detach(), slope, math)


# ---------------------------------------------------------------------------------------------- head
def head_params_struct(p):
    """p: dict with conv_w[5], ln_w[5], ln_b[5], cls_pred_w/b, reg_pred_w/b tensors."""
    hp = HeadParams()
    for i in range(5):
        hp.conv_w[i] = p["conv_w"][i].data_ptr()
        hp.ln_w[i] = p["ln_w"][i].data_ptr()
        hp.ln_b[i] = p["ln_b"][i].data_ptr()
    hp.cls_pred_w = p["cls_pred_w"].data_ptr(); hp.cls_pred_b = p["cls_pred_b"].data_ptr()
    hp.reg_pred_w = p["reg_pred_w"].data_ptr(); hp.reg_pred_b = p["reg_pred_b"].data_ptr()
    return hp


HEAD_ORDER = ["conv_w", "ln_w", "ln_b"]


def _head_flat(p):
    return [*p["conv_w"], *p["ln_w"], *p["ln_b"], p["cls_pred_w"], p["cls_pred_b"], p["reg_pred_w"], p["reg_pred_b"]]


def _head_dict(flat):
    return {"conv_w": flat[0:5], "ln_w": flat[5:10], "ln_b": flat[10:15],
            "cls_pred_w": flat[15], "cls_pred_b": flat[16], "reg_pred_w": flat[17], "reg_pred_b": flat[18]}


def head_fwd(feat, p, cls_thred, ln_eps=LN_EPS, slope=LRELU_SLOPE):
    """feat (16,H,W) -> (2,H,W): [masked depth, wet probability] (reference head/flood_head.py:131-202)."""
    lib = _capi.load()
    feat = _chk(feat, "feat")
    C16, H, W = feat.shape
    if C16 != 16:
        raise ValueError(f"head: expected 16 feature channels, got {C16}")
    p = dict(p)                                  # the checked (contiguous) tensors are the ones whose pointers are passed
    p["conv_w"] = [_chk(t, "head conv weight", (16, 16, 1, 1)) for t in p["conv_w"]]
    p["ln_w"] = [_chk(t, "head ln weight", (16, H, W)) for t in p["ln_w"]]
    p["ln_b"] = [_chk(t, "head ln bias", (16, H, W)) for t in p["ln_b"]]
    out = torch.empty((2, H, W), dtype=torch.float32, device=feat.device)
    ws = _ws(lib.urnn_head_fwd_workspace_bytes(H, W), feat.device)
    hp = head_params_struct(p)
    _capi.check(lib.urnn_head_fwd(H, W, cls_thred, ln_eps, slope, C.byref(hp), _p(feat), _p(out), _p(ws), ws.numel(),
                                  _stream()), "urnn_head_fwd")
    return out


class HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, cls_thred, *flat):
        feat = _c(feat)
        flat = tuple(_c(t) for t in flat)
        ctx.cls_thred = cls_thred
        ctx.save_for_backward(feat, *flat)
        with torch.no_grad():
            return head_fwd(feat, _head_dict([t.detach() for t in flat]), cls_thred)

    @staticmethod
    def backward(ctx, dout):
        lib = _capi.load()
        feat, *flat = ctx.saved_tensors
        _, H, W = feat.shape
        dout = _chk(dout, "dout", (2, H, W))
        dfeat = _empty_like_c(feat)
        gflat = [torch.zeros_like(t) for t in flat]
        hp = head_params_struct(_head_dict(flat))
        hg = HeadGrads()
        gd = _head_dict(gflat)
        for i in range(5):
            hg.conv_w[i] = gd["conv_w"][i].data_ptr(); hg.ln_w[i] = gd["ln_w"][i].data_ptr(); hg.ln_b[i] = gd["ln_b"][i].data_ptr()
        hg.cls_pred_w = gd["cls_pred_w"].data_ptr(); hg.cls_pred_b = gd["cls_pred_b"].data_ptr()
        hg.reg_pred_w = gd["reg_pred_w"].data_ptr(); hg.reg_pred_b = gd["reg_pred_b"].data_ptr()
        ws = _ws(lib.urnn_head_bwd_workspace_bytes(H, W), feat.device)
        _capi.check(lib.urnn_head_bwd(H, W, ctx.cls_thred, LN_EPS, LRELU_SLOPE, C.byref(hp), _p(feat), _p(dout), _p(dfeat),
                                      C.byref(hg), _p(ws), ws.numel(), _stream()), "urnn_head_bwd")
        return (dfeat, None, *gflat)


def head(feat, p, cls_thred):
    flat = _head_flat(p)
    if torch.is_grad_enabled() and (feat.requires_grad or any(t.requires_grad for t in flat)):
        return HeadFn.apply(feat, cls_thred, *flat)
    return head_fwd(feat, _head_dict([t.detach() for t in flat]), cls_thred)


# ---------------------------------------------------------------------------------------------- whole step
def make_ed_desc(H, W, Cin, enc_conv, enc_gru, dec_gru, dec_conv, cls_thred, math=None, ksize=1):
    d = EdDesc()
    d.H, d.W, d.Cin = H, W, Cin
    for i in range(3):
        d.enc_conv[i] = enc_conv[i]; d.enc_gru[i] = enc_gru[i]
        d.dec_gru[i] = dec_gru[i]; d.dec_conv[i] = dec_conv[i]
    d.ksize = ksize
    d.math = _math_code(math)
    d.cls_thred, d.gn_eps, d.ln_eps, d.lrelu_slope = cls_thred, GN_EPS, LN_EPS, LRELU_SLOPE
    return d


def ed_workspace_bytes(desc):
    n = _capi.load().urnn_ed_step_workspace_bytes(C.byref(desc))
    if n == 0:
        raise RuntimeError("urnn_ed_step_workspace_bytes: " + _capi.load().urnn_last_error().decode())
    return n


def ed_step_fwd(desc, params, x, states_in, states_out, out, ws):
    """Enqueue one whole encoder-decoder step (reference model.py:65-121).  All buffers preallocated."""
    lib = _capi.load()
    sin = (C.c_void_p * 6)(*[s.data_ptr() for s in states_in])
    sout = (C.c_void_p * 6)(*[s.data_ptr() for s in states_out])
    _capi.check(lib.urnn_ed_step_fwd(C.byref(desc), C.byref(params), _p(x), sin, sout, _p(out), _p(ws), ws.numel(),
                                     _stream()), "urnn_ed_step_fwd")
