"""ctypes binding of liburnn_b200.so -- mirrors include/urnn_b200.h one to one.

The library is the product: importing this module without a built liburnn_b200.so raises, and so does
every op when handed anything but CUDA tensors.  There is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("URNN_B200_LIB", os.path.join(HERE, "liburnn_b200.so"))   # override: bring-up only

URNN_CELL_ENCODER, URNN_CELL_DECODER = 0, 1
MATH_FP32, MATH_BF16, MATH_F16X3 = 0, 2, 3
# fp32: FFMA parity mode | f16x3: tcgen05, fp16 hi+lo split operands (the fast mode that meets the T=180 tolerance) |
# bf16: tcgen05 single-pass bf16 operands (round-1 kernels; short horizons only)
MATH_BY_NAME = {"fp32": MATH_FP32, "bf16": MATH_BF16, "f16x3": MATH_F16X3}

fp = C.c_void_p   # device pointers travel as integers


class CellDesc(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("Cx", C.c_int32), ("F", C.c_int32),
                ("ksize", C.c_int32), ("variant", C.c_int32), ("math", C.c_int32), ("eps", C.c_float)]


class CellParams(C.Structure):
    _fields_ = [(n, fp) for n in ("w1", "b1", "gn1_w", "gn1_b", "w2", "b2", "gn2_w", "gn2_b")]


class CellGrads(C.Structure):
    _fields_ = [(n, fp) for n in ("w1", "b1", "gn1_w", "gn1_b", "w2", "b2", "gn2_w", "gn2_b")]


class HeadParams(C.Structure):
    _fields_ = [("conv_w", fp * 5), ("ln_w", fp * 5), ("ln_b", fp * 5),
                ("cls_pred_w", fp), ("cls_pred_b", fp), ("reg_pred_w", fp), ("reg_pred_b", fp)]


class HeadGrads(C.Structure):
    _fields_ = HeadParams._fields_


class EdDesc(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("Cin", C.c_int32),
                ("enc_conv", C.c_int32 * 3), ("enc_gru", C.c_int32 * 3),
                ("dec_gru", C.c_int32 * 3), ("dec_conv", C.c_int32 * 3),
                ("ksize", C.c_int32), ("math", C.c_int32),
                ("cls_thred", C.c_float), ("gn_eps", C.c_float), ("ln_eps", C.c_float), ("lrelu_slope", C.c_float)]


class EventDesc(C.Structure):
    _fields_ = [("T", C.c_int32), ("hist", C.c_int32), ("rain_max", C.c_float), ("cumsum_rain_max", C.c_float),
                ("dem_min", C.c_float), ("dem_max", C.c_float)]


class EdParams(C.Structure):
    _fields_ = [("enc_stem_w", fp * 3), ("enc_stem_b", fp * 3),
                ("enc_cell", CellParams * 3), ("dec_cell", CellParams * 3),
                ("dec_stem_w", fp * 3), ("dec_stem_b", fp * 3),
                ("head", HeadParams)]


# name -> (restype, argtypes); every symbol include/urnn_b200.h declares
i32, f32, sz, vp = C.c_int32, C.c_float, C.c_size_t, C.c_void_p
SIGNATURES = {
    "urnn_abi_version": (C.c_int, []),
    "urnn_last_error": (C.c_char_p, []),
    "urnn_launch_count": (C.c_uint64, []),
    "urnn_cgru_fwd_workspace_bytes": (sz, [C.POINTER(CellDesc)]),
    "urnn_cgru_fwd": (C.c_int, [C.POINTER(CellDesc), C.POINTER(CellParams), fp, fp, fp, fp, vp, sz, vp]),
    "urnn_cgru_bwd_workspace_bytes": (sz, [C.POINTER(CellDesc)]),
    "urnn_cgru_bwd": (C.c_int, [C.POINTER(CellDesc), C.POINTER(CellParams), fp, fp, fp, fp, fp, fp, fp,
                                C.POINTER(CellGrads), vp, sz, vp]),
    "urnn_conv1x1_lrelu_fwd": (C.c_int, [i32, i32, i32, i32, i32, f32, i32, fp, fp, fp, fp, vp]),
    "urnn_conv1x1_lrelu_bwd_workspace_bytes": (sz, [i32, i32, i32, i32, i32]),
    "urnn_conv1x1_lrelu_bwd": (C.c_int, [i32, i32, i32, i32, i32, f32, fp, fp, fp, fp, fp, fp, fp, vp, sz, vp]),
    "urnn_deconv2x2_lrelu_fwd": (C.c_int, [i32, i32, i32, i32, f32, i32, fp, fp, fp, fp, vp]),
    "urnn_deconv2x2_lrelu_bwd_workspace_bytes": (sz, [i32, i32, i32, i32]),
    "urnn_deconv2x2_lrelu_bwd": (C.c_int, [i32, i32, i32, i32, f32, fp, fp, fp, fp, fp, fp, fp, vp, sz, vp]),
    "urnn_head_fwd_workspace_bytes": (sz, [i32, i32]),
    "urnn_head_fwd": (C.c_int, [i32, i32, f32, f32, f32, C.POINTER(HeadParams), fp, fp, vp, sz, vp]),
    "urnn_head_bwd_workspace_bytes": (sz, [i32, i32]),
    "urnn_head_bwd": (C.c_int, [i32, i32, f32, f32, f32, C.POINTER(HeadParams), fp, fp, fp,
                                C.POINTER(HeadGrads), vp, sz, vp]),
    "urnn_ed_step_workspace_bytes": (sz, [C.POINTER(EdDesc)]),
    "urnn_ed_step_fwd": (C.c_int, [C.POINTER(EdDesc), C.POINTER(EdParams), fp, C.POINTER(fp), C.POINTER(fp), fp,
                                   vp, sz, vp]),
    "urnn_ed_sequence_dev_workspace_bytes": (sz, [C.POINTER(EdDesc)]),
    "urnn_ed_sequence_dev": (C.c_int, [C.POINTER(EdDesc), C.POINTER(EdParams), i32, vp, vp, vp, C.POINTER(fp),
                                       vp, sz, vp]),
    "urnn_layout_index": (C.c_int64, [i32, i32, i32, i32, i32, C.POINTER(C.c_int64)]),
    "urnn_ed_profile_dev": (C.c_int, [C.POINTER(EdDesc), C.POINTER(EdParams), i32, vp, C.POINTER(fp), vp, sz, vp,
                                      vp, vp, i32, vp]),
    "urnn_ed_sequence_host_workspace_bytes": (sz, [C.POINTER(EdDesc)]),
    "urnn_ed_sequence_host": (C.c_int, [C.POINTER(EdDesc), C.POINTER(EdParams), i32, vp, vp, C.POINTER(fp),
                                        vp, sz, vp]),
    "urnn_ed_event_host_workspace_bytes": (sz, [C.POINTER(EdDesc), C.POINTER(EventDesc)]),
    "urnn_ed_event_host": (C.c_int, [C.POINTER(EdDesc), C.POINTER(EdParams), C.POINTER(EventDesc), vp, vp, vp, vp, vp, vp,
                                     C.POINTER(fp), vp, sz, vp]),
    "urnn_metrics_workspace_bytes": (sz, [i32, i32, i32]),
    "urnn_metrics_reset": (C.c_int, [i32, i32, i32, vp, sz, vp]),
    "urnn_metrics_accumulate": (C.c_int, [i32, i32, i32, i32, i32, vp, vp, f32, vp, sz, vp]),
    "urnn_metrics_finalize": (C.c_int, [i32, i32, i32, f32, vp, sz, vp, vp]),
    "urnn_comm_local_init": (C.c_int, [i32, i32, vp]),
    "urnn_comm_connect": (C.c_int, [vp]),
    "urnn_comm_destroy": (C.c_int, []),
    "urnn_comm_world": (C.c_int, []),
}

_lib = None


def load():
    """Load (once) and return the ctypes handle; raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"liburnn_b200.so not found at {LIB_PATH}: build it first "
            "(python -c 'import __graft_entry__ as g; g.build()' or python u-rnn_b200/urnn_b200/build.py). "
            "urnn_b200 has no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is missing: loud by design
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().urnn_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
