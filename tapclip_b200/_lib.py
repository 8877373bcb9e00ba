"""ctypes binding of libtapclip.so (the C ABI declared in include/tapclip.h).

There is no fallback: if the shared library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TAPCLIP_LIB") or os.path.join(_HERE, "lib", "libtapclip.so")   # TAPCLIP_LIB: developer override (kernel variants)

ACT = {"gelu_erf": 0, "quick_gelu": 1}
DTYPE = {"fp32": 0, "bf16": 1, "mixed": 2, "fp16": 2}     # engine precision / element type (include/tapclip.h)
ATTR_MODE = {"literal": 0, "intended": 1, "attribution_only": 2}
EPI_ACT, EPI_F32, EPI_F32_ADD = 0, 1, 2
PROBE_NONE, PROBE_TEXT_COL, PROBE_CLS_ROW = 0, 1, 2


class TapclipConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "image_size", "patch_size", "vision_width", "vision_layers", "vision_heads",
        "text_width", "text_layers", "text_heads", "embed_dim", "context_length", "act", "dtype")]


class TapclipError(RuntimeError):
    pass


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol include/tapclip.h declares
PROTOTYPES = {
    "tapclip_create": (C.c_int, [C.POINTER(TapclipConfig), C.POINTER(_vp)]),
    "tapclip_destroy": (C.c_int, [_vp]),
    "tapclip_last_error": (C.c_char_p, []),
    "tapclip_version": (C.c_char_p, []),
    "tapclip_load_weight": (C.c_int, [_vp, C.c_char_p, _vp, _i32, C.POINTER(_i64), _vp]),
    "tapclip_weights_complete": (C.c_int, [_vp]),
    "tapclip_encode_image": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "tapclip_encode_text": (C.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "tapclip_text_forward": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, C.POINTER(_i64), _i64, _i32, _vp]),
    "tapclip_text_gather_config": (C.c_int, [_vp, C.POINTER(_vp), _i32, _i32, _i64]),
    "tapclip_op_gemm_stats_parts": (_i32, [_i64]),
    "tapclip_op_gemm_resid": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _i64, _i64, _i32, _vp]),
    "tapclip_op_gemm_fold": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _vp]),
    "tapclip_op_fold_ln_weight": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _i32, _i32, _vp]),
    "tapclip_op_row_stats_cast": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _i64, _i32, _vp]),
    "tapclip_op_preprocess": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _vp, _vp]),
    "tapclip_logits": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp, _vp, _vp, _vp, _i32, _vp]),
    "tapclip_logits_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp]),
    "tapclip_text_backward": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "tapclip_adamw_step": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i32, _vp]),
    "tapclip_argmax_count": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "tapclip_workspace_bytes": (_i64, [_vp]),
    "tapclip_launch_count": (_i64, [_vp]),
    "tapclip_profile": (C.c_int, [_vp, _i32]),
    "tapclip_profile_report": (C.c_char_p, [_vp]),
    "tapclip_op_gemm": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _vp]),
    "tapclip_op_layernorm": (C.c_int, [_vp, _i64, _vp, _vp, _vp, _i32, _vp, _i64, _i32, _vp]),
    "tapclip_op_layernorm_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp]),
    "tapclip_op_attention": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _i64, _vp]),
    "tapclip_op_attention_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tapclip_op_attention_lse": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "tapclip_op_rollout_step": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "tapclip_op_attribution": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "tapclip_op_cast": (C.c_int, [_vp, _vp, _i32, _i64, _vp]),
}

_lib = None


def load():
    """Load libtapclip.so (built by ``__graft_entry__.build()`` / ``make -C tapclip_b200/csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TapclipError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C tapclip_b200/csrc`). tapclip_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def last_error() -> str:
    return load().tapclip_last_error().decode("utf-8", "replace")


def check(rc: int):
    if rc != 0:
        msg = last_error()
        if "unknown weight" in msg or "unexpected shape" in msg or "must be" in msg:
            raise ValueError(msg)
        raise TapclipError(msg)


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
