"""ctypes binding of libspef_b200.so (include/spef_b200.h).

There is deliberately no fallback: if the shared library is missing or a call fails, an exception is
raised.  The product path never runs on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPEF_DEV_LIB") or os.path.join(_PKG_DIR, "libspef_b200.so")   # SPEF_DEV_LIB: developer A/B builds (tools_dev/build_variant.sh)

SPEF_FP32, SPEF_BF16 = 0, 1
FLAG_ORI_NAN, FLAG_POS_ZERO_SUM, FLAG_POS_NAN, FLAG_DOT_GT_1_01, FLAG_ENC_NAN = 1, 2, 4, 8, 16
KIND_STEM, KIND_PW, KIND_DW, KIND_POOL, KIND_HEAD = 0, 1, 2, 3, 4


class SpefError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libspef_b200 error {code}: {msg}")
        self.code = code


class SpefConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "struct_size", "device", "img_h", "img_w", "n_ori", "n_pos", "pos_classification", "precision",
        "max_batch", "pw_impl")]


class SpefTemporalOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "still_ori_soft", "still_pos_soft", "still_quat", "still_pos", "video_ori_soft", "video_pos_soft",
        "video_quat", "video_pos", "ori_distance", "pos_distance", "flags")]


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64

# name -> (restype, argtypes).  Every symbol declared in include/spef_b200.h appears here.
SIGNATURES = {
    "spef_abi_version": (C.c_int, []),
    "spef_create": (C.c_int, [C.POINTER(_vp), C.POINTER(SpefConfig)]),
    "spef_destroy": (None, [_vp]),
    "spef_last_error": (C.c_char_p, [_vp]),
    "spef_load_tensor": (C.c_int, [_vp, C.c_char_p, _vp, C.POINTER(_i64), _i32]),
    "spef_finalize_weights": (C.c_int, [_vp]),
    "spef_set_ori_histogram": (C.c_int, [_vp, _vp, _i32]),
    "spef_set_pos_histogram": (C.c_int, [_vp, _vp, _i32]),
    "spef_set_image_dtype": (C.c_int, [_vp, _i32]),
    "spef_set_host_pack": (C.c_int, [_vp, _i32]),
    "spef_pack_bf16_host": (C.c_int, [_vp, _vp, _i64]),
    "spef_host_pack_info": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), _vp]),
    "spef_resize_frames": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "spef_forward": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp]),
    "spef_num_layers": (C.c_int, [_vp]),
    "spef_layer_info": (C.c_int, [_vp, _i32] + [C.POINTER(_i32)] * 10),
    "spef_layer_forward": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _vp]),
    "spef_num_blocks": (C.c_int, [_vp]),
    "spef_block_info": (C.c_int, [_vp, _i32] + [C.POINTER(_i32)] * 8),
    "spef_set_fusion": (C.c_int, [_vp, _i32]),
    "spef_block_forward": (C.c_int, [_vp, _i32, _vp, _vp, _i32, _vp]),
    "spef_set_stem_fusion": (C.c_int, [_vp, _i32]),
    "spef_stem_fusion_active": (C.c_int, [_vp]),
    "spef_pool_fusion_active": (C.c_int, [_vp]),
    "spef_stem_block_forward": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "spef_decode_ori": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spef_decode_pos": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "spef_score": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "spef_encode_ori": (C.c_int, [_vp, _vp, _i32, _i32, C.c_double, _vp, _vp, _vp, _vp]),
    "spef_encode_pos": (C.c_int, [_vp, _vp, _i32, _i32, C.c_double, _vp, _vp, _vp]),
    "spef_error_stats": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp]),
    "spef_predict": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spef_predict_host": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spef_eval_reset": (C.c_int, [_vp, _vp]),
    "spef_eval_batch_host": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "spef_eval_batch": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "spef_eval_submit_host": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "spef_eval_wait": (C.c_int, [_vp, _vp]),
    "spef_eval_read": (C.c_int, [_vp, _vp, _vp]),
    "spef_eval_sums_dev": (_vp, [_vp]),
    "spef_decode_ori_host": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "spef_decode_pos_host": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "spef_score_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "spef_temporal_reset": (C.c_int, [_vp, _i32, _vp]),
    "spef_temporal_step_logits": (C.c_int, [_vp, _vp, _vp, _i32, C.POINTER(SpefTemporalOut), _vp]),
    "spef_temporal_step": (C.c_int, [_vp, _vp, _i32, _i32, C.POINTER(SpefTemporalOut), _vp]),
    "spef_pdf_filter": (C.c_int, [_vp, _vp, _i32, _i32, _vp, _vp, C.c_float, C.c_float, _vp, _vp, _vp]),
    "spef_launch_count": (_i64, [_vp]),
    "spef_forward_cost": (C.c_int, [_vp, _i32, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "spef_forward_timed": (C.c_int, [_vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "spef_debug_jacobi4_host": (C.c_int, [_vp, _vp, _vp]),
    "spef_debug_decode_solve_host": (C.c_int, [_vp, _i32, _vp, _vp]),
    "spef_debug_resize_taps_host": (C.c_int, [_i32, _i32, _vp, _vp, _vp, _i32, _vp]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libspef_b200.so (once).  Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a).  spef_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.spef_abi_version() != 1:
            raise ImportError("libspef_b200.so ABI version mismatch: rebuild the library")
        _lib = handle
    return _lib


def check(ctx: Optional[int], rc: int) -> None:
    if rc != 0:
        msg = lib().spef_last_error(C.c_void_p(ctx) if ctx else None)
        raise SpefError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")


def ptr(t) -> Optional[int]:
    """Device/host address of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data
