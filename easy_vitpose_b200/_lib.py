"""ctypes binding of libvitpose_b200.so (C ABI declared in include/vitpose_b200.h).

The library is the product; this module only declares argument types and turns non-zero return codes
into RuntimeError.  A missing library is a hard error: there is no Python / CPU fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB, build, is_stale

_lib = None


class VpbConfig(C.Structure):
    _fields_ = [("embed_dim", C.c_int32), ("depth", C.c_int32), ("num_heads", C.c_int32),
                ("num_keypoints", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32)]


EXPORTS = {
    # name: (restype, argtypes)
    "vpb_last_error": (C.c_char_p, []),
    "vpb_create": (C.c_int, [C.POINTER(VpbConfig), C.POINTER(C.c_void_p)]),
    "vpb_destroy": (None, [C.c_void_p]),
    "vpb_load_tensor": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "vpb_finalize": (C.c_int, [C.c_void_p]),
    "vpb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vpb_forward_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vpb_decode": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "vpb_head": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vpb_flip_back": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "vpb_decode_modes": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpb_decode_modes_ex": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "vpb_infer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpb_infer_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpb_submit_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "vpb_wait_host": (C.c_int, [C.c_void_p, C.c_int32]),
    "vpb_preprocess": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpb_decode_frame": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "vpb_infer_frame": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpb_frame_status": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32)]),
    "vpb_infer_frame_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpb_submit_frame_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "vpb_host_alloc": (C.c_void_p, [C.c_int64]),
    "vpb_host_free": (None, [C.c_void_p]),
    "vpb_kernel_launches": (C.c_int, [C.c_void_p, C.c_int32]),
    "vpb_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "vpb_profile_classes": (C.c_int, []),
    "vpb_profile_class_name": (C.c_char_p, [C.c_int32]),
    "vpb_profile_collect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vpb_read_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int64]),
    "vpb_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                           C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "vpb_debug_gemm": (C.c_int, [C.c_int32, C.c_void_p]),
    "vpb_attention": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "vpb_debug_attention": (C.c_int, [C.c_int32]),
    "vpb_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
}


def lib():
    """Loads the shared library once.  Raises if it has not been built (python -m easy_vitpose_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or is_stale():
            try:
                build()                              # nvcc cross-compiles sm_100a anywhere; seconds
            except Exception as exc:
                # never load a library older than its sources: its ABI / kernels may no longer match the header and
                # the argtypes below.  VPB_ALLOW_STALE=1 is the explicit escape hatch (e.g. a box without nvcc).
                if not os.path.exists(LIB) or os.environ.get("VPB_ALLOW_STALE", "0") != "1":
                    raise RuntimeError(f"{LIB} is missing or older than its sources and could not be rebuilt ({exc}); build it "
                                       "with `python -m easy_vitpose_b200.build`; there is no fallback implementation") from exc
                import warnings
                warnings.warn(f"loading a STALE {LIB} (VPB_ALLOW_STALE=1): {exc}")
        handle = C.CDLL(LIB)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(handle, name)          # AttributeError here = header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(code: int) -> None:
    if code != 0:
        raise RuntimeError(f"vitpose_b200 error {code}: {lib().vpb_last_error().decode()}")


def check_value(code: int) -> None:
    """Like check(), but a bad ARGUMENT (code 1: e.g. a box that is empty after clipping, where the reference raises from
    pad_image / cv2.resize) surfaces as ValueError."""
    if code == 1:
        raise ValueError(lib().vpb_last_error().decode())
    check(code)
