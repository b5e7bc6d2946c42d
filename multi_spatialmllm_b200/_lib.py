"""ctypes binding of libmsp.so (the C ABI declared in include/msp.h).

There is no fallback: if the shared library is missing or a symbol cannot be
resolved, importing this module raises.  Build it with
``python __graft_entry__.py`` or ``make -C multi_spatialmllm_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MSP_LIB", os.path.join(_HERE, "libmsp.so"))   # MSP_LIB: alternative build (tuning experiments)

MSP_PTS_F32X3, MSP_PTS_F32X4, MSP_PTS_F64X3 = 0, 1, 2


class MspCamera(C.Structure):
    _fields_ = [("image_h", C.c_int32), ("image_w", C.c_int32),
                ("depth_h", C.c_int32), ("depth_w", C.c_int32),
                ("depth_scale", C.c_double)]


class MspError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: the CUDA extension is not built. "
        "Run `python __graft_entry__.py` (or `make -C multi_spatialmllm_b200/csrc`). "
        "There is deliberately no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_P = C.c_void_p
_SIGNATURES = {
    "msp_version": (C.c_int, []),
    "msp_last_error": (C.c_char_p, []),
    "msp_debug_set_cull_stats": (C.c_int, [_P]),
    "msp_pose_prepare": (C.c_int, [_P, _P, C.c_int32, _P, _P, _P, _P, _P]),
    "msp_project_visibility": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, C.c_int32, _P, C.POINTER(MspCamera),
                                         _P, C.c_int64, _P, _P, _P]),
    "msp_project_visibility_objects": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, C.c_int32, _P, C.POINTER(MspCamera),
                                                 _P, C.c_int64, _P, C.c_int32, _P, _P]),
    "msp_zbuffer_render": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, C.c_int32, C.POINTER(MspCamera), _P, _P]),
    "msp_depth_to_u16": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_double, _P, _P]),
    "msp_pair_overlap": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int64, _P, _P, _P, _P]),
    "msp_pair_annotate": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int64, _P, C.c_int32, _P, C.c_int32,
                                    _P, _P, _P, _P, _P, C.POINTER(MspCamera), _P, _P]),
    "msp_project_points": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "msp_annotate_points": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, C.c_int64, _P, _P, _P,
                                      C.POINTER(MspCamera), _P, _P, _P, _P, _P, _P]),
    "msp_check_visibility": (C.c_int, [_P, _P, C.c_int64, _P, C.POINTER(MspCamera), C.c_int32, _P, _P]),
    "msp_object_counts": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int32, _P, _P]),
    "msp_object_extents": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, _P]),
    "msp_track_project": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    "msp_track_cumulative_loss": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_double, _P, _P]),
    "msp_track_pair_displacement": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, C.c_int64, _P, _P]),
    "msp_track_all_pair_distances": (C.c_int, [_P, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, _P, C.c_int64, _P, _P]),
    "msp_jpeg_info": (C.c_int, [_P, C.c_size_t, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "msp_jpeg_decode_bgr": (C.c_int, [_P, C.c_size_t, _P, C.c_int32, C.c_int32, _P]),
    "msp_jpeg_encode_bgr": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_size_t, C.POINTER(C.c_size_t), _P]),
    "msp_visibility_json_points_size": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int32, _P, _P]),
    "msp_visibility_json_points_write": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int32, _P, _P, _P]),
    "msp_visibility_json_images_size": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int32, _P, _P, _P]),
    "msp_visibility_json_images_write": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int32, _P, C.c_int32, _P, _P, _P, _P]),
    "msp_draw_marks": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P, _P, C.c_int32, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError if the symbol is missing: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise MspError(f"{what or 'libmsp'} failed ({rc}): {lib.msp_last_error().decode()}")


def version() -> int:
    return lib.msp_version()
