"""ctypes binding of libviddet_b200.so (the C ABI declared in include/viddet_b200.h).

There is NO fallback: if the library is missing or a call fails, a VidDetError is raised.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("VD_LIB") or os.path.join(HERE, "libviddet_b200.so")     # VD_LIB: tuning variants (build.py --variant)

VD_MAX_SCALES = 3
VD_MAX_TOPK = 1024
VD_MAX_MIRRORS = 7
VD_MODE_INFER, VD_MODE_TRAIN, VD_MODE_AGNOSTIC = 0, 1, 2
VD_JOIN_NONE, VD_JOIN_CAT, VD_JOIN_MAX, VD_JOIN_MEAN = 0, 1, 2, 3
VD_STAGE_TCONV, VD_STAGE_HEAD, VD_STAGE_NMS, VD_STAGE_ALL = 1, 2, 4, 7
VD_HEAD_NO_FUSED_TIP, VD_HEAD_NO_PAIR_KERNEL = 1, 2          # VdHeadParams.flags
VD_PREC_BF16, VD_PREC_FP32_SPLIT, VD_PREC_BF16X2 = 0, 1, 2
PLANES = {VD_PREC_BF16: 1, VD_PREC_FP32_SPLIT: 3, VD_PREC_BF16X2: 2}
ERR_NAMES = {-1: "VD_ERR_INVALID_ARG", -2: "VD_ERR_UNSUPPORTED", -3: "VD_ERR_WORKSPACE", -4: "VD_ERR_CUDA"}


class VidDetError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "VD_ERR"), code, msg))
        self.code = code


class VdHeadScale(ctypes.Structure):
    _fields_ = [
        ("tip_nhwc_bf16", ctypes.c_void_p), ("weight_bf16", ctypes.c_void_p), ("bias", ctypes.c_void_p),
        ("H", ctypes.c_int), ("W", ctypes.c_int), ("Cin", ctypes.c_int),
        ("stride", ctypes.c_float), ("anchors", ctypes.c_float * 6),
        ("tconv_weight_bf16", ctypes.c_void_p), ("tconv_scale", ctypes.c_void_p),
        ("tconv_shift", ctypes.c_void_p), ("tconv_out_nhwc_bf16", ctypes.c_void_p),
        ("tip_window_stride_frames", ctypes.c_int), ("reserved", ctypes.c_int),
    ]


class VdHeadParams(ctypes.Structure):
    _fields_ = [
        ("num_scales", ctypes.c_int), ("num_class", ctypes.c_int), ("frames", ctypes.c_int),
        ("T", ctypes.c_int), ("K_frames", ctypes.c_int), ("join", ctypes.c_int),
        ("nms_thresh", ctypes.c_float), ("valid_thresh", ctypes.c_float),
        ("nms_topk", ctypes.c_int), ("post_nms", ctypes.c_int),
        ("precision", ctypes.c_int), ("flags", ctypes.c_int),
        ("scale", VdHeadScale * VD_MAX_SCALES),
        ("n_mirrors", ctypes.c_int), ("reserved1", ctypes.c_int),
        ("mirror_delta", ctypes.c_longlong * VD_MAX_MIRRORS),
    ]


# symbol -> (restype, argtypes); every symbol declared in include/viddet_b200.h must be here
_vp, _i, _i64, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
_ip = ctypes.POINTER(ctypes.c_int)
_fp = ctypes.POINTER(ctypes.c_float)
SIGNATURES = {
    "vd_version": (_i, []),
    "vd_last_error": (ctypes.c_char_p, []),
    "vd_device_info": (_i, [_i, _ip, _ip, _ip]),
    "vd_box_nms_workspace_bytes": (_sz, [_i64, _i64, _i, _i]),
    "vd_box_nms": (_i, [_vp, _i64, _i64, _i, _f, _f, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "vd_yolo_decode": (_i, [_vp, _i, _i, _i, _i, _i, _fp, _f, _i, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "vd_repack_nchw_f32_to_nhwc_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "vd_pred_conv": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "vd_pred_conv_ex": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "vd_repack_nchw_f32_to_nhwc_split": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "vd_split_f32_rows": (_i, [_vp, _vp, _i64, _i64, _i, _vp]),
    "vd_sizeof": (_sz, [_i]),
    "vd_head_workspace_bytes": (_sz, [ctypes.POINTER(VdHeadParams)]),
    "vd_head_forward": (_i, [ctypes.POINTER(VdHeadParams), _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "vd_head_forward_stages": (_i, [ctypes.POINTER(VdHeadParams), _vp, _vp, _vp, _vp, _vp, _sz, _vp, _i]),
    "vd_head_launch_count": (_i, [ctypes.POINTER(VdHeadParams)]),
    "vd_head_fused_tip": (_i, [ctypes.POINTER(VdHeadParams)]),
    "vd_head_fused_tip_plan": (_i, [ctypes.POINTER(VdHeadParams), _i, _ip, _ip, _ip]),
    "vd_head_stats_offset": (_sz, [ctypes.POINTER(VdHeadParams)]),
    "vd_head_debug_offset": (_sz, [ctypes.POINTER(VdHeadParams)]),
    "vd_head_detections": (_i, [ctypes.POINTER(VdHeadParams), _vp, _vp, _sz, _vp]),
    "vd_ipc_alloc": (_i, [_sz, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_ubyte)]),
    "vd_ipc_open": (_i, [ctypes.POINTER(ctypes.c_ubyte), ctypes.POINTER(ctypes.c_void_p)]),
    "vd_ipc_close": (_i, [_vp]),
    "vd_ipc_free": (_i, [_vp]),
    "vd_temporal_conv": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp]),
    "vd_temporal_conv_ex": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _i, _i, _vp]),
    "vd_conv_bn_lrelu": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _f, _vp]),
    "vd_upsample_concat": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "vd_conv_tile_box": (_i, [_i, _i, _i, _ip, _ip, _ip]),
    "vd_temporal_pool": (_i, [_vp, _vp, _i, _i, _i64, _i, _vp]),
    "vd_prefetch_targets": (_i, [_i, _i, _i, _i, _i, _ip, _fp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vd_target_merge": (_i, [_i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vd_postprocess_detections": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp]),
    "vd_hierarchical_nms": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _vp, ctypes.c_double, ctypes.c_double, _i, _i, _vp, _vp, _vp]),
    "vd_yolo3_loss_workspace_bytes": (_sz, [_i, _i]),
    "vd_yolo3_loss": (_i, [_i, _i, _i] + [_vp] * 14 + [_vp, _sz, _vp]),
}

_lib = None


def load():
    """Load the shared library (fails loudly if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise VidDetError(-4, "%s not found -- run `python -m viddet_b200.build` (there is no CPU fallback)" % SO_PATH)
        lib = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.vd_sizeof(0) != ctypes.sizeof(VdHeadScale) or lib.vd_sizeof(1) != ctypes.sizeof(VdHeadParams):
            raise VidDetError(-1, "struct mirror out of date: library VdHeadScale/VdHeadParams = %d/%d bytes, ctypes %d/%d"
                              % (lib.vd_sizeof(0), lib.vd_sizeof(1), ctypes.sizeof(VdHeadScale), ctypes.sizeof(VdHeadParams)))
        _lib = lib
    return _lib


def check(code):
    if code != 0:
        raise VidDetError(code, load().vd_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
