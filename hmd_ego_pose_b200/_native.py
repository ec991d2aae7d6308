"""ctypes binding of libhmdpose.so (include/hmdpose.h).  There is NO fallback: if the shared
library is missing or cannot be loaded this module raises, and every compute entry point fails
loudly when no CUDA device is present."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libhmdpose.so")

ABI_VERSION = 1
PRECISION_PARITY = 0
PRECISION_FAST = 1
NUM_HAND = 63
BEST_LEN = 11


class Config(ctypes.Structure):
    _fields_ = [("abi_version", c_int), ("image_size", c_int), ("max_batch", c_int), ("device", c_int),
                ("precision", c_int), ("num_classes", c_int), ("score_threshold", c_float),
                ("iou_threshold", c_float), ("max_detections", c_int), ("micro_batch", c_int),
                ("use_graph", c_int)]


# every symbol include/hmdpose.h declares: name -> (restype, argtypes)
FP = c_void_p  # float* / int32_t* passed as raw addresses (host or device)
SYMBOLS = {
    "hmdpose_version": (c_char_p, []),
    "hmdpose_default_config": (None, [POINTER(Config)]),
    "hmdpose_create": (c_int, [c_char_p, c_int, c_int, c_int, c_float, c_float, c_int, POINTER(c_void_p)]),
    "hmdpose_create_ex": (c_int, [POINTER(Config), c_char_p, POINTER(c_void_p)]),
    "hmdpose_create_from_memory": (c_int, [POINTER(Config), c_void_p, c_size_t, POINTER(c_void_p)]),
    "hmdpose_destroy": (None, [c_void_p]),
    "hmdpose_last_error": (c_char_p, [c_void_p]),
    "hmdpose_num_anchors": (c_int, [c_void_p]),
    "hmdpose_num_classes": (c_int, [c_void_p]),
    "hmdpose_get_anchors": (c_int, [c_void_p, FP, FP]),
    "hmdpose_compute_anchors": (c_int, [c_int, FP, FP, c_int]),
    "hmdpose_preprocess": (c_int, [c_void_p, FP, c_int, c_int, c_int, FP, FP]),
    "hmdpose_run_detect_u8": (c_int, [c_void_p, FP, c_int, c_int, c_int, FP, FP, FP, FP, FP, FP, FP, FP, FP]),
    "hmdpose_run_best_u8": (c_int, [c_void_p, FP, c_int, c_int, FP, FP, FP]),
    "hmdpose_preprocess_i420": (c_int, [c_void_p, FP, c_int, c_int, c_int, c_int, c_int, FP, FP]),
    "hmdpose_run_best_i420": (c_int, [c_void_p, FP, c_int, c_int, c_int, c_int, FP, FP, FP]),
    "hmdpose_pose_packet": (c_int, [FP, FP]),
    "hmdpose_run_packet": (c_int, [c_void_p, FP, FP, FP, FP]),
    "hmdpose_compute_anchors_d0": (c_int, [c_int, FP, c_int]),
    "hmdpose_run_d0": (c_int, [c_void_p, FP, c_int, c_float, c_float, c_int, FP, FP, FP, FP, FP]),
    "hmdpose_d0_postprocess": (c_int, [c_void_p, FP, FP, c_int, c_float, c_float, c_int, FP, FP, FP, FP, FP]),
    "hmdpose_run_raw": (c_int, [c_void_p, FP, c_int, FP, FP, FP, FP, FP]),
    "hmdpose_run_detect": (c_int, [c_void_p, FP, FP, c_int, FP, FP, FP, FP, FP, FP, FP]),
    "hmdpose_run_best": (c_int, [c_void_p, FP, FP, FP]),
    "hmdpose_postprocess": (c_int, [c_void_p, FP, FP, FP, FP, FP, FP, c_int, FP, FP, FP, FP, FP, FP, FP]),
    "hmdpose_filter_boxes": (c_int, [c_void_p, FP, FP, FP, FP, FP, c_int, FP, FP, FP, FP, FP, FP, FP]),
    "hmdpose_best_from_raw": (c_int, [c_void_p, FP, FP, FP, FP, FP, FP]),
    "hmdpose_run_raw_device": (c_int, [c_void_p, FP, c_int64, c_int64, c_int64, c_int64, c_int, FP, FP, FP, FP, FP,
                                       c_void_p]),
    "hmdpose_run_detect_device": (c_int, [c_void_p, FP, c_int64, c_int64, c_int64, c_int64, FP, c_int, FP, FP, FP, FP,
                                          FP, FP, FP, c_void_p]),
    "hmdpose_debug_read": (c_int64, [c_void_p, c_char_p, FP, c_int64]),
    "hmdpose_profile_steps": (c_int, [c_void_p, c_int, c_int, c_int, c_char_p, c_char_p, FP, FP, FP, c_int]),
    "hmdpose_last_launch_count": (c_int, [c_void_p]),
    "hmdpose_last_gpu_ms": (c_float, [c_void_p]),
    "hmdpose_test_gemm": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, FP, FP, FP, FP, c_int, FP, c_int, FP,
                                  POINTER(c_float)]),
}

_lib = None


def load() -> ctypes.CDLL:
    """Load libhmdpose.so (building it in-tree first if nvcc is available and sources changed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) or os.environ.get("HMDPOSE_REBUILD"):
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m hmd_ego_pose_b200.build` "
                          "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class HmdPoseError(RuntimeError):
    pass


def check(rc: int, handle=None) -> None:
    if rc != 0:
        msg = load().hmdpose_last_error(handle)
        raise HmdPoseError(f"libhmdpose error {rc}: {msg.decode() if msg else ''}")
