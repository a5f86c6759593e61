"""ctypes binding of libpbnet_b200.so (include/pbnet_b200.h).  There is no CPU fallback: a missing
library or a missing CUDA device raises."""
from __future__ import annotations

import ctypes
import os

from . import build as _build

_lib = None

ERR_NAMES = {1: "PB_ERR_ARG", 2: "PB_ERR_CUDA", 3: "PB_ERR_SEM_RANGE", 4: "PB_ERR_NONFINITE", 5: "PB_ERR_RANGE",
             6: "PB_ERR_MIXED_CLASS", 7: "PB_ERR_CAPACITY", 8: "PB_ERR_NOMEM"}

# every symbol include/pbnet_b200.h declares
SYMBOLS = ["pb_create", "pb_destroy", "pb_last_error", "pb_last_launch_count", "pb_binary_cluster",
           "pb_binary_cluster_batched", "pb_set_profiling", "pb_stage_count", "pb_stage_name", "pb_stage_ms",
           "pb_counter", "pb_set_chunk_points", "pb_set_small_calls", "pb_selftest_division", "pb_voxelize", "pb_voxel_rows", "pb_devoxelize", "pb_get_iou", "pb_cal_iou_and_masklabel",
           "pb_group_front", "pb_local_scenes_plan", "pb_local_scenes_fill", "pb_get_proposal", "pb_scene_features", "pb_eval_postprocess", "pb_cal_normal_line"]


class PBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR_NAMES.get(code, code)}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    so = _build.SO
    if not os.path.exists(so):
        raise ImportError(f"{so} is missing — run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(pbnet_b200 has no CPU fallback)")
    L = ctypes.CDLL(so)
    vp, fp, ip = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p  # raw addresses (host or device)
    L.pb_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    L.pb_create.restype = ctypes.c_int
    L.pb_destroy.argtypes = [vp]
    L.pb_destroy.restype = None
    L.pb_last_error.argtypes = [vp]
    L.pb_last_error.restype = ctypes.c_char_p
    L.pb_last_launch_count.argtypes = [vp]
    L.pb_last_launch_count.restype = ctypes.c_int64
    common_tail = [fp, ip, ctypes.c_float, ctypes.c_int, ip, ip, ip, fp, ctypes.c_int64, ip, ctypes.c_int64,
                   ctypes.POINTER(ctypes.c_int64)]
    L.pb_binary_cluster.argtypes = [vp, fp, fp, fp, fp, fp, fp, ip, ip, ctypes.c_int32, ctypes.c_int64,
                                    *common_tail, ctypes.c_int, vp]
    L.pb_binary_cluster.restype = ctypes.c_int
    L.pb_binary_cluster_batched.argtypes = [vp, fp, fp, fp, fp, fp, fp, ip, ip, ctypes.c_int32, ip, ctypes.c_int32,
                                            ctypes.c_int64, *common_tail, ctypes.c_void_p, ctypes.c_int, vp]
    L.pb_binary_cluster_batched.restype = ctypes.c_int
    L.pb_set_profiling.argtypes = [vp, ctypes.c_int]
    L.pb_set_profiling.restype = None
    L.pb_set_chunk_points.argtypes = [vp, ctypes.c_int64]
    L.pb_set_chunk_points.restype = None
    L.pb_set_small_calls.argtypes = [vp, ctypes.c_int]
    L.pb_set_small_calls.restype = None
    L.pb_selftest_division.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64)]
    L.pb_selftest_division.restype = ctypes.c_int
    L.pb_get_iou.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int32, ctypes.c_int32, vp]
    L.pb_get_iou.restype = ctypes.c_int
    L.pb_cal_iou_and_masklabel.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int32, ctypes.c_int32, vp, vp, ctypes.c_int, vp]
    L.pb_cal_iou_and_masklabel.restype = ctypes.c_int
    i64p = ctypes.POINTER(ctypes.c_int64)
    L.pb_local_scenes_plan.argtypes = [vp, vp, vp, ctypes.c_int32, vp, vp, ctypes.c_int32, ctypes.c_int64, vp, vp, ctypes.c_int64,
                                       vp, vp, vp, i64p, i64p, vp]
    L.pb_local_scenes_plan.restype = ctypes.c_int
    L.pb_local_scenes_fill.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.pb_local_scenes_fill.restype = ctypes.c_int
    L.pb_get_proposal.argtypes = [vp, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, ctypes.c_float, vp, vp, vp, vp, i64p, i64p, vp]
    L.pb_get_proposal.restype = ctypes.c_int
    L.pb_scene_features.argtypes = [vp, vp, ctypes.c_int32, vp, ctypes.c_int32, vp, vp, vp, vp, ctypes.c_int64, vp, vp]
    L.pb_scene_features.restype = ctypes.c_int
    L.pb_eval_postprocess.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.c_int64, vp, vp, ctypes.c_int64, ctypes.c_int32, vp,
                                      ctypes.c_int64, vp, ctypes.c_int32, ctypes.c_float, ctypes.c_int32, ctypes.c_float, vp, vp, vp,
                                      vp, ctypes.c_int64, i64p, vp]
    L.pb_eval_postprocess.restype = ctypes.c_int
    L.pb_cal_normal_line.argtypes = [vp, vp, vp, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int, vp]
    L.pb_cal_normal_line.restype = ctypes.c_int
    L.pb_group_front.argtypes = [vp, vp, vp, vp, vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                                 i64p, vp]
    L.pb_group_front.restype = ctypes.c_int
    L.pb_stage_count.argtypes = []
    L.pb_stage_count.restype = ctypes.c_int
    L.pb_stage_name.argtypes = [ctypes.c_int]
    L.pb_stage_name.restype = ctypes.c_char_p
    L.pb_stage_ms.argtypes = [vp, ctypes.c_int]
    L.pb_stage_ms.restype = ctypes.c_float
    L.pb_counter.argtypes = [vp, ctypes.c_int]
    L.pb_counter.restype = ctypes.c_int64
    _lib = L
    return L
