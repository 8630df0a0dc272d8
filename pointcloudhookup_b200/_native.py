"""ctypes binding of libpch_b200.so (the C ABI declared in include/pch_b200.h).

There is NO CPU fallback: if the shared library is missing or a call fails, this raises.
ctypes releases the GIL for the duration of each call, so the reference GUI's worker threads
(pyGUI_towers_test.py:385) stay responsive.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PCH_LIB_PATH") or os.path.join(_HERE, "libpch_b200.so")   # override: tuning variants only


class NativeError(RuntimeError):
    pass


class VoxelPlan(C.Structure):
    _fields_ = [("bits_x", C.c_int32), ("bits_y", C.c_int32), ("bits_z", C.c_int32), ("bits_idx", C.c_int32),
                ("key_bits", C.c_int32), ("n_passes", C.c_int32), ("status", C.c_int32), ("reserved", C.c_int32)]


_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32
_f64 = C.c_double
_sz = C.c_size_t
_d3 = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol include/pch_b200.h declares
SIGNATURES = {
    "pch_last_error": (C.c_char_p, []),
    "pch_version": (C.c_int, []),
    "pch_launch_count": (C.c_longlong, []),
    "pch_profile_enable": (None, [C.c_int]),
    "pch_profile_report": (C.c_int, [C.c_char_p, _sz]),
    "pch_las_chunk_minmax": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _p]),
    "pch_voxel_keys_xyz16": (C.c_int, [_p, _i64, _i64, _d3, _d3, _f64, _p, C.POINTER(VoxelPlan), _p, _p]),
    "pch_las_decode_f64": (C.c_int, [_p, _i64, _i32, _d3, _d3, _p, _p]),
    "pch_las_decode_f32": (C.c_int, [_p, _i64, _i32, _d3, _d3, _p, _p]),
    "pch_las_quantise": (C.c_int, [_p, _i64, _d3, _d3, _p, _p]),
    "pch_las_encode": (C.c_int, [_p, _i64, _i32, _p, _p, _p]),
    "pch_voxel_plan_build": (C.c_int, [_p, _i64, _i64, _d3, _d3, _f64, _p, _p, _p]),
    "pch_voxel_keys": (C.c_int, [_p, _i64, _i32, _i64, _d3, _d3, _f64, _p, C.POINTER(VoxelPlan), _p, _p, _p]),
    "pch_voxel_plan_build_f64": (C.c_int, [_p, _i64, _i64, _f64, _p, _p, _p, _p]),
    "pch_voxel_keys_f64": (C.c_int, [_p, _i64, _i64, _f64, _p, C.POINTER(VoxelPlan), _p, _p]),
    "pch_sort_workspace_bytes": (_sz, [_i64, _i64, _i32, _i32]),
    "pch_sort_u64_segmented": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _p, _sz, _p]),
    "pch_voxel_reduce_workspace_bytes": (_sz, [_i64, _i64]),
    "pch_voxel_reduce": (C.c_int, [_p, _i64, _i64, _i32, _p, _i32, _p, _p, _d3, _d3, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "pch_voxel_downsample_las_workspace_bytes": (_sz, [_i64, _i64]),
    "pch_voxel_downsample_las": (C.c_int, [_p, _i64, _i32, _i64, _d3, _d3, _f64, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p,
                                           _sz, _p]),
    "pch_voxel_index3_f64": (C.c_int, [_p, _i64, _i64, _f64, _p, _p, _p]),
    "pch_voxel_wide_words": (C.c_int, [_p, _p, _i64, _i64, _i32, _i32, _p, _p]),
    "pch_selftest_fastdiv": (C.c_int, [_p, _i64, _f64, _p, _p]),
    "pch_selftest_fastdiv_f32": (C.c_int, [_p, _i64, C.c_float, _p, _p]),
}

class ClusterStats(C.Structure):
    _fields_ = [("count", C.c_int64), ("min", C.c_float * 3), ("max", C.c_float * 3), ("sum", C.c_double * 3)]


SIGNATURES.update({
    "pch_f32_centroid_workspace_bytes": (_sz, [_i64]),
    "pch_f32_centroid": (C.c_int, [_p, _i64, _p, _p, _p, _sz, _p]),
    "pch_f32_shift": (C.c_int, [_p, _i64, _p, _p, _p, _p]),
    "pch_f32_column": (C.c_int, [_p, _i64, _i32, _p, _p]),
    "pch_select_workspace_bytes": (_sz, []),
    "pch_select_f32": (C.c_int, [_p, _i64, _i64, _i64, _p, _p, _sz, _p]),
    "pch_compact_workspace_bytes": (_sz, [_i64]),
    "pch_compact_points": (C.c_int, [_p, _p, _p, _i64, _p, C.c_float, _p, _p, _p, _p, _p, _sz, _p]),
    "pch_grid_min": (C.c_int, [_p, _i64, _p, C.c_float, C.c_float, C.c_float, _i32, _i32, _p, _p]),
    "pch_compact_points_grid": (C.c_int, [_p, _i64, _p, C.c_float, C.c_float, C.c_float, _i32, _i32, C.c_float, _p, _p, _p, _p, _p,
                                          _p, _sz, _p]),
    "pch_grid_min_ground": (C.c_int, [_p, _i64, C.c_float, C.c_float, C.c_float, _i32, _i32, C.c_float, _p, _p, _p, _p]),
    "pch_f32_minmax": (C.c_int, [_p, _i64, _p, _p]),
    "pch_label_words": (C.c_int, [_p, _i64, _p, _p]),
    "pch_gather_rows_f32": (C.c_int, [_p, _p, _i64, _p, _p, _p]),
    "pch_dbscan_plan": (C.c_int, [_p, _i64, _i64, _f64, _p, _p, _p]),
    "pch_dbscan_workspace_bytes": (_sz, [_i64, _i64, C.POINTER(VoxelPlan), _i64]),
    "pch_dbscan_fused_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "pch_dbscan": (C.c_int, [_p, _i64, _i64, _f64, _i32, _p, _p, _i64, _p, _sz, _p]),
    "pch_dbscan_cores": (C.c_int, [_p, _i64, _i64, _f64, _i32, _p, _i64, _p, _sz, _p]),
    "pch_dbscan_acc_bytes": (_sz, [_i64]),
    "pch_dbscan_finish": (C.c_int, [_p, _i64, _i64, _f64, _i32, _p, _i64, _i64, _i64, _p, _p, _p, _i64, _p, _sz, _p]),
    "pch_label_min_index": (C.c_int, [_p, _i64, _i64, _i64, _i64, _p, _p]),
    "pch_axis_extent": (C.c_int, [_p, _i64, _f64, _f64, _p, _p]),
    "pch_axis_band_mask": (C.c_int, [_p, _i64, _f64, _f64, _f64, _f64, _p, _p, _p]),
    "pch_xy_minmax_f64": (C.c_int, [_p, _i64, _p, _p, _p]),
    "pch_ransac_tile_words": (C.c_int, [_p, _i64, C.POINTER(C.c_double), _i32, C.POINTER(C.c_double), _i32, _p, _p]),
    "pch_gather_rows_f64": (C.c_int, [_p, _p, _i64, _p, _p]),
    "pch_ransac_tiles": (C.c_int, [_p, _p, _i32, _f64, _i32, _f64, C.c_uint64, _p, _i32, _p, _p, _p]),
    "pch_ransac_split": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p, _p]),
    "pch_obb_workspace_bytes": (_sz, [_i32]),
    "pch_obb_batch": (C.c_int, [_p, _p, _i32, _p, _p, _sz, _p]),
    "pch_dbscan_run": (C.c_int, [_p, _i64, _i64, _f64, _i32, _p, C.POINTER(VoxelPlan), _p, _p, _p, _i64, _p, _sz, _p]),
})

class GeoidGrid(C.Structure):
    _fields_ = [("ll_lat", C.c_double), ("ll_lon", C.c_double), ("dlat", C.c_double), ("dlon", C.c_double),
                ("rows", C.c_int32), ("cols", C.c_int32), ("pitch", C.c_int32), ("is_global", C.c_int32)]


class TmParams(C.Structure):
    _fields_ = [("rect_radius", C.c_double), ("beta", C.c_double * 6), ("ecc", C.c_double),
                ("lon0_deg", C.c_double), ("k0", C.c_double), ("fe", C.c_double), ("fn", C.c_double)]


SIGNATURES.update({
    "pch_geoid_shift": (C.c_int, [_p, _p, _p, _i64, _p, C.POINTER(GeoidGrid), _f64, _p, _p, _p]),
    "pch_haversine_matrix": (C.c_int, [_p, _p, _i64, _p, _p, _i64, _p, _p]),
    "pch_gk_inverse": (C.c_int, [_p, _p, _i64, C.POINTER(TmParams), _p, _p, _p]),
    "pch_las_geodetic": (C.c_int, [_p, _i64, _i32, _d3, _d3, C.POINTER(TmParams), _p, C.POINTER(GeoidGrid),
                                   _i32, _i32, _i32, _i32, _f64, _p, _p]),
})

SIGNATURES.update({
    "pch_host_pack_xyz": (C.c_int, [_p, _i64, _i32, _p, _i32]),
    "pch_las_box_crop": (C.c_int, [_p, _i64, _i32, _d3, _d3, _p, _i32, _p, _i64, _p, _p]),
    "pch_word_bounds": (C.c_int, [_p, _i64, _i32, _p, _p]),
    "pch_las_gather_f64": (C.c_int, [_p, _i64, _i32, _d3, _d3, _p, _i64, _p, _p]),
    "pch_sample_indices": (C.c_int, [_i64, _i64, C.c_uint64, _p, _p]),
})

_lib = None
_lock = threading.Lock()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NativeError(
                        f"{LIB_PATH} is not built (run `python -m pointcloudhookup_b200.build`); "
                        "pointcloudhookup_b200 has no CPU fallback")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.restype = res
                    fn.argtypes = args
                _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pch_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what or 'pch call'} failed ({rc}): {msg}")


def d3(values) -> "C.Array":
    return (C.c_double * 3)(*[float(v) for v in values])
