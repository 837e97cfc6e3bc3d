"""ctypes loader for the product library ``liblfx.so`` (C ABI in include/lfx.h).

There is no CPU fallback: if the CUDA library is missing the import fails loudly, and
``lfx_create`` fails with ``LFX_E_CUDA`` when no device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.environ.get("LFX_LIB") or os.path.join(PKG_DIR, "liblfx.so")  # LFX_LIB: A/B builds of the same ABI
HEADER = os.path.join(ROOT, "include", "lfx.h")

(LFX_OK, LFX_E_BAD_PARAM, LFX_E_NOT_DENSE, LFX_E_NO_RING, LFX_E_BAD_LAYOUT, LFX_E_CAPACITY, LFX_E_CUDA, LFX_E_STATE,
 LFX_E_CONVERT) = range(9)
STATUS_NAMES = ["LFX_OK", "LFX_E_BAD_PARAM", "LFX_E_NOT_DENSE", "LFX_E_NO_RING", "LFX_E_BAD_LAYOUT", "LFX_E_CAPACITY",
                "LFX_E_CUDA", "LFX_E_STATE", "LFX_E_CONVERT"]
CONVERT_STATUS_NAMES = ["LFX_CONVERT_OK", "LFX_CONVERT_E_SIZE", "LFX_CONVERT_E_DATATYPE", "LFX_CONVERT_E_LAYOUT",
                        "LFX_CONVERT_E_FEW_FIELDS", "LFX_CONVERT_E_FIELD_COUNT", "LFX_CONVERT_E_OVERFLOW",
                        "LFX_CONVERT_E_RING_TYPE", "LFX_CONVERT_E_RING_RANGE"]
LFX_MEM_HOST, LFX_MEM_DEVICE = 0, 1
LFX_RING_U8, LFX_RING_U16, LFX_RING_U32 = 2, 4, 6
LFX_RING_OK, LFX_RING_SPARSE, LFX_RING_SKIPPED, LFX_RING_TOO_LONG = range(4)
LFX_WORLD_ROOM, LFX_WORLD_TUNNEL = 0, 1
LFX_TOPIC_SCAN_EDGE, LFX_TOPIC_SCAN_SURFACE, LFX_TOPIC_COLORED_SCAN = 0, 1, 2
LFX_LOC_EDGE, LFX_LOC_SURFACE = 0, 1


class Params(C.Structure):
    """lfx_params == HyperParameters (hyper_parameter.hpp:32-65)."""

    _fields_ = [
        ("padding", C.c_int),
        ("neighbor_degree_threshold", C.c_double),
        ("distance_diff_threshold", C.c_double),
        ("parallel_beam_min_range_ratio", C.c_double),
        ("edge_threshold", C.c_double),
        ("surface_threshold", C.c_double),
        ("min_range", C.c_double),
        ("max_range", C.c_double),
        ("n_blocks", C.c_int),
    ]


class Options(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("max_ring_points", C.c_int),
        ("max_rings", C.c_int),
        ("want_sorted_src", C.c_int),
        ("want_curvature", C.c_int),
        ("force_order_path", C.c_int),
        ("stream", C.c_void_p),
        ("use_graph", C.c_int),
    ]


class CloudView(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("n_points", C.c_uint32),
        ("point_step", C.c_uint32),
        ("off_x", C.c_uint32),
        ("off_y", C.c_uint32),
        ("off_z", C.c_uint32),
        ("off_ring", C.c_uint32),
        ("ring_datatype", C.c_uint8),
        ("has_ring", C.c_uint8),
        ("is_dense", C.c_uint8),
        ("memory", C.c_uint8),
    ]


class RingInfo(C.Structure):
    _fields_ = [
        ("count", C.c_uint32),
        ("offset", C.c_uint32),
        ("n_edge", C.c_uint32),
        ("n_surface", C.c_uint32),
        ("status", C.c_uint32),
        ("order_path", C.c_uint32),
    ]


class BatchResult(C.Structure):
    _fields_ = [
        ("n_scans", C.c_int),
        ("total_points", C.c_uint64),
        ("d_edge_xyz", C.c_void_p),
        ("d_surface_xyz", C.c_void_p),
        ("d_counts", C.c_void_p),
        ("d_offsets", C.c_void_p),
        ("d_labels", C.c_void_p),
        ("d_sorted_src", C.c_void_p),
        ("d_curvature", C.c_void_p),
        ("d_rings", C.c_void_p),
        ("d_point_base", C.c_void_p),
        ("max_rings", C.c_int),
    ]


class ScanOutput(C.Structure):
    _fields_ = [
        ("edge_xyz", C.c_void_p),
        ("surface_xyz", C.c_void_p),
        ("n_edge", C.c_uint32),
        ("n_surface", C.c_uint32),
        ("labels", C.c_void_p),
        ("sorted_src", C.c_void_p),
        ("n_points", C.c_uint32),
    ]


class ShardResult(C.Structure):
    """lfx_shard_result."""

    _fields_ = [("n_frames", C.c_uint64), ("first_frame", C.c_uint64), ("last_frame", C.c_uint64),
                ("d_counts_all", C.c_void_p), ("d_offsets_all", C.c_void_p)]


SHARD_ID_BYTES = 128


class BatchStats(C.Structure):
    _fields_ = [("fast_rings", C.c_uint32 * 3), ("general_scans", C.c_uint32), ("general_rings", C.c_uint32),
                ("indexed_rings", C.c_uint32 * 3)]


class PointFieldC(C.Structure):
    """lfx_point_field == sensor_msgs/PointField."""

    _fields_ = [("name", C.c_char_p), ("offset", C.c_uint32), ("datatype", C.c_uint8), ("count", C.c_uint32)]


class RawCloud(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_bytes", C.c_uint64), ("point_step", C.c_uint32),
                ("fields", C.POINTER(PointFieldC)), ("n_fields", C.c_uint32), ("is_bigendian", C.c_uint8),
                ("memory", C.c_uint8)]


class ConvertResult(C.Structure):
    _fields_ = [("n_clouds", C.c_int), ("d_points", C.c_void_p), ("point_base", C.POINTER(C.c_uint64)),
                ("kept", C.POINTER(C.c_uint32)), ("status", C.POINTER(C.c_uint32))]


class Pose(C.Structure):
    """lfx_pose == geometry_msgs/Pose (orientation x, y, z, w)."""

    _fields_ = [("position", C.c_double * 3), ("orientation", C.c_double * 4)]


class ColoredResult(C.Structure):
    _fields_ = [("n_scans", C.c_int), ("d_points", C.c_void_p), ("point_base", C.POINTER(C.c_uint64)),
                ("counts", C.POINTER(C.c_uint32))]


class SynthSpec(C.Structure):
    _fields_ = [
        ("n_rings", C.c_int),
        ("n_cols", C.c_int),
        ("elev_lo_deg", C.c_float),
        ("elev_hi_deg", C.c_float),
        ("world", C.c_int),
        ("range_noise", C.c_float),
        ("dropout_prob", C.c_float),
        ("dropout_burst", C.c_float),
        ("near_prob", C.c_float),
        ("seed", C.c_uint64),
    ]


def build(verbose: bool = False) -> str:
    """Compile liblfx.so for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(PKG_DIR, "csrc")]
    if not verbose:
        cmd.insert(1, "-s")
    subprocess.run(cmd, check=True)
    return LIB_PATH


def declared_symbols() -> list[str]:
    """Every function include/lfx.h declares (used by the symbol-export test)."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lfx_[a-z0-9_]+)\s*\(", text)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.lfx_default_params.argtypes = [C.POINTER(Params)]
    L.lfx_default_params.restype = None
    L.lfx_launch_yaml_params.argtypes = [C.POINTER(Params)]
    L.lfx_launch_yaml_params.restype = None
    L.lfx_create.argtypes = [C.POINTER(Params), C.POINTER(Options), C.POINTER(H)]
    L.lfx_destroy.argtypes = [H]
    L.lfx_destroy.restype = None
    L.lfx_last_error.argtypes = [H]
    L.lfx_last_error.restype = C.c_char_p
    L.lfx_get_params.argtypes = [H, C.POINTER(Params)]
    L.lfx_device.argtypes = [H]
    L.lfx_stream.argtypes = [H]
    L.lfx_stream.restype = C.c_void_p
    L.lfx_extract_batch.argtypes = [H, C.POINTER(CloudView), C.c_int, C.POINTER(BatchResult)]
    L.lfx_synchronize.argtypes = [H]
    L.lfx_batch_status.argtypes = [H]
    L.lfx_fetch_counts.argtypes = [H, C.c_void_p, C.c_void_p]
    L.lfx_fetch_features.argtypes = [H, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    L.lfx_fetch_points.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lfx_fetch_rings.argtypes = [H, C.c_void_p]
    L.lfx_extract_scan.argtypes = [H, C.POINTER(CloudView), C.POINTER(ScanOutput)]
    L.lfx_label_to_color.argtypes = [C.c_uint8, C.c_void_p]
    L.lfx_host_alloc.argtypes = [C.c_size_t]
    L.lfx_host_alloc.restype = C.c_void_p
    L.lfx_host_alloc_on.argtypes = [H, C.c_size_t, C.POINTER(C.c_int)]
    L.lfx_host_alloc_on.restype = C.c_void_p
    L.lfx_host_free.argtypes = [C.c_void_p]
    L.lfx_host_free.restype = None
    L.lfx_device_alloc.argtypes = [H, C.c_size_t, C.POINTER(C.c_void_p)]
    L.lfx_device_free.argtypes = [H, C.c_void_p]
    L.lfx_memcpy_h2d.argtypes = [H, C.c_void_p, C.c_void_p, C.c_size_t]
    L.lfx_memcpy_d2h.argtypes = [H, C.c_void_p, C.c_void_p, C.c_size_t]
    L.lfx_kernel_launch_count.argtypes = [H]
    L.lfx_kernel_launch_count.restype = C.c_uint64
    L.lfx_set_stage_timing.argtypes = [H, C.c_int]
    L.lfx_last_stage_ms.argtypes = [H, C.c_void_p]
    L.lfx_last_batch_stats.argtypes = [H, C.POINTER(BatchStats)]
    L.lfx_convert_batch.argtypes = [H, C.POINTER(RawCloud), C.c_int, C.POINTER(ConvertResult)]
    L.lfx_last_convert_ms.argtypes = [H, C.POINTER(C.c_float)]
    L.lfx_converted_view.argtypes = [H, C.c_int, C.POINTER(CloudView)]
    L.lfx_converted_views.argtypes = [H, C.POINTER(CloudView), C.c_int, C.POINTER(C.c_int)]
    L.lfx_fetch_converted.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    L.lfx_color_batch.argtypes = [H, C.POINTER(ColoredResult)]
    L.lfx_fetch_colored.argtypes = [H, C.c_int, C.c_void_p, C.c_size_t]
    L.lfx_topic_layout.argtypes = [C.c_int, C.POINTER(PointFieldC), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.lfx_map_add_batch.argtypes = [H, C.POINTER(Pose), C.c_int, C.c_void_p, C.POINTER(C.c_uint64)]
    L.lfx_map_size.argtypes = [H, C.POINTER(C.c_uint64)]
    L.lfx_map_fetch.argtypes = [H, C.c_uint64, C.c_uint64, C.c_void_p]
    L.lfx_map_clear.argtypes = [H]
    L.lfx_map_gate.argtypes = [C.POINTER(Pose), C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(Pose), C.c_void_p]
    L.lfx_map_set_state.argtypes = [H, C.c_int, C.POINTER(Pose)]
    L.lfx_pose_diff_is_small.argtypes = [C.POINTER(Pose), C.POINTER(Pose), C.c_double, C.c_double]
    L.lfx_synth_named.argtypes = [C.c_char_p, C.POINTER(SynthSpec)]
    L.lfx_synth_scan_host.argtypes = [C.POINTER(SynthSpec), C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32)]
    L.lfx_synth_batch_device.argtypes = [H, C.POINTER(SynthSpec), C.c_uint64, C.c_int, C.c_void_p]
    L.lfx_loc_set_map.argtypes = [H, C.c_int, C.c_void_p, C.c_uint64, C.c_int]
    for fn in (L.lfx_loc_edge, L.lfx_loc_surface):
        fn.argtypes = [H, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(Pose), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.lfx_loc_release.argtypes = [H]
    S = C.c_void_p
    L.lfx_shard_range.argtypes = [C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.lfx_shard_unique_id.argtypes = [C.c_void_p]
    L.lfx_shard_create.argtypes = [H, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.POINTER(S)]
    L.lfx_shard_create_local.argtypes = [C.POINTER(H), C.c_int, C.c_uint64, C.POINTER(S)]
    L.lfx_shard_exchange.argtypes = [S]
    L.lfx_shard_finish.argtypes = [S, C.POINTER(ShardResult)]
    L.lfx_shard_fetch.argtypes = [S, C.c_void_p, C.c_void_p]
    L.lfx_shard_info.argtypes = [S, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.lfx_shard_last_error.argtypes = [S]
    L.lfx_shard_last_error.restype = C.c_char_p
    L.lfx_shard_destroy.argtypes = [S]
    L.lfx_shard_destroy.restype = None
    _lib = L
    return L
