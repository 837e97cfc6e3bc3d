"""ctypes bindings for the CHECKERS: liblfx_oracle.so (plain-C restatement) and, when built,
oracle/_ref/libref_{verbatim,stable}.so (the reference's own sources compiled in place).

TEST INFRASTRUCTURE, NOT PRODUCT. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` leg may import this module. The CUDA product path
(lidar_feature_extraction_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liblfx_oracle.so")
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"


class Params(C.Structure):
    """Layout shared by lfxo_params, ref_params and lfx_params (hyper_parameter.hpp:32-65)."""

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


def default_params(**kw) -> Params:
    """Compiled defaults, hyper_parameter.hpp:35-43."""
    p = Params(5, 2.0, 0.3, 0.02, 0.05, 0.05, 0.1, 100.0, 6)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def launch_yaml_params(**kw) -> Params:
    """lidar_feature_launch/config/lidar_feature_extraction.param.yaml:3-10 (surface stays default)."""
    p = Params(2, 3.0, 0.3, 0.02, 50.0, 0.05, 0.1, 1000.0, 6)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


class Cloud(C.Structure):
    _fields_ = [
        ("data", C.c_void_p),
        ("n_points", C.c_int),
        ("point_step", C.c_int),
        ("off_x", C.c_int),
        ("off_y", C.c_int),
        ("off_z", C.c_int),
        ("off_ring", C.c_int),
        ("ring_datatype", C.c_int),
    ]


@dataclass
class ScanResult:
    ring_ids: np.ndarray
    ring_sizes: np.ndarray
    ring_skipped: np.ndarray
    sorted_src: np.ndarray   # source index of every kept point, (ring asc, angle asc)
    labels: np.ndarray       # u8 per kept point, 255 for skipped rings
    curvature: np.ndarray    # f64 per kept point
    edge_idx: np.ndarray     # positions into sorted order
    surface_idx: np.ndarray


def build(ref: bool | None = None) -> None:
    """Build the C restatement; and the reference-backed libraries when /root/reference exists."""
    subprocess.run(["make", "-s", "-f", os.path.join(HERE, "Makefile"), "oracle"], check=True)
    if ref is None:
        ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "extraction", "src"))
    if ref:
        subprocess.run(["make", "-s", "-j4", "-f", os.path.join(HERE, "Makefile"), "ref"], check=True)


_I = C.c_int
_D = C.c_double
_F = C.c_float
_P = C.c_void_p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """liblfx_oracle.so."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.lfxo_polar_less_f32.argtypes = [_F] * 4
        L.lfxo_polar_less_f64.argtypes = [_D] * 4
        L.lfxo_sort_by_polar_angle_f32.argtypes = [_P, _P, _P, _I]
        L.lfxo_sort_by_polar_angle_f64.argtypes = [_P, _P, _P, _I]
        L.lfxo_xy_norm.argtypes = [_D, _D]
        L.lfxo_xy_norm.restype = _D
        L.lfxo_calc_radian.argtypes = [_D, _D, _D, _D, C.POINTER(_D)]
        L.lfxo_is_neighbor.argtypes = [_F, _F, _F, _F, _D]
        L.lfxo_degree_to_radian.argtypes = [_D]
        L.lfxo_degree_to_radian.restype = _D
        L.lfxo_make_weight.argtypes = [_I, _P]
        L.lfxo_convolution_1d.argtypes = [_P, _I, _P, _I, _P]
        L.lfxo_curvature.argtypes = [_P, _I, _I, _P]
        L.lfxo_index_range.argtypes = [_I, _I, _I, _P]
        L.lfxo_padded_index_range.argtypes = [_I, _I, _I, _P]
        L.lfxo_argsort.argtypes = [_P, _I, _P]
        L.lfxo_fill_from_left.argtypes = [_P, _P, _I, _I, _I, C.c_uint8]
        L.lfxo_fill_from_right.argtypes = [_P, _P, _I, _I, _I, C.c_uint8]
        L.lfxo_fill_neighbors.argtypes = [_P, _P, _I, _I, _I, C.c_uint8]
        L.lfxo_edge_assign.argtypes = [_P, _P, _P, _I, _I, _D]
        L.lfxo_surface_assign.argtypes = [_P, _P, _P, _I, _I, _D]
        L.lfxo_occlusion_from_left.argtypes = [_P, _P, _P, _I, _I, _D]
        L.lfxo_occlusion_from_right.argtypes = [_P, _P, _P, _I, _I, _D]
        L.lfxo_out_of_range.argtypes = [_P, _P, _I, _D, _D]
        L.lfxo_parallel_beam.argtypes = [_P, _P, _I, _D]
        L.lfxo_label_to_color.argtypes = [C.c_uint8, _P]
        L.lfxo_extract_ring.argtypes = [_P, _P, _I, C.POINTER(Params), _P, _P]
        L.lfxo_extract_scan.argtypes = [C.POINTER(Cloud), C.POINTER(Params), _I] + [_P] * 7 + [_P, _P, _P, _P]
        L.lfxo_extract_batch_counts.argtypes = [C.POINTER(Cloud), _I, C.POINTER(Params), _I, _P]

    # -- helpers mirroring the reference's test fixtures --
    @staticmethod
    def links_from_groups(groups) -> np.ndarray:
        """NeighborCheckDebug (neighbor.hpp:116-136): neighbours iff same group id."""
        g = np.asarray(groups)
        link = np.zeros(max(len(g), 1), dtype=np.uint8)
        link[: len(g) - 1] = g[:-1] == g[1:]
        return link

    def links_from_points(self, x, y, radian_threshold: float) -> np.ndarray:
        n = len(x)
        link = np.zeros(max(n, 1), dtype=np.uint8)
        for i in range(n - 1):
            r = self.lib.lfxo_is_neighbor(float(x[i]), float(y[i]), float(x[i + 1]), float(y[i + 1]), radian_threshold)
            if r < 0:
                raise ValueError("both norms zero")
            link[i] = r
        return link

    def curvature(self, ranges, padding: int):
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        out = np.zeros(len(r), dtype=np.float64)
        rc = self.lib.lfxo_curvature(_ptr(r), len(r), padding, _ptr(out))
        return None if rc else out

    def padded_index_range(self, size: int, n_blocks: int, padding: int):
        out = np.zeros(n_blocks + 1, dtype=np.int32)
        rc = self.lib.lfxo_padded_index_range(size, n_blocks, padding, _ptr(out))
        return None if rc else out

    def index_range(self, start: int, end: int, n_blocks: int):
        out = np.zeros(n_blocks + 1, dtype=np.int32)
        rc = self.lib.lfxo_index_range(start, end, n_blocks, _ptr(out))
        return None if rc else out

    def argsort(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        out = np.zeros(len(v), dtype=np.int32)
        self.lib.lfxo_argsort(_ptr(v), len(v), _ptr(out))
        return out

    def sort_by_polar_angle(self, x, y, dtype=np.float32):
        x = np.ascontiguousarray(x, dtype=dtype)
        y = np.ascontiguousarray(y, dtype=dtype)
        idx = np.arange(len(x), dtype=np.int32)
        f = self.lib.lfxo_sort_by_polar_angle_f32 if dtype == np.float32 else self.lib.lfxo_sort_by_polar_angle_f64
        f(_ptr(x), _ptr(y), _ptr(idx), len(x))
        return idx

    def extract_ring(self, x, y, prm: Params):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        n = len(x)
        labels = np.zeros(max(n, 1), dtype=np.uint8)
        curv = np.zeros(max(n, 1), dtype=np.float64)
        rc = self.lib.lfxo_extract_ring(_ptr(x), _ptr(y), n, C.byref(prm), _ptr(labels), _ptr(curv))
        return rc, labels[:n], curv[:n]

    def extract_scan(self, cloud_bytes: np.ndarray, prm: Params, point_step=32, off_x=0, off_y=4, off_z=8,
                     off_ring=20, ring_datatype=4) -> ScanResult:
        data = np.ascontiguousarray(cloud_bytes).view(np.uint8).reshape(-1)
        n = data.size // point_step
        cloud = Cloud(data.ctypes.data, n, point_step, off_x, off_y, off_z, off_ring, ring_datatype)
        return _run_scan(lambda *a: self.lib.lfxo_extract_scan(C.byref(cloud), C.byref(prm), *a), n)

    def extract_batch_counts(self, scans: list[np.ndarray], prm: Params, n_threads: int, point_step=32,
                             off_x=0, off_y=4, off_z=8, off_ring=20, ring_datatype=4) -> np.ndarray:
        arr = (Cloud * len(scans))()
        keep = []
        for i, s in enumerate(scans):
            d = np.ascontiguousarray(s).view(np.uint8).reshape(-1)
            keep.append(d)
            arr[i] = Cloud(d.ctypes.data, d.size // point_step, point_step, off_x, off_y, off_z, off_ring, ring_datatype)
        counts = np.zeros(2 * len(scans), dtype=np.int32)
        rc = self.lib.lfxo_extract_batch_counts(arr, len(scans), C.byref(prm), n_threads, _ptr(counts))
        if rc:
            raise RuntimeError(f"oracle batch failed rc={rc}")
        return counts.reshape(-1, 2)


def _run_scan(call, n: int) -> ScanResult:
    cap = max(n, 1)
    n_rings = C.c_int(0)
    n_e = C.c_int(0)
    n_s = C.c_int(0)
    ring_ids = np.zeros(cap, dtype=np.int32)
    ring_sizes = np.zeros(cap, dtype=np.int32)
    ring_skipped = np.zeros(cap, dtype=np.int32)
    sorted_src = np.zeros(cap, dtype=np.int32)
    labels = np.zeros(cap, dtype=np.uint8)
    curv = np.zeros(cap, dtype=np.float64)
    e_idx = np.zeros(cap, dtype=np.int32)
    s_idx = np.zeros(cap, dtype=np.int32)
    m = call(cap, C.addressof(n_rings), _ptr(ring_ids), _ptr(ring_sizes), _ptr(ring_skipped), _ptr(sorted_src),
             _ptr(labels), _ptr(curv), C.addressof(n_e), _ptr(e_idx), C.addressof(n_s), _ptr(s_idx))
    if m < 0:
        raise RuntimeError("scan extraction failed (capacity)")
    r = n_rings.value
    return ScanResult(ring_ids[:r].copy(), ring_sizes[:r].copy(), ring_skipped[:r].copy(), sorted_src[:m].copy(),
                      labels[:m].copy(), curv[:m].copy(), e_idx[: n_e.value].copy(), s_idx[: n_s.value].copy())


class Reference:
    """oracle/_ref/libref_{verbatim,stable}.so — the reference's own code (32-byte PointXYZIR input only)."""

    def __init__(self, variant: str = "stable"):
        path = os.path.join(REF_DIR, f"libref_{variant}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.variant = variant
        self.lib = L = C.CDLL(path)
        L.ref_extract_scan.argtypes = [_P, _I, C.POINTER(Params), _I] + [_P] * 7 + [_P, _P, _P, _P]
        L.ref_color_scan.argtypes = [_P, _I, C.POINTER(Params), _P, _P]
        L.ref_polar_less.argtypes = [_F] * 4
        L.ref_curvature.argtypes = [_P, _I, _I, _P]
        L.ref_boundaries.argtypes = [_I, _I, _I, _P]
        L.ref_is_neighbor.argtypes = [_F, _F, _F, _F, _D]
        L.ref_variant.restype = C.c_char_p
        assert L.ref_variant().decode() == variant

    @staticmethod
    def available(variant: str = "stable") -> bool:
        return os.path.exists(os.path.join(REF_DIR, f"libref_{variant}.so"))

    def extract_scan(self, cloud_bytes: np.ndarray, prm: Params) -> ScanResult:
        data = np.ascontiguousarray(cloud_bytes).view(np.uint8).reshape(-1)
        assert data.size % 32 == 0
        n = data.size // 32
        return _run_scan(lambda *a: self.lib.ref_extract_scan(data.ctypes.data, n, C.byref(prm), *a), n)

    def color_scan(self, cloud_bytes: np.ndarray, prm: Params):
        """colored_scan of the reference (ColorPointsByLabel per ring, feature_extraction.cpp:153): (xyz [m, 3] f32,
        rgb [m, 3] u8), rings ascending."""
        data = np.ascontiguousarray(cloud_bytes).view(np.uint8).reshape(-1)
        n = data.size // 32
        xyz = np.zeros((max(n, 1), 3), np.float32)
        rgb = np.zeros((max(n, 1), 3), np.uint8)
        m = self.lib.ref_color_scan(data.ctypes.data, n, C.byref(prm), _ptr(xyz), _ptr(rgb))
        return xyz[:m].copy(), rgb[:m].copy()

    def curvature(self, ranges, padding: int):
        r = np.ascontiguousarray(ranges, dtype=np.float64)
        out = np.zeros(len(r), dtype=np.float64)
        rc = self.lib.ref_curvature(_ptr(r), len(r), padding, _ptr(out))
        return None if rc else out

    def boundaries(self, size: int, n_blocks: int, padding: int):
        out = np.zeros(n_blocks + 1, dtype=np.int32)
        rc = self.lib.ref_boundaries(size, n_blocks, padding, _ptr(out))
        return None if rc else out
