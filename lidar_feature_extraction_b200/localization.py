"""Host-side mirror of the localization package's residual build over the CUDA library (first
slice): ``Edge<..>::Make`` (localization/include/lidar_feature_localization/edge.hpp:88-124) and
``Surface<..>::MakeFromDownsampled`` (surface.hpp:116-139) for all features of a scan at once. Nothing under
``oracle/`` is imported; the CUDA library does all the work (lfx_loc_*)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .extraction import ExtractionError, FeatureExtraction, _is_cuda_tensor


def _points(arr):
    """[n, 4] float32 x,y,z,1 (pcl::PointXYZ) as (pointer, n, memory, owner)."""
    if _is_cuda_tensor(arr):
        assert arr.dim() == 2 and arr.shape[1] == 4 and arr.element_size() == 4 and arr.is_contiguous()
        return arr.data_ptr(), int(arr.shape[0]), N.LFX_MEM_DEVICE, arr
    a = np.ascontiguousarray(arr, np.float32)
    if a.ndim != 2 or a.shape[1] not in (3, 4):
        raise ValueError("points must be [n, 3] or [n, 4] float32")
    if a.shape[1] == 3:
        a = np.concatenate([a, np.ones((len(a), 1), np.float32)], axis=1)
    return a.ctypes.data, len(a), N.LFX_MEM_HOST, a


class LoamProblem:
    """LOAMOptimizationProblem (loam_optimization_problem.hpp:50-89) for maps held on the device: ``make_edge`` /
    ``make_surface`` return what Edge::Make / Surface::MakeFromDownsampled return, for all features of the scan."""

    def __init__(self, extraction: FeatureExtraction, edge_map, surface_map, n_neighbors: int = 15):
        self.fe = extraction
        self._lib = N.lib()
        self.n_neighbors = int(n_neighbors)
        for kind, m in ((N.LFX_LOC_EDGE, edge_map), (N.LFX_LOC_SURFACE, surface_map)):
            if m is None:
                continue
            ptr, n, mem, keep = _points(m)
            rc = self._lib.lfx_loc_set_map(self.fe.handle, kind, ptr, n, mem)
            if rc != N.LFX_OK:
                raise ExtractionError(rc, self._lib.lfx_last_error(self.fe.handle).decode())

    def _run(self, fn, scan, q_xyzw, t, jw, rw, want_neighbors):
        ptr, n, mem, keep = _points(scan)
        pose = N.Pose((C.c_double * 3)(*[float(v) for v in t]), (C.c_double * 4)(*[float(v) for v in q_xyzw]))
        J = np.zeros((n,) + jw, np.float64)
        r = np.zeros((n,) + rw, np.float64)
        nb = np.zeros((n, self.n_neighbors), np.uint32) if want_neighbors else None
        rc = fn(self.fe.handle, ptr, n, mem, C.byref(pose), self.n_neighbors, J.ctypes.data, r.ctypes.data,
                nb.ctypes.data if nb is not None else None)
        if rc != N.LFX_OK:
            raise ExtractionError(rc, self._lib.lfx_last_error(self.fe.handle).decode())
        return (J, r, nb) if want_neighbors else (J, r)

    def make_edge(self, scan, q_xyzw, t, want_neighbors: bool = False):
        """Jacobians [n, 3, 7] (columns q_w,q_x,q_y,q_z,t_x,t_y,t_z), residuals [n, 3] (+ neighbour indices [n, k])."""
        return self._run(self._lib.lfx_loc_edge, scan, q_xyzw, t, (3, 7), (3,), want_neighbors)

    def make_surface(self, scan, q_xyzw, t, want_neighbors: bool = False):
        """Jacobians [n, 7], residuals [n] for an already down-sampled surface scan."""
        return self._run(self._lib.lfx_loc_surface, scan, q_xyzw, t, (7,), (), want_neighbors)
