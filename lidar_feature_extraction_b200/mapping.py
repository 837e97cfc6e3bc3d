"""Host-side mirror of the mapping package's accumulate step over the CUDA library.

Mirrors ``MapBuilder`` / ``Map`` (mapping/include/lidar_feature_mapping/map.hpp:62-145) for the offline, batched
sequence: ``add_batch(poses)`` does for every frame of the last extracted batch what ``MapBuilder::Callback`` does
for one (scan_edge, pose) pair - pose gate on the host, transform + append in ``k_map_transform_add``. Saving the
map as a PCD file (``Map::Save``, pcl::io) is storage and stays with the caller: ``points()`` returns the x,y,z,1
float32 array.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _native as N
from .extraction import ExtractionError, FeatureExtraction


def make_pose(position, orientation_xyzw) -> N.Pose:
    p = N.Pose()
    p.position[:] = [float(v) for v in position]
    p.orientation[:] = [float(v) for v in orientation_xyzw]
    return p


def pose_diff_is_sufficiently_small(pose0: N.Pose, pose1: N.Pose, translation_threshold: float, rotation_threshold: float) -> bool:
    """map.hpp:50-60."""
    return N.lib().lfx_pose_diff_is_small(C.byref(pose0), C.byref(pose1), translation_threshold, rotation_threshold) == 1


def gate_frames(poses: Sequence[N.Pose], n_edge, map_empty: bool = True, prev: N.Pose | None = None):
    """MapBuilder::Callback's decisions for a frame sequence (lfx_map_gate, host only).
    Returns (selected bool array, map_empty, prev pose)."""
    n = len(poses)
    arr = (N.Pose * max(n, 1))(*poses)
    sizes = np.ascontiguousarray(n_edge, np.uint32)
    sel = np.zeros(n, np.uint8)
    empty = C.c_int(1 if map_empty else 0)
    p = N.Pose()
    if prev is not None:
        C.memmove(C.byref(p), C.byref(prev), C.sizeof(N.Pose))
    rc = N.lib().lfx_map_gate(arr, sizes.ctypes.data, n, C.byref(empty), C.byref(p), sel.ctypes.data)
    if rc != N.LFX_OK:
        raise ValueError("lfx_map_gate: bad arguments")
    return sel.astype(bool), bool(empty.value), p


class MapBuilder:
    def __init__(self, extraction: FeatureExtraction):
        self.extraction = extraction
        self._lib = N.lib()

    def _check(self, rc):
        if rc != N.LFX_OK:
            raise ExtractionError(rc, self._lib.lfx_last_error(self.extraction.handle).decode())

    def add_batch(self, poses: Sequence[N.Pose]) -> np.ndarray:
        """One pose per scan of the last extracted batch; returns the bool array of frames that were added."""
        arr = (N.Pose * max(len(poses), 1))(*poses)
        sel = np.zeros(len(poses), np.uint8)
        n = C.c_uint64()
        self._check(self._lib.lfx_map_add_batch(self.extraction.handle, arr, len(poses), sel.ctypes.data, C.byref(n)))
        return sel.astype(bool)

    def is_empty(self) -> bool:
        return len(self) == 0

    def __len__(self) -> int:
        n = C.c_uint64()
        self._check(self._lib.lfx_map_size(self.extraction.handle, C.byref(n)))
        return int(n.value)

    def points(self) -> np.ndarray:
        out = np.zeros((len(self), 4), np.float32)
        self._check(self._lib.lfx_map_fetch(self.extraction.handle, 0, out.shape[0], out.ctypes.data))
        return out

    def set_state(self, map_empty: bool, prev: N.Pose | None = None):
        """Install the gate state (sharded driver: the state after the frames before this rank's shard)."""
        self._check(self._lib.lfx_map_set_state(self.extraction.handle, 1 if map_empty else 0, C.byref(prev) if prev is not None else None))

    def clear(self):
        self._check(self._lib.lfx_map_clear(self.extraction.handle))
