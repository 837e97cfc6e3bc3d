"""CPU restatement (numpy / plain Python doubles) of the mapping package's accumulate step — TEST INFRASTRUCTURE,
never the product.

Follows mapping/include/lidar_feature_mapping/map.hpp: PoseDiffIsSufficientlySmall (:50-60), Map::TransformAdd
(:68-74), MapBuilder::Callback (:104-133, thresholds :92-93) and GetIsometry3d (lib/src/ros_msg.cpp:33-38).

Third-party arithmetic the reference calls and /root/reference does not contain (named, restated from their
published sources, parity with them UNPINNED here except for the reference's own known-answer vectors,
mapping/test/test_map.cpp:33-100, replayed in tests/test_map_oracle.py):
  * tf2_eigen (ros-humble) fromMsg(Pose, Isometry3d): Translation3d(p) * Quaterniond(w, x, y, z);
  * Eigen 3.4 QuaternionBase::toRotationMatrix, the matrix -> quaternion assignment (trace branch / largest
    diagonal branch), Isometry inverse (R^T, -R^T t) and product, Vector3d::norm;
  * PCL 1.12 pcl::transformPointCloud(cloud, out, Affine3d): detail::Transformer<double>::se3 -
    out = (float)(m0*x + m1*y + m2*z + m3) per row, evaluated left to right in double, data[3] = 1
    (pcl/common/impl/transforms.hpp); the host build has no FMA contraction (no -march flags,
    mapping/CMakeLists.txt).
"""
from __future__ import annotations

import math

import numpy as np

TRANSLATION_THRESHOLD = 1.0   # map.hpp:92
ROTATION_THRESHOLD = 0.1      # map.hpp:93


def pose_to_matrix(position, orientation_xyzw) -> np.ndarray:
    """tf2::fromMsg + Quaterniond::toRotationMatrix -> 3x4 [R | t] in double."""
    x, y, z, w = (float(v) for v in orientation_xyzw)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    m = np.zeros((3, 4))
    m[0, 0] = 1.0 - (tyy + tzz); m[0, 1] = txy - twz; m[0, 2] = txz + twy
    m[1, 0] = txy + twz; m[1, 1] = 1.0 - (txx + tzz); m[1, 2] = tyz - twx
    m[2, 0] = txz - twy; m[2, 1] = tyz + twx; m[2, 2] = 1.0 - (txx + tyy)
    m[:, 3] = [float(v) for v in position]
    return m


def _rotation_to_quat_vec(r):
    """Eigen's matrix -> quaternion assignment; returns (x, y, z) = dq.vec()."""
    t = r[0][0] + r[1][1] + r[2][2]
    q = [0.0, 0.0, 0.0]
    if t > 0.0:
        t = math.sqrt(t + 1.0)
        t = 0.5 / t
        q[0] = (r[2][1] - r[1][2]) * t
        q[1] = (r[0][2] - r[2][0]) * t
        q[2] = (r[1][0] - r[0][1]) * t
    else:
        i = 0
        if r[1][1] > r[0][0]:
            i = 1
        if r[2][2] > r[i][i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = math.sqrt(r[i][i] - r[j][j] - r[k][k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[j] = (r[j][i] + r[i][j]) * t
        q[k] = (r[k][i] + r[i][k]) * t
    return q


def pose_diff_is_sufficiently_small(m0, m1, translation_threshold, rotation_threshold) -> bool:
    """map.hpp:50-60 on 3x4 [R | t] matrices: d = pose0^-1 * pose1."""
    r0 = [[float(m0[a][b]) for b in range(3)] for a in range(3)]
    r1 = [[float(m1[a][b]) for b in range(3)] for a in range(3)]
    t0 = [float(m0[a][3]) for a in range(3)]
    t1 = [float(m1[a][3]) for a in range(3)]
    # inverse of an isometry: linear = R0^T, translation = -(R0^T t0)
    rt = [[r0[b][a] for b in range(3)] for a in range(3)]
    it = [-((rt[a][0] * t0[0] + rt[a][1] * t0[1]) + rt[a][2] * t0[2]) for a in range(3)]
    d = [[(rt[a][0] * r1[0][b] + rt[a][1] * r1[1][b]) + rt[a][2] * r1[2][b] for b in range(3)] for a in range(3)]
    dt = [((rt[a][0] * t1[0] + rt[a][1] * t1[1]) + rt[a][2] * t1[2]) + it[a] for a in range(3)]
    q = _rotation_to_quat_vec(d)
    dt_norm = math.sqrt((dt[0] * dt[0] + dt[1] * dt[1]) + dt[2] * dt[2])
    dq_norm = math.sqrt((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2])
    return dt_norm < translation_threshold and dq_norm < rotation_threshold


def gate_frames(matrices, sizes, map_empty=True, prev=None):
    """MapBuilder::Callback, map.hpp:104-127, over a sequence: frame i is added iff its cloud is not empty and
    (the map is empty or the pose moved enough since the last ADDED frame). prev_transform_ starts as an
    uninitialised Eigen::Isometry3d in the reference; it is only read once the map is non-empty, i.e. after
    it has been assigned. Returns (selected bool array, map_empty, prev)."""
    sel = np.zeros(len(sizes), bool)
    for i, (m, n) in enumerate(zip(matrices, sizes)):
        if n == 0:
            continue
        if not map_empty and pose_diff_is_sufficiently_small(prev, m, TRANSLATION_THRESHOLD, ROTATION_THRESHOLD):
            continue
        sel[i] = True
        prev = m
        map_empty = False
    return sel, map_empty, prev


def transform_points(m, xyz: np.ndarray) -> np.ndarray:
    """pcl::transformPointCloud with an Affine3d (detail::Transformer<double>::se3) -> [n, 4] float32 x,y,z,1."""
    p = np.asarray(xyz, np.float32)[:, :3].astype(np.float64)
    out = np.ones((p.shape[0], 4), np.float32)
    for a in range(3):
        out[:, a] = (((m[a, 0] * p[:, 0] + m[a, 1] * p[:, 1]) + m[a, 2] * p[:, 2]) + m[a, 3]).astype(np.float32)
    return out


def build_map(matrices, clouds, map_empty=True, prev=None):
    """clouds: list of [n_i, >=3] float32 scan_edge clouds. Returns ([N, 4] float32 map, selected)."""
    sel, _, _ = gate_frames(matrices, [len(c) for c in clouds], map_empty, prev)
    parts = [transform_points(m, c) for m, c, s in zip(matrices, clouds, sel) if s]
    return (np.concatenate(parts, axis=0) if parts else np.zeros((0, 4), np.float32)), sel
