"""CPU restatement (numpy, plain loops) of the localization package's residual build — TEST INFRASTRUCTURE, never
the product. SURVEY.md 8(f-4): per edge / surface feature of a scan, k nearest map points -> line / plane model ->
one Jacobian block + residual of the LOAM problem (x 40 iterations per scan in the consumer).

Follows, function by function:
  knn                 localization/src/kdtree.cpp:42-55 (nanoflann findNeighbors, metric_L2 over doubles: the squared
                      distance is the left-to-right sum of the squared coordinate differences; ascending distance)
  transform           Eigen::Isometry3d * p with the rotation of Eigen::Quaterniond::toRotationMatrix
  center / mean_cov   localization/src/edge.cpp:37-48
  principal           Eigen::SelfAdjointEigenSolver<Matrix3d>::computeDirect (edge.hpp:108-111: column 2 = eigenvector
                      of the largest eigenvalue)
  drp_dq              rotationlib/src/jacobian/quaternion.cpp:36-52 (Sola, eq. 174)
  edge_row            localization/src/edge.cpp:64-83 (MakeEdgeJacobianRow, MakeEdgeResidual), edge.hpp:100-119
  plane_coefficients  localization/include/lidar_feature_localization/surface.hpp:78-82 + math.hpp:36-40
                      (HouseholderQR::solve of X w = -1)
  surface_row         surface.hpp:84-92 (MakeJacobianRow), :46-60 (signed distance), :116-137

Pinned: neighbour INDEX SETS against the vendored nanoflann compiled in place (oracle/_ref/libref_knn.so,
tests/test_loc_oracle.py) and against the reference's own known-answer vectors (test_kdtree.cpp, test_edge.cpp:
Center, CalcMeanAndCovariance, TripletCross; test_math.cpp: SolveLinear). Eigen (mean / covariance reductions,
computeDirect, HouseholderQR) is third party and absent from /root/reference and from this image: its published
algorithms are restated here and parity with Eigen's exact rounding is UNPINNED — the tests compare within 1e-9 and,
for the principal axis, up to the sign (which cancels in J^T J and J^T r)."""
from __future__ import annotations

import numpy as np


def knn(map_xyz: np.ndarray, queries: np.ndarray, k: int):
    """Brute force; ties broken by the smaller index. Returns (idx [n, k] int64, d2 [n, k])."""
    m = np.asarray(map_xyz, np.float64)
    q = np.asarray(queries, np.float64)
    idx = np.zeros((len(q), k), np.int64)
    d2s = np.zeros((len(q), k), np.float64)
    for i in range(len(q)):
        d = q[i] - m
        d2 = d[:, 0] * d[:, 0]
        d2 = d2 + d[:, 1] * d[:, 1]
        d2 = d2 + d[:, 2] * d[:, 2]
        order = np.lexsort((np.arange(len(m)), d2))[:k]
        idx[i], d2s[i] = order, d2[order]
    return idx, d2s


def rotation_matrix(q_xyzw) -> np.ndarray:
    """Eigen::QuaternionBase::toRotationMatrix."""
    x, y, z, w = (float(v) for v in q_xyzw)
    tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1.0 - (txx + tyy)]])


def transform(q_xyzw, t, p) -> np.ndarray:
    return rotation_matrix(q_xyzw) @ np.asarray(p, np.float64) + np.asarray(t, np.float64)


def center(X) -> np.ndarray:
    X = np.asarray(X, np.float64)
    return X.sum(axis=0) / X.shape[0]


def mean_cov(X):
    X = np.asarray(X, np.float64)
    mean = center(X)
    D = X - mean
    return mean, (D.T @ D) / X.shape[0]


def triplet_cross(p0, p1, p2) -> np.ndarray:
    p0, p1, p2 = (np.asarray(v, np.float64) for v in (p0, p1, p2))
    return np.cross(p2 - p1, np.cross(p0 - p1, p0 - p2))


def _roots(m):
    """internal::direct_selfadjoint_eigenvalues<...,3>::computeRoots (trigonometric closed form, ascending)."""
    s_inv3, s_sqrt3 = 1.0 / 3.0, np.sqrt(3.0)
    c0 = (m[0, 0] * m[1, 1] * m[2, 2] + 2.0 * m[1, 0] * m[2, 0] * m[2, 1] - m[0, 0] * m[2, 1] * m[2, 1]
          - m[1, 1] * m[2, 0] * m[2, 0] - m[2, 2] * m[1, 0] * m[1, 0])
    c1 = m[0, 0] * m[1, 1] - m[1, 0] * m[1, 0] + m[0, 0] * m[2, 2] - m[2, 0] * m[2, 0] + m[1, 1] * m[2, 2] - m[2, 1] * m[2, 1]
    c2 = m[0, 0] + m[1, 1] + m[2, 2]
    c2_over_3 = c2 * s_inv3
    a_over_3 = max((c2 * c2_over_3 - c1) * s_inv3, 0.0)
    half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1))
    q = max(a_over_3 * a_over_3 * a_over_3 - half_b * half_b, 0.0)
    rho = np.sqrt(a_over_3)
    theta = np.arctan2(np.sqrt(q), half_b) * s_inv3
    ct, st = np.cos(theta), np.sin(theta)
    return np.array([c2_over_3 - rho * (ct + s_sqrt3 * st), c2_over_3 - rho * (ct - s_sqrt3 * st), c2_over_3 + 2.0 * rho * ct])


def _extract_kernel(mat):
    """extract_kernel: the null-space direction of a rank-2 symmetric 3x3 from the cross products of its columns."""
    i0 = int(np.argmax(np.abs(np.diag(mat))))
    rep = mat[:, i0].copy()
    c0 = np.cross(rep, mat[:, (i0 + 1) % 3])
    c1 = np.cross(rep, mat[:, (i0 + 2) % 3])
    n0, n1 = float(c0 @ c0), float(c1 @ c1)
    res = c0 / np.sqrt(n0) if n0 > n1 else c1 / np.sqrt(n1)
    return res, rep


def eigen_direct(C):
    """SelfAdjointEigenSolver<Matrix3d>::computeDirect: (eigenvalues ascending, eigenvectors in columns)."""
    C = np.asarray(C, np.float64)
    eps = np.finfo(np.float64).eps
    shift = np.trace(C) / 3.0
    sm = C - shift * np.eye(3)
    sm = np.tril(sm) + np.tril(sm, -1).T          # the solver reads the lower triangle
    scale = np.abs(sm).max()
    if scale > 0.0:
        sm = sm / scale
    ev = _roots(sm)
    vec = np.eye(3)
    if ev[2] - ev[0] > eps:
        d0, d1 = ev[2] - ev[1], ev[1] - ev[0]
        k, l = 0, 2
        if d0 > d1:
            k, l = 2, 0
            d0 = d1
        tmp = sm - ev[k] * np.eye(3)
        vec[:, k], rep = _extract_kernel(tmp)
        if d0 <= 2.0 * eps * d1:
            v = rep - (vec[:, k] @ rep) * rep
            vec[:, l] = v / np.linalg.norm(v)
        else:
            tmp = sm - ev[l] * np.eye(3)
            vec[:, l], _ = _extract_kernel(tmp)
        v1 = np.cross(vec[:, 2], vec[:, 0])
        vec[:, 1] = v1 / np.linalg.norm(v1)
    return ev * scale + shift, vec


def hat(p) -> np.ndarray:
    x, y, z = (float(v) for v in p)
    return np.array([[0.0, -z, y], [z, 0.0, -x], [-y, x, 0.0]])


def drp_dq(q_xyzw, p) -> np.ndarray:
    """[3, 4]: derivative of R(q) p by (w, x, y, z)."""
    p = np.asarray(p, np.float64)
    v = np.asarray(q_xyzw[:3], np.float64)
    w = float(q_xyzw[3])
    half = np.zeros((3, 4))
    half[:, 0] = w * p + np.cross(v, p)
    half[:, 1:] = (v @ p) * np.eye(3) + np.outer(v, p) - np.outer(p, v) - w * hat(p)
    return 2.0 * half


def edge_row(q_xyzw, t, p0, neighbors):
    """One edge feature: (jacobian [3, 7], residual [3]). neighbors: [k, 3] map points (doubles)."""
    mean, cov = mean_cov(neighbors)
    _, vec = eigen_direct(cov)
    principal = vec[:, 2]
    p0 = np.asarray(p0, np.float64)
    p1, p2 = mean - principal, mean + principal
    K = hat(p2 - p1)
    J = np.concatenate([K @ drp_dq(q_xyzw, p0), K], axis=1)
    p = transform(q_xyzw, t, p0)
    return J, np.cross(p - p1, p - p2)


def householder_qr_solve(A, b) -> np.ndarray:
    """Eigen::HouseholderQR<MatrixXd>(A).solve(b) for a full-column-rank m x n system, m >= n (no pivoting)."""
    A = np.array(A, np.float64)
    c = np.array(b, np.float64)
    m, n = A.shape
    for k in range(n):
        x = A[k:, k].copy()
        tail2 = float(x[1:] @ x[1:])
        c0 = x[0]
        if tail2 <= np.finfo(np.float64).tiny:        # makeHouseholder: nothing to annihilate
            tau, beta, ess = 0.0, c0, np.zeros(len(x) - 1)
        else:
            beta = np.sqrt(c0 * c0 + tail2)
            if c0 >= 0.0:
                beta = -beta
            ess = x[1:] / (c0 - beta)
            tau = (beta - c0) / beta
        v = np.concatenate([[1.0], ess])
        for j in range(k + 1, n):                     # applyHouseholderOnTheLeft on the remaining columns
            A[k:, j] -= tau * v * float(v @ A[k:, j])
        c[k:] -= tau * v * float(v @ c[k:])           # ... and on the right-hand side (Q^T b)
        A[k, k] = beta
        A[k + 1:, k] = 0.0
    x = np.zeros(n)
    for i in range(n - 1, -1, -1):
        x[i] = (c[i] - A[i, i + 1:n] @ x[i + 1:]) / A[i, i]
    return x


def plane_coefficients(X) -> np.ndarray:
    X = np.asarray(X, np.float64)
    return householder_qr_solve(X, -np.ones(X.shape[0]))


def surface_row(q_xyzw, t, p, neighbors):
    """One surface feature: (jacobian [7], residual scalar)."""
    p = np.asarray(p, np.float64)
    w = plane_coefficients(neighbors)
    norm = np.sqrt(float(w @ w))
    u = w / norm
    J = np.concatenate([u @ drp_dq(q_xyzw, p), u])
    x = transform(q_xyzw, t, p)
    return J, (float(w @ x) + 1.0) / norm


def edge_problem(map_xyz, scan_xyz, q_xyzw, t, k=15):
    """Edge<...>::Make, edge.hpp:88-124. Returns (neighbour idx [n, k], J [n, 3, 7], r [n, 3])."""
    m = np.asarray(map_xyz, np.float64)
    queries = np.array([transform(q_xyzw, t, p) for p in np.asarray(scan_xyz, np.float64)])
    idx, _ = knn(m, queries, k)
    J = np.zeros((len(queries), 3, 7))
    r = np.zeros((len(queries), 3))
    for i, p0 in enumerate(np.asarray(scan_xyz, np.float64)):
        J[i], r[i] = edge_row(q_xyzw, t, p0, m[idx[i]])
    return idx, J, r


def surface_problem(map_xyz, scan_xyz, q_xyzw, t, k=15):
    """Surface<...>::MakeFromDownsampled, surface.hpp:116-139 (the scan as it is AFTER the voxel down-sampling of
    :106-112, which is PCL's and not restated). Returns (neighbour idx [n, k], J [n, 7], r [n])."""
    m = np.asarray(map_xyz, np.float64)
    queries = np.array([transform(q_xyzw, t, p) for p in np.asarray(scan_xyz, np.float64)])
    idx, _ = knn(m, queries, k)
    J = np.zeros((len(queries), 7))
    r = np.zeros(len(queries))
    for i, p in enumerate(np.asarray(scan_xyz, np.float64)):
        J[i], r[i] = surface_row(q_xyzw, t, p, m[idx[i]])
    return idx, J, r
