"""The localization residual-build oracle (oracle/loc_oracle.py): neighbour sets against the vendored nanoflann
compiled in place (oracle/_ref/libref_knn.so), the reference's own known-answer vectors, and the restated Eigen
algorithms against numpy's. CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import loc_oracle as lo

REF_KNN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libref_knn.so")


def _ref_knn(map_xyz, queries, k, leaf=10):
    L = C.CDLL(REF_KNN)
    L.ref_knn.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    m = np.ascontiguousarray(map_xyz, np.float64)
    q = np.ascontiguousarray(queries, np.float64)
    idx = np.zeros((len(q), k), np.uint64)
    d2 = np.zeros((len(q), k), np.float64)
    assert L.ref_knn(m.ctypes.data, len(m), 3, leaf, q.ctypes.data, len(q), k, idx.ctypes.data, d2.ctypes.data) == 0
    return idx.astype(np.int64), d2


def test_kdtree_vector_of_the_reference_test():
    # localization/test/test_kdtree.cpp:38-77
    pts = np.array([[2, 0, 1], [2, 0, 0], [0, 0, 4], [0, 2, 4]], np.float64)
    idx, d2 = lo.knn(pts, np.zeros((1, 3)), 4)
    assert idx[0].tolist() == [1, 0, 2, 3] and d2[0].tolist() == [4.0, 5.0, 16.0, 20.0]
    idx, d2 = lo.knn(pts, np.zeros((1, 3)), 2)
    assert idx[0].tolist() == [1, 0] and d2[0].tolist() == [4.0, 5.0]


@pytest.mark.skipif(not os.path.exists(REF_KNN), reason="oracle/_ref is not built (no /root/reference here)")
@pytest.mark.parametrize("seed", range(4))
def test_knn_equals_nanoflann_compiled_in_place(seed):
    """Index sets AND order AND squared distances (bit for bit) of kdtree.cpp:42-55 on float-valued maps."""
    rng = np.random.default_rng(seed)
    n_map = [500, 3000, 20000, 64][seed]
    m = rng.normal(0, [20, 20, 2], size=(n_map, 3)).astype(np.float32).astype(np.float64)
    q = rng.normal(0, [20, 20, 2], size=(200, 3))
    for k in (1, 5, 15):
        got_i, got_d = lo.knn(m, q, k)
        want_i, want_d = _ref_knn(m, q, k)
        assert np.array_equal(got_d.view(np.uint64), want_d.view(np.uint64))
        assert np.array_equal(got_i, want_i)


def test_edge_vectors_of_the_reference_test():
    # localization/test/test_edge.cpp: TripletCross :45-65, Center :98-110, CalcMeanAndCovariance :112-135
    assert lo.triplet_cross([1, 2, 3], [4, 5, 6], [7, 8, 9]).tolist() == [0.0, 0.0, 0.0]
    assert lo.triplet_cross([0, 2, 1], [2, 0, 1], [1, 3, 0]).tolist() == [14.0, 2.0, -8.0]
    A = np.array([[4, 5, 1], [2, 0, 4], [6, 2, 2], [7, 5, 9], [0, 7, 7]], np.float64)
    assert lo.center(A).tolist() == [3.8, 3.8, 4.6]
    X = np.array([[2, 8, 9], [3, 5, 0], [6, 5, 5], [5, 2, 2]], np.float64)
    mean, cov = lo.mean_cov(X)
    assert np.array_equal(cov, np.array([[10., -9., -6.], [-9., 18., 21.], [-6., 21., 46.]]) / 4.0)
    # PrincipalComponents :67-96: points on the x axis
    Xl = np.array([[0.1 * i, 0.0, 0.0] for i in range(10)])
    ev, vec = lo.eigen_direct(lo.mean_cov(Xl)[1])
    assert np.linalg.norm(vec[:, 2] - [1, 0, 0]) <= 1e-8 and abs(ev[0]) <= 1e-8 and abs(ev[1]) <= 1e-8 and ev[2] > 0


def test_solve_linear_vectors_of_the_reference_test():
    # localization/test/test_math.cpp:36-70
    assert np.linalg.norm(lo.householder_qr_solve([[3, -1], [2, 3]], [7, 1]) - [2, -1]) <= 1e-7
    assert np.linalg.norm(lo.householder_qr_solve([[2, 1], [3, 3], [2, 4], [1, 2]], [7, 9, 4, 2]) - [4, -1]) <= 1e-7


def test_restated_eigen_algorithms_against_numpy():
    rng = np.random.default_rng(1)
    for _ in range(200):
        A = rng.normal(size=(3, 3)) * rng.uniform(0.01, 10)
        Cm = A @ A.T
        ev, vec = lo.eigen_direct(Cm)
        w, v = np.linalg.eigh(Cm)
        np.testing.assert_allclose(ev, w, rtol=1e-9, atol=1e-9 * abs(w).max())
        assert abs(abs(vec[:, 2] @ v[:, 2]) - 1.0) < 1e-8
        assert np.linalg.norm(Cm @ vec[:, 2] - ev[2] * vec[:, 2]) <= 1e-8 * max(1.0, abs(ev[2]))
    for _ in range(100):
        X = rng.normal(size=(5, 3)) + rng.normal(size=3) * 5
        np.testing.assert_allclose(lo.plane_coefficients(X), np.linalg.lstsq(X, -np.ones(5), rcond=None)[0], rtol=1e-8, atol=1e-10)


def test_jacobians_are_the_derivatives_of_the_residuals():
    """edge.cpp:64-83 / surface.hpp:84-92: the 3 x 7 (1 x 7) block is d residual / d (q_w, q_x, q_y, q_z, t), with the
    line / plane model held fixed (finite differences on the same neighbours)."""
    rng = np.random.default_rng(2)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    t = rng.normal(size=3)
    p0 = rng.normal(size=3) * 5
    nb = rng.normal(size=(15, 3)) * [3, 0.1, 0.1] + 4
    J, r = lo.edge_row(q, t, p0, nb)
    mean, cov = lo.mean_cov(nb)
    principal = lo.eigen_direct(cov)[1][:, 2]
    p1, p2 = mean - principal, mean + principal

    def res_e(qq, tt):
        w, v = qq[3], qq[:3]
        # R(q) p for a not necessarily unit quaternion, as the derivative DRpDq is taken (Sola eq. 174)
        p = (w * w - v @ v) * p0 + 2 * (v @ p0) * v + 2 * w * np.cross(v, p0) + tt
        return np.cross(p - p1, p - p2)

    h = 1e-6
    for c, (dq, dt) in enumerate([(np.eye(4)[[3, 0, 1, 2][c]] if c < 4 else np.zeros(4), np.eye(3)[c - 4] if c >= 4 else np.zeros(3)) for c in range(7)]):
        num = (res_e(q + h * dq, t + h * dt) - res_e(q - h * dq, t - h * dt)) / (2 * h)
        np.testing.assert_allclose(J[:, c], num, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r, res_e(q, t), rtol=1e-12, atol=1e-12)
    nbp = rng.normal(size=(5, 3)) * [2, 2, 0.01] + [1, 2, 3]
    Js, rs = lo.surface_row(q, t, p0, nbp)
    w_pl = lo.plane_coefficients(nbp)

    def res_s(qq, tt):
        w, v = qq[3], qq[:3]
        p = (w * w - v @ v) * p0 + 2 * (v @ p0) * v + 2 * w * np.cross(v, p0) + tt
        return (w_pl @ p + 1.0) / np.linalg.norm(w_pl)

    for c in range(7):
        dq = np.eye(4)[[3, 0, 1, 2][c]] if c < 4 else np.zeros(4)
        dt = np.eye(3)[c - 4] if c >= 4 else np.zeros(3)
        num = (res_s(q + h * dq, t + h * dt) - res_s(q - h * dq, t - h * dt)) / (2 * h)
        assert abs(Js[c] - num) <= 1e-5 * max(1.0, abs(num))
    assert abs(rs - res_s(q, t)) <= 1e-12


def test_golden_neighbour_lists_equal_the_oracle():
    """tests/golden/loc_knn.npz (nanoflann's lists, travels to the GPU box) against the brute-force oracle."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loc_knn.npz"))
    q, t = z["q_xyzw"], z["t"]
    for name in ("edge", "surface"):
        qs = np.array([lo.transform(q, t, p) for p in z[f"{name}.scan"].astype(np.float64)])
        for k in (5, 15):
            idx, d2 = lo.knn(z[f"{name}_map"], qs, k)
            assert np.array_equal(idx, z[f"{name}.idx{k}"].astype(np.int64))
            assert np.array_equal(d2.view(np.uint64), z[f"{name}.d2_{k}"].view(np.uint64))
