"""Localization residual build on the device (lfx_loc_*): neighbour lists against the reference's kd-tree (golden
vectors produced by the vendored nanoflann compiled in place, tests/golden/make_loc_golden.py), Jacobians and residuals
against the oracle (oracle/loc_oracle.py; Eigen's algorithms restated, tolerance 1e-9, principal axis up to its sign)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loc_knn.npz")


def _same_up_to_sign(J, r, Jw, rw, tol=1e-9):
    """An edge feature's block may come out negated (p1 and p2 swap with the sign of the principal axis)."""
    for i in range(len(J)):
        scale = max(1.0, np.abs(Jw[i]).max(), np.abs(rw[i]).max())
        same = np.abs(J[i] - Jw[i]).max() <= tol * scale and np.abs(r[i] - rw[i]).max() <= tol * scale
        flip = np.abs(J[i] + Jw[i]).max() <= tol * scale and np.abs(r[i] + rw[i]).max() <= tol * scale
        assert same or flip, (i, J[i], Jw[i])


@pytest.mark.parametrize("k", [5, 15])
def test_neighbours_equal_the_reference_kdtree_and_rows_equal_the_oracle(k):
    from lidar_feature_extraction_b200 import FeatureExtraction, LoamProblem
    from oracle import loc_oracle as lo

    z = np.load(GOLD)
    q, t = z["q_xyzw"], z["t"]
    with FeatureExtraction() as fe:
        prob = LoamProblem(fe, z["edge_map"], z["surface_map"], n_neighbors=k)
        J, r, nb = prob.make_edge(z["edge.scan"], q, t, want_neighbors=True)
        assert np.array_equal(nb, z[f"edge.idx{k}"]), "edge neighbour lists differ from nanoflann's"
        idx, Jw, rw = lo.edge_problem(z["edge_map"], z["edge.scan"], q, t, k)
        assert np.array_equal(idx, nb.astype(np.int64))
        _same_up_to_sign(J, r, Jw, rw)
        J, r, nb = prob.make_surface(z["surface.scan"], q, t, want_neighbors=True)
        assert np.array_equal(nb, z[f"surface.idx{k}"]), "surface neighbour lists differ from nanoflann's"
        idx, Jw, rw = lo.surface_problem(z["surface_map"], z["surface.scan"], q, t, k)
        for i in range(len(J)):
            scale = max(1.0, np.abs(Jw[i]).max(), abs(rw[i]))
            assert np.abs(J[i] - Jw[i]).max() <= 1e-8 * scale and abs(r[i] - rw[i]) <= 1e-8 * scale, i


def test_device_resident_inputs_ties_and_small_maps():
    """Maps / scans given as CUDA tensors; duplicate map points (ties go to the smaller index); a map of exactly k points."""
    import torch

    from lidar_feature_extraction_b200 import FeatureExtraction, LoamProblem
    from oracle import loc_oracle as lo

    rng = np.random.default_rng(3)
    base = rng.normal(0, 3, size=(700, 3)).astype(np.float32)
    m = np.concatenate([base, base[:300]])                 # 300 exact duplicates
    m4 = np.concatenate([m, np.ones((len(m), 1), np.float32)], axis=1)
    scan = rng.normal(0, 3, size=(257, 3)).astype(np.float32)
    s4 = np.concatenate([scan, np.ones((len(scan), 1), np.float32)], axis=1)
    q, t = np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3)
    with FeatureExtraction() as fe:
        prob = LoamProblem(fe, torch.from_numpy(m4).cuda(), torch.from_numpy(m4[:15].copy()).cuda(), n_neighbors=15)
        J, r, nb = prob.make_edge(torch.from_numpy(s4).cuda(), q, t, want_neighbors=True)
        idx, _ = lo.knn(m, scan.astype(np.float64), 15)
        assert np.array_equal(nb.astype(np.int64), idx)
        J, r, nb = prob.make_surface(s4, q, t, want_neighbors=True)
        assert (np.sort(nb, axis=1) == np.arange(15)).all()


def test_errors():
    from lidar_feature_extraction_b200 import ExtractionError, FeatureExtraction, LoamProblem

    with FeatureExtraction() as fe:
        prob = LoamProblem(fe, np.zeros((4, 3), np.float32), None, n_neighbors=15)
        with pytest.raises(ExtractionError):
            prob.make_edge(np.zeros((2, 3), np.float32), [0, 0, 0, 1], [0, 0, 0])   # fewer map points than neighbours
        with pytest.raises(ExtractionError):
            LoamProblem(fe, np.zeros((40, 3), np.float32), None, n_neighbors=17).make_edge(np.zeros((2, 3), np.float32), [0, 0, 0, 1], [0, 0, 0])
