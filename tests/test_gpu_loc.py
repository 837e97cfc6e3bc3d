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


@pytest.mark.parametrize("cell", ["", "0.25", "7.0"])
def test_grid_search_returns_the_exhaustive_neighbour_lists(cell, monkeypatch):
    """k_loc_knn_grid (uniform grid, shell-by-shell with an exact stopping rule) against the exhaustive k_loc_knn on a
    map with dense clusters, thin structures, duplicates, far outliers, and queries inside, at the rim and far outside
    the map's bounding box; also with cells much smaller / larger than the default (LFX_LOC_CELL)."""
    from lidar_feature_extraction_b200 import FeatureExtraction, LoamProblem

    rng = np.random.default_rng(11)
    wall = np.stack([rng.uniform(-40, 40, 20000), np.full(20000, 12.5), rng.uniform(-2, 6, 20000)], axis=1)
    blobs = np.concatenate([rng.normal(c, 0.7, size=(3000, 3)) for c in ([0, 0, 0], [25, -8, 1], [-30, 3, 2])])
    ground = np.stack([rng.uniform(-60, 60, 30000), rng.uniform(-60, 60, 30000), rng.normal(-1.8, 0.02, 30000)], axis=1)
    far = rng.uniform(-900, 900, size=(40, 3))
    m = np.concatenate([wall, blobs, ground, far, blobs[:500]]).astype(np.float32)
    scan = np.concatenate([rng.uniform(-50, 50, size=(3000, 3)), rng.normal([25, -8, 1], 0.5, size=(500, 3)),
                           rng.uniform(-3000, 3000, size=(200, 3)), m[::997].astype(np.float64)]).astype(np.float32)
    q = np.array([0.02, -0.01, 0.3, 1.0])
    q /= np.linalg.norm(q)
    t = np.array([1.5, -2.0, 0.3])
    lists = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("LFX_LOC_EXHAUSTIVE", mode)
        if cell:
            monkeypatch.setenv("LFX_LOC_CELL", cell)
        with FeatureExtraction() as fe:
            prob = LoamProblem(fe, m, m[:4000], n_neighbors=15)
            Je, re, nbe = prob.make_edge(scan, q, t, want_neighbors=True)
            Js, rs, nbs = prob.make_surface(scan[:500], q, t, want_neighbors=True)
        lists[mode] = (Je, re, nbe, Js, rs, nbs)
    for a, b in zip(lists["1"], lists["0"]):
        assert np.array_equal(a, b)
