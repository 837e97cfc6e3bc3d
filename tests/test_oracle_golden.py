"""Pins the C oracle (oracle/lfx_oracle.c) against the known-answer vectors of the reference's own
gtest files (extraction/test/*.cpp, cited per test). The vectors are data transcribed from those
tests; the functions under test are the oracle's restatements."""
import ctypes as C
import math

import numpy as np
import pytest

D, E, EN, S, SN, OOR, OCC, PB = range(8)  # point_label.hpp:32-42


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- test_curvature.cpp:34-66
def test_make_weight(oracle):
    for p, want in ((2, [1, 1, -4, 1, 1]), (3, [1, 1, 1, -6, 1, 1, 1])):
        out = np.zeros(2 * p + 1)
        oracle.lib.lfxo_make_weight(p, _ptr(out))
        assert out.tolist() == want


def test_calc_curvature(oracle):
    r = [1., 1., 2., 0., 1., 1., 0.]
    e0 = 1 + 1 + 2 * (-4) + 0 + 1
    e1 = 1 + 2 + 0 * (-4) + 1 + 1
    e2 = 2 + 0 + 1 * (-4) + 1 + 0
    assert oracle.curvature(r, 2).tolist() == [0., 0., e0 * e0, e1 * e1, e2 * e2, 0., 0.]
    r = [4., 4., 1., 2., 0., 5., 3., 6.]
    e0 = 4 + 4 + 1 + 2 * (-6) + 0 + 5 + 3
    e1 = 4 + 1 + 2 + 0 * (-6) + 5 + 3 + 6
    assert oracle.curvature(r, 3).tolist() == [0., 0., 0., e0 * e0, e1 * e1, 0., 0., 0.]


# ---------------------------------------------------------------- test_convolution.cpp:37-71
def test_convolution_1d(oracle):
    def conv(inp, w):
        inp, w = np.array(inp, float), np.array(w, float)
        out = np.zeros(len(inp))
        rc = oracle.lib.lfxo_convolution_1d(_ptr(inp), len(inp), _ptr(w), len(w), _ptr(out))
        return rc, out.tolist()

    assert conv([1., -1, 2., 0., 1], [1., 0., -1]) == (0, [0., -1., -1., 1., 0.])
    assert conv([1., -1, 2.], [1., 0., -1]) == (0, [0., -1., 0.])
    assert conv([2., 0.], [1., 0., -1])[0] == 1  # throws std::invalid_argument


# ---------------------------------------------------------------- test_algorithm.cpp:36-49
def test_argsort(oracle):
    assert oracle.argsort([0.3, 0.2, 1.0, 0.2, 0.0, 0.1]).tolist() == [4, 5, 1, 3, 0, 2]
    assert oracle.argsort([0.0] * 5).tolist() == [0, 1, 2, 3, 4]
    assert oracle.argsort([0.0] * 100).tolist() == list(range(100))  # index tie-break beyond n = 16


# ---------------------------------------------------------------- test_index_range.cpp:38-176
def test_index_range(oracle):
    assert oracle.index_range(0, 12, 3).tolist() == [0, 4, 8, 12]
    assert oracle.index_range(0, 14, 4).tolist() == [0, 3, 7, 10, 14]
    assert oracle.index_range(0, 3, 3).tolist() == [0, 1, 2, 3]
    assert oracle.index_range(1, 3, 3) is None  # ctor throws
    assert oracle.padded_index_range(17, 3, 1).tolist() == [1, 6, 11, 16]
    assert oracle.padded_index_range(20, 4, 2).tolist() == [2, 6, 10, 14, 18]


def test_sector_tables_of_survey(oracle):
    # boundaries at the BASELINE sensor widths, defaults P=5, 6 blocks
    assert oracle.padded_index_range(1800, 6, 5).tolist() == [5, 303, 601, 900, 1198, 1496, 1795]
    assert oracle.padded_index_range(2170, 6, 5).tolist() == [5, 365, 725, 1085, 1445, 1805, 2165]
    assert oracle.padded_index_range(2048, 6, 5).tolist() == [5, 344, 684, 1024, 1363, 1703, 2043]


# ---------------------------------------------------------------- test_fill.cpp:40-273
def _fill(oracle, fn, groups, a, b, label, init=None):
    link = oracle.links_from_groups(groups)
    labels = np.zeros(len(groups), np.uint8) if init is None else np.array(init, np.uint8)
    rc = getattr(oracle.lib, fn)(_ptr(labels), _ptr(link), len(groups), a, b, label)
    return rc, labels.tolist()


def test_fill_from_left(oracle):
    assert _fill(oracle, "lfxo_fill_from_left", [0] * 5, 1, 4, E) == (0, [D, E, E, E, D])
    assert _fill(oracle, "lfxo_fill_from_left", [0, 0, 0, 1, 1], 1, 5, E) == (0, [D, E, E, D, D])
    assert _fill(oracle, "lfxo_fill_from_left", [0, 0, 0], 1, 3, D)[0] == 0
    assert _fill(oracle, "lfxo_fill_from_left", [0, 0, 0], 1, 4, D)[0] == 1   # end_index > size
    assert _fill(oracle, "lfxo_fill_from_left", [0, 0, 0], 0, 2, D)[0] == 0
    assert _fill(oracle, "lfxo_fill_from_left", [0, 0, 0], -1, 2, D)[0] == 1  # begin_index < 0


def test_fill_from_right(oracle):
    assert _fill(oracle, "lfxo_fill_from_right", [0] * 5, 0, 2, E) == (0, [D, E, E, D, D])
    assert _fill(oracle, "lfxo_fill_from_right", [0] * 5, 1, 3, E) == (0, [D, D, E, E, D])
    assert _fill(oracle, "lfxo_fill_from_right", [0, 0, 0, 1, 1], 1, 4, E) == (0, [D, D, D, E, E])
    assert _fill(oracle, "lfxo_fill_from_right", [0, 0, 0], 1, 2, D)[0] == 0
    assert _fill(oracle, "lfxo_fill_from_right", [0, 0, 0], 1, 3, D)[0] == 1   # end_index >= size
    assert _fill(oracle, "lfxo_fill_from_right", [0, 0, 0], -1, 2, D)[0] == 0
    assert _fill(oracle, "lfxo_fill_from_right", [0, 0, 0], -2, 2, D)[0] == 1  # begin_index < -1


def test_fill_neighbors(oracle):
    f = lambda g, i: _fill(oracle, "lfxo_fill_neighbors", g, i, 2, EN)  # noqa: E731
    assert f([0] * 6, 3) == (0, [D, EN, EN, EN, EN, EN])
    assert f([0, 0, 1, 1, 1, 1, 2, 2], 3) == (0, [D, D, EN, EN, EN, EN, D, D])
    assert f([0] * 6, 1) == (0, [EN, EN, EN, EN, D, D])
    assert f([1, 1, 1, 0, 0, 0], 4) == (0, [D, D, D, EN, EN, EN])
    assert f([1, 1, 1, 0, 0, 0], 2) == (0, [EN, EN, EN, D, D, D])


# ---------------------------------------------------------------- test_label.cpp:76-129
def test_edge_label(oracle):
    def assign(curv, groups):
        link = oracle.links_from_groups(groups)
        labels = np.zeros(len(curv), np.uint8)
        c = np.array(curv, float)
        oracle.lib.lfxo_edge_assign(_ptr(labels), _ptr(c), _ptr(link), len(curv), 2, 1.5)
        return labels.tolist()

    assert assign([3, 1, 2, 1, 1, 4, 1, 1], [0] * 8) == [E, EN, EN, EN, EN, E, EN, EN]
    assert assign([0, 2, 4, 1, 0, 2, 1, 1], [0, 0, 0, 0, 1, 1, 1, 1]) == [EN, EN, E, EN, EN, E, EN, EN]


def test_surface_label_mirror(oracle):
    # no reference vector exists for SurfaceLabel::Assign; this is the mirrored
    # edge case: picking ascending with <= threshold
    link = oracle.links_from_groups([0] * 8)
    labels = np.zeros(8, np.uint8)
    c = np.array([1, 3, 2, 3, 3, 0, 3, 3], float)
    oracle.lib.lfxo_surface_assign(_ptr(labels), _ptr(c), _ptr(link), 8, 2, 2.5)
    assert labels.tolist() == [S, SN, SN, SN, SN, S, SN, SN]


# ---------------------------------------------------------------- test_occlusion.cpp:38-198
def _occl(oracle, pts, padding):
    x = np.array([p[0] for p in pts], np.float32)  # pcl::PointXYZ holds floats
    y = np.array([p[1] for p in pts], np.float32)
    link = oracle.links_from_points(x, y, 0.2)
    r = np.array([oracle.lib.lfxo_xy_norm(float(a), float(b)) for a, b in zip(x, y)])
    labels = np.zeros(len(pts), np.uint8)
    oracle.lib.lfxo_occlusion_from_left(_ptr(labels), _ptr(link), _ptr(r), len(pts), padding, 2.0)
    oracle.lib.lfxo_occlusion_from_right(_ptr(labels), _ptr(link), _ptr(r), len(pts), padding, 2.0)
    return labels.tolist()


def test_occlusion_from_left(oracle):
    assert _occl(oracle, [(4.03, 1.0), (8.04, 2.0), (8.05, 2.0), (8.06, 2.0)], 2) == [D, OCC, OCC, OCC]
    pts = [(4.00, 1.0), (4.01, 1.0), (4.02, 1.0), (4.03, 1.0), (8.04, 2.0), (8.05, 2.0), (8.06, 2.0), (8.07, 8.0), (8.08, 8.0)]
    assert _occl(oracle, pts, 1) == [D, D, D, D, OCC, OCC, D, D, D]
    assert _occl(oracle, pts, 3) == [D, D, D, D, OCC, OCC, OCC, D, D]


def test_occlusion_from_right(oracle):
    assert _occl(oracle, [(8.06, 2.0), (8.07, 2.0), (8.08, 2.0), (4.09, 1.0)], 2) == [OCC, OCC, OCC, D]
    pts = [(8.03, 8.0), (8.04, 2.0), (8.05, 2.0), (8.06, 2.0), (8.07, 2.0), (8.08, 2.0), (4.09, 1.0), (4.10, 1.0), (4.11, 1.0), (4.12, 1.0)]
    assert _occl(oracle, pts, 1) == [D, D, D, D, OCC, OCC, D, D, D, D]
    assert _occl(oracle, pts, 3) == [D, D, OCC, OCC, OCC, OCC, D, D, D, D]


# ---------------------------------------------------------------- test_out_of_range.cpp:34-56
def test_out_of_range(oracle):
    pts = [(1.9, 0.0), (2.0, 0.0), (0.0, 5.0), (0.0, 8.0), (0.0, 8.1)]
    r = np.array([oracle.lib.lfxo_xy_norm(float(np.float32(a)), float(np.float32(b))) for a, b in pts])
    labels = np.zeros(5, np.uint8)
    oracle.lib.lfxo_out_of_range(_ptr(labels), _ptr(r), 5, 2.0, 8.0)
    assert labels.tolist() == [OOR, D, D, D, OOR]


# ---------------------------------------------------------------- test_parallel_beam.cpp:35-74
def test_parallel_beam(oracle):
    r = np.array([8.0, 8.0, 2.0, 8.0, 8.0])
    for thr, want in ((3.0, [D] * 5), (2.9, [D, D, PB, D, D])):
        labels = np.zeros(5, np.uint8)
        oracle.lib.lfxo_parallel_beam(_ptr(labels), _ptr(r), 5, thr)
        assert labels.tolist() == want


# ---------------------------------------------------------------- test_neighbor.cpp:38-86, test_math.cpp:35-78
def test_is_neighbor_and_calc_radian(oracle):
    L = oracle.lib
    assert L.lfxo_is_neighbor(1., 1., 1., 1., 1e-7) == 1
    assert L.lfxo_is_neighbor(0., 1., 1., 0., math.pi / 2 + 1e-3) == 1
    assert L.lfxo_is_neighbor(0., 1., 1., 0., math.pi / 2 - 1e-3) == 0
    # NeighborCheckXY vector: (1,1) (0,1) (1,0) (1,0)
    assert L.lfxo_is_neighbor(1., 0., 1., 0., 1e-3) == 1
    assert L.lfxo_is_neighbor(1., 1., 0., 1., math.pi / 4 + 1e-3) == 1
    assert L.lfxo_is_neighbor(0., 1., 1., 0., math.pi / 4 + 1e-3) == 0
    out = C.c_double()
    cases = [((1, 1, 1, 1), 0.), ((-1, 1, -1, 1), 0.), ((1, 0, 0, 1), math.pi / 2), ((0, 1, 1, 0), math.pi / 2),
             ((1, -1, 1, 1), math.pi / 2), ((-1, -1, 1, 1), math.pi), ((1, 1, -1, -1), math.pi),
             ((-1, 1, 0, 1), math.pi / 4), ((0, 1, -1, 1), math.pi / 4)]
    for args, want in cases:
        assert L.lfxo_calc_radian(*map(float, args), C.byref(out)) == 0
        assert abs(out.value - want) < 1e-7
    assert L.lfxo_calc_radian(0., 0., 0., 0., C.byref(out)) == 1  # throws
    assert L.lfxo_xy_norm(0., 0.) == 0. and L.lfxo_xy_norm(-1., 0.) == 1. and L.lfxo_xy_norm(3., 4.) == 5.


# ---------------------------------------------------------------- test_ring.cpp:47-216
SPECIFIC = [((0, 0), (0, 0)), ((0, 0), (0, 1)), ((0, 0), (1, 0)), ((0, 1), (0, 0)), ((1, 0), (0, 0)), ((0, 0), (0, -1)),
            ((0, 0), (-1, 0)), ((0, -1), (0, 0)), ((-1, 0), (0, 0)), ((-1, 1), (-1, 1)), ((1, 1), (1, 1)),
            ((1, -1), (1, -1)), ((-1, -1), (-1, -1)), ((-1, 0), (-1, 0)), ((0, 1), (0, 1)), ((1, 0), (1, 0)),
            ((0, -1), (0, -1)), ((1, 0), (1, 1)), ((1, 0), (1, -1)), ((-1, 0), (-1, 1)), ((-1, 0), (-1, -1)),
            ((1, 1), (1, 0)), ((1, -1), (1, 0)), ((-1, 1), (-1, 0)), ((-1, -1), (-1, 0)), ((-1, 1), (1, 1)),
            ((1, -1), (1, 1)), ((1, 1), (-1, 1)), ((1, 1), (1, -1)), ((-1, 1), (-1, -1)), ((1, -1), (-1, -1)),
            ((-1, -1), (-1, 1)), ((-1, -1), (1, -1))]


def test_polar_comparator_specific_values(oracle):
    for (ax, ay), (bx, by) in SPECIFIC:
        want = int(math.atan2(ay, ax) < math.atan2(by, bx))
        assert oracle.lib.lfxo_polar_less_f64(float(ax), float(ay), float(bx), float(by)) == want
        assert oracle.lib.lfxo_polar_less_f32(float(ax), float(ay), float(bx), float(by)) == want


def test_polar_comparator_random(oracle):
    rng = np.random.default_rng(0)
    pts = rng.uniform(-1, 1, size=(10000, 2))
    idx = oracle.sort_by_polar_angle(pts[:, 0], pts[:, 1], dtype=np.float64)
    want = np.argsort(np.arctan2(pts[:, 1], pts[:, 0]), kind="stable")
    assert np.array_equal(idx, want)


def test_sort_by_atan2(oracle):
    x = [1., 1., 1., 0., 0., -1.]
    y = [1., 0., -1., 1., -1., -1.]
    assert oracle.sort_by_polar_angle(x, y, dtype=np.float64).tolist() == [5, 4, 2, 1, 0, 3]


def test_extract_angle_sorted_rings(oracle):
    # test_ring.cpp:191-216 through the scan-level entry (rings {0:2, 1:3, 2:3} with padding 1 -> all kept)
    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    ring = [0, 0, 1, 1, 1, 2, 2, 2]
    x = [1., 1., 1., 1., 0., 1., 0., -1.]
    y = [1., 0., -1., 0., 1., 1., -1., -1.]
    cloud = synth.make_cloud(x, y, np.zeros(8), ring)
    res = oracle.extract_scan(cloud, ob.default_params(padding=1, n_blocks=1))
    assert res.ring_ids.tolist() == [0, 1, 2]
    assert res.sorted_src.tolist() == [1, 0, 2, 3, 4, 7, 6, 5]


# ---------------------------------------------------------------- test_color_points.cpp:40-78
def test_label_to_color(oracle):
    want = {D: (255, 255, 255), E: (255, 0, 0), EN: (255, 63, 0), OOR: (127, 127, 127), OCC: (255, 0, 255), PB: (0, 255, 0),
            S: (255, 0, 0), SN: (255, 63, 0)}
    for label, rgb in want.items():
        out = np.zeros(3, np.uint8)
        assert oracle.lib.lfxo_label_to_color(label, _ptr(out)) == 0
        assert tuple(out) == rgb
    assert oracle.lib.lfxo_label_to_color(8, _ptr(np.zeros(3, np.uint8))) == 1
