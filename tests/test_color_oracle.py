"""The colored_scan oracle against the known-answer vector of the reference's own test
(extraction/test/test_color_points.cpp:40-78). CPU only."""
import struct

import numpy as np

from oracle import color_oracle as co


def test_color_points_by_label_vector_of_reference_test():
    labels = np.array([0, 1, 2, 5, 6, 7], np.uint8)   # Default, Edge, EdgeNeighbor, OutOfRange, Occluded, ParallelBeam
    xyz = np.array([[k, 0, 0] for k in range(6)], np.float32)
    out = co.color_points_by_label(xyz, labels)
    got = []
    for rec in out:
        x, y, z, w, b, g, r, a = struct.unpack("<ffffBBBB12x", rec.tobytes())
        assert w == 1.0 and a == 255 and not rec[20:].any()
        got.append((x, y, z, r, g, b))
    assert got == [(0., 0., 0., 255, 255, 255), (1., 0., 0., 255, 0, 0), (2., 0., 0., 255, 63, 0),
                   (3., 0., 0., 127, 127, 127), (4., 0., 0., 255, 0, 255), (5., 0., 0., 0, 255, 0)]


def test_surface_labels_share_the_edge_colours():
    # color_points.cpp:49-57: Surface == Edge colour, SurfaceNeighbor == EdgeNeighbor colour
    assert tuple(co.LABEL_RGB[3]) == tuple(co.LABEL_RGB[1]) and tuple(co.LABEL_RGB[4]) == tuple(co.LABEL_RGB[2])


def test_invalid_label_raises():
    import pytest

    with pytest.raises(ValueError):
        co.color_points_by_label(np.zeros((1, 3), np.float32), np.array([8], np.uint8))
