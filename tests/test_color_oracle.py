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


# ---- pinned to the reference's own ColorPointsByLabel (oracle/ref_driver.cpp: ref_color_scan) ------------------

def _xyzrgb(records):
    """x,y,z [n,3] f32 and r,g,b [n,3] u8 of [n,32] pcl::PointXYZRGB records (b,g,r,a at bytes 16..19)."""
    rec = np.ascontiguousarray(records).reshape(-1, 32)
    return rec[:, 0:12].copy().view(np.float32).reshape(-1, 3), rec[:, [18, 17, 16]]


def test_colored_scan_oracle_reproduces_the_reference_on_the_golden_scans(oracle):
    """tests/golden/*.npz hold colored_scan as the compiled reference produced it (x,y,z,r,g,b per point)."""
    import glob
    import os

    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob
    from test_golden import FIXTURES, PARAMSETS

    assert FIXTURES
    for path in FIXTURES:
        z = np.load(path)
        cloud = z["cloud"]
        x, y, zz, _, _ = (np.ascontiguousarray(a) for a in synth.fields(cloud))
        for pname, kw in PARAMSETS.items():
            got = co.colored_scan(x, y, zz, oracle.extract_scan(cloud, ob.default_params(**kw)))
            gx, grgb = _xyzrgb(got)
            assert np.array_equal(gx.view(np.uint32), z[f"{pname}.colored_xyz"].view(np.uint32)), (os.path.basename(path), pname)
            assert np.array_equal(grgb, z[f"{pname}.colored_rgb"]), (os.path.basename(path), pname)
            assert (got[:, 12:16].copy().view(np.float32) == 1.0).all() and (got[:, 19] == 255).all() and not got[:, 20:].any()


def test_colored_scan_oracle_equals_the_compiled_reference_on_sensor_scans(oracle, reference_stable):
    """Whole sensor-shaped scans (incl. the tunnel scene, which carries all eight labels) through ref_color_scan."""
    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    for name, frame in (("vlp16", 1), ("hdl64", 2), ("hdl32", 3)):
        cloud = synth.scan_host(synth.spec(name), frame)
        x, y, z, _, _ = (np.ascontiguousarray(a) for a in synth.fields(cloud))
        for kw in (dict(), dict(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0)):
            prm = ob.default_params(**kw)
            want_xyz, want_rgb = reference_stable.color_scan(cloud, prm)
            gx, grgb = _xyzrgb(co.colored_scan(x, y, z, oracle.extract_scan(cloud, prm)))
            assert np.array_equal(gx.view(np.uint32), want_xyz.view(np.uint32)) and np.array_equal(grgb, want_rgb), (name, kw)
