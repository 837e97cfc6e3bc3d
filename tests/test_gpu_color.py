"""colored_scan on the device (k_color_scan through lfx_color_batch) and the topic layouts against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(clouds, hp=None):
    from helpers import oracle_params
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, synth
    from oracle import binding as ob
    from oracle import color_oracle as co

    hp = hp or HyperParameters()
    oracle = ob.Oracle()
    with FeatureExtraction(hp, want_sorted_src=True) as fe:
        fe.extract_batch(clouds)
        got = fe.colored_scans()
    assert len(got) == len(clouds)
    for cloud, g in zip(clouds, got):
        x, y, z, _, _ = (np.ascontiguousarray(a) for a in synth.fields(cloud))
        want = co.colored_scan(x, y, z, oracle.extract_scan(cloud, oracle_params(ob, hp)))
        assert g.shape == want.shape
        assert np.array_equal(g, want)


def test_colored_scan_regular_and_ragged_scans():
    from lidar_feature_extraction_b200 import synth

    clouds = [synth.scan_host(synth.spec("vlp16"), frame=1), synth.scan_host(synth.spec("hdl64"), frame=2),
              synth.scan_host(synth.spec("hdl32"), frame=3)]
    _run(clouds)


def test_colored_scan_drops_sparse_and_skipped_rings():
    """Rings with < padding + 1 points are removed (ring.cpp:46-59), rings too short for the convolution / sectors
    and rings with two adjacent zero-XY points throw (feature_extraction.cpp:154-156): none of them is coloured."""
    import adversarial as adv

    clouds = [adv.ragged_scan(21, [700, 4, 12, 900, 0, 300, 9], shuffle="interleave"),
              adv.ragged_scan(22, [400, 500, 640], shuffle="random", ring_ids=[7, 2, 90], zero_xy=0),
              adv.ragged_scan(23, [300, 800, 5], shuffle="rotate", zero_xy=40)]
    _run(clouds)


def test_topic_layouts_and_callback_messages():
    from lidar_feature_extraction_b200 import FeatureExtraction, PointCloud2, synth
    from lidar_feature_extraction_b200 import _native as N
    from lidar_feature_extraction_b200.extraction import topic_fields

    f, step = topic_fields(N.LFX_TOPIC_SCAN_EDGE)
    assert [(x.name, x.offset, x.datatype) for x in f] == [("x", 0, 7), ("y", 4, 7), ("z", 8, 7)] and step == 16
    assert topic_fields(N.LFX_TOPIC_SCAN_SURFACE) == (f, 16)
    f, step = topic_fields(N.LFX_TOPIC_COLORED_SCAN)
    assert [(x.name, x.offset, x.datatype) for x in f] == [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("rgb", 16, 7)] and step == 32
    cloud = synth.scan_host(synth.spec("vlp16"), frame=0)
    with FeatureExtraction(want_sorted_src=True) as fe:
        out = fe.callback(PointCloud2.from_wire(cloud, stamp=(12, 34)))
    for topic in ("scan_edge", "scan_surface", "colored_scan"):
        m = out[topic]
        assert m.frame_id == "lidar_feature_base_link" and m.stamp == (12, 34) and m.height == 1
        assert m.data.shape == (m.width, m.point_step)
    assert out["colored_scan"].width == len(cloud)
