"""Mapping accumulate on the device (lfx_map_add_batch: host pose gate + k_map_transform_add) against the oracle:
same frames selected, map points bit-identical, state carried across batches."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _trajectory(n, rng):
    """Poses along a curve: stretches of < 1 m steps (gated out), larger steps and turns in place."""
    poses, x, y, yaw = [], 0.0, 0.0, 0.0
    for i in range(n):
        step = rng.choice([0.0, 0.2, 0.45, 1.3])
        yaw += rng.choice([0.0, 0.0, 0.03, 0.25])
        x += step * np.cos(yaw); y += step * np.sin(yaw)
        roll = 0.02 * np.sin(i)
        q = np.array([np.sin(roll / 2) * np.cos(yaw / 2), np.sin(roll / 2) * np.sin(yaw / 2), np.cos(roll / 2) * np.sin(yaw / 2), np.cos(roll / 2) * np.cos(yaw / 2)])
        poses.append(((x, y, 0.1 * i), q / np.linalg.norm(q)))
    return poses


def test_map_accumulates_like_the_reference_map_builder():
    from lidar_feature_extraction_b200 import FeatureExtraction, MapBuilder, make_pose, synth
    from oracle import map_oracle as mo

    rng = np.random.default_rng(1)
    sp = synth.spec("vlp16")
    n1, n2 = 14, 9
    clouds = [synth.scan_host(sp, f) for f in range(n1 + n2)]
    clouds[3] = clouds[3][:0]          # an empty cloud: "Do nothing and continue" (map.hpp:117-120)
    traj = _trajectory(n1 + n2, rng)
    with FeatureExtraction() as fe:
        mb = MapBuilder(fe)
        assert mb.is_empty()
        edges, sel_got = [], []
        for lo, hi in ((0, n1), (n1, n1 + n2)):                      # two batches: the gate state carries over
            out = fe.extract_batch(clouds[lo:hi], fetch_points=False)
            edges += [out.scan_edges(s) for s in range(hi - lo)]
            sel_got += list(mb.add_batch([make_pose(*traj[i]) for i in range(lo, hi)]))
        got = mb.points()
        mats = [mo.pose_to_matrix(*p) for p in traj]
        want, sel = mo.build_map(mats, edges)
        assert list(sel) == sel_got and 3 < sel.sum() < n1 + n2 and not sel[3]
        assert got.shape == want.shape and len(mb) == want.shape[0]
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
        mb.clear()
        assert mb.is_empty()
        fe.extract_batch(clouds[:2], fetch_points=False)
        assert mb.add_batch([make_pose(*traj[0]), make_pose(*traj[0])]).tolist() == [True, False]   # first frame always added


def test_map_of_reference_test_vector():
    """test_map.cpp:65-98 through the device path: the clouds are whatever scan_edge holds, so the check is that an
    identity pose reproduces the edge cloud and a pure translation shifts it exactly."""
    from lidar_feature_extraction_b200 import FeatureExtraction, MapBuilder, make_pose, synth

    cloud = synth.scan_host(synth.spec("vlp16"), 0)
    with FeatureExtraction() as fe:
        mb = MapBuilder(fe)
        out = fe.extract_batch([cloud, cloud], fetch_points=False)
        e = out.scan_edges(0)
        mb.add_batch([make_pose((0, 0, 0), (0, 0, 0, 1)), make_pose((3, 0, 0), (0, 0, 0, 1))])
        m = mb.points()
        assert np.array_equal(m[: len(e)], e)
        shifted = e.copy()
        shifted[:, 0] = (e[:, 0].astype(np.float64) + 3.0).astype(np.float32)
        assert np.array_equal(m[len(e):], shifted)


def test_sharded_map_equals_the_sequential_map():
    """Two 'ranks' (two handles here) own frames [0, 9) and [9, 16): each installs the gate state of the frames
    before its shard (lfx_map_gate over the gathered sizes) and builds its part; the parts concatenate to the map
    one sequential MapBuilder produces."""
    from lidar_feature_extraction_b200 import FeatureExtraction, MapBuilder, make_pose, synth
    from lidar_feature_extraction_b200.mapping import gate_frames

    rng = np.random.default_rng(2)
    sp = synth.spec("vlp16")
    n, cut = 16, 9
    clouds = [synth.scan_host(sp, f) for f in range(n)]
    poses = [make_pose(*p) for p in _trajectory(n, rng)]
    with FeatureExtraction() as fe:
        mb = MapBuilder(fe)
        out = fe.extract_batch(clouds, fetch_points=False)
        sizes = out.counts[:, 0].copy()
        mb.add_batch(poses)
        whole = mb.points()
    parts = []
    for lo, hi in ((0, cut), (cut, n)):
        with FeatureExtraction() as fe:
            mb = MapBuilder(fe)
            _, empty, prev = gate_frames(poses[:lo], sizes[:lo])
            mb.set_state(empty, None if empty else prev)
            fe.extract_batch(clouds[lo:hi], fetch_points=False)
            mb.add_batch(poses[lo:hi])
            parts.append(mb.points())
    assert np.array_equal(np.concatenate(parts).view(np.uint32), whole.view(np.uint32))
