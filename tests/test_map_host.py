"""Host half of the mapping row (pose gate arithmetic in liblfx.so, no GPU work) against the oracle and the
reference's own vectors (mapping/test/test_map.cpp:33-63). CPU only."""
import numpy as np

from oracle import map_oracle as mo


def _q(w, x, y, z):
    q = np.array([x, y, z, w], float)
    return q / np.linalg.norm(q)


def _compose(q0, q1):
    x0, y0, z0, w0 = q0
    x1, y1, z1, w1 = q1
    return np.array([w0 * x1 + x0 * w1 + y0 * z1 - z0 * y1, w0 * y1 - x0 * z1 + y0 * w1 + z0 * x1,
                     w0 * z1 + x0 * y1 - y0 * x1 + z0 * w1, w0 * w1 - x0 * x1 - y0 * y1 - z0 * z1])


def test_reference_vectors():
    from lidar_feature_extraction_b200 import make_pose, pose_diff_is_sufficiently_small as small

    q0 = _q(1.0, 0.1, 0.1, -0.1)
    p0 = make_pose((2.0, 1.0, -1.0), q0)
    p1 = make_pose((3.0, 1.0, -1.0), q0)
    assert not small(p0, p1, 0.999999, 1e-8)
    assert small(p0, p1, 1.1, 1e-8)
    p1 = make_pose((2.0, 1.0, -1.0), _compose(q0, _q(1.0, 0.1, 0.1, 0.1)))
    assert not small(p0, p1, 1e-8, 0.1)
    assert small(p0, p1, 1e-8, 0.2)


def test_random_pose_pairs_agree_with_the_oracle():
    from lidar_feature_extraction_b200 import make_pose, pose_diff_is_sufficiently_small as small

    rng = np.random.default_rng(0)
    n_small = 0
    for _ in range(2000):
        q0 = rng.normal(size=4); q0 /= np.linalg.norm(q0)
        dq = np.array([*rng.normal(0, rng.choice([0.02, 0.08, 0.5]), 3), 1.0]); dq /= np.linalg.norm(dq)
        q1 = _compose(q0, dq)
        t0 = rng.normal(0, 10, 3)
        t1 = t0 + rng.normal(0, rng.choice([0.3, 0.6, 3.0]), 3)
        want = mo.pose_diff_is_sufficiently_small(mo.pose_to_matrix(t0, q0), mo.pose_to_matrix(t1, q1), 1.0, 0.1)
        assert small(make_pose(t0, q0), make_pose(t1, q1), 1.0, 0.1) == want
        n_small += want
    assert 100 < n_small < 1900


def test_gate_sequence_and_shard_state_agree_with_the_oracle():
    """lfx_map_gate over a whole sequence == the oracle's MapBuilder::Callback loop, and gating a shard after
    installing the state of its prefix gives the same decisions (what a rank of the sharded driver does)."""
    from lidar_feature_extraction_b200 import make_pose
    from lidar_feature_extraction_b200.mapping import gate_frames

    rng = np.random.default_rng(4)
    n = 400
    pos = np.cumsum(rng.choice([0.0, 0.3, 0.6, 1.5], size=(n, 1)) * rng.normal(size=(n, 3)), axis=0)
    yaw = np.cumsum(rng.choice([0.0, 0.02, 0.3], size=n))
    quats = [np.array([0, 0, np.sin(a / 2), np.cos(a / 2)]) for a in yaw]
    sizes = rng.choice([0, 7, 300], size=n, p=[0.1, 0.45, 0.45])
    poses = [make_pose(p, q) for p, q in zip(pos, quats)]
    want, _, _ = mo.gate_frames([mo.pose_to_matrix(p, q) for p, q in zip(pos, quats)], sizes)
    got, empty, _ = gate_frames(poses, sizes)
    assert got.tolist() == want.tolist() and not empty and 20 < got.sum() < n
    for lo in (0, 1, 137, 399):
        _, e0, p0 = gate_frames(poses[:lo], sizes[:lo])
        tail, _, _ = gate_frames(poses[lo:], sizes[lo:], map_empty=e0, prev=p0)
        assert tail.tolist() == want[lo:].tolist()
