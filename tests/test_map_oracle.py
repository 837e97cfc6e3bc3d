"""The mapping oracle against the known-answer vectors of the reference's own test (mapping/test/test_map.cpp). CPU only."""
import numpy as np

from oracle import map_oracle as mo


def _normalized(w, x, y, z):
    q = np.array([x, y, z, w], float)
    return q / np.linalg.norm(q)


def _mul(m, r):
    out = m.copy()
    out[:, :3] = m[:, :3] @ r
    return out


def test_pose_diff_is_sufficiently_small_vectors():
    """test_map.cpp:33-63."""
    pose0 = mo.pose_to_matrix((2.0, 1.0, -1.0), _normalized(1.0, 0.1, 0.1, -0.1))
    pose1 = pose0.copy()
    pose1[:, 3] += (1.0, 0.0, 0.0)
    assert not mo.pose_diff_is_sufficiently_small(pose0, pose1, 0.999999, 1e-8)
    assert mo.pose_diff_is_sufficiently_small(pose0, pose1, 1.1, 1e-8)
    dq = mo.pose_to_matrix((0, 0, 0), _normalized(1.0, 0.1, 0.1, 0.1))[:, :3]
    pose1 = _mul(pose0, dq)
    assert not mo.pose_diff_is_sufficiently_small(pose0, pose1, 1e-8, 0.1)
    assert mo.pose_diff_is_sufficiently_small(pose0, pose1, 1e-8, 0.2)


def test_transform_add_vector():
    """test_map.cpp:65-98: identity then a translation by (3, 0, 0)."""
    ident = mo.pose_to_matrix((0, 0, 0), (0, 0, 0, 1))
    shift = mo.pose_to_matrix((3, 0, 0), (0, 0, 0, 1))
    c0 = np.array([[0, 1, 0], [0, 1, 0]], np.float32)
    c1 = np.array([[0, 1, 0]], np.float32)
    got = np.concatenate([mo.transform_points(ident, c0), mo.transform_points(shift, c1)])
    assert got[:, :3].tolist() == [[0., 1., 0.], [0., 1., 0.], [3., 1., 0.]]
    assert (got[:, 3] == 1.0).all()


def test_gate_follows_map_builder_callback():
    """map.hpp:104-127: empty clouds are ignored, the first non-empty cloud is always added, later ones only after
    the pose moved >= 1 m or rotated (|dq.vec| >= 0.1) since the last ADDED frame."""
    def pose(x, yaw=0.0):
        return mo.pose_to_matrix((x, 0, 0), (0, 0, np.sin(yaw / 2), np.cos(yaw / 2)))

    ms = [pose(0.0), pose(0.0), pose(0.5), pose(0.99), pose(1.0), pose(1.2), pose(1.2, 0.3), pose(2.5), pose(9.0)]
    sizes = [0, 5, 5, 5, 5, 5, 5, 5, 0]
    sel, empty, prev = mo.gate_frames(ms, sizes)
    assert sel.tolist() == [False, True, False, False, True, False, True, True, False]
    assert not empty and prev is ms[7]
