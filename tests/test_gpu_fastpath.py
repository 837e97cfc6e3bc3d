"""GPU parity of the fast path (k_probe_layout + k_extract_sectors: one warp per ring-sector) against the
oracle, through the C ABI. Bit-exact labels, order and feature clouds; curvature within 1e-6 relative.
Regular scans (sensor firing order, rotated monotone rings) must take the fast path; scans that break
one of its hypotheses must be caught by its checks, redone by the general path, and still match."""
import numpy as np
import pytest

import adversarial as adv
from helpers import compare_scan, oracle_params

pytestmark = pytest.mark.gpu


def _fe(hp=None, diag=True, **kw):
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters

    return FeatureExtraction(hp or HyperParameters(), device=0, want_sorted_src=diag, want_curvature=diag, **kw)


def _hp(**kw):
    from lidar_feature_extraction_b200 import HyperParameters

    return HyperParameters(**kw)


def regular_scan(seed, n_rings, width, kinds=None, direction="cw", ring_ids=None, zero_xy=0):
    """Firing-order scan: column-major, every ring `width` points, each ring a rotation of an ascending
    ('ccw') or descending ('cw') azimuth sequence with its own rotation offset."""
    from lidar_feature_extraction_b200 import synth

    rng = np.random.default_rng(seed)
    X = np.zeros((width, n_rings), np.float32)
    Y = np.zeros((width, n_rings), np.float32)
    for k in range(n_rings):
        kind = kinds[k % len(kinds)] if kinds else adv.KINDS[(seed + k) % len(adv.KINDS)]
        x, y = adv._ring_points(rng, width, kind)
        for _ in range(zero_xy):
            j = int(rng.integers(0, width))
            x[j] = 0.0
            y[j] = 0.0
        order = np.arange(width)
        if direction == "cw":
            order = order[::-1]
        order = np.roll(order, int(rng.integers(0, width)))
        X[:, k], Y[:, k] = x[order], y[order]
    ids = np.arange(n_rings) if ring_ids is None else np.asarray(ring_ids)
    R = np.broadcast_to(ids.astype(np.uint16), (width, n_rings))
    Z = rng.normal(0, 1, size=(width, n_rings)).astype(np.float32)
    return synth.make_cloud(X.ravel(), Y.ravel(), Z.ravel(), np.ascontiguousarray(R).ravel())


def _check(oracle, hp, clouds, diag=True, **kw):
    from oracle import binding as ob

    with _fe(hp, diag=diag, **kw) as fe:
        out = fe.extract_batch(clouds)
        stats = fe.batch_stats()
    for s, cloud in enumerate(clouds):
        compare_scan(out, s, cloud, oracle.extract_scan(cloud, oracle_params(ob, hp)))
    return out, stats


@pytest.mark.parametrize("paramset", ["default", "yaml"])
@pytest.mark.parametrize("sensor,kidx", [("vlp16", 0), ("hdl32", 2), ("os128", 1)])
def test_sensor_shapes_take_the_fast_path(oracle, sensor, kidx, paramset):
    from lidar_feature_extraction_b200 import default_params, launch_yaml_params, synth

    hp = default_params() if paramset == "default" else launch_yaml_params()
    sp = synth.spec(sensor)
    clouds = [synth.scan_host(sp, f) for f in range(3)]
    out, stats = _check(oracle, hp, clouds)
    assert stats["general_scans"] == 0 and sum(stats["fast_rings"]) == 3 * sp.n_rings, stats
    if paramset == "default":
        assert stats["fast_rings"][kidx] == 3 * sp.n_rings, stats
    # and without the diagnostic outputs (the production instantiation of the kernel)
    _, stats2 = _check(oracle, hp, clouds, diag=False)
    assert stats2 == stats


@pytest.mark.parametrize("direction", ["cw", "ccw"])
@pytest.mark.parametrize("width", [64, 100, 299, 640, 1024, 1800, 2048, 2232])
def test_regular_scans_of_every_kind(oracle, width, direction):
    """Every ring kind (ramp = deepest selection chains, plateau = all ties, gaps = broken links, near/far =
    out of range, steps = occlusion) in firing order, all window classes, both spin directions."""
    clouds = [regular_scan(width + s, len(adv.KINDS), width, kinds=adv.KINDS, direction=direction) for s in range(2)]
    for hp in (_hp(), _hp(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0),
               _hp(surface_threshold=1e9, edge_threshold=1e-9)):
        out, stats = _check(oracle, hp, clouds)
        assert stats["general_scans"] == 0 and sum(stats["fast_rings"]) == 2 * len(adv.KINDS), (width, stats)
        assert (out.rings["order_path"] == 0).all()


@pytest.mark.parametrize("blocks", [1, 2, 5, 6, 7, 31])
def test_regular_scans_other_sector_counts(oracle, blocks):
    width = {1: 300, 2: 640, 5: 1500, 6: 2100, 7: 2400, 31: 2304}[blocks]
    clouds = [regular_scan(blocks, 7, width, direction="cw")]
    for P in (5, 2):
        out, stats = _check(oracle, _hp(n_blocks=blocks, padding=P), clouds)
        assert stats["general_scans"] == 0, stats


def test_sparse_ring_ids_and_ring_datatypes(oracle):
    from lidar_feature_extraction_b200 import PointCloud2, PointField, synth
    from oracle import binding as ob

    base = regular_scan(9, 5, 700, ring_ids=[40, 3, 17, 99, 4])
    hp = _hp()
    out, stats = _check(oracle, hp, [base])
    assert stats["general_scans"] == 0 and sum(stats["fast_rings"]) == 5
    x, y, z, _, ring = (np.ascontiguousarray(a) for a in synth.fields(base))
    want = oracle.extract_scan(base, oracle_params(ob, hp))
    # (point_step, x offset, ring offset, ring datatype, regular path?): the regular (strided) path fetches whole
    # 32-byte points, so it needs 32-byte aligned records with the ring word at +20 from x (the deployed layout);
    # every other layout is bucketed and runs on the indexed sector path - same results
    for step, ox, oring, rdt, npdt, fast in ((32, 0, 20, 4, np.uint16, True), (64, 32, 52, 6, np.uint32, True),
                                             (32, 0, 21, 2, np.uint8, True), (48, 16, 44, 6, np.uint32, False),
                                             (32, 0, 13, 2, np.uint8, False), (64, 32, 18, 4, np.uint16, False),
                                             (64, 32, 2, 4, np.uint16, False)):
        buf = np.zeros((len(x), step), np.uint8)
        for k, a in enumerate((x, y, z)):
            buf[:, ox + 4 * k: ox + 4 * k + 4] = a.view(np.uint8).reshape(-1, 4)
        rb = ring.astype(npdt).view(np.uint8).reshape(len(x), -1)
        buf[:, oring:oring + rb.shape[1]] = rb
        msg = PointCloud2(data=buf, point_step=step, width=len(x),
                          fields=[PointField("x", ox, 7), PointField("y", ox + 4, 7), PointField("z", ox + 8, 7), PointField("ring", oring, rdt)])
        with _fe(hp) as fe:
            out = fe.extract_batch([msg])
            st = fe.batch_stats()
        compare_scan(out, 0, base, want)
        assert st["general_scans"] == (0 if fast else 1), (step, oring, st)


def _points(cloud):
    return cloud.view(np.float32).reshape(-1, 8)


def test_broken_hypotheses_fall_back_to_the_general_path(oracle):
    """Each corruption keeps the scan looking regular to the probe (or not) but violates what the sector
    kernel verifies; the scan must be redone by the general path and match the oracle."""
    hp = _hp()
    n_rings, width = 6, 900
    good = regular_scan(1, n_rings, width)
    cases = {}
    c = good.copy()   # a ring id that breaks the period somewhere in the middle of the scan
    c.view(np.uint16).reshape(-1, 16)[n_rings * 400 + 2, 10] = 4
    cases["ring_id"] = c
    c = good.copy()   # two points of one ring exchanged: the ring is no longer a rotated monotone sequence
    p = _points(c)
    i, j = n_rings * 300 + 1, n_rings * 310 + 1
    p[[i, j], 0:3] = p[[j, i], 0:3]
    cases["swap"] = c
    c = good.copy()   # exact duplicate of the previous point of the ring: equal angles
    p = _points(c)
    p[n_rings * 500 + 3, 0:2] = p[n_rings * 499 + 3, 0:2]
    cases["duplicate"] = c
    c = good.copy()   # two adjacent zero-XY points: CalcRadian throws, the ring is skipped (math.cpp:40-42)
    p = _points(c)
    p[n_rings * 200 + 5, 0:2] = 0.0
    p[n_rings * 201 + 5, 0:2] = 0.0
    cases["zero_pair"] = c
    cases["truncated"] = good[:-3].copy()   # n_points no longer a multiple of the period
    c = good.copy()   # a fully shuffled ring order inside the scan
    rng = np.random.default_rng(0)
    cases["shuffled"] = c[rng.permutation(len(c))]
    for name, cloud in cases.items():
        out, stats = _check(oracle, hp, [good, cloud, good])
        assert stats["general_scans"] == 1, (name, stats)
        assert sum(stats["fast_rings"]) in (2 * n_rings, 3 * n_rings), (name, stats)
    # zero pair: the ring is skipped, its neighbours are not
    from lidar_feature_extraction_b200 import _native as N

    out, _ = _check(oracle, hp, [cases["zero_pair"]])
    assert out.rings["status"][0][5] == N.LFX_RING_SKIPPED and out.rings["status"][0][4] == N.LFX_RING_OK


def test_isolated_zero_xy_points(oracle):
    """An isolated zero-XY point sorts to polar angle 0 (ring.hpp:69-80), i.e. out of its firing slot: the
    ring is no longer a rotation, the order check catches it and the general path sorts it."""
    clouds = [regular_scan(s, 4, 800, zero_xy=1) for s in range(3)]
    out, stats = _check(oracle, _hp(), clouds)
    assert stats["general_scans"] >= 1, stats


def test_mixed_batches(oracle):
    from lidar_feature_extraction_b200 import synth

    clouds = [regular_scan(1, 8, 1200), adv.ragged_scan(2, [300, 17, 2048, 64], shuffle="random"),
              synth.scan_host(synth.spec("vlp16"), 3), np.zeros((0, 32), np.uint8),
              adv.ragged_scan(4, [500, 500, 500], shuffle="interleave"),   # regular, ascending, unrotated
              regular_scan(5, 3, 50),                                        # too short for the fast path
              synth.scan_host(synth.spec("hdl64"), 1)]                       # drop-outs: not regular
    out, stats = _check(oracle, _hp(), clouds)
    assert stats["general_scans"] == 4, stats
    assert sum(stats["fast_rings"]) == 8 + 16 + 3, stats
    out2, stats2 = _check(oracle, _hp(), clouds[::-1], diag=False)
    assert stats2["general_scans"] == 4
    for s in range(len(clouds)):
        assert np.array_equal(out.scan_edges(s), out2.scan_edges(len(clouds) - 1 - s))
        assert np.array_equal(out.scan_surfaces(s), out2.scan_surfaces(len(clouds) - 1 - s))


def test_other_paddings_use_the_general_kernel(oracle):
    """The sector kernel is compiled for convolution_padding 1..8 (tests/test_gpu_envelope.py); 9..15 run on the on-chip
    per-ring kernel with run-time loops."""
    clouds = [regular_scan(3, 4, 900)]
    for padding in (9, 12):
        out, stats = _check(oracle, _hp(padding=padding), clouds)
        assert sum(stats["fast_rings"]) == 0 and stats["general_scans"] == 1


def test_full_size_batch_properties(oracle):
    """BASELINE-sized device-resident batch (os128 x 64): counts/offsets/labels consistent, a sample of
    scans matches the oracle exactly, and the production (non-diagnostic) kernel gives identical clouds."""
    import ctypes as C

    import torch

    from lidar_feature_extraction_b200 import _native as N
    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    sp = synth.spec("os128")
    n_scans, per = 64, sp.n_rings * sp.n_cols
    hp = _hp()
    with _fe(hp) as fe:
        d_in = torch.empty((n_scans * per, 32), dtype=torch.uint8, device="cuda")
        assert N.lib().lfx_synth_batch_device(fe.handle, C.byref(sp), 500, n_scans, d_in.data_ptr()) == 0
        torch.cuda.synchronize()
        views = [fe.wire_view((d_in.data_ptr() + s * per * 32, per)) for s in range(n_scans)]
        fe.extract_views(views, keep=d_in)
        out = fe.fetch()
        stats = fe.batch_stats()
        host = d_in.cpu().numpy().reshape(n_scans, per, 32)
    assert stats["general_scans"] == 0 and stats["fast_rings"][1] == n_scans * sp.n_rings, stats
    assert np.array_equal(out.offsets[1:], np.cumsum(out.counts, axis=0))
    hist = np.bincount(out.labels, minlength=256)
    assert hist[1] == out.counts[:, 0].sum() and hist[3] == out.counts[:, 1].sum() and hist[8:].sum() == 0
    for s in (0, 31, 63):
        compare_scan(out, s, host[s], oracle.extract_scan(host[s], oracle_params(ob, hp)))
    with _fe(hp, diag=False) as fe:
        fe.extract_views(views, keep=d_in)
        out2 = fe.fetch()
    assert np.array_equal(out.labels, out2.labels) and np.array_equal(out.edge_xyz, out2.edge_xyz)
    assert np.array_equal(out.surface_xyz, out2.surface_xyz) and np.array_equal(out.counts, out2.counts)


# ---------------------------------------------------------------------------------------------------------
# Indexed sector path: scans that are not regular are bucketed by ring id; every bucket that is a rotated
# monotone sequence of polar angles runs on the sector kernel through its index list (k_probe_rings +
# k_extract_sectors<indexed>), everything else on the per-ring kernel.

def dropout_scan(seed, n_rings, width, drop=0.03, **kw):
    """Firing-order scan with returns removed at random - what the converter's zero-point filter
    (point_type_converter/convert.py:201) leaves of a real sweep: ragged rings, no fixed ring period."""
    cloud = regular_scan(seed, n_rings, width, **kw)
    keep = np.random.default_rng(seed + 1000).random(len(cloud)) >= drop
    keep[:n_rings] = True
    return np.ascontiguousarray(cloud[keep])


@pytest.mark.parametrize("direction", ["cw", "ccw"])
@pytest.mark.parametrize("width", [100, 299, 640, 1024, 1800, 2048, 2232])
def test_dropout_scans_take_the_indexed_sector_path(oracle, width, direction):
    clouds = [dropout_scan(width + s, len(adv.KINDS), width, kinds=adv.KINDS, direction=direction) for s in range(3)]
    for hp in (_hp(), _hp(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0),
               _hp(surface_threshold=1e9, edge_threshold=1e-9)):
        out, stats = _check(oracle, hp, clouds)
        assert stats["general_scans"] == 3 and sum(stats["fast_rings"]) == 0, (width, stats)
        assert sum(stats["indexed_rings"]) == 3 * len(adv.KINDS) and stats["general_rings"] == 0, (width, stats)
        assert (out.rings["order_path"] == 0).all()
        _, stats2 = _check(oracle, hp, clouds, diag=False)   # the production instantiation
        assert stats2 == stats


def test_synthetic_tunnel_scans_take_the_indexed_sector_path(oracle):
    from lidar_feature_extraction_b200 import synth

    sp = synth.spec("hdl64")   # 1 % drop-outs: ragged rings
    clouds = [synth.scan_host(sp, f) for f in range(4)]
    for hp in (_hp(), _hp(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0)):
        out, stats = _check(oracle, hp, clouds)
        assert stats["general_scans"] == 4 and sum(stats["indexed_rings"]) == 4 * sp.n_rings and stats["general_rings"] == 0, stats


def test_indexed_rings_that_fail_a_check_are_redone_per_ring(oracle):
    """Corruptions of single rings of a drop-out scan: only the damaged ring goes to the per-ring kernel."""
    hp = _hp()
    n_rings, width = 6, 900
    good = dropout_scan(1, n_rings, width)
    ring = good.view(np.uint16).reshape(-1, 16)[:, 10]
    cases = {}
    for name, target in (("swap", 1), ("duplicate", 3), ("zero_pair", 5)):
        c = good.copy()
        p = _points(c)
        members = np.nonzero(ring == target)[0]
        i, j = members[300], members[310]
        if name == "swap":
            p[[i, j], 0:3] = p[[j, i], 0:3]
        elif name == "duplicate":
            p[members[500], 0:2] = p[members[499], 0:2]
        else:
            p[members[200], 0:2] = 0.0
            p[members[201], 0:2] = 0.0
        cases[name] = c
    for name, cloud in cases.items():
        out, stats = _check(oracle, hp, [good, cloud, good])
        assert stats["general_scans"] == 3, (name, stats)
        assert stats["general_rings"] == 1 and sum(stats["indexed_rings"]) == 3 * n_rings, (name, stats)
    out, _ = _check(oracle, hp, [cases["zero_pair"]])
    from lidar_feature_extraction_b200 import _native as N

    assert out.rings["status"][0][5] == N.LFX_RING_SKIPPED and out.rings["status"][0][4] == N.LFX_RING_OK


def test_indexed_path_mixed_ring_lengths_and_sparse_rings(oracle):
    """Buckets of very different lengths in one scan (each its own sector table and window class), short and
    sparse rings next to them, ring-major and shuffled layouts."""
    lengths = [2232, 64, 300, 17, 1800, 3, 2048, 640, 63, 1024]
    for shuffle, want_indexed in (("interleave", True), ("rotate", True), ("rotate_reverse", True), ("random", False)):
        clouds = [adv.ragged_scan(7 + s, lengths, shuffle=shuffle) for s in range(2)]
        # seven rings have at least 64 points; with 4 sectors the three longest exceed the 384-position window
        for hp, n_idx in ((_hp(), 7), (_hp(padding=2, n_blocks=4), 4)):
            out, stats = _check(oracle, hp, clouds)
            assert stats["general_scans"] == 2, (shuffle, stats)
            if want_indexed:
                assert sum(stats["indexed_rings"]) == 2 * n_idx and stats["general_rings"] == 2 * (10 - n_idx), (shuffle, stats)
            else:
                assert stats["general_rings"] >= 2 * 3, (shuffle, stats)


@pytest.mark.parametrize("max_rings", [8, 128, 1024])
def test_bucketing_kernels_for_small_and_large_ring_id_spaces(oracle, max_rings):
    """max_rings <= ~500 buckets with the bitmap scatter, larger id spaces with the match-based scatter."""
    ids = [5, 0, 7, 3] if max_rings == 8 else ([100, 3, 64, 127] if max_rings == 128 else [1000, 3, 512, 700])
    clouds = [dropout_scan(40 + s, 4, 777, ring_ids=ids) for s in range(3)] + [adv.ragged_scan(9, [300, 70, 1500, 20], shuffle="random", ring_ids=ids)]
    out, stats = _check(oracle, _hp(), clouds, max_rings=max_rings)
    assert stats["general_scans"] == 4 and sum(stats["indexed_rings"]) >= 12, stats
