"""GPU parity on ragged / adversarial inputs, through the C ABI, against the oracle (bit-exact labels,
order and feature sets; curvature within 1e-6 relative). Edge cases follow the reference's own
failure modes: sparse rings (ring.cpp:46-59), rings that throw (feature_extraction.cpp:154-156),
zero-XY pairs (math.cpp:40-42), missing ring field / non-dense clouds (feature_extraction.cpp:96-108)."""
import ctypes as C

import numpy as np
import pytest

import adversarial as adv
from helpers import compare_scan, oracle_params

pytestmark = pytest.mark.gpu


def _fe(hp=None, **kw):
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters

    kw.setdefault("want_sorted_src", True)
    kw.setdefault("want_curvature", True)
    return FeatureExtraction(hp or HyperParameters(), device=0, **kw)


def _hp(**kw):
    from lidar_feature_extraction_b200 import HyperParameters

    return HyperParameters(**kw)


PARAMSETS = {
    "default": {},
    "yaml": dict(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0),
    "p2b4": dict(padding=2, n_blocks=4, edge_threshold=0.02, surface_threshold=0.2),
    "p8b3": dict(padding=8, n_blocks=3, neighbor_degree_threshold=1.0),
    "p15b1": dict(padding=15, n_blocks=1, edge_threshold=0.5, surface_threshold=0.5),
    "p1b64": dict(padding=1, n_blocks=64, neighbor_degree_threshold=5.0),
}


def _check(oracle, hp, clouds, **kw):
    from oracle import binding as ob

    with _fe(hp, **kw) as fe:
        out = fe.extract_batch(clouds)
    for s, cloud in enumerate(clouds):
        compare_scan(out, s, cloud, oracle.extract_scan(cloud, oracle_params(ob, hp)))
    return out


@pytest.mark.parametrize("pname", sorted(PARAMSETS))
@pytest.mark.parametrize("shuffle", ["none", "interleave", "random", "rotate", "rotate_reverse", "reverse"])
def test_ragged_scans(oracle, pname, shuffle):
    hp = _hp(**PARAMSETS[pname])
    clouds = []
    for seed in range(5):
        rng = np.random.default_rng(100 + seed)
        lengths = [int(v) for v in rng.choice([0, 1, 3, 6, 11, 12, 17, 23, 40, 97, 300, 777, 2048, 2304], size=9)]
        clouds.append(adv.ragged_scan(seed, lengths, shuffle=shuffle, zero_xy=seed % 3))
    _check(oracle, hp, clouds)


@pytest.mark.parametrize("path", [1, 2])
def test_forced_sort_paths(oracle, path):
    """force_order_path=1: key sort + exact verification; 2: exact comparator sort. Same results."""
    clouds = [adv.ragged_scan(s, [500, 33, 1024, 2000, 64], shuffle=sh) for s, sh in enumerate(["random", "none", "rotate_reverse", "interleave"])]
    out = _check(oracle, _hp(), clouds, force_order_path=path)
    used = out.rings["order_path"][out.rings["count"] > 5]
    # the key sort may hand rings with near-coincident azimuths to the exact sort (path 2); never the reverse
    assert (used >= path).all() and (used == path).mean() > 0.5


def test_auto_order_path_selection(oracle):
    sorted_scan = adv.ragged_scan(1, [400, 400], shuffle="rotate")
    shuffled = adv.ragged_scan(2, [400, 400], shuffle="random")
    out = _check(oracle, _hp(), [sorted_scan, shuffled])
    assert (out.rings["order_path"][0][:2] == 0).all()      # rotated-monotone fast path
    assert (out.rings["order_path"][1][:2] >= 1).all()      # needed a sort


def test_exact_curvature_ties_follow_index_tiebreak(oracle):
    hp = _hp(edge_threshold=0.01, surface_threshold=0.5)
    _check(oracle, hp, [adv.symmetric_ties_scan(s) for s in range(4)])
    _check(oracle, _hp(padding=2, n_blocks=4, surface_threshold=5.0), [adv.symmetric_ties_scan(7, n_quarter=60, n_rings=5)])


@pytest.mark.parametrize("kind", adv.KINDS)
def test_each_ring_kind_at_capacity(oracle, kind):
    """one long ring per kind (ramp = worst-case selection depth; plateau = all-equal curvature)."""
    clouds = [adv.ragged_scan(3, [2304, 1800], kinds=[kind], shuffle="rotate_reverse")]
    _check(oracle, _hp(), clouds)
    _check(oracle, _hp(surface_threshold=1e9, edge_threshold=1e-9), clouds)  # everything is a candidate


def test_sparse_and_skipped_rings_contribute_nothing(oracle):
    from lidar_feature_extraction_b200 import _native as N

    # P=5: n<6 sparse; 6..10 too short for the convolution; 11..16 too short for 6 sectors; 17..21 sector length 1
    lengths = [5, 6, 10, 11, 16, 17, 21, 22, 23, 300]
    out = _check(oracle, _hp(), [adv.ragged_scan(11, lengths, kinds=["mixed"])])
    st = out.rings["status"][0][: len(lengths)].tolist()
    assert st[0] == N.LFX_RING_SPARSE and st[-1] == N.LFX_RING_OK
    assert all(v == N.LFX_RING_SKIPPED for v in st[1:7]), st
    # two adjacent zero-XY points => CalcRadian throws => ring skipped
    c = adv.ragged_scan(12, [200, 200], kinds=["mixed"])
    x = c.view(np.float32).reshape(-1, 8)
    x[10:12, 0:2] = 0.0
    out = _check(oracle, _hp(), [c])
    assert out.rings["status"][0][0] == N.LFX_RING_SKIPPED and out.rings["status"][0][1] == N.LFX_RING_OK


def test_empty_inputs(oracle):
    with _fe() as fe:
        out = fe.extract_batch([])
        assert out.counts.shape == (0, 2) and out.offsets.tolist() == [[0, 0]]
        empty = np.zeros((0, 32), np.uint8)
        one = adv.ragged_scan(0, [100])
        out = fe.extract_batch([empty, one, empty])
        assert out.counts[0].tolist() == [0, 0] and out.counts[2].tolist() == [0, 0]
        assert out.counts[1].sum() > 0


def test_sparse_ring_ids_and_general_layouts(oracle):
    """ring ids need not be dense; fields may sit anywhere in the point (lookup by name, ros_msg.hpp:73-79)."""
    from lidar_feature_extraction_b200 import PointCloud2, PointField
    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    base = adv.ragged_scan(5, [150, 90, 400], ring_ids=[3, 77, 100], shuffle="interleave")
    x, y, z, _, ring = (np.ascontiguousarray(a) for a in synth.fields(base))
    hp = _hp()
    want = oracle.extract_scan(base, oracle_params(ob, hp))
    for step, ox, oy, oz, oring, rdt, npdt in ((48, 16, 4, 32, 44, 6, np.uint32), (20, 0, 4, 8, 13, 2, np.uint8), (32, 12, 8, 4, 0, 4, np.uint16)):
        buf = np.zeros((len(x), step), np.uint8)
        buf[:, ox:ox + 4] = x.view(np.uint8).reshape(-1, 4)
        buf[:, oy:oy + 4] = y.view(np.uint8).reshape(-1, 4)
        buf[:, oz:oz + 4] = z.view(np.uint8).reshape(-1, 4)
        rb = ring.astype(npdt).view(np.uint8).reshape(len(x), -1)
        buf[:, oring:oring + rb.shape[1]] = rb
        msg = PointCloud2(data=buf, point_step=step, width=len(x),
                          fields=[PointField("x", ox, 7), PointField("y", oy, 7), PointField("z", oz, 7), PointField("ring", oring, rdt)])
        with _fe(hp) as fe:
            out = fe.extract_batch([msg])
        compare_scan(out, 0, base, want)


def test_device_resident_input_and_callback_mirror(oracle):
    import torch

    from lidar_feature_extraction_b200 import PointCloud2, synth
    from oracle import binding as ob

    cloud = synth.scan_host(synth.spec("vlp16"), 5)
    hp = _hp()
    want = oracle.extract_scan(cloud, oracle_params(ob, hp))
    with _fe(hp) as fe:
        out = fe.extract_batch([torch.from_numpy(cloud).cuda()])
        compare_scan(out, 0, cloud, want)
        res = fe.callback(PointCloud2.from_wire(cloud, stamp=(12, 34)))
    assert res["stamp"] == (12, 34) and res["frame_id"] == "lidar_feature_base_link"  # feature_extraction.cpp:159-166
    assert res["scan_edge"].point_step == 16 and res["scan_edge"].width == len(want.edge_idx)
    assert res["scan_surface"].width == len(want.surface_idx)
    got = res["scan_edge"].data.view(np.float32).reshape(-1, 4)
    x, y, z, _, _ = synth.fields(cloud)
    src = want.sorted_src[want.edge_idx]
    assert np.array_equal(got[:, :3], np.stack([x[src], y[src], z[src]], axis=1))


def test_error_codes():
    from lidar_feature_extraction_b200 import ExtractionError, PointCloud2, PointField
    from lidar_feature_extraction_b200 import _native as N

    cloud = adv.ragged_scan(0, [100, 100])
    with _fe() as fe:
        msg = PointCloud2.from_wire(cloud)
        msg.is_dense = False
        with pytest.raises(ExtractionError) as e:
            fe.callback(msg)
        assert e.value.code == N.LFX_E_NOT_DENSE     # feature_extraction.cpp:96-101
        with pytest.raises(ExtractionError) as e:
            fe.extract_batch([msg])
        assert e.value.code == N.LFX_E_NOT_DENSE
        msg = PointCloud2.from_wire(cloud)
        msg.fields = [f for f in msg.fields if f.name != "ring"]
        with pytest.raises(ExtractionError) as e:
            fe.callback(msg)
        assert e.value.code == N.LFX_E_NO_RING       # feature_extraction.cpp:103-108
        with pytest.raises(ExtractionError) as e:
            fe.extract_batch([msg])
        assert e.value.code == N.LFX_E_NO_RING
        msg = PointCloud2.from_wire(cloud)
        msg.fields[0] = PointField("x", 30, 7)       # x would run past point_step
        with pytest.raises(ExtractionError) as e:
            fe.extract_batch([msg])
        assert e.value.code == N.LFX_E_BAD_LAYOUT
        fe.extract_batch([cloud])                    # the handle stays usable after errors
    # (a ring longer than max_ring_points is not an error: it runs on k_extract_rings_big, tests/test_gpu_envelope.py)
    with _fe(max_rings=8) as fe:
        with pytest.raises(ExtractionError) as e:
            fe.extract_batch([adv.ragged_scan(0, [50, 50], ring_ids=[1, 9])])
        assert e.value.code == N.LFX_E_CAPACITY


def test_batches_are_deterministic_and_scan_independent(oracle):
    from lidar_feature_extraction_b200 import synth

    sp = synth.spec("hdl32")
    clouds = [synth.scan_host(sp, f) for f in range(6)]
    with _fe() as fe:
        a = fe.extract_batch(clouds)
        b = fe.extract_batch(clouds)                     # same graph replayed
        c = fe.extract_batch(clouds[::-1])               # different batch composition
        single = [fe.extract_batch([cl]) for cl in clouds[:2]]
    assert np.array_equal(a.labels, b.labels) and np.array_equal(a.edge_xyz, b.edge_xyz) and np.array_equal(a.surface_xyz, b.surface_xyz)
    for s in range(6):
        assert np.array_equal(a.scan_edges(s), c.scan_edges(5 - s)) and np.array_equal(a.scan_surfaces(s), c.scan_surfaces(5 - s))
    for s in range(2):
        assert np.array_equal(a.scan_edges(s), single[s].scan_edges(0))


def test_full_size_properties(oracle):
    """BASELINE-sized batch (hdl32 x 256 on device): size-independent checks - every selected point is a
    source point of the right label, counts sum to offsets, labels histogram is consistent, and a sample
    of scans matches the oracle exactly."""
    import torch

    from lidar_feature_extraction_b200 import _native as N
    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    sp = synth.spec("hdl32")
    n_scans, per = 256, sp.n_rings * sp.n_cols
    with _fe() as fe:
        d_in = torch.empty((n_scans * per, 32), dtype=torch.uint8, device="cuda")
        assert N.lib().lfx_synth_batch_device(fe.handle, C.byref(sp), 1000, n_scans, d_in.data_ptr()) == 0
        torch.cuda.synchronize()
        views = [fe.wire_view((d_in.data_ptr() + s * per * 32, per)) for s in range(n_scans)]
        fe.extract_views(views, keep=d_in)
        out = fe.fetch()
        host = d_in.cpu().numpy().reshape(n_scans, per, 32)
    assert np.array_equal(out.offsets[1:], np.cumsum(out.counts, axis=0))
    hist = np.bincount(out.labels, minlength=256)
    assert hist[1] == out.counts[:, 0].sum() and hist[3] == out.counts[:, 1].sum() and hist[8:].sum() == 0
    assert (out.edge_xyz[:, 3] == 1.0).all() and (out.surface_xyz[:, 3] == 1.0).all()
    hp = _hp()
    for s in (0, 101, 255):
        compare_scan(out, s, host[s], oracle.extract_scan(host[s], oracle_params(ob, hp)))
    # sorted_src is a permutation of every scan's point indices
    for s in (3, 200):
        seg = out.sorted_src[int(out.point_base[s]): int(out.point_base[s + 1])]
        assert np.array_equal(np.sort(seg), np.arange(per))
