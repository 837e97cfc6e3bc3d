"""The device converter (k_convert through lfx_convert_batch) against the golden fixtures made by
the reference's own convert.py (tests/golden/convert_*.npz) and against the numpy oracle on random layouts:
output bytes identical, and an error exactly where the reference raises."""
import glob
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "convert_*.npz")))
NP_TYPES = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 8: np.float64}


def _msg(raw, fields, step, big):
    from lidar_feature_extraction_b200 import PointCloud2, PointField

    return PointCloud2(data=np.ascontiguousarray(raw).reshape(-1).view(np.uint8), point_step=step, is_bigendian=big,
                       fields=[PointField(f.name, f.offset, f.datatype) for f in fields])


def _oracle(raw, fields, step, big):
    from oracle import convert_oracle as co

    try:
        out, _ = co.convert(raw, fields, step, big)
        return out
    except co.ConvertError:
        return None


@pytest.fixture(scope="module")
def conv():
    from lidar_feature_extraction_b200 import PointTypeConverter

    with PointTypeConverter() as c:
        yield c


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_golden_fixtures_of_the_reference_converter(conv, path):
    from lidar_feature_extraction_b200 import ConvertError
    from oracle import convert_oracle as co

    z = np.load(path)
    fields = [co.Field(str(n), int(o), int(d)) for n, o, d in zip(z["field_names"], z["field_offsets"], z["field_datatypes"])]
    msg = _msg(z["raw"], fields, int(z["point_step"]), bool(z["is_bigendian"]))
    if str(z["error"]):
        with pytest.raises(ConvertError):
            conv.callback(msg)
        return
    out = conv.callback(msg)
    assert out.width == z["out"].shape[0] and out.point_step == 32 and out.is_dense and not out.is_bigendian
    assert [(f.name, f.offset, f.datatype) for f in out.fields] == [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("padding", 12, 7), ("intensity", 16, 7), ("ring", 20, 4)]
    assert np.array_equal(out.data, z["out"])


def _random_case(rng, n_points, tame=True):
    """A random field table (possibly overlapping / unordered / unaligned offsets) and matching random data."""
    from oracle import convert_oracle as co

    names = ["x", "y", "z", "intensity", "ring"]
    extra = ["t", "reflectivity", "noise", "range", "padding", "ambient"]
    if rng.random() < 0.15:
        names.remove(rng.choice(names))                      # a missing retained field -> struct.pack count mismatch
    names += list(rng.choice(extra, size=rng.integers(0, 4), replace=False))
    order = list(names)
    if rng.random() < 0.3:
        rng.shuffle(order)
    fields, off = [], 0
    for nm in order:
        if nm == "ring":
            dt = int(rng.choice([2, 4, 4, 5, 6, 3, 7] if not tame else [2, 4, 4, 6, 5]))
        elif nm in ("x", "y", "z", "intensity"):
            dt = int(rng.choice([7, 7, 7, 8, 5, 3]))
        else:
            dt = int(rng.integers(1, 9))
        size = np.dtype(NP_TYPES[dt]).itemsize
        if rng.random() < 0.7:
            off = (off + size - 1) // size * size            # natural alignment, the common case
        if nm == "padding" or (off <= 12 < off + size) or (12 <= off < 16 and rng.random() < 0.8):
            off = max(off, 16) if rng.random() < 0.9 else off   # usually leave the hole the appended padding field reads
        fields.append(co.Field(nm, off, dt))
        off += size + int(rng.choice([0, 0, 0, 1, 2, 4]))
        if rng.random() < 0.05:
            off = max(off - int(rng.integers(1, 6)), 0)      # overlapping declaration: effective offsets kick in
    end = max(f.offset + np.dtype(NP_TYPES[f.datatype]).itemsize for f in fields)
    step = max(end, 16) + int(rng.choice([0, 0, 2, 4, 12, 16]))
    if rng.random() < 0.05:
        step = max(end - 2, 1)                               # fields run past point_step
    big = bool(rng.random() < 0.25)
    raw = np.zeros((n_points, step), np.uint8)
    bo = ">" if big else "<"
    for f in fields:
        dt = np.dtype(NP_TYPES[f.datatype]).newbyteorder(bo)
        if f.offset + dt.itemsize > step:
            continue
        if dt.kind == "f":
            v = rng.normal(0, 30, n_points)
            v[rng.random(n_points) < 0.1] = 0.0
            v[rng.random(n_points) < 0.01] = -0.0
            col = v.astype(dt)
            if not tame or rng.random() < 0.3:
                bits = rng.integers(0, 256, (n_points, dt.itemsize), dtype=np.uint8)
                sel = rng.random(n_points) < (0.02 if dt.itemsize == 4 or not tame else 0.0)
                col = col.copy()
                col.view(np.uint8).reshape(n_points, dt.itemsize)[sel] = bits[sel]
        elif f.name == "ring":
            hi = 64 if tame else (70000 if dt.itemsize >= 4 else 64)
            lo = 0 if tame or dt.kind == "u" else -2
            col = rng.integers(lo, hi, n_points).astype(dt)
        else:
            info = np.iinfo(NP_TYPES[f.datatype])
            col = rng.integers(info.min, int(info.max) + 1, n_points).astype(dt)
            col[rng.random(n_points) < 0.1] = 0
        raw[:, f.offset:f.offset + dt.itemsize] = col.view(np.uint8).reshape(n_points, dt.itemsize)
    dead = rng.random(n_points) < 0.2                          # returns the driver zeroed out
    raw[dead] = 0
    return raw, fields, step, big


@pytest.mark.parametrize("seed", range(12))
def test_random_layouts_match_the_oracle(conv, seed):
    from lidar_feature_extraction_b200 import ConvertError

    rng = np.random.default_rng(1000 + seed)
    n_ok = n_err = 0
    for it in range(25):
        n = int(rng.choice([0, 1, 7, 255, 256, 257, 1000, 5000]))
        raw, fields, step, big = _random_case(rng, n, tame=(it % 3 != 0))
        want = _oracle(raw, fields, step, big)
        msg = _msg(raw, fields, step, big)
        if want is None:
            with pytest.raises(ConvertError):
                conv.callback(msg)
            n_err += 1
        else:
            got = conv.callback(msg)
            assert got.data.shape == want.shape, (seed, it, fields, step, big)
            assert np.array_equal(got.data, want), (seed, it, fields, step, big)
            n_ok += 1
    assert n_ok >= 5


def test_batch_of_clouds_with_many_tiles_and_failures(conv):
    """One launch, several clouds: long ones (look-back over hundreds of tiles), an empty one, failing ones."""
    from oracle import convert_oracle as co

    rng = np.random.default_rng(7)
    cases = []
    for n in (100_003, 0, 65_536, 1, 300_000):
        cases.append(_random_case(rng, n, tame=True))
    bad = [co.Field("x", 0, 7), co.Field("y", 4, 7), co.Field("z", 8, 7), co.Field("intensity", 16, 7), co.Field("ring", 20, 7)]
    cases.insert(2, (rng.normal(0, 1, (50, 6)).astype("<f4").view(np.uint8).reshape(50, 24), bad, 24, False))   # float ring
    msgs = [_msg(*c) for c in cases]
    res = conv.convert_batch(msgs)
    assert res.n_clouds == len(cases)
    for i, c in enumerate(cases):
        want = _oracle(*c)
        if want is None:
            assert conv.status_of(i) != 0 and conv.kept(i) == 0
        else:
            assert conv.status_of(i) == 0, (i, conv.status_of(i))
            assert np.array_equal(conv.fetch(i), want), i


def test_device_resident_input_and_chaining_into_the_extraction():
    """Raw Ouster-like cloud on the device -> converter -> extraction without leaving the device equals the
    oracle's extraction of the oracle-converted cloud."""
    import torch

    from helpers import compare_scan, oracle_params
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, PointCloud2, PointField, PointTypeConverter, synth
    from oracle import binding as ob
    from oracle import convert_oracle as co

    wire = synth.scan_host(synth.spec("vlp16"), frame=3)
    x, y, z, inten, ring = (np.ascontiguousarray(a) for a in synth.fields(wire))
    n = len(x)
    rng = np.random.default_rng(5)
    raw = np.zeros((n, 48), np.uint8)   # test_convert.py:177-187: x,y,z,intensity,t,reflectivity,ring(u8),noise,range
    for off, a in ((0, x), (4, y), (8, z), (16, inten), (20, rng.integers(0, 2**32, n).astype("<u4")),
                   (24, rng.integers(0, 65536, n).astype("<u2")), (26, ring.astype(np.uint8)),
                   (28, rng.integers(0, 65536, n).astype("<u2")), (32, rng.integers(0, 2**32, n).astype("<u4"))):
        raw[:, off:off + a.dtype.itemsize] = a.view(np.uint8).reshape(n, -1)
    dead = rng.random(n) < 0.03
    raw[dead, 0:12] = 0
    fields = [co.Field("x", 0, 7), co.Field("y", 4, 7), co.Field("z", 8, 7), co.Field("intensity", 16, 7), co.Field("t", 20, 6),
              co.Field("reflectivity", 24, 4), co.Field("ring", 26, 2), co.Field("noise", 28, 4), co.Field("range", 32, 6)]
    want_cloud, keep = co.convert(raw, fields, 48, False)
    assert keep.sum() == n - dead.sum()
    hp = HyperParameters()
    oracle = ob.Oracle()
    want = oracle.extract_scan(want_cloud, oracle_params(ob, hp))
    d_raw = torch.from_numpy(raw).cuda()
    with FeatureExtraction(hp, want_sorted_src=True, want_curvature=True) as fe, PointTypeConverter(fe) as conv:
        msg = PointCloud2(data=d_raw, point_step=48, fields=[PointField(f.name, f.offset, f.datatype) for f in fields])
        conv.convert_batch([msg, msg])
        assert np.array_equal(conv.fetch(0), want_cloud) and np.array_equal(conv.fetch(1), want_cloud)
        fe.extract_views([conv.view(0), conv.view(1)])
        out = fe.fetch()
        compare_scan(out, 0, want_cloud, want)
        compare_scan(out, 1, want_cloud, want)
        # the bulk form of the same chain (one marshalled batch, one call for all views): the bucketing then reads the
        # converter's ring-id by-product instead of the points, and a failed cloud in the middle is left out
        bad = PointCloud2(data=d_raw, point_step=48, fields=[PointField(f.name, f.offset, f.datatype) for f in fields[:5]])  # no ring
        batch = conv.marshal([msg, bad, msg])
        for _ in range(2):
            conv.convert_batch(batch)
            views = conv.views()
            assert views.n_views == 2 and conv.status_of(1) != 0
            fe.extract_views(views)
            out = fe.fetch()
            assert fe.batch_stats()["general_scans"] == 2
            compare_scan(out, 0, want_cloud, want)
            compare_scan(out, 1, want_cloud, want)


def test_common_plan_fast_kernel_special_values(conv):
    """The all-float32 little-endian plan runs on the specialised instantiation: signalling NaNs come out quiet
    with their payload, -0.0 counts as zero, a uint32 ring above 65535 fails like struct.pack('H')."""
    from lidar_feature_extraction_b200 import ConvertError
    from oracle import convert_oracle as co

    rng = np.random.default_rng(11)
    n = 3000
    fields = [co.Field("x", 0, 7), co.Field("y", 4, 7), co.Field("z", 8, 7), co.Field("intensity", 16, 7), co.Field("ring", 20, 6)]
    raw = np.zeros((n, 24), np.uint8)
    f = rng.normal(0, 10, (n, 4)).astype("<f4")
    bits = f.view(np.uint32)
    bits[rng.random((n, 4)) < 0.05] = 0x7F800001 + rng.integers(0, 0x3FFFFF)       # signalling NaNs with payloads
    bits[rng.random((n, 4)) < 0.05] = 0xFFC12345
    bits[rng.random((n, 4)) < 0.02] = 0x7F800000                                     # inf
    bits[rng.random((n, 4)) < 0.02] = 0x00000001                                     # denormal
    zero = rng.random(n) < 0.2
    bits[zero, :3] = rng.choice(np.array([0, 0x80000000], np.uint32), (int(zero.sum()), 3))   # +-0.0
    raw[:, 0:12] = f[:, :3].view(np.uint8).reshape(n, 12)
    raw[:, 16:20] = f[:, 3:].view(np.uint8).reshape(n, 4)
    raw[:, 20:24] = rng.integers(0, 65536, n).astype("<u4").view(np.uint8).reshape(n, 4)
    want = _oracle(raw, fields, 24, False)
    assert want is not None and 0 < want.shape[0] < n
    assert np.array_equal(conv.callback(_msg(raw, fields, 24, False)).data, want)
    raw[n // 2, 20:24] = np.array([65536], "<u4").view(np.uint8)
    raw[n // 2, 0:4] = np.array([1.0], "<f4").view(np.uint8)
    assert _oracle(raw, fields, 24, False) is None
    with pytest.raises(ConvertError):
        conv.callback(_msg(raw, fields, 24, False))
