"""The converter oracle (oracle/convert_oracle.py) against (1) the known-answer vectors of the reference's own
point_type_converter/test/test_convert.py and (2) tests/golden/convert_*.npz, outputs of the reference's
unmodified convert.py (tests/golden/make_convert_golden.py). CPU only."""
import glob
import os
import struct

import numpy as np
import pytest

from oracle import convert_oracle as co

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "convert_*.npz")))


def load_case(path):
    z = np.load(path)
    fields = [co.Field(str(n), int(o), int(d)) for n, o, d in zip(z["field_names"], z["field_offsets"], z["field_datatypes"])]
    return z["raw"], fields, int(z["point_step"]), bool(z["is_bigendian"]), z["out"], str(z["error"])


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 19


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[8:-4] for p in GOLDEN])
def test_oracle_matches_reference_converter(path):
    raw, fields, step, big, want, error = load_case(path)
    if error:
        with pytest.raises(co.ConvertError):
            co.convert(raw, fields, step, big)
        return
    got, keep = co.convert(raw, fields, step, big)
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    assert int(keep.sum()) == want.shape[0]


# ---- known-answer vectors of the reference's own tests

OUSTER = [co.Field("x", 0, 7), co.Field("y", 4, 7), co.Field("z", 8, 7), co.Field("intensity", 16, 7), co.Field("t", 20, 6),
          co.Field("reflectivity", 24, 4), co.Field("ring", 26, 2), co.Field("noise", 28, 4), co.Field("range", 32, 6)]


def test_effective_offsets_follow_create_point_format():
    """test_convert.py:41-58: 'fffxxxxfIHBxHxxIxxxxxxxxxxxx' for the 48-byte Ouster layout (without padding)."""
    plan = co.make_plan(OUSTER, 48)
    # the plan contains the appended padding field at 12; drop it to compare with the reference's string
    offs = [o for n, o in zip(plan.names, plan.eff_offsets) if n != "padding"]
    assert offs == [0, 4, 8, 16, 20, 24, 26, 28, 32]
    assert plan.point_size == 48 and plan.layout_error is None
    assert [plan.names[i] for i in plan.retained] == ["x", "y", "z", "padding", "intensity", "ring"]


def test_unpack_vector_of_reference_test():
    """test_convert.py:89-120: two 32-byte points -> (2,4,6,0,10,22), (1,3,5,0,20,11)."""
    data = (b"\x00\x00\x00@\x00\x00\x80@\x00\x00\xc0@\x00\x00\x00\x00\x00\x00 A\x16" + b"\x00" * 11
            + b"\x00\x00\x80?\x00\x00@@\x00\x00\xa0@\x00\x00\x00\x00\x00\x00\xa0A\x0b" + b"\x00" * 11)
    fields = [co.Field("x", 0, 7), co.Field("y", 4, 7), co.Field("z", 8, 7), co.Field("intensity", 16, 7), co.Field("ring", 20, 2)]
    out, keep = co.convert(data, fields, 32, False)
    assert keep.all()
    vals = [struct.unpack("<fffffH10x", out[i].tobytes()) for i in range(2)]
    assert vals == [(2., 4., 6., 0., 10., 22), (1., 3., 5., 0., 20., 11)]


def test_node_vector_of_reference_test():
    """make_input_cloud + check_output_cloud, test_convert.py:175-216: ((1,2,3,0,10,2), (4,6,8,0,20,8))."""
    pts = ((1., 2., 3., 10., 50, 100, 2), (4., 6., 8., 20., 40, 200, 8))
    data = b"".join(struct.pack("<fff4xfIHB5x", *p) for p in pts)
    out, _ = co.convert(data, OUSTER[:7], 32, False)
    vals = [struct.unpack("<fffffH10x", out[i].tobytes()) for i in range(2)]
    assert vals == [(1., 2., 3., 0., 10., 2), (4., 6., 8., 0., 20., 8)]
    assert out.shape == (2, 32)


def test_data_size_must_be_multiple_of_point_step():
    with pytest.raises(co.ConvertError):
        co.convert(np.zeros(33, np.uint8), OUSTER[:7], 32, False)
