"""CPU restatement (numpy) of the upstream point-type converter — TEST INFRASTRUCTURE, never the product.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module. It restates
PointTypeConverter.callback (point_type_converter/point_type_converter/convert.py:183-212) of the
reference, including the quirks of its struct-format construction:

  * a FLOAT32 field 'padding' at offset 12 is appended to the input fields, then the fields are sorted by
    offset with a stable sort (convert.py:185-186);
  * create_point_format (convert.py:69-81) never steps backwards: a field whose declared offset lies before
    the end of the previous field is read where the previous field ended (effective offset), and the
    per-point format is max(end of last field, point_step) bytes long, so struct.unpack rejects any
    non-empty cloud whose fields run past point_step (convert.py:90-97);
  * `count` is ignored (one value per field);
  * a point is dropped when the FIRST THREE fields in offset order compare equal to 0 (convert.py:165-166,
    192), whatever their names or types (-0.0 is zero, NaN is not);
  * the retained fields are those whose NAME is one of x, y, z, padding, intensity, ring, in input offset
    order (convert.py:118-126), and they are packed POSITIONALLY into '<fffffH' + 10 pad bytes
    (convert.py:137-145, 195-196): five float32 slots and one uint16 slot, little-endian, bytes 22..31 zero;
  * struct.pack raises when a kept point does not carry exactly six values, when the sixth is not an integer
    in [0, 65535], or when a finite value overflows float32 ('f' packing goes double -> float, round to
    nearest even; NaN payloads keep their top bits and become quiet).

Pinned against the reference's own code: tests/golden/convert_*.npz are outputs of the unmodified
convert.py functions (imported with stub rclpy / sensor_msgs modules by tests/golden/make_convert_golden.py),
and tests/test_convert_oracle.py replays the known-answer vectors of point_type_converter/test/test_convert.py.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# sensor_msgs/PointField datatype ids, convert.py:40-53
NP_TYPES = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 8: np.float64}
OUTPUT_NAMES = ("x", "y", "z", "padding", "intensity", "ring")   # make_fields, convert.py:137-145
OUTPUT_POINT_STEP = 32                                            # convert.py:134


class ConvertError(ValueError):
    """What surfaces in the reference as struct.error / IndexError / OverflowError inside the callback."""


@dataclass
class Field:
    name: str
    offset: int
    datatype: int
    count: int = 1


@dataclass
class Plan:
    """Resolved read plan of one cloud: (effective offset, datatype) of every field in offset order."""
    names: list
    eff_offsets: list
    datatypes: list
    point_size: int          # bytes struct.unpack expects per point
    retained: list           # indices (into the sorted fields) of the fields packed into the output
    layout_error: str | None


def make_plan(fields, point_step: int) -> Plan:
    """convert.py:184-186 (append padding, sort) + create_point_format (convert.py:69-81)."""
    fs = list(fields) + [Field("padding", 12, 7, 1)]
    fs = sorted(fs, key=lambda f: f.offset)   # stable, like Python's sorted in the reference
    eff, index = [], 0
    for f in fs:
        if f.datatype not in NP_TYPES:
            raise ConvertError(f"unknown datatype {f.datatype}")
        if index < f.offset:
            index = f.offset
        eff.append(index)
        index += np.dtype(NP_TYPES[f.datatype]).itemsize
    size = max(index, point_step)
    err = None
    if size != point_step:
        err = f"fields end at byte {index}, beyond point_step {point_step} (struct.unpack size mismatch)"
    retained = [i for i, f in enumerate(fs) if f.name in OUTPUT_NAMES]   # find_indices, convert.py:118-119
    return Plan([f.name for f in fs], eff, [f.datatype for f in fs], size, retained, err)


def _read(data: np.ndarray, n: int, step: int, off: int, datatype: int, big: bool) -> np.ndarray:
    dt = np.dtype(NP_TYPES[datatype]).newbyteorder(">" if big else "<")
    raw = np.lib.stride_tricks.as_strided(data[off:], shape=(n, dt.itemsize), strides=(step, 1))
    return np.ascontiguousarray(raw).view(dt).reshape(n)


def convert(data, fields, point_step: int, is_bigendian: bool = False):
    """Returns (out [n_kept, 32] uint8, kept_mask [n]). Raises ConvertError where the reference raises."""
    data = np.ascontiguousarray(np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8))
    if point_step <= 0 or data.size % point_step != 0:
        raise ConvertError("Data size must be mutiple of point step")   # convert.py:91-92
    n = data.size // point_step
    plan = make_plan(fields, point_step)
    if n == 0:
        return np.zeros((0, OUTPUT_POINT_STEP), np.uint8), np.zeros(0, bool)
    if plan.layout_error:
        raise ConvertError(plan.layout_error)
    if len(plan.names) < 3:
        raise ConvertError("nonzero() indexes the first three fields")   # IndexError in convert.py:166
    vals = [_read(data, n, point_step, plan.eff_offsets[i], plan.datatypes[i], is_bigendian) for i in range(len(plan.names))]
    with np.errstate(invalid="ignore"):
        keep = ~((vals[0] == 0) & (vals[1] == 0) & (vals[2] == 0))     # nonzero, convert.py:165-166
    nk = int(keep.sum())
    out = np.zeros((nk, OUTPUT_POINT_STEP), np.uint8)
    if nk == 0:
        return out, keep
    if len(plan.retained) != 6:
        raise ConvertError(f"{len(plan.retained)} retained fields, struct.pack expects 6")
    for slot, i in enumerate(plan.retained[:5]):
        v = vals[i][keep]
        with np.errstate(over="ignore", invalid="ignore"):
            d = v.astype(np.float64)       # struct.unpack yields Python floats / ints: exact for every type of the table
            f32 = d.astype(np.float32)     # PyFloat_Pack4: (float)x, round to nearest even
        if np.any(np.isinf(f32) & ~np.isinf(d)):
            raise ConvertError("float too large to pack with f format")   # OverflowError in the reference
        out[:, 4 * slot: 4 * slot + 4] = f32.astype("<f4").view(np.uint8).reshape(nk, 4)
    i = plan.retained[5]
    v = vals[i][keep]
    if v.dtype.kind == "f":
        raise ConvertError("required argument is not an integer")          # struct.error, 'H' format
    v64 = v.astype(np.int64)
    if np.any((v64 < 0) | (v64 > 65535)):
        raise ConvertError("ushort format requires 0 <= number <= 65535")
    out[:, 20:22] = v64.astype("<u2").view(np.uint8).reshape(nk, 2)
    return out, keep


def output_fields():
    """make_fields, convert.py:137-145."""
    return [Field("x", 0, 7), Field("y", 4, 7), Field("z", 8, 7), Field("padding", 12, 7), Field("intensity", 16, 7), Field("ring", 20, 4)]
