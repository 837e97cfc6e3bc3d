#!/usr/bin/env python
"""Generates tests/golden/convert_*.npz: raw clouds and what the REFERENCE's own converter makes of them.

The outputs come from the unmodified point_type_converter/point_type_converter/convert.py of /root/reference,
imported here with stand-in `rclpy` / `sensor_msgs` modules (ROS 2 is not installed): the stand-ins only provide
the Node base class, the QoS names and the two message containers; every byte of the conversion is the
reference's own struct-based code (PointTypeConverter.callback, convert.py:183-212).
Run where /root/reference exists:   python tests/golden/make_convert_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LFX_REFERENCE", "/root/reference")


def install_stubs():
    class PointField:
        def __init__(self, name="", offset=0, datatype=0, count=0):
            self.name, self.offset, self.datatype, self.count = name, offset, datatype, count

    class PointCloud2:
        def __init__(self):
            self.header = None
            self.height = self.width = 0
            self.fields = []
            self.is_bigendian = False
            self.point_step = self.row_step = 0
            self.data = b""
            self.is_dense = False

    class _Publisher:
        def __init__(self):
            self.messages = []

        def publish(self, msg):
            self.messages.append(msg)

    class Node:
        def __init__(self, name, *a, **k):
            pass

        def create_subscription(self, *a, **k):
            return object()

        def create_publisher(self, *a, **k):
            return _Publisher()

    rclpy = types.ModuleType("rclpy")
    node = types.ModuleType("rclpy.node")
    node.Node = Node
    qos = types.ModuleType("rclpy.qos")

    class _Enum:
        BEST_EFFORT = RELIABLE = KEEP_ALL = 0

    qos.QoSHistoryPolicy = qos.QoSReliabilityPolicy = _Enum
    qos.QoSProfile = lambda **k: k
    rclpy.node, rclpy.qos = node, qos
    sm = types.ModuleType("sensor_msgs")
    msg = types.ModuleType("sensor_msgs.msg")
    msg.PointCloud2, msg.PointField = PointCloud2, PointField
    sm.msg = msg
    sys.modules.update({"rclpy": rclpy, "rclpy.node": node, "rclpy.qos": qos, "sensor_msgs": sm, "sensor_msgs.msg": msg})
    return PointCloud2, PointField


def load_reference_converter():
    PointCloud2, PointField = install_stubs()
    path = os.path.join(REF, "point_type_converter", "point_type_converter", "convert.py")
    spec = importlib.util.spec_from_file_location("ref_convert", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, PointCloud2, PointField


OUSTER = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("t", 20, 6), ("reflectivity", 24, 4),
          ("ring", 26, 2), ("noise", 28, 4), ("range", 32, 6)]          # test_convert.py:43-53, point_step 48
REFTEST = OUSTER[:7]                                                      # test_convert.py:177-187, point_step 32
PLAIN = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("ring", 20, 4)]
INTS = [("x", 0, 3), ("y", 2, 3), ("z", 4, 5), ("intensity", 16, 4), ("ring", 20, 6)]
DOUBLES = [("x", 0, 8), ("y", 8, 8), ("z", 16, 8), ("intensity", 24, 7), ("ring", 28, 4)]
VELODYNE = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 12, 7), ("ring", 16, 4), ("time", 18, 7)]
RING_F = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("ring", 20, 7)]
RING_I = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 7), ("ring", 20, 5)]
WITH_PAD = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("padding", 12, 7), ("intensity", 16, 7), ("ring", 20, 4)]
F64_INT = [("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("intensity", 16, 8), ("ring", 24, 2)]

NP = {1: "i1", 2: "u1", 3: "i2", 4: "u2", 5: "i4", 6: "u4", 7: "f4", 8: "f8"}


def random_cloud(seed, fields, step, n, big=False, zero_frac=0.15, specials=True, ring_max=128, huge_f64=False):
    """n points of `step` bytes: random junk everywhere, then plausible values in every declared field."""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, step), dtype=np.uint8)
    raw[:, 12:16] = 0 if step >= 16 else raw[:, 12:16]    # what drivers leave in the PCL padding lane
    e = ">" if big else "<"
    for name, off, dt in fields:
        t = np.dtype(e + NP[dt])
        if off + t.itemsize > step:
            continue
        if t.kind == "f":
            v = rng.normal(0, 20, n)
            if name == "intensity":
                v = rng.uniform(0, 255, n)
            if huge_f64 and t.itemsize == 8 and name == "intensity":
                v[::7] = 1e39
        elif name == "ring":
            v = rng.integers(min(0, ring_max), max(ring_max, 1), n)
        else:
            info = np.iinfo(t)
            v = rng.integers(max(info.min, -30000), min(info.max, 30000), n)
        raw[:, off: off + t.itemsize] = v.astype(t).view(np.uint8).reshape(n, t.itemsize)
    zero = rng.random(n) < zero_frac
    for name, off, dt in fields[:3]:
        t = np.dtype(e + NP[dt])
        raw[zero, off: off + t.itemsize] = 0
    if specials and fields[0][2] == 7 and n >= 16:
        f = np.dtype(e + "f4")
        u = np.dtype(e + "u4")
        def put(i, off, val, typ):
            raw[i, off: off + 4] = np.array([val], typ).view(np.uint8)
        put(1, 0, -0.0, f); put(1, 4, 0.0, f); put(1, 8, -0.0, f)          # all "zero": dropped
        put(2, 0, np.nan, f); put(2, 4, 0.0, f); put(2, 8, 0.0, f)         # NaN != 0: kept
        put(3, 0, 0x7FA00001, u)                                            # signalling NaN in x: comes out quiet
        put(4, 0, 0.0, f); put(4, 4, 0.0, f); put(4, 8, 1e-45, f)          # denormal z: kept
        put(5, 0, np.inf, f); put(6, 4, -np.inf, f)
        put(7, 16, 0xFFC12345, u) if step >= 20 else None                   # NaN payload in intensity
    return np.ascontiguousarray(raw)


def cases():
    yield "ouster48", OUSTER, 48, False, random_cloud(1, OUSTER, 48, 700)
    yield "ouster48_be", OUSTER, 48, True, random_cloud(2, OUSTER, 48, 300, big=True)
    yield "reftest32", REFTEST, 32, False, random_cloud(3, REFTEST, 32, 515)
    yield "plain32", PLAIN, 32, False, random_cloud(4, PLAIN, 32, 1025)
    yield "plain24_tight", PLAIN, 24, False, random_cloud(5, PLAIN, 24, 333)
    yield "ints24", INTS, 24, False, random_cloud(6, INTS, 24, 400, specials=False)
    yield "doubles40_quirk", DOUBLES, 40, False, random_cloud(7, DOUBLES, 40, 200, specials=False)
    yield "doubles32_overrun", DOUBLES, 32, False, random_cloud(8, DOUBLES, 32, 50, specials=False)
    yield "velodyne22_overrun", VELODYNE, 22, False, random_cloud(9, VELODYNE, 22, 64)
    yield "ring_float_error", RING_F, 32, False, random_cloud(10, RING_F, 32, 64)
    yield "ring_i32_negative_error", RING_I, 32, False, random_cloud(11, RING_I, 32, 64, ring_max=-5)
    yield "ring_i32_ok", RING_I, 32, False, random_cloud(12, RING_I, 32, 64, ring_max=65536)
    yield "already_converted_error", WITH_PAD, 32, False, random_cloud(13, WITH_PAD, 32, 64)
    yield "f64_intensity_overflow_error", F64_INT, 32, False, random_cloud(14, F64_INT, 32, 64, huge_f64=True)
    yield "f64_intensity_ok", F64_INT, 32, False, random_cloud(15, F64_INT, 32, 257)
    yield "all_zero", OUSTER, 48, False, random_cloud(16, OUSTER, 48, 100, zero_frac=2.0, specials=False)
    yield "all_zero_bad_fields", WITH_PAD, 32, False, random_cloud(17, WITH_PAD, 32, 40, zero_frac=2.0, specials=False)
    yield "empty_overrun", VELODYNE, 22, False, np.zeros((0, 22), np.uint8)
    yield "one_point", OUSTER, 48, False, random_cloud(18, OUSTER, 48, 1, zero_frac=0.0, specials=False)


def main():
    mod, PointCloud2, PointField = load_reference_converter()
    for name, fields, step, big, raw in cases():
        conv = mod.PointTypeConverter()
        msg = PointCloud2()
        msg.fields = [PointField(name=n, offset=o, datatype=d, count=1) for n, o, d in fields]
        msg.point_step, msg.is_bigendian = step, big
        msg.height, msg.width, msg.row_step = 1, raw.shape[0], step * raw.shape[0]
        msg.data = raw.tobytes()
        msg.is_dense = True
        error, out = "", np.zeros((0, 32), np.uint8)
        try:
            conv.callback(msg)
            o = conv.publisher.messages[-1]
            assert o.point_step == 32 and not o.is_bigendian and o.is_dense and o.height == 1
            out = np.frombuffer(bytes(o.data), np.uint8).reshape(-1, 32).copy()
            assert o.width == out.shape[0] and o.row_step == 32 * o.width
            assert [(f.name, f.offset, f.datatype) for f in o.fields] == [
                ("x", 0, 7), ("y", 4, 7), ("z", 8, 7), ("padding", 12, 7), ("intensity", 16, 7), ("ring", 20, 4)]
        except Exception as ex:  # struct.error, IndexError, OverflowError: the ROS callback would raise
            error = f"{type(ex).__name__}: {ex}"
        np.savez_compressed(
            os.path.join(HERE, f"convert_{name}.npz"), raw=raw, point_step=step, is_bigendian=big,
            field_names=np.array([f[0] for f in fields]), field_offsets=np.array([f[1] for f in fields]),
            field_datatypes=np.array([f[2] for f in fields]), out=out, error=np.array(error))
        print(f"{name:32s} n={raw.shape[0]:5d} kept={out.shape[0]:5d} {error}")


if __name__ == "__main__":
    main()
