"""Host-side mirror of the reference extraction node's interface, on top of the C ABI (include/lfx.h).

Mirrors, for the one path this repository replaces:
  * ``HyperParameters``      extraction/include/lidar_feature_extraction/hyper_parameter.hpp:32-65
  * ``FeatureExtraction``    extraction/app/feature_extraction.cpp:62-176 (``callback`` == ``Callback``)
  * ``PointCloud2``/``PointField``: the sensor_msgs subset the callback touches.

All compute happens in the CUDA library; nothing here computes features, and nothing under
``oracle/`` is ever imported.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Sequence

import numpy as np

from . import _native as N

POINT_STEP = 32  # deployed wire layout, point_type_converter/convert.py:134
FRAME_ID = "lidar_feature_base_link"  # feature_extraction.cpp:159

# sensor_msgs/PointField datatype ids
INT8, UINT8, INT16, UINT16, INT32, UINT32, FLOAT32, FLOAT64 = range(1, 9)


class ExtractionError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{N.STATUS_NAMES[code] if 0 <= code < len(N.STATUS_NAMES) else code}: {message}")
        self.code = code


@dataclass
class PointField:
    name: str
    offset: int
    datatype: int
    count: int = 1


@dataclass
class PointCloud2:
    """sensor_msgs/msg/PointCloud2 subset. ``data`` is a uint8 numpy array (host) or a CUDA tensor."""

    data: object
    fields: list = field(default_factory=list)
    point_step: int = POINT_STEP
    width: int = 0
    height: int = 1
    is_dense: bool = True
    is_bigendian: bool = False
    stamp: object = None
    frame_id: str = ""

    @staticmethod
    def from_wire(cloud: np.ndarray, stamp=None, frame_id: str = "lidar") -> "PointCloud2":
        """Wrap a [n, 32] uint8 cloud in the deployed layout (convert.py:137-145)."""
        c = np.ascontiguousarray(cloud).reshape(-1, POINT_STEP)
        return PointCloud2(
            data=c, point_step=POINT_STEP, width=c.shape[0], height=1, stamp=stamp, frame_id=frame_id,
            fields=[PointField("x", 0, FLOAT32), PointField("y", 4, FLOAT32), PointField("z", 8, FLOAT32),
                    PointField("padding", 12, FLOAT32), PointField("intensity", 16, FLOAT32),
                    PointField("ring", 20, UINT16)])


@dataclass
class HyperParameters:
    """Same nine parameters, names and defaults as hyper_parameter.hpp:35-43 (ROS name in comments)."""

    padding: int = 5                              # convolution_padding
    neighbor_degree_threshold: float = 2.0
    distance_diff_threshold: float = 0.3
    parallel_beam_min_range_ratio: float = 0.02
    edge_threshold: float = 0.05
    surface_threshold: float = 0.05
    min_range: float = 0.1
    max_range: float = 100.0
    n_blocks: int = 6

    def to_c(self) -> N.Params:
        return N.Params(self.padding, self.neighbor_degree_threshold, self.distance_diff_threshold,
                        self.parallel_beam_min_range_ratio, self.edge_threshold, self.surface_threshold,
                        self.min_range, self.max_range, self.n_blocks)

    @staticmethod
    def from_c(p: N.Params) -> "HyperParameters":
        return HyperParameters(*(getattr(p, n) for n, _ in N.Params._fields_))


def default_params() -> HyperParameters:
    p = N.Params()
    N.lib().lfx_default_params(C.byref(p))
    return HyperParameters.from_c(p)


def launch_yaml_params() -> HyperParameters:
    """lidar_feature_launch/config/lidar_feature_extraction.param.yaml:3-10."""
    p = N.Params()
    N.lib().lfx_launch_yaml_params(C.byref(p))
    return HyperParameters.from_c(p)


def label_to_color(label: int) -> tuple:
    rgb = (C.c_uint8 * 3)()
    rc = N.lib().lfx_label_to_color(label, rgb)
    if rc != N.LFX_OK:
        raise ValueError(f"Invalid label {label}")  # color_points.cpp:33-37
    return tuple(rgb)


def topic_fields(topic: int) -> tuple:
    """(fields, point_step) of one of the node's output topics (lfx_topic_layout; ros_msg.hpp:53-71)."""
    arr = (N.PointFieldC * 4)()
    n, step = C.c_uint32(), C.c_uint32()
    rc = N.lib().lfx_topic_layout(topic, arr, C.byref(n), C.byref(step))
    if rc != N.LFX_OK:
        raise ValueError(f"unknown topic {topic}")
    return [PointField(arr[k].name.decode(), arr[k].offset, arr[k].datatype, arr[k].count) for k in range(n.value)], step.value


def _is_cuda_tensor(obj) -> bool:
    return hasattr(obj, "is_cuda") and bool(obj.is_cuda)


@dataclass
class BatchOutput:
    """Host copies of one batch's results (what lfx_fetch_* return)."""

    counts: np.ndarray          # [n_scans, 2] (n_edge, n_surface)
    offsets: np.ndarray         # [n_scans + 1, 2]
    edge_xyz: np.ndarray        # [sum n_edge, 4] f32
    surface_xyz: np.ndarray     # [sum n_surface, 4] f32
    labels: np.ndarray | None   # [total_points] u8, ring-sorted order
    sorted_src: np.ndarray | None
    curvature: np.ndarray | None
    rings: np.ndarray | None    # structured [n_scans, max_rings]
    point_base: np.ndarray      # [n_scans + 1]

    def scan_edges(self, s: int) -> np.ndarray:
        return self.edge_xyz[self.offsets[s, 0]: self.offsets[s, 0] + self.counts[s, 0]]

    def scan_surfaces(self, s: int) -> np.ndarray:
        return self.surface_xyz[self.offsets[s, 1]: self.offsets[s, 1] + self.counts[s, 1]]


RING_DTYPE = np.dtype([("count", "<u4"), ("offset", "<u4"), ("n_edge", "<u4"), ("n_surface", "<u4"),
                       ("status", "<u4"), ("order_path", "<u4")])


class FeatureExtraction:
    """Mirror of the reference node class (feature_extraction.cpp:62-176) over the CUDA library.

    ``callback(msg)`` is the per-scan entry the ROS shim calls; ``extract_batch`` is the batched
    entry used offline (scans are independent: the reference callback is const and stateless).
    """

    def __init__(self, params: HyperParameters | None = None, device: int = 0, max_ring_points: int = 0,
                 max_rings: int = 0, want_sorted_src: bool = False, want_curvature: bool = False,
                 force_order_path: int = 0, stream: int | None = None, use_graph: bool = True):
        self._lib = N.lib()
        self.params = params or HyperParameters()
        opt = N.Options(device, max_ring_points, max_rings, int(want_sorted_src), int(want_curvature),
                        force_order_path, stream, 0 if use_graph else -1)
        h = C.c_void_p()
        p = self.params.to_c()
        rc = self._lib.lfx_create(C.byref(p), C.byref(opt), C.byref(h))
        if rc != N.LFX_OK:
            raise ExtractionError(rc, self._lib.lfx_last_error(None).decode())
        self._h = h
        self.want_sorted_src = want_sorted_src
        self.want_curvature = want_curvature
        self.max_rings = max_rings or 128
        self._last = None
        # owners of the input buffers of the last TWO batches: lfx_extract_batch returns before batch k's host-to-device
        # copies have finished (two descriptor slots, lfx.h), so batch k's owners may only go when batch k+2 is enqueued
        self._keep = None
        self._keep_prev = None

    # -- lifecycle
    def close(self):
        if getattr(self, "_h", None):
            self._lib.lfx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != N.LFX_OK:
            raise ExtractionError(rc, self._lib.lfx_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        """cudaStream_t (as an integer) every kernel of this handle runs on: the one passed in, else the handle's own."""
        return int(self._lib.lfx_stream(self._h) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.lfx_kernel_launch_count(self._h))

    # -- views
    @staticmethod
    def view_of(msg: PointCloud2) -> N.CloudView:
        """Field lookup by name, as pcl::fromROSMsg does for PointXYZIR (point_type.hpp:83-86)."""
        by_name = {f.name: f for f in msg.fields}
        for name in ("x", "y", "z"):
            if name not in by_name or by_name[name].datatype != FLOAT32:
                raise ExtractionError(N.LFX_E_BAD_LAYOUT, f"field {name!r} (FLOAT32) is required")
        ring = by_name.get("ring")  # RingIsAvailable, ring.cpp:36-44
        data = msg.data
        if _is_cuda_tensor(data):
            ptr, mem = data.data_ptr(), N.LFX_MEM_DEVICE
            nbytes = data.numel() * data.element_size()
        else:
            data = np.ascontiguousarray(data)
            ptr, mem, nbytes = data.ctypes.data, N.LFX_MEM_HOST, data.nbytes
        n = msg.width * msg.height if msg.width else nbytes // msg.point_step
        if msg.point_step <= 0 or n * msg.point_step > nbytes:   # an inconsistent message must not become an out-of-bounds read
            raise ExtractionError(N.LFX_E_BAD_LAYOUT, f"width * height * point_step = {n * msg.point_step} exceeds the {nbytes} data bytes")
        for f in (by_name["x"], by_name["y"], by_name["z"], ring):
            if f is not None and f.offset + (4 if f is not ring else {N.LFX_RING_U8: 1, N.LFX_RING_U16: 2}.get(f.datatype, 4)) > msg.point_step:
                raise ExtractionError(N.LFX_E_BAD_LAYOUT, f"field {f.name!r} at offset {f.offset} does not fit point_step {msg.point_step}")
        return N.CloudView(ptr, n, msg.point_step, by_name["x"].offset, by_name["y"].offset, by_name["z"].offset,
                           ring.offset if ring else 0, ring.datatype if ring else N.LFX_RING_U16,
                           1 if ring else 0, 1 if msg.is_dense else 0, mem)

    @staticmethod
    def wire_view(data, n_points: int | None = None) -> N.CloudView:
        """View of a buffer in the deployed 32-byte layout (numpy uint8 array or CUDA tensor / raw pointer)."""
        if isinstance(data, tuple):  # (device_ptr, n_points)
            return N.CloudView(data[0], data[1], POINT_STEP, 0, 4, 8, 20, N.LFX_RING_U16, 1, 1, N.LFX_MEM_DEVICE)
        if _is_cuda_tensor(data):
            n = data.numel() * data.element_size() // POINT_STEP if n_points is None else n_points
            return N.CloudView(data.data_ptr(), n, POINT_STEP, 0, 4, 8, 20, N.LFX_RING_U16, 1, 1, N.LFX_MEM_DEVICE)
        n = data.nbytes // POINT_STEP if n_points is None else n_points
        return N.CloudView(data.ctypes.data, n, POINT_STEP, 0, 4, 8, 20, N.LFX_RING_U16, 1, 1, N.LFX_MEM_HOST)

    # -- batched entry
    @staticmethod
    def view_array(views: Sequence[N.CloudView]):
        """The C array lfx_extract_batch takes; build it once when the same views are extracted repeatedly
        (marshalling a thousand structs costs more host time than the GPU needs for the batch)."""
        arr = (N.CloudView * max(len(views), 1))(*views)
        arr.n_views = len(views)
        return arr

    def extract_views(self, views, keep=None) -> N.BatchResult:
        """Enqueue one batch (asynchronous). ``views``: a sequence of CloudView or a ``view_array``.
        ``keep`` pins Python owners of the input buffers."""
        arr = views if hasattr(views, "n_views") else self.view_array(views)
        res = N.BatchResult()
        self._check(self._lib.lfx_extract_batch(self._h, arr, arr.n_views, C.byref(res)))
        self._last = res
        self._keep_prev, self._keep = self._keep, keep
        return res

    def extract_batch(self, scans: Sequence, fetch_points: bool = True) -> BatchOutput:
        """scans: wire-layout uint8 arrays / CUDA tensors, or PointCloud2 messages."""
        keep, views = [], []
        for s in scans:
            if isinstance(s, PointCloud2):
                if not _is_cuda_tensor(s.data):
                    s.data = np.ascontiguousarray(s.data)
                keep.append(s.data)
                views.append(self.view_of(s))
            else:
                if not _is_cuda_tensor(s):
                    s = np.ascontiguousarray(s)
                keep.append(s)
                views.append(self.wire_view(s))
        self.extract_views(views, keep)
        return self.fetch(fetch_points=fetch_points)

    def synchronize(self):
        self._check(self._lib.lfx_synchronize(self._h))

    def fetch(self, fetch_points: bool = True, fetch_features: bool = True) -> BatchOutput:
        res = self._last
        if res is None:
            raise ExtractionError(N.LFX_E_STATE, "no batch has been extracted")
        ns = res.n_scans
        counts = np.zeros((ns, 2), dtype=np.uint32)
        offsets = np.zeros((ns + 1, 2), dtype=np.uint32)
        self._check(self._lib.lfx_fetch_counts(self._h, counts.ctypes.data, offsets.ctypes.data))
        self._check(self._lib.lfx_batch_status(self._h))
        ne, nsf = int(offsets[ns, 0]), int(offsets[ns, 1])
        edge = np.zeros((ne, 4), dtype=np.float32)
        surf = np.zeros((nsf, 4), dtype=np.float32)
        if fetch_features:
            self._check(self._lib.lfx_fetch_features(self._h, edge.ctypes.data, max(ne, 1), surf.ctypes.data, max(nsf, 1)))
        labels = sorted_src = curv = rings = None
        tp = int(res.total_points)
        if fetch_points:
            labels = np.zeros(tp, dtype=np.uint8)
            sorted_src = np.zeros(tp, dtype=np.uint32) if self.want_sorted_src else None
            curv = np.zeros(tp, dtype=np.float64) if self.want_curvature else None
            self._check(self._lib.lfx_fetch_points(
                self._h, labels.ctypes.data, sorted_src.ctypes.data if sorted_src is not None else None,
                curv.ctypes.data if curv is not None else None))
            rings = np.zeros((ns, res.max_rings), dtype=RING_DTYPE)
            self._check(self._lib.lfx_fetch_rings(self._h, rings.ctypes.data))
        pb = np.zeros(ns + 1, dtype=np.uint64)
        if ns >= 0:
            self._check(self._lib.lfx_memcpy_d2h(self._h, pb.ctypes.data, res.d_point_base, pb.nbytes))
        return BatchOutput(counts, offsets, edge, surf, labels, sorted_src, curv, rings, pb)

    # -- colored_scan (feature_extraction.cpp:153,161,168)
    def colored_scans(self) -> list:
        """[n_i, 32] uint8 PointXYZRGB records of every scan of the last batch (needs want_sorted_src)."""
        res = N.ColoredResult()
        self._check(self._lib.lfx_color_batch(self._h, C.byref(res)))
        out = []
        for s in range(res.n_scans):
            a = np.zeros((int(res.counts[s]), 32), np.uint8)
            self._check(self._lib.lfx_fetch_colored(self._h, s, a.ctypes.data, a.nbytes))
            out.append(a)
        return out

    # -- the ROS-callback-shaped entry (feature_extraction.cpp:92-171)
    def callback(self, msg: PointCloud2) -> dict:
        """One PointCloud2 in; returns {"scan_edge", "scan_surface", "colored_scan"} PointCloud2 messages
        stamped with the input stamp and frame ``lidar_feature_base_link`` (feature_extraction.cpp:159-166)."""
        if not msg.is_dense:  # :96-101
            raise ExtractionError(N.LFX_E_NOT_DENSE, "Point cloud is not in dense format, please remove NaN points first!")
        if not any(f.name == "ring" for f in msg.fields):  # :103-108
            raise ExtractionError(N.LFX_E_NO_RING, "Ring channel could not be found")
        if not _is_cuda_tensor(msg.data):
            msg.data = np.ascontiguousarray(msg.data)
        view = self.view_of(msg)
        out = N.ScanOutput()
        self._check(self._lib.lfx_extract_scan(self._h, C.byref(view), C.byref(out)))

        def xyz_msg(ptr, n, topic):
            pts = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(max(n, 1), 4))[:n].copy()
            fields, step = topic_fields(topic)
            return PointCloud2(data=pts.view(np.uint8).reshape(n, step), point_step=step, width=n, height=1,
                               stamp=msg.stamp, frame_id=FRAME_ID, fields=fields)

        labels = np.ctypeslib.as_array(C.cast(out.labels, C.POINTER(C.c_uint8)), shape=(max(out.n_points, 1),))[: out.n_points].copy()
        res = {"scan_edge": xyz_msg(out.edge_xyz, out.n_edge, N.LFX_TOPIC_SCAN_EDGE),
               "scan_surface": xyz_msg(out.surface_xyz, out.n_surface, N.LFX_TOPIC_SCAN_SURFACE),
               "labels": labels, "stamp": msg.stamp, "frame_id": FRAME_ID}
        if self.want_sorted_src:   # colored_scan is a debug topic: only built when the index map is kept
            colored = self.colored_scans()[0]
            fields, step = topic_fields(N.LFX_TOPIC_COLORED_SCAN)
            res["colored_scan"] = PointCloud2(data=colored, point_step=step, width=colored.shape[0], height=1, stamp=msg.stamp,
                                              frame_id=FRAME_ID, fields=fields)
        return res

    def batch_stats(self) -> dict:
        """Which path the last batch took: rings on the sector kernel per lane class, scans/rings on the general path."""
        st = N.BatchStats()
        self._check(self._lib.lfx_last_batch_stats(self._h, C.byref(st)))
        return {"fast_rings": list(st.fast_rings), "general_scans": int(st.general_scans), "general_rings": int(st.general_rings),
                "indexed_rings": list(st.indexed_rings)}

    # -- stage timing
    def set_stage_timing(self, enabled, in_graph: bool = False):
        """enabled: False/0 off, True/1 eager launches with events, 2 (or in_graph=True) events inside the graph."""
        mode = 2 if (enabled and in_graph) else int(enabled)
        self._check(self._lib.lfx_set_stage_timing(self._h, mode))

    def last_stage_ms(self):
        ms = (C.c_float * 6)()   # LFX_N_STAGES
        self._check(self._lib.lfx_last_stage_ms(self._h, ms))
        return tuple(ms)


class PipelinedExtraction:
    """Offline batches from HOST memory through two handles in turn (mirror of ``lfx::Pipeline``, include/lfx.hpp).

    ``submit(views)`` enqueues a batch on the idle handle and returns; ``collect(...)`` waits for the OLDEST batch in
    flight and copies its counts and feature clouds to the host. While batch k uploads on one stream, batch k-1's
    features download on the other (PCIe is full duplex) and batch k's kernels later hide under the upload of batch
    k+1. Scans are independent (feature_extraction.cpp:92,173-175): results equal one handle called batch by batch.
    The input buffers of a batch must stay valid until it has been collected.
    """

    def __init__(self, params: HyperParameters | None = None, device: int = 0, handles=None, **kw):
        """``handles``: two existing FeatureExtraction objects of the same device to run on (else two are created)."""
        self._own = handles is None
        self.fe = tuple(handles) if handles is not None else (FeatureExtraction(params, device, **kw), FeatureExtraction(params, device, **kw))
        assert len(self.fe) == 2 and self.fe[0] is not self.fe[1]
        self._head = 0
        self._in_flight = 0
        self._n_scans = [0, 0]

    def close(self):
        if self._own:
            for fe in self.fe:
                fe.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def in_flight(self) -> int:
        return self._in_flight

    def submit(self, views, keep=None) -> int:
        """Enqueue a batch; returns the slot (0 / 1) of the handle that took it."""
        if self._in_flight == 2:
            raise ExtractionError(N.LFX_E_STATE, "two batches in flight: collect() first")
        arr = views if hasattr(views, "n_views") else FeatureExtraction.view_array(views)
        slot = self._head
        self.fe[slot].extract_views(arr, keep=keep)
        self._n_scans[slot] = arr.n_views
        self._head ^= 1
        self._in_flight += 1
        return slot

    def collect(self, edge_ptr: int = 0, edge_capacity: int = 0, surface_ptr: int = 0, surface_capacity: int = 0):
        """Counts [n_scans, 2], offsets [n_scans + 1, 2] and, when destination pointers (pinned memory for full PCIe
        speed, ``lfx_host_alloc_on``) are given, the edge / surface clouds of the oldest batch in flight."""
        if self._in_flight == 0:
            raise ExtractionError(N.LFX_E_STATE, "no batch in flight")
        slot = self._head if self._in_flight == 2 else self._head ^ 1
        fe, ns = self.fe[slot], self._n_scans[slot]
        counts = np.zeros((ns, 2), np.uint32)
        offsets = np.zeros((ns + 1, 2), np.uint32)
        self._in_flight -= 1
        fe._check(fe._lib.lfx_fetch_counts(fe.handle, counts.ctypes.data, offsets.ctypes.data))
        fe._check(fe._lib.lfx_batch_status(fe.handle))
        if edge_ptr or surface_ptr:
            fe._check(fe._lib.lfx_fetch_features(fe.handle, edge_ptr or None, edge_capacity, surface_ptr or None, surface_capacity))
        return counts, offsets

    def collect_output(self) -> BatchOutput:
        """The oldest batch in flight as a full BatchOutput (numpy arrays; convenience for tests)."""
        if self._in_flight == 0:
            raise ExtractionError(N.LFX_E_STATE, "no batch in flight")
        slot = self._head if self._in_flight == 2 else self._head ^ 1
        self._in_flight -= 1
        return self.fe[slot].fetch(fetch_points=True)
