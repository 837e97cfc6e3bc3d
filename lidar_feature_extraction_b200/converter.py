"""Host-side mirror of the reference's point_type_converter node over the CUDA library.

Mirrors ``PointTypeConverter`` (point_type_converter/point_type_converter/convert.py:171-212): ``callback(msg)``
takes the raw driver PointCloud2 published on /points_raw and returns the PointCloud2 the reference publishes on
/points_converted (convert.py:198-212: height 1, width = kept points, the six fields of make_fields, point_step
32, little-endian, is_dense). Where the reference's callback raises (struct.error, IndexError, OverflowError,
KeyError) this raises ``ConvertError``. All conversion work happens in ``k_convert`` (csrc/lfx_convert.cuh);
nothing under ``oracle/`` is imported.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _native as N
from .extraction import FLOAT32, POINT_STEP, UINT16, ExtractionError, FeatureExtraction, PointCloud2, PointField, _is_cuda_tensor


class ConvertError(ValueError):
    """The reference converter raises for this cloud; ``status`` is the LFX_CONVERT_* reason."""

    def __init__(self, status: int):
        super().__init__(N.CONVERT_STATUS_NAMES[status] if 0 <= status < len(N.CONVERT_STATUS_NAMES) else str(status))
        self.status = status


def output_fields() -> list:
    """make_fields, convert.py:137-145."""
    return [PointField("x", 0, FLOAT32), PointField("y", 4, FLOAT32), PointField("z", 8, FLOAT32),
            PointField("padding", 12, FLOAT32), PointField("intensity", 16, FLOAT32), PointField("ring", 20, UINT16)]


class PointTypeConverter:
    """``extraction``: the FeatureExtraction whose handle (device, stream, buffers) the converter shares, so that
    converted clouds can be handed to ``extract_views`` without leaving the device; one is created if omitted."""

    def __init__(self, extraction: FeatureExtraction | None = None, device: int = 0):
        self._own = extraction is None
        self.extraction = extraction or FeatureExtraction(device=device)
        self._lib = N.lib()
        self._keep = None

    def close(self):
        if self._own and self.extraction is not None:
            self.extraction.close()
        self.extraction = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- batched entry
    def marshal(self, msgs: Sequence[PointCloud2]):
        """The C array of lfx_raw_cloud for ``msgs`` (with the Python owners of everything it points to). Build it once
        when the same messages are converted repeatedly: marshalling a thousand structs costs milliseconds."""
        keep, raws = self._marshal(msgs)
        carr = (N.RawCloud * max(len(raws), 1))(*raws)
        carr.n_clouds = len(raws)
        carr.keep = keep
        return carr

    def convert_batch(self, msgs) -> N.ConvertResult:
        """One launch for all clouds (``msgs``: PointCloud2 messages, or the result of ``marshal``). Returns the
        lfx_convert_result (device pointers + per-cloud status); does not raise for per-cloud failures - see
        ``status_of`` / ``fetch``."""
        h = self.extraction.handle
        carr = msgs if hasattr(msgs, "n_clouds") else self.marshal(msgs)
        res = N.ConvertResult()
        rc = self._lib.lfx_convert_batch(h, carr, carr.n_clouds, C.byref(res))
        if rc not in (N.LFX_OK, N.LFX_E_CONVERT):
            raise ExtractionError(rc, self._lib.lfx_last_error(h).decode())
        self._keep = carr
        self._res = res
        return res

    def views(self):
        """view_array of every successfully converted cloud of the last batch (one C call), for ``extract_views``."""
        n = len(self._keep.keep) if self._keep is not None else 0
        arr = (N.CloudView * max(n, 1))()
        got = C.c_int(0)
        rc = self._lib.lfx_converted_views(self.extraction.handle, arr, n, C.byref(got))
        if rc != N.LFX_OK:
            raise ExtractionError(rc, self._lib.lfx_last_error(self.extraction.handle).decode())
        arr.n_views = got.value
        return arr

    def _marshal(self, msgs: Sequence[PointCloud2]):
        keep, raws = [], []
        field_cache = {}   # clouds of one driver share their field table: marshal it once
        for m in msgs:
            key = id(m.fields)
            hit = field_cache.get(key)
            if hit is None:
                names = [f.name.encode() for f in m.fields]
                arr = (N.PointFieldC * max(len(m.fields), 1))(*[N.PointFieldC(nm, f.offset, f.datatype, f.count) for nm, f in zip(names, m.fields)])
                hit = field_cache[key] = (names, arr, len(m.fields))
            names, arr, nf = hit
            data = m.data
            if _is_cuda_tensor(data):
                ptr, nbytes, mem = data.data_ptr(), data.numel() * data.element_size(), N.LFX_MEM_DEVICE
            else:
                data = np.ascontiguousarray(np.frombuffer(data, np.uint8) if isinstance(data, (bytes, bytearray)) else data).reshape(-1).view(np.uint8)
                ptr, nbytes, mem = data.ctypes.data, data.nbytes, N.LFX_MEM_HOST
            keep.append((names, arr, data, m.fields))
            raws.append(N.RawCloud(ptr, nbytes, m.point_step, arr, nf, 1 if m.is_bigendian else 0, mem))
        return keep, raws

    def status_of(self, cloud: int) -> int:
        return int(self._res.status[cloud])

    def kept(self, cloud: int) -> int:
        return int(self._res.kept[cloud])

    def view(self, cloud: int) -> N.CloudView:
        """Device view of a converted cloud, for ``FeatureExtraction.extract_views``."""
        v = N.CloudView()
        rc = self._lib.lfx_converted_view(self.extraction.handle, cloud, C.byref(v))
        if rc == N.LFX_E_CONVERT:
            raise ConvertError(self.status_of(cloud))
        if rc != N.LFX_OK:
            raise ExtractionError(rc, self._lib.lfx_last_error(self.extraction.handle).decode())
        return v

    def fetch(self, cloud: int) -> np.ndarray:
        """[kept, 32] uint8: PointCloud2.data of the converted cloud."""
        st = self.status_of(cloud)
        if st != 0:
            raise ConvertError(st)
        out = np.zeros((self.kept(cloud), POINT_STEP), np.uint8)
        rc = self._lib.lfx_fetch_converted(self.extraction.handle, cloud, out.ctypes.data, out.nbytes)
        if rc != N.LFX_OK:
            raise ExtractionError(rc, self._lib.lfx_last_error(self.extraction.handle).decode())
        return out

    # -- the ROS-callback-shaped entry (convert.py:183-212)
    def callback(self, msg: PointCloud2) -> PointCloud2:
        self.convert_batch([msg])
        data = self.fetch(0)
        return PointCloud2(data=data, fields=output_fields(), point_step=POINT_STEP, width=data.shape[0], height=1,
                           is_dense=True, is_bigendian=False, stamp=msg.stamp, frame_id=msg.frame_id)
