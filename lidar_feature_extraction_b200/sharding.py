"""Multi-GPU driver logic: frames shard by index, one process per GPU, and the path's only exchange is an
all-gather of the per-scan (n_edge, n_surface) counts, from which every rank derives the global offsets of
the concatenated feature clouds (BASELINE.json north_star). No feature data crosses GPUs.

torch.distributed is plumbing here: backend "nccl" on the GPU box (the gather runs on the extraction
stream, over NVLink), "gloo" in the CPU tests of this logic (tests/test_sharding.py, world_size 2).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of frame indices owned by `rank`: [rank * F / G, (rank + 1) * F / G)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def shard_sizes(n_frames: int, world: int) -> list[int]:
    return [shard_range(n_frames, r, world)[1] - shard_range(n_frames, r, world)[0] for r in range(world)]


def global_offsets(counts_all: np.ndarray) -> np.ndarray:
    """counts_all: [n_frames, 2] in frame order -> [n_frames + 1, 2] exclusive prefix (last row = totals),
    i.e. where each scan's edge / surface cloud starts in the frame-ordered concatenation."""
    c = np.asarray(counts_all, dtype=np.int64).reshape(-1, 2)
    out = np.zeros((c.shape[0] + 1, 2), dtype=np.int64)
    np.cumsum(c, axis=0, out=out[1:])
    return out


def gather_counts(counts_local, n_frames: int, group=None, out=None):
    """All-gather of the per-scan counts. counts_local: int32 tensor [shard size, 2] on the device the backend
    works with (CUDA for nccl, CPU for gloo). Returns an int32 tensor [n_frames, 2] in frame order on the same
    device. Shards differ by at most one frame, so every rank contributes ceil(F / G) rows (zero padded)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_frames, world)
    if counts_local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {counts_local.shape[0]} scans, its shard has {sizes[rank]}")
    width = max(sizes) if sizes else 0
    send = counts_local.to(torch.int32).contiguous()
    if send.shape[0] < width:
        pad = torch.zeros((width - send.shape[0], 2), dtype=torch.int32, device=send.device)
        send = torch.cat([send, pad], dim=0)
    recv = out if out is not None else torch.empty((world * width, 2), dtype=torch.int32, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if all(s == width for s in sizes):
        return recv
    return torch.cat([recv[r * width: r * width + sizes[r]] for r in range(world)], dim=0)


_COUNT_VIEWS: dict = {}


def device_counts_tensor(res, device):
    """Zero-copy int32 view [n_scans, 2] of the library-owned per-scan counts of a batch result."""
    import torch

    class _View:
        pass

    key = (int(res.d_counts), int(res.n_scans), str(device))
    hit = _COUNT_VIEWS.get(key)
    if hit is not None:
        return hit
    v = _View()
    v.__cuda_array_interface__ = {"shape": (int(res.n_scans), 2), "typestr": "<i4", "data": (int(res.d_counts), False),
                                  "version": 3, "strides": None}
    t = torch.as_tensor(v, device=device)
    _COUNT_VIEWS.clear()   # the library's array only moves when it grows: one live view is enough
    _COUNT_VIEWS[key] = t
    return t


class ShardedExtraction:
    """One rank of the sharded offline driver: owns one FeatureExtraction handle on its GPU and the frames
    [lo, hi) of an n_frames sequence. `step(views)` enqueues the extraction of the shard and the count all-gather
    on the extraction stream; `offsets()` gives the global frame-ordered offsets of the last step.

    overlap=True moves the gather to a side stream (counts copied into one of two rotating buffers first, because
    the next batch overwrites the library's array). Measured on 2 x B200 it is SLOWER (4.91 vs 4.47 ms per step):
    the persistent sector kernel fills every SM, so the NCCL kernel cannot run beside it and only delays the
    next batch; kept for experiments, off by default."""

    def __init__(self, fe, n_frames: int, device, group=None, overlap: bool = False):
        import torch.distributed as dist

        self.fe = fe
        self.n_frames = n_frames
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_range(n_frames, self.rank, self.world)
        self.counts_all = None
        self.overlap = overlap and self.world > 1
        self._side = None
        self._recv = None
        self._slots = [None, None]
        self._done = [None, None]
        self._k = 0
        self._main = None
        self.time_gather = False      # diagnosis: CUDA events around the exchange on the extraction stream
        self.gather_events = []

    def _extraction_stream(self):
        """torch view of the stream the handle launches on: the gather is enqueued THERE (not on whatever torch's
        current stream happens to be), so it is ordered after the batch that writes the counts."""
        import torch

        if self._main is None:
            self._main = torch.cuda.ExternalStream(self.fe.stream, device=self.device)
        return self._main

    def step(self, views, keep=None):
        import torch

        res = self.fe.extract_views(views, keep=keep)
        cnt = device_counts_tensor(res, self.device)
        if self.world == 1:
            self.counts_all = cnt
            return res
        with torch.cuda.stream(self._extraction_stream()):
            if not self.time_gather:
                return self._step_gather(res, cnt)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            self._step_gather(res, cnt)
            e1.record()
            self.gather_events.append((e0, e1))
            return res

    def gather_ms(self):
        """Device time per step between the end of the batch and the end of the exchange (time_gather=True)."""
        self.fe.synchronize()
        return [a.elapsed_time(b) for a, b in self.gather_events]

    def _step_gather(self, res, cnt):
        import torch

        if not self.overlap:
            width = max(shard_sizes(self.n_frames, self.world))
            if self._recv is None or self._recv.shape[0] != self.world * width:
                self._recv = torch.empty((self.world * width, 2), dtype=torch.int32, device=cnt.device)
            self.counts_all = gather_counts(cnt, self.n_frames, self.group, out=self._recv)
            return res
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.device)
        k = self._k & 1
        self._k += 1
        if self._done[k] is not None:
            main.wait_event(self._done[k])          # the gather that read this slot two steps ago is finished
        if self._slots[k] is None or self._slots[k].shape != cnt.shape:
            self._slots[k] = torch.empty_like(cnt)
        self._slots[k].copy_(cnt, non_blocking=True)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            self.counts_all = gather_counts(self._slots[k], self.n_frames, self.group)
            self._done[k] = torch.cuda.Event()
            self._done[k].record(self._side)
        return res

    def join(self):
        """Make the extraction stream wait for the outstanding gathers (call before reading counts_all / timing)."""
        import torch

        if self._side is not None:
            self._extraction_stream().wait_stream(self._side)

    def offsets(self) -> np.ndarray:
        self.join()
        self.fe.synchronize()      # the counts (and the gather, enqueued on the same stream) are complete
        return global_offsets(self.counts_all.cpu().numpy())


# ---------------------------------------------------------------------------------------------------------------
# The same driver over the C ABI (include/lfx.h: lfx_shard_*). The exchange is the library's own: a one-CTA kernel that
# stores the counts into every peer's receive buffer through peer-mapped memory (NVLink) and raises a flag; NCCL only
# sets the group up. torch.distributed is not needed at all (the unique id can travel by any means); when a process
# group exists it is used to hand the id around.

class ShardError(RuntimeError):
    pass


class AbiShard:
    """One rank of the sharded driver through lfx_shard_create / lfx_shard_exchange / lfx_shard_fetch.
    ``unique_id``: the LFX_SHARD_ID_BYTES bytes rank 0 got from ``AbiShard.unique_id()`` (None: world == 1, or taken
    from rank 0 through torch.distributed's default group when one is initialised)."""

    def __init__(self, fe, n_frames: int, rank: int = 0, world: int = 1, unique_id: bytes | None = None):
        import ctypes as C

        from . import _native as N

        self._lib = N.lib()
        self.fe, self.n_frames, self.rank, self.world = fe, int(n_frames), int(rank), int(world)
        self.lo, self.hi = shard_range(n_frames, rank, world)
        if world > 1 and unique_id is None:
            import torch.distributed as dist

            if not dist.is_initialized():
                raise ShardError("a unique id is needed (AbiShard.unique_id() on rank 0, handed to every rank)")
            box = [self.unique_id() if rank == 0 else None]
            dist.broadcast_object_list(box, src=0)
            unique_id = box[0]
        self._s = C.c_void_p()
        buf = C.create_string_buffer(bytes(unique_id), N.SHARD_ID_BYTES) if unique_id is not None else None
        rc = self._lib.lfx_shard_create(fe.handle, buf, rank, world, self.n_frames, C.byref(self._s))
        if rc != N.LFX_OK:
            raise ShardError(f"lfx_shard_create: {N.STATUS_NAMES[rc]}: {self._lib.lfx_last_error(fe.handle).decode()}")

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C

        from . import _native as N

        buf = C.create_string_buffer(N.SHARD_ID_BYTES)
        rc = N.lib().lfx_shard_unique_id(buf)
        if rc != N.LFX_OK:
            raise ShardError(f"lfx_shard_unique_id: {N.STATUS_NAMES[rc]}: {N.lib().lfx_last_error(None).decode()}")
        return buf.raw

    def _check(self, rc: int, what: str):
        from . import _native as N

        if rc != N.LFX_OK:
            raise ShardError(f"{what}: {N.STATUS_NAMES[rc]}: {self._lib.lfx_shard_last_error(self._s).decode()}")

    def step(self, views, keep=None):
        """Extraction of this rank's shard + publication of its counts, both enqueued on the handle's stream."""
        res = self.fe.extract_views(views, keep=keep)
        self._check(self._lib.lfx_shard_exchange(self._s), "lfx_shard_exchange")
        return res

    def exchange(self):
        self._check(self._lib.lfx_shard_exchange(self._s), "lfx_shard_exchange")

    def finish(self):
        """Enqueue the consumer side of the last exchange; returns the device-side tables (lfx_shard_result)."""
        import ctypes as C

        from . import _native as N

        r = N.ShardResult()
        self._check(self._lib.lfx_shard_finish(self._s, C.byref(r)), "lfx_shard_finish")
        return r

    def join(self):
        self.finish()

    def fetch(self):
        """(counts [n_frames, 2] uint32, offsets [n_frames + 1, 2] uint64) of the last exchange, frame order; synchronises."""
        counts = np.zeros((self.n_frames, 2), np.uint32)
        offsets = np.zeros((self.n_frames + 1, 2), np.uint64)
        self._check(self._lib.lfx_shard_fetch(self._s, counts.ctypes.data, offsets.ctypes.data), "lfx_shard_fetch")
        return counts, offsets

    def offsets(self) -> np.ndarray:
        return self.fetch()[1].astype(np.int64)

    def info(self) -> dict:
        import ctypes as C

        p2p, n = C.c_int(0), C.c_int(0)
        self._check(self._lib.lfx_shard_info(self._s, C.byref(p2p), C.byref(n)), "lfx_shard_info")
        return {"exchange": "peer stores over NVLink" if p2p.value else "ncclAllGather", "nccl_ranks": int(n.value)}

    def close(self):
        if getattr(self, "_s", None):
            self._lib.lfx_shard_destroy(self._s)
            self._s = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def local_group(extractions, n_frames: int) -> list:
    """All ranks in this process (lfx_shard_create_local): one AbiShard per FeatureExtraction, rank = position."""
    import ctypes as C

    from . import _native as N

    lib = N.lib()
    world = len(extractions)
    hs = (C.c_void_p * world)(*[fe.handle for fe in extractions])
    out = (C.c_void_p * world)()
    rc = lib.lfx_shard_create_local(hs, world, int(n_frames), out)
    if rc != N.LFX_OK:
        raise ShardError(f"lfx_shard_create_local: {N.STATUS_NAMES[rc]}: {lib.lfx_last_error(extractions[0].handle).decode()}")
    shards = []
    for g, fe in enumerate(extractions):
        s = AbiShard.__new__(AbiShard)
        s._lib, s.fe, s.n_frames, s.rank, s.world = lib, fe, int(n_frames), g, world
        s.lo, s.hi = shard_range(n_frames, g, world)
        s._s = C.c_void_p(out[g])
        shards.append(s)
    return shards
