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


def gather_counts(counts_local, n_frames: int, group=None):
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
    recv = torch.empty((world * width, 2), dtype=torch.int32, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    if all(s == width for s in sizes):
        return recv
    return torch.cat([recv[r * width: r * width + sizes[r]] for r in range(world)], dim=0)


def device_counts_tensor(res, device):
    """Zero-copy int32 view [n_scans, 2] of the library-owned per-scan counts of a batch result."""
    import torch

    class _View:
        pass

    v = _View()
    v.__cuda_array_interface__ = {"shape": (int(res.n_scans), 2), "typestr": "<i4", "data": (int(res.d_counts), False),
                                  "version": 3, "strides": None}
    return torch.as_tensor(v, device=device)


class ShardedExtraction:
    """One rank of the sharded offline driver: owns one FeatureExtraction handle on its GPU and the frames
    [lo, hi) of an n_frames sequence. `step(views)` enqueues the extraction of the shard and the count
    all-gather on the same stream; `offsets()` gives the global frame-ordered offsets."""

    def __init__(self, fe, n_frames: int, device, group=None):
        import torch.distributed as dist

        self.fe = fe
        self.n_frames = n_frames
        self.device = device
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_range(n_frames, self.rank, self.world)
        self.counts_all = None

    def step(self, views, keep=None):
        res = self.fe.extract_views(views, keep=keep)
        cnt = device_counts_tensor(res, self.device)
        self.counts_all = gather_counts(cnt, self.n_frames, self.group) if self.world > 1 else cnt
        return res

    def offsets(self) -> np.ndarray:
        return global_offsets(self.counts_all.cpu().numpy())
