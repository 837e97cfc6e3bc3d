"""Host logic of the multi-GPU path (lidar_feature_extraction_b200/sharding.py) on CPU: frame sharding, and
the count all-gather + global offsets with world_size 2 over gloo (the GPU box runs the same code over nccl)."""
import os
import socket

import numpy as np
import pytest

from lidar_feature_extraction_b200 import sharding


def test_shard_ranges_partition_the_sequence():
    for n, world in ((10000, 8), (50000, 8), (7, 3), (1, 4), (0, 2), (1250, 1)):
        ranges = [sharding.shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = sharding.shard_sizes(n, world)
        assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    assert sharding.shard_range(10000, 3, 8) == (3750, 5000)   # BASELINE config 4: 1250 scans per GPU
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_global_offsets():
    counts = np.array([[3, 10], [0, 0], [5, 1]])
    assert sharding.global_offsets(counts).tolist() == [[0, 0], [3, 10], [3, 10], [8, 11]]
    assert sharding.global_offsets(np.zeros((0, 2))).tolist() == [[0, 0]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, out_dir):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(n_frames, rank, world)
        frames = np.arange(lo, hi)
        local = torch.from_numpy(np.stack([frames * 3 + 1, frames * 7 + 2], axis=1).astype(np.int32).reshape(-1, 2))
        allc = sharding.gather_counts(local, n_frames)
        np.save(os.path.join(out_dir, f"counts_{rank}.npy"), allc.numpy())
        bad = None
        try:
            sharding.gather_counts(local[:-1] if len(local) else torch.zeros((1, 2), dtype=torch.int32), n_frames)
        except ValueError as e:
            bad = str(e)
        assert bad is not None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [10, 7, 1])
def test_count_all_gather_world_size_2_gloo(tmp_path, n_frames):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_frames, str(tmp_path)), nprocs=world, join=True)
    frames = np.arange(n_frames)
    want = np.stack([frames * 3 + 1, frames * 7 + 2], axis=1)
    for r in range(world):
        got = np.load(tmp_path / f"counts_{r}.npy")
        assert np.array_equal(got, want), (r, got)     # every rank ends up with the frame-ordered table
    off = sharding.global_offsets(want)
    assert off[-1].tolist() == want.sum(axis=0).tolist()


def _gpu_worker(rank, world, port, out_dir, overlap=False):
    import torch
    import torch.distributed as dist

    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n_frames = 5
        sp = synth.spec("vlp16")
        # no stream given: the handle creates its own (non-blocking) stream, and the driver must enqueue the
        # gather on THAT stream (lfx_stream), not on torch's current one
        fe = FeatureExtraction(HyperParameters(), device=rank)
        assert fe.stream != 0
        drv = sharding.ShardedExtraction(fe, n_frames, dev, overlap=overlap)
        clouds = [torch.from_numpy(synth.scan_host(sp, f)).to(dev) for f in range(drv.lo, drv.hi)]
        for _ in range(3):
            drv.step([fe.wire_view(c) for c in clouds], keep=clouds)
        offs = drv.offsets()   # joins and synchronises the extraction stream
        assert offs.shape == (n_frames + 1, 2)
        np.save(os.path.join(out_dir, f"gpu_counts_{rank}.npy"), drv.counts_all.cpu().numpy())
        fe.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("overlap", [False, True], ids=["gather_on_extraction_stream", "gather_on_side_stream"])
def test_sharded_extraction_two_gpus_nccl(tmp_path, oracle, overlap):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    mp.spawn(_gpu_worker, args=(2, _free_port(), str(tmp_path), overlap), nprocs=2, join=True)
    sp = synth.spec("vlp16")
    want = []
    for f in range(5):
        r = oracle.extract_scan(synth.scan_host(sp, f), ob.default_params())
        want.append([len(r.edge_idx), len(r.surface_idx)])
    for r in range(2):
        assert np.load(tmp_path / f"gpu_counts_{r}.npy").tolist() == want


# ---- the C-ABI driver (lfx_shard_*): NCCL set-up, counts pushed through peer-mapped memory --------------------------

def _oracle_counts(n_frames, oracle):
    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    sp = synth.spec("vlp16")
    want = []
    for f in range(n_frames):
        r = oracle.extract_scan(synth.scan_host(sp, f), ob.default_params())
        want.append([len(r.edge_idx), len(r.surface_idx)])
    return np.array(want, np.uint32)


@pytest.mark.gpu
def test_abi_shard_single_rank(oracle):
    """world == 1 goes through the same kernels (a group of one)."""
    import torch

    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, synth

    sp = synth.spec("vlp16")
    n_frames = 3
    with FeatureExtraction(HyperParameters(), device=0) as fe:
        drv = sharding.AbiShard(fe, n_frames)
        clouds = [torch.from_numpy(synth.scan_host(sp, f)).cuda() for f in range(n_frames)]
        for _ in range(3):
            drv.step([fe.wire_view(c) for c in clouds], keep=clouds)
        counts, offsets = drv.fetch()
        want = _oracle_counts(n_frames, oracle)
        assert np.array_equal(counts, want)
        assert np.array_equal(offsets, sharding.global_offsets(want).astype(np.uint64))
        drv.close()


@pytest.mark.gpu
def test_abi_shard_local_group_two_gpus(oracle):
    """Both ranks in this process (lfx_shard_create_local): 5 frames over 2 GPUs, several batches back to back."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, synth

    sp = synth.spec("vlp16")
    n_frames = 5
    fes = [FeatureExtraction(HyperParameters(), device=g) for g in range(2)]
    shards = sharding.local_group(fes, n_frames)
    assert shards[0].info()["nccl_ranks"] == 2
    assert shards[0].info()["exchange"].startswith("peer stores")
    clouds = [[torch.from_numpy(synth.scan_host(sp, f)).to(f"cuda:{g}") for f in range(s.lo, s.hi)] for g, s in enumerate(shards)]
    for _ in range(4):
        for g, s in enumerate(shards):
            s.step([fes[g].wire_view(c) for c in clouds[g]], keep=clouds[g])
    want = _oracle_counts(n_frames, oracle)
    for s in shards:
        s.finish()
    for s in shards:
        counts, offsets = s.fetch()
        assert np.array_equal(counts, want), s.rank
        assert np.array_equal(offsets, sharding.global_offsets(want).astype(np.uint64))
    for s in shards:
        s.close()
    for fe in fes:
        fe.close()


def _abi_worker(rank, world, out_dir, id_path, exchange):
    import time

    if exchange == "nccl":
        os.environ["LFX_SHARD_EXCHANGE"] = "nccl"
    import torch

    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, synth

    torch.cuda.set_device(rank)
    # the unique id travels through a file: the C-ABI driver needs no torch.distributed
    if rank == 0:
        uid = sharding.AbiShard.unique_id()
        with open(id_path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(id_path + ".tmp", id_path)
    else:
        for _ in range(600):
            if os.path.exists(id_path):
                break
            time.sleep(0.05)
        uid = open(id_path, "rb").read()
    n_frames = 5
    sp = synth.spec("vlp16")
    fe = FeatureExtraction(HyperParameters(), device=rank)
    drv = sharding.AbiShard(fe, n_frames, rank, world, unique_id=uid)
    clouds = [torch.from_numpy(synth.scan_host(sp, f)).to(f"cuda:{rank}") for f in range(drv.lo, drv.hi)]
    for _ in range(4):
        drv.step([fe.wire_view(c) for c in clouds], keep=clouds)
    counts, offsets = drv.fetch()
    np.save(os.path.join(out_dir, f"abi_counts_{rank}.npy"), counts)
    np.save(os.path.join(out_dir, f"abi_info_{rank}.npy"), np.array([drv.info()["nccl_ranks"], int(drv.info()["exchange"].startswith("peer"))]))
    drv.close()
    fe.close()


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_abi_shard_one_process_per_gpu(tmp_path, oracle, exchange):
    """One process per GPU (lfx_shard_create with a unique id): counts pushed into buffers mapped through CUDA IPC, or
    (LFX_SHARD_EXCHANGE=nccl) gathered by ncclAllGather."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_abi_worker, args=(2, str(tmp_path), str(tmp_path / "nccl_id.bin"), exchange), nprocs=2, join=True)
    want = _oracle_counts(5, oracle)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"abi_counts_{r}.npy"), want)
        info = np.load(tmp_path / f"abi_info_{r}.npy")
        assert info[0] == 2 and info[1] == (1 if exchange == "p2p" else 0)
