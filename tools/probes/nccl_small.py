"""Latency of a 10 KB all_gather_into_tensor on this box (torchrun, 2+ ranks)."""
import os
import time
import torch
import torch.distributed as dist

rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
w = dist.get_world_size()
send = torch.zeros((1250, 2), dtype=torch.int32, device=dev)
recv = torch.empty((w * 1250, 2), dtype=torch.int32, device=dev)
for _ in range(20):
    dist.all_gather_into_tensor(recv, send)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(200):
    dist.all_gather_into_tensor(recv, send)
e1.record()
torch.cuda.synchronize()
if rank == 0:
    print("all_gather 10KB: %.1f us per call (device time, back to back)" % (e0.elapsed_time(e1) / 200 * 1e3))
# interleaved with a 4 ms busy kernel, like the extraction
a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    for _ in range(4):
        a.mul_(1.0001)
    dist.all_gather_into_tensor(recv, send)
e1.record()
torch.cuda.synchronize()
t_with = e0.elapsed_time(e1) / 20
e0.record()
for _ in range(20):
    for _ in range(4):
        a.mul_(1.0001)
e1.record()
torch.cuda.synchronize()
t_without = e0.elapsed_time(e1) / 20
if rank == 0:
    print("busy step %.3f ms, with gather %.3f ms" % (t_without, t_with))
dist.destroy_process_group()
