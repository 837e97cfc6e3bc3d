"""Where does the per-step cost of a small all-gather between long kernels come from? (torchrun, 2 ranks)"""
import os
import torch
import torch.distributed as dist

rank = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
w = dist.get_world_size()
send = torch.zeros((1250, 2), dtype=torch.int32, device=dev)
recv = torch.empty((w * 1250, 2), dtype=torch.int32, device=dev)
a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def run(name, busy, gather, n=20):
    for _ in range(3):
        busy(); gather()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        busy(); gather()
    e1.record()
    torch.cuda.synchronize()
    if rank == 0:
        print(f"{name}: {e0.elapsed_time(e1) / n:.3f} ms per step", flush=True)


def busy_mem():
    for _ in range(4):
        a.mul_(1.0001)


def busy_sleep():
    torch.cuda._sleep(int(1.25e-3 * 1.9e9))


def g_none():
    pass


def g_sync():
    dist.all_gather_into_tensor(recv, send)


works = []


def g_async():
    works.append(dist.all_gather_into_tensor(recv, send, async_op=True))


def g_allreduce():
    dist.all_reduce(send)


run("mem busy, no gather", busy_mem, g_none)
run("mem busy, all_gather", busy_mem, g_sync)
run("mem busy, all_gather async_op (no wait)", busy_mem, g_async)
run("mem busy, all_reduce", busy_mem, g_allreduce)
run("sleep busy, no gather", busy_sleep, g_none)
run("sleep busy, all_gather", busy_sleep, g_sync)
for wk in works:
    wk.wait()
dist.destroy_process_group()
