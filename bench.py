#!/usr/bin/env python
"""bench.py — scan-points/sec of the extraction hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU)
    python bench.py --impl reference --steps K --warmup W     # the reference's own CPU code, host cores

A step = one pass of the whole hot path (ingest, order, curvature, masks, selection, packing, and at
N>1 the NCCL all-gather of per-scan counts) over one batch of synthetic scans.
Default workload (weak scaling): BASELINE.json configs[3], OS1-128-shaped scans, 1250 scans per GPU
(= the 10k-scan batch sharded across 8 B200). `value` is measured with inputs resident in HBM;
`e2e` is the same metric through the public C ABI with pinned HOST buffers (H2D of the scans and D2H
of counts + feature clouds inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "scan_points_per_sec"
UNIT = "points/s"
WORKLOADS = {  # name -> (sensor, scans per GPU)
    "os128x1250": ("os128", 1250),   # BASELINE configs[3] sharded over 8 (10k scans / 8)
    "hdl32x1000": ("hdl32", 1000),   # BASELINE configs[1]
    "hdl64x256": ("hdl64", 256),     # BASELINE configs[2] (host-generated: has drop-outs)
    "vlp16x6250": ("vlp16", 6250),   # BASELINE configs[4] sharded over 8
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """SM clock / throttle-reason sampler running DURING the timed region: an NVML polling thread (5 ms
    period, so that even a 40 ms timed region yields several samples under load); falls back to an
    `nvidia-smi -lms` child process when NVML is unusable."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"sw_power_cap": 0x4, "hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
            "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index: int, uuid: str | None = None, period_s: float = 0.005):
        self.gpu = gpu_index
        self.uuid = uuid
        self.period = period_s
        self.nvml = None
        self.handle = None
        self.thread = None
        self.stop_flag = threading.Event()
        self.sm, self.reason_bits, self.mx = [], 0, None
        self.proc = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nvml = pynvml
            if uuid:
                try:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid if isinstance(uuid, bytes) else uuid.encode())
                except Exception:
                    self.handle = None
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self._poll()   # first call pays NVML's lazy initialisation outside the timed region
            self.sm, self.reason_bits = [], 0
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            self.reason_bits |= int(get(self.handle))
        except Exception:
            pass

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                self._poll()
            except Exception:
                pass
            self.stop_flag.wait(self.period)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
            return
        try:
            self.path = tempfile.mktemp(prefix="lfx_clocks_", suffix=".csv")
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if not self.sm:
                return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": ["unsampled"]}
            reasons = sorted(k for k, b in self.BITS.items() if self.reason_bits & b)
            return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(self.sm), "sm_min_mhz": float(min(self.sm)), "source": "nvml, 5 ms period, timed region only"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [t.strip() for t in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    sm.append(float(p[1]))
                    mx.append(float(p[2]))
                except ValueError:
                    continue
                for k, nm in enumerate(names):
                    if p[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 20"}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(sensor: str, n_scans: int, threads: int, first_frame: int = 0, repeats: int = 1, loops: int = 0):
    """Time the reference's CPU extraction (oracle/_ref if built, else the C port) on `n_scans` synthetic
    scans, frames round-robin over `threads` host threads. Returns (points_per_sec, kind, n_points, secs).
    loops > 0: total time of `loops` passes over the sample instead of the best of `repeats`."""
    from concurrent.futures import ThreadPoolExecutor

    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob

    sp = synth.spec(sensor)
    clouds = [synth.scan_host(sp, first_frame + f) for f in range(n_scans)]
    n_points = int(sum(len(c) for c in clouds))
    prm = ob.default_params()
    if ob.Reference.available("stable"):
        ref = ob.Reference("stable")
        kind = "reference"

        def run_all():
            with ThreadPoolExecutor(max_workers=threads) as ex:
                list(ex.map(lambda c: ref.extract_scan(c, prm), clouds))
    else:
        orc = ob.Oracle()
        kind = "port"

        def run_all():
            orc.extract_batch_counts(clouds, prm, threads)
    if loops > 0:
        run_all()   # warm-up pass (page faults, thread pool)
        t0 = time.perf_counter()
        for _ in range(loops):
            run_all()
        dt = time.perf_counter() - t0
        return n_points * loops / dt, kind, n_points * loops, dt
    best = None
    for _ in range(max(repeats, 1)):
        t0 = time.perf_counter()
        run_all()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return n_points / best, kind, n_points, best


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sensor, _ = WORKLOADS[args.workload]
    cores = host_cores()
    n_sample = max(2 * cores, 8)
    if sensor in ("vlp16",):
        n_sample *= 4
    for _ in range(args.warmup):
        cpu_reference_run(sensor, min(n_sample, cores), cores)
    times, pts = [], 0
    kind = "port"
    for k in range(args.steps):
        v, kind, n_points, secs = cpu_reference_run(sensor, n_sample, cores, first_frame=k * n_sample)
        times.append(secs)
        pts += n_points
    total = float(sum(times))
    value = pts / total
    sample = f"{n_sample} synthetic {sensor} scans per step, frames round-robin over {cores} host threads, extraction only"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "sensor": sensor, "params": "compiled defaults (hyper_parameter.hpp:35-43)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="os128x1250", choices=sorted(WORKLOADS))
    ap.add_argument("--scans", type=int, default=0, help="override scans per GPU")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, sharding, synth
    from lidar_feature_extraction_b200 import _native as N

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    sensor, scans_per_gpu = WORKLOADS[args.workload]
    if args.scans:
        scans_per_gpu = args.scans
    sp = synth.spec(sensor)
    per_scan = sp.n_rings * sp.n_cols
    n_frames = world * scans_per_gpu    # weak scaling: the per-GPU shard is fixed
    first_frame, last_frame = sharding.shard_range(n_frames, rank, world)  # rank g owns [g*F/G, (g+1)*F/G)
    assert last_frame - first_frame == scans_per_gpu

    # one explicit stream for everything that is timed: the library launches on it, NCCL (torch's current stream)
    # runs on it and the timing events are recorded on it. (torch's default stream has handle 0, which the C ABI
    # reads as "create your own stream": events on the default stream would then not bracket the kernels.)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    fe = FeatureExtraction(HyperParameters(), device=local_rank, stream=stream.cuda_stream,
                           max_rings=max(128, sp.n_rings))
    lib = N.lib()

    # ---- inputs resident in HBM (generated on device; host-generated when the sensor has drop-outs)
    if sp.dropout_prob > 0:
        clouds = [synth.scan_host(sp, first_frame + f) for f in range(scans_per_gpu)]
        sizes = [len(c) for c in clouds]
        d_in = torch.from_numpy(np.concatenate(clouds, axis=0)).to(dev)
    else:
        sizes = [per_scan] * scans_per_gpu
        d_in = torch.empty((scans_per_gpu * per_scan, 32), dtype=torch.uint8, device=dev)
        rc = lib.lfx_synth_batch_device(fe.handle, C.byref(sp), first_frame, scans_per_gpu, d_in.data_ptr())
        assert rc == 0, rc
    torch.cuda.synchronize()
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n_points = int(offs[-1])
    dev_views = FeatureExtraction.view_array(
        [FeatureExtraction.wire_view((d_in.data_ptr() + int(offs[s]) * 32, sizes[s])) for s in range(scans_per_gpu)])
    # the path's only exchange: all-gather of per-scan (n_edge, n_surface) over NVLink on a side stream (it overlaps
    # the next batch); global offsets follow from a local exclusive scan. join() puts it back on the timed stream.
    gather_mode = os.environ.get("LFX_BENCH_GATHER", "sync")   # diagnosis only: sync (default) | overlap | none
    sharded = sharding.ShardedExtraction(fe, n_frames, dev, overlap=(gather_mode == "overlap"))

    def step_device():
        if gather_mode == "none":
            return fe.extract_views(dev_views, keep=d_in)
        return sharded.step(dev_views, keep=d_in)

    for _ in range(max(args.warmup, 3)):
        step_device()
    sharded.join()
    torch.cuda.synchronize()
    launches0 = fe.kernel_launches
    # NVML is initialised (first query included) BEFORE the barrier: it takes milliseconds and a different time on
    # every rank, and whatever start skew the ranks have after the barrier is paid by all of them at the first exchange
    try:
        uuid = "GPU-" + str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        uuid = None
    sampler = ClockSampler(local_rank, uuid)
    no_clocks = os.environ.get("LFX_BENCH_NO_CLOCKS") == "1"    # diagnosis only
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if not no_clocks:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    sharded.join()
    e1.record(stream)
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if not no_clocks else None
    ms_total = e0.elapsed_time(e1)
    # the events sit on the stream every kernel of the step is launched on, so the host clock around the same
    # region (synchronize included) can only be slightly larger; anything else means the events missed work
    if not (0.8 * wall_ms - 1.0 <= ms_total <= wall_ms * 1.001 + 0.05):
        print(f"warning: rank {rank}: event time {ms_total:.3f} ms vs host clock {wall_ms:.3f} ms", file=sys.stderr)
    launches = fe.kernel_launches - launches0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = n_points * world / (ms_per_step * 1e-3)

    # ---- the exchange alone (N > 1): device time from the end of the batch to the end of the all-gather
    exchange = None
    if world > 1 and gather_mode == "sync":
        sharded.time_gather = True
        for _ in range(6):
            step_device()
        g = sharded.gather_ms()[1:]
        sharded.time_gather = False
        t = torch.tensor([float(np.mean(g)), float(np.max(g))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exchange = {"all_gather_ms_mean": float(t[0].item()), "all_gather_ms_max": float(t[1].item()),
                    "bytes_per_rank": 8 * scans_per_gpu, "how": "CUDA events around the collective on the extraction stream, max over ranks"}

    # ---- dominant kernel (k_extract_sectors) timed live with CUDA events on the launching stream
    counts, offsets = np.zeros((scans_per_gpu, 2), np.uint32), np.zeros((scans_per_gpu + 1, 2), np.uint32)
    lib.lfx_fetch_counts(fe.handle, counts.ctypes.data, offsets.ctypes.data)
    n_feat = int(offsets[-1, 0]) + int(offsets[-1, 1])
    alg_bytes = 32 * n_points + n_points + 16 * n_feat + 8 * scans_per_gpu  # SURVEY.md 8(d) / BASELINE.md 4
    # (a) as in the timed region: the batch runs as its CUDA graph, steps back to back, the stage events are
    #     event-record nodes of that graph; (b) eager launches with a synchronize between steps, for comparison
    def stage_loop(mode, sync):
        fe.set_stage_timing(mode)
        rows = []
        for _ in range(max(3, min(args.steps, 10)) + 1):
            fe.extract_views(dev_views, keep=d_in)
            if sync:
                fe.synchronize()
            rows.append(fe.last_stage_ms())
        fe.set_stage_timing(False)
        return np.array(rows[1:])
    stage = stage_loop(2, False)
    stage_eager = stage_loop(1, True)
    # k_extract_sectors: launches on regular scans (stage 1) + launches on bucketed rings (stage 3); on a given
    # workload one instantiation holds (nearly) the whole batch
    ring_ms = float((stage[:, 1] + stage[:, 3]).mean())
    peaks, peak_kind = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    achieved = alg_bytes / (ring_ms * 1e-3) / 1e9
    traffic = None
    try:
        # dram__bytes_read + dram__bytes_write of one launch of the sector kernel (ncu --set full), per point
        with open(os.path.join(ROOT, "profiles", "sector_kernel_traffic.json")) as f:
            tj = json.load(f)
        traffic = float(tj["dram_bytes_per_point"][sensor]) * n_points
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_extract_sectors", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_kind": f"of {peak_kind}",
                "kernel_ms": ring_ms, "algorithmic_bytes": alg_bytes,
                "stage_ms": {"probe": float(stage[:, 0].mean()), "sectors": float(stage[:, 1].mean()),
                             "bucketing": float(stage[:, 2].mean()), "sectors_indexed": float(stage[:, 3].mean()),
                             "rings": float(stage[:, 4].mean()), "pack": float(stage[:, 5].mean())},
                "stage_ms_how": "event-record nodes inside the batch's CUDA graph, steps back to back as in the timed region",
                "kernel_ms_eager": float((stage_eager[:, 1] + stage_eager[:, 3]).mean()),
                "paths": fe.batch_stats(),
                "pipeline_frac": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak}

    # ---- e2e: same metric through the public C ABI with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        nbytes = n_points * 32
        hptr = lib.lfx_host_alloc(nbytes)
        assert hptr, "pinned allocation failed"
        lib.lfx_memcpy_d2h(fe.handle, hptr, d_in.data_ptr(), nbytes)
        host_views = [FeatureExtraction.wire_view((hptr + int(offs[s]) * 32, sizes[s])) for s in range(scans_per_gpu)]
        for v in host_views:
            v.memory = N.LFX_MEM_HOST
        host_views = FeatureExtraction.view_array(host_views)
        cap = n_points
        h_edge = lib.lfx_host_alloc(16 * max(n_feat, 1) * 2)
        h_surf = lib.lfx_host_alloc(16 * max(n_feat, 1) * 2)
        h_counts = np.zeros((scans_per_gpu, 2), np.uint32)
        h_offsets = np.zeros((scans_per_gpu + 1, 2), np.uint32)

        def step_e2e():
            res = fe.extract_views(host_views)
            if world > 1:
                sharding.gather_counts(sharding.device_counts_tensor(res, dev), n_frames)
            rc = lib.lfx_fetch_counts(fe.handle, h_counts.ctypes.data, h_offsets.ctypes.data)
            rc |= lib.lfx_fetch_features(fe.handle, h_edge, 2 * max(n_feat, 1), h_surf, 2 * max(n_feat, 1))
            assert rc == 0, rc

        k_e2e = args.e2e_steps or max(2, min(args.steps, 5))
        for _ in range(2):
            step_e2e()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(k_e2e):
            step_e2e()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms_e2e = max(e0.elapsed_time(e1), wall)  # host-side copies and syncs are part of the call
        if world > 1:
            dist.barrier()
            t = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item())
        d2h = int(h_counts.nbytes + h_offsets.nbytes + 16 * (int(h_offsets[-1, 0]) + int(h_offsets[-1, 1])))
        e2e = {"value": n_points * world / (ms_e2e / k_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": d2h, "steps": k_e2e, "ms_per_step": ms_e2e / k_e2e}
        lib.lfx_host_free(hptr)
        lib.lfx_host_free(h_edge)
        lib.lfx_host_free(h_surf)

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only; bounded sample)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        # bounded sample: ~10 s of CPU work (the reference runs ~2 Mpts/s per thread)
        n_sample = max(int(4.0e6 * cores / per_scan), cores)
        loops = 5
        v, kind, pts, secs = cpu_reference_run(sensor, n_sample, cores, loops=loops)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_sample} synthetic {sensor} scans x {loops} passes ({pts} points, {secs:.2f} s), frames round-robin over {cores} host threads, extraction only"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "sensor": sensor, "rings": sp.n_rings, "cols": sp.n_cols,
                       "scans_per_gpu": scans_per_gpu, "points_per_gpu": n_points, "params": "compiled defaults (hyper_parameter.hpp:35-43)",
                       "sharding": "frames by index, no data-path collective; NCCL all-gather of per-scan counts" if world > 1 else "single GPU",
                       "l2": f"inputs {n_points * 32 / 1e9:.2f} GB per GPU, larger than the 126 MB L2 (no flush needed)",
                       "selected_fraction": n_feat / max(n_points, 1)},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "exchange": exchange,
        }
        print(json.dumps(line), flush=True)
    fe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
