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


PARAMS_NOTE = "compiled defaults (hyper_parameter.hpp:35-43)"


def workload_config(workload: str, sensor: str, rings: int, cols: int, scans_per_gpu: int, points_per_gpu: int, world: int) -> dict:
    """The `config` object of the JSON line: built by this one function for BOTH arms, so that the driver's
    same-config check compares like with like."""
    return {"workload": workload, "sensor": sensor, "rings": rings, "cols": cols, "scans_per_gpu": scans_per_gpu,
            "points_per_gpu": points_per_gpu, "params": PARAMS_NOTE,
            "sharding": "frames by index, no data-path collective; per-scan counts exchanged between the ranks (lfx_shard_*)" if world > 1 else "single GPU",
            "l2": f"inputs {points_per_gpu * 32 / 1e9:.2f} GB per GPU, larger than the 126 MB L2 (no flush needed)"}


def git_blob_sha1(path: str) -> str:
    """`git hash-object` of a file (what keys profiles/sector_kernel_traffic.json to the kernel source)."""
    import hashlib

    data = open(path, "rb").read()
    return hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()


def cpu_reference_pass(clouds, threads: int):
    """One pass of the reference's CPU extraction (oracle/_ref if built, else the C port) over `clouds`, scans
    round-robin over `threads` host threads. Returns (seconds, kind, per-scan results or None for the port)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import binding as ob

    prm = ob.default_params()
    if ob.Reference.available("stable"):
        ref = ob.Reference("stable")
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as ex:
            results = list(ex.map(lambda c: ref.extract_scan(c, prm), clouds))
        return time.perf_counter() - t0, "reference", results
    orc = ob.Oracle()
    t0 = time.perf_counter()
    orc.extract_batch_counts(clouds, prm, threads)
    return time.perf_counter() - t0, "port", None


def cpu_reference_run(clouds, threads: int, loops: int = 1):
    """`loops` timed passes after one warm-up pass (page faults, thread pool). Returns (points_per_sec, kind,
    points processed, seconds, results of the warm-up pass)."""
    n_points = int(sum(len(c) for c in clouds))
    _, kind, results = cpu_reference_pass(clouds, threads)
    total = 0.0
    for _ in range(max(loops, 1)):
        dt, kind, _ = cpu_reference_pass(clouds, threads)
        total += dt
    return n_points * max(loops, 1) / total, kind, n_points * max(loops, 1), total, results


def parity_check(out, clouds, results, kind: str) -> dict:
    """The GPU batch `out` (BatchOutput) against the reference's own results on the first len(clouds) scans of the
    same batch: ring tables, every label byte, per-scan counts, and the edge / surface clouds byte for byte."""
    from lidar_feature_extraction_b200 import _native as N
    from lidar_feature_extraction_b200 import synth

    bad_scans, n_pts = 0, 0
    hist_gpu, hist_ref = np.zeros(256, np.int64), np.zeros(256, np.int64)
    first_bad = None
    for s, (cloud, ref) in enumerate(zip(clouds, results)):
        why = None
        base = int(out.point_base[s])
        rings = out.rings[s]
        kept = [r for r in range(len(rings)) if rings[r]["count"] > 0 and rings[r]["status"] != N.LFX_RING_SPARSE]
        if kept != [int(r) for r in ref.ring_ids]:
            why = "ring ids"
        pos = 0
        for k, r in enumerate(kept):
            if why:
                break
            cnt, off = int(rings[r]["count"]), int(rings[r]["offset"])
            if cnt != int(ref.ring_sizes[k]):
                why = f"ring {r} size"
                break
            g, w = out.labels[base + off: base + off + cnt], ref.labels[pos: pos + cnt]
            hist_gpu += np.bincount(g, minlength=256)
            hist_ref += np.bincount(w, minlength=256)
            if not np.array_equal(g, w):
                why = f"ring {r} labels"
            pos += cnt
        n_pts += pos
        if why is None:
            x, y, z, _, _ = synth.fields(cloud)
            src = ref.sorted_src
            if int(out.counts[s, 0]) != len(ref.edge_idx) or int(out.counts[s, 1]) != len(ref.surface_idx):
                why = "counts"
            else:
                for got, idx, name in ((out.scan_edges(s), ref.edge_idx, "edge"), (out.scan_surfaces(s), ref.surface_idx, "surface")):
                    want = np.stack([x[src[idx]], y[src[idx]], z[src[idx]], np.ones(len(idx), np.float32)], axis=1)
                    if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
                        why = f"{name} cloud"
        if why is not None:
            bad_scans += 1
            first_bad = first_bad or f"scan {s}: {why}"
    return {"scans": len(clouds), "points": n_pts, "mismatches": bad_scans, "against": kind, "first_mismatch": first_bad,
            "checked": "ring tables, label bytes, per-scan counts, edge and surface clouds (bitwise), in canonical order",
            "label_histogram": {str(k): int(v) for k, v in enumerate(hist_gpu[:8])},
            "label_histogram_equal": bool(np.array_equal(hist_gpu, hist_ref))}


def run_reference_arm(args):
    """The reference's own CPU extraction on the box's host cores; scans come from libsynth.so (host-only), so the
    CUDA product library is never mapped by this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lidar_feature_extraction_b200 import synth

    sensor, scans_per_gpu = WORKLOADS[args.workload]
    if args.scans:
        scans_per_gpu = args.scans
    world = int(os.environ.get("WORLD_SIZE", "1"))
    sp = synth.spec(sensor)
    cores = host_cores()
    n_sample = min(max(2 * cores, 8) * (4 if sensor == "vlp16" else 1), scans_per_gpu)
    if sp.dropout_prob > 0:   # ragged scans: the batch's point count is a property of the data
        points_per_gpu = int(sum(len(synth.scan_host(sp, f)) for f in range(scans_per_gpu)))
    else:
        points_per_gpu = scans_per_gpu * sp.n_rings * sp.n_cols
    kind = "port"
    for _ in range(args.warmup):
        cpu_reference_pass([synth.scan_host(sp, f) for f in range(min(n_sample, cores))], cores)
    total, pts = 0.0, 0
    for k in range(args.steps):   # step k: the k-th block of n_sample frames of rank 0's shard (wrapping around)
        clouds = [synth.scan_host(sp, (k * n_sample + f) % scans_per_gpu) for f in range(n_sample)]
        secs, kind, _ = cpu_reference_pass(clouds, cores)
        total += secs
        pts += int(sum(len(c) for c in clouds))
    value = pts / total
    sample = (f"{n_sample} of the workload's {scans_per_gpu} synthetic {sensor} scans per step (consecutive frames of rank 0's shard), "
              f"scans round-robin over {cores} host threads, extraction only (generation outside the timed region)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, sensor, sp.n_rings, sp.n_cols, scans_per_gpu, points_per_gpu, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


STAGES = ("probe", "sectors", "bucketing", "sectors_indexed", "rings", "pack")


def make_inputs(fe, lib, sp, first_frame: int, scans: int, dev):
    """`scans` synthetic scans resident in HBM: generated on the device, or on the host when the sensor has drop-outs
    (ragged scans). Returns (uint8 tensor [points, 32], per-scan sizes, host clouds or None)."""
    import torch

    from lidar_feature_extraction_b200 import synth

    per_scan = sp.n_rings * sp.n_cols
    if sp.dropout_prob > 0:
        clouds = [synth.scan_host(sp, first_frame + f) for f in range(scans)]
        sizes = [len(c) for c in clouds]
        return torch.from_numpy(np.concatenate(clouds, axis=0)).to(dev), sizes, clouds
    d_in = torch.empty((scans * per_scan, 32), dtype=torch.uint8, device=dev)
    rc = lib.lfx_synth_batch_device(fe.handle, C.byref(sp), first_frame, scans, d_in.data_ptr())
    assert rc == 0, rc
    return d_in, [per_scan] * scans, None


def stage_times(fe, views, keep, mode: int, sync: bool, iters: int):
    fe.set_stage_timing(mode)
    rows = []
    for _ in range(iters + 1):
        fe.extract_views(views, keep=keep)
        if sync:
            fe.synchronize()
        rows.append(fe.last_stage_ms())
    fe.set_stage_timing(False)
    return np.array(rows[1:])


def measure_workload(name: str, local_rank: int, dev, stream, steps: int, peak: float, rank: int = 0, world: int = 1,
                     parity_scans: int = -1) -> dict:
    """One of the other BASELINE.json configs, device-resident: same timing rules as the headline (CUDA events on the
    launching stream, >= 3 warm-up steps, inputs larger than L2), fewer extras. world > 1: every rank extracts its own
    shard of world x scans frames and the per-scan counts are exchanged each step (lfx_shard_*); time = max over ranks,
    value = all ranks' points / that time; the stage times and paths are rank 0's."""
    import torch

    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, sharding, synth
    from lidar_feature_extraction_b200 import _native as N

    sensor, scans = WORKLOADS[name]
    sp = synth.spec(sensor)
    lib = N.lib()
    fe = FeatureExtraction(HyperParameters(), device=local_rank, stream=stream.cuda_stream, max_rings=max(128, sp.n_rings))
    d_in, sizes, _ = make_inputs(fe, lib, sp, rank * scans, scans, dev)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n_points = int(offs[-1])
    views = FeatureExtraction.view_array(
        [FeatureExtraction.wire_view((d_in.data_ptr() + int(offs[s]) * 32, sizes[s])) for s in range(scans)])
    shard = sharding.AbiShard(fe, world * scans, rank, world) if world > 1 else None

    def step():
        if shard is not None:
            shard.step(views, keep=d_in)
        else:
            fe.extract_views(views, keep=d_in)

    for _ in range(3):
        step()
    if shard is not None:
        shard.join()
    torch.cuda.synchronize()
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    if shard is not None:
        shard.join()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    counts, offsets = np.zeros((scans, 2), np.uint32), np.zeros((scans + 1, 2), np.uint32)
    lib.lfx_fetch_counts(fe.handle, counts.ctypes.data, offsets.ctypes.data)
    n_feat = int(offsets[-1, 0]) + int(offsets[-1, 1])
    alg = 32 * n_points + n_points + 16 * n_feat + 8 * scans
    if world > 1:
        t = torch.tensor([ms, float(n_points), float(alg)], dtype=torch.float64, device=dev)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms = float(tmax[0].item())
        n_points_all, alg_all = int(t[1].item()), int(t[2].item())
    else:
        n_points_all, alg_all = n_points, alg
    stage = stage_times(fe, views, d_in, 2, False, 3)
    kernel_ms = float((stage[:, 1] + stage[:, 3]).mean())
    out = {"ms_per_step": ms, "value": n_points_all / (ms * 1e-3), "unit": UNIT, "steps": steps, "points_per_step": n_points_all,
           "scans": scans * world, "n_gpus": world, "input_gb": n_points_all * 32 / 1e9, "algorithmic_bytes": alg_all,
           "pipeline_frac": alg_all / (ms * 1e-3) / 1e9 / (peak * world), "kernel_frac": alg / (kernel_ms * 1e-3) / 1e9 / peak,
           "stage_ms": {k: float(stage[:, i].mean()) for i, k in enumerate(STAGES)}, "paths": fe.batch_stats(),
           "selected_fraction": n_feat / max(n_points, 1)}
    # parity of THIS batch against the reference itself on its first scans (rank 0; CPU work bounded to ~4 M points)
    if rank == 0 and parity_scans != 0:
        n_par = max(2, min(scans, int(4.0e6 // max(n_points // scans, 1)))) if parity_scans < 0 else min(parity_scans, scans)
        sample = d_in[: int(offs[n_par])].cpu().numpy()
        clouds = [sample[int(offs[s]): int(offs[s + 1])] for s in range(n_par)]
        _, kind, results = cpu_reference_pass(clouds, host_cores())
        if results is not None:
            fe.extract_views(views, keep=d_in)
            out["parity"] = parity_check(fe.fetch(fetch_points=True), clouds, results, kind)
            out["parity"].pop("label_histogram", None)
        del sample, clouds
    if shard is not None:
        torch.cuda.synchronize()
        dist.barrier()              # no peer may still be pushing into this rank's buffers when they are released
        shard.close()
    fe.close()
    del d_in
    torch.cuda.empty_cache()
    return out


def measure_chain(local_rank: int, dev, stream, steps: int, peak: float, scans: int = 1250) -> dict:
    """/points_raw -> k_convert -> extraction without leaving the device: Ouster-like 48-byte driver clouds (3 % of the
    returns zeroed, so the converted clouds are ragged and take the bucketing + indexed sector path)."""
    import torch

    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters, PointCloud2, PointField, PointTypeConverter, synth
    from lidar_feature_extraction_b200 import _native as N

    sp = synth.spec("os128")
    per = sp.n_rings * sp.n_cols
    lib = N.lib()
    fe = FeatureExtraction(HyperParameters(), device=local_rank, stream=stream.cuda_stream, max_rings=128)
    wire, _, _ = make_inputs(fe, lib, sp, 0, scans, dev)
    torch.cuda.synchronize()
    # Ouster driver layout (test_convert.py:177-187): x,y,z f32 @0, intensity f32 @16, t u32 @20, reflectivity u16 @24,
    # ring u8 @26, noise u16 @28, range u32 @32, point_step 48
    raw = torch.zeros((scans * per, 48), dtype=torch.uint8, device=dev)
    raw[:, 0:12] = wire[:, 0:12]
    raw[:, 16:20] = wire[:, 16:20]
    raw[:, 26] = wire[:, 20]
    g = torch.Generator(device=dev)
    g.manual_seed(0xC0FFEE)
    dead = torch.rand(scans * per, device=dev, generator=g) < 0.03
    raw[dead] = 0
    del wire, dead
    fields = [PointField("x", 0, 7), PointField("y", 4, 7), PointField("z", 8, 7), PointField("intensity", 16, 7), PointField("t", 20, 6),
              PointField("reflectivity", 24, 4), PointField("ring", 26, 2), PointField("noise", 28, 4), PointField("range", 32, 6)]
    raw3 = raw.view(scans, per, 48)
    msgs = [PointCloud2(data=raw3[s], point_step=48, fields=fields) for s in range(scans)]
    conv = PointTypeConverter(fe)
    raw_batch = conv.marshal(msgs)   # the C array of raw clouds, built once (the clouds are the same every step)

    def chain_step():
        conv.convert_batch(raw_batch)
        fe.extract_views(conv.views())

    for _ in range(3):
        chain_step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        chain_step()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    ms = max(e0.elapsed_time(e1) / steps, wall)   # the converter's call is synchronous (it returns the kept counts)
    cms = C.c_float()
    lib.lfx_last_convert_ms(fe.handle, C.byref(cms))
    kept = sum(int(conv.kept(s)) for s in range(scans))
    counts, offsets = np.zeros((scans, 2), np.uint32), np.zeros((scans + 1, 2), np.uint32)
    lib.lfx_fetch_counts(fe.handle, counts.ctypes.data, offsets.ctypes.data)
    n_feat = int(offsets[-1, 0]) + int(offsets[-1, 1])
    n = scans * per
    alg = 48 * n + 32 * kept + 32 * kept + kept + 16 * n_feat + 8 * scans   # converter (raw in, 32 B out) + extraction
    fe.set_stage_timing(True)
    chain_step()
    fe.synchronize()
    st = fe.last_stage_ms()
    fe.set_stage_timing(False)
    out = {"ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "raw points/s", "steps": steps, "points_per_step": n, "scans": scans,
           "kept_fraction": kept / n, "algorithmic_bytes": alg, "pipeline_frac": alg / (ms * 1e-3) / 1e9 / peak,
           "convert_kernel_ms": float(cms.value), "stage_ms": {k: float(st[i]) for i, k in enumerate(STAGES)}, "paths": fe.batch_stats()}
    # parity of THIS chain batch: the converted bytes of the first clouds against the converter's CPU restatement
    # (oracle/convert_oracle.py, pinned to the reference's convert.py), and the features extracted from them against
    # the reference's extraction of those converted clouds
    try:
        from oracle import convert_oracle as co

        n_par = max(2, min(scans, int(4.0e6 // per)))
        host_conv, conv_bad = [], 0
        ofields = [co.Field(f.name, f.offset, f.datatype) for f in fields]
        for s in range(n_par):
            got = conv.fetch(s)
            want, _ = co.convert(raw3[s].cpu().numpy(), ofields, 48)
            conv_bad += 0 if np.array_equal(got, want) else 1
            host_conv.append(got)
        _, kind, results = cpu_reference_pass(host_conv, host_cores())
        if results is not None:
            out["parity"] = parity_check(fe.fetch(fetch_points=True), host_conv, results, kind)
            out["parity"].pop("label_histogram", None)
            out["parity"]["converter"] = {"clouds": n_par, "mismatches": conv_bad, "against": "port (oracle/convert_oracle.py, pinned to convert.py's outputs)"}
    except Exception as e:   # the checker must not take the measurement down
        out["parity"] = {"error": repr(e)}
    conv.close()
    fe.close()
    del raw, raw3, msgs
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--no-workloads", action="store_true", help="skip the other BASELINE.json configs and the converter chain")
    ap.add_argument("--e2e-steps", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from lidar_feature_extraction_b200 import FeatureExtraction, PipelinedExtraction, HyperParameters, sharding, synth
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
    d_in, sizes, _ = make_inputs(fe, lib, sp, first_frame, scans_per_gpu, dev)
    torch.cuda.synchronize()
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n_points = int(offs[-1])
    dev_views = FeatureExtraction.view_array(
        [FeatureExtraction.wire_view((d_in.data_ptr() + int(offs[s]) * 32, sizes[s])) for s in range(scans_per_gpu)])
    # the path's only exchange: all-gather of per-scan (n_edge, n_surface) over NVLink on a side stream (it overlaps
    # the next batch); global offsets follow from a local exclusive scan. join() puts it back on the timed stream.
    # The path's only exchange: every rank's per-scan (n_edge, n_surface) reach every rank, global offsets follow from a
    # scan. Default: the library's own driver (lfx_shard_*: NCCL set-up, counts pushed through peer-mapped memory over
    # NVLink by a one-CTA kernel at the tail of the batch). Diagnosis: LFX_BENCH_GATHER=torch (all_gather_into_tensor
    # on the extraction stream) | overlap (the same on a side stream) | none.
    gather_mode = os.environ.get("LFX_BENCH_GATHER", "abi") if world > 1 else "none"
    if gather_mode == "none":
        class _NoExchange:      # a single GPU has nobody to tell
            def join(self):
                pass
        sharded = _NoExchange()
    elif gather_mode == "abi":
        sharded = sharding.AbiShard(fe, n_frames, rank, world)
    else:
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
    if world > 1 and gather_mode == "abi":
        # device time of what the exchange adds to a step on the extraction stream: the scan of the previous exchange
        # (its peers' flags arrived a batch ago) + the push of this one
        evs = []
        for _ in range(6):
            fe.extract_views(dev_views, keep=d_in)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            sharded.exchange()
            b.record(stream)
            evs.append((a, b))
        sharded.join()
        torch.cuda.synchronize()
        g = [a.elapsed_time(b) for a, b in evs][1:]
        t = torch.tensor([float(np.mean(g)), float(np.max(g))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        counts_all, offsets_all = sharded.fetch()
        info = sharded.info()
        exchange = {"ms_mean": float(t[0].item()), "ms_max": float(t[1].item()), "bytes_per_rank": 8 * scans_per_gpu,
                    "mechanism": info["exchange"], "nccl_ranks": info["nccl_ranks"], "frames_in_global_table": int(counts_all.shape[0]),
                    "global_features": [int(offsets_all[-1, 0]), int(offsets_all[-1, 1])],
                    "how": "CUDA events around lfx_shard_exchange on the extraction stream (scan of the previous exchange + push of this one), max over ranks"}
    elif world > 1 and gather_mode == "torch":
        sharded.time_gather = True
        for _ in range(6):
            step_device()
        g = sharded.gather_ms()[1:]
        sharded.time_gather = False
        t = torch.tensor([float(np.mean(g)), float(np.max(g))], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        exchange = {"ms_mean": float(t[0].item()), "ms_max": float(t[1].item()), "mechanism": "torch.distributed all_gather_into_tensor (NCCL)",
                    "bytes_per_rank": 8 * scans_per_gpu, "how": "CUDA events around the collective on the extraction stream, max over ranks"}

    # ---- dominant kernel (k_extract_sectors) timed live with CUDA events on the launching stream
    counts, offsets = np.zeros((scans_per_gpu, 2), np.uint32), np.zeros((scans_per_gpu + 1, 2), np.uint32)
    lib.lfx_fetch_counts(fe.handle, counts.ctypes.data, offsets.ctypes.data)
    n_feat = int(offsets[-1, 0]) + int(offsets[-1, 1])
    alg_bytes = 32 * n_points + n_points + 16 * n_feat + 8 * scans_per_gpu  # SURVEY.md 8(d) / BASELINE.md 4
    # (a) as in the timed region: the batch runs as its CUDA graph, steps back to back, the stage events are
    #     event-record nodes of that graph; (b) eager launches with a synchronize between steps, for comparison
    stage = stage_times(fe, dev_views, d_in, 2, False, max(3, min(args.steps, 10)))
    stage_eager = stage_times(fe, dev_views, d_in, 1, True, max(3, min(args.steps, 10)))
    # k_extract_sectors: launches on regular scans (stage 1) + launches on bucketed rings (stage 3); on a given
    # workload one instantiation holds (nearly) the whole batch
    ring_ms = float((stage[:, 1] + stage[:, 3]).mean())
    peaks, peak_kind = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    achieved = alg_bytes / (ring_ms * 1e-3) / 1e9
    traffic, traffic_note = None, "no ncu capture of this kernel source / sensor under profiles/"
    try:
        # dram__bytes_read + dram__bytes_write of one launch of the sector kernel (ncu --set full), per point. It cannot
        # be measured inside this run (it needs the profiler), so the committed figure is keyed to the git blob of the
        # kernel source it was captured from and is reported only while that source is unchanged
        with open(os.path.join(ROOT, "profiles", "sector_kernel_traffic.json")) as f:
            tj = json.load(f)
        blob = git_blob_sha1(os.path.join(ROOT, "lidar_feature_extraction_b200", "csrc", "lfx_sector.cuh"))
        if tj.get("sector_cuh_blob") == blob and sensor in tj["dram_bytes_per_point"]:
            traffic = float(tj["dram_bytes_per_point"][sensor]) * n_points
            traffic_note = f"ncu capture {tj.get('capture', '?')} of lfx_sector.cuh blob {blob[:12]}"
        elif tj.get("sector_cuh_blob") != blob:
            traffic_note = f"lfx_sector.cuh (blob {blob[:12]}) changed since the capture (blob {str(tj.get('sector_cuh_blob'))[:12]})"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_extract_sectors", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_note, "peak_kind": f"of {peak_kind}",
                "kernel_ms": ring_ms, "algorithmic_bytes": alg_bytes,
                "stage_ms": {k: float(stage[:, i].mean()) for i, k in enumerate(STAGES)},
                "stage_ms_how": "event-record nodes inside the batch's CUDA graph, steps back to back as in the timed region",
                "kernel_ms_eager": float((stage_eager[:, 1] + stage_eager[:, 3]).mean()),
                "paths": fe.batch_stats(), "selected_fraction": n_feat / max(n_points, 1),
                "pipeline_frac": (alg_bytes / (ms_per_step * 1e-3) / 1e9) / peak}

    # ---- e2e: same metric through the public C ABI with pinned HOST buffers
    e2e = None
    if not args.no_e2e:
        nbytes = n_points * 32
        numa = C.c_int(-1)
        hptr = lib.lfx_host_alloc_on(fe.handle, nbytes, C.byref(numa))   # pinned, on the NUMA node of this rank's GPU
        assert hptr, "pinned allocation failed"
        lib.lfx_memcpy_d2h(fe.handle, hptr, d_in.data_ptr(), nbytes)
        host_views = [FeatureExtraction.wire_view((hptr + int(offs[s]) * 32, sizes[s])) for s in range(scans_per_gpu)]
        for v in host_views:
            v.memory = N.LFX_MEM_HOST
        host_views = FeatureExtraction.view_array(host_views)
        cap = n_points
        h_edge = lib.lfx_host_alloc_on(fe.handle, 16 * max(n_feat, 1) * 2, None)
        h_surf = lib.lfx_host_alloc_on(fe.handle, 16 * max(n_feat, 1) * 2, None)
        h_counts = np.zeros((scans_per_gpu, 2), np.uint32)
        h_offsets = np.zeros((scans_per_gpu + 1, 2), np.uint32)

        def step_e2e():
            res = fe.extract_views(host_views)
            if world > 1 and gather_mode == "abi":
                sharded.exchange()
            elif world > 1:
                sharding.gather_counts(sharding.device_counts_tensor(res, dev), n_frames)
            rc = lib.lfx_fetch_counts(fe.handle, h_counts.ctypes.data, h_offsets.ctypes.data)
            rc |= lib.lfx_fetch_features(fe.handle, h_edge, 2 * max(n_feat, 1), h_surf, 2 * max(n_feat, 1))
            assert rc == 0, rc

        k_e2e = args.e2e_steps or max(2, min(args.steps, 5))

        def timed(step_fn, drain_fn=None):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for k in range(k_e2e):
                step_fn(k)
            if drain_fn:
                drain_fn()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) * 1e3     # host clock: copies, syncs and both streams are part of the call
            if world > 1:
                dist.barrier()
                t = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms

        # (a) one handle, call by call: upload, kernels and download of a batch in series
        for _ in range(2):
            step_e2e()
        ms_serial = timed(lambda k: step_e2e())
        d2h = int(h_counts.nbytes + h_offsets.nbytes + 16 * (int(h_offsets[-1, 0]) + int(h_offsets[-1, 1])))

        # (b) the pipelined host API (PipelinedExtraction = lfx::Pipeline): two handles take the batches in turn, batch
        #     k-1's features download while batch k uploads on the other stream. Every batch is uploaded, extracted,
        #     exchanged and downloaded inside the timed region, the last one drained before the clock stops.
        fe2 = FeatureExtraction(HyperParameters(), device=local_rank, max_rings=max(128, sp.n_rings))
        shards = (sharded, sharding.AbiShard(fe2, n_frames, rank, world)) if (world > 1 and gather_mode == "abi") else None
        pipe = PipelinedExtraction(handles=(fe, fe2))

        def pipe_collect():
            pipe.collect(h_edge, 2 * max(n_feat, 1), h_surf, 2 * max(n_feat, 1))

        def pipe_step(k):
            slot = pipe.submit(host_views)
            if shards:
                shards[slot].exchange()
            if pipe.in_flight == 2:
                pipe_collect()

        for k in range(3):
            pipe_step(k)
        pipe_collect()
        ms_e2e = timed(pipe_step, pipe_collect)
        if shards:
            shards[1].join()
            torch.cuda.synchronize()
            dist.barrier()          # no peer may still be pushing into this rank's buffers when they are released
            shards[1].close()
        e2e = {"value": n_points * world / (ms_e2e / k_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": d2h, "steps": k_e2e, "ms_per_step": ms_e2e / k_e2e,
               "api": "PipelinedExtraction (lfx::Pipeline): two handles in turn; all uploads, kernels and downloads of the timed batches inside the timed region",
               "serial": {"value": n_points * world / (ms_serial / k_e2e * 1e-3), "ms_per_step": ms_serial / k_e2e,
                          "api": "one handle: lfx_extract_batch then lfx_fetch_counts + lfx_fetch_features, batch by batch"},
               "host_buffers": f"pinned, NUMA node {numa.value} (node of the rank's GPU)" if numa.value >= 0 else "pinned (no NUMA node reported)"}
        fe2.close()
        lib.lfx_host_free(hptr)
        lib.lfx_host_free(h_edge)
        lib.lfx_host_free(h_surf)

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only; bounded sample) and, with its results, parity of
    #      THIS run's GPU batch: the reference extracts the first n_sample scans of the very buffer the GPU timed
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = host_cores()
        # bounded sample: ~10 s of CPU work (the reference runs ~2 Mpts/s per thread)
        n_sample = min(max(int(4.0e6 * cores / per_scan), cores), scans_per_gpu)
        loops = 5
        sample = d_in[: int(offs[n_sample])].cpu().numpy()
        clouds = [sample[int(offs[s]): int(offs[s + 1])] for s in range(n_sample)]
        v, kind, pts, secs, results = cpu_reference_run(clouds, cores, loops=loops)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"the first {n_sample} scans of the timed batch x {loops} passes ({pts} points, {secs:.2f} s), scans round-robin over {cores} host threads, extraction only"}
        if results is not None:
            fe.extract_views(dev_views, keep=d_in)
            parity = parity_check(fe.fetch(fetch_points=True), clouds, results, kind)
        del sample, clouds

    # ---- the other BASELINE.json configs and the deployed chain, a few steps each (rank 0, N=1 only)
    workloads = None
    if not args.no_workloads:
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            if hasattr(sharded, "close"):
                sharded.close()
        fe.close()
        del d_in
        torch.cuda.empty_cache()
        workloads = {}
        for name in sorted(WORKLOADS):
            if name != args.workload:
                workloads[name] = measure_workload(name, local_rank, dev, stream, max(5, min(args.steps, 10)), peak, rank, world)
        if world == 1:
            workloads["os128_raw_convert_extract"] = measure_chain(local_rank, dev, stream, max(3, min(args.steps, 5)), peak)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": workload_config(args.workload, sensor, sp.n_rings, sp.n_cols, scans_per_gpu, n_points, world),
            "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "exchange": exchange, "workloads": workloads,
        }
        print(json.dumps(line), flush=True)
    fe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
