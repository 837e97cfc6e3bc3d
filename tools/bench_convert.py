#!/usr/bin/env python3
"""Throughput of the device converter (k_convert) on Ouster-like 48-byte raw clouds, device-resident, timed with
CUDA events; prints one JSON line with the HBM roofline fraction (algorithmic bytes = point_step read + 32 B
written per kept point) and, beside it, the numpy port of the reference's converter on a bounded sample.
usage: python tools/bench_convert.py [--scans 1250] [--steps 10]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scans", type=int, default=1250)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--chain", action="store_true", help="also time raw clouds -> converter -> extraction on the device")
    args = ap.parse_args()
    import torch

    from lidar_feature_extraction_b200 import FeatureExtraction, PointCloud2, PointField, PointTypeConverter, synth
    from lidar_feature_extraction_b200 import _native as N

    sp = synth.spec("os128")
    per = sp.n_rings * sp.n_cols
    dev = torch.device("cuda", 0)
    fe = FeatureExtraction(max_rings=128)
    lib = N.lib()
    wire = torch.empty((args.scans * per, 32), dtype=torch.uint8, device=dev)
    assert lib.lfx_synth_batch_device(fe.handle, C.byref(sp), 0, args.scans, wire.data_ptr()) == 0
    # Ouster driver layout (test_convert.py:177-187): x,y,z f32 @0, intensity f32 @16, t u32 @20, reflectivity u16 @24,
    # ring u8 @26, noise u16 @28, range u32 @32, point_step 48; 3 % of the returns zeroed like a driver does
    raw = torch.zeros((args.scans * per, 48), dtype=torch.uint8, device=dev)
    raw[:, 0:12] = wire[:, 0:12]
    raw[:, 16:20] = wire[:, 16:20]
    raw[:, 26] = wire[:, 20]
    raw[:, 20:24] = torch.randint(0, 256, (args.scans * per, 4), dtype=torch.uint8, device=dev)
    dead = torch.rand(args.scans * per, device=dev) < 0.03
    raw[dead] = 0
    del wire
    fields = [PointField("x", 0, 7), PointField("y", 4, 7), PointField("z", 8, 7), PointField("intensity", 16, 7), PointField("t", 20, 6),
              PointField("reflectivity", 24, 4), PointField("ring", 26, 2), PointField("noise", 28, 4), PointField("range", 32, 6)]
    raw3 = raw.view(args.scans, per, 48)
    msgs = [PointCloud2(data=raw3[s], point_step=48, fields=fields) for s in range(args.scans)]
    conv = PointTypeConverter(fe)
    for _ in range(args.warmup):
        conv.convert_batch(msgs)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    kms = []
    ms = C.c_float()
    for _ in range(args.steps):
        res = conv.convert_batch(msgs)
        assert lib.lfx_last_convert_ms(fe.handle, C.byref(ms)) == 0
        kms.append(ms.value)
    torch.cuda.synchronize()
    kernel_ms = float(np.mean(kms))
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps     # includes the host-side plans, H2D of the plans, D2H of the widths
    kept = sum(int(res.kept[s]) for s in range(args.scans))
    n = args.scans * per
    alg = 48 * n + 32 * kept
    # roofline: the kernel alone (CUDA events on the launching stream, lfx_last_convert_ms); value: the synchronous call
    peak = 6551.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    # CPU: numpy port of the reference converter on a bounded sample (the reference itself is pure-Python struct code)
    from oracle import convert_oracle as co

    ofields = [co.Field(f.name, f.offset, f.datatype) for f in fields]
    sample = raw3[:8].cpu().numpy()
    t0 = time.perf_counter()
    for s in range(sample.shape[0]):
        co.convert(sample[s], ofields, 48, False)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"metric": "points_converted_per_sec", "value": n / (wall_ms * 1e-3), "unit": "points/s", "ms_per_call": wall_ms,
                      "config": {"workload": f"os128 raw 48-byte clouds x {args.scans}", "kept_fraction": kept / n},
                      "roofline": {"bound": "hbm", "kernel": "k_convert", "achieved": alg / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                   "frac": alg / (kernel_ms * 1e-3) / 1e9 / peak, "algorithmic_bytes": alg, "kernel_ms": kernel_ms,
                                   "points_per_sec_kernel": n / (kernel_ms * 1e-3)},
                      "cpu_baseline": {"value": sample.shape[0] * per / cpu_s, "unit": "points/s", "cores": 1, "kind": "port",
                                       "sample": f"{sample.shape[0]} clouds, numpy port of convert.py"}}))
    if args.chain:
        # /points_raw -> features without leaving the device: the converted clouds (3 % of the returns removed, so rings
        # are ragged and the scans take the bucketing + indexed sector path) feed lfx_extract_batch directly
        def chain_step():
            conv.convert_batch(msgs)
            fe.extract_views(fe.view_array([conv.view(s) for s in range(args.scans)]))
        for _ in range(2):
            chain_step()
        fe.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            chain_step()
        fe.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / args.steps
        fe.set_stage_timing(True)
        chain_step()
        fe.synchronize()
        st = fe.last_stage_ms()
        print(json.dumps({"metric": "raw_points_per_sec_converted_and_extracted", "value": n / (ms * 1e-3), "unit": "points/s",
                          "ms_per_step": ms, "convert_kernel_ms": kernel_ms,
                          "extract_stage_ms": dict(zip(["probe", "sectors", "bucketing", "sectors_indexed", "rings", "pack"], [round(v, 3) for v in st])),
                          "paths": fe.batch_stats()}))
    conv.close()
    fe.close()


if __name__ == "__main__":
    main()
