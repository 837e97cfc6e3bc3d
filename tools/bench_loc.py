#!/usr/bin/env python3
"""Throughput of the localization residual build (lfx_loc_edge / lfx_loc_surface) on a synthetic map, next to the
reference's own neighbour search (vendored nanoflann compiled in place, single thread) when oracle/_ref is present.
One JSON line. usage: python tools/bench_loc.py [--edge-map 100000] [--surface-map 400000] [--edge 2000] [--surface 8000]"""
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
    ap.add_argument("--edge-map", type=int, default=100000)
    ap.add_argument("--surface-map", type=int, default=400000)
    ap.add_argument("--edge", type=int, default=2000)
    ap.add_argument("--surface", type=int, default=8000)
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    import torch

    from lidar_feature_extraction_b200 import FeatureExtraction, LoamProblem

    rng = np.random.default_rng(0)

    def cloud(n):
        a = rng.uniform([-60, -60, -2], [60, 60, 6], size=(n, 3)).astype(np.float32)
        return np.concatenate([a, np.ones((n, 1), np.float32)], axis=1)

    em, sm = cloud(args.edge_map), cloud(args.surface_map)
    es, ss = cloud(args.edge), cloud(args.surface)
    q, t = np.array([0.0, 0.0, 0.0, 1.0]), np.zeros(3)
    out = {"metric": "features_per_sec_residual_build", "unit": "features/s",
           "config": {"edge_map": args.edge_map, "surface_map": args.surface_map, "edge_features": args.edge,
                      "surface_features": args.surface, "n_neighbors": 15,
                      "search": "exhaustive, exact" if os.environ.get("LFX_LOC_EXHAUSTIVE", "0") not in ("", "0") else "uniform grid (1 m cells), exact"}}
    with FeatureExtraction() as fe:
        prob = LoamProblem(fe, torch.from_numpy(em).cuda(), torch.from_numpy(sm).cuda(), 15)
        d_es, d_ss = torch.from_numpy(es).cuda(), torch.from_numpy(ss).cuda()
        for _ in range(3):
            prob.make_edge(d_es, q, t)
            prob.make_surface(d_ss, q, t)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prob.make_edge(d_es, q, t)
        te = (time.perf_counter() - t0) / args.steps
        t0 = time.perf_counter()
        for _ in range(args.steps):
            prob.make_surface(d_ss, q, t)
        ts = (time.perf_counter() - t0) / args.steps
    out["edge_ms"], out["surface_ms"] = te * 1e3, ts * 1e3
    out["value"] = (args.edge + args.surface) / (te + ts)
    out["pairs_per_sec"] = (args.edge * args.edge_map + args.surface * args.surface_map) / (te + ts)
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_knn.so")
    if os.path.exists(ref):
        L = C.CDLL(ref)
        L.ref_knn.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        secs, build = 0.0, 0.0
        for m, s in ((em, es), (sm, ss)):
            md, qd = np.ascontiguousarray(m[:, :3], np.float64), np.ascontiguousarray(s[:, :3], np.float64)
            idx, d2 = np.zeros((len(qd), 15), np.uint64), np.zeros((len(qd), 15))
            t0 = time.perf_counter()
            L.ref_knn(md.ctypes.data, len(md), 3, 10, qd.ctypes.data, len(qd), 15, idx.ctypes.data, d2.ctypes.data)
            secs += time.perf_counter() - t0
            t0 = time.perf_counter()
            L.ref_knn(md.ctypes.data, len(md), 3, 10, qd.ctypes.data, 0, 15, idx.ctypes.data, d2.ctypes.data)   # the build alone
            build += time.perf_counter() - t0
        out["cpu_baseline"] = {"value": (args.edge + args.surface) / max(secs - build, 1e-9), "unit": "features/s", "cores": 1, "kind": "reference",
                               "sample": f"nanoflann 15-NN of the same queries, queries only (the kd-tree build, {build:.3f} s per map pair, is "
                                         "excluded: the reference builds it once per map, not per iteration); neighbour search only, no residuals",
                               "with_build": (args.edge + args.surface) / secs}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
