#!/usr/bin/env python
"""Generates tests/golden/loc_knn.npz: a small edge / surface map, a scan, a pose, and the neighbour index lists and
squared distances the REFERENCE's kd-tree returns for them (vendored nanoflann compiled in place:
oracle/_ref/libref_knn.so, driven like localization/src/kdtree.cpp:42-55). Run where /root/reference exists."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loc_oracle as lo  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def scene(seed=7, n_edge=6000, n_surf=20000):
    """Edge map: points along a few hundred vertical / horizontal line segments; surface map: points on walls and floor."""
    rng = np.random.default_rng(seed)
    segs = []
    for _ in range(300):
        a = rng.uniform([-40, -40, -1], [40, 40, 4])
        d = rng.normal(size=3) * [0.2, 0.2, 1.0] if rng.random() < 0.6 else rng.normal(size=3) * [1.0, 1.0, 0.1]
        d /= np.linalg.norm(d)
        s = rng.uniform(0, 3, size=n_edge // 300)
        segs.append(a + s[:, None] * d + rng.normal(0, 0.01, size=(len(s), 3)))
    edge = np.concatenate(segs).astype(np.float32)
    planes = []
    for _ in range(40):
        o = rng.uniform([-40, -40, -1], [40, 40, 1])
        u = rng.normal(size=3); u /= np.linalg.norm(u)
        v = np.cross(u, rng.normal(size=3)); v /= np.linalg.norm(v)
        ab = rng.uniform(0, 8, size=(n_surf // 40, 2))
        planes.append(o + ab[:, :1] * u + ab[:, 1:] * v + rng.normal(0, 0.01, size=(len(ab), 3)))
    surf = np.concatenate(planes).astype(np.float32)
    return edge, surf, rng


def main():
    lib = os.path.join(ROOT, "oracle", "_ref", "libref_knn.so")
    if not os.path.exists(lib):
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    L = C.CDLL(lib)
    L.ref_knn.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    edge, surf, rng = scene()
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    t = rng.normal(size=3) * 2
    out = {"edge_map": edge, "surface_map": surf, "q_xyzw": q, "t": t}
    for name, m in (("edge", edge), ("surface", surf)):
        # scan features: map points seen from the sensor frame (inverse pose) with a little noise
        pick = rng.choice(len(m), size=400, replace=False)
        R = lo.rotation_matrix(q)
        scan = ((m[pick].astype(np.float64) + rng.normal(0, 0.05, size=(400, 3)) - t) @ R).astype(np.float32)
        queries = np.array([lo.transform(q, t, p) for p in scan.astype(np.float64)])
        md = np.ascontiguousarray(m, np.float64)
        for k in (5, 15):
            idx = np.zeros((len(queries), k), np.uint64)
            d2 = np.zeros((len(queries), k), np.float64)
            assert L.ref_knn(md.ctypes.data, len(md), 3, 10, queries.ctypes.data, len(queries), k, idx.ctypes.data, d2.ctypes.data) == 0
            out[f"{name}.idx{k}"], out[f"{name}.d2_{k}"] = idx.astype(np.uint32), d2
        out[f"{name}.scan"] = scan
    path = os.path.join(HERE, "loc_knn.npz")
    np.savez_compressed(path, **out)
    print(f"loc_knn.npz: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
