#!/usr/bin/env python
"""Generates tests/golden/*.npz: inputs and the outputs of the REFERENCE ITSELF on them.

The outputs come from oracle/_ref/libref_stable.so, i.e. the reference's own extraction sources compiled in
place from /root/reference (oracle/Makefile target `ref`) with Argsort in its (value, index) tie-break form.
Run it where /root/reference exists:   python tests/golden/make_golden.py
The fixtures are small on purpose (a few rings per scan); they travel with the repository, so the GPU box -
which has no /root/reference - can check the CUDA path against reference outputs, not only against the port.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import adversarial as adv  # noqa: E402
from lidar_feature_extraction_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

PARAMSETS = {
    "default": dict(),
    "yaml": dict(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0),
    "p2b4": dict(padding=2, n_blocks=4, edge_threshold=0.02, surface_threshold=0.2),
}


def firing_order(seed, n_rings, width, direction="cw"):
    """column-major scan, every ring a rotated monotone azimuth sequence (what a spinning sensor emits)"""
    rng = np.random.default_rng(seed)
    X = np.zeros((width, n_rings), np.float32)
    Y = np.zeros((width, n_rings), np.float32)
    for k in range(n_rings):
        x, y = adv._ring_points(rng, width, adv.KINDS[(seed + k) % len(adv.KINDS)])
        order = np.arange(width)[::-1] if direction == "cw" else np.arange(width)
        order = np.roll(order, int(rng.integers(0, width)))
        X[:, k], Y[:, k] = x[order], y[order]
    R = np.ascontiguousarray(np.broadcast_to(np.arange(n_rings, dtype=np.uint16), (width, n_rings)))
    Z = rng.normal(0, 1, size=(width, n_rings)).astype(np.float32)
    return synth.make_cloud(X.ravel(), Y.ravel(), Z.ravel(), R.ravel())


def subsample_rings(cloud, keep):
    ring = synth.fields(cloud)[4]
    return np.ascontiguousarray(cloud[np.isin(ring, keep)])


def cases():
    vlp = synth.scan_host(synth.spec("vlp16"), 7)
    yield "vlp16_rings_0_5_15", subsample_rings(vlp, [0, 5, 15])          # sensor scene, firing order, 3 x 1800
    hdl = synth.scan_host(synth.spec("hdl64"), 2)
    yield "hdl64_tunnel_rings_3_40", subsample_rings(hdl, [3, 40])         # tunnel + clutter + drop-outs (ragged)
    yield "firing_order_cw_9x640", firing_order(11, 9, 640, "cw")          # every ring kind, clockwise
    yield "firing_order_ccw_4x2048", firing_order(12, 4, 2048, "ccw")
    yield "ragged_random", adv.ragged_scan(5, [0, 3, 6, 11, 17, 23, 97, 300, 777], shuffle="random", zero_xy=1)
    yield "exact_ties", adv.symmetric_ties_scan(3, n_quarter=60, n_rings=3)


def main():
    if not ob.Reference.available("stable"):
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    ref = ob.Reference("stable")
    for name, cloud in cases():
        out = {"cloud": cloud}
        for pname, kw in PARAMSETS.items():
            r = ref.extract_scan(cloud, ob.default_params(**kw))
            for f in ("ring_ids", "ring_sizes", "ring_skipped", "sorted_src", "labels", "curvature", "edge_idx", "surface_idx"):
                out[f"{pname}.{f}"] = np.asarray(getattr(r, f))
            # colored_scan by the reference's own ColorPointsByLabel (ref_color_scan): x,y,z and r,g,b per point
            out[f"{pname}.colored_xyz"], out[f"{pname}.colored_rgb"] = ref.color_scan(cloud, ob.default_params(**kw))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(cloud)} points -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
