#!/usr/bin/env python
"""Generates tests/golden/envelope_*.npz: outputs of the REFERENCE ITSELF (oracle/_ref/libref_stable.so, the reference's
own extraction sources compiled in place) for parameters and ring lengths beyond what the on-chip CUDA kernels are
compiled for - convolution_padding above 15, more than 64 sectors, rings above 8192 points. The reference has no such
bounds (hyper_parameter.hpp:45-53, index_range.cpp:32-66); the library runs these on k_extract_rings_big (lfx_big.cuh).
Run where /root/reference exists:   python tests/golden/make_envelope_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import adversarial as adv  # noqa: E402
from oracle import binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

PARAMSETS = {
    "default": dict(),                      # (the long ring alone is outside the on-chip kernels)
    "p20": dict(padding=20),
    "b100": dict(n_blocks=100),
    "p33b70": dict(padding=33, n_blocks=70),
}
FIELDS = ("ring_ids", "ring_sizes", "ring_skipped", "sorted_src", "labels", "curvature", "edge_idx", "surface_idx")


def cases():
    yield "envelope_long_ring_interleaved", adv.ragged_scan(21, [9000, 300, 64, 1500], shuffle="interleave", zero_xy=1)
    yield "envelope_ragged_random", adv.ragged_scan(22, [0, 6, 23, 97, 640, 2500, 40], shuffle="random")


def main():
    if not ob.Reference.available("stable"):
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    ref = ob.Reference("stable")
    for name, cloud in cases():
        out = {"cloud": cloud}
        for pname, kw in PARAMSETS.items():
            r = ref.extract_scan(cloud, ob.default_params(**kw))
            for f in FIELDS:
                out[f"{pname}.{f}"] = np.asarray(getattr(r, f))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: {len(cloud)} points -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
