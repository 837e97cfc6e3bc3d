"""Adversarial / ragged synthetic scans for parity tests (numpy, deterministic). Every scan is emitted in
the deployed 32-byte wire layout. Distinct azimuths inside a ring are guaranteed (the reference's own
order is unspecified otherwise) unless a builder says differently."""
from __future__ import annotations

import numpy as np

from lidar_feature_extraction_b200 import synth


def _ring_points(rng, n, kind):
    """n points of one ring with strictly increasing azimuth in (-pi, pi); returns x, y (float32)."""
    # strictly distinct azimuths: jittered grid, optional big gaps (broken links)
    base = np.sort(rng.uniform(-np.pi * 0.999, np.pi * 0.999, size=n)) if kind == "random_az" else \
        -np.pi * 0.999 + (np.arange(n) + rng.uniform(-0.3, 0.3, size=n)) * (2 * np.pi * 0.999 / max(n, 1))
    if kind == "gaps" and n > 8:
        # remove azimuth mass to create > threshold gaps
        cut = rng.integers(1, n - 1, size=max(1, n // 50))
        shift = np.zeros(n)
        shift[cut] = rng.uniform(0.03, 0.2, size=len(cut))
        base = base + np.cumsum(shift)
        base = -np.pi * 0.999 + (base - base.min()) * (2 * np.pi * 0.998 / max(base.max() - base.min(), 1e-9))
    az = base
    t = np.linspace(0, 1, n)
    if kind == "plateau":
        r = np.full(n, 7.5)
    elif kind == "ramp":
        r = 5.0 + 3.0 * t                       # monotone: worst case for selection depth
    elif kind == "steps":
        r = 4.0 + np.floor(t * 12) * 0.7        # occlusion jumps both directions
        r[::2 if n < 40 else 37] += 0.0
    elif kind == "far":
        r = np.where(rng.uniform(size=n) < 0.1, rng.uniform(100.5, 150, size=n), rng.uniform(3, 90, size=n))
    elif kind == "near":
        r = np.where(rng.uniform(size=n) < 0.1, rng.uniform(0.001, 0.099, size=n), rng.uniform(0.2, 9, size=n))
    elif kind == "spiky":
        r = 6.0 + np.where(np.arange(n) % 3 == 1, 0.4, 0.0) + rng.normal(0, 0.002, size=n)
    else:
        r = 6.0 + 2.0 * np.sin(6 * az) + rng.normal(0, 0.01, size=n)
        jumps = rng.uniform(size=n) < 0.02
        r = r + np.cumsum(np.where(jumps, rng.normal(0, 1.0, size=n), 0.0))
        r = np.clip(np.abs(r), 0.05, 140)
    x = (r * np.cos(az)).astype(np.float32)
    y = (r * np.sin(az)).astype(np.float32)
    return x, y


KINDS = ["mixed", "gaps", "plateau", "ramp", "steps", "far", "near", "spiky", "random_az"]


def ragged_scan(seed, ring_lengths, kinds=None, shuffle="none", ring_ids=None, zero_xy=0):
    """Scan with the given ring lengths. shuffle: 'none' (ring-major, ascending azimuth), 'interleave'
    (column-major like a sensor), 'reverse' (descending azimuth), 'rotate', 'random' (fully shuffled)."""
    rng = np.random.default_rng(seed)
    xs, ys, zs, rs = [], [], [], []
    for k, n in enumerate(ring_lengths):
        kind = (kinds[k % len(kinds)] if kinds else KINDS[(seed + k) % len(KINDS)])
        x, y = _ring_points(rng, n, kind)
        if zero_xy and n > 4:
            for _ in range(zero_xy):
                j = int(rng.integers(0, n))
                x[j] = 0.0
                y[j] = 0.0
        order = np.arange(n)
        if shuffle == "reverse":
            order = order[::-1]
        elif shuffle == "rotate" and n > 0:
            order = np.roll(order, int(rng.integers(0, n)))
        elif shuffle == "rotate_reverse" and n > 0:
            order = np.roll(order[::-1], int(rng.integers(0, n)))
        xs.append(x[order])
        ys.append(y[order])
        zs.append(rng.normal(0, 1, size=n).astype(np.float32))
        rid = k if ring_ids is None else ring_ids[k]
        rs.append(np.full(n, rid, dtype=np.uint16))
    x, y, z, r = (np.concatenate(a) if a else np.zeros(0, dt) for a, dt in ((xs, np.float32), (ys, np.float32), (zs, np.float32), (rs, np.uint16)))
    n = len(x)
    if shuffle == "random":
        p = rng.permutation(n)
        x, y, z, r = x[p], y[p], z[p], r[p]
    elif shuffle == "interleave" and n:
        # stable interleave: sort by position-within-ring so rings alternate, like column-major firing order
        pos = np.concatenate([np.arange(m) for m in ring_lengths])
        p = np.argsort(pos, kind="stable")
        x, y, z, r = x[p], y[p], z[p], r[p]
    return synth.make_cloud(x, y, z, r)


def symmetric_ties_scan(seed, n_quarter=120, n_rings=3):
    """Rings that are exactly 4-fold rotation symmetric and mirror symmetric about the quadrant diagonals:
    XY ranges repeat bit-for-bit, so curvature has many exact ties, also inside a +-P window
    (exercises the (value, index) tie-break of Argsort, algorithm.hpp:65-71)."""
    rng = np.random.default_rng(seed)
    xs, ys, rs = [], [], []
    for k in range(n_rings):
        m = n_quarter // 2
        az = (np.arange(m) + 0.5) * (np.pi / 4) / m * 0.999          # (0, pi/4)
        rad = 5.0 + 0.5 * np.round(np.sin(9 * az + k) * 3) / 3 + np.where(rng.uniform(size=m) < 0.1, 0.4, 0.0)
        x = (rad * np.cos(az)).astype(np.float32)
        y = (rad * np.sin(az)).astype(np.float32)
        qx = np.concatenate([x, y[::-1]])       # mirror about the diagonal: (x,y) -> (y,x), exact
        qy = np.concatenate([y, x[::-1]])
        fx = np.concatenate([qx, -qy, -qx, qy])  # rotate by 90 degrees three times, exact
        fy = np.concatenate([qy, qx, -qy, -qx])
        xs.append(fx)
        ys.append(fy)
        rs.append(np.full(len(fx), k, np.uint16))
    x, y, r = np.concatenate(xs), np.concatenate(ys), np.concatenate(rs)
    return synth.make_cloud(x, y, np.zeros_like(x), r)
