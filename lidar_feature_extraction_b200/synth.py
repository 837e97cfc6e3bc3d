"""Synthetic scans for the BASELINE.json sensor shapes (bindings of lfx_synth_* in include/lfx.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

NAMES = ("vlp16", "hdl32", "hdl64", "os128")


def spec(name: str, **overrides) -> N.SynthSpec:
    s = N.SynthSpec()
    if N.lib().lfx_synth_named(name.encode(), C.byref(s)) != N.LFX_OK:
        raise ValueError(f"unknown synthetic sensor {name!r}; known: {NAMES}")
    for k, v in overrides.items():
        setattr(s, k, v)
    return s


def scan_host(sp: N.SynthSpec, frame: int) -> np.ndarray:
    """One scan as PointCloud2 payload bytes: uint8 array [n_points, 32] (deployed wire layout)."""
    buf = np.zeros((sp.n_rings * sp.n_cols, 32), dtype=np.uint8)
    n = C.c_uint32(0)
    rc = N.lib().lfx_synth_scan_host(C.byref(sp), frame, buf.ctypes.data, C.byref(n))
    if rc != N.LFX_OK:
        raise RuntimeError(f"lfx_synth_scan_host failed: {rc}")
    return buf[: n.value]


def fields(cloud: np.ndarray):
    """Decode (x, y, z, intensity, ring) columns of a [n, 32] uint8 wire cloud."""
    c = np.ascontiguousarray(cloud).reshape(-1, 32)
    f = c.view(np.float32).reshape(-1, 8)
    ring = c.view(np.uint16).reshape(-1, 16)[:, 10]
    return f[:, 0], f[:, 1], f[:, 2], f[:, 4], ring


def make_cloud(x, y, z, ring, intensity=None) -> np.ndarray:
    """Assemble the deployed 32-byte layout (convert.py:137-145) from columns."""
    n = len(x)
    out = np.zeros((n, 32), dtype=np.uint8)
    f = out.view(np.float32).reshape(n, 8)
    f[:, 0] = x
    f[:, 1] = y
    f[:, 2] = z
    f[:, 3] = 1.0
    f[:, 4] = 0.0 if intensity is None else intensity
    out.view(np.uint16).reshape(n, 16)[:, 10] = ring
    return out
