"""Synthetic scans for the BASELINE.json sensor shapes (bindings of lfx_synth_* in include/lfx.h).

The host generator is loaded from ``libsynth.so``, a host-only build of the same code (csrc/lfx_synth_host.cpp), so that
generating scans does not map the CUDA product library: bench.py's reference arm and the oracle-only tests run
without ``liblfx.so`` in their address space. The device generator (``lfx_synth_batch_device``) lives in ``liblfx.so``."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N

NAMES = ("vlp16", "hdl32", "hdl64", "os128")
SYNTH_LIB_PATH = os.path.join(N.PKG_DIR, "libsynth.so")
_synth = None


def _lib() -> C.CDLL:
    global _synth
    if _synth is None:
        if not os.path.exists(SYNTH_LIB_PATH):
            raise ImportError(f"{SYNTH_LIB_PATH} is missing (run `python -c 'import __graft_entry__ as g; g.build()'`)")
        L = C.CDLL(SYNTH_LIB_PATH)
        L.lfx_synth_named.argtypes = [C.c_char_p, C.POINTER(N.SynthSpec)]
        L.lfx_synth_scan_host.argtypes = [C.POINTER(N.SynthSpec), C.c_uint64, C.c_void_p, C.POINTER(C.c_uint32)]
        _synth = L
    return _synth


def spec(name: str, **overrides) -> N.SynthSpec:
    s = N.SynthSpec()
    if _lib().lfx_synth_named(name.encode(), C.byref(s)) != N.LFX_OK:
        raise ValueError(f"unknown synthetic sensor {name!r}; known: {NAMES}")
    for k, v in overrides.items():
        setattr(s, k, v)
    return s


def scan_host(sp: N.SynthSpec, frame: int) -> np.ndarray:
    """One scan as PointCloud2 payload bytes: uint8 array [n_points, 32] (deployed wire layout)."""
    buf = np.zeros((sp.n_rings * sp.n_cols, 32), dtype=np.uint8)
    n = C.c_uint32(0)
    rc = _lib().lfx_synth_scan_host(C.byref(sp), frame, buf.ctypes.data, C.byref(n))
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
