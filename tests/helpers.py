"""Shared helpers for the parity tests: compare one scan of CUDA output with a checker's ScanResult."""
from __future__ import annotations

import numpy as np

from lidar_feature_extraction_b200 import synth
from lidar_feature_extraction_b200 import _native as N


def oracle_params(ob, hp):
    return ob.Params(hp.padding, hp.neighbor_degree_threshold, hp.distance_diff_threshold,
                     hp.parallel_beam_min_range_ratio, hp.edge_threshold, hp.surface_threshold,
                     hp.min_range, hp.max_range, hp.n_blocks)


def compare_scan(out, s, cloud, ref, curv_rtol=1e-6, check_order=True):
    """out: BatchOutput; s: scan index in the batch; cloud: [n,32] wire bytes; ref: checker ScanResult
    (rings removed as sparse are absent from ref; present with label 255 in `out`)."""
    base = int(out.point_base[s])
    rings = out.rings[s]
    x, y, z, _, _ = synth.fields(cloud)
    pos = 0
    e_want, s_want = [], []
    present = [r for r in range(len(rings)) if rings[r]["count"] > 0]
    kept = [r for r in present if rings[r]["status"] != N.LFX_RING_SPARSE]
    assert kept == list(ref.ring_ids), (kept, list(ref.ring_ids))
    for k, r in enumerate(kept):
        cnt, off = int(rings[r]["count"]), int(rings[r]["offset"])
        assert cnt == ref.ring_sizes[k]
        sl = slice(base + off, base + off + cnt)
        rl = slice(pos, pos + cnt)
        assert bool(ref.ring_skipped[k]) == (rings[r]["status"] == N.LFX_RING_SKIPPED), f"ring {r} skip status"
        if check_order and out.sorted_src is not None and not ref.ring_skipped[k]:
            assert np.array_equal(out.sorted_src[sl].astype(np.int64), ref.sorted_src[rl].astype(np.int64)), f"ring {r}: angle order differs"
        assert np.array_equal(out.labels[sl], ref.labels[rl]), f"ring {r}: labels differ at {np.nonzero(out.labels[sl] != ref.labels[rl])[0][:10]}"
        if out.curvature is not None:
            np.testing.assert_allclose(out.curvature[sl], ref.curvature[rl], rtol=curv_rtol, atol=0, err_msg=f"ring {r} curvature")
        pos += cnt
    src = ref.sorted_src
    for idx, want in ((ref.edge_idx, e_want), (ref.surface_idx, s_want)):
        want.append(np.stack([x[src[idx]], y[src[idx]], z[src[idx]], np.ones(len(idx), np.float32)], axis=1))
    assert out.counts[s, 0] == len(ref.edge_idx) and out.counts[s, 1] == len(ref.surface_idx), (out.counts[s], len(ref.edge_idx), len(ref.surface_idx))
    assert np.array_equal(out.scan_edges(s), e_want[0]), "edge cloud differs"
    assert np.array_equal(out.scan_surfaces(s), s_want[0]), "surface cloud differs"
    for r in present:
        if rings[r]["status"] == N.LFX_RING_SPARSE:
            cnt, off = int(rings[r]["count"]), int(rings[r]["offset"])
            assert (out.labels[base + off: base + off + cnt] == 255).all()
