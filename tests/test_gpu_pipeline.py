"""PipelinedExtraction / lfx::Pipeline: two handles take the batches in turn (download of batch k-1 overlaps the upload of
batch k). Scans are independent (feature_extraction.cpp:92,173-175), so every batch must come out exactly as the oracle
has it, in submission order."""
import ctypes as C

import numpy as np
import pytest

from helpers import compare_scan, oracle_params

pytestmark = pytest.mark.gpu


def test_pipelined_batches_match_oracle_in_submission_order(oracle):
    from lidar_feature_extraction_b200 import FeatureExtraction, PipelinedExtraction, default_params, synth
    from oracle import binding as ob

    hp = default_params()
    batches = [[synth.scan_host(synth.spec(sensor), f) for f in range(b, b + 2)]
               for b, sensor in enumerate(["vlp16", "hdl32", "hdl64", "vlp16", "os128"])]
    outs = []
    with PipelinedExtraction(hp, device=0, want_sorted_src=True, want_curvature=True) as pipe:
        for k, clouds in enumerate(batches):
            views = [FeatureExtraction.wire_view(c) for c in clouds]
            assert pipe.submit(views, keep=clouds) == k % 2
            if pipe.in_flight == 2:
                outs.append(pipe.collect_output())
        while pipe.in_flight:
            outs.append(pipe.collect_output())
        with pytest.raises(Exception):
            pipe.collect_output()
    assert len(outs) == len(batches)
    for clouds, out in zip(batches, outs):
        for s, cloud in enumerate(clouds):
            compare_scan(out, s, cloud, oracle.extract_scan(cloud, oracle_params(ob, hp)), 1e-6)


def test_pipelined_collect_into_pinned_memory_equals_single_handle():
    from lidar_feature_extraction_b200 import FeatureExtraction, PipelinedExtraction, default_params, synth
    from lidar_feature_extraction_b200 import _native as N

    lib = N.lib()
    clouds = [synth.scan_host(synth.spec("hdl32"), f) for f in range(6)]
    with FeatureExtraction(default_params(), device=0) as fe:
        want = fe.extract_batch(clouds, fetch_points=False)
    with PipelinedExtraction(default_params(), device=0) as pipe:
        cap = sum(len(c) for c in clouds)
        numa = C.c_int(-7)
        h_edge = lib.lfx_host_alloc_on(pipe.fe[0].handle, 16 * cap, C.byref(numa))
        h_surf = lib.lfx_host_alloc_on(pipe.fe[0].handle, 16 * cap, None)
        assert h_edge and h_surf and numa.value >= -1
        try:
            views = FeatureExtraction.view_array([FeatureExtraction.wire_view(c) for c in clouds])
            for k in range(4):
                pipe.submit(views, keep=clouds)
                if pipe.in_flight == 2:
                    counts, offsets = pipe.collect(h_edge, cap, h_surf, cap)
                    assert np.array_equal(counts, want.counts) and np.array_equal(offsets, want.offsets)
            counts, offsets = pipe.collect(h_edge, cap, h_surf, cap)
            ne, ns = int(offsets[-1, 0]), int(offsets[-1, 1])
            edge = np.ctypeslib.as_array(C.cast(h_edge, C.POINTER(C.c_float)), shape=(cap, 4))[:ne]
            surf = np.ctypeslib.as_array(C.cast(h_surf, C.POINTER(C.c_float)), shape=(cap, 4))[:ns]
            assert np.array_equal(edge, want.edge_xyz) and np.array_equal(surf, want.surface_xyz)
        finally:
            lib.lfx_host_free(h_edge)
            lib.lfx_host_free(h_surf)
