"""Golden fixtures (tests/golden/*.npz): outputs of the REFERENCE's own extraction code (compiled in place,
see tests/golden/make_golden.py) on small scans. CPU: the oracle must reproduce them bit for bit. GPU: the
CUDA path, through the C ABI, must reproduce them - labels, ring order and feature index sets bit-exact,
curvature within 1e-6 relative (north_star tolerance; it is bit-exact in practice)."""
import glob
import os
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import compare_scan, oracle_params

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                  if not os.path.basename(p).startswith(("convert_", "loc_", "envelope_")))  # converter / localization / envelope fixtures have their own tests
PARAMSETS = {
    "default": dict(),
    "yaml": dict(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0),
    "p2b4": dict(padding=2, n_blocks=4, edge_threshold=0.02, surface_threshold=0.2),
}
FIELDS = ("ring_ids", "ring_sizes", "ring_skipped", "sorted_src", "labels", "curvature", "edge_idx", "surface_idx")


def _load(path, pname):
    z = np.load(path)
    return z["cloud"], SimpleNamespace(**{f: z[f"{pname}.{f}"] for f in FIELDS})


def test_fixtures_present():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize("pname", sorted(PARAMSETS))
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-4] for p in FIXTURES])
def test_oracle_reproduces_reference_outputs(oracle, path, pname):
    from oracle import binding as ob

    cloud, want = _load(path, pname)
    got = oracle.extract_scan(cloud, ob.default_params(**PARAMSETS[pname]))
    for f in ("ring_ids", "ring_sizes", "ring_skipped", "labels", "edge_idx", "surface_idx"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    assert np.array_equal(got.curvature.view(np.uint64), want.curvature.view(np.uint64)), "curvature bits"
    pos = 0
    for n, skipped in zip(want.ring_sizes, want.ring_skipped):
        if not skipped:  # a skipped ring contributes nothing; its internal order is not part of the contract
            assert np.array_equal(got.sorted_src[pos:pos + n], want.sorted_src[pos:pos + n])
        pos += n


@pytest.mark.gpu
@pytest.mark.parametrize("diag", [True, False])
@pytest.mark.parametrize("pname", sorted(PARAMSETS))
def test_cuda_reproduces_reference_outputs(pname, diag):
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters

    hp = HyperParameters(**PARAMSETS[pname])
    loaded = [_load(p, pname) for p in FIXTURES]
    with FeatureExtraction(hp, device=0, want_sorted_src=diag, want_curvature=diag) as fe:
        out = fe.extract_batch([c for c, _ in loaded])
        stats = fe.batch_stats()
    for s, (cloud, want) in enumerate(loaded):
        compare_scan(out, s, cloud, want)
    assert sum(stats["fast_rings"]) > 0 and stats["general_scans"] > 0  # both pipelines are exercised


# ---- beyond the compiled envelope (tests/golden/make_envelope_golden.py): paddings above 15, more than 64 sectors, a
#      9000-point ring; the CUDA side runs these on k_extract_rings_big
ENVELOPE = sorted(glob.glob(os.path.join(HERE, "golden", "envelope_*.npz")))
ENVELOPE_PARAMS = {
    "default": dict(),
    "p20": dict(padding=20),
    "b100": dict(n_blocks=100),
    "p33b70": dict(padding=33, n_blocks=70),
}


@pytest.mark.parametrize("pname", sorted(ENVELOPE_PARAMS))
@pytest.mark.parametrize("path", ENVELOPE, ids=[os.path.basename(p)[:-4] for p in ENVELOPE])
def test_oracle_reproduces_reference_outputs_beyond_the_envelope(oracle, path, pname):
    from oracle import binding as ob

    cloud, want = _load(path, pname)
    got = oracle.extract_scan(cloud, ob.default_params(**ENVELOPE_PARAMS[pname]))
    for f in ("ring_ids", "ring_sizes", "ring_skipped", "labels", "edge_idx", "surface_idx"):
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    assert np.array_equal(got.curvature.view(np.uint64), want.curvature.view(np.uint64)), "curvature bits"


@pytest.mark.gpu
@pytest.mark.parametrize("pname", sorted(ENVELOPE_PARAMS))
def test_cuda_reproduces_reference_outputs_beyond_the_envelope(pname):
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters

    assert len(ENVELOPE) >= 2
    hp = HyperParameters(**ENVELOPE_PARAMS[pname])
    loaded = [_load(p, pname) for p in ENVELOPE]
    with FeatureExtraction(hp, device=0, want_sorted_src=True, want_curvature=True) as fe:
        out = fe.extract_batch([c for c, _ in loaded])
    for s, (cloud, want) in enumerate(loaded):
        compare_scan(out, s, cloud, want)
