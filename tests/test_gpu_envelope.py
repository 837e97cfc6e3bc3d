"""The reference has no upper bound on ring length, convolution_padding or n_blocks (hyper_parameter.hpp:45-53 asserts
"> 0" only, index_range.cpp:32-66). Whatever the two on-chip kernels do not hold runs on k_extract_rings_big
(lfx_big.cuh) and must come out exactly as the oracle has it (the oracle is pinned to the compiled reference for these
parameter sets in tests/test_oracle_vs_ref.py)."""
import numpy as np
import pytest

import adversarial as adv
from helpers import compare_scan, oracle_params

pytestmark = pytest.mark.gpu


def _fe(hp=None, **kw):
    from lidar_feature_extraction_b200 import FeatureExtraction, HyperParameters

    kw.setdefault("want_sorted_src", True)
    kw.setdefault("want_curvature", True)
    return FeatureExtraction(hp or HyperParameters(), device=0, **kw)


def _hp(**kw):
    from lidar_feature_extraction_b200 import HyperParameters

    return HyperParameters(**kw)


def _check(oracle, hp, clouds, **kw):
    from oracle import binding as ob

    with _fe(hp, **kw) as fe:
        out = fe.extract_batch(clouds)
        stats = fe.batch_stats()
    for s, cloud in enumerate(clouds):
        compare_scan(out, s, cloud, oracle.extract_scan(cloud, oracle_params(ob, hp)))
    return out, stats


@pytest.mark.parametrize("shuffle", ["none", "interleave", "random", "rotate", "rotate_reverse", "reverse"])
def test_rings_longer_than_the_on_chip_capacity(oracle, shuffle):
    """Rings above lfx_options.max_ring_points (here 256) and above the hard 8192 of the on-chip kernel."""
    clouds = [adv.ragged_scan(3, [100, 400, 3, 0, 1500], shuffle=shuffle),
              adv.ragged_scan(4, [10000, 300, 9001], shuffle=shuffle, zero_xy=1)]
    _check(oracle, _hp(), clouds[:1], max_ring_points=256)
    _check(oracle, _hp(), clouds)
    _check(oracle, _hp(padding=2, neighbor_degree_threshold=3.0, edge_threshold=50.0, max_range=1000.0), clouds)


@pytest.mark.parametrize("kind", ["ramp", "plateau", "gaps", "steps", "spiky"])
def test_long_rings_of_every_kind(oracle, kind):
    """Selection depth (ramp), exact ties (plateau), broken links (gaps), occlusion (steps) on rings of 9000+ points."""
    clouds = [adv.ragged_scan(7, [9000, 12000], kinds=[kind], shuffle="interleave")]
    _check(oracle, _hp(), clouds)
    _check(oracle, _hp(n_blocks=1), clouds)


PARAMSETS = {
    "p20": dict(padding=20),
    "p40b3": dict(padding=40, n_blocks=3, edge_threshold=0.5, surface_threshold=0.5),
    "b100": dict(n_blocks=100),
    "p1b300": dict(padding=1, n_blocks=300, neighbor_degree_threshold=5.0),
    "p33b70": dict(padding=33, n_blocks=70),
}


@pytest.mark.parametrize("pname", sorted(PARAMSETS))
@pytest.mark.parametrize("shuffle", ["none", "interleave", "random", "rotate_reverse"])
def test_parameters_beyond_the_compiled_envelope(oracle, pname, shuffle):
    hp = _hp(**PARAMSETS[pname])
    clouds = []
    for seed in range(4):
        rng = np.random.default_rng(300 + seed)
        lengths = [int(v) for v in rng.choice([0, 1, 6, 23, 40, 67, 97, 300, 777, 2048, 2500], size=8)]
        clouds.append(adv.ragged_scan(seed, lengths, shuffle=shuffle, zero_xy=seed % 3))
    _check(oracle, hp, clouds)


def test_sensor_shaped_scans_with_a_long_window(oracle):
    """A regular scan under padding 20: the sector kernel is not compiled for it, every ring takes the big kernel."""
    from lidar_feature_extraction_b200 import synth

    clouds = [synth.scan_host(synth.spec("vlp16"), f) for f in range(2)]
    _, stats = _check(oracle, _hp(padding=20), clouds)
    assert stats["fast_rings"] == [0, 0, 0] and stats["general_scans"] == 2


@pytest.mark.parametrize("padding", [1, 3, 4, 6, 7, 8])
@pytest.mark.parametrize("sensor", ["vlp16", "hdl64", "os128"])
def test_every_padding_up_to_eight_runs_on_the_sector_kernel(oracle, padding, sensor):
    """convolution_padding 1..8 are compiled instantiations of k_extract_sectors (5 and 2 in lfx_api.cu, the others in
    lfx_sector_extra.cu): regular scans take the strided sector path, drop-out scans (hdl64) the indexed one; no ring
    is left to the per-ring kernels."""
    from lidar_feature_extraction_b200 import synth

    clouds = [synth.scan_host(synth.spec(sensor), f) for f in range(2)]
    _, stats = _check(oracle, _hp(padding=padding), clouds)
    assert stats["general_rings"] == 0, stats
    assert sum(stats["fast_rings"]) + sum(stats["indexed_rings"]) == 2 * synth.spec(sensor).n_rings, stats
    if sensor != "hdl64":
        assert stats["general_scans"] == 0, stats


def test_a_4096_column_sweep_runs_on_chip_by_default(oracle):
    """HDL-64 at 5 Hz / a 4096-column sweep: rings of ~4000 points have sectors too long for the sector kernel (683
    positions at six sectors) but fit the on-chip per-ring kernel of a default handle (max_ring_points 4096); rings beyond
    that take the unbounded kernel. Either way the result is the oracle's."""
    clouds = [adv.ragged_scan(5, [4096, 4000, 3600, 4097, 64], shuffle="interleave")]
    _check(oracle, _hp(), clouds)
    _, stats = _check(oracle, _hp(n_blocks=12), clouds)      # twelve sectors: 341 positions each, the sector kernel's size
    assert sum(stats["indexed_rings"]) == 4, stats           # (the 64-point ring has sectors of 4 points: fine as well)


@pytest.mark.parametrize("n_ragged", [0, 1, 17, 48])
def test_batches_big_enough_for_the_conditional_node(oracle, n_ragged):
    """From 32 scans on, the batch graph wraps the general path in an IF node whose condition a kernel sets on the
    device (lfx_api.cu, COND_MIN_SCANS): batches with no, one, some and only ragged scans, replayed from the cached
    graph with the ragged scans at other places, must all come out as the oracle has them."""
    from lidar_feature_extraction_b200 import synth

    sp = synth.spec("vlp16")
    regular = [synth.scan_host(sp, f) for f in range(6)]
    ragged = [adv.ragged_scan(40 + k, [300, 97, 0, 640, 23, 1200][k % 6:] + [150], shuffle=["interleave", "random", "rotate"][k % 3])
              for k in range(6)]
    n = 48
    order_a = [i < n_ragged for i in range(n)]
    order_b = [i >= n - n_ragged for i in range(n)]
    hp = _hp()
    from oracle import binding as ob

    with _fe(hp) as fe:
        for order in (order_a, order_b, order_a):
            clouds = [ragged[i % 6] if is_r else regular[i % 6] for i, is_r in enumerate(order)]
            out = fe.extract_batch(clouds)
            stats = fe.batch_stats()
            assert stats["general_scans"] == n_ragged, stats
            for s in (0, 1, 16, 17, 30, 31, 46, 47):
                compare_scan(out, s, clouds[s], oracle.extract_scan(clouds[s], oracle_params(ob, hp)))
