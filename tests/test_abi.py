"""Host-side checks that run without a GPU: the C-ABI library loads, exports every symbol the header
declares, mirrors the reference's parameter sets, and FAILS LOUDLY without a device (no CPU fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

from lidar_feature_extraction_b200 import _native as N


def test_library_exports_every_declared_symbol():
    lib = N.lib()
    names = N.declared_symbols()
    assert len(names) >= 28
    missing = [s for s in names if not hasattr(lib, s)]
    assert not missing, missing


def test_parameter_sets_match_reference():
    from lidar_feature_extraction_b200 import default_params, launch_yaml_params

    d = default_params()   # hyper_parameter.hpp:35-43
    assert (d.padding, d.neighbor_degree_threshold, d.distance_diff_threshold, d.parallel_beam_min_range_ratio,
            d.edge_threshold, d.surface_threshold, d.min_range, d.max_range, d.n_blocks) == (5, 2.0, 0.3, 0.02, 0.05, 0.05, 0.1, 100.0, 6)
    y = launch_yaml_params()  # lidar_feature_extraction.param.yaml:3-10
    assert (y.padding, y.neighbor_degree_threshold, y.distance_diff_threshold, y.parallel_beam_min_range_ratio,
            y.edge_threshold, y.surface_threshold, y.min_range, y.max_range, y.n_blocks) == (2, 3.0, 0.3, 0.02, 50.0, 0.05, 0.1, 1000.0, 6)


def test_struct_layouts_match_header():
    assert C.sizeof(N.Params) == 72
    assert C.sizeof(N.CloudView) == 40
    assert C.sizeof(N.RingInfo) == 24
    assert C.sizeof(N.SynthSpec) == 48


def test_label_to_color_matches_reference_table():
    from lidar_feature_extraction_b200 import label_to_color

    # color_points.cpp:39-68
    assert [label_to_color(k) for k in range(8)] == [(255, 255, 255), (255, 0, 0), (255, 63, 0), (255, 0, 0), (255, 63, 0),
                                                     (127, 127, 127), (255, 0, 255), (0, 255, 0)]
    with pytest.raises(ValueError):
        label_to_color(8)


def test_invalid_parameters_are_rejected_before_touching_cuda():
    lib = N.lib()
    for field in ("padding", "n_blocks", "edge_threshold", "min_range"):
        p = N.Params()
        lib.lfx_default_params(C.byref(p))
        setattr(p, field, 0)
        h = C.c_void_p()
        assert lib.lfx_create(C.byref(p), None, C.byref(h)) == N.LFX_E_BAD_PARAM  # hyper_parameter.hpp:45-53
        assert b"> 0" in lib.lfx_last_error(None)


def test_large_padding_and_sector_counts_pass_validation():
    """The reference has no upper bound on convolution_padding / n_blocks (hyper_parameter.hpp:45-53): such handles are
    created (k_extract_rings_big runs them); without a device the only error left is LFX_E_CUDA."""
    import torch

    lib = N.lib()
    for padding, n_blocks in ((20, 6), (5, 100), (64, 300)):
        p = N.Params()
        lib.lfx_default_params(C.byref(p))
        p.padding, p.n_blocks = padding, n_blocks
        h = C.c_void_p()
        rc = lib.lfx_create(C.byref(p), None, C.byref(h))
        assert rc == (N.LFX_OK if torch.cuda.is_available() else N.LFX_E_CUDA), (rc, lib.lfx_last_error(None))
        if rc == N.LFX_OK:
            lib.lfx_destroy(h)


def test_no_gpu_means_loud_failure_not_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from lidar_feature_extraction_b200 import ExtractionError, FeatureExtraction

    with pytest.raises(ExtractionError) as e:
        FeatureExtraction()
    assert e.value.code == N.LFX_E_CUDA
    assert "no CPU fallback" in str(e.value)
    from lidar_feature_extraction_b200 import PipelinedExtraction

    with pytest.raises(ExtractionError) as e:      # the two-handle pipeline has no fallback either
        PipelinedExtraction()
    assert e.value.code == N.LFX_E_CUDA


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "lidar_feature_extraction_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "lfx_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_synthetic_generator_shapes_and_determinism():
    from lidar_feature_extraction_b200 import synth

    for name, (r, w) in {"vlp16": (16, 1800), "hdl32": (32, 2170), "os128": (128, 2048)}.items():
        sp = synth.spec(name)
        a, b = synth.scan_host(sp, 7), synth.scan_host(sp, 7)
        assert a.shape == (r * w, 32) and np.array_equal(a, b)
        x, y, z, _, ring = synth.fields(a)
        assert np.array_equal(ring.reshape(w, r)[0], np.arange(r))  # column-major firing order
        assert not ((x == 0) & (y == 0) & (z == 0)).any()            # convert.py drops (0,0,0) upstream
        az = np.arctan2(y, x).reshape(w, r)
        for k in (0, r - 1):                                         # strictly distinct azimuths per ring
            assert len(np.unique(az[:, k])) == w
    hd = synth.scan_host(synth.spec("hdl64"), 0)
    assert 0 < len(hd) < 64 * 2048                                   # drop-outs make it ragged


def test_cpp_host_mirror_compiles_links_and_fails_loudly(tmp_path):
    """include/lfx.hpp (the C++ side of the boundary) against the built library: parameter mirror, field
    lookup by name, and - on a box without a GPU - LFX_E_CUDA from the constructor instead of any fallback."""
    import shutil
    import subprocess

    import torch

    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "host.cpp"
    src.write_text(r'''
#include <cstdio>
#include "lfx.hpp"
struct float4buf { float p[16]; };
int main() {
  lfx::HyperParameters d;
  lfx::HyperParameters y = lfx::HyperParameters::LaunchYaml();
  if (d.padding != 5 || d.n_blocks != 6 || y.padding != 2 || y.edge_threshold != 50.0) { return 10; }
  std::vector<lfx::PointField> f = {{"x", 0, 7}, {"y", 4, 7}, {"z", 8, 7}, {"intensity", 16, 7}, {"ring", 20, LFX_RING_U16}};
  float buf[8] = {1, 2, 3, 1, 0, 0, 0, 0};
  lfx_cloud_view v = lfx::MakeView(buf, 1, 32, f, true);
  if (!v.has_ring || v.off_ring != 20 || v.off_y != 4 || v.ring_datatype != LFX_RING_U16) { return 11; }
  f.pop_back();
  if (lfx::MakeView(buf, 1, 32, f, true).has_ring) { return 12; }   // RingIsAvailable, ring.cpp:36-44
  try {
    lfx::FeatureExtraction fe(d, 0);
    lfx_scan_output out = fe.Extract(v);                                // a GPU is present: one sparse ring, no features
    std::printf("gpu n_edge=%u n_surface=%u\n", out.n_edge, out.n_surface);
    if (out.n_edge != 0 || out.n_surface != 0) { return 13; }
    lfx::Pipeline pipe(d, 0);                                           // two handles in turn, oldest batch first
    float4buf e, s;
    pipe.Submit({v}); pipe.Submit({v, v});
    const lfx::Pipeline::Output o1 = pipe.Collect(e.p, 4, s.p, 4), o2 = pipe.Collect(e.p, 4, s.p, 4);
    return (o1.counts.size() == 2 && o2.counts.size() == 4 && o2.n_edge == 0 && pipe.InFlight() == 0) ? 0 : 15;
  } catch (const lfx::Error & e) {
    std::printf("error %d: %s\n", e.code, e.what());
    return e.code == LFX_E_CUDA ? 42 : 14;
  }
}
''')
    exe = tmp_path / "host"
    pkg = os.path.join(root, "lidar_feature_extraction_b200")
    subprocess.run([cxx, "-std=c++17", "-Wall", "-Werror", f"-I{root}/include", str(src), "-o", str(exe),
                    f"-L{pkg}", "-llfx", f"-Wl,-rpath,{pkg}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == (0 if torch.cuda.is_available() else 42), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_host_mirror_extracts_a_scan_on_the_gpu(tmp_path):
    """The success path of include/lfx.hpp on the device: a C++ program (no Python, no ctypes) feeds one VLP-16-shaped
    scan through lfx::FeatureExtraction::Extract and ColoredScan; its clouds and labels must equal the oracle's."""
    import shutil
    import subprocess

    from lidar_feature_extraction_b200 import synth
    from oracle import binding as ob
    from oracle import color_oracle as co

    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cloud = synth.scan_host(synth.spec("vlp16"), frame=5)
    (tmp_path / "scan.bin").write_bytes(cloud.tobytes())
    src = tmp_path / "host.cpp"
    src.write_text(r'''
#include <cstdio>
#include <fstream>
#include <iterator>
#include "lfx.hpp"
static void dump(const char * path, const void * p, size_t n) { std::ofstream f(path, std::ios::binary); f.write(static_cast<const char *>(p), n); }
int main(int argc, char ** argv) {
  std::ifstream in(argv[1], std::ios::binary);
  std::vector<char> bytes((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  const std::vector<lfx::PointField> f = {{"x", 0, 7}, {"y", 4, 7}, {"z", 8, 7}, {"intensity", 16, 7}, {"ring", 20, LFX_RING_U16}};
  try {
    lfx_options opt;
    std::memset(&opt, 0, sizeof(opt));
    opt.want_sorted_src = 1;                                            // colored_scan needs the index map
    lfx::FeatureExtraction fe(lfx::HyperParameters(), opt);
    const lfx_cloud_view v = lfx::MakeView(bytes.data(), (uint32_t)(bytes.size() / 32), 32, f, true);
    const lfx_scan_output out = fe.Extract(v);
    dump(argv[2], out.edge_xyz, 16 * (size_t)out.n_edge);
    dump(argv[3], out.surface_xyz, 16 * (size_t)out.n_surface);
    dump(argv[4], out.labels, out.n_points);
    const std::vector<uint8_t> colored = fe.ColoredScan(0);
    dump(argv[5], colored.data(), colored.size());
    std::printf("%u %u %u\n", out.n_points, out.n_edge, out.n_surface);
    return 0;
  } catch (const lfx::Error & e) {
    std::printf("error %d: %s\n", e.code, e.what());
    return 14;
  }
}
''')
    exe = tmp_path / "host"
    pkg = os.path.join(root, "lidar_feature_extraction_b200")
    subprocess.run([cxx, "-std=c++17", "-Wall", "-Werror", f"-I{root}/include", str(src), "-o", str(exe),
                    f"-L{pkg}", "-llfx", f"-Wl,-rpath,{pkg}"], check=True)
    outs = [str(tmp_path / n) for n in ("edge.bin", "surface.bin", "labels.bin", "colored.bin")]
    r = subprocess.run([str(exe), str(tmp_path / "scan.bin")] + outs, capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    ref = ob.Oracle().extract_scan(cloud, ob.default_params())
    x, y, z, _, _ = (np.ascontiguousarray(a) for a in synth.fields(cloud))
    src_idx = ref.sorted_src
    for path, idx in ((outs[0], ref.edge_idx), (outs[1], ref.surface_idx)):
        got = np.fromfile(path, np.float32).reshape(-1, 4)
        want = np.stack([x[src_idx[idx]], y[src_idx[idx]], z[src_idx[idx]], np.ones(len(idx), np.float32)], axis=1)
        assert np.array_equal(got, want)
    assert np.array_equal(np.fromfile(outs[2], np.uint8), ref.labels)
    assert np.array_equal(np.fromfile(outs[3], np.uint8).reshape(-1, 32), co.colored_scan(x, y, z, ref))


@pytest.mark.gpu
def test_cpp_host_mirror_classes_run_on_the_gpu(tmp_path):
    """Every class of include/lfx.hpp on the device, from a C++ program: lfx::Pipeline (two handles in turn),
    lfx::PointTypeConverter -> ExtractBatch (the deployed chain), lfx::MapBuilder, lfx::LoamProblem and a one-rank
    lfx::Shard. The program prints what it got; the same calls through the Python mirrors must give the same numbers."""
    import shutil
    import subprocess

    from lidar_feature_extraction_b200 import (FeatureExtraction, LoamProblem, MapBuilder, default_params, make_pose, synth)

    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    clouds = [synth.scan_host(synth.spec("vlp16"), f) for f in (3, 4)]
    for k, c in enumerate(clouds):
        (tmp_path / f"scan{k}.bin").write_bytes(c.tobytes())
    src = tmp_path / "host.cpp"
    src.write_text(r'''
#include <cstdio>
#include <fstream>
#include <iterator>
#include "lfx.hpp"
static std::vector<char> slurp(const char * p) { std::ifstream in(p, std::ios::binary); return std::vector<char>((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>()); }
int main(int, char ** argv) {
  try {
    const std::vector<char> a = slurp(argv[1]), b = slurp(argv[2]);
    const std::vector<lfx::PointField> f = {{"x", 0, 7}, {"y", 4, 7}, {"z", 8, 7}, {"intensity", 16, 7}, {"ring", 20, LFX_RING_U16}};
    const lfx_cloud_view va = lfx::MakeView(a.data(), (uint32_t)(a.size() / 32), 32, f, true), vb = lfx::MakeView(b.data(), (uint32_t)(b.size() / 32), 32, f, true);
    // ---- Pipeline: {a}, {a, b}, {b} through two handles, collected oldest first
    lfx::Pipeline pipe;
    std::vector<float> edge(4 * (a.size() + b.size()) / 32), surf(edge.size());
    pipe.Submit({va}); pipe.Submit({va, vb});
    const lfx::Pipeline::Output o1 = pipe.Collect(edge.data(), edge.size() / 4, surf.data(), surf.size() / 4);
    pipe.Submit({vb});
    const lfx::Pipeline::Output o2 = pipe.Collect(edge.data(), edge.size() / 4, surf.data(), surf.size() / 4);
    const lfx::Pipeline::Output o3 = pipe.Collect(edge.data(), edge.size() / 4, surf.data(), surf.size() / 4);
    std::printf("pipeline %u %u | %u %u %u %u | %u %u\n", o1.counts[0], o1.counts[1], o2.counts[0], o2.counts[1], o2.counts[2], o2.counts[3], o3.counts[0], o3.counts[1]);
    // ---- converter -> extraction on the device (the wire layout is a raw layout like any other), then the map
    lfx::FeatureExtraction fe;
    lfx::PointTypeConverter conv(fe);
    const lfx_point_field pf[5] = {{"x", 0, 7, 1}, {"y", 4, 7, 1}, {"z", 8, 7, 1}, {"intensity", 16, 7, 1}, {"ring", 20, 4, 1}};
    std::vector<lfx_raw_cloud> raw(2);
    raw[0] = {a.data(), a.size(), 32, pf, 5, 0, LFX_MEM_HOST};
    raw[1] = {b.data(), b.size(), 32, pf, 5, 0, LFX_MEM_HOST};
    const lfx_convert_result cr = conv.Convert(raw);
    fe.ExtractBatch(conv.Views(cr.n_clouds));
    std::vector<uint32_t> counts(4), offsets(6);
    if (lfx_fetch_counts(fe.handle(), counts.data(), offsets.data()) != LFX_OK) { return 20; }
    std::printf("chain %u %u | %u %u %u %u\n", cr.kept[0], cr.kept[1], counts[0], counts[1], counts[2], counts[3]);
    // ---- one-rank shard over the same batch
    lfx::Shard shard(fe, {}, 0, 1, 2);
    shard.Exchange();
    std::vector<uint32_t> gc; std::vector<uint64_t> go;
    shard.Fetch(gc, go, 2);
    std::printf("shard %u %u %u %u | %llu %llu\n", gc[0], gc[1], gc[2], gc[3], (unsigned long long)go[4], (unsigned long long)go[5]);
    lfx::MapBuilder map(fe);
    const lfx_pose p0{{0, 0, 0}, {0, 0, 0, 1}}, p1{{2.5, 0, 0}, {0, 0, 0.0998334166, 0.9950041653}};
    const std::vector<uint8_t> sel = map.AddBatch({p0, p1});
    const std::vector<float> pts = map.Points();
    double sum = 0;
    for (float v : pts) { sum += v; }
    std::printf("map %d %d %llu %.9e\n", (int)sel[0], (int)sel[1], (unsigned long long)map.Size(), sum);
    // ---- residual rows of scan b's edges against the map built from both frames
    if (lfx_fetch_features(fe.handle(), edge.data(), edge.size() / 4, surf.data(), surf.size() / 4) != LFX_OK) { return 21; }
    lfx::LoamProblem prob(fe, pts.data(), pts.size() / 4, surf.data(), offsets[5], 15);
    std::vector<double> J, r;
    prob.Edge(edge.data() + 4 * offsets[2], counts[2], p1, J, r);
    double sj = 0, sr = 0;
    for (double v : J) { sj += v < 0 ? -v : v; }
    for (double v : r) { sr += v < 0 ? -v : v; }
    std::printf("loam %zu %.9e %.9e\n", r.size(), sj, sr);
    return 0;
  } catch (const lfx::Error & e) {
    std::printf("error %d: %s\n", e.code, e.what());
    return 14;
  }
}
''')
    exe = tmp_path / "host"
    pkg = os.path.join(root, "lidar_feature_extraction_b200")
    subprocess.run([cxx, "-std=c++17", "-Wall", "-Werror", f"-I{root}/include", str(src), "-o", str(exe),
                    f"-L{pkg}", "-llfx", f"-Wl,-rpath,{pkg}"], check=True)
    r = subprocess.run([str(exe), str(tmp_path / "scan0.bin"), str(tmp_path / "scan1.bin")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    got = {line.split()[0]: line.split()[1:] for line in r.stdout.strip().splitlines()}

    q1 = [0.0, 0.0, 0.0998334166, 0.9950041653]
    with FeatureExtraction(default_params(), device=0) as fe:
        out = fe.extract_batch(clouds, fetch_points=False)
        ca, cb = [int(v) for v in out.counts[0]], [int(v) for v in out.counts[1]]
        mb = MapBuilder(fe)
        sel = mb.add_batch([make_pose([0, 0, 0], [0, 0, 0, 1]), make_pose([2.5, 0, 0], q1)])
        pts = mb.points()
        prob = LoamProblem(fe, pts, out.surface_xyz, n_neighbors=15)
        J, res = prob.make_edge(out.scan_edges(1), q1, [2.5, 0, 0])
    assert [int(v) for v in got["pipeline"] if v != "|"] == ca + ca + cb + cb
    assert [int(v) for v in got["chain"] if v != "|"] == [len(clouds[0]), len(clouds[1])] + ca + cb   # no (0,0,0) points: all kept
    assert [int(v) for v in got["shard"] if v != "|"] == ca + cb + [ca[0] + cb[0], ca[1] + cb[1]]
    assert [int(got["map"][0]), int(got["map"][1]), int(got["map"][2])] == [int(sel[0]), int(sel[1]), len(pts)]
    assert abs(float(got["map"][3]) - float(np.asarray(pts, np.float64).sum())) <= 1e-3 * max(1.0, abs(float(np.asarray(pts, np.float64).sum())))
    assert int(got["loam"][0]) == res.size
    assert abs(float(got["loam"][1]) - np.abs(J).sum()) <= 1e-6 * np.abs(J).sum() + 1e-9
    assert abs(float(got["loam"][2]) - np.abs(res).sum()) <= 1e-6 * np.abs(res).sum() + 1e-9


def test_library_is_built_for_sm_100a_only_and_the_hot_kernels_use_the_blackwell_paths():
    """The shipped liblfx.so holds sm_100a code and nothing else (no multi-arch fat binary, no PTX fallback path to
    another architecture), the sector kernel fetches its points with 32-byte loads (LDG.E...256, sm_100 only) and the
    converter stages its tiles with the TMA unit's bulk copy (UBLKCP + mbarrier transaction SYNCS)."""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    lib = N.LIB_PATH
    elfs = subprocess.run([cuobjdump, "-lelf", lib], capture_output=True, text=True, check=True).stdout.split("\n")
    elfs = [e for e in elfs if "ELF file" in e]
    assert elfs and all(e.strip().endswith(".sm_100a.cubin") for e in elfs), elfs

    def sass(mangled):
        return subprocess.run([cuobjdump, "-sass", "-fun", mangled, lib], capture_output=True, text=True).stdout

    sector = sass("_ZN4lfxk17k_extract_sectorsILi5ELi11ELb0ELb0EEEvNS_10SectorArgsE")
    assert sector.count("LDG.E.NA.ENL2.256") >= 11, "one 32-byte load per window position of a lane"
    assert "DFMA" in sector and "SHFL" in sector
    conv = sass("_ZN4lfxk9k_convertILb1EEEvNS_8ConvArgsE")
    assert "UBLKCP" in conv and "SYNCS" in conv
