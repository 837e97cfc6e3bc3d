"""bench.py's contract, as far as a box without a GPU can check it: the reference arm (the reference's own CPU code on
the host cores, no CUDA library involved) prints ONE JSON line with the agreed keys and the same `config` builder as the
GPU arm; the GPU arm's helpers that need no device behave."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_line(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--scans", "8", *extra], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_prints_one_contract_line():
    d = _reference_line("--workload", "vlp16x6250")
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["metric"] == "scan_points_per_sec" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["workload"] == "vlp16x6250" and d["config"]["rings"] == 16 and d["config"]["scans_per_gpu"] == 8
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 0


def test_reference_arm_never_maps_the_cuda_library():
    """The arm must not load liblfx.so (its scans come from libsynth.so): a ratio against an arm that touches the
    product would be void."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--scans', '4', "
            "'--workload', 'vlp16x6250']; runpy.run_path('bench.py', run_name='__main__'); "
            "maps = open('/proc/self/maps').read(); print('LIBLFX' if 'liblfx.so' in maps else 'CLEAN')")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "CLEAN", r.stdout[-500:]


def test_workload_config_is_shared_by_both_arms():
    sys.path.insert(0, ROOT)
    import bench

    cfg = bench.workload_config("os128x1250", "os128", 128, 2048, 1250, 1250 * 128 * 2048, 8)
    assert cfg["workload"] == "os128x1250" and cfg["points_per_gpu"] == 327680000 and "sharding" in cfg and "l2" in cfg
    assert set(bench.WORKLOADS) == {"os128x1250", "hdl32x1000", "hdl64x256", "vlp16x6250"}
