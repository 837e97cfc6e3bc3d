set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r1g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1g_bench_ref.json 2>> gpurun_out/r1g_bench.err
for w in hdl32x1000 hdl64x256 vlp16x6250; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e >> gpurun_out/r1g_bench_other.json 2>>gpurun_out/r1g_bench.err; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extract_sectors -s 13 -c 1 -o gpurun_out/r1g_sector -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out
cat gpurun_out/r1g_pytest.log gpurun_out/r1g_bench.json gpurun_out/r1g_bench_other.json
