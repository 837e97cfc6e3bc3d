# Evidence of a state whose kernels did not change since the last full tools/gpu_round_check.sh: parity tests,
# bench lines (default, reference arm, other workloads), ncu launch list of the bench command, smoke.
tag=${1:-r1s}
nvidia-smi -L
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
rm -f gpurun_out/${tag}_bench_other.json
for w in hdl32x1000 hdl64x256 vlp16x6250; do timeout 200 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e >> gpurun_out/${tag}_bench_other.json 2>>gpurun_out/${tag}_bench.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
cat gpurun_out/${tag}_pytest.log gpurun_out/${tag}_smoke.log; cut -c1-300 gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench_other.json gpurun_out/${tag}_bench_ref.json
