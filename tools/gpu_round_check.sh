# One GPU-box call that produces the round's evidence under gpurun_out/ (tag = $1, default r1k):
# parity tests, the bench line + reference arm, the other workloads, the converter row, launch list and
# ncu --set full captures of the two dominant kernels.
tag=${1:-r1k}
set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err
rm -f gpurun_out/${tag}_bench_other.json
for w in hdl32x1000 hdl64x256 vlp16x6250; do timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e >> gpurun_out/${tag}_bench_other.json 2>>gpurun_out/${tag}_bench.err; done
timeout 300 python tools/bench_convert.py --scans 1250 > gpurun_out/${tag}_bench_convert.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extract_sectors -s 13 -c 1 -o gpurun_out/${tag}_sector -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_convert -s 3 -c 1 -o gpurun_out/${tag}_convert -f python tools/bench_convert.py --scans 1250 --steps 1 > /dev/null 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
ls -la gpurun_out
cat gpurun_out/${tag}_pytest.log gpurun_out/${tag}_smoke.log gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench_other.json gpurun_out/${tag}_bench_convert.json
