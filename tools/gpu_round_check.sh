# One GPU-box call that produces the round's evidence under gpurun_out/ (tag = $1, default r2):
# parity tests, the bench line (with parity / workloads blocks) + reference arm, the converter row, the localization row,
# launch lists (default workload, ragged workload), ncu --set full captures of the dominant kernels and ncu summaries
# of the small ones, smoke.
tag=${1:-r2}
set -x
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python tools/bench_convert.py --scans 1250 > gpurun_out/${tag}_bench_convert.json 2>> gpurun_out/${tag}_bench.err
timeout 300 python tools/bench_loc.py > gpurun_out/${tag}_bench_loc.json 2>> gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_os128x1250.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-workloads > /dev/null 2>&1
# the same with the conditional node off (LFX_NO_COND=1): every kernel of the general path is launched and listed
LFX_NO_COND=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_os128x1250_nocond.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-workloads > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_hdl64x256.csv python bench.py --workload hdl64x256 --steps 2 --warmup 3 --no-e2e --no-cpu --no-workloads > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extract_sectors -s 7 -c 1 -o gpurun_out/${tag}_sector -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-workloads > /dev/null 2>&1
# the small kernels of the regular path and the ragged path's kernels: one full-set pass each, summaries only
timeout 900 ncu --set full --clock-control none -k regex:'k_pack_fast|k_probe_layout' -s 2 -c 2 -o gpurun_out/${tag}_small -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-workloads > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_ring_hist|k_ring_scatter|k_ring_plan|k_probe_rings|k_general_list' -s 5 -c 5 -o gpurun_out/${tag}_bucketing -f python bench.py --workload hdl64x256 --steps 1 --warmup 1 --no-e2e --no-cpu --no-workloads > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none -k regex:k_extract_sectors -s 10 -c 1 -o gpurun_out/${tag}_sector_indexed -f python bench.py --workload hdl64x256 --steps 1 --warmup 1 --no-e2e --no-cpu --no-workloads > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_convert -s 3 -c 1 -o gpurun_out/${tag}_convert -f python tools/bench_convert.py --scans 1250 --steps 1 > /dev/null 2>&1
# summaries are made here (gpurun brings back at most 64 MiB): text for every capture, only the sector report itself is kept
for r in sector small bucketing sector_indexed convert; do python tools/ncu_summary.py gpurun_out/${tag}_${r}.ncu-rep > gpurun_out/${tag}_${r}_ncu.txt 2>/dev/null; done
python tools/ncu_lines.py gpurun_out/${tag}_sector.ncu-rep 40 > gpurun_out/${tag}_sector_lines.txt 2>/dev/null
python tools/ncu_lines.py gpurun_out/${tag}_convert.ncu-rep 25 > gpurun_out/${tag}_convert_lines.txt 2>/dev/null
rm -f gpurun_out/${tag}_small.ncu-rep gpurun_out/${tag}_bucketing.ncu-rep gpurun_out/${tag}_sector_indexed.ncu-rep gpurun_out/${tag}_convert.ncu-rep
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
cat gpurun_out/${tag}_pytest.log gpurun_out/${tag}_smoke.log; cut -c1-400 gpurun_out/${tag}_bench.json gpurun_out/${tag}_bench_ref.json gpurun_out/${tag}_bench_convert.json gpurun_out/${tag}_bench_loc.json
