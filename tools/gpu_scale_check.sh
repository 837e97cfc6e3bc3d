# N-GPU bench lines of one box (N = $1), our arm and the reference arm, launched like the driver does
N=${1:-8}; tag=${2:-r2s}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 2> gpurun_out/${tag}_bench_n$N.err | tail -1 > gpurun_out/${tag}_bench_n$N.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_bench_n$N.json").read())
print("N=$N value", round(d["value"]/1e9,1), "Gpts/s ms", round(d["ms_per_step"],3), "pipe", round(d["roofline"]["pipeline_frac"],3), "exchange", d["exchange"] and {k:d["exchange"][k] for k in ("ms_mean","ms_max","mechanism","nccl_ranks")}, "e2e", d["e2e"] and round(d["e2e"]["value"]/1e9,2))
PY
