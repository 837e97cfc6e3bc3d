# the four informative variants of tools/ab_gather.sh at N = 8 (GPU-minutes are charged N times)
n=8
run() { tag=$1; shift; env "$@" timeout 60 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n --steps 20 --warmup 3 --no-e2e 2>gpurun_out/abg8_$tag.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$tag', 'n=%d' % d['n_gpus'], 'ms/step %.3f' % d['ms_per_step'], 'Gpts/s %.1f' % (d['value'] / 1e9), d.get('exchange'))
"; }
run sync A=1
run overlap LFX_BENCH_GATHER=overlap
run sync_noclk LFX_BENCH_NO_CLOCKS=1
run overlap_r2 LFX_BENCH_GATHER=overlap LFX_RESERVE_SMS=2
