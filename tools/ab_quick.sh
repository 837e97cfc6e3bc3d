#!/bin/bash
# Quick A/B of library builds with the same ABI: ms per step of bench.py on one or more workloads.
# usage: WORKLOADS="os128x1250 hdl32x1000" tools/ab_quick.sh lib.so ...   ("default" = the in-tree library)
for lib in "$@"; do
  if [ "$lib" = default ]; then unset LFX_LIB; else export LFX_LIB=$PWD/$lib; fi
  for w in ${WORKLOADS:-os128x1250}; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', '$w', 'ms_per_step', round(d['ms_per_step'],3), 'Gpts/s', round(d['value']/1e9,2), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
  done
done
