#!/bin/bash
# A/B of library builds with the same ABI: time per step (bench.py) and DRAM bytes of the sector kernel (ncu).
# usage: tools/ab_variants.sh <tag> [lib.so ...]   ("default" = the in-tree library)
tag=$1; shift
for lib in "$@"; do
  if [ "$lib" = default ]; then unset LFX_LIB; else export LFX_LIB=$PWD/$lib; fi
  echo "== $lib"
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step', round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()})"
  timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct -k regex:k_extract_sectors -s 13 -c 1 --csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu 2>/dev/null | grep -E "dram__bytes|time_duration|hit_rate" | awk -F'","' '{print "   " $(NF-2), $(NF-1), $NF}'
done
