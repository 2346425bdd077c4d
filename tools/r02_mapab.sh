#!/bin/bash
out=gpurun_out/${1:-r02r}; mkdir -p $out
for w in ale_freesurface heat3d; do
 for v in 0 1 0 1; do
  ( PB2_MAP_ASYNC=$v timeout 600 python bench.py --workload $w --n $([ $w = heat3d ] && echo 126 || echo 512) --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w async=$v ms', round(d['ms_per_step'],3))" ) >> $out/ab.log 2>&1
 done
done
cat $out/ab.log
