#!/bin/bash
# schedule sweep on one GPU at the per-rank mesh sizes of the 2/4/8-GPU runs: patch width x patches per unit
out=gpurun_out/${1:-r02sched}
mkdir -p $out
for n in 362 512 724; do
  for pw in 8 4; do
    for up in 4 2 1; do
      PB2_PATCH_WIDTH=$pw PB2_UNIT_PATCHES=$up timeout 200 python bench.py --n $n --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('n $n pw $pw up $up ms', round(d['ms_per_step'], 4), 'Mel/s', round(d['value'] / 1e6, 1), 'tiles', d['config']['tiles'])" >> $out/sweep.log 2>&1
    done
  done
done
cat $out/sweep.log
