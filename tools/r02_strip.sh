#!/bin/bash
# the per-rank mesh of the 8-GPU run (a 128 x 1024 strip) on one GPU: patch shapes x patches per unit
out=gpurun_out/${1:-r02strip}
mkdir -p $out
for pw in 8 8x4 4x8 8x16 16x8; do
  for up in auto 1 2 4; do
    if [ $up = auto ]; then unset PB2_UNIT_PATCHES; else export PB2_UNIT_PATCHES=$up; fi
    PB2_BENCH_NY=1024 PB2_PATCH_WIDTH=$pw timeout 200 python bench.py --n 128 --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('strip 128x1024 pw $pw up $up ms', round(d['ms_per_step'], 4), 'Mel/s', round(d['value'] / 1e6, 1))" >> $out/strip.log 2>&1
  done
done
cat $out/strip.log
