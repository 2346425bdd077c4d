#!/bin/bash
# GPU call: dependency gates -- all GPU tests under a hard timeout, then single-GPU bench at the per-rank mesh sizes and the full size
tag=${1:-r02d}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 900 python -m pytest tests -m gpu -q -x ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
for n in 362 512 724 1024; do
  timeout 200 python bench.py --n $n --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('n $n ms', round(d['ms_per_step'], 4), 'Mel/s', round(d['value'] / 1e6, 1), 'frac', round(d['roofline']['frac'], 4))" >> $out/sizes.log 2>&1
done
timeout 300 python bench.py --workload heat3d --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 >> $out/sizes.log
timeout 300 python bench.py --workload poisson --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 >> $out/sizes.log
grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; cat $out/sizes.log
