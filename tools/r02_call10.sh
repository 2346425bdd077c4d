#!/bin/bash
# GPU call: point blocks padded to an odd number of 16-byte units (phase-1 STS.128 bank conflicts): full suite + bench with extras, A/B on heat3d
tag=${1:-r02k}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err
for w in heat3d poisson; do
  ( PB2_PB_PAD=0 timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w nopad ms', round(d['ms_per_step'],3))" ) >> $out/ab.log 2>&1
  ( timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w pad   ms', round(d['ms_per_step'],3))" ) >> $out/ab.log 2>&1
done
grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; cat $out/ab.log; python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[0])
print('ns ms', d['ms_per_step'], [ (e['workload'][:12], round(e['ms_per_step'],3)) for e in d['extra_workloads']])"
