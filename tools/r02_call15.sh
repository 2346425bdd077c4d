#!/bin/bash
# GPU call: asynchronous map prefetch for odd ndof^2: full suite + default bench (all extras)
tag=${1:-r02q}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err
grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -4 $out/bench.err
python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[0])
print('ns ms', d['ms_per_step'])
for e in d['extra_workloads']: print(e['workload'][:70], '| ms', round(e['ms_per_step'],3), '| Mel/s', round(e['value']/1e6,1), '| hbm', round(e['roofline']['frac'],3))"
