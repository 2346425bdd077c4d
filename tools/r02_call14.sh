#!/bin/bash
# GPU call: coupled free-surface Newton test (new), whole GPU suite, smoke, default bench with all extra workloads
tag=${1:-r02o}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "coupled_free_surface" ) > $out/pytest_new.log 2>&1
echo "rc=$?" >> $out/pytest_new.log
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err
tail -30 $out/pytest_new.log | cut -c1-600; grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -3 $out/smoke.log; tail -5 $out/bench.err
python -c "
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[0])
print('ns ms', d['ms_per_step'])
for e in d['extra_workloads']: print(e['workload'][:70], '| ms', round(e['ms_per_step'],3), '| Mel/s', round(e['value']/1e6,1), '| hbm', round(e['roofline']['frac'],3), '| setup', e['setup_s'], e['pattern_setup_s'])"
