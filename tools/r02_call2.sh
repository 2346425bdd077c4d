#!/bin/bash
# Round 2, GPU call 2 (1 GPU): all GPU parity tests (no -x: every case's CSR statistics), smoke, cooperative vs plain launch timing,
# config-3 bench with the sum-factorised columns as default
tag=${1:-r02b}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 900 python -m pytest tests -m gpu -q -s ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra"
for coop in 1 0 1 0; do
  PB2_COOP=$coop timeout 300 $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('coop=$coop ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'])" >> $out/coop.log 2>&1
done
timeout 400 python bench.py --workload heat3d --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/bench_heat3d.json 2> $out/bench_heat3d.err
grep -E "passed|failed" $out/pytest.log | tail -3; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -3 $out/smoke.log; cat $out/coop.log; cut -c1-300 $out/bench_heat3d.json
