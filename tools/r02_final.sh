#!/bin/bash
# last call of the round on one GPU: what the driver runs at round end (GPU tests, smoke, reference arm, default bench)
tag=${1:-r02final}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 1200 python -m pytest tests -x -q -m gpu ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/bench_reference.json 2> $out/bench_reference.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err
grep -E "passed|failed" $out/pytest.log | tail -2; tail -3 $out/smoke.log; cut -c1-250 $out/bench_reference.json; cut -c1-400 $out/bench.json; tail -4 $out/bench.err
