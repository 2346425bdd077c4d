#!/bin/bash
# GPU call: all GPU tests (new: transposed Hessian products, Hessian tensor, C host program), smoke, default bench (setup timing on)
tag=${1:-r02c}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
( time PB2_SETUP_TIMING=1 timeout 900 python bench.py --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err
grep -E "passed|failed" $out/pytest.log | tail -3; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -2 $out/smoke.log; cat $out/bench.json | cut -c1-1500; grep "pb2 setup" $out/bench.err | head -40
