#!/bin/bash
tag=${1:-r02y}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "integral_expressions" ) > $out/pytest_new.log 2>&1
echo "rc=$?" >> $out/pytest_new.log
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
tail -30 $out/pytest_new.log | cut -c1-500; grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -3 $out/smoke.log
