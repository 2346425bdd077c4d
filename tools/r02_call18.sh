#!/bin/bash
# final confirmation after the oracle plugin tag change: the GPU suite and the smoke test
tag=${1:-r02al}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
ls oracle/_build/*.so | wc -l > $out/oracle_plugins.txt
grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -3 $out/smoke.log; cat $out/oracle_plugins.txt
