#!/bin/bash
# GPU call: full GPU suite (new: point expressions, device solver, transposed Hessians, tensor, C host), smoke
tag=${1:-r02f}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -3 $out/smoke.log
