#!/bin/bash
# GPU-box verification pass: GPU parity tests, smoke, both bench arms.  usage (under gpurun): bash tools/verify.sh <tag>
tag=${1:-r01v}
out=gpurun_out/$tag
mkdir -p $out
nproc > $out/host.txt; nvidia-smi -L >> $out/host.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
tail -3 $out/pytest.log; cat $out/smoke.log | tail -2; cat $out/bench_reference.json; cat $out/bench.json
