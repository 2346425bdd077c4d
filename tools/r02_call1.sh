#!/bin/bash
# Round 2, first GPU call (1 GPU): GPU parity tests (incl. the bit-exact CSR statistics), smoke, both bench arms (default run with the
# extra workloads), full ncu capture of the config-2 kernel, and the sum-factorised 3D columns prepared in round 1.
tag=${1:-r02a}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
nproc > $out/host.txt; nvidia-smi -L >> $out/host.txt; free -g >> $out/host.txt
( time timeout 900 python -m pytest tests -m gpu -x -q -s ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $out/bench_reference.json 2> $out/bench_reference.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err
B2="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_ns.csv $B2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pb2_ns -s 4 -c 1 -f -o $out/ns_prof $B2 > $out/ns_prof.log 2>&1
( PB2_SUMFAC=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "heat3d or config3" 2>&1 | tail -3
  PB2_SUMFAC=1 timeout 300 python bench.py --workload heat3d --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400 ) > $out/sumfac.log 2>&1
tail -4 $out/pytest.log; tail -3 $out/smoke.log; cut -c1-300 $out/bench_reference.json; cat $out/bench.json; cat $out/bench.err | tail -5; cat $out/sumfac.log
