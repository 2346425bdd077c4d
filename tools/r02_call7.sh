#!/bin/bash
# GPU call (1 GPU): full GPU suite incl. interface element classes and child problems, smoke, both bench arms as the driver runs them,
# launch lists and full ncu captures of the three bench kernels (configs 2, 3, 1) in their final form.
tag=${1:-r02h}
out=gpurun_out/$tag
mkdir -p $out
rm -f gpurun_out/csr_parity_stats.jsonl
nproc > $out/host.txt; nvidia-smi -L >> $out/host.txt
( time timeout 1200 python -m pytest tests -m gpu -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
mv gpurun_out/csr_parity_stats.jsonl $out/ 2>/dev/null
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke rc=$?" >> $out/smoke.log
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > $out/bench_reference.json 2> $out/bench_reference.err
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
for w in ns_cavity heat3d poisson; do
  B2="python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$w.csv $B2 > /dev/null 2>&1
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:_r0_f1 -s 4 -c 1 -f -o $out/${w}_prof $B2 > $out/${w}_prof.log 2>&1
  timeout 400 ncu --metrics $MET --clock-control none -k regex:_r0_f1 -s 4 -c 1 --csv --log-file $out/${w}_traffic.csv $B2 > /dev/null 2>&1
done
ls -la $out
grep -E "passed|failed" $out/pytest.log | tail -2; grep -E "^FAILED|^ERROR" $out/pytest.log | head; tail -3 $out/smoke.log; cut -c1-300 $out/bench_reference.json; cat $out/bench.json; tail -5 $out/bench.err
