#!/bin/bash
# final ncu captures of the three bench kernels (launch lists, full captures, traffic / flop metrics)
tag=${1:-r02v}
out=gpurun_out/$tag
mkdir -p $out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum
for w in ns_cavity heat3d poisson; do
  B2="python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extra"
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$w.csv $B2 > /dev/null 2>&1
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:_r0_f1 -s 4 -c 1 -f -o $out/${w}_prof $B2 > $out/${w}_prof.log 2>&1
  timeout 400 ncu --metrics $MET --clock-control none -k regex:_r0_f1 -s 4 -c 1 --csv --log-file $out/${w}_traffic.csv $B2 > /dev/null 2>&1
done
ls -la $out
