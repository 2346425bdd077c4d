#!/bin/bash
# kernel experiment under an environment switch: parity of the touched classes, then bench lines.  usage: r02_exp.sh <tag> "<ENV=1 ...>" [workloads]
tag=$1; envs=$2; wls=${3:-"ns_cavity poisson"}
out=gpurun_out/$tag
mkdir -p $out
( env $envs timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "residual_jacobian_mass or full_size_config2 or hessian_vector" 2>&1 | tail -4 ) > $out/parity.log 2>&1
for wl in $wls; do
  env $envs timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4))" >> $out/bench.log 2>&1
  timeout 400 python bench.py --workload $wl --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('$wl baseline', 'ms', round(d['ms_per_step'], 4), 'frac', round(d['roofline']['frac'], 4))" >> $out/bench.log 2>&1
done
cat $out/parity.log; cat $out/bench.log
