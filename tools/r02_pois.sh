#!/bin/bash
# Config 1 (Poisson Q9, 2048^2) with more than one resident block per SM: the kernel is bound by the hand-offs between its roles
# (fp64 pipe 12 %, L1 40 %, DRAM 11 %; barrier stalls 5.8 warps per issue), 63 KB of shared memory per block.
tag=${1:-r02pois}
out=gpurun_out/$tag
mkdir -p $out
run() { # name, env...
  name=$1; shift
  ( env "$@" timeout 600 python bench.py --workload poisson --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$name', 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), d['config'].get('tiles'))" ) >> $out/sweep.log 2>&1
}
run base PB2_PIPE_MINBLOCKS=1
run mb2 PB2_PIPE_MINBLOCKS=2
run mb2_ns64 PB2_PIPE_MINBLOCKS=2 PB2_PIPE_NS=64 PB2_PIPE_NG=32
run mb3_ns64 PB2_PIPE_MINBLOCKS=3 PB2_PIPE_NS=64 PB2_PIPE_NG=32
run mb2_ng32 PB2_PIPE_MINBLOCKS=2 PB2_PIPE_NS=128 PB2_PIPE_NG=32
run mb2_ns96 PB2_PIPE_MINBLOCKS=2 PB2_PIPE_NS=96 PB2_PIPE_NG=32
( PB2_PIPE_MINBLOCKS=2 PB2_PIPE_NS=64 PB2_PIPE_NG=32 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "poisson or robin" 2>&1 | tail -3 ) >> $out/sweep.log 2>&1
cat $out/sweep.log
