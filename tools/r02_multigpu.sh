#!/bin/bash
# Round 2 multi-GPU call (gpurun --gpus N): the 2-GPU parity tests inside pytest, then REPS consecutive torchrun bench runs at N GPUs
# (each with the oracle window check of the distributed matrix), exactly as the driver launches them.
# usage: bash tools/r02_multigpu.sh <tag> <N> <reps>
tag=${1:-r02mg}; N=${2:-2}; REPS=${3:-5}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/host.txt
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) > $out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $out/pytest_multi.log
for i in $(seq 1 $REPS); do
  port=$((29500 + i))
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 ) > $out/bench_n${N}_run$i.json 2> $out/bench_n${N}_run$i.err
  echo "run $i rc=$?" >> $out/runs.log
done
tail -3 $out/pytest_multi.log; cat $out/runs.log
for i in $(seq 1 $REPS); do tail -1 $out/bench_n${N}_run$i.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('ms', round(d['ms_per_step'], 4), 'value', round(d['value'] / 1e6, 1), 'M el/s  e2e ms', round(d['e2e']['ms_per_step'], 1), 'parity', d.get('multi_gpu_parity'), 'exch', d['config']['exchange_bytes_per_step_max_rank'])
except Exception as e:
    print('unparsable', e)
"; done
