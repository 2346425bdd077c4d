#!/bin/bash
# 8-GPU box: the driver's scaling sequence (N = 1, 2, 4, 8 back to back, both arms at N=1 only for ours), with the window parity check
tag=${1:-r02scale}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/host.txt
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) > $out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $out/pytest_multi.log
( timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-extra ) > $out/bench_n1.json 2> $out/bench_n1.err; echo "N=1 rc=$?" >> $out/runs.log
for N in 2 4 8; do
  ( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 5 ) > $out/bench_n$N.json 2> $out/bench_n$N.err
  echo "N=$N rc=$?" >> $out/runs.log
done
tail -3 $out/pytest_multi.log; cat $out/runs.log
for N in 1 2 4 8; do tail -1 $out/bench_n$N.json | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read())
    print('N', d['n_gpus'], 'ms', round(d['ms_per_step'], 4), 'value', round(d['value'] / 1e6, 1), 'M el/s  e2e ms', round(d['e2e']['ms_per_step'], 1), 'parity', d.get('multi_gpu_parity'), 'exch', d['config']['exchange_bytes_per_step_max_rank'])
except Exception as e:
    print('unparsable', e)
"; done
