#!/bin/bash
# 2-GPU box, final tree: the multi-GPU tests and the driver's N=2 bench call (both arms)
tag=${1:-r02ao}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi -L > $out/host.txt
( time timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q -s ) > $out/pytest_multi.log 2>&1
echo "pytest rc=$?" >> $out/pytest_multi.log
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 ) > $out/bench_reference_n2.json 2> $out/bench_reference_n2.err
echo "reference N=2 rc=$?" >> $out/runs.log
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 ) > $out/bench_n2.json 2> $out/bench_n2.err
echo "N=2 rc=$?" >> $out/runs.log
tail -3 $out/pytest_multi.log; cat $out/runs.log; cut -c1-300 $out/bench_reference_n2.json
tail -1 $out/bench_n2.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'ms', round(d['ms_per_step'], 4), 'value', round(d['value'] / 1e6, 1), 'M el/s  e2e ms', round(d['e2e']['ms_per_step'], 1), 'parity', d.get('multi_gpu_parity'))
"
