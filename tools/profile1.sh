#!/bin/bash
# one full ncu capture of the assembly kernel under the env given as $2..; usage: bash tools/profile1.sh <tag> ENV=..
tag=$1; shift
mkdir -p gpurun_out
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:pb2_ -s 4 -c 1 -f -o gpurun_out/${tag}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_prof.log 2>&1
env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
