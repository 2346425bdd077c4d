#!/bin/bash
# GPU-box profiling pass (B200_PROFILING.md recipe): launch list, one full capture of the assembly kernel, per-role cycle timing
# usage (under gpurun): bash tools/profile.sh <tag> [workload] [n]
tag=${1:-r01}
wl=${2:-ns_cavity}
n=${3:-0}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload $wl --n $n"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $B > gpurun_out/${tag}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pb2_ -s 4 -c 1 -f -o gpurun_out/${tag}_prof $B > gpurun_out/${tag}_prof.log 2>&1
PB2_TIMING=1 timeout 600 $B 2>&1 | grep -E "pb2 timing|metric" > gpurun_out/${tag}_timing.log
