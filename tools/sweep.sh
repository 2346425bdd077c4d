#!/bin/bash
# usage: tools_sweep.sh "ENV1=.. ENV2=.." ...   runs bench for each env setting, prints ms_per_step
for cfg in "$@"; do
  out=$(env $cfg timeout 150 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1)
  echo "$cfg => $(echo "$out" | python -c 'import sys,json
try:
    d=json.loads(sys.stdin.read()); print("ms_per_step=%.3f frac=%.3f launches=%d" % (d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["launches_per_step"]))
except Exception as e: print("FAILED", e)')"
done
