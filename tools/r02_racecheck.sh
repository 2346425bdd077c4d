#!/bin/bash
out=gpurun_out/${1:-r02race}; mkdir -p $out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 1 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "residual_jacobian_mass_parity and (poisson-5 or ns-12-0.0)" > $out/racecheck.log 2>&1
echo "rc=$?" >> $out/racecheck.log
grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed|rc=|hazard" $out/racecheck.log | tail -12
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 1 --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "residual_jacobian_mass_parity and (poisson-5 or ns-12-0.0)" > $out/synccheck.log 2>&1
echo "rc=$?" >> $out/synccheck.log
grep -E "ERROR SUMMARY|passed|failed|rc=" $out/synccheck.log | tail -6
