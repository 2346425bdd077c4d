#!/bin/bash
# compute-sanitizer memcheck over small cases of every kernel family (asynchronous map prefetch with its covering words, child
# problems, constraints, point / integral kernels)
out=gpurun_out/${1:-r02san}; mkdir -p $out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
  -k "(residual_jacobian_mass_parity and (poisson-5 or heat3d-3-0.0 or ns-12 or poisson_tri-9 or poisson_tet-3-0.0 or ns_axi_swirl-7)) or (interface_element_classes_parity and (robin_if-6 or nitsche_face-5)) or hanging_nodes_parity and poisson_hang or integral_expressions_parity and ns_obs" > $out/memcheck.log 2>&1
echo "rc=$?" >> $out/memcheck.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Invalid|out of bounds" $out/memcheck.log | tail -12
