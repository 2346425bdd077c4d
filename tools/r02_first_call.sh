#!/bin/bash
# First GPU call of a new round (under gpurun, ~5 min): verification pass, launch list + full ncu capture + role timing of the two bench
# kernels, and the opt-in paths that round 1 could only test on the CPU.  usage: bash tools/r02_first_call.sh [tag]
tag=${1:-r02a}
out=gpurun_out/$tag
mkdir -p $out
bash tools/verify.sh $tag/verify > $out/verify.log 2>&1
mv gpurun_out/$tag/verify/* $out/ 2>/dev/null
# config 2: launch list, full capture, role cycles
B2="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_ns.csv $B2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pb2_ns -s 4 -c 1 -f -o $out/ns_prof $B2 > $out/ns_prof.log 2>&1
# config 3 at full size: bench line + full capture on the 64^3 sample
timeout 300 python bench.py --workload heat3d --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $out/bench_heat3d.json 2> $out/bench_heat3d.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pb2_heat3d -s 4 -c 1 -f -o $out/heat3d_prof python bench.py --workload heat3d --n 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $out/heat3d_prof.log 2>&1
# opt-in Morton patch hint on a relabelled mesh: parity against the oracle (CPU-tested only in round 1)
timeout 200 python - > $out/patch_hint.log 2>&1 <<'PY'
import sys; sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np
from problems import compare_matrix, csr_to_sorted, make_oracle, make_problem
from pyoomph_b200.assembly import B200Assembly
pb = make_problem("ns_unsteady", 24, distortion=0.1, unstructured=True)
op = make_oracle(pb)
r_ref, mats = op.assemble(flag=1)
n = pb["dofmap"].n_dof
for hint in (None, "spatial"):
    asm = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, patch_hint=hint)
    for t in range(pb["vals"].shape[0]):
        asm.set_nodal_values(t, pb["vals"][t])
    from problems import TIME
    asm.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
    asm.assemble(flag=1)
    r, jac, _ = asm.fetch()
    err, missing = compare_matrix(csr_to_sorted(n, asm.indptr, asm.indices, jac), csr_to_sorted(n, *mats[0]))
    print("patch_hint", hint, "tiles", asm.num_launches(), "jac err", err, "missing", missing, "res err", np.abs(r - r_ref).max() / np.abs(r_ref).max())
PY
# flagged round-1 experiment: sum-factorised 3D columns (compiled + algebra-checked on the CPU only): parity, then the bench line
( PB2_SUMFAC=1 timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "heat3d or config3" 2>&1 | tail -3
  PB2_SUMFAC=1 timeout 300 python bench.py --workload heat3d --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | cut -c1-220 ) > $out/sumfac.log 2>&1
cat $out/sumfac.log
tail -3 $out/verify.log; cat $out/patch_hint.log | tail -3; cat $out/bench_heat3d.json | cut -c1-200
