"""Generates tests/golden/*.json from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).
The reference itself cannot be executed in this environment (no GiNaC), so these are oracle outputs."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from problems import csr_to_sorted, make_oracle, make_problem  # noqa: E402

for kind, N in [("poisson", 5), ("ns_unsteady", 4), ("heat3d", 2), ("ale", 4), ("ns_axi_swirl", 3), ("ale_axi_obs", 3), ("heat3d_obs", 2)]:
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    r, mats = op.assemble(flag=2)
    n = pb["dofmap"].n_dof
    g = {"kind": kind, "N": N, "n_dof": int(n), "res_l1": float(np.abs(r).sum())}
    v = np.cos(np.arange(n))
    for m, key in zip(mats, ("jac", "mass")):
        A = csr_to_sorted(n, *m)
        g[key + "_nnz"] = int(A.nnz)
        g[key + "_l1"] = float(abs(A).sum())
        g[key + "_matvec_l1"] = float(np.abs(A @ v).sum())
    if pb["code"].integral_expressions:
        g["integrals"] = op.evaluate_integral_expressions()
    json.dump(g, open(os.path.join(HERE, "%s_%d.json" % (kind, N)), "w"), indent=1)
    print(g)
