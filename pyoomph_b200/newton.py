"""The callers of the assembly path: Newton's method and one implicit time step, as oomph-lib's ``Problem::newton_solve`` /
``Problem::unsteady_newton_solve`` drive them (oomph-lib problem.cc; pyoomph ``Problem.solve`` / ``Problem.run``,
/root/reference/pyoomph/generic/problem.py) -- SURVEY section 8(f): "the callers either side of the path".

Works with every assembler of this package that offers ``set_dofs`` / ``assemble`` / ``fetch`` (``B200Assembly``, a parent with child
element classes, ``HangingNodeAssembly``).  Two linear-solver routes:

* host: the CSR values travel to the host and SuperLU (scipy, the reference's default direct solver) solves -- what pyoomph does today;
* device: a ``DeviceLinearSystemSolver`` works on the device-resident matrix (``solvers.newton_step_on_device``), nothing but scalars
  crosses the host link.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


class NewtonConvergenceError(RuntimeError):
    """oomph-lib's NewtonSolverError: the residual did not fall below the tolerance within max_iter iterations (or grew beyond
    max_residual)"""


def newton_solve(asm, dofs: np.ndarray, *, tol: float = 1e-8, max_iter: int = 10, max_residual: float = 1e10, residual: str = "",
                 device_solver=None) -> Tuple[np.ndarray, List[float]]:
    """U <- U - J(U)^-1 R(U) until max|R| < tol (Problem::newton_solve: Newton_solver_tolerance 1e-8, Max_newton_iterations 10,
    Max_residuals 1e10).  Returns the converged dof vector and the history of max|R|.  The history levels, the time weights and the
    parameters of `asm` are used as they are (set them with set_nodal_values / set_unsteady / set_parameters)."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.linalg import splu
    U = np.array(dofs, dtype=np.float64, copy=True)
    history: List[float] = []
    if device_solver is not None:
        asm.set_dofs(U)
        for _ in range(max_iter + 1):
            rmax, _stats = asm.newton_step_on_device(device_solver, residual=residual)       # assemble, solve, update: all on the device
            history.append(rmax)
            if rmax < tol:
                # the step just applied belongs to a converged residual: harmless (its size is below the tolerance of the linear solve)
                return asm.fetch_dofs(), history
            if not np.isfinite(rmax) or rmax > max_residual:
                break
        raise NewtonConvergenceError("Newton's method did not converge: max|R| = %s" % history)
    n = asm.n_dof
    for _ in range(max_iter + 1):
        asm.set_dofs(U)
        asm.assemble(flag=1, residual=residual)
        r, jac, _ = asm.fetch(True, False)
        history.append(float(np.abs(r).max()) if n else 0.0)
        if history[-1] < tol:
            return U, history
        if not np.isfinite(history[-1]) or history[-1] > max_residual:
            break
        U = U - splu(csr_matrix((jac, asm.indices, asm.indptr), shape=(n, n)).tocsc()).solve(r)
    raise NewtonConvergenceError("Newton's method did not converge: max|R| = %s" % history)


def unsteady_newton_solve(asm, dofs: np.ndarray, t: float, dt: float, dtprev: Optional[float], unsteady_steps_done: int, **newton_kw):
    """One implicit step from t to t + dt (Problem::unsteady_newton_solve): shift the history levels on the device
    (``shift_time_values``), set the time weights (BDF2 with the BDF1-degraded first step, src/elements.cpp:4603-4626), solve.  The dof
    vector of the previous step is the initial guess.  Returns (dofs at t + dt, Newton history)."""
    asm.set_dofs(np.asarray(dofs, dtype=np.float64))
    asm.shift_time_values()
    asm.set_unsteady(t + dt, dt, dt if dtprev is None else dtprev, unsteady_steps_done)
    return newton_solve(asm, dofs, **newton_kw)
