"""Linear-solver plugins: the registry of pyoomph's ``GenericLinearSystemSolver`` and a DEVICE-RESIDENT hand-off next to it (SURVEY N-d).

Mirrors /root/reference/pyoomph/solvers/generic.py:64-118: ``GenericLinearSystemSolver`` with ``register_solver`` / ``factory_solver``
and the ``solve_serial(op_flag, n, nnz, nrhs, values, rowind, colptr, b, ldb, transpose)`` contract through which oomph-lib hands HOST arrays
to SuperLU / MUMPS / Pardiso (src/pybind/solver.cpp:104-140).  The sparse direct solve stays the reference's own (BASELINE north_star) --
what this module adds is the other direction: ``DeviceLinearSystemSolver`` receives DEVICE pointers of the matrix the assembly kernels
just wrote (values and residual never leave HBM, the pattern is uploaded once) and returns the Newton correction on the device, where
``B200Assembly.newton_step_on_device`` applies it.  Per Newton iteration the host link then carries nothing but scalars: the
3.1 GB device-to-host copy of the end-to-end path at BASELINE config 2 (61.9 ms of 64) disappears.

``TorchKrylovSolver`` (idname "torch_krylov") is the plugin the tests use: Jacobi-preconditioned BiCGStab on a zero-copy
``torch.sparse_csr_tensor`` view of the engine's buffers (cuSPARSE SpMV through PyTorch: library plumbing, not a product kernel).
It stands where cuDSS / AmgX / a user's own solver would be registered; it is not part of the timed assembly path.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Dict, Optional, Type

import numpy as np


class GenericLinearSystemSolver:
    """pyoomph/solvers/generic.py:64 -- host-array solver plugins (interface kept so that pyoomph's own plugins register unchanged)."""
    _registered_solvers: Dict[str, Type["GenericLinearSystemSolver"]] = {}
    idname: str = ""

    def __init__(self, problem=None):
        self.problem = problem

    def setup_solver(self) -> None:
        pass

    def solve_serial(self, op_flag: int, n: int, nnz: int, nrhs: int, values, rowind, colptr, b, ldb: int, transpose: int) -> int:
        raise NotImplementedError("You need to specialise the function 'solve_serial'")

    def distributed_possible(self) -> bool:
        return False

    def set_num_threads(self, nthreads: Optional[int]) -> None:
        pass

    @classmethod
    def register_solver(cls, *, override: bool = False) -> Callable:
        def decorator(subclass):
            name = subclass.idname
            if not name:
                raise RuntimeError("solver class needs an idname")
            if name in cls._registered_solvers and not override:
                raise RuntimeError("You tried to register the solver " + name + ", but there is already one defined. Please add override=True "
                                   "to the arguments of @GenericLinearSystemSolver.register_solver(override=True)")
            cls._registered_solvers[name] = subclass
            return subclass
        return decorator

    @staticmethod
    def factory_solver(name: str, problem=None) -> "GenericLinearSystemSolver":
        if name in GenericLinearSystemSolver._registered_solvers:
            return GenericLinearSystemSolver._registered_solvers[name](problem)
        raise RuntimeError("Unknown Linear Algebra solver: '" + name + "'. Following are defined (and included): " +
                           str(list(GenericLinearSystemSolver._registered_solvers.keys())))


class DeviceLinearSystemSolver(GenericLinearSystemSolver):
    """A solver that works on the device-resident CSR matrix.  Specialise ``solve_device``."""

    def solve_device(self, n: int, nnz: int, row_start_ptr: int, col_index_ptr: int, values_ptr: int, rhs_ptr: int, device: int):
        """Solve A x = rhs for the CSR matrix (int32 row_start[n+1], int32 col_index[nnz], float64 values[nnz]) and the right-hand side
        (float64[n]) at the given DEVICE addresses; returns an object with ``data_ptr()`` (a device float64[n]: torch tensor or similar)
        that stays alive until the caller has consumed it, plus a dict of solver statistics."""
        raise NotImplementedError("You need to specialise the function 'solve_device'")


class _DeviceArray:
    """zero-copy typed view of engine-owned device memory (``__cuda_array_interface__``)"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


@GenericLinearSystemSolver.register_solver()
class TorchKrylovSolver(DeviceLinearSystemSolver):
    """Jacobi-preconditioned BiCGStab (van der Vorst) on the device; float64 throughout."""
    idname = "torch_krylov"

    def __init__(self, problem=None, rtol: float = 1e-12, max_iter: int = 5000):
        super().__init__(problem)
        self.rtol, self.max_iter = rtol, max_iter

    def solve_device(self, n, nnz, row_start_ptr, col_index_ptr, values_ptr, rhs_ptr, device):
        import torch
        dev = torch.device("cuda", device)
        crow = torch.as_tensor(_DeviceArray(row_start_ptr, n + 1, "<i4"), device=dev)
        col = torch.as_tensor(_DeviceArray(col_index_ptr, nnz, "<i4"), device=dev)
        val = torch.as_tensor(_DeviceArray(values_ptr, nnz, "<f8"), device=dev)
        b = torch.as_tensor(_DeviceArray(rhs_ptr, n, "<f8"), device=dev)
        A = torch.sparse_csr_tensor(crow, col, val, size=(n, n), device=dev)      # views: no copy of the 3 GB value array
        rows = torch.repeat_interleave(torch.arange(n, device=dev), (crow[1:] - crow[:-1]).to(torch.int64))
        diag = torch.zeros(n, dtype=torch.float64, device=dev)
        on_diag = rows == col.to(torch.int64)
        diag.index_add_(0, rows[on_diag], val[on_diag])
        minv = torch.where(diag != 0, 1.0 / diag, torch.ones_like(diag))
        x = torch.zeros(n, dtype=torch.float64, device=dev)
        r = b.clone()
        r0 = r.clone()
        bnorm = float(torch.linalg.vector_norm(b))
        if bnorm == 0.0:
            return x, {"iterations": 0, "relative_residual": 0.0}
        rho = alpha = omega = 1.0
        v = torch.zeros_like(x)
        p = torch.zeros_like(x)
        it, rel = 0, 1.0
        for it in range(1, self.max_iter + 1):
            rho_new = float(torch.dot(r0, r))
            if rho_new == 0.0:
                break
            beta = (rho_new / rho) * (alpha / omega)
            p = r + beta * (p - omega * v)
            ph = minv * p
            v = torch.mv(A, ph)
            alpha = rho_new / float(torch.dot(r0, v))
            s = r - alpha * v
            if float(torch.linalg.vector_norm(s)) <= self.rtol * bnorm:
                x = x + alpha * ph
                rel = float(torch.linalg.vector_norm(s)) / bnorm
                break
            sh = minv * s
            t = torch.mv(A, sh)
            omega = float(torch.dot(t, s)) / float(torch.dot(t, t))
            x = x + alpha * ph + omega * sh
            r = s - omega * t
            rho = rho_new
            rel = float(torch.linalg.vector_norm(r)) / bnorm
            if rel <= self.rtol:
                break
        return x, {"iterations": it, "relative_residual": rel}


def newton_step_on_device(asm, solver: DeviceLinearSystemSolver, residual: str = ""):
    """One Newton iteration without moving the matrix: assemble R, J on the device (one launch), hand the device CSR to ``solver``,
    apply  dofs -= dx  on the device (pb2_problem_update_dofs_device).  Returns (max |R| before the step, solver statistics); the only
    device-to-host traffic is that scalar.  The current dofs must have been set with ``asm.set_dofs`` (they live in the engine's
    device dof vector from then on)."""
    import torch
    lib = asm.lib
    asm.assemble(flag=1, residual=residual)
    rs, ci = ctypes.c_void_p(), ctypes.c_void_p()
    if lib.pb2_problem_device_pattern(asm.prob, ctypes.byref(rs), ctypes.byref(ci)) != 0:
        raise RuntimeError("pyoomph_b200: " + lib.pb2_last_error().decode())
    r_ptr, j_ptr, _ = asm.device_outputs()
    dev = asm._device
    res = torch.as_tensor(_DeviceArray(r_ptr, asm.n_dof, "<f8"), device=torch.device("cuda", dev))
    torch.cuda.synchronize(dev)
    rmax = float(res.abs().max()) if asm.n_dof else 0.0
    dx, stats = solver.solve_device(asm.n_dof, asm.nnz, rs.value, ci.value, j_ptr, r_ptr, dev)
    torch.cuda.synchronize(dev)
    if lib.pb2_problem_update_dofs_device(asm.prob, ctypes.c_void_p(dx.data_ptr()), ctypes.c_double(-1.0), None) != 0:
        raise RuntimeError("pyoomph_b200: " + lib.pb2_last_error().decode())
    return rmax, stats
