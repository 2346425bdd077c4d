"""Global rank-3 Hessian tensor: the host mirror of ``pyoomph::SparseRank3Tensor`` (/root/reference/src/hessian_tensor.hpp:31-82,
src/hessian_tensor.cpp:28-96) and of ``Problem::assemble_hessian_tensor`` (src/problem.cpp:1530-1560) for the GPU path.

The reference fills the tensor element by element from the ``ndof^3`` buffer of ``HessianVectorProduct<i>`` with flag 3
(src/elements.cpp:5243-5247): ``T(iG, jG, kG) += H_e[i][k][j]`` for entries with ``fabs(value) > 0``.  On the GPU the same numbers come
from the Hessian-vector kernels: with a unit vector ``Y = e_k`` the matrix ``N = d(J.Y)/dU`` has the entries ``N[i, j] = sum_j' H_{i j' j} Y_j'
= H_{i k j}``, i.e. one slice ``T(:, :, k)`` of the tensor per launch -- no ``ndof^3`` buffer per element, the fixed CSR pattern is reused
for every slice.  That is ``n_dof`` launches: like the reference's own tensor path (a ``std::map`` per row), a small-problem tool; the
bifurcation trackers use the vector products directly (``MultiAssembleRequest.dJdU``).
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np


class SparseRank3Tensor:
    """``[i](j,k) -> value`` with the reference's interface: accumulate, finalize_for_vector_product, right_vector_mult, get_entries."""

    def __init__(self, size: int, symmetric: bool = False):
        if symmetric:
            # the reference's symmetric branch of right_vector_mult indexes the result with a k index (src/hessian_tensor.cpp:89):
            # nothing to be faithful to; the trackers build the tensor with symmetric = false (pyoomph/generic/problem.py)
            raise NotImplementedError("symmetric storage of SparseRank3Tensor is not mirrored")
        self._size = int(size)
        self._i: List[np.ndarray] = []
        self._j: List[np.ndarray] = []
        self._k: List[np.ndarray] = []
        self._v: List[np.ndarray] = []
        self._final = None

    def size(self) -> int:
        return self._size

    def accumulate(self, i, j, k, val) -> None:
        """scalar or array arguments; repeated (i, j, k) add up (src/hessian_tensor.hpp:58-74)"""
        self._i.append(np.atleast_1d(np.asarray(i, dtype=np.int64)))
        self._j.append(np.atleast_1d(np.asarray(j, dtype=np.int64)))
        self._k.append(np.atleast_1d(np.asarray(k, dtype=np.int64)))
        self._v.append(np.atleast_1d(np.asarray(val, dtype=np.float64)))
        self._final = None

    def _sorted(self):
        if not self._i:
            z = np.zeros(0, dtype=np.int64)
            return z, z, z, np.zeros(0)
        i, j, k, v = (np.concatenate(a) for a in (self._i, self._j, self._k, self._v))
        order = np.lexsort((k, j, i))
        i, j, k, v = i[order], j[order], k[order], v[order]
        # merge repeated (i, j, k): std::map semantics, contributions added in insertion order
        new = np.ones(i.size, dtype=bool)
        new[1:] = (i[1:] != i[:-1]) | (j[1:] != j[:-1]) | (k[1:] != k[:-1])
        seg = np.cumsum(new) - 1
        vs = np.zeros(int(seg[-1]) + 1 if seg.size else 0)
        np.add.at(vs, seg, v)
        return i[new], j[new], k[new], vs

    def get_entries(self) -> List[Tuple[int, int, int, float]]:
        i, j, k, v = self._sorted()
        return list(zip(i.tolist(), j.tolist(), k.tolist(), v.tolist()))

    def finalize_for_vector_product(self):
        """(matrix_col_index, matrix_row_start) of the CSR matrix M_ij = T_ijk v_k (src/hessian_tensor.cpp:33-60)"""
        i, j, k, v = self._sorted()
        newc = np.ones(i.size, dtype=bool)
        newc[1:] = (i[1:] != i[:-1]) | (j[1:] != j[:-1])
        ent = np.cumsum(newc) - 1                              # CSR entry of every (i, j, k)
        col_index = j[newc].astype(np.int32)
        rows = i[newc]
        row_start = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=self._size))]).astype(np.int32)
        self._final = (ent, k, v, col_index.size)
        return col_index, row_start

    def right_vector_mult(self, vec) -> np.ndarray:
        """values of M_ij = sum_k T_ijk v_k on the pattern of finalize_for_vector_product (src/hessian_tensor.cpp:67-82)"""
        vec = np.asarray(vec, dtype=np.float64)
        if vec.shape != (self._size,):
            raise RuntimeError("Mismatch in tensor and vector size")
        if self._final is None:
            raise RuntimeError("Call finalize_for_vector_product first")
        ent, k, v, n = self._final
        vals = np.zeros(n)
        np.add.at(vals, ent, vec[k] * v)
        return vals


def assemble_hessian_tensor(asm, symmetric: bool = False, residual: str = "", block: int = 8) -> SparseRank3Tensor:
    """Problem::assemble_hessian_tensor for a B200Assembly: slice k of the tensor = d(J.e_k)/dU from one Hessian-vector launch."""
    n = asm.n_dof
    T = SparseRank3Tensor(n, symmetric)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(asm.indptr))
    cols = asm.indices.astype(np.int64)
    for k0 in range(0, n, block):
        ks = np.arange(k0, min(n, k0 + block))
        Y = np.zeros((ks.size, n))
        Y[np.arange(ks.size), ks] = 1.0
        Jv, _ = asm.assemble_hessian(Y, flag=1, residual=residual)
        for kk, vals in zip(ks, Jv):
            nz = np.abs(vals) > 0.0                            # Numerical_zero_for_sparse_assembly (src/problem.cpp:1548)
            T.accumulate(rows[nz], cols[nz], np.full(int(nz.sum()), kk, dtype=np.int64), vals[nz])
    return T
