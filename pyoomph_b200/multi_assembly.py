"""Single-request multi-assembly: the host mirror of pyoomph's ``MultiAssembleRequest``
(/root/reference/pyoomph/generic/bifurcation_tools.py:449-531), which the bifurcation trackers (fold, pitchfork, Hopf,
azimuthal) use to get R, J, M, parameter derivatives and Hessian-vector products of several residual contributions from ONE
pass over the elements (``BulkElementBase::get_multi_assembly``, src/elements.cpp:4776-5003, and
``Problem::sparse_assemble_row_or_column_compressed_base_problem``, src/problem.cpp:2056-2284).

On the GPU the "one pass" becomes the smallest set of launches that covers the request: everything asked of one
(contribution, parameter) pair comes from one launch at the highest flag needed (flag 1 also yields the residual, flag 2
also the Jacobian), and the Hessian-vector products of one contribution and vector (d(J.Y)/dU and d(M.Y)/dU together) come
from one launch per vector, the vectors travelling to the device in blocks of ``PB2_MAX_HVEC``.  Results come back in request order, vectors as float64 arrays and matrices as scipy CSR
matrices over the fixed pattern of the assembler, exactly the list ``MultiAssembleRequest.assemble`` returns.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

PB2_MAX_HVEC = 4

_VECTOR_KINDS = ("residuals", "dresiduals_dparameter")
_FLAG = {"residuals": 0, "jacobian": 1, "mass_matrix": 2,
         "dresiduals_dparameter": 0, "djacobian_dparameter": 1, "dmass_matrix_dparameter": 2}


class MultiAssembleRequest:
    """``MultiAssembleRequest(asm).R().J().M().dRdp("mu").dJdU(Y).assemble()``; ``asm`` is a ``B200Assembly``."""

    def __init__(self, assembler):
        self.assembler = assembler
        self._what: List[str] = []
        self._contributions: List[str] = []
        self._parameters: List[str] = []          # one per d*dp entry, in request order
        self._hessian_vectors: List[np.ndarray] = []
        self._hessian_vector_indices: List[int] = []

    def _resolve_hessian_vector_index(self, V) -> int:
        for i, w in enumerate(self._hessian_vectors):
            if V is w:
                return i
        self._hessian_vectors.append(V)
        return len(self._hessian_vectors) - 1

    def _add(self, what: str, contribution: str):
        if contribution not in self.assembler.residual_names:
            raise RuntimeError("unknown residual contribution '%s'" % contribution)
        self._what.append(what)
        self._contributions.append(contribution)
        return self

    def R(self, contribution: str = ""):
        return self._add("residuals", contribution)

    def J(self, contribution: str = ""):
        return self._add("jacobian", contribution)

    def M(self, contribution: str = ""):
        return self._add("mass_matrix", contribution)

    def _param(self, parameter) -> str:
        name = parameter if isinstance(parameter, str) else parameter.get_name()
        if name not in self.assembler.param_names:
            raise RuntimeError("unknown global parameter '%s'" % name)
        return name

    def dRdp(self, parameter, contribution: str = ""):
        self._parameters.append(self._param(parameter))
        return self._add("dresiduals_dparameter", contribution)

    def dJdp(self, parameter, contribution: str = ""):
        self._parameters.append(self._param(parameter))
        return self._add("djacobian_dparameter", contribution)

    def dMdp(self, parameter, contribution: str = ""):
        self._parameters.append(self._param(parameter))
        return self._add("dmass_matrix_dparameter", contribution)

    def _hvp(self, what: str, vector, contribution: str, transposed: bool):
        """transposed: the contraction sum_j H_jik Y_j (flags 4 / 5 of HessianVectorProduct; the reference's multi-assembly passes
        hessian_vector_transposed along, bifurcation_tools.py:505-515 -> src/elements.cpp:4983-4988)"""
        v = np.asarray(vector, dtype=np.float64)
        if v.shape != (self.assembler.n_dof,):
            raise RuntimeError("Hessian vector must have one entry per dof")
        self._hessian_vector_indices.append(self._resolve_hessian_vector_index(vector))
        self.__dict__.setdefault("_hessian_transposed", []).append(bool(transposed))
        return self._add(what, contribution)

    def dJdU(self, vector, contribution: str = "", transposed: bool = False):
        return self._hvp("hessian_vector_product", vector, contribution, transposed)

    def dMdU(self, vector, contribution: str = "", transposed: bool = False):
        return self._hvp("mass_matrix_hessian_vector_product", vector, contribution, transposed)

    # ------------------------------------------------------------------------------------------------
    def assemble(self) -> list:
        from scipy.sparse import csr_matrix
        asm = self.assembler
        n = asm.n_dof
        # 1. group: (contribution, parameter or None) -> highest flag; contribution -> (vector index -> needs mass Hessian)
        rjm: Dict[Tuple[str, object], int] = {}
        hvp: Dict[str, Dict[int, bool]] = {}
        pi = hi = 0
        keys = []
        for what, contrib in zip(self._what, self._contributions):
            if what in _FLAG:
                par = None
                if what.startswith("d"):
                    par = self._parameters[pi]
                    pi += 1
                k = (contrib, par)
                rjm[k] = max(rjm.get(k, 0), _FLAG[what])
                keys.append(k)
            else:
                vi = self._hessian_vector_indices[hi]
                tr = self.__dict__.get("_hessian_transposed", [])[hi]
                hi += 1
                d = hvp.setdefault((contrib, tr), {})
                d[vi] = d.get(vi, False) or what.startswith("mass_matrix")
                keys.append((contrib, tr, vi))
        # 2. launches
        self.launches = 0
        got: Dict[Tuple[str, object], tuple] = {}
        for (contrib, par), flag in rjm.items():
            asm.assemble(flag=flag, residual=contrib, parameter=par)
            self.launches += asm.launch_count()
            got[(contrib, par)] = asm.fetch(want_jacobian=flag >= 1, want_mass=flag >= 2)
        hgot: Dict[Tuple[str, bool, int], tuple] = {}
        for (contrib, tr), vecs in hvp.items():
            idx = sorted(vecs)
            for b in range(0, len(idx), PB2_MAX_HVEC):
                blk = idx[b:b + PB2_MAX_HVEC]
                flag = 2 if any(vecs[i] for i in blk) else 1
                Y = np.stack([np.asarray(self._hessian_vectors[i], dtype=np.float64) for i in blk])
                Jv, Mv = asm.assemble_hessian(Y, flag=flag, residual=contrib, transposed=tr)
                self.launches += asm.launch_count()
                for k, i in enumerate(blk):
                    hgot[(contrib, tr, i)] = (Jv[k], Mv[k])
        # 3. results in request order

        def mat(values):
            return csr_matrix((values, asm.indices, asm.indptr), shape=(n, n))
        out = []
        for what, k in zip(self._what, keys):
            if what in _VECTOR_KINDS:
                out.append(got[k][0])
            elif what in ("jacobian", "djacobian_dparameter"):
                out.append(mat(got[k][1]))
            elif what in ("mass_matrix", "dmass_matrix_dparameter"):
                out.append(mat(got[k][2]))
            elif what == "hessian_vector_product":
                out.append(mat(hgot[k][0]))
            else:
                out.append(mat(hgot[k][1]))
        return out
