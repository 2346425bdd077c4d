"""Hanging nodes: the reference distributes every element contribution of a hanging value to the equations of its
master values inside the generated routine (`*_CONTINUOUS_SPACE` / `*_HANG` macros, /root/reference/src/jitbridge_hang.h:104-184, with
the local numbering of `RefineableElement::assign_hanging_local_eqn_numbers`, oomph-lib refineable_elements.cc:312-470, and
`BulkElementBase::fill_hang_info_with_equations`, src/elements.cpp:812-1160).

On the GPU the element kernels stay as they are.  Every hanging value gets a VIRTUAL equation behind the real ones; the kernels assemble
the extended system  J_ext, R_ext  over  n_ext = n_dof + n_virtual  equations with their usual position maps, and one reduction pass on
the device applies the constraint  u_ext = P u  (P: n_ext x n_dof, identity on the real equations, master weights in the virtual rows):

    J = P^T J_ext P,    R = P^T R_ext

Only entries in virtual rows or columns move: every target entry (a real row and column) sums its weighted sources in a fixed order
(`pb2_problem_set_constraints` / the reduction kernels in pb2_core.cu: no atomics, deterministic); targets the element pattern does not
hold enter it as extra pattern entries.  Afterwards the virtual rows and columns are cleared and get a unit diagonal, so the device holds
J (+) I over n_ext equations -- a device solver can use it as it is, `fetch()` returns the n_dof x n_dof block."""
from __future__ import annotations

import ctypes
import dataclasses
from typing import Optional

import numpy as np

from .assembly import B200Assembly, _check, c_double_p, c_int_p
from .codegen import FiniteElementCode
from .meshes import DofMap, HangingNodes


@dataclasses.dataclass
class ExtendedNumbering:
    dofmap: DofMap                 # node_eqn with virtual equations on the hanging values, n_dof = n_ext
    n_real: int
    P_rows: np.ndarray             # COO of the virtual rows of P: virtual equation, master equation (>= 0), weight
    P_cols: np.ndarray
    P_vals: np.ndarray

    @property
    def n_ext(self) -> int:
        return self.dofmap.n_dof

    def prolongation(self):
        """P as scipy CSR (n_ext x n_real)"""
        from scipy.sparse import coo_matrix
        n = self.n_real
        rows = np.concatenate([np.arange(n), self.P_rows])
        cols = np.concatenate([np.arange(n), self.P_cols])
        vals = np.concatenate([np.ones(n), self.P_vals])
        return coo_matrix((vals, (rows, cols)), shape=(self.n_ext, n)).tocsr()


def extend_numbering(code: FiniteElementCode, dofmap: DofMap, hanging: HangingNodes) -> ExtendedNumbering:
    """virtual equations n_dof, n_dof+1, ... for the hanging values (node order, then value index); values whose masters are all pinned
    stay without an equation (they contribute to nothing)"""
    node_eqn = dofmap.node_eqn.copy()
    pos_eqn = None if dofmap.pos_eqn is None else dofmap.pos_eqn.copy()
    n = dofmap.n_dof
    nval = node_eqn.shape[1]
    todo = []
    for f in code.nodal_fields():
        for node, (masters, weights) in hanging.of_space(f.space).items():
            todo.append((int(node), f.index, masters, weights))
    if pos_eqn is not None:
        # moving mesh: the position dofs of a geometrically hanging node (C2 hang info) hang on its masters' position dofs; they are
        # addressed here as value indices nval, nval + 1, ... of the node
        for node, (masters, weights) in hanging.C2.items():
            for d in range(pos_eqn.shape[1]):
                todo.append((int(node), nval + d, masters, weights))
    todo.sort(key=lambda t: (t[0], t[1]))
    all_eqn = node_eqn if pos_eqn is None else np.concatenate([dofmap.node_eqn, dofmap.pos_eqn], axis=1)
    ext_eqn = all_eqn.copy()
    hang_by_field = {}
    for node, fi, _, _ in todo:
        hang_by_field.setdefault(fi, set()).add(node)
    pr, pc, pv = [], [], []
    nxt = n
    for node, fi, masters, weights in todo:
        assert all_eqn[node, fi] < 0, "a hanging value must not have an equation of its own"
        if any(int(m) in hang_by_field[fi] for m in masters):
            raise NotImplementedError("masters that hang themselves (more than one refinement level across an edge)")
        meq = all_eqn[np.asarray(masters), fi]
        live = meq >= 0
        if not live.any():
            continue
        ext_eqn[node, fi] = nxt
        pr += [nxt] * int(live.sum())
        pc += [int(g) for g in meq[live]]
        pv += [float(w) for w in np.asarray(weights)[live]]
        nxt += 1
    node_eqn = np.ascontiguousarray(ext_eqn[:, :nval])
    pos_ext = None if pos_eqn is None else np.ascontiguousarray(ext_eqn[:, nval:])
    return ExtendedNumbering(DofMap(node_eqn, pos_ext, nxt), n, np.array(pr, dtype=np.int64), np.array(pc, dtype=np.int64), np.array(pv))


def constraint_lists(indptr: np.ndarray, indices: np.ndarray, ext: ExtendedNumbering):
    """For the CSR pattern of the extended system: the reduction lists of P^T J P.
    Returns dict with
      target_pos [nt], src_start [nt+1], src_pos [ns], src_w [ns]   matrix: J[target] += sum_k w_k J[src_k]   (sources in ascending position)
      res_row [nr], res_start [nr+1], res_src [..], res_w [..]        residual: R[row] += sum_k w_k R[src_k]
      clear_pos                                                       all entries of virtual rows and columns
      diag_pos, virt_rows                                             diagonal entries / rows of the virtual equations
    Raises if a target entry is missing from the pattern (it must have been added as an extra pattern entry)."""
    n, n_ext = ext.n_real, ext.n_ext
    nv = n_ext - n
    # what an equation of the extended system stands for: itself (weight 1) or, for a virtual one, its masters
    order = np.argsort(ext.P_rows, kind="stable")
    prow, pcol, pval = ext.P_rows[order] - n, ext.P_cols[order], ext.P_vals[order]
    cnt = np.ones(n_ext, dtype=np.int64)
    cnt[n:] = np.bincount(prow, minlength=nv)
    lst_start = np.concatenate([[0], np.cumsum(cnt)])
    lst_eq = np.concatenate([np.arange(n, dtype=np.int64), pcol])
    lst_w = np.concatenate([np.ones(n), pval])
    rows_of_pos = np.repeat(np.arange(n_ext, dtype=np.int64), np.diff(indptr))
    cols_of_pos = indices.astype(np.int64)
    src = np.nonzero((rows_of_pos >= n) | (cols_of_pos >= n))[0]
    r_eq, c_eq = rows_of_pos[src], cols_of_pos[src]
    r_cnt, c_cnt = cnt[r_eq], cnt[c_eq]
    tot = r_cnt * c_cnt
    S = np.repeat(np.arange(src.size), tot)                       # source index of every (target, weight) pair
    off = np.arange(int(tot.sum())) - np.repeat(np.cumsum(tot) - tot, tot)
    ri, ci = off // c_cnt[S], off % c_cnt[S]
    t_row, t_col = lst_eq[lst_start[r_eq[S]] + ri], lst_eq[lst_start[c_eq[S]] + ci]
    w = lst_w[lst_start[r_eq[S]] + ri] * lst_w[lst_start[c_eq[S]] + ci]
    # CSR position of every target
    t_pos = np.empty(t_row.size, dtype=np.int64)
    for k in range(t_row.size):          # targets are few (entries next to hanging nodes): a plain loop over searchsorted is fine
        a, b = indptr[t_row[k]], indptr[t_row[k] + 1]
        j = a + np.searchsorted(indices[a:b], t_col[k])
        if j >= b or indices[j] != t_col[k]:
            raise RuntimeError("constraint target (%d, %d) is not in the CSR pattern" % (t_row[k], t_col[k]))
        t_pos[k] = j
    o2 = np.lexsort((src[S], t_pos))
    t_sorted, s_sorted, w_sorted = t_pos[o2], src[S][o2], w[o2]
    target_pos, first = np.unique(t_sorted, return_index=True)
    src_start = np.concatenate([first, [t_sorted.size]])
    # residual
    res_row_all, res_src_all, res_w_all = pcol, prow + n, pval
    o3 = np.lexsort((res_src_all, res_row_all))
    rr, rs, rw = res_row_all[o3], res_src_all[o3], res_w_all[o3]
    res_row, rfirst = np.unique(rr, return_index=True)
    res_start = np.concatenate([rfirst, [rr.size]])
    virt_rows = np.arange(n, n_ext, dtype=np.int64)
    diag_pos = np.empty(nv, dtype=np.int64)
    for k, r in enumerate(virt_rows):
        a, b = indptr[r], indptr[r + 1]
        j = a + np.searchsorted(indices[a:b], r)
        assert j < b and indices[j] == r
        diag_pos[k] = j
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return dict(target_pos=i32(target_pos), src_start=i32(src_start), src_pos=i32(s_sorted), src_w=np.ascontiguousarray(w_sorted, dtype=np.float64),
                res_row=i32(res_row), res_start=i32(res_start), res_src=i32(rs), res_w=np.ascontiguousarray(rw, dtype=np.float64),
                clear_pos=i32(src), diag_pos=i32(diag_pos), virt_rows=i32(virt_rows))


def extra_pattern_for_constraints(code: FiniteElementCode, mesh, ext: ExtendedNumbering):
    """(rows, cols) of the entries of P^T S P (S = the element pattern of the extended system) that S does not hold: the master-master
    couplings the hanging contributions are redirected to"""
    from scipy.sparse import csr_matrix
    from .distributed import element_dof_table, structural_pattern
    ed = element_dof_table(code, mesh, ext.dofmap, np.arange(mesh.n_elem))
    ip, ix = structural_pattern(ed, ext.n_ext)
    S = csr_matrix((np.ones(ix.size), ix, ip), shape=(ext.n_ext, ext.n_ext))
    P = ext.prolongation()
    P.data[:] = 1.0
    T = (P.T @ S @ P).tocoo()                      # n_real x n_real pattern of the reduced matrix
    Tfull = csr_matrix((np.ones(T.nnz), (T.row, T.col)), shape=(ext.n_ext, ext.n_ext))
    Tfull.data[:] = 1.0
    S.data[:] = 1.0
    D = (Tfull - Tfull.multiply(S)).tocoo()
    keep = D.data > 0.5
    return D.row[keep].astype(np.int32), D.col[keep].astype(np.int32)


class HangingNodeAssembly:
    """`B200Assembly` of an element class on a mesh with hanging nodes.  Same calls; dof vectors and results are those of the REAL
    equations (`dofmap`), the virtual ones stay inside."""

    def __init__(self, code: FiniteElementCode, mesh, dofmap: DofMap, *, name: str = "elem", device: int = 0, **kw):
        hanging = getattr(mesh, "hanging", None)
        if hanging is None:
            raise ValueError("the mesh has no hanging nodes: use B200Assembly")
        self.code, self.mesh, self.dofmap, self.hanging = code, mesh, dofmap, hanging
        self.ext = extend_numbering(code, dofmap, hanging)
        self._pinned_offset = {}
        self._pinned_pos_offset = {}
        self.P = self.ext.prolongation()
        extra = extra_pattern_for_constraints(code, mesh, self.ext)
        if "patch_hint" not in kw and not hasattr(mesh, "element_patches"):
            kw["patch_hint"] = "spatial"          # refined meshes have no lattice order
        self.asm = B200Assembly(code, mesh, self.ext.dofmap, name=name, device=device, extra_pattern=extra, **kw)
        self.n_dof, self.n_ext = self.ext.n_real, self.ext.n_ext
        self.lists = constraint_lists(self.asm.indptr, self.asm.indices, self.ext)
        # the n_dof x n_dof block of the extended pattern: what fetch() hands out
        from scipy.sparse import csr_matrix
        A = csr_matrix((np.arange(1, self.asm.nnz + 1, dtype=np.float64), self.asm.indices, self.asm.indptr), shape=(self.n_ext, self.n_ext))
        B = A[:self.n_dof, :self.n_dof].tocsr()
        B.sort_indices()
        self.indptr, self.indices = B.indptr.astype(np.int32), B.indices.astype(np.int32)
        self._block_pos = (B.data - 1).astype(np.int64)
        self.nnz = int(self.indices.size)
        if device >= 0:
            L = self.lists
            ip = lambda a: a.ctypes.data_as(c_int_p)
            dp = lambda a: a.ctypes.data_as(c_double_p)
            _check(self.asm.lib.pb2_problem_set_constraints(
                self.asm.prob, ctypes.c_longlong(L["target_pos"].size), ip(L["target_pos"]), ip(L["src_start"]), ip(L["src_pos"]), dp(L["src_w"]),
                ctypes.c_longlong(L["res_row"].size), ip(L["res_row"]), ip(L["res_start"]), ip(L["res_src"]), dp(L["res_w"]),
                ctypes.c_longlong(L["clear_pos"].size), ip(L["clear_pos"]), ctypes.c_longlong(L["virt_rows"].size), ip(L["virt_rows"]), ip(L["diag_pos"])))

    # ---- data: hanging values are the interpolation of their masters
    def _interpolate(self, values: np.ndarray) -> np.ndarray:
        v = np.array(values, dtype=np.float64, copy=True)
        for f in self.code.nodal_fields():
            for n, (m, w) in self.hanging.of_space(f.space).items():
                v[n, f.index] = v[np.asarray(m), f.index] @ np.asarray(w)
        return v

    def set_nodal_values(self, t: int, values: np.ndarray):
        v = self._interpolate(values)
        self.asm.set_nodal_values(t, v)
        # what pinned masters (Dirichlet values) add to the hanging values: needed when a dof vector is scattered later
        off = np.zeros(self.n_ext)
        ne = self.ext.dofmap.node_eqn
        for f in self.code.nodal_fields():
            for n, (m, w) in self.hanging.of_space(f.space).items():
                g = ne[n, f.index]
                if g >= 0:
                    pinned = self.dofmap.node_eqn[np.asarray(m), f.index] < 0
                    off[g] = float(np.dot(np.asarray(w)[pinned], v[np.asarray(m)[pinned], f.index]))
        self._pinned_offset[t] = off

    def set_nodal_positions(self, t: int, pos: np.ndarray):
        """positions of level t; (geometrically) hanging nodes are placed on their masters' interpolation
        (BulkElementBase::interpolate_hang_values, src/elements.cpp:448)"""
        x = np.array(pos, dtype=np.float64, copy=True)
        for n, (m, w) in self.hanging.C2.items():
            x[n] = np.asarray(w) @ x[np.asarray(m)]
        self.asm.set_nodal_positions(t, x)
        if self.dofmap.pos_eqn is not None:
            off = np.zeros(self.n_ext)
            pe = self.ext.dofmap.pos_eqn
            for n, (m, w) in self.hanging.C2.items():
                for d in range(x.shape[1]):
                    g = pe[n, d]
                    if g >= 0:
                        pinned = self.dofmap.pos_eqn[np.asarray(m), d] < 0
                        off[g] = float(np.dot(np.asarray(w)[pinned], x[np.asarray(m)[pinned], d]))
            self._pinned_pos_offset[t] = off

    def set_lagrangian_positions(self, pos: np.ndarray):
        x = np.array(pos, dtype=np.float64, copy=True)
        for n, (m, w) in self.hanging.C2.items():
            x[n] = np.asarray(w) @ x[np.asarray(m)]
        self.asm.set_lagrangian_positions(x)

    def set_dofs(self, dofs: np.ndarray, t: int = 0):
        u = self.P @ np.asarray(dofs, dtype=np.float64)
        if t in self._pinned_offset:
            u = u + self._pinned_offset[t]
        if t in self._pinned_pos_offset:
            u = u + self._pinned_pos_offset[t]
        self.asm.set_dofs(u, t)

    def set_parameters(self, **values: float):
        self.asm.set_parameters(**values)

    def shift_time_values(self):
        self.asm.shift_time_values()

    def set_steady(self):
        self.asm.set_steady()

    def set_unsteady(self, *a, **kw):
        self.asm.set_unsteady(*a, **kw)

    # ---- assembly: the engine applies the reduction behind every R/J/M assembly of a problem with constraints
    def assemble(self, flag: int = 1, residual: str = "", parameter: Optional[str] = None, stream: int = 0):
        self.asm.assemble(flag=flag, residual=residual, parameter=parameter, stream=stream)

    def fetch(self, want_jacobian: bool = True, want_mass: bool = False):
        r, j, m = self.asm.fetch(want_jacobian, want_mass)
        return (r[:self.n_dof].copy(), None if j is None else j[self._block_pos], None if m is None else m[self._block_pos])

    def fetch_extended(self, want_jacobian: bool = True, want_mass: bool = False):
        """the device-resident system over n_ext equations: J (+) I, residual zero in the virtual rows"""
        return self.asm.fetch(want_jacobian, want_mass)

    def launch_count(self) -> int:
        return self.asm.launch_count()

    def close(self):
        self.asm.close()
