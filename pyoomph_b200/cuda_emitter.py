"""CUDA emission backend: element-class code model -> one .cu file of batched sm_100a kernels.

This is the counterpart of ``FiniteElementCode::write_code`` (/root/reference/src/codegen.cpp:4684) and
``write_generic_RJM`` (:3912) for the GPU.  Where the reference emits, per (test field, unknown field), one big
expression evaluated inside ``for l_test { for l_shape {...} }`` with a host callback per Gauss point
(src/codegen.cpp:3207-3238 -> src/elements.cpp:3593), this backend emits per routine ONE kernel with three
phases per batch of elements (DESIGN.md "Kernel"):

  phase 0  gather      element nodes -> shared memory (positions, nodal values, BDF/Newmark time combinations)
  phase 1  points      one thread per (element, Gauss point): geometry (the restatement of
                       fill_shape_info_at_s), field interpolation, and the problem-specific straight-line code
                       for the pointwise coefficients R_s, C_{s,(G,a)} (sympy CSE of ResidualForm)
  phase 2  contract    one thread per (element, test node, field group): register-tiled
                       J[row][col] += T_b[l_test] * C * S_a[l_shape] with reference-element shape tables read
                       from the constant bank, then a coloured, first-touch-aware scatter through the
                       element -> CSR position map.

No tensor cores: per-element work is small irregular fp64 contraction (BASELINE north_star).
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List, Optional, Tuple

import sympy as sp
from sympy.printing.c import C99CodePrinter

from . import expressions as ex
from .codegen import AtomInfo, FiniteElementCode, ResidualForm

K3 = 0.774596669241483
K3T = 0.774596662941483   # sic: oomph-lib integral.cc:87-93 (mistyped literal, kept for parity)
K33 = 0.77459666924148


def gauss_rule(dim: int):
    """Literal oomph tables Gauss<2,3> (integral.cc:84-102) and Gauss<3,3> (:169-207)."""
    if dim == 2:
        kn = [(-K3, -K3), (-K3, 0.0), (-K3, K3T), (0.0, -K3), (0.0, 0.0), (0.0, K3T), (K3T, -K3), (K3T, 0.0), (K3T, K3T)]
        w = [25.0 / 81.0, 40.0 / 81.0, 25.0 / 81.0, 40.0 / 81.0, 64.0 / 81.0, 40.0 / 81.0, 25.0 / 81.0, 40.0 / 81.0, 25.0 / 81.0]
        return kn, w
    k = [-K33, 0.0, K33]
    kn = [(k[i], k[j], k[l]) for i in range(3) for j in range(3) for l in range(3)]
    wa, wb, wc, wd = 0.17146776406035, 0.27434842249657, 0.43895747599451, 0.70233196159122
    w1 = {0: 0, 1: 1, 2: 0}
    w = []
    for i in range(3):
        for j in range(3):
            for l in range(3):
                w.append([wa, wb, wc, wd][w1[i] + w1[j] + w1[l]])
    return kn, w


def tgauss_rule():
    """Literal oomph table TGauss<2,3> (integral.cc:355-369): 7 points, degree 5 (Bathe)."""
    a, b, c, d, e = 0.1012865073235, 0.7974269853531, 0.4701420641051, 0.0597158717898, 0.3333333333333
    kn = [(a, a), (b, a), (a, b), (c, d), (c, c), (d, c), (e, e)]
    w = [0.5 * 0.1259391805448] * 3 + [0.5 * 0.1323941527885] * 3 + [0.5 * 0.225]
    return kn, w


def triangle_shape_tables(order: int, knots):
    """psi[ipt][l], dpsi[ipt][l][b] of TElementShape<2,3> (Telements.h:627-664) / TElementShape<2,2> (:519-545), same operation order"""
    psi, dpsi = [], []
    for (s0, s1) in knots:
        if order == 3:
            s2 = 1.0 - s0 - s1
            ps = [2.0 * s0 * (s0 - 0.5), 2.0 * s1 * (s1 - 0.5), 2.0 * s2 * (s2 - 0.5), 4.0 * s0 * s1, 4.0 * s1 * s2, 4.0 * s2 * s0]
            ds = [(4.0 * s0 - 1.0, 0.0), (0.0, 4.0 * s1 - 1.0), (2.0 * (2.0 * s0 - 1.5 + 2.0 * s1), 2.0 * (2.0 * s0 - 1.5 + 2.0 * s1)),
                  (4.0 * s1, 4.0 * s0), (-4.0 * s1, 4.0 * (1.0 - s0 - 2.0 * s1)), (4.0 * (1.0 - 2.0 * s0 - s1), -4.0 * s0)]
        else:
            ps = [s0, s1, 1.0 - s0 - s1]
            ds = [(1.0, 0.0), (0.0, 1.0), (-1.0, -1.0)]
        psi.append(ps)
        dpsi.append(ds)
    return psi, dpsi


TRIANGLE_NODE_COORDS = [(1.0, 0.0), (0.0, 1.0), (0.0, 0.0), (0.5, 0.5), (0.0, 0.5), (0.5, 0.0)]    # Telements.h:575-621


def tgauss3_rule():
    """Literal oomph table TGauss<3,3> (integral.cc:752-776): 11 points (one weight negative), degree 4 (Keast)."""
    a, b, c, d = 0.785714285714286, 0.071428571428571, 0.399403576166799, 0.100596423833201
    kn = [(0.25, 0.25, 0.25), (a, b, b), (b, b, b), (b, a, b), (b, b, a), (c, c, d), (c, d, c), (d, c, c), (c, d, d), (d, c, d), (d, d, c)]
    w = [-0.01315555555556] + [0.00762222222222] * 4 + [0.02488888888889] * 6
    return kn, w


def tetra_shape_tables(order: int, knots):
    """psi[ipt][l], dpsi[ipt][l][b] of TElementShape<3,3> (Telements.h:2133-2197) / TElementShape<3,2> (:1978-2012), same operation order"""
    psi, dpsi = [], []
    for (s0, s1, s2) in knots:
        s3 = 1.0 - s0 - s1 - s2
        if order == 3:
            ps = [(2.0 * s0 - 1.0) * s0, (2.0 * s1 - 1.0) * s1, (2.0 * s2 - 1.0) * s2, (2.0 * s3 - 1.0) * s3,
                  4.0 * s0 * s1, 4.0 * s0 * s2, 4.0 * s0 * s3, 4.0 * s1 * s2, 4.0 * s2 * s3, 4.0 * s1 * s3]
            q = -4.0 * s3 + 1.0
            ds = [(4.0 * s0 - 1.0, 0.0, 0.0), (0.0, 4.0 * s1 - 1.0, 0.0), (0.0, 0.0, 4.0 * s2 - 1.0), (q, q, q),
                  (4.0 * s1, 4.0 * s0, 0.0), (4.0 * s2, 0.0, 4.0 * s0), (4.0 * (s3 - s0), -4.0 * s0, -4.0 * s0),
                  (0.0, 4.0 * s2, 4.0 * s1), (-4.0 * s2, -4.0 * s2, 4.0 * (s3 - s2)), (-4.0 * s1, 4.0 * (s3 - s1), -4.0 * s1)]
        else:
            ps = [s0, s1, s2, 1.0 - s0 - s1 - s2]
            ds = [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (-1.0, -1.0, -1.0)]
        psi.append(ps)
        dpsi.append(ds)
    return psi, dpsi


TETRA_NODE_COORDS = [(1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.0, 0.0, 0.0), (0.5, 0.5, 0.0), (0.5, 0.0, 0.5), (0.5, 0.0, 0.0),
                     (0.0, 0.5, 0.5), (0.0, 0.0, 0.5), (0.0, 0.5, 0.0)]                       # Telements.h:2051-2127


def gauss_rule_1d():
    """Literal oomph table Gauss<1,3> (integral.cc:50-53; these knots are the correct ones)."""
    return [(-0.774596669241483,), (0.0,), (0.774596669241483,)], [5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0]


def line_shape_tables(order: int, knots):
    """QElement<1,3> / QElement<1,2>: the 1D Lagrange polynomials themselves (shape.h:604-650)"""
    psi, dpsi = [], []
    for (s0,) in knots:
        P, D = _lag(order, s0)
        psi.append(list(P))
        dpsi.append([(d,) for d in D])
    return psi, dpsi


def element_rule(et):
    """(knots, weights) of the element type's default integration scheme"""
    if et.name.startswith("QuadFace"):
        k1, w1 = gauss_rule_1d()          # the face s1 = -1 of the bulk element, Gauss<1,3> along it
        return [(k[0], -1.0) for k in k1], list(w1)
    if et.name.startswith("Tri"):
        return tgauss_rule()
    if et.name.startswith("Tetra"):
        return tgauss3_rule()
    if et.elem_dim == 1:
        return gauss_rule_1d()
    return gauss_rule(et.nodal_dim)


def element_shape_tables(et, order: int, knots):
    if et.name.startswith("Tri"):
        return triangle_shape_tables(order, knots)
    if et.name.startswith("Tetra"):
        return tetra_shape_tables(order, knots)
    if et.elem_dim == 1:
        return line_shape_tables(order, knots)
    return shape_tables(et.nodal_dim, order, knots)


def element_node_coords(et):
    """local coordinates of the element's nodes (local_coordinate_of_node)"""
    if et.name.startswith("Tri"):
        return list(TRIANGLE_NODE_COORDS)
    if et.name.startswith("Tetra"):
        return list(TETRA_NODE_COORDS)
    grid = (-1.0, 0.0, 1.0)
    return [tuple(grid[(l // 3 ** d) % 3] for d in range(et.elem_dim)) for l in range(et.nnode)]


def _lag(order: int, s: float):
    # oomph-lib shape.h:604-650, same operation order
    if order == 3:
        return [0.5 * s * (s - 1.0), 1.0 - s * s, 0.5 * s * (s + 1.0)], [s - 0.5, -2.0 * s, s + 0.5]
    return [0.5 * (1.0 - s), 0.5 * (1.0 + s)], [-0.5, 0.5]


def shape_tables(dim: int, order: int, knots):
    """psi[ipt][l], dpsi[ipt][l][b] in oomph's tensor-product order (Qelements.cc:348-377, :621-660)."""
    psi, dpsi = [], []
    for s in knots:
        P, D = zip(*[_lag(order, s[a]) for a in range(dim)])
        ps, ds = [], []
        if dim == 2:
            for i in range(order):
                for j in range(order):
                    ds.append((P[1][i] * D[0][j], D[1][i] * P[0][j]))
                    ps.append(P[1][i] * P[0][j])
        else:
            for i in range(order):
                for j in range(order):
                    for k in range(order):
                        ds.append((P[2][i] * P[1][j] * D[0][k], P[2][i] * D[1][j] * P[0][k], D[2][i] * P[1][j] * P[0][k]))
                        ps.append(P[2][i] * P[1][j] * P[0][k])
        psi.append(ps)
        dpsi.append(ds)
    return psi, dpsi


class _CudaPrinter(C99CodePrinter):
    def __init__(self, names):
        super().__init__({"precision": 17})
        self.names = names

    def _print_Symbol(self, e):
        return self.names.get(e, e.name)

    def _print_Float(self, e):
        return repr(float(e))

    def _print_Rational(self, e):
        return "(%d.0/%d.0)" % (e.p, e.q)

    def _print_Integer(self, e):
        return "%d.0" % int(e)

    def _print_Pow(self, e):
        if e.exp.is_Integer:
            n = int(e.exp)
            b = self.parenthesize(e.base, sp.printing.precedence.PRECEDENCE["Mul"])
            if 1 < abs(n) <= 4:
                s = "*".join([b] * abs(n))
                return "(%s)" % s if n > 0 else "(1.0/(%s))" % s
            if n == -1:
                return "(1.0/%s)" % b
            return "pow(%s,%d.0)" % (self._print(e.base), n)
        return super()._print_Pow(e)


@dataclasses.dataclass
class RowGroup:
    space: str
    fields: List[str]
    rows_per_thread: int
    threads_per_elem: int
    thread_off: int = 0      # first thread of the group in the block
    nthreads: int = 0        # threads reserved (warp padded)


@dataclasses.dataclass
class RoutinePlan:
    key: str                 # C identifier suffix
    form: ResidualForm
    res_index: int
    param_index: int


class CudaEmitter:
    def __init__(self, code: FiniteElementCode, name: str = "elem", *, elems_per_block: Optional[int] = None,
                 acc_max: int = 96, ipt_unroll: Optional[int] = None, table_source: Optional[str] = None):
        self.code = code
        self.name = name
        self.et = code.etype
        self.dim = code.nodal_dim
        self.edim = self.et.elem_dim          # local coordinates of the element: < dim for interface elements (lines in 2D)
        self.NN = self.et.nnode
        self.NN1 = self.et.nnode_C1
        self.NIPT = self.et.n_int_pt
        self.layout = code.dof_layout()                      # [(field, space-local node)]
        self.ndof = len(self.layout)
        self.nval = code.n_nodal_values
        import os
        if self.dim == 3:
            acc_max = min(acc_max, 60)        # 3 rows x 27 columns per thread spills at the 168-register cap of the pipelined kernel
        self.acc_max = int(os.environ.get("PB2_ACC_MAX", acc_max))
        if elems_per_block is None and os.environ.get("PB2_EPB"):
            elems_per_block = int(os.environ["PB2_EPB"])
        self.min_blocks = int(os.environ.get("PB2_MINBLOCKS", "2" if self.dim == 2 else "1"))
        import os as _os0
        self.table_source = table_source or _os0.environ.get("PB2_TABLES") or ("const" if self.dim == 2 else "smem")
        import os as _os
        if ipt_unroll is None and _os.environ.get("PB2_IPT_UNROLL"):
            ipt_unroll = int(_os.environ["PB2_IPT_UNROLL"])
        self.ipt_unroll = ipt_unroll if ipt_unroll is not None else 1
        self.routines: List[RoutinePlan] = []
        for i, rn in enumerate(code.residual_names()):
            self.routines.append(RoutinePlan("r%d" % i, code.derive(rn), i, -1))
            for k, p in enumerate(code.global_params):
                self.routines.append(RoutinePlan("r%d_dp%d" % (i, k), code.derive(rn, p), i, k))
        self.hessian = os.environ.get("PB2_NO_HESSIAN", "0") != "1" and (not code.coordinates_as_dofs or code.etype.elem_dim == code.nodal_dim)
        self.hroutines: List[RoutinePlan] = []
        if self.hessian:
            for i, rn in enumerate(code.residual_names()):
                self.hroutines.append(RoutinePlan("h%d" % i, code.hessian_form(rn), i, -1))
                # transposed contraction (flags 4 / 5 of HessianVectorProduct, src/jitbridge.h:637-691): plugin query kind 3
                self.hroutines.append(RoutinePlan("ht%d" % i, code.hessian_form(rn, transposed=True), i, -1))
        self.T_val = code.history_levels()
        self.T_pos = self.T_val if code.coordinates_as_dofs else 1
        self._plan_groups()
        self.EPB_max = elems_per_block or self._default_epb()
        self.smem_budget = int(os.environ.get("PB2_SMEM_BUDGET", str(100 * 1024 if self.dim == 2 else 110 * 1024)))
        self.EPB = self.EPB_max
        self._layout_threads()
        self._kernel_cfg: Dict[str, Tuple[int, int, int]] = {}
        self.pipeline = os.environ.get("PB2_PIPELINE", "1") != "0"
        self.timing = os.environ.get("PB2_TIMING", "0") == "1"
        self.pipe_smem_budget = int(os.environ.get("PB2_PIPE_SMEM", str(200 * 1024)))
        self.pipe_gather_threads = int(os.environ.get("PB2_PIPE_NG", "64"))
        self.pipe_scatter_threads = int(os.environ.get("PB2_PIPE_NS", "128"))
        # 3D: the shape side of the contraction from the 1D factors of the tensor-product basis (18 uniform constants per Gauss point
        # instead of 108 shared-memory table loads for the 27 columns of a Q27 field)
        self.brick = self.dim == 3 and self.NN == 27            # tensor-product 3D elements: the 1D-factor forms below; tetrahedra use the tables
        self.tensor_columns = self.brick and os.environ.get("PB2_TP3D", "1") != "0"
        # ... and phase 1 (geometry + interpolation of the Q27 / position fields) in ONE node loop with psi_l, dpsi_l built from the same
        # 1D factors: 18 table loads per point instead of one per node, table and quantity
        self.tensor_points = self.brick and os.environ.get("PB2_TP3D_POINTS", "1") != "0"
        # round-2 experiment (off: compiled and algebra-checked on the CPU only, not yet run on a GPU): sum factorisation of the column
        # side over the Gauss points, DESIGN.md section 9 item 4
        self.sum_factorise = self.tensor_columns and os.environ.get("PB2_SUMFAC", "1") != "0"
        # 2D analogue: the Q9 column side of a row contracted direction by direction (135 instead of 243 DFMA per Q9 block and row);
        # Q4 column blocks keep the table form
        self.sum_factorise_2d = self.dim == 2 and os.environ.get("PB2_SUMFAC2D", "0") == "1"

    # ------------------------------------------------------------------ planning
    def _col_index(self, field: str, lnode: int) -> int:
        return self.layout.index((field, lnode))

    def _nnode_space(self, space: str) -> int:
        return self.NN if space in ("C2", "Pos") else self.NN1

    def _plan_groups(self):
        code = self.code
        by_space: Dict[str, List[str]] = {}
        for f in code.unknown_field_names():
            by_space.setdefault(code.fields[f].space, []).append(f)
        groups: List[RowGroup] = []
        for space, fl in by_space.items():
            nn = self._nnode_space(space)
            # rows per thread: whole fields per group; RB>1 only for 3D single-field groups
            per_field = self.ndof
            maxf = max(1, self.acc_max // per_field)
            for i in range(0, len(fl), maxf):
                chunk = fl[i:i + maxf]
                rb = 1
                if self.dim == 3:
                    for cand in (3,):
                        if nn % cand == 0 and cand * len(chunk) * per_field <= self.acc_max:
                            rb = cand
                groups.append(RowGroup(space, chunk, rb, nn // rb))
        self.groups = groups

    def _default_epb(self) -> int:
        tpe = sum(g.threads_per_elem for g in self.groups)
        target_threads = 182 if self.dim == 2 else 192
        return max(4, target_threads // max(1, tpe))

    def _layout_threads(self):
        off = 0
        for g in self.groups:
            g.thread_off = off
            g.nthreads = ((self.EPB * g.threads_per_elem + 31) // 32) * 32
            off += g.nthreads
        self.NT = max(off, 64)

    # ------------------------------------------------------------------ shared-memory plan of one routine
    def _plan_smem(self, form: ResidualForm, what: int):
        """Returns dict with element-local offsets (in doubles)."""
        code, dim, NN = self.code, self.dim, self.NN
        need_lagr = form.uses_dX or any(a.deriv.startswith("dX") for a in form.atoms) or \
            any(k[2].startswith("dX") for k in list(form.J) + list(form.M)) or any(s.deriv.startswith("dX") for s in form.slots)
        # nodal source arrays: (field, kind) kind = ("cur", past) | ("dt", order, scheme)
        sources: List[Tuple[str, tuple]] = []
        for a in form.atoms:
            if a.field.startswith("lagrangian_"):
                continue
            kind = ("dt", a.dt_order, a.scheme) if a.dt_order else ("cur", a.past)
            if a.field.startswith("coordinate_") and kind == ("cur", 0):
                continue  # current positions are always staged
            if (a.field, kind) not in sources:
                sources.append((a.field, kind))
        off = 0
        plan = {"need_lagr": need_lagr, "sources": sources}
        plan["xpos"] = off
        off += NN * dim
        if need_lagr:
            plan["xlag"] = off
            off += NN * dim
        plan["src_off"] = {}
        for (f, kind) in sources:
            plan["src_off"][(f, kind)] = off
            off += self._nnode_space(code.fields[f].space)
        plan["EL0"] = off
        # per-point block
        poff = 0
        plan["gg"] = poff
        poff += self.edim * dim
        if need_lagr:
            plan["ggL"] = poff
            poff += self.edim * dim
        ncoef = len(form.slots)
        jkeys = sorted(form.J.keys()) if what >= 1 else []
        mkeys = sorted(form.M.keys()) if what >= 2 else []
        plan["R_off"] = poff
        poff += ncoef
        # Jacobian / mass coefficients: symbolically identical expressions share one slot of the point block (NS: 17 -> 9), every
        # slot costs two shared-memory wavefronts per warp and Gauss point in the contraction (profiles/r01_notes.md)
        dedup = os.environ.get("PB2_DEDUP", "1") != "0"
        seen: Dict[object, int] = {}

        def slot_of(expr):
            nonlocal poff
            key = sp.srepr(expr) if dedup else object()
            if key not in seen:
                seen[key] = poff
                poff += 1
            return seen[key]
        plan["J_off"] = {k: slot_of(form.J[k]) for k in jkeys}
        plan["M_off"] = {k: slot_of(form.M[k]) for k in mkeys}
        plan["PB"] = poff
        els = plan["EL0"] + self.NIPT * poff
        nstage = self.ndof * self.ndof + self.ndof if what >= 1 else self.ndof
        if what >= 2:
            # the mass pass needs the point data again after the Jacobian pass was staged: no aliasing
            plan["SJ_off"] = els
            plan["stage_alias"] = False
            els += nstage
        else:
            plan["SJ_off"] = 0
            plan["stage_alias"] = True
            els = max(els, nstage)
        if els % 2 == 0:
            els += 1          # odd stride: elements of one warp fall into different banks
        plan["ELS"] = els
        return plan

    # ------------------------------------------------------------------ code pieces
    def _w_name(self, scheme: str, order: int) -> str:
        return "a.ti.w_%s_%s" % ("dt" if order == 1 else "d2t", scheme)

    def _emit_tables(self, o: List[str]):
        kn, w = element_rule(self.et)
        psi2, dpsi2 = element_shape_tables(self.et, 3, kn)
        psi1, dpsi1 = element_shape_tables(self.et, 2, kn)

        def arr(vals):
            return ", ".join(repr(float(v)) for v in vals)
        t1 = []
        if self.brick:
            # 1D factors per Gauss point: [ipt][dir][L_0..L_2, L'_0..L'_2] at the knot of that direction (psi_c = L_i(s0) L_j(s1) L_k(s2),
            # c = i + 3j + 9k, Qelements.cc:621-660)
            for sk in kn:
                for d in range(3):
                    P, D = _lag(3, sk[d])
                    t1 += list(P) + list(D)
        t1_smem = t1
        if self.dim == 2 and self.sum_factorise_2d:
            # the same for quads: [ipt][dir][L_0..L_2, L'_0..L'_2]; psi_c = L_a(s0) L_b(s1), c = a + 3b (Qelements.cc:348-377); constant bank only
            t2 = []
            for sk in kn:
                for d in range(2):
                    P, D = _lag(3, sk[d])
                    t2 += list(P) + list(D)
            o.append("__constant__ double c_t1d[%d] = {%s};" % (len(t2), ", ".join(repr(float(v)) for v in t2)))
        o.append("// reference-element tables at the oomph Gauss points (integral.cc literals, shape.h polynomials)")
        o.append("__constant__ double c_w[%d] = {%s};" % (self.NIPT, arr(w)))
        o.append("__constant__ double c_psi2[%d] = {%s};" % (self.NIPT * self.NN, arr(v for p in psi2 for v in p)))
        o.append("__constant__ double c_dpsi2[%d] = {%s};" % (self.NIPT * self.NN * self.edim, arr(v for p in dpsi2 for l in p for v in l)))
        o.append("__constant__ double c_psi1[%d] = {%s};" % (self.NIPT * self.NN1, arr(v for p in psi1 for v in p)))
        o.append("__constant__ double c_dpsi1[%d] = {%s};" % (self.NIPT * self.NN1 * self.edim, arr(v for p in dpsi1 for l in p for v in l)))
        o.append("__device__ const double g_tables[%d] = {%s};" % (self._tables_smem_size(), arr(
            [v for p in psi2 for v in p] + [v for p in dpsi2 for l in p for v in l] + [v for p in psi1 for v in p] + [v for p in dpsi1 for l in p for v in l] + t1_smem)))
        if self.brick:
            o.append("__constant__ double c_t1d[%d] = {%s};" % (len(t1), arr(t1)))
        if self.code.point_expression_names():
            # the same tables at the element's NODES (local coordinates -1, 0, 1 per direction, oomph node order): point expressions are
            # evaluated at the integration points or at the nodes (eval_local_expression_at_node, src/elements.cpp:4659)
            nk = element_node_coords(self.et)
            npsi2, ndpsi2 = element_shape_tables(self.et, 3, nk)
            npsi1, ndpsi1 = element_shape_tables(self.et, 2, nk)
            nt1 = []
            if self.brick:
                for sk in nk:
                    for d in range(3):
                        P, D = _lag(3, sk[d])
                        nt1 += list(P) + list(D)
            o.append("__device__ const double g_tables_nodes[%d] = {%s};" % (self._tables_smem_size(self.NN), arr(
                [v for p in npsi2 for v in p] + [v for p in ndpsi2 for l in p for v in l] + [v for p in npsi1 for v in p] + [v for p in ndpsi1 for l in p for v in l] + nt1)))
        o.append("__constant__ int c_c1node[%d] = {%s};" % (self.NN1, ", ".join(str(n) for n in self.et.c1_nodes)))
        # row dof index of (field, space-local node)
        for f in self.code.unknown_field_names():
            nn = self._nnode_space(self.code.fields[f].space)
            o.append("__constant__ int c_row_%s[%d] = {%s};" % (f, nn, ", ".join(str(self._col_index(f, l)) for l in range(nn))))
        o.append("")

    def _tables_smem_size(self, npt: Optional[int] = None) -> int:
        """doubles of shape tables staged in shared memory (needed by phase 1 always, phase 2 in smem mode) for npt points
        (default: the integration points)"""
        npt = self.NIPT if npt is None else npt
        return npt * (self.NN * (1 + self.edim) + self.NN1 * (1 + self.edim)) + (npt * 18 if self.brick else 0)

    def _emit_kernel(self, o: List[str], rp: RoutinePlan, what: int):
        code, dim, NN, NN1, NIPT = self.code, self.dim, self.NN, self.NN1, self.NIPT
        form = rp.form
        plan = self._plan_smem(form, what)
        ELS, EL0, PB = plan["ELS"], plan["EL0"], plan["PB"]
        kname = "pb2_%s_%s_f%d" % (self.name, rp.key, what)
        tab_n = self._tables_smem_size()
        ND, ND2 = self.ndof, self.ndof * self.ndof
        # elements per block of THIS kernel: as many as the thread budget allows, capped by the shared-memory budget
        per_el = ELS * 8 + 2 * ND * 4 + (2 * ND2 if what >= 1 else 0)
        self.EPB = max(2, min(self.EPB_max, (self.smem_budget - tab_n * 8) // per_el))
        self._layout_threads()
        smem_doubles = tab_n + self.EPB * ELS
        # staged scatter maps: rowstart[EPB][ND] + res[EPB][ND] ints, then the (8|16 bit) offset bytes
        map_ints = 2 * self.EPB * ND
        map_bytes = 2 * self.EPB * ND2 if what >= 1 else 0
        self._kernel_smem[kname] = smem_doubles * 8 + map_ints * 4 + ((map_bytes + 15) // 16) * 16
        self._kernel_cfg[kname] = (self.EPB, self.NT, self._kernel_smem[kname])
        w = o.append
        T2 = "s_psi2" if self.table_source == "smem" else "c_psi2"
        w("// %s  what=%d : %d slots, %d Jacobian coefficients, %d mass coefficients, element smem %d doubles" % (
            form.name or "<default residual>", what, len(form.slots), len(form.J) if what else 0, len(form.M) if what > 1 else 0, ELS))
        w("extern \"C\" __global__ void __launch_bounds__(%d, %d) %s(const pb2_kernel_args a)" % (self.NT, self.min_blocks, kname))
        w("{")
        w("  extern __shared__ double smem[];")
        w("  double* const s_psi2 = smem;")
        w("  double* const s_dpsi2 = s_psi2 + %d;" % (NIPT * NN))
        w("  double* const s_psi1 = s_dpsi2 + %d;" % (NIPT * NN * self.edim))
        w("  double* const s_dpsi1 = s_psi1 + %d;" % (NIPT * NN1))
        w("  const double* const s_t1d = s_dpsi1 + %d; (void)s_t1d;" % (NIPT * NN1 * self.edim))
        w("  double* const s_el = smem + %d;" % tab_n)
        w("  int* const s_rowstart = (int*)(smem + %d);" % (tab_n + self.EPB * ELS))
        w("  int* const s_resmap = s_rowstart + %d;" % (self.EPB * self.ndof))
        w("  unsigned char* const s_map = (unsigned char*)(s_resmap + %d);" % (self.EPB * self.ndof))
        w("  const int tid = threadIdx.x;")
        w("  for (int i = tid; i < %d; i += %d) smem[i] = g_tables[i];   // psi2 | dpsi2 | psi1 | dpsi1, once per persistent block" % (tab_n, self.NT))
        w("  const int nbatch = (a.n_elem + %d - 1) / %d;" % (self.EPB, self.EPB))
        w("  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x)")
        w("  {")
        w("    const int e0 = batch * %d;" % self.EPB)
        w("    const int nel = min(%d, a.n_elem - e0);" % self.EPB)
        w("    __syncthreads();")
        # ---------------- phase 0
        self._emit_gather_sync(o, plan, ELS)
        w("    // scatter maps of the batch: issued together with the gather, consumed in phase 3 (no dependent global load there)")
        w("    {")
        w("      const long long eg0 = (long long)(a.elem_begin + e0);")
        w("      for (int i = tid; i < nel * %d; i += %d) { s_rowstart[i] = __ldg(a.elem_rowstart + eg0 * %d + i); s_resmap[i] = __ldg(a.elem_res + eg0 * %d + i); }" % (self.ndof, self.NT, self.ndof, self.ndof))
        if what >= 1:
            nd2 = self.ndof * self.ndof
            w("      const int mbytes = nel * %d * (a.map_bits >> 3);" % nd2)
            w("      const unsigned char* __restrict__ gmap = (const unsigned char*)a.elem_off + eg0 * %d * (a.map_bits >> 3);" % nd2)
            if nd2 % 4 == 0:
                w("      for (int i = tid; i < (mbytes >> 2); i += %d) ((unsigned*)s_map)[i] = __ldg((const unsigned*)gmap + i);" % self.NT)
            else:
                w("      for (int i = tid; i < mbytes; i += %d) s_map[i] = __ldg(gmap + i);" % self.NT)
        w("    }")
        w("    __syncthreads();")
        # ---------------- phase 1
        w("    // ---- phase 1: one thread per (element, Gauss point): geometry + interpolation + pointwise coefficients")
        w("    for (int i = tid; i < nel * %d; i += %d)" % (NIPT, self.NT))
        w("    {")
        w("      const int el = i / %d, ipt = i - el * %d;" % (NIPT, NIPT))
        w("      double* E = s_el + el * %d;" % ELS)
        w("      double* P = E + %d + ipt * %d;" % (EL0, PB))
        self._emit_phase1_body(o, rp, plan, what)
        w("    }")
        w("    __syncthreads();")
        # ---------------- phase 2 + 3, once per output matrix ("J": residual + Jacobian, "M": mass matrix)
        passes = [("J", form.J, plan["J_off"], "a.jac_vals", True)] if what >= 1 else [("R", {}, {}, None, True)]
        if what >= 2:
            passes.append(("M", form.M, plan["M_off"], "a.mass_vals", False))
        nacc = max(self._group_nacc(form, g, coef) for g in self.groups for (_, coef, _, _, _) in passes)
        w("    double acc[%d];" % max(1, nacc))
        for pi_, (pname, coef, coff, target, with_res) in enumerate(passes):
            w("    // ---- phase 2 (%s): register-tiled contraction over (l_test, l_shape)" % pname)
            for g in self.groups:
                self._emit_group_compute(o, rp, plan, g, pname, coef, coff, with_res)
            if plan["stage_alias"]:
                w("    __syncthreads();   // point data is dead from here on: the staging area aliases it")
            w("    // ---- phase 3a (%s): element matrices -> shared memory, dense [row][col]" % pname)
            for g in self.groups:
                self._emit_group_stage(o, rp, plan, g, pname, coef, with_res, target is not None)
            w("    __syncthreads();")
            w("    // ---- phase 3b (%s): cooperative coloured scatter, consecutive threads = consecutive (row,col) entries" % pname)
            self._emit_scatter(o, plan, target, with_res)
            if pi_ + 1 < len(passes):
                w("    __syncthreads();")
        w("  }")
        w("}")
        w("")
        return kname

    def _emit_gather_sync(self, o: List[str], plan, ELS: int):
        """phase 0 of the phase-synchronous kernels: element nodes -> s_el (expects e0, nel, tid, s_el)"""
        code, dim, NN, NN1 = self.code, self.dim, self.NN, self.NN1
        w = o.append
        w("    // ---- phase 0: gather element data (fill_element_info's pointer tables become one staged copy)")
        w("    for (int i = tid; i < nel * %d; i += %d)" % (NN, self.NT))
        w("    {")
        w("      const int el = i / %d, l = i - el * %d;" % (NN, NN))
        w("      const long long node = a.elem_nodes[(long long)(a.elem_begin + e0 + el) * %d + l];" % NN)
        w("      double* E = s_el + el * %d;" % ELS)
        for d in range(dim):
            w("      E[%d + l * %d + %d] = a.node_pos[node * %d + %d];" % (plan["xpos"], dim, d, dim, d))
        if plan["need_lagr"]:
            for d in range(dim):
                w("      E[%d + l * %d + %d] = a.node_lagr[node * %d + %d];" % (plan["xlag"], dim, d, dim, d))
        c1_sources = []
        for (f, kind) in plan["sources"]:
            fld = code.fields[f]
            soff = plan["src_off"][(f, kind)]
            if fld.space == "Pos":
                dd = ex.DIRS.index(f[-1])
                base = "a.node_pos"
                stride, comp, nh = dim, dd, "a.n_hist_pos"
            else:
                base = "a.node_val"
                stride, comp, nh = self.nval, fld.index, "a.n_hist_val"
            if fld.space == "C1":
                c1_sources.append((f, kind, soff, base, stride, comp))
                continue
            if kind[0] == "cur":
                w("      E[%d + l] = %s[((long long)%d * a.n_node + node) * %d + %d];" % (soff, base, kind[1], stride, comp))
            else:
                wn = self._w_name(kind[2], kind[1])
                w("      { double s = 0.0; for (int t = 0; t < a.ti.ntstorage; ++t) s += %s[t] * %s[((long long)t * a.n_node + node) * %d + %d]; E[%d + l] = s; }" % (
                    wn, base, stride, comp, soff))
        w("    }")
        if c1_sources:
            w("    for (int i = tid; i < nel * %d; i += %d)" % (NN1, self.NT))
            w("    {")
            w("      const int el = i / %d, l = i - el * %d;" % (NN1, NN1))
            w("      const long long node = a.elem_nodes[(long long)(a.elem_begin + e0 + el) * %d + c_c1node[l]];" % NN)
            w("      double* E = s_el + el * %d;" % ELS)
            for (f, kind, soff, base, stride, comp) in c1_sources:
                if kind[0] == "cur":
                    w("      E[%d + l] = %s[((long long)%d * a.n_node + node) * %d + %d];" % (soff, base, kind[1], stride, comp))
                else:
                    wn = self._w_name(kind[2], kind[1])
                    w("      { double s = 0.0; for (int t = 0; t < a.ti.ntstorage; ++t) s += %s[t] * %s[((long long)t * a.n_node + node) * %d + %d]; E[%d + l] = s; }" % (
                        wn, base, stride, comp, soff))
            w("    }")

    def _emit_integral_kernel(self, o: List[str]) -> str:
        """EvalIntegralExpression for all integral expressions at once (src/codegen.cpp:4125-4364): gather, one thread per
        (element, Gauss point) for geometry + interpolation + integrands (they carry their measure), then one thread per
        (element, expression) sums the Gauss points in order and writes the per-element value."""
        form = self.code.integral_form()
        rp = RoutinePlan("integrals", form, 0, -1)
        plan = self._plan_smem(form, 0)
        ELS, EL0, PB = plan["ELS"], plan["EL0"], plan["PB"]
        NI, NIPT, NN, NN1, dim = len(form.slots), self.NIPT, self.NN, self.NN1, self.dim
        tab_n = self._tables_smem_size()
        NT = 256
        EPB = max(2, min(NT // NIPT if NIPT <= NT else 2, (self.smem_budget - tab_n * 8) // (ELS * 8)))
        kname = "pb2_%s_integrals" % self.name
        self._kernel_smem[kname] = (tab_n + EPB * ELS) * 8
        self._kernel_cfg[kname] = (EPB, NT, self._kernel_smem[kname])
        w = o.append
        w("// integral expressions: %s" % ", ".join(self.code.integral_expression_names()))
        w("extern \"C\" __global__ void __launch_bounds__(%d) %s(const pb2_kernel_args a)" % (NT, kname))
        w("{")
        w("  extern __shared__ double smem[];")
        w("  double* const s_psi2 = smem;")
        w("  double* const s_dpsi2 = s_psi2 + %d;" % (NIPT * NN))
        w("  double* const s_psi1 = s_dpsi2 + %d;" % (NIPT * NN * self.edim))
        w("  double* const s_dpsi1 = s_psi1 + %d;" % (NIPT * NN1))
        w("  const double* const s_t1d = s_dpsi1 + %d; (void)s_t1d;" % (NIPT * NN1 * self.edim))
        w("  double* const s_el = smem + %d;" % tab_n)
        w("  const int tid = threadIdx.x;")
        w("  for (int i = tid; i < %d; i += %d) smem[i] = g_tables[i];" % (tab_n, NT))
        w("  const int nbatch = (a.n_elem + %d - 1) / %d;" % (EPB, EPB))
        w("  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x)")
        w("  {")
        w("    const int e0 = batch * %d;" % EPB)
        w("    const int nel = min(%d, a.n_elem - e0);" % EPB)
        w("    __syncthreads();")
        saved = self.NT
        self.NT = NT
        try:
            self._emit_gather_sync(o, plan, ELS)
        finally:
            self.NT = saved
        w("    __syncthreads();")
        w("    for (int i = tid; i < nel * %d; i += %d)" % (NIPT, NT))
        w("    {")
        w("      const int el = i / %d, ipt = i - el * %d;" % (NIPT, NIPT))
        w("      double* E = s_el + el * %d;" % ELS)
        w("      double* P = E + %d + ipt * %d;" % (EL0, PB))
        self._emit_phase1_body(o, rp, plan, 0)
        w("    }")
        w("    __syncthreads();")
        w("    for (int i = tid; i < nel * %d; i += %d)" % (NI, NT))
        w("    {")
        w("      const int el = i / %d, k = i - el * %d;" % (NI, NI))
        w("      const double* P = s_el + el * %d + %d + k;" % (ELS, EL0 + plan["R_off"]))
        w("      double s = 0.0;")
        w("      for (int ipt = 0; ipt < %d; ++ipt) s += P[ipt * %d];" % (NIPT, PB))
        w("      a.integrals[(long long)(a.elem_begin + e0 + el) * %d + k] = s;" % NI)
        w("    }")
        w("  }")
        w("}")
        w("")
        return kname

    def _emit_point_kernel(self, o: List[str], at_nodes: bool) -> str:
        """EvalLocalExpression / EvalExtremumExpression / GetZ2Fluxes for all point expressions of the class (src/codegen.cpp:4366-4453)
        at one point set: gather, then one thread per (element, point) for geometry + interpolation + the expressions, written to
        integrals[element][point][expression].  The point set only changes the reference-element tables."""
        form = self.code.point_form()
        rp = RoutinePlan("points", form, 0, -1)
        plan = self._plan_smem(form, 0)
        ELS, EL0, PB = plan["ELS"], plan["EL0"], plan["PB"]
        NE, NN, NN1, dim = len(form.slots), self.NN, self.NN1, self.dim
        NPT = NN if at_nodes else self.NIPT
        tab_n = self._tables_smem_size(NPT)
        NT = 256
        EPB = max(2, min(NT // NPT if NPT <= NT else 2, (self.smem_budget - tab_n * 8) // (ELS * 8)))
        kname = "pb2_%s_points_%s" % (self.name, "n" if at_nodes else "g")
        self._kernel_smem[kname] = (tab_n + EPB * ELS) * 8
        self._kernel_cfg[kname] = (EPB, NT, self._kernel_smem[kname])
        w = o.append
        w("// point expressions at the %s: %s" % ("nodes" if at_nodes else "integration points", ", ".join(n for _, n in self.code.point_expression_names())))
        w("extern \"C\" __global__ void __launch_bounds__(%d) %s(const pb2_kernel_args a)" % (NT, kname))
        w("{")
        w("  extern __shared__ double smem[];")
        w("  double* const s_psi2 = smem;")
        w("  double* const s_dpsi2 = s_psi2 + %d;" % (NPT * NN))
        w("  double* const s_psi1 = s_dpsi2 + %d;" % (NPT * NN * self.edim))
        w("  double* const s_dpsi1 = s_psi1 + %d;" % (NPT * NN1))
        w("  const double* const s_t1d = s_dpsi1 + %d; (void)s_t1d;" % (NPT * NN1 * self.edim))
        w("  double* const s_el = smem + %d;" % tab_n)
        w("  const int tid = threadIdx.x;")
        w("  for (int i = tid; i < %d; i += %d) smem[i] = %s[i];" % (tab_n, NT, "g_tables_nodes" if at_nodes else "g_tables"))
        w("  const int nbatch = (a.n_elem + %d - 1) / %d;" % (EPB, EPB))
        w("  for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x)")
        w("  {")
        w("    const int e0 = batch * %d;" % EPB)
        w("    const int nel = min(%d, a.n_elem - e0);" % EPB)
        w("    __syncthreads();")
        saved = self.NT
        self.NT = NT
        try:
            self._emit_gather_sync(o, plan, ELS)
        finally:
            self.NT = saved
        w("    __syncthreads();")
        w("    for (int i = tid; i < nel * %d; i += %d)" % (NPT, NT))
        w("    {")
        w("      const int el = i / %d, ipt = i - el * %d;" % (NPT, NPT))
        w("      double* E = s_el + el * %d;" % ELS)
        w("      double* P = E + %d + ipt * %d;" % (EL0, PB))
        self._emit_phase1_body(o, rp, plan, 0)
        w("      double* const out = a.integrals + ((long long)(a.elem_begin + e0 + el) * %d + ipt) * %d;" % (NPT, NE))
        w("      for (int k = 0; k < %d; ++k) out[k] = P[%d + k];" % (NE, plan["R_off"]))
        w("    }")
        w("  }")
        w("}")
        w("")
        return kname

    # ------------------------------------------------------------------ pipelined, warp-specialised kernel
    def _emit_kernel_pipe(self, o: List[str], rp: RoutinePlan, what: int):
        """One persistent kernel per assembly.  Warps are specialised:
             G (gather) warps   : element nodes -> IN[slot]   (positions, nodal values, time combinations)
             C (compute) warps  : phase 1 (points) + phase 2 (register-tiled contraction) -> OUT[slot]
             S (scatter) warps  : position maps -> MAPS, OUT[slot] -> CSR values / residual (stores + REDs)
           connected by double-buffered shared-memory slots and named barriers (bar.sync / bar.arrive), so global-memory
           latency of gather and scatter is hidden behind the fp64 work.  Batches are ordered (chunk, colour); the scatter
           of the first batch of a tile waits on a global counter until the previous tile is complete, which keeps the
           deterministic colour order and lets the CSR rows of a chunk be finished while they are resident in L2."""
        code, dim, NN, NN1, NIPT = self.code, self.dim, self.NN, self.NN1, self.NIPT
        form = rp.form
        plan = self._plan_smem(form, what)
        EL0, PB = plan["EL0"], plan["PB"]
        ND, ND2 = self.ndof, self.ndof * self.ndof
        kname = "pb2_%s_%s_f%d" % (self.name, rp.key, what)
        tab_n = self._tables_smem_size()
        IN_S = EL0 | 1
        vecP = os.environ.get("PB2_VEC_P", "1") != "0"
        if vecP:
            PB = (PB + 1) // 2 * 2                       # even point block: coefficients are read and written as double2
            if (PB // 2) % 2 == 0 and os.environ.get("PB2_PB_PAD", "1") != "0":
                PB += 2                                  # odd number of 16-byte units per point: the (element, point) threads of phase 1 write their
                #                                          blocks with STS.128 at a stride of PB doubles -- an even count puts all 8 lanes of a quarter
                #                                          warp on one bank group (measured on Q27 heat: 174 of 201 wavefronts per element were conflicts)
            PT_S = NIPT * PB
            if (PT_S // 2) % 2 == 0:
                PT_S += 2                                # element stride = odd number of 16-byte units: elements of a warp hit different bank groups
        else:
            PT_S = (NIPT * PB) | 1
        RS = (ND | 1) if os.environ.get("PB2_RS_PAD", "0") != "0" else ND   # optional odd row stride of the staged matrix: measured slower (scatter reads), off
        OUT_S = ((ND * RS + ND) | 1) if what >= 1 else ND
        ROFF = ND * RS if what >= 1 else 0             # residual behind the matrix
        npass = 2 if what >= 2 else 1
        # helper mode: the row group with the least contraction work (the C1 rows of a Taylor-Hood class) also computes phase 1 of
        # the NEXT batch, so the warps of the heavy row groups run the fp64-dense contraction back to back (point data double-buffered)
        helper = self.groups[-1] if (len(self.groups) >= 2 and os.environ.get("PB2_HELPER", "0") != "0") else None
        n_pts_slots = 2 if helper is not None else 1
        per_el = 8 * (2 * IN_S + n_pts_slots * PT_S + 2 * OUT_S) + 2 * (2 * ND * 4 + (2 * ND2 if what >= 1 else 0)) + 2 * NN * 4
        budget = (self.pipe_smem_budget if helper is None else max(self.pipe_smem_budget, 226 * 1024)) - tab_n * 8 - 256
        self.EPB = max(2, min(self.EPB_max, 63, budget // per_el))
        self._layout_threads()
        NC = sum(g.nthreads for g in self.groups)
        NH = helper.nthreads if helper is not None else 0
        NG, NS = self.pipe_gather_threads, self.pipe_scatter_threads
        if NC + NG + NS > 384 and "PB2_PIPE_NS" not in os.environ:
            NG, NS = 32, max(64, 384 - NC - 32)   # keep the register cap (65536 / block size) at 170 for the compute warps
        NT = NC + NG + NS
        EPB = self.EPB
        off_in = tab_n
        off_pts = (off_in + 2 * EPB * IN_S + 1) // 2 * 2   # 16-byte aligned
        off_out = off_pts + n_pts_slots * EPB * PT_S
        off_maps = off_out + 2 * EPB * OUT_S          # doubles; ints/bytes follow
        map_slot_bytes = 2 * EPB * ND * 4 + ((EPB * ND2 * 2 if what >= 1 else 0) + 15) // 16 * 16 + 16   # + the leading bytes of an unaligned map slice
        smem_bytes = off_maps * 8 + 2 * map_slot_bytes + 2 * EPB * NN * 4
        self._kernel_smem[kname] = smem_bytes
        self._kernel_cfg[kname] = (EPB, NT, smem_bytes)
        plan = dict(plan)
        plan["RS"] = RS
        plan["PB"] = PB
        plan["vecP"] = vecP
        plan["PT_expr"] = "s_pts + el * %d" % PT_S
        plan["SJ_expr"] = "s_out + el * %d" % OUT_S
        w = o.append
        w("// %s  what=%d : pipelined; %d compute + %d gather + %d scatter threads, %d elements per batch, %d B smem" % (
            form.name or "<default residual>", what, NC, NG, NS, EPB, smem_bytes))
        # blocks per SM: light element classes (one scalar field on Q9: 63 KB of shared memory, 10 accumulators per thread) are bound by the
        # hand-offs between the roles, not by a pipe -- a second resident block fills the gaps if the registers allow it
        minb = int(os.environ.get("PB2_PIPE_MINBLOCKS", "1"))
        if minb > 1 and (smem_bytes + 1024) * minb > 227 * 1024:
            minb = 1
        w("extern \"C\" __global__ void __launch_bounds__(%d, %d) %s(const pb2_kernel_args a)" % (NT, minb, kname))
        w("{")
        w("  extern __shared__ double smem[];")
        w("  double* const s_psi2 = smem;")
        w("  double* const s_dpsi2 = s_psi2 + %d;" % (NIPT * NN))
        w("  double* const s_psi1 = s_dpsi2 + %d;" % (NIPT * NN * self.edim))
        w("  double* const s_dpsi1 = s_psi1 + %d;" % (NIPT * NN1))
        w("  const double* const s_t1d = s_dpsi1 + %d; (void)s_t1d;" % (NIPT * NN1 * self.edim))
        w("  (void)s_psi1; (void)s_dpsi1;")
        w("  const int tid = threadIdx.x;")
        w("  for (int i = tid; i < %d; i += %d) smem[i] = g_tables[i];" % (tab_n, NT))
        w("  __syncthreads();")
        w("  const int ib0 = a.block_begin[blockIdx.x], ib1 = a.block_begin[blockIdx.x + 1];")
        # ---------------------------------------------------------------- gather warps
        w("  if (tid >= %d && tid < %d)" % (NC, NC + NG))
        w("  {")
        w("    const int gt = tid - %d;" % NC)
        w("    int it = 0; long long dbg0 = 0, dbg1 = 0; (void)dbg0; (void)dbg1;")
        w("    int* const s_idx0 = (int*)((unsigned char*)(smem + %d) + %d);" % (off_maps, 2 * map_slot_bytes))
        w("    // element -> node indices travel one batch ahead (cp.async), so the nodal data of a batch needs one memory round trip")
        # Batch table (first element, size | tile, conflict mask) two batches ahead in registers instead of global loads at the top of
        # every batch.  A/B on one box (tools/r02_btab.py): in the SCATTER role it shortens config 2 by 1.9 % (2.703 -> 2.653 ms: the
        # critical role no longer waits for an L2 round trip per batch) but costs the classes with little scatter work per batch
        # 0.3-1.3 % (Poisson, Q27 heat); in the GATHER role it changes nothing.  Default: scatter role for batches of >= 6000 matrix entries.
        bt_pipe_g = os.environ.get("PB2_BT_PIPE_G", "0") != "0"
        bt_pipe_s = os.environ.get("PB2_BT_PIPE_S", "1" if ND2 * EPB >= 6000 else "0") != "0"
        w("    int gt_m0 = 0, gt_m1 = 0, gt_m2 = 0, gt_e0 = 0, gt_e1 = 0, gt_e2 = 0;     // batch table two batches ahead in registers")
        w("    if (ib0 < ib1) { gt_m0 = __ldg(a.batch_meta + ib0); gt_e0 = __ldg(a.batch_elem + ib0); }")
        w("    if (ib0 + 1 < ib1) { gt_m1 = __ldg(a.batch_meta + ib0 + 1); gt_e1 = __ldg(a.batch_elem + ib0 + 1); }")
        w("    if (ib0 < ib1) { const int pe0 = gt_e0, pn = (gt_m0 & 63) * %d; for (int i = gt; i < pn; i += %d) pb2_cp_async4(s_idx0 + i, a.elem_nodes + (long long)pe0 * %d + i); }" % (NN, NG, NN))
        w("    asm volatile(\"cp.async.commit_group;\" ::: \"memory\");")
        w("    for (int batch = ib0; batch < ib1; ++batch, ++it)")
        w("    {")
        w("      const int slot = it & 1;")
        if bt_pipe_g:
            w("      const int nel = gt_m0 & 63, e0 = gt_e0; (void)e0;")
            w("      if (batch + 2 < ib1) { gt_m2 = __ldg(a.batch_meta + batch + 2); gt_e2 = __ldg(a.batch_elem + batch + 2); }")
        else:
            w("      const int nel = a.batch_meta[batch] & 63, e0 = a.batch_elem[batch]; (void)e0;")
            w("      if (batch + 1 < ib1) { gt_m1 = a.batch_meta[batch + 1]; gt_e1 = a.batch_elem[batch + 1]; }")
        w("      const int* const s_idx = s_idx0 + slot * %d;" % (EPB * NN))
        w("      asm volatile(\"cp.async.wait_all;\" ::: \"memory\");")
        w("      pb2_bar_sync(13, %d);                        // indices of this batch visible to all gather warps" % NG)
        if self.timing: w("      long long tg0 = clock64();")
        w("      if (it >= 2) pb2_bar_sync(%d + slot, %d);   // IN[slot] released by the phase-1 threads" % (3, (NH or NC) + NG))
        if self.timing: w("      long long tg1 = clock64(); dbg0 += tg1 - tg0;")
        w("      double* const s_in = smem + %d + slot * %d;" % (off_in, EPB * IN_S))
        w("      for (int i = gt; i < nel * %d; i += %d)" % (NN, NG))
        w("      {")
        w("        const int el = i / %d, l = i - el * %d;" % (NN, NN))
        w("        const long long node = s_idx[i];")
        w("        double* E = s_in + el * %d;" % IN_S)
        self._emit_gather_node(o, plan, c1=False, use_async=True)
        w("      }")
        if any(code.fields[f].space == "C1" for (f, kind) in plan["sources"]):
            w("      for (int i = gt; i < nel * %d; i += %d)" % (NN1, NG))
            w("      {")
            w("        const int el = i / %d, l = i - el * %d;" % (NN1, NN1))
            w("        const long long node = s_idx[el * %d + c_c1node[l]];" % NN)
            w("        double* E = s_in + el * %d;" % IN_S)
            self._emit_gather_node(o, plan, c1=True, use_async=True)
            w("      }")
        w("      asm volatile(\"cp.async.commit_group;\" ::: \"memory\");")
        w("      if (batch + 1 < ib1) { const int pe0 = gt_e1, pn = (gt_m1 & 63) * %d; int* const nx = s_idx0 + (slot ^ 1) * %d; for (int i = gt; i < pn; i += %d) pb2_cp_async4(nx + i, a.elem_nodes + (long long)pe0 * %d + i); }" % (NN, EPB * NN, NG, NN))
        w("      asm volatile(\"cp.async.commit_group;\" ::: \"memory\");")
        w("      asm volatile(\"cp.async.wait_group 1;\" ::: \"memory\");   // nodal data of this batch has landed; next indices may be in flight")
        w("      __threadfence_block();")
        if self.timing: w("      dbg1 += clock64() - tg1;")
        w("      pb2_bar_arrive(%d + slot, %d);            // IN[slot] full" % (1, (NH or NC) + NG))
        w("      gt_m0 = gt_m1; gt_e0 = gt_e1; gt_m1 = gt_m2; gt_e1 = gt_e2;")
        w("    }")
        if self.timing: w("    if (gt == 0 && a.debug) { atomicAdd(a.debug + 0, (unsigned long long)dbg0); atomicAdd(a.debug + 1, (unsigned long long)dbg1); atomicAdd(a.debug + 2, (unsigned long long)it); }")
        w("  }")
        # ---------------------------------------------------------------- scatter warps
        w("  else if (tid >= %d)" % (NC + NG))
        w("  {")
        w("    const int st = tid - %d;" % (NC + NG))
        w("    long long dbs0 = 0, dbs1 = 0, dbs2 = 0, dbs3 = 0; (void)dbs0; (void)dbs1; (void)dbs2; (void)dbs3;")
        w("    int it = 0, item = 0, gated_tile = 0, prev_tile = -1, pending = 0;")
        w("    unsigned char* const maps0 = (unsigned char*)(smem + %d);" % off_maps)
        # prefetch helper (lambda-like macro through a local struct is overkill: emit the loop twice)
        # position maps into shared memory: 8-byte asynchronous copies over the words that cover a batch's slice.  A/B switch
        # PB2_MAP_ASYNC=0: the round-1 form (cp.async of 4 bytes when ndof^2 is a multiple of 4, else byte loads through registers).
        # Classes with large maps (ndof^2 > 1024 bytes per element: the 49-dof moving-mesh class) keep the round-1 form: measured
        # 4 % faster there, while Poisson gains 23 %, the Q27 and Taylor-Hood classes 1-5 % (profiles/r02_notes.md).
        map_async = os.environ.get("PB2_MAP_ASYNC", "1") != "0" and ND2 <= 1024

        def emit_prefetch(indent, pe0_expr, meta_expr, slot_expr):
            w(indent + "{")
            w(indent + "  const int pe0 = %s, pnel = (%s) & 63;" % (pe0_expr, meta_expr))
            w(indent + "  unsigned char* const pm = maps0 + (%s) * %d;" % (slot_expr, map_slot_bytes))
            w(indent + "  int* const prs = (int*)pm; int* const pres = prs + %d;" % (EPB * ND))
            w(indent + "  for (int i = st; i < pnel * %d; i += %d) { pb2_cp_async4(prs + i, a.elem_rowstart + (long long)pe0 * %d + i); pb2_cp_async4(pres + i, a.elem_res + (long long)pe0 * %d + i); }" % (ND, NS, ND, ND))
            if what >= 1:
                w(indent + "  const int mbytes = pnel * %d * (a.map_bits >> 3);" % ND2)
                w(indent + "  unsigned char* const pmap = (unsigned char*)(pres + %d);" % (EPB * ND))
                if not map_async:
                    w(indent + "  const unsigned char* __restrict__ gmap = (const unsigned char*)a.elem_off + (long long)pe0 * %d * (a.map_bits >> 3);" % ND2)
                    if ND2 % 4 == 0:
                        w(indent + "  for (int i = st; i < (mbytes >> 2); i += %d) pb2_cp_async4(pmap + 4 * i, gmap + 4 * i);" % NS)
                    else:
                        w(indent + "  for (int i = st; i < mbytes; i += %d) pmap[i] = __ldg(gmap + i);" % NS)
                else:
                    # The batch's bytes start at any alignment when ndof^2 is odd (9, 27, 31, 49 dofs).  Copy the 8-byte words that cover
                    # them asynchronously (the consumer skips the leading `shift` bytes); byte loads through registers put a global
                    # round trip per batch on the scatter warps' critical path (14 % of all samples of the Q27 kernel, 22 % of the
                    # Poisson kernel's time).
                    w(indent + "  const long long goff = (long long)pe0 * %d * (a.map_bits >> 3);" % ND2)
                    w(indent + "  const int shift = (int)(goff & 7);")
                    w(indent + "  const unsigned char* __restrict__ gal = (const unsigned char*)a.elem_off + (goff - shift);")
                    w(indent + "  for (int i = st; i < ((shift + mbytes + 7) >> 3); i += %d) pb2_cp_async8(pmap + 8 * i, gal + 8 * i);" % NS)
            w(indent + "  asm volatile(\"cp.async.commit_group;\" ::: \"memory\");")
            w(indent + "}")
        w("    // position maps travel one batch ahead of the scatter (cp.async into the other MAPS slot)")
        w("    // the batch table (first element, size | tile, conflict mask) travels TWO batches ahead in registers: the scatter warps are the")
        w("    // critical role and would otherwise wait for an L2 round trip at the top of every batch")
        w("    int bt_m0 = 0, bt_m1 = 0, bt_m2 = 0, bt_e0 = 0, bt_e1 = 0, bt_e2 = 0; unsigned long long bt_b0 = 0, bt_b1 = 0, bt_b2 = 0;")
        w("    if (ib0 < ib1) { bt_m0 = __ldg(a.batch_meta + ib0); bt_e0 = __ldg(a.batch_elem + ib0); bt_b0 = __ldg(a.batch_bar + ib0); }")
        w("    if (ib0 + 1 < ib1) { bt_m1 = __ldg(a.batch_meta + ib0 + 1); bt_e1 = __ldg(a.batch_elem + ib0 + 1); bt_b1 = __ldg(a.batch_bar + ib0 + 1); }")
        w("    if (ib0 < ib1)")
        emit_prefetch("    ", "bt_e0", "bt_m0", "0")
        w("    for (int batch = ib0; batch < ib1; ++batch, ++it)")
        w("    {")
        if bt_pipe_s:
            w("      const int meta = bt_m0, nel = meta & 63, tile = meta >> 7, e0_this = bt_e0; (void)e0_this;")
            w("      const unsigned long long bmask = bt_b0;")
            w("      if (batch + 2 < ib1) { bt_m2 = __ldg(a.batch_meta + batch + 2); bt_e2 = __ldg(a.batch_elem + batch + 2); bt_b2 = __ldg(a.batch_bar + batch + 2); }")
        else:
            w("      const int meta = a.batch_meta[batch], nel = meta & 63, tile = meta >> 7, e0_this = a.batch_elem[batch]; (void)e0_this;")
            w("      const unsigned long long bmask = a.batch_bar[batch];")
            w("      if (batch + 1 < ib1) { bt_m1 = a.batch_meta[batch + 1]; bt_e1 = a.batch_elem[batch + 1]; }")
        w("      unsigned char* const mbase = maps0 + (it & 1) * %d;" % map_slot_bytes)
        w("      int* const s_rowstart = (int*)mbase; int* const s_resmap = s_rowstart + %d;" % (EPB * ND))
        if not map_async:
            w("      unsigned char* const s_map = (unsigned char*)(s_resmap + %d);" % (EPB * ND))
        else:
            w("      unsigned char* const s_map = (unsigned char*)(s_resmap + %d) + (int)(((long long)e0_this * %d * (a.map_bits >> 3)) & 7);" % (EPB * ND, ND2))
        w("      (void)s_map;")
        if self.timing: w("      long long ts0 = clock64();")
        w("      asm volatile(\"cp.async.wait_all;\" ::: \"memory\");")
        w("      const bool publish = prev_tile >= 0 && (tile != prev_tile || (meta & 64));")
        w("      if (publish) __threadfence();               // everything scattered so far becomes globally visible")
        w("      pb2_bar_sync(11, %d);                       // maps of this batch visible to all scatter warps; previous batch fully issued" % NS)
        w("      if (batch + 1 < ib1)")
        emit_prefetch("      ", "bt_e1", "bt_m1", "(it + 1) & 1")
        w("      if (publish && tile != prev_tile) { if (st == 0) atomicAdd(a.tile_done + prev_tile, pending); pending = 0; }")
        w("      prev_tile = tile; ++pending;")
        if self.timing: w("      long long ts1 = clock64(); dbs0 += ts1 - ts0;")
        w("      // stream order of the colours: everything of the previous tile must have been scattered")
        w("      if (tile > gated_tile)")
        w("      {")
        w("        if (st == 0) pb2_gate_wait(a.tile_done + tile - 1, a.tile_nbatch[tile - 1], a.status);")
        w("        gated_tile = tile;")
        w("        pb2_bar_sync(11, %d);                     // gate passed (tile, gated_tile are uniform over the scatter warps)" % NS)
        w("        __threadfence();")
        w("      }")
        is_h = rp.key.startswith("h")      # Hessian-vector routine: matrices only, no residual
        passes = [("J", "a.jac_vals", not is_h)] if what >= 1 else [("R", None, True)]
        if what >= 2:
            passes.append(("M", "a.mass_vals", False))
        w("      #pragma unroll 1")
        w("      for (int pass = 0; pass < %d; ++pass, ++item)" % npass)
        w("      {")
        w("        const int oslot = item & 1;")
        if self.timing: w("        long long ts2 = clock64(); if (pass == 0) dbs1 += ts2 - ts1;")
        w("        pb2_bar_sync(%d + oslot, %d);             // OUT[oslot] full" % (5, NC + NS))
        if self.timing: w("        long long ts3 = clock64(); dbs2 += ts3 - ts2;")
        w("        const double* const s_out = smem + %d + oslot * %d;" % (off_out, EPB * OUT_S))
        for pi_, (pname, target, with_res) in enumerate(passes):
            w("        %sif (pass == %d)" % ("" if pi_ == 0 else "else ", pi_))
            w("        {")
            targs = "%d, %d, %d, %d, %d, %s, %s, %d" % (ND, RS, OUT_S, NS, EPB, "true" if target is not None else "false", "true" if with_res else "false",
                                                         ROFF if target is not None else 0)
            tail = "s_rowstart, s_resmap, s_out, %s, a.residual, nel, st, bmask" % (target if target is not None else "(double*)0")
            if target is not None:
                w("          if (a.map_bits == 8)")
                w("            pb2_scatter_batch<unsigned char, 0x80u, 0xFFu, %s>((const unsigned char*)s_map, %s);" % (targs, tail))
                w("          else")
                w("            pb2_scatter_batch<unsigned short, 0x8000u, 0xFFFFu, %s>((const unsigned short*)s_map, %s);" % (targs, tail))
            else:
                w("          pb2_scatter_batch<unsigned char, 0x80u, 0xFFu, %s>((const unsigned char*)s_map, %s);" % (targs, tail))
            w("        }")
        w("        __threadfence_block();")
        if self.timing: w("        dbs3 += clock64() - ts3;")
        w("        pb2_bar_arrive(%d + oslot, %d);           // OUT[oslot] free again" % (7, NC + NS))
        w("      }")
        w("      bt_m0 = bt_m1; bt_e0 = bt_e1; bt_b0 = bt_b1; bt_m1 = bt_m2; bt_e1 = bt_e2; bt_b1 = bt_b2;")
        w("    }")
        w("    if (prev_tile >= 0) { __threadfence(); pb2_bar_sync(11, %d); if (st == 0) atomicAdd(a.tile_done + prev_tile, pending); }" % NS)
        if self.timing: w("    if (st == 0 && a.debug) { atomicAdd(a.debug + 4, (unsigned long long)dbs0); atomicAdd(a.debug + 5, (unsigned long long)dbs1); atomicAdd(a.debug + 6, (unsigned long long)dbs2); atomicAdd(a.debug + 7, (unsigned long long)dbs3); }")
        w("  }")
        # ---------------------------------------------------------------- compute warps
        w("  else")
        w("  {")
        cpasses = [("J", form.J, plan["J_off"], not is_h, True)] if what >= 1 else [("R", {}, {}, True, False)]
        if what >= 2:
            cpasses.append(("M", form.M, plan["M_off"], False, True))
        nacc = max(self._group_nacc(form, g, coef) for g in self.groups for (_, coef, _, _, _) in cpasses)
        w("    double acc[%d];" % max(1, nacc))
        if helper is not None:
            self._emit_compute_helper_mode(o, rp, plan, what, cpasses, helper, dict(NC=NC, NH=NH, NG=NG, NS=NS, EPB=EPB, IN_S=IN_S, PT_S=PT_S, PB=PB, OUT_S=OUT_S,
                                                                                   off_in=off_in, off_pts=off_pts, off_out=off_out, vecP=vecP))
            w("  }")
            w("}")
            w("")
            return kname
        w("    double* const s_pts = smem + %d;" % off_pts)
        w("    int it = 0, item = 0;")
        w("    long long dbc0 = 0, dbc1 = 0, dbc2 = 0, dbc3 = 0, dbc4 = 0; (void)dbc0; (void)dbc1; (void)dbc2; (void)dbc3; (void)dbc4;")
        w("    for (int batch = ib0; batch < ib1; ++batch, ++it)")
        w("    {")
        w("      const int slot = it & 1;")
        w("      const int nel = a.batch_meta[batch] & 63;")
        if self.timing: w("      long long tc0 = clock64();")
        w("      pb2_bar_sync(%d + slot, %d);                // IN[slot] full" % (1, NC + NG))
        if self.timing: w("      long long tc1 = clock64(); dbc0 += tc1 - tc0;")
        w("      const double* const s_in = smem + %d + slot * %d;" % (off_in, EPB * IN_S))
        w("      // ---- phase 1: one thread per (element, Gauss point)")
        w("      for (int i = tid; i < nel * %d; i += %d)" % (NIPT, NC))
        w("      {")
        w("        const int el = i / %d, ipt = i - el * %d;" % (NIPT, NIPT))
        w("        const double* E = s_in + el * %d;" % IN_S)
        w("        double* P = s_pts + el * %d + ipt * %d;" % (PT_S, PB))
        sub: List[str] = []
        self._emit_phase1_body(sub, rp, plan, what)
        if vecP:
            sub = self._vectorise_point_writes(sub, PB)
        for ln in sub:
            w("  " + ln)
        w("      }")
        w("      __threadfence_block();")
        w("      if (batch + 2 < ib1) pb2_bar_arrive(%d + slot, %d);   // IN[slot] may be refilled" % (3, NC + NG))
        w("      pb2_bar_sync(9, %d);                        // point data complete" % NC)
        if self.timing: w("      long long tc2 = clock64(); dbc1 += tc2 - tc1;")
        for pi_, (pname, coef, coff, with_res, with_matrix) in enumerate(cpasses):
            w("      { // ---- phase 2 (%s): register-tiled contraction, then staging into OUT" % pname)
            sub = []
            for g in self.groups:
                gsub: List[str] = []
                self._emit_group_compute(gsub, rp, plan, g, pname, coef, coff, with_res)
                sub += self._vectorise_point_reads(gsub) if vecP else gsub
            for ln in sub:
                w("  " + ln)
            w("        const int oslot = item & 1;")
            if self.timing: w("        long long tc3 = clock64(); dbc2 += tc3 - tc2;")
            w("        if (item >= 2) pb2_bar_sync(%d + oslot, %d);  // OUT[oslot] drained by the scatter warps" % (7, NC + NS))
            if self.timing: w("        long long tc4 = clock64(); dbc3 += tc4 - tc3;")
            w("        double* const s_out = smem + %d + oslot * %d;" % (off_out, EPB * OUT_S))
            sub = []
            for g in self.groups:
                self._emit_group_stage(sub, rp, plan, g, pname, coef, with_res, with_matrix)
            for ln in sub:
                w("  " + ln)
            w("        __threadfence_block();")
            if self.timing: w("        tc2 = clock64(); dbc4 += tc2 - tc4;")
            w("        pb2_bar_arrive(%d + oslot, %d);           // OUT[oslot] full" % (5, NC + NS))
            w("        ++item;")
            w("      }")
        w("      pb2_bar_sync(10, %d);                       // everybody done with the point data" % NC)
        w("    }")
        if self.timing: w("    if (tid == 0 && a.debug) { atomicAdd(a.debug + 8, (unsigned long long)dbc0); atomicAdd(a.debug + 9, (unsigned long long)dbc1); atomicAdd(a.debug + 10, (unsigned long long)dbc2); atomicAdd(a.debug + 11, (unsigned long long)dbc3); atomicAdd(a.debug + 12, (unsigned long long)dbc4); }")
        w("  }")
        w("}")
        w("")
        return kname

    def _emit_compute_helper_mode(self, o: List[str], rp: RoutinePlan, plan, what: int, cpasses, helper: RowGroup, k):
        """compute warps when one row group (`helper`) doubles as the phase-1 producer of the next batch; barrier ids:
        1,2 IN full / 3,4 IN free (gather <-> helper), 10,15 PTS full / 12,14 PTS free (helper <-> heavy groups), 5,6 OUT full /
        7,8 OUT free (all compute <-> scatter), 9 helper-internal"""
        w = o.append
        NC, NH, NG, NS, EPB = k["NC"], k["NH"], k["NG"], k["NS"], k["EPB"]
        NIPT = self.NIPT
        heavy = [g for g in self.groups if g is not helper]
        PTS_FULL, PTS_FREE = (10, 15), (12, 14)

        def stage_and_publish(groups, pname, coef, with_res, with_matrix, ind):
            w(ind + "const int oslot = item & 1;")
            w(ind + "if (item >= 2) pb2_bar_sync(%d + oslot, %d);  // OUT[oslot] drained by the scatter warps" % (7, NC + NS))
            w(ind + "double* const s_out = smem + %d + oslot * %d;" % (k["off_out"], EPB * k["OUT_S"]))
            sub: List[str] = []
            for g in groups:
                self._emit_group_stage(sub, rp, plan, g, pname, coef, with_res, with_matrix)
            for ln in sub:
                w(ind[:-6] + ln if len(ind) >= 6 else ln)
            w(ind + "__threadfence_block();")
            w(ind + "pb2_bar_arrive(%d + oslot, %d);           // OUT[oslot] full" % (5, NC + NS))
            w(ind + "++item;")

        def compute(groups, pname, coef, coff, with_res, ind):
            sub: List[str] = []
            for g in groups:
                gsub: List[str] = []
                self._emit_group_compute(gsub, rp, plan, g, pname, coef, coff, with_res)
                sub += self._vectorise_point_reads(gsub) if k["vecP"] else gsub
            for ln in sub:
                w(ind[:-6] + ln if len(ind) >= 6 else ln)

        def phase1(batch_expr, slot_expr, ind):
            w(ind + "{ // ---- phase 1 of batch %s: one thread per (element, Gauss point), into PTS[%s]" % (batch_expr, slot_expr))
            w(ind + "  const int nel1 = a.batch_meta[%s] & 63;" % batch_expr)
            w(ind + "  const double* const s_in = smem + %d + (%s) * %d;" % (k["off_in"], slot_expr, EPB * k["IN_S"]))
            w(ind + "  double* const s_ptw = smem + %d + (%s) * %d;" % (k["off_pts"], slot_expr, EPB * k["PT_S"]))
            npar = int(os.environ.get("PB2_P1_ILP", "2"))
            w(ind + "  // %d points per thread in one basic block: their dependent chains (9-term sums, reciprocal, sqrt) interleave" % npar)
            w(ind + "  for (int i0 = tid - %d; i0 < nel1 * %d; i0 += %d)" % (helper.thread_off, NIPT, npar * NH))
            w(ind + "  {")
            for rep in range(npar):
                w(ind + "    {")
                w(ind + "      const int i = min(i0 + %d, nel1 * %d - 1);   // a clamped duplicate rewrites identical values" % (rep * NH, NIPT))
                w(ind + "      const int el = i / %d, ipt = i - el * %d;" % (NIPT, NIPT))
                w(ind + "      const double* E = s_in + el * %d;" % k["IN_S"])
                w(ind + "      double* P = s_ptw + el * %d + ipt * %d;" % (k["PT_S"], k["PB"]))
                sub: List[str] = []
                self._emit_phase1_body(sub, rp, plan, what)
                if k["vecP"]:
                    sub = self._vectorise_point_writes(sub, k["PB"])
                for ln in sub:
                    w(ind + ln)
                w(ind + "    }")
            w(ind + "  }")
            w(ind + "  __threadfence_block();")
            w(ind + "}")
        # ------------------------------------------------ heavy row groups: contraction only
        w("    if (tid < %d)" % helper.thread_off)
        w("    {")
        w("      int it = 0, item = 0;")
        w("      for (int batch = ib0; batch < ib1; ++batch, ++it)")
        w("      {")
        w("        const int slot = it & 1;")
        w("        const int nel = a.batch_meta[batch] & 63;")
        w("        pb2_bar_sync(slot ? %d : %d, %d);          // PTS[slot] full" % (PTS_FULL[1], PTS_FULL[0], NC))
        w("        const double* const s_pts = smem + %d + slot * %d;" % (k["off_pts"], EPB * k["PT_S"]))
        for pi_, (pname, coef, coff, with_res, with_matrix) in enumerate(cpasses):
            w("        { // ---- phase 2 (%s)" % pname)
            compute(heavy, pname, coef, coff, with_res, "          ")
            if pi_ + 1 == len(cpasses):
                w("          __threadfence_block();")
                w("          if (batch + 2 < ib1) pb2_bar_arrive(slot ? %d : %d, %d);   // PTS[slot] may be refilled (results are in registers)" % (PTS_FREE[1], PTS_FREE[0], NC))
            stage_and_publish(heavy, pname, coef, with_res, with_matrix, "          ")
            w("        }")
        w("      }")
        w("    }")
        # ------------------------------------------------ helper row group: its own rows + phase 1 of the next batch
        w("    else")
        w("    {")
        w("      int it = 0, item = 0;")
        w("      if (ib0 < ib1)")
        w("      {")
        w("        pb2_bar_sync(1, %d);                       // IN[0] full" % (NH + NG))
        phase1("ib0", "0", "        ")
        w("        if (ib0 + 2 < ib1) pb2_bar_arrive(3, %d);  // IN[0] may be refilled" % (NH + NG))
        w("      }")
        w("      for (int batch = ib0; batch < ib1; ++batch, ++it)")
        w("      {")
        w("        const int slot = it & 1;")
        w("        const int nel = a.batch_meta[batch] & 63;")
        w("        pb2_bar_sync(9, %d);                        // phase 1 of this batch complete (all helper threads)" % NH)
        w("        pb2_bar_arrive(slot ? %d : %d, %d);        // PTS[slot] full: the heavy row groups may start" % (PTS_FULL[1], PTS_FULL[0], NC))
        w("        const double* const s_pts = smem + %d + slot * %d;" % (k["off_pts"], EPB * k["PT_S"]))
        w("        if (batch + 1 < ib1)   // first the point data the heavy groups wait for next, then this group's own rows")
        w("        {")
        w("          if (it >= 1) pb2_bar_sync(slot ? %d : %d, %d);   // PTS[slot ^ 1] released by the heavy row groups" % (PTS_FREE[0], PTS_FREE[1], NC))
        w("          pb2_bar_sync(slot ? 1 : 2, %d);          // IN[slot ^ 1] full" % (NH + NG))
        phase1("batch + 1", "slot ^ 1", "          ")
        w("          if (batch + 3 < ib1) pb2_bar_arrive(slot ? 3 : 4, %d);   // IN[slot ^ 1] may be refilled" % (NH + NG))
        w("        }")
        for pi_, (pname, coef, coff, with_res, with_matrix) in enumerate(cpasses):
            w("        { // ---- phase 2 (%s), helper rows" % pname)
            compute([helper], pname, coef, coff, with_res, "          ")
            stage_and_publish([helper], pname, coef, with_res, with_matrix, "          ")
            w("        }")
        w("      }")
        w("    }")

    def _emit_gather_node(self, o: List[str], plan, c1: bool, use_async: bool = False):
        code, dim = self.code, self.dim
        w = o.append

        def copy(dst, src):
            if use_async:
                w("        pb2_cp_async8(&%s, &%s);" % (dst, src))
            else:
                w("        %s = %s;" % (dst, src))
        if not c1:
            for d in range(dim):
                copy("E[%d + l * %d + %d]" % (plan["xpos"], dim, d), "a.node_pos[node * %d + %d]" % (dim, d))
            if plan["need_lagr"]:
                for d in range(dim):
                    copy("E[%d + l * %d + %d]" % (plan["xlag"], dim, d), "a.node_lagr[node * %d + %d]" % (dim, d))
        for (f, kind) in plan["sources"]:
            fld = code.fields[f]
            if (fld.space == "C1") != c1:
                continue
            soff = plan["src_off"][(f, kind)]
            if fld.aux_of:
                # Hessian direction vector on the dofs of the underlying field (pinned dofs carry no direction)
                w("        { const int yq = __ldg(a.elem_eqn + (long long)(e0 + el) * %d + c_row_%s[l]); E[%d + l] = yq >= 0 ? __ldg(a.hvec + yq) : 0.0; }" % (
                    self.ndof, fld.aux_of, soff))
                continue
            if fld.space == "Pos":
                base, stride, comp = "a.node_pos", dim, ex.DIRS.index(f[-1])
            else:
                base, stride, comp = "a.node_val", self.nval, fld.index
            if kind[0] == "cur":
                copy("E[%d + l]" % soff, "%s[((long long)%d * a.n_node + node) * %d + %d]" % (base, kind[1], stride, comp))
            else:
                wn = self._w_name(kind[2], kind[1])
                w("        { double s = 0.0; for (int t = 0; t < a.ti.ntstorage; ++t) s += %s[t] * %s[((long long)t * a.n_node + node) * %d + %d]; E[%d + l] = s; }" % (
                    wn, base, stride, comp, soff))

    def _emit_phase1_body(self, o: List[str], rp: RoutinePlan, plan, what: int):
        """geometry + interpolation + CSE'd pointwise coefficients of one (element, Gauss point); expects E, P, ipt"""
        code, dim, NN, NN1, NIPT = self.code, self.dim, self.NN, self.NN1, self.NIPT
        form = rp.form
        w = o.append
        edim = self.edim
        w("      const double* ps2 = s_psi2 + ipt * %d; const double* dp2 = s_dpsi2 + ipt * %d;" % (NN, NN * edim))
        w("      const double* ps1 = s_psi1 + ipt * %d; const double* dp1 = s_dpsi1 + ipt * %d;" % (NN1, NN1 * edim))
        w("      (void)ps1; (void)dp1; (void)ps2;")
        # (field, kind) -> derivatives needed, for all atoms of the form
        needed: Dict[Tuple[str, tuple], set] = {}
        for at in form.atoms:
            kind = ("dt", at.dt_order, at.scheme) if at.dt_order else ("cur", at.past)
            needed.setdefault((at.field, kind), set()).add(at.deriv)

        def src_of(f, kind):
            if f.startswith("lagrangian_"):
                return "E[%d + l * %d + %d]" % (plan["xlag"], dim, ex.DIRS.index(f[-1]))
            if f.startswith("coordinate_") and kind == ("cur", 0):
                return "E[%d + l * %d + %d]" % (plan["xpos"], dim, ex.DIRS.index(f[-1]))
            return "E[%d + l]" % plan["src_off"][(f, kind)]

        def tag_of(f, kind):
            return "%s_%s" % (f, "_".join(str(k) for k in kind))
        fused = set()
        if self.tensor_points:
            # ONE loop over the 27 nodes for the tangents and every Q27 / position quantity; psi_l and its local derivatives are built
            # from the 1D factors of the Gauss point (psi_l = a_i b_j c_k, l = i + 3j + 9k)
            geos = [("xpos", "gg")] + ([("xlag", "ggL")] if plan["need_lagr"] else [])
            for (src, gname) in geos:
                w("      double %s;" % ", ".join("t_%s%d%d = 0.0" % (gname, a, i) for a in range(dim) for i in range(dim)))
            for (f, kind), derivs in needed.items():
                if code.fields[f].space == "C1":
                    continue
                fused.add((f, kind))
                tag = tag_of(f, kind)
                decl = (["v_%s = 0.0" % tag] if "d0" in derivs else []) + (["s%d_%s = 0.0" % (b, tag) for b in range(dim)] if any(d != "d0" for d in derivs) else [])
                w("      double %s;" % ", ".join(decl))
            w("      {")
            w("        const double* t1 = s_t1d + ipt * 18;")
            w("        const double fa0 = t1[0], fa1 = t1[1], fa2 = t1[2], fda0 = t1[3], fda1 = t1[4], fda2 = t1[5];")
            w("        const double fb0 = t1[6], fb1 = t1[7], fb2 = t1[8], fdb0 = t1[9], fdb1 = t1[10], fdb2 = t1[11];")
            w("        #pragma unroll 1")
            w("        for (int kk = 0; kk < 3; ++kk)")
            w("        {")
            w("          const double fc = t1[12 + kk], fdc = t1[15 + kk];")
            for jj in range(3):
                w("          { const double bc = fb%d * fc, dbc = fdb%d * fc, bdc = fb%d * fdc;" % (jj, jj, jj))
                for ii in range(3):
                    w("            { const int l = %d + 9 * kk; const double psi = fa%d * bc, d0 = fda%d * bc, d1 = fa%d * dbc, d2 = fa%d * bdc; (void)psi;" % (ii + 3 * jj, ii, ii, ii, ii))
                    for (src, gname) in geos:
                        for i in range(dim):
                            w("              { const double x = E[%d + l * %d + %d]; %s }" % (plan[src], dim, i, " ".join("t_%s%d%d += x * d%d;" % (gname, a, i, a) for a in range(dim))))
                    for (f, kind), derivs in needed.items():
                        if (f, kind) not in fused:
                            continue
                        tag = tag_of(f, kind)
                        upd = (["v_%s += u * psi;" % tag] if "d0" in derivs else []) + (["s%d_%s += u * d%d;" % (b, tag, b) for b in range(dim)] if any(d != "d0" for d in derivs) else [])
                        w("              { const double u = %s; %s }" % (src_of(f, kind), " ".join(upd)))
                    w("            }")
                w("          }")
            w("        }")
            w("      }")
        self._emit_geometry(o, plan, "xpos", "gg", "detE", sums_done=self.tensor_points)
        if plan["need_lagr"]:
            self._emit_geometry(o, plan, "xlag", "ggL", "detL", sums_done=self.tensor_points)
        w("      const double dx = c_w[ipt] * detE;")
        if plan["need_lagr"]:
            w("      const double dX = c_w[ipt] * detL;")
        # interpolation
        names: Dict[sp.Symbol, str] = {ex.DX_EUL: "dx", ex.DX_LAG: "dX", ex.TIME: "a.ti.t[0]", ex.pi: "3.14159265359",
                                       ex.ELEMSIZE_EUL: "esz_eul", ex.ELEMSIZE_EUL_CART: "esz_cart"}
        for i_, ns_ in enumerate(ex.NORMAL):
            names[ns_] = "nrm%d" % i_          # computed by _emit_geometry on interface elements
        for k, p in enumerate(code.global_params):
            names[code._param_syms[p]] = "a.params[%d]" % k
        all_exprs = list(form.R) + list(form.J.values()) + list(form.M.values())
        names[ex.ELEMSIZE_LAG], names[ex.ELEMSIZE_LAG_CART] = "eszL_eul", "eszL_cart"
        for (sym_e, sym_c, xkey, v_e, v_c) in ((ex.ELEMSIZE_EUL, ex.ELEMSIZE_EUL_CART, "xpos", "esz_eul", "esz_cart"),
                                               (ex.ELEMSIZE_LAG, ex.ELEMSIZE_LAG_CART, "xlag", "eszL_eul", "eszL_cart")):
            if not any(e_.has(sym_e) or e_.has(sym_c) for e_ in all_exprs if hasattr(e_, "has")):
                continue
            # element sizes (fill_shape_info_element_sizes, src/elements.cpp:3527-3568): sum over ALL integration points of w * J, with the
            # coordinate system's JacobianForElementSize (2 Pi r when axisymmetric) at the point for the non-Cartesian one; every
            # (element, point) thread forms the sum itself (a few hundred flops; only classes that use the symbols pay)
            if edim != dim or dim not in (2, 3):
                raise NotImplementedError("element sizes: bulk elements (two- or three-dimensional) only")
            axi = code.coordinate_system.get_id_name() == "Axisymmetric"
            w("      double %s = 0.0, %s = 0.0;" % (v_c, v_e))
            w("      for (int q = 0; q < %d; ++q)" % NIPT)
            w("      {")
            w("        const double* dq = s_dpsi2 + q * %d; const double* pq = s_psi2 + q * %d; (void)pq;" % (NN * edim, NN))
            if dim == 2:
                w("        double e00 = 0.0, e01 = 0.0, e10 = 0.0, e11 = 0.0, ex0 = 0.0;")
                w("        for (int l = 0; l < %d; ++l)" % NN)
                w("        {")
                w("          const double X0 = E[%d + l * 2], X1 = E[%d + l * 2 + 1];" % (plan[xkey], plan[xkey]))
                w("          e00 += X0 * dq[l * 2]; e01 += X1 * dq[l * 2]; e10 += X0 * dq[l * 2 + 1]; e11 += X1 * dq[l * 2 + 1]; ex0 += X0 * pq[l];")
                w("        }")
                w("        const double a00 = e00 * e00 + e01 * e01, a01 = e00 * e10 + e01 * e11, a11 = e10 * e10 + e11 * e11;")
                w("        const double Jq = c_w[q] * sqrt(a00 * a11 - a01 * a01);")
            else:
                # three dimensions: the metric a_ab = t_a . t_b of the three tangents and the root of its determinant, as
                # fill_shape_info_at_s forms it (src/elements.cpp:3801)
                w("        double e[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}}; const double ex0 = 0.0; (void)ex0;")
                w("        for (int l = 0; l < %d; ++l)" % NN)
                w("        {")
                w("          const double X0 = E[%d + l * 3], X1 = E[%d + l * 3 + 1], X2 = E[%d + l * 3 + 2];" % ((plan[xkey],) * 3))
                w("          #pragma unroll")
                w("          for (int b = 0; b < 3; ++b) { const double d = dq[l * 3 + b]; e[b][0] += X0 * d; e[b][1] += X1 * d; e[b][2] += X2 * d; }")
                w("        }")
                w("        double m[3][3];")
                w("        #pragma unroll")
                w("        for (int a_ = 0; a_ < 3; ++a_)")
                w("          #pragma unroll")
                w("          for (int b = 0; b < 3; ++b) m[a_][b] = e[a_][0] * e[b][0] + e[a_][1] * e[b][1] + e[a_][2] * e[b][2];")
                w("        const double detm = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])")
                w("                          + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);")
                w("        const double Jq = c_w[q] * sqrt(detm);")
            w("        %s += Jq; %s += Jq * %s;" % (v_c, v_e, "(2.0 * 3.14159265359 * ex0)" if axi else "1.0"))
            w("      }")
            w("      (void)%s; (void)%s;" % (v_c, v_e))
        # local-derivative sums per (field, kind)
        for (f, kind), derivs in needed.items():
            fld = code.fields[f]
            nn = self._nnode_space(fld.space)
            ps, dp = ("ps1", "dp1") if fld.space == "C1" else ("ps2", "dp2")
            if f.startswith("lagrangian_"):
                srcexpr = "E[%d + l * %d + %d]" % (plan["xlag"], dim, ex.DIRS.index(f[-1]))
            elif f.startswith("coordinate_") and kind == ("cur", 0):
                srcexpr = "E[%d + l * %d + %d]" % (plan["xpos"], dim, ex.DIRS.index(f[-1]))
            else:
                srcexpr = "E[%d + l]" % plan["src_off"][(f, kind)]
            tag = "%s_%s" % (f, "_".join(str(k) for k in kind))
            want_val = "d0" in derivs
            want_grad = any(d != "d0" for d in derivs)
            if (f, kind) not in fused:
                decl = []
                if want_val:
                    decl.append("v_%s = 0.0" % tag)
                if want_grad:
                    decl += ["s%d_%s = 0.0" % (b, tag) for b in range(edim)]
                w("      double %s;" % ", ".join(decl))
                w("      #pragma unroll%s" % self._node_unroll(nn))
                w("      for (int l = 0; l < %d; ++l) { const double u = %s;" % (nn, srcexpr))
                if want_val:
                    w("        v_%s += u * %s[l];" % (tag, ps))
                if want_grad:
                    for b in range(edim):
                        w("        s%d_%s += u * %s[l * %d + %d];" % (b, tag, dp, edim, b))
                w("      }")
            for d in sorted(derivs):
                at = AtomInfo(f, kind[1] if kind[0] == "dt" else 0, kind[2] if kind[0] == "dt" else "", d, kind[1] if kind[0] == "cur" else 0)
                sym = code.atom_symbol(at)
                cn = "A_%s_%s" % (tag, d)
                names[sym] = cn
                if d == "d0":
                    w("      const double %s = v_%s;" % (cn, tag))
                else:
                    g = "gg" if d[1] == "x" else "ggL"
                    i = int(d[2:])
                    w("      const double %s = %s;" % (cn, " + ".join("%s%d%d * s%d_%s" % (g, b, i, b, tag) for b in range(edim))))
        for sch in ("BDF1", "BDF2", "Newmark2", "BDF2_degr", "Newmark2_degr"):
            names[sp.Symbol("W__%s__1" % sch, real=True)] = "a.ti.w_dt_%s[0]" % sch
        names[sp.Symbol("W__Newmark2__2", real=True)] = "a.ti.w_d2t_Newmark2[0]"
        # coefficients through CSE
        exprs, targets = [], []
        for si, r in enumerate(form.R):
            exprs.append(r)
            targets.append(plan["R_off"] + si)
        if what >= 1:
            for k in sorted(form.J.keys()):
                exprs.append(form.J[k])
                targets.append(plan["J_off"][k])
        if what >= 2:
            for k in sorted(form.M.keys()):
                exprs.append(form.M[k])
                targets.append(plan["M_off"][k])
        repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("cse"), optimizations="basic")
        pr = _CudaPrinter(names)
        for s, e in repl:
            w("      const double %s = %s;" % (s.name, pr.doprint(e)))
        written = set()
        for tgt, e in zip(targets, red):
            if tgt in written:
                continue          # shared slot of identical coefficients
            written.add(tgt)
            w("      P[%d] = %s;" % (tgt, pr.doprint(e)))
        self._flops_phase1 = sum(int(sp.count_ops(e)) for _, e in repl) + sum(int(sp.count_ops(e)) for e in red)

    @staticmethod
    def _node_unroll(nn: int) -> str:
        """node loops of phase 1: fully unrolled for 2D (9 nodes); for 27-node bricks ptxas hoists every load of a fully unrolled
        loop (> 255 registers, accumulators of phase 2 end up in local memory), so the loop is unrolled by 3"""
        return "" if nn <= 9 else " %d" % int(os.environ.get("PB2_NODE_UNROLL", "3"))

    def _emit_geometry(self, o: List[str], plan, src: str, gname: str, detname: str, sums_done: bool = False):
        """Restates fill_shape_info_at_s (src/elements.cpp:3604-3626 tangents; metric, inverse and gab_gai for el_dim 1 :3651-3672,
        2 :3677-3703, 3 :3804-3836) with the same operation order; stores gab_gai[b][i] to the point block.  el_dim < nodal_dim
        (interface elements): the same metric form gives the surface gradient, sqrt(det) the surface measure."""
        dim, NN, edim = self.dim, self.NN, self.edim
        w = o.append
        t = "t_" + gname
        if not sums_done:
            w("      double %s;" % ", ".join("%s%d%d = 0.0" % (t, a, i) for a in range(edim) for i in range(dim)))
            w("      #pragma unroll%s" % self._node_unroll(NN))
            w("      for (int l = 0; l < %d; ++l) {" % NN)
            for i in range(dim):
                for a in range(edim):
                    w("        %s%d%d += E[%d + l * %d + %d] * dp2[l * %d + %d];" % (t, a, i, plan[src], dim, i, edim, a))
            w("      }")
        for al in range(edim):
            for be in range(edim):
                terms = ["%s%d%d * %s%d%d" % (t, al, i, t, be, i) for i in range(dim)]
                # amet += in order i=0.. starting from 0.0
                w("      const double am_%s%d%d = %s;" % (gname, al, be, " + ".join(terms)))
        am = lambda a, b: "am_%s%d%d" % (gname, a, b)
        if edim == 1:
            w("      const double det_%s = %s;" % (gname, am(0, 0)))
            w("      const double rdet_%s = 1.0 / det_%s;" % (gname, gname))
            up = {(0, 0): "rdet_%s" % gname}
        elif edim == 2:
            w("      const double det_%s = %s * %s - %s * %s;" % (gname, am(0, 0), am(1, 1), am(0, 1), am(1, 0)))
            # one IEEE reciprocal instead of the reference's four divisions by det (src/elements.cpp:3690-3703): <=1 ulp apart
            w("      const double rdet_%s = 1.0 / det_%s;" % (gname, gname))
            up = {(0, 0): "%s * rdet_%s" % (am(1, 1), gname), (0, 1): "-%s * rdet_%s" % (am(0, 1), gname),
                  (1, 0): "-%s * rdet_%s" % (am(1, 0), gname), (1, 1): "%s * rdet_%s" % (am(0, 0), gname)}
        else:
            w("      const double det_%s = %s * %s * %s + %s * %s * %s + %s * %s * %s - %s * %s * %s - %s * %s * %s - %s * %s * %s;" % (
                gname, am(0, 0), am(1, 1), am(2, 2), am(0, 1), am(1, 2), am(2, 0), am(0, 2), am(1, 0), am(2, 1),
                am(0, 0), am(1, 2), am(2, 1), am(0, 1), am(1, 0), am(2, 2), am(0, 2), am(1, 1), am(2, 0)))
            w("      const double rdet_%s = 1.0 / det_%s;" % (gname, gname))
            D = "rdet_" + gname
            up = {
                (0, 0): "(%s * %s - %s * %s) * %s" % (am(1, 1), am(2, 2), am(1, 2), am(2, 1), D),
                (0, 1): "-(%s * %s - %s * %s) * %s" % (am(0, 1), am(2, 2), am(0, 2), am(2, 1), D),
                (0, 2): "(%s * %s - %s * %s) * %s" % (am(0, 1), am(1, 2), am(0, 2), am(1, 1), D),
                (1, 0): "-(%s * %s - %s * %s) * %s" % (am(1, 0), am(2, 2), am(1, 2), am(2, 0), D),
                (1, 1): "(%s * %s - %s * %s) * %s" % (am(0, 0), am(2, 2), am(0, 2), am(2, 0), D),
                (1, 2): "-(%s * %s - %s * %s) * %s" % (am(0, 0), am(1, 2), am(0, 2), am(1, 0), D),
                (2, 0): "(%s * %s - %s * %s) * %s" % (am(1, 0), am(2, 1), am(1, 1), am(2, 0), D),
                (2, 1): "-(%s * %s - %s * %s) * %s" % (am(0, 0), am(2, 1), am(0, 1), am(2, 0), D),
                (2, 2): "(%s * %s - %s * %s) * %s" % (am(0, 0), am(1, 1), am(0, 1), am(1, 0), D),
            }
        for (al, be), e in up.items():
            w("      const double up_%s%d%d = %s;" % (gname, al, be, e))
        for b in range(edim):
            for i in range(dim):
                w("      const double %s%d%d = %s;" % (gname, b, i, " + ".join("up_%s%d%d * %s%d%d" % (gname, a_, b, t, a_, i) for a_ in range(edim))))
                w("      P[%d] = %s%d%d;" % (plan[gname] + b * dim + i, gname, b, i))
        if self.et.name.startswith("QuadFace") and gname == "gg":
            # face of a bulk element: the integral runs over s0 on the face s1 = -1, measure |dx/ds0|; outer normal of a counter-clockwise
            # element (t_y, -t_x)/|t| (the reference: FaceElement::outer_unit_normal with the bulk element's normal_sign)
            w("      const double %s = sqrt(am_%s00);" % (detname, gname))
            w("      const double nrm0 = %s01 / %s, nrm1 = -%s00 / %s; (void)nrm0; (void)nrm1;" % (t, detname, t, detname))
        elif self.et.name.startswith("QuadFace"):
            w("      const double %s = sqrt(am_%s00);" % (detname, gname))
        else:
            w("      const double %s = sqrt(det_%s);" % (detname, gname))
        if edim == 1 and dim == 2 and gname == "gg":
            # unit normal of a line element (BulkElementBase::get_normal_at_s, src/elements.cpp:1730-1752): (-t_y, t_x) / |t|
            w("      const double nrm_len = (det_gg < 1e-20) ? 1.0 : sqrt(det_gg);")
            w("      const double nrm0 = -%s01 / nrm_len, nrm1 = %s00 / nrm_len; (void)nrm0; (void)nrm1;" % (t, t))

    def _group_pairs(self, form: ResidualForm, g: RowGroup, coef):
        fields = [f for f in g.fields if any(s.field == f for s in form.slots)]
        pairs: Dict[Tuple[str, str], List[Tuple[int, str]]] = {}
        for (si, G, a_) in coef.keys():
            F = form.slots[si].field
            if F in fields:
                pairs.setdefault((F, G), []).append((si, a_))
        return fields, pairs

    def _group_acc_layout(self, form: ResidualForm, g: RowGroup, coef):
        """index of every accumulator of the group inside the kernel-wide acc[] array"""
        fields, pairs = self._group_pairs(form, g, coef)
        base: Dict[Tuple[str, str], int] = {}
        n = 0
        for F in fields:
            base[(F, "__res")] = n
            n += g.rows_per_thread
            for G in self.code.unknown_field_names():
                if (F, G) in pairs:
                    base[(F, G)] = n
                    n += g.rows_per_thread * self._nnode_space(self.code.fields[G].space)
        return base, n

    def _group_nacc(self, form, g, coef) -> int:
        return self._group_acc_layout(form, g, coef)[1]

    @staticmethod
    def _vectorise_point_reads(lines: List[str]) -> List[str]:
        """P[n] reads of the contraction become halves of 128-bit shared-memory loads (one wavefront serves two coefficients:
        the shared-memory pipe, not the fp64 pipe, is the busiest unit of the kernel, profiles/r01_notes.md)"""
        import re
        out: List[str] = []
        ptr_line = -1
        used = set()
        for ln in lines:
            if "const double* P = " in ln:
                ptr_line = len(out)
                out.append(ln)
                continue
            if ptr_line >= 0:
                def rep(m):
                    n = int(m.group(1))
                    used.add(n // 2)
                    return "pq%d.%s" % (n // 2, "xy"[n % 2])
                ln = re.sub(r"\bP\[(\d+)\]", rep, ln)
            out.append(ln)
        if ptr_line >= 0 and used:
            ind = " " * (len(out[ptr_line]) - len(out[ptr_line].lstrip()))
            loads = [ind + "const double2 pq%d = *reinterpret_cast<const double2*>(P + %d);" % (k, 2 * k) for k in sorted(used)]
            out[ptr_line + 1:ptr_line + 1] = loads
        return out

    @staticmethod
    def _vectorise_point_writes(lines: List[str], npoint: int) -> List[str]:
        """P[n] = expr of phase 1 become registers that are stored pairwise (128-bit) at the end of the point"""
        import re
        out: List[str] = []
        seen = set()
        ind = "      "
        for ln in lines:
            m = re.match(r"^(\s*)P\[(\d+)\] = (.*);$", ln)
            if m:
                ind = m.group(1)
                seen.add(int(m.group(2)))
                out.append("%sconst double pw%s = %s;" % (m.group(1), m.group(2), m.group(3)))
            else:
                out.append(ln)
        for k in range((npoint + 1) // 2):
            a_, b_ = ("pw%d" % (2 * k) if 2 * k in seen else "0.0"), ("pw%d" % (2 * k + 1) if 2 * k + 1 in seen else "0.0")
            if 2 * k in seen or 2 * k + 1 in seen:
                out.append("%s*reinterpret_cast<double2*>(P + %d) = make_double2(%s, %s);" % (ind, 2 * k, a_, b_))
        return out

    def _emit_group_compute(self, o: List[str], rp: RoutinePlan, plan, g: RowGroup, pname, coef, coff, with_res):
        code, dim, NIPT = self.code, self.dim, self.NIPT
        form = rp.form
        ELS, EL0, PB = plan["ELS"], plan["EL0"], plan["PB"]
        w = o.append
        nn_g = self._nnode_space(g.space)
        RB, TPE = g.rows_per_thread, g.threads_per_elem
        fields, pairs = self._group_pairs(form, g, coef)
        base, nacc = self._group_acc_layout(form, g, coef)
        unknowns = code.unknown_field_names()
        if not fields:
            return
        w("    if (tid >= %d && tid < %d)" % (g.thread_off, g.thread_off + g.nthreads))
        w("    {")
        w("      const int tl = tid - %d;" % g.thread_off)
        w("      const int el = tl / %d, q = tl - el * %d;" % (TPE, TPE))
        w("      if (el < nel)")
        w("      {")
        w("        #pragma unroll")
        w("        for (int i = 0; i < %d; ++i) acc[i] = 0.0;" % nacc)
        sf = self.sum_factorise and RB == 1 and bool(pairs) and all(code.fields[G].space != "C1" for (F_, G) in pairs)
        sf_pairs: List[Tuple[str, bool, bool, int]] = []
        sf2 = self.sum_factorise_2d and RB == 1 and any(code.fields[G].space != "C1" for (F_, G) in pairs)
        sf2_pairs: List[Tuple[str, bool, bool, int, int]] = []
        if sf2:
            # Gauss point (sp, sq) = ipt sp*3 + sq (s0 outer); the s1 direction is contracted point by point into X0 / X1, the s0 direction
            # once per sp
            w("        #pragma unroll 1")
            w("        for (int sp = 0; sp < 3; ++sp)")
            w("        {")
            for (F_, G_) in pairs:
                if code.fields[G_].space != "C1":
                    w("          double sfX0_%s_%s[3] = {0.0, 0.0, 0.0}, sfX1_%s_%s[3] = {0.0, 0.0, 0.0};" % (F_, G_, F_, G_))
            w("          #pragma unroll")
            w("          for (int sq = 0; sq < 3; ++sq)")
            w("        {")
            w("          const int ipt = sp * 3 + sq;")
        elif sf:
            # Gauss point (sp, sq, sr) = ipt sp*9 + sq*3 + sr; sr is contracted point by point into U, sq and sp after the inner loops
            w("        #pragma unroll 1")
            w("        for (int sp = 0; sp < 3; ++sp)")
            w("        {")
            for (F_, G_) in pairs:
                w("          double sfU03_%s_%s[9], sfU1_%s_%s[9], sfU2_%s_%s[9];" % (F_, G_, F_, G_, F_, G_))
                w("          #pragma unroll")
                w("          for (int i = 0; i < 9; ++i) { sfU03_%s_%s[i] = 0.0; sfU1_%s_%s[i] = 0.0; sfU2_%s_%s[i] = 0.0; }" % (F_, G_, F_, G_, F_, G_))
            w("          #pragma unroll")
            w("          for (int sq = 0; sq < 3; ++sq)")
            w("          #pragma unroll 1")
            w("          for (int sr = 0; sr < 3; ++sr)")
            w("        {")
            w("          const int ipt = sp * 9 + sq * 3 + sr;")
        else:
            if self.ipt_unroll >= NIPT:
                w("        #pragma unroll")
            else:
                w("        #pragma unroll %d" % self.ipt_unroll)
            w("        for (int ipt = 0; ipt < %d; ++ipt)" % NIPT)
            w("        {")
        w("          const double* P = %s + ipt * %d;" % (plan.get("PT_expr", "s_el + el * %d + %d" % (ELS, EL0)), PB))
        need_x = any(s.deriv.startswith("dx") for s in form.slots if s.field in fields) or any(a_.startswith("dx") for (F, G), l in pairs.items() for (_, a_) in l)
        need_X = any(s.deriv.startswith("dX") for s in form.slots if s.field in fields) or any(a_.startswith("dX") for (F, G), l in pairs.items() for (_, a_) in l)
        if need_x:
            w("          " + " ".join("const double gg%d%d = P[%d];" % (b, i, plan["gg"] + b * dim + i) for b in range(self.edim) for i in range(dim)))
        if need_X:
            w("          " + " ".join("const double ggL%d%d = P[%d];" % (b, i, plan["ggL"] + b * dim + i) for b in range(self.edim) for i in range(dim)))
        if sf2:
            w("          " + " ".join("const double tL1%d = c_t1d[ipt * 12 + %d]; const double tD1%d = c_t1d[ipt * 12 + %d];" % (n, 6 + n, n, 9 + n) for n in range(3)))
        tp = self.tensor_columns and any(code.fields[G].space != "C1" for (F_, G) in pairs)
        if tp:
            for d in ((2,) if sf else range(3)):
                w("          " + " ".join("const double tL%d%d = c_t1d[ipt * 18 + %d]; const double tD%d%d = c_t1d[ipt * 18 + %d];" % (
                    d, n, d * 6 + n, d, n, d * 6 + 3 + n) for n in range(3)))
        w("          #pragma unroll")
        w("          for (int k = 0; k < %d; ++k)" % RB)
        w("          {")
        w("            const int lt = q + k * %d;" % TPE)
        sps, sdp = ("s_psi1", "s_dpsi1") if g.space == "C1" else ("s_psi2", "s_dpsi2")
        w("            const double T_d0 = %s[ipt * %d + lt]; (void)T_d0;" % (sps, nn_g))
        test_derivs = {s.deriv for s in form.slots if s.field in fields}
        if any(d != "d0" for d in test_derivs):
            for b in range(self.edim):
                w("            const double Ts%d = %s[(ipt * %d + lt) * %d + %d];" % (b, sdp, nn_g, self.edim, b))
        for d in sorted(test_derivs):
            if d == "d0":
                continue
            gn = "gg" if d[1] == "x" else "ggL"
            i = int(d[2:])
            w("            const double T_%s = %s;" % (d, " + ".join("%s%d%d * Ts%d" % (gn, b, i, b) for b in range(self.edim))))
        for F in fields:
            fslots = [(si, s) for si, s in enumerate(form.slots) if s.field == F]
            if with_res:
                terms = ["T_%s * P[%d]" % (s.deriv, plan["R_off"] + si) for si, s in fslots if form.R[si] != 0]
                if terms:
                    expr = "acc[%d + k]" % base[(F, "__res")]
                    for si, sl in fslots:
                        if form.R[si] != 0:
                            expr = "fma(T_%s, P[%d], %s)" % (sl.deriv, plan["R_off"] + si, expr)
                    w("            acc[%d + k] = %s;" % (base[(F, "__res")], expr))
            for G in unknowns:
                if (F, G) not in pairs:
                    continue
                Gs = code.fields[G].space
                nnG = self._nnode_space(Gs)
                by_atom: Dict[str, List[int]] = {}
                for (si, a_) in pairs[(F, G)]:
                    by_atom.setdefault(a_, []).append(si)
                for a_, sis in sorted(by_atom.items()):
                    w("            const double W_%s_%s_%s = %s;" % (F, G, a_, " + ".join(
                        "T_%s * P[%d]" % (form.slots[si].deriv, coff[(si, G, a_)]) for si in sis)))
                have_s = False
                sterms = {b: [] for b in range(self.edim)}
                for a_ in sorted(by_atom):
                    if a_ == "d0":
                        continue
                    gn = "gg" if a_[1] == "x" else "ggL"
                    i = int(a_[2:])
                    for b in range(self.edim):
                        sterms[b].append("W_%s_%s_%s * %s%d%d" % (F, G, a_, gn, b, i))
                    have_s = True
                if have_s:
                    for b in range(self.edim):
                        w("            const double Ws%d_%s_%s = %s;" % (b, F, G, " + ".join(sterms[b])))
                if Gs == "C1":
                    tp, td = ("s_psi1", "s_dpsi1") if self.table_source == "smem" else ("c_psi1", "c_dpsi1")
                else:
                    tp, td = ("s_psi2", "s_dpsi2") if self.table_source == "smem" else ("c_psi2", "c_dpsi2")
                if sf2 and Gs != "C1":
                    # X0[b] += W0 L_b(q) + Ws1 L'_b(q);  X1[b] += Ws0 L_b(q)   (b: basis index of the s1 direction)
                    pf = "%s_%s" % (F, G)
                    w0 = "W_%s_d0" % pf if "d0" in by_atom else None
                    for b in range(3):
                        terms = ([("%s * tL1%d" % (w0, b))] if w0 else []) + ([("Ws1_%s * tD1%d" % (pf, b))] if have_s else [])
                        w("            sfX0_%s[%d] += %s;" % (pf, b, " + ".join(terms)))
                        if have_s:
                            w("            sfX1_%s[%d] = fma(Ws0_%s, tL1%d, sfX1_%s[%d]);" % (pf, b, pf, b, pf, b))
                    sf2_pairs.append((pf, have_s, base[(F, G)], nnG, 0))
                    continue
                if sf:
                    # contraction of the third direction, point by point: U03 collects what is later multiplied by L_b(q) L_a(p),
                    # U1 by L_b(q) L'_a(p), U2 by L'_b(q) L_a(p)
                    pf = "%s_%s" % (F, G)
                    w0 = "W_%s_d0" % pf if "d0" in by_atom else None
                    for c in range(3):
                        terms = ([("%s * tL2%d" % (w0, c))] if w0 else []) + ([("Ws2_%s * tD2%d" % (pf, c))] if have_s else [])
                        w("            sfU03_%s[sq * 3 + %d] += %s;" % (pf, c, " + ".join(terms)))
                        if have_s:
                            w("            sfU1_%s[sq * 3 + %d] = fma(Ws0_%s, tL2%d, sfU1_%s[sq * 3 + %d]);" % (pf, c, pf, c, pf, c))
                            w("            sfU2_%s[sq * 3 + %d] = fma(Ws1_%s, tL2%d, sfU2_%s[sq * 3 + %d]);" % (pf, c, pf, c, pf, c))
                    sf_pairs.append((pf, have_s, base[(F, G)], nnG))
                    continue
                if self.tensor_columns and Gs != "C1":
                    # psi_c = a_i b_j c_k: W0 psi + Ws.dpsi = (W0 a_i + Ws0 a'_i) b_j c_k + a_i (Ws1 b'_j c_k + Ws2 b_j c'_k)
                    pf = "%s_%s" % (F, G)
                    w0 = "W_%s_d0" % pf if "d0" in by_atom else None
                    for i in range(3):
                        terms = ([("%s * tL0%d" % (w0, i))] if w0 else []) + ([("Ws0_%s * tD0%d" % (pf, i))] if have_s else [])
                        w("            const double tpA%d_%s = %s;" % (i, pf, " + ".join(terms)))
                    for j in range(3):
                        for i in range(3):
                            w("            const double tpAB%d%d_%s = tpA%d_%s * tL1%d;" % (i, j, pf, i, pf, j))
                    if have_s:
                        for j in range(3):
                            w("            const double tpX%d_%s = Ws1_%s * tD1%d; const double tpY%d_%s = Ws2_%s * tL1%d;" % (j, pf, pf, j, j, pf, pf, j))
                        for k_ in range(3):
                            for j in range(3):
                                w("            const double tpC%d%d_%s = tL2%d * tpX%d_%s + tD2%d * tpY%d_%s;" % (j, k_, pf, k_, j, pf, k_, j, pf))
                    for k_ in range(3):
                        for j in range(3):
                            for i in range(3):
                                c = i + 3 * j + 9 * k_
                                an = "acc[%d + k * %d + %d]" % (base[(F, G)], nnG, c)
                                e = "fma(tpAB%d%d_%s, tL2%d, %s)" % (i, j, pf, k_, an)
                                if have_s:
                                    e = "fma(tL0%d, tpC%d%d_%s, %s)" % (i, j, k_, pf, e)
                                w("            %s = %s;" % (an, e))
                    continue
                parts = []
                if "d0" in by_atom:
                    parts.append(("W_%s_%s_d0" % (F, G), "%s[ipt * %d + c]" % (tp, nnG)))
                if have_s:
                    for b in range(self.edim):
                        parts.append(("Ws%d_%s_%s" % (b, F, G), "%s[(ipt * %d + c) * %d + %d]" % (td, nnG, self.edim, b)))
                accname = "acc[%d + k * %d + c]" % (base[(F, G)], nnG)
                expr = accname
                for (wn, tn) in parts:          # one DFMA per term, accumulated in place
                    expr = "fma(%s, %s, %s)" % (wn, tn, expr)
                w("            #pragma unroll")
                w("            for (int c = 0; c < %d; ++c)" % nnG)
                w("              %s = %s;" % (accname, expr))
        w("          }")
        w("        }")
        if sf2:
            # first direction: J[a + 3b] += L_a(p) X0[b] + L'_a(p) X1[b] with the factors of the s0 knot of this sp
            w("          {")
            w("            const int k = 0; (void)k;")
            w("            " + " ".join("const double sfLa%d = c_t1d[sp * 36 + %d]; const double sfDa%d = c_t1d[sp * 36 + %d];" % (a_, a_, a_, 3 + a_) for a_ in range(3)))
            for (pf, have_s, b0, nnG, _) in sf2_pairs:
                for b in range(3):
                    for a_ in range(3):
                        an = "acc[%d + k * %d + %d]" % (b0, nnG, a_ + 3 * b)
                        e = "fma(sfX0_%s[%d], sfLa%d, %s)" % (pf, b, a_, an)
                        if have_s:
                            e = "fma(sfX1_%s[%d], sfDa%d, %s)" % (pf, b, a_, e)
                        w("            %s = %s;" % (an, e))
            w("          }")
            w("        }")
        if sf:
            # second direction (constant factors, q unrolled), then first direction (factors of plane sp)
            w("          {")
            w("            const int k = 0; (void)k;")
            w("            " + " ".join("const double sfLa%d = c_t1d[sp * 162 + %d]; const double sfDa%d = c_t1d[sp * 162 + %d];" % (a_, a_, a_, 3 + a_) for a_ in range(3)))
            for (pf, have_s, b0, nnG) in sf_pairs:
                for c in range(3):
                    for b in range(3):
                        m_terms = ["sfU03_%s[%d] * c_t1d[%d]" % (pf, q_ * 3 + c, q_ * 54 + 6 + b) for q_ in range(3)]
                        if have_s:
                            m_terms += ["sfU2_%s[%d] * c_t1d[%d]" % (pf, q_ * 3 + c, q_ * 54 + 9 + b) for q_ in range(3)]
                        w("            const double sfM%d%d_%s = %s;" % (b, c, pf, " + ".join(m_terms)))
                        if have_s:
                            w("            const double sfV%d%d_%s = %s;" % (b, c, pf, " + ".join("sfU1_%s[%d] * c_t1d[%d]" % (pf, q_ * 3 + c, q_ * 54 + 6 + b) for q_ in range(3))))
                for c in range(3):
                    for b in range(3):
                        for a_ in range(3):
                            an = "acc[%d + k * %d + %d]" % (b0, nnG, a_ + 3 * b + 9 * c)
                            e = "fma(sfM%d%d_%s, sfLa%d, %s)" % (b, c, pf, a_, an)
                            if have_s:
                                e = "fma(sfV%d%d_%s, sfDa%d, %s)" % (b, c, pf, a_, e)
                            w("            %s = %s;" % (an, e))
            w("          }")
            w("        }")
        w("      }")
        w("    }")

    def _emit_group_stage(self, o: List[str], rp: RoutinePlan, plan, g: RowGroup, pname, coef, with_res, with_matrix):
        """accumulators -> dense element matrix/residual in shared memory (structurally empty blocks are zeros: the
        fixed CSR pattern still owns those entries)"""
        code = self.code
        form = rp.form
        w = o.append
        ND = self.ndof
        RB, TPE = g.rows_per_thread, g.threads_per_elem
        fields, pairs = self._group_pairs(form, g, coef)
        base, nacc = self._group_acc_layout(form, g, coef)
        unknowns = code.unknown_field_names()
        w("    if (tid >= %d && tid < %d)" % (g.thread_off, g.thread_off + g.nthreads))
        w("    {")
        w("      const int tl = tid - %d;" % g.thread_off)
        w("      const int el = tl / %d, q = tl - el * %d;" % (TPE, TPE))
        w("      if (el < nel)")
        w("      {")
        w("        double* SJ = %s;" % plan.get("SJ_expr", "s_el + el * %d + %d" % (plan["ELS"], plan["SJ_off"])))
        w("        #pragma unroll")
        w("        for (int k = 0; k < %d; ++k)" % RB)
        w("        {")
        w("          const int lt = q + k * %d;" % TPE)
        for F in g.fields:
            w("          {")
            w("            const int row = c_row_%s[lt];" % F)
            if with_res:
                roff = ND * plan.get("RS", ND) if with_matrix else 0
                if F in fields:
                    w("            SJ[%d + row] = acc[%d + k];" % (roff, base[(F, "__res")]))
                else:
                    w("            SJ[%d + row] = 0.0;" % roff)
            if with_matrix:
                w("            double* srow = SJ + row * %d;" % plan.get("RS", ND))
                for G in unknowns:
                    nnG = self._nnode_space(code.fields[G].space)
                    for c in range(nnG):
                        col = self._col_index(G, c)
                        if (F, G) in pairs:
                            w("            srow[%d] = acc[%d + k * %d + %d];" % (col, base[(F, G)], nnG, c))
                        else:
                            w("            srow[%d] = 0.0;" % col)
            w("          }")
        w("        }")
        w("      }")
        w("    }")

    def _emit_scatter(self, o: List[str], plan, target, with_res):
        w = o.append
        ND, ND2 = self.ndof, self.ndof * self.ndof
        ELS, SJ = plan["ELS"], plan["SJ_off"]
        w("    {")
        if target is not None:
            w("      if (a.map_bits == 8)")
            w("        pb2_scatter_matrix<unsigned char, 0x80u, 0xFFu, %d, %d, %d, %d>((const unsigned char*)s_map, s_rowstart, s_el + %d, %s, nel, tid);" % (
                ND, ELS, self.NT, self.EPB, SJ, target))
            w("      else")
            w("        pb2_scatter_matrix<unsigned short, 0x8000u, 0xFFFFu, %d, %d, %d, %d>((const unsigned short*)s_map, s_rowstart, s_el + %d, %s, nel, tid);" % (
                ND, ELS, self.NT, self.EPB, SJ, target))
        if with_res:
            w("      for (int idx = tid; idx < nel * %d; idx += %d)" % (ND, self.NT))
            w("      {")
            w("        const int el = idx / %d, row = idx - el * %d;" % (ND, ND))
            w("        pb2_put(a.residual, s_resmap[idx], s_el[el * %d + %d + row]);" % (ELS, SJ + (ND2 if target is not None else 0)))
            w("      }")
        w("    }")

    # ------------------------------------------------------------------ whole file
    def emit(self) -> str:
        import os
        code = self.code
        self._kernel_smem: Dict[str, int] = {}
        o: List[str] = []
        w = o.append
        w("// generated by pyoomph_b200.cuda_emitter for element class '%s' (%s) -- sm_100a" % (self.name, self.et.name))
        w("#include <cuda_runtime.h>")
        w("#include <string.h>")
        w("#include <stdlib.h>")
        w('#include "pb2_jit_cuda.h"')
        w("")
        w("static __device__ __forceinline__ void pb2_cp_async4(void* smem_dst, const void* gsrc)")
        w("{")
        w("  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);")
        w("  asm volatile(\"cp.async.ca.shared.global [%0], [%1], 4;\" :: \"r\"(sa), \"l\"(gsrc) : \"memory\");")
        w("}")
        w("static __device__ __forceinline__ void pb2_cp_async8(void* smem_dst, const void* gsrc)")
        w("{")
        w("  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);")
        w("  asm volatile(\"cp.async.ca.shared.global [%0], [%1], 8;\" :: \"r\"(sa), \"l\"(gsrc) : \"memory\");")
        w("}")
        w("static __device__ __forceinline__ void pb2_bar_sync(const int id, const int nthreads) { asm volatile(\"bar.sync %0, %1;\" :: \"r\"(id), \"r\"(nthreads) : \"memory\"); }")
        w("static __device__ __forceinline__ void pb2_bar_arrive(const int id, const int nthreads) { asm volatile(\"bar.arrive %0, %1;\" :: \"r\"(id), \"r\"(nthreads) : \"memory\"); }")
        w("// tile gate of the persistent kernels: wait until `done` reaches `need`.  The launch is cooperative (all blocks resident, or the")
        w("// launch fails), so the gate opens; the deadline only turns an impossible wait (a lost block) into an error the host reports")
        w("static __device__ __forceinline__ void pb2_gate_wait(const int* done, const int need, int* status)")
        w("{")
        w("  if (*(volatile const int*)done >= need) return;")
        w("  unsigned long long t0; asm volatile(\"mov.u64 %0, %%globaltimer;\" : \"=l\"(t0));")
        w("  unsigned spins = 0;")
        w("  while (*(volatile const int*)done < need)")
        w("  {")
        w("    __nanosleep(64);")
        w("    if ((++spins & 4095u) != 0u) continue;     // the deadline and the error word (host memory) are looked at every ~0.3 ms only")
        w("    unsigned long long t1; asm volatile(\"mov.u64 %0, %%globaltimer;\" : \"=l\"(t1));")
        w("    if (t1 - t0 > PB2_GATE_TIMEOUT_NS || *(volatile int*)status != 0) { atomicExch(status, PB2_STATUS_GATE_TIMEOUT); break; }")
        w("  }")
        w("}")
        w("// branch-free: first touch of a CSR entry stores, later colours reduce (fire-and-forget red, no return value)")
        w("static __device__ __forceinline__ void pb2_store_or_red(double* dst, const double v, const bool do_store, const bool do_red)")
        w("{")
        w("  asm volatile(\"{\\n\\t.reg .pred ps, pr;\\n\\tsetp.ne.u32 ps, %2, 0;\\n\\tsetp.ne.u32 pr, %3, 0;\\n\\t@ps st.global.f64 [%0], %1;\\n\\t@pr red.global.add.f64 [%0], %1;\\n\\t}\" :: \"l\"(dst), \"d\"(v), \"r\"((unsigned)do_store), \"r\"((unsigned)do_red) : \"memory\");")
        w("}")
        w("static __device__ __forceinline__ void pb2_put(double* __restrict__ dst, const int p, const double v)")
        w("{")
        w("  pb2_store_or_red(dst + (p >= 0 ? p : ~p), v, p < 0 && p != PB2_MAP_SKIP, p >= 0);")
        w("}")
        w("// cooperative scatter of dense element matrices staged in shared memory (element stride ELS doubles) into the CSR")
        w("// value array: entry idx of the batch <-> byte idx of the position map, so map reads are perfectly coalesced and")
        w("// neighbouring lanes hit neighbouring columns of the same CSR row.")
        w("// one CSR entry: first touch stores, later colours reduce; pinned rows/columns are skipped.  Hand-written so that the")
        w("// whole decision is predicates (no branches, no bool materialisation): 9 instructions next to the 3 shared loads.")
        w("template <unsigned FIRST, unsigned SKIP>")
        w("static __device__ __forceinline__ void pb2_scatter_entry(double* vals, const unsigned code, const int r0, const double v)")
        w("{")
        w("  asm volatile(\"{\\n\\t.reg .pred pok, ps, pr;\\n\\t.reg .b32 off, fb;\\n\\t.reg .s32 pos;\\n\\t.reg .b64 ad;\\n\\t\"")
        w("               \"setp.ne.u32 pok, %2, %5;\\n\\tsetp.ge.and.s32 pok, %3, 0, pok;\\n\\t\"")
        w("               \"and.b32 off, %2, %6;\\n\\tadd.s32 pos, %3, off;\\n\\tmad.wide.s32 ad, pos, 8, %0;\\n\\t\"")
        w("               \"and.b32 fb, %2, %4;\\n\\tsetp.ne.and.u32 ps, fb, 0, pok;\\n\\tsetp.eq.and.u32 pr, fb, 0, pok;\\n\\t\"")
        dbg = os.environ.get("PB2_DEBUG_SCATTER", "")
        if dbg == "nostore":      # development experiment: no global writes at all (results are wrong)
            w("               \"@ps add.f64 %1, %1, %1;\\n\\t}\"")
        elif dbg == "nored":      # development experiment: reductions replaced by plain stores (results are wrong)
            w("               \"@ps st.global.f64 [ad], %1;\\n\\t@pr st.global.f64 [ad], %1;\\n\\t}\"")
        else:
            w("               \"@ps st.global.f64 [ad], %1;\\n\\t@pr red.global.add.f64 [ad], %1;\\n\\t}\"")
        w("               :: \"l\"(vals), \"d\"(v), \"r\"(code), \"r\"(r0), \"n\"(FIRST), \"n\"(SKIP), \"n\"(FIRST - 1u) : \"memory\");")
        w("}")
        w("template <typename MapT, unsigned FIRST, unsigned SKIP, int ND, int ELS, int NT, int EPB>")
        w("static __device__ __forceinline__ void pb2_scatter_matrix(const MapT* __restrict__ mp, const int* __restrict__ rowstart,")
        w("                                                          const double* __restrict__ sj, double* __restrict__ vals, const int nel, const int tid)")
        w("{")
        w("  // all operands are in shared memory; stores and reductions are fire-and-forget (colouring => one add per entry and launch,")
        w("  // stream order between launches => the summation order is fixed).  Lane <-> consecutive (row,col) entries of one element,")
        w("  // so neighbouring lanes hit neighbouring columns of the same CSR row.  A thread keeps its (row,col) slots for all elements")
        w("  // of the batch: the element loop is unrolled and every shared-memory address is base + immediate.")
        w("  constexpr int ND2 = ND * ND, NJ = (ND2 + NT - 1) / NT;")
        w("  const MapT* m[NJ]; const int* rs[NJ]; const double* sv[NJ];")
        w("  #pragma unroll")
        w("  for (int j = 0; j < NJ; ++j)")
        w("  {")
        w("    const int k = min(tid + j * NT, ND2 - 1);   // clamped slots are masked below")
        w("    m[j] = mp + k; rs[j] = rowstart + k / ND; sv[j] = sj + k;")
        w("  }")
        w("  #pragma unroll")
        w("  for (int el = 0; el < EPB; ++el)")
        w("  {")
        w("    if (el < nel)")
        w("    {")
        w("      #pragma unroll")
        w("      for (int j = 0; j < NJ; ++j)")
        w("      {")
        w("        const bool live = (j + 1 < NJ) || (tid + j * NT < ND2);")
        w("        const unsigned code = live ? (unsigned)m[j][el * ND2] : SKIP;")
        w("        pb2_scatter_entry<FIRST, SKIP>(vals, code, rs[j][el * ND], sv[j][el * ELS]);")
        w("      }")
        w("    }")
        w("  }")
        w("}")
        w("// scatter of one batch by the scatter warps of the pipelined kernels: matrix entries as above, the residual entry of row r by")
        w("// thread NT-1-r.  Elements of a batch may share CSR entries: bit el of bmask = 'synchronise the scatter warps before element el'")
        w("// (bar.sync orders the first-touch store of one thread before the reduction of another: same block, same address), so every")
        w("// entry receives its contributions in element order and neighbouring elements complete it while the line is still in L2.")
        w("template <typename MapT, unsigned FIRST, unsigned SKIP, int ND, int RS, int ELS, int NT, int EPB, bool MAT, bool RES, int ROFF>")
        w("static __device__ __forceinline__ void pb2_scatter_batch(const MapT* __restrict__ mp, const int* __restrict__ rowstart, const int* __restrict__ resmap,")
        w("                                                         const double* __restrict__ sj, double* __restrict__ vals, double* __restrict__ residual,")
        w("                                                         const int nel, const int tid, const unsigned long long bmask)")
        w("{")
        w("  constexpr int ND2 = ND * ND, NJ = (ND2 + NT - 1) / NT;")
        w("  const MapT* m[NJ]; const int* rs[NJ]; const double* sv[NJ];")
        w("  #pragma unroll")
        w("  for (int j = 0; j < NJ; ++j)")
        w("  {")
        w("    const int k = min(tid + j * NT, ND2 - 1);   // clamped slots are masked below")
        w("    m[j] = mp + k; rs[j] = rowstart + k / ND; sv[j] = sj + (k / ND) * RS + k % ND;")
        w("  }")
        w("  const int rrow = NT - 1 - tid;   // the residual rows go to the threads with the fewest matrix slots")
        w("  #pragma unroll")
        w("  for (int el = 0; el < EPB; ++el)")
        w("  {")
        w("    if (el < nel)")
        w("    {")
        w("      if ((bmask >> el) & 1ull) pb2_bar_sync(11, NT);")
        w("      if (MAT)")
        w("      {")
        w("        #pragma unroll")
        w("        for (int j = 0; j < NJ; ++j)")
        w("        {")
        w("          const bool live = (j + 1 < NJ) || (tid + j * NT < ND2);")
        w("          const unsigned code = live ? (unsigned)m[j][el * ND2] : SKIP;")
        w("          pb2_scatter_entry<FIRST, SKIP>(vals, code, rs[j][el * ND], sv[j][el * ELS]);")
        w("        }")
        w("      }")
        w("      if (RES)")
        w("      {")
        w("        #pragma unroll")
        w("        for (int r = rrow; r < ND; r += NT) pb2_put(residual, resmap[el * ND + r], sj[el * ELS + ROFF + r]);")
        w("      }")
        w("    }")
        w("  }")
        w("}")
        w("")
        self._emit_tables(o)
        kernels: Dict[Tuple[str, int], str] = {}
        for rp in self.routines:
            for what in (0, 1, 2):
                kernels[(rp.key, what)] = self._emit_kernel_pipe(o, rp, what) if self.pipeline else self._emit_kernel(o, rp, what)
        if self.hessian and self.pipeline:
            for rp in self.hroutines:
                for what in (1, 2):
                    kernels[(rp.key, what)] = self._emit_kernel_pipe(o, rp, what)
        integral_kernel = self._emit_integral_kernel(o) if self.code.integral_expressions else None
        point_kernels = [self._emit_point_kernel(o, False), self._emit_point_kernel(o, True)] if self.code.point_expression_names() else []
        # host side: launchers + table
        w("static int pb2_query(int kind, int residual_index, int param_index, unsigned flag, pb2_kernel_cfg* out)")
        w("{")
        w("  memset(out, 0, sizeof(*out));")
        w("  if (kind < 0 || kind > 4 || flag > 2u) return 1;")
        if integral_kernel:
            w("  if (kind == 2) { out->func = (const void*)%s; out->smem_bytes = %d; out->elems_per_batch = %d; out->threads = %d; }" % (
                integral_kernel, self._kernel_smem[integral_kernel], self._kernel_cfg[integral_kernel][0], self._kernel_cfg[integral_kernel][1]))
        w("  if (kind == 2 && !out->func) return 2;")
        for fl, kn in enumerate(point_kernels):
            w("  if (kind == 4 && flag == %du) { out->func = (const void*)%s; out->smem_bytes = %d; out->elems_per_batch = %d; out->threads = %d; }" % (
                fl, kn, self._kernel_smem[kn], self._kernel_cfg[kn][0], self._kernel_cfg[kn][1]))
        w("  if (kind == 4 && !out->func) return 2;")
        if self.hessian and self.pipeline:
            for rp in self.hroutines:
                for what in (1, 2):
                    kn = kernels[(rp.key, what)]
                    w("  if (kind == %d && residual_index == %d && flag == %du) { out->func = (const void*)%s; out->smem_bytes = %d; out->elems_per_batch = %d; out->threads = %d; }" % (
                        3 if rp.key.startswith("ht") else 1, rp.res_index, what, kn, self._kernel_smem[kn], self._kernel_cfg[kn][0], self._kernel_cfg[kn][1]))
        for rp in self.routines:
            for what in (0, 1, 2):
                kn = kernels[(rp.key, what)]
                w("  if (kind == 0 && residual_index == %d && param_index == %d && flag == %du) { out->func = (const void*)%s; out->smem_bytes = %d; out->elems_per_batch = %d; out->threads = %d; }" % (
                    rp.res_index, rp.param_index, what, kn, self._kernel_smem[kn], self._kernel_cfg[kn][0], self._kernel_cfg[kn][1]))
        w("  if (!out->func) return 2;")
        w("  out->pipelined = (kind == 2 || kind == 4) ? 0 : %d;   // kinds 0, 1, 3 are persistent pipelined kernels" % (1 if self.pipeline else 0))
        w("  cudaError_t err = cudaFuncSetAttribute(out->func, cudaFuncAttributeMaxDynamicSharedMemorySize, out->smem_bytes);")
        w("  if (err != cudaSuccess) return 100 + (int)err;")
        w("  int per_sm = 0;")
        w("  err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, out->func, out->threads, (size_t)out->smem_bytes);")
        w("  if (err != cudaSuccess) return 100 + (int)err;")
        w("  if (per_sm < 1) return 3;")
        w("  out->blocks_per_sm = per_sm;")
        w("  return 0;")
        w("}")
        w("")
        w("static int pb2_launch(const pb2_kernel_cfg* cfg, const pb2_kernel_args* args, int grid, void* stream)")
        w("{")
        w("  if (grid <= 0) return 0;")
        w("  void* params[1] = {(void*)args};")
        w("  // persistent kernels wait on each other at the tile gates: cooperative launch = every block resident at once, or an error")
        w("  static const int coop = getenv(\"PB2_COOP\") ? atoi(getenv(\"PB2_COOP\")) : 1;   // development switch: 0 = plain launch (no residency guarantee)")
        w("  cudaError_t err = (cfg->pipelined && coop)")
        w("    ? cudaLaunchCooperativeKernel(cfg->func, dim3((unsigned)grid), dim3((unsigned)cfg->threads), params, (size_t)cfg->smem_bytes, (cudaStream_t)stream)")
        w("    : cudaLaunchKernel(cfg->func, dim3((unsigned)grid), dim3((unsigned)cfg->threads), params, (size_t)cfg->smem_bytes, (cudaStream_t)stream);")
        w("  if (err == cudaSuccess) err = cudaGetLastError();")
        w("  return err == cudaSuccess ? 0 : 100 + (int)err;")
        w("}")
        w("")
        w('extern "C" void JIT_ELEMENT_init_cuda(pb2_cuda_table_t* table)')
        w("{")
        w("  memset(table, 0, sizeof(*table));")
        w("  pb2_class_info* ci = &table->info;")
        w("  ci->abi_version = PB2_ABI_VERSION;")
        w('  strncpy(ci->name, "%s", sizeof(ci->name) - 1);' % self.name)
        w("  ci->nodal_dim = %d; ci->elem_dim = %d; ci->nnode = %d; ci->nnode_C1 = %d; ci->n_int_pt = %d;" % (
            self.dim, self.et.elem_dim, self.NN, self.NN1, self.NIPT))
        for i, n in enumerate(self.et.c1_nodes):
            w("  ci->c1_nodes[%d] = %d;" % (i, n))
        w("  ci->nval = %d;" % self.nval)
        nf = code.nodal_fields()
        w("  ci->n_fields = %d;" % len(nf))
        for i, f in enumerate(nf):
            w('  strncpy(ci->field_names[%d], "%s", 47); ci->field_space[%d] = %d; ci->field_index[%d] = %d;' % (
                i, f.name, i, 2 if f.space == "C2" else 1, i, f.index))
        w("  ci->moving_nodes = %d;" % (1 if code.coordinates_as_dofs else 0))
        w("  ci->ndof_el = %d;" % self.ndof)
        for k, (f, l) in enumerate(self.layout):
            fld = code.fields[f]
            node = l if fld.space != "C1" else self.et.c1_nodes[l]
            if fld.space == "Pos":
                w("  ci->dof_node[%d] = %d; ci->dof_kind[%d] = 0; ci->dof_index[%d] = %d;" % (k, node, k, k, fld.index))
            else:
                w("  ci->dof_node[%d] = %d; ci->dof_kind[%d] = 1; ci->dof_index[%d] = %d;" % (k, node, k, k, fld.index))
        rn = code.residual_names()
        w("  ci->n_residuals = %d;" % len(rn))
        for i, n in enumerate(rn):
            w('  strncpy(ci->residual_names[%d], "%s", 47);' % (i, n))
        w("  ci->n_params = %d;" % len(code.global_params))
        for i, n in enumerate(code.global_params):
            w('  strncpy(ci->param_names[%d], "%s", 47);' % (i, n))
        w("  ci->n_hist_val = %d; ci->n_hist_pos = %d; ci->max_dt_order = %d;" % (self.T_val, self.T_pos, code.max_dt_order()))
        w("  ci->elems_per_block = %d; ci->threads_per_block = %d; ci->smem_bytes = %d;" % (
            self._kernel_cfg[kernels[(self.routines[0].key, 1)]][0], self._kernel_cfg[kernels[(self.routines[0].key, 1)]][1], max(self._kernel_smem.values())))
        w("  ci->hessian_generated = %d;" % (1 if (self.hessian and self.pipeline) else 0))
        for what in (0, 1, 2):
            w("  ci->alg_bytes_per_elem[%d] = %r;" % (what, self.algorithmic_bytes(what)))
        w("  ci->alg_bytes_per_hist_level = %r;" % float(8 * sum(self._nnode_space(f.space) for f in code.nodal_fields())))
        inames = code.integral_expression_names()
        if len(inames) > 16:
            raise RuntimeError("more than PB2_MAX_INTEGRALS integral expressions")
        w("  ci->n_integrals = %d;" % len(inames))
        for i, n in enumerate(inames):
            w("  strncpy(ci->integral_names[%d], \"%s\", 47);" % (i, n))
        pnames = code.point_expression_names()
        if len(pnames) > 32:
            raise RuntimeError("more than PB2_MAX_POINT_EXPRS point expressions")
        w("  ci->n_point_exprs = %d;" % len(pnames))
        for i, (kind_, n) in enumerate(pnames):
            w("  strncpy(ci->point_names[%d], \"%s\", 47); ci->point_kind[%d] = %d;" % (i, n, i, {"local": 0, "extremum": 1, "z2": 2}[kind_]))
        w("  table->query = &pb2_query;")
        w("  table->launch = &pb2_launch;")
        w("}")
        self._emit_jit_element_init(o)
        return "\n".join(o) + "\n"

    def _emit_jit_element_init(self, o: List[str]):
        """The reference's own plugin entry point, ``JIT_ELEMENT_init(JITFuncSpec_Table_FiniteElement_t*)`` (src/jitbridge.h:499, emitted by
        FiniteElementCode::write_code_info, src/codegen.cpp:6398-7090), next to JIT_ELEMENT_init_cuda in the SAME shared object, so that the
        host's loader (DynamicBulkElementCode ctor, src/problem.cpp:101-142) accepts a CUDA plugin unchanged: the check_compiler_size
        handshake, every metadata field the host reads (dimensions, per-space field counts / names / nodal and buffer offsets, residual
        names, required shapes, global parameter indices, moving_nodes, max_dt_order, integration_order, dominant_space, integral
        expression names, domain name) and clean_up.  The per-element CPU routines of the table do not exist in a CUDA plugin (no CPU
        fallback): their slots hold a function that reports this and aborts; assembly goes through the batched launchers of the GPU block.
        Compiled only when the reference's jitbridge.h is on the include path (-DPB2_WITH_JITBRIDGE, CudaCCompiler.jitbridge_include):
        the table layout is the reference's own header, never a restatement."""
        code = self.code
        w = o.append
        nf = code.nodal_fields()
        c2 = [f for f in nf if f.space == "C2"]
        c1 = [f for f in nf if f.space == "C1"]
        rn = code.residual_names()
        dim = self.dim
        nres = max(1, len(rn))
        w("")
        w("#ifdef PB2_WITH_JITBRIDGE")
        w("// ---- the reference's plugin contract (src/jitbridge.h:329-499), compiled against the reference's own header")
        w("#define JIT_ELEMENT_SHARED_LIB")
        w("#include <stdio.h>")
        w('#include "jitbridge.h"')
        w("static JITFuncSpec_Table_FiniteElement_t* my_func_table;")
        w("static void pb2_no_cpu_rjm(const JITElementInfo_t*, const JITShapeInfo_t*, double*, double*, double*, unsigned)")
        w('{ fprintf(stderr, "pyoomph_b200: element class \'%s\' is a CUDA plugin: per-element CPU routines do not exist, assemble through JIT_ELEMENT_init_cuda / libpyoomph_b200\\n"); abort(); }' % self.name)
        w("static void pb2_no_cpu_hvp(const JITElementInfo_t*, const JITShapeInfo_t*, const double*, double*, double*, unsigned, unsigned)")
        w('{ fprintf(stderr, "pyoomph_b200: element class \'%s\' is a CUDA plugin: per-element CPU routines do not exist, assemble through JIT_ELEMENT_init_cuda / libpyoomph_b200\\n"); abort(); }' % self.name)
        w("static double pb2_no_cpu_integral(const JITElementInfo_t*, const JITShapeInfo_t*, unsigned) { pb2_no_cpu_rjm(0, 0, 0, 0, 0, 0u); return 0.0; }")
        w("static void pb2_free_names(char** tab, unsigned n) { if (!tab) return; for (unsigned i = 0; i < n; ++i) pyoomph_tested_free(tab[i]); free(tab); }")
        w("static void clean_up(JITFuncSpec_Table_FiniteElement_t* functable)")
        w("{")
        w("  pb2_free_names(functable->fieldnames_C2, functable->numfields_C2); functable->fieldnames_C2 = 0;")
        w("  pb2_free_names(functable->fieldnames_C1, functable->numfields_C1); functable->fieldnames_C1 = 0;")
        w("  pb2_free_names(functable->fieldnames_Pos, functable->numfields_Pos); functable->fieldnames_Pos = 0;")
        w("  pb2_free_names(functable->res_jac_names, functable->num_res_jacs); functable->res_jac_names = 0;")
        w("  pb2_free_names(functable->integral_expressions_names, functable->numintegral_expressions); functable->integral_expressions_names = 0;")
        w("  if (functable->ParameterDerivative) { for (unsigned i = 0; i < functable->num_res_jacs; ++i) pyoomph_tested_free(functable->ParameterDerivative[i]); free(functable->ParameterDerivative); functable->ParameterDerivative = 0; }")
        for nm in ("global_paramindices", "global_parameters", "ResidualAndJacobian", "ResidualAndJacobianSteady", "ResidualAndJacobian_NoHang",
                   "shapes_required_ResJac", "shapes_required_Hessian", "HessianVectorProduct", "missing_residual_assembly",
                   "has_constant_mass_matrix_for_sure", "temporal_error_scales", "discontinuous_refinement_exponents", "dominant_space", "domain_name"):
            w("  pyoomph_tested_free(functable->%s); functable->%s = 0;" % (nm, nm))
        w("}")
        w("")
        w('extern "C" JIT_API void JIT_ELEMENT_init(JITFuncSpec_Table_FiniteElement_t* functable)')
        w("{")
        w("  // the size handshake of src/codegen.cpp:6403-6421: the host compares with its own sizeof and refuses a mismatching compiler")
        for t, nm in (("char", "char"), ("unsigned short", "unsigned short"), ("unsigned int", "unsigned int"), ("unsigned long int", "unsigned long int"),
                      ("unsigned long long int", "unsigned long long int"), ("float", "float"), ("double", "double"), ("size_t", "size_t"),
                      ("struct JITElementInfo", "struct JITElementInfo"), ("struct JITHangInfoEntry", "struct JITHangInfoEntry"),
                      ("struct JITHangInfo", "struct JITHangInfo"), ("struct JITShapeInfo", "struct JITShapeInfo"),
                      ("struct JITFuncSpec_RequiredShapes_FiniteElement", "struct JITFuncSpec_RequiredShapes_FiniteElement"),
                      ("struct JITFuncSpec_Callback_Entry", "struct JITFuncSpec_Callback_Entry"),
                      ("struct JITFuncSpec_MultiRet_Entry", "struct JITFuncSpec_MultiRet_Entry"),
                      ("struct JITFuncSpec_Table_FiniteElement", "struct JITFuncSpec_Table_FiniteElement")):
            w('  if (functable->check_compiler_size) functable->check_compiler_size(sizeof(%s), sizeof(%s), (char*)"%s");' % (t, t, nm))
        w("  functable->nodal_dim = %d; functable->lagr_dim = %d;" % (dim, dim))
        w("  functable->fd_jacobian = false; functable->fd_position_jacobian = false; functable->with_adaptivity = false;")
        w("  functable->debug_jacobian_epsilon = 0.0; functable->stop_on_jacobian_difference = false;")
        # position space: coordinate_*, lagrangian_* (src/codegen.cpp:2816-2835); continuous spaces in the order C2TB|C2|C1TB|C1
        pos_names = ["coordinate_" + d for d in ex.DIRS[:dim]] + ["lagrangian_" + d for d in ex.DIRS[:dim]]
        w("  functable->numfields_Pos = %d;" % len(pos_names))
        w("  functable->fieldnames_Pos = (char**)malloc(sizeof(char*) * %d);" % len(pos_names))
        for i, n in enumerate(pos_names):
            w('  SET_INTERNAL_FIELD_NAME(functable->fieldnames_Pos, %d, "%s");' % (i, n))
        off = 0
        for sp_name, fl in (("C2", c2), ("C1", c1)):
            w("  functable->numfields_%s = functable->numfields_%s_bulk = functable->numfields_%s_basebulk = functable->numfields_%s_new = %d;" % (
                sp_name, sp_name, sp_name, sp_name, len(fl)))
            w("  functable->nodal_offset_%s_basebulk = %d; functable->buffer_offset_%s_basebulk = %d;" % (sp_name, off, sp_name, off))
            if fl:
                w("  functable->fieldnames_%s = (char**)malloc(sizeof(char*) * %d);" % (sp_name, len(fl)))
                for f in fl:
                    w('  SET_INTERNAL_FIELD_NAME(functable->fieldnames_%s, %d, "%s");' % (sp_name, f.index - off, f.name))
            off += len(fl)
        w("  functable->hangindex_C1 = functable->hangindex_C2 = functable->hangindex_C1TB = functable->hangindex_C2TB = functable->hangindex_Pos = -1;")
        w("  functable->num_res_jacs = %d; functable->current_res_jac = 0;" % len(rn))
        w("  functable->res_jac_names = (char**)calloc(%d, sizeof(char*));" % nres)
        for i, n in enumerate(rn):
            w('  SET_INTERNAL_FIELD_NAME(functable->res_jac_names, %d, "%s");' % (i, n))
        for nm, ty in (("ResidualAndJacobian", "JITFuncSpec_ResidualAndJacobian_FiniteElement"), ("ResidualAndJacobianSteady", "JITFuncSpec_ResidualAndJacobian_FiniteElement"),
                       ("ResidualAndJacobian_NoHang", "JITFuncSpec_ResidualAndJacobian_FiniteElement"), ("HessianVectorProduct", "JITFuncSpec_HessianVectorProduct_FiniteElement"),
                       ("shapes_required_ResJac", "JITFuncSpec_RequiredShapes_FiniteElement_t"), ("shapes_required_Hessian", "JITFuncSpec_RequiredShapes_FiniteElement_t")):
            w("  functable->%s = (%s*)calloc(%d, sizeof(%s));" % (nm, ty, nres, ty))
        w("  functable->missing_residual_assembly = (bool*)calloc(%d, sizeof(bool));" % nres)
        w("  functable->has_constant_mass_matrix_for_sure = (bool*)calloc(%d, sizeof(bool));" % nres)
        npar = len(code.global_params)
        w("  functable->numglobal_params = %d;" % npar)
        w("  functable->global_paramindices = (unsigned*)malloc(sizeof(unsigned) * %d);" % max(1, npar))
        w("  functable->global_parameters = (double**)calloc(%d, sizeof(double*));" % max(1, npar))
        for k in range(npar):
            w("  functable->global_paramindices[%d] = %d;   // '%s': the host resolves names to its parameter table (src/codegen.cpp:6700-6706)" % (k, k, code.global_params[k]))
        w("  functable->ParameterDerivative = (JITFuncSpec_ResidualAndJacobian_FiniteElement**)calloc(%d, sizeof(JITFuncSpec_ResidualAndJacobian_FiniteElement*));" % nres)
        for i, n in enumerate(rn):
            form = self.routines[[r.key for r in self.routines].index("r%d" % i)].form
            need_lagr = any(a.deriv.startswith("dX") or a.field.startswith("lagrangian_") for a in form.atoms) or form.uses_dX
            spaces_used = {code.fields[a.field].space for a in form.atoms} | {code.fields[s_.field].space for s_ in form.slots}
            for S in ("C2", "C1"):
                if S in spaces_used or (S == "C2" and "Pos" in spaces_used):
                    w("  functable->shapes_required_ResJac[%d].psi_%s = true; functable->shapes_required_ResJac[%d].dx_psi_%s = true;%s" % (
                        i, S, i, S, (" functable->shapes_required_ResJac[%d].dX_psi_%s = true;" % (i, S)) if need_lagr else ""))
            w("  functable->shapes_required_ResJac[%d].psi_Pos = true; functable->shapes_required_ResJac[%d].dx_psi_Pos = true;%s" % (
                i, i, (" functable->shapes_required_ResJac[%d].dX_psi_Pos = true;" % i) if need_lagr else ""))
            w("  functable->shapes_required_Hessian[%d] = functable->shapes_required_ResJac[%d];" % (i, i))
            w("  functable->ResidualAndJacobian[%d] = functable->ResidualAndJacobianSteady[%d] = functable->ResidualAndJacobian_NoHang[%d] = &pb2_no_cpu_rjm;" % (i, i, i))
            w("  functable->HessianVectorProduct[%d] = &pb2_no_cpu_hvp;" % i)
            w("  functable->ParameterDerivative[%d] = (JITFuncSpec_ResidualAndJacobian_FiniteElement*)calloc(%d, sizeof(JITFuncSpec_ResidualAndJacobian_FiniteElement));" % (i, max(1, npar)))
            for k in range(npar):
                w("  functable->ParameterDerivative[%d][%d] = &pb2_no_cpu_rjm;" % (i, k))
        w("  functable->hessian_generated = %s;" % ("true" if (self.hessian and self.pipeline) else "false"))
        w("  functable->use_shared_shape_buffer_during_multi_assemble = true;")
        w("  functable->temporal_error_scales = (double*)calloc(%d, sizeof(double));" % max(1, len(nf)))
        w("  functable->discontinuous_refinement_exponents = (double*)calloc(%d, sizeof(double));" % max(1, len(nf)))
        inames = code.integral_expression_names()
        w("  functable->numintegral_expressions = %d;" % len(inames))
        if inames:
            w("  functable->integral_expressions_names = (char**)malloc(sizeof(char*) * %d);" % len(inames))
            for i, n in enumerate(inames):
                w('  SET_INTERNAL_FIELD_NAME(functable->integral_expressions_names, %d, "%s");' % (i, n))
            w("  functable->EvalIntegralExpression = &pb2_no_cpu_integral;")
            w("  functable->shapes_required_IntegralExprs = functable->shapes_required_ResJac[0];")
        w("  functable->max_dt_order = %d;" % code.max_dt_order())
        w("  functable->moving_nodes = %s;" % ("true" if code.coordinates_as_dofs else "false"))
        w("  functable->integration_order = 0;     // the element's default scheme: Gauss<DIM,3> (src/elements.cpp:184-249)")
        w('  SET_INTERNAL_NAME(functable->dominant_space, "C2");')
        w('  SET_INTERNAL_NAME(functable->domain_name, "%s");' % self.name)
        w("  functable->clean_up = &clean_up;")
        w("  my_func_table = functable;")
        w("}")
        w("#endif  // PB2_WITH_JITBRIDGE")

    def algorithmic_bytes(self, what: int) -> float:
        """B_el of SURVEY 8(d): gather (positions, nodal values x history, local->global map) + scatter
        (residual add, Jacobian values written once, element->CSR position map read)."""
        code = self.code
        c_pos = 2 if any(a.field.startswith("lagrangian") or a.deriv.startswith("dX") for rp in self.routines for a in rp.form.atoms) else 1
        gather = 8 * self.NN * self.dim * c_pos
        gather += 8 * self.T_val * sum(self._nnode_space(f.space) for f in code.nodal_fields())
        gather += 4 * self.ndof
        scatter = 8 * self.ndof
        if what >= 1:
            scatter += 12 * self.ndof * self.ndof
        if what >= 2:
            scatter += 12 * self.ndof * self.ndof
        return float(gather + scatter)


def emit_cuda_source(code: FiniteElementCode, name: str = "elem", **kw) -> str:
    return CudaEmitter(code, name, **kw).emit()
