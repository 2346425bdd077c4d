"""Element-class code model: fields, spaces, residual bookkeeping and symbolic derivation.

Host-side mirror of ``pyoomph::FiniteElementCode`` (/root/reference/src/codegen.cpp) and of the
Python ``Equations`` front end (/root/reference/pyoomph/generic/codegen.py:1852).  The reference
derives, per test field and per unknown field, a *full* expression ``GiNaC::diff(var_part, field)``
(src/codegen.cpp:1956) that the generated C evaluates for every (l_test, l_shape) pair.  The B200
design instead exploits that a weak form is linear in the test function and its gradient, and
that every shape expansion enters through ``psi_l`` or ``d psi_l/dx``:

    E  = sum_s  T_s[l_test] * R_s(point)                 s = (test field, test atom)
    J  = sum_s sum_a T_s[l_test] * C_{s,(G,a)}(point) * S_a[l_shape]     a = shape atom of unknown G

so the only problem-specific code is the *pointwise* evaluation of R_s and C_{s,(G,a)} (straight-line
code after CSE); the (l_test,l_shape) contraction is a fixed register-tiled kernel.  Moving-mesh
columns (src/codegen.cpp:8454-8586, src/elements.cpp:3051) fit the same form through
``d(d psi_m/dx_d)/dX^l_j = -(d psi_m/dx_j)(d psi_l/dx_d)`` and ``d(dx)/dX^l_j = dx * d psi_l/dx_j``
(valid for elements without co-dimension, which is all bulk elements).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import sympy as sp
from sympy.core.function import AppliedUndef

from . import expressions as ex

# ---------------------------------------------------------------------------------------------
# Element types (src/elements.hpp:556-1377).  Node order = oomph QElement tensor-product order,
# first local coordinate fastest (oomph-lib Qelements.cc:348-377); C1 nodes = vertices
# (src/elements.cpp:8733, :10764).
# ---------------------------------------------------------------------------------------------


@dataclasses.dataclass(frozen=True)
class ElementType:
    name: str
    nodal_dim: int
    elem_dim: int
    nnode: int               # nodes of the dominant (geometry) space
    order: int               # 1D nodes per direction of the geometry space (3 = quadratic)
    c1_nodes: Tuple[int, ...]  # element-local node numbers carrying the C1 space
    n_int_pt: int            # default oomph integration scheme (integration_order==0)

    @property
    def nnode_C1(self):
        return len(self.c1_nodes)


ELEMENT_TYPES: Dict[str, ElementType] = {
    "Quad2dC2": ElementType("Quad2dC2", 2, 2, 9, 3, (0, 2, 6, 8), 9),
    "Brick3dC2": ElementType("Brick3dC2", 3, 3, 27, 3, (0, 2, 6, 8, 18, 20, 24, 26), 27),
    # BulkElementTri2dC2 = oomph TElement<2,3> (src/elements.hpp:990, src/elements.cpp:9844-9856): 6 nodes (vertices 0,1,2 then the
    # mid-side nodes 3: 0-1, 4: 1-2, 5: 2-0, Telements.h:575-621), C1 on the vertices, default scheme TGauss<2,3> with 7 points
    "Tri2dC2": ElementType("Tri2dC2", 2, 2, 6, 3, (0, 1, 2), 7),
    # InterfaceElementLine1dC2 (src/elements.hpp:1435-2298): the three nodes of a Q9 / T6 edge in a 2D space; QElement<1,3> shapes,
    # C1 on the end nodes, Gauss<1,3>
    "Line1dC2": ElementType("Line1dC2", 2, 1, 3, 3, (0, 2), 3),
    # BulkElementTetra3dC2 = TElement<3,3>: ten-node tetrahedra (vertices 0-3, then the mid-edge nodes in oomph order, Telements.h:2051),
    # C1 on the four vertices, TGauss<3,3> (11 points)
    "Tetra3dC2": ElementType("Tetra3dC2", 3, 3, 10, 3, (0, 1, 2, 3), 11),
    # A FACE of a Q9 bulk element seen through the bulk element (the reference's bulk_eleminfo / opposite_eleminfo of an interface element,
    # src/jitbridge.h:88-120: `blk_` / `oppblk_` quantities of the generated code): the nine nodes of the bulk element, rotated so that the
    # face is s1 = -1; the integral runs over the face (Gauss<1,3> in s0, measure |dx/ds0|, outer normal (t_y, -t_x)/|t|) while shape
    # functions and their Eulerian gradients are the BULK ones evaluated on the face -- normal derivatives of bulk fields are available.
    "QuadFace2dC2": ElementType("QuadFace2dC2", 2, 2, 9, 3, (0, 2, 6, 8), 3),
}

SPACE_ORDER = ("C2TB", "C2", "C1TB", "C1")  # nodal_data index order (src/codegen.cpp:2367-2380)


@dataclasses.dataclass
class Field:
    name: str
    space: str          # "C2", "C1" or "Pos"
    index: int = -1     # index into nodal_data / nodal_coords (src/codegen.cpp:2795-2835)
    aux_of: str = ""    # "Y__<field>": values of a Hessian direction vector on the dofs of <field> (not a nodal value)


@dataclasses.dataclass(frozen=True)
class AtomInfo:
    """One interpolated quantity at an integration point (a printed ShapeExpansion)."""
    field: str
    dt_order: int        # 0,1,2
    scheme: str          # "", "BDF1", "BDF2", "Newmark2" (+"_degr")
    deriv: str           # "d0", "dx0".."dx2", "dX0".."dX2"
    past: int = 0        # history index (evaluate_in_past)

    @property
    def cname(self) -> str:
        """Name in the reference's generated code (src/codegen.cpp:943-1032)."""
        d = {"d0": "d0x"}.get(self.deriv, "d1" + self.deriv[1] + self.deriv[2:])
        return "intrp_d%dt%d%s_%s_%s" % (self.dt_order, self.past, self.scheme if self.dt_order else "", d, self.field)


@dataclasses.dataclass(frozen=True)
class TestSlot:
    field: str
    deriv: str           # "d0", "dxN", "dXN"


@dataclasses.dataclass
class ResidualForm:
    """Coefficient form of one generated routine (ResidualAndJacobian<i>, dResidual<i>dParameter_<p>)."""
    name: str
    slots: List[TestSlot]
    R: List[sp.Expr]                                          # per slot, measure weights folded in
    J: Dict[Tuple[int, str, str], sp.Expr]                    # (slot, unknown field, shape atom) -> coefficient
    M: Dict[Tuple[int, str, str], sp.Expr]                    # mass-matrix coefficients (flag==2)
    atoms: List[AtomInfo]                                     # point inputs needed
    uses_dx: bool = True
    uses_dX: bool = False

    def unknown_fields(self) -> List[str]:
        seen: List[str] = []
        for (_, g, _a) in list(self.J.keys()) + list(self.M.keys()):
            if g not in seen:
                seen.append(g)
        return seen


class FiniteElementCode:
    """Per-element-class code object (FiniteElementCode, src/codegen.hpp; FiniteElementCodeGenerator,
    pyoomph/generic/codegen.py:60)."""

    def __init__(self, element_type: str, equations: "Equations", *, default_timestepping_scheme: str = "BDF2",
                 name: str = "domain", coordinate_system=None):
        self.etype = ELEMENT_TYPES[element_type]
        self.nodal_dim = self.etype.nodal_dim
        # Problem.set_coordinate_system (pyoomph/generic/problem.py) / Equations.get_coordinate_system: Cartesian unless told otherwise
        if isinstance(coordinate_system, str):
            coordinate_system = {"cartesian": ex.cartesian, "axisymmetric": ex.axisymmetric}[coordinate_system]
        self.coordinate_system = coordinate_system or ex.cartesian
        self.name = name
        self.default_timestepping_scheme = default_timestepping_scheme
        self.fields: Dict[str, Field] = {}
        self.vector_fields: Dict[str, List[str]] = {}
        self.coordinates_as_dofs = False
        self.global_params: List[str] = []
        self._param_syms: Dict[str, sp.Symbol] = {}
        self.residuals: Dict[str, sp.Expr] = {}
        self.integral_expressions: Dict[str, sp.Expr] = {}     # FiniteElementCode::integral_expressions (src/codegen.hpp:677)
        # expressions evaluated at a local coordinate of an element (no measure): local expressions (output at the nodes), extremum
        # expressions (sampled at Gauss points and nodes), Z2 flux terms of the error estimator (src/codegen.cpp:4125-4453)
        self.local_expressions: Dict[str, sp.Expr] = {}
        self.extremum_expressions: Dict[str, sp.Expr] = {}
        self.Z2_fluxes: List[sp.Expr] = []
        self._atom_syms: Dict[sp.Symbol, AtomInfo] = {}
        self._atom_by_info: Dict[AtomInfo, sp.Symbol] = {}
        self._test_syms: Dict[sp.Symbol, TestSlot] = {}
        self.equations = equations
        self._defining_fields = False
        # position fields always exist (src/codegen.cpp:2816-2835)
        for i, d in enumerate(ex.DIRS[:self.nodal_dim]):
            self.fields["coordinate_" + d] = Field("coordinate_" + d, "Pos", i)
        for i, d in enumerate(ex.DIRS[:self.nodal_dim]):
            self.fields["lagrangian_" + d] = Field("lagrangian_" + d, "Pos", self.nodal_dim + i)
        ex._Context.stack.append(self)
        try:
            equations._code = self
            self._defining_fields = True
            equations.define_fields()
            self._defining_fields = False
            self._index_fields()
            equations.define_residuals()
            equations.define_additional_functions()
        finally:
            ex._Context.stack.pop()
        self._forms: Dict[str, ResidualForm] = {}

    # -- field bookkeeping ---------------------------------------------------------------------
    def define_scalar_field(self, name: str, space: str):
        if space not in ("C2", "C1"):
            raise RuntimeError("space %s is outside the scope of the CUDA assembly path (C2/C1 only)" % space)
        if space == "C2" and self.etype.order < 3:
            raise RuntimeError("C2 field on a first-order element")
        self.fields[name] = Field(name, space)

    def define_vector_field(self, name: str, space: str, dim: Optional[int] = None):
        dim = dim or self.nodal_dim
        if dim > self.nodal_dim and self.coordinate_system.get_id_name() == "Axisymmetric":
            comps = [name + "_x", name + "_y", name + "_phi"][:dim]       # azimuthal (swirl) component
        else:
            comps = [name + "_" + d for d in ex.DIRS[:dim]]
        for c in comps:
            self.define_scalar_field(c, space)
        self.vector_fields[name] = comps

    def _index_fields(self):
        idx = 0
        for sp_name in SPACE_ORDER:
            for f in self.fields.values():
                if f.space == sp_name and not f.aux_of:
                    f.index = idx
                    idx += 1
        self.n_nodal_values = idx

    def nodal_fields(self) -> List[Field]:
        return sorted([f for f in self.fields.values() if f.space != "Pos" and not f.aux_of], key=lambda f: f.index)

    def _require_field(self, name: str):
        if name.startswith("coordinate_") or name.startswith("lagrangian_"):
            if name not in self.fields:
                raise RuntimeError("no such position field " + name)
            return
        if name not in self.fields:
            raise RuntimeError("field '%s' is not defined on element class '%s'" % (name, self.name))

    def _global_param_symbol(self, name: str) -> sp.Symbol:
        if name not in self._param_syms:
            self._param_syms[name] = sp.Symbol("P__" + name, real=True)
            self.global_params.append(name)
        return self._param_syms[name]

    def add_residual(self, expr, destination: str = ""):
        expr = sp.sympify(expr)
        cs = self.coordinate_system
        if getattr(cs, "has_normal_mode_expansion", False):
            # Problem._residual_mapping_functions of define_problem_for_axial_symmetry_breaking_investigation
            # (pyoomph/generic/problem.py:4543-4558): one residual -> base mode + real / imaginary part of the angular eigenproblem
            parts = {destination: cs.map_residual_on_base_mode(expr),
                     cs.real_contribution_name + destination: cs.map_residual_on_angular_eigenproblem_real(expr),
                     cs.imag_contribution_name + destination: cs.map_residual_on_angular_eigenproblem_imag(expr)}
        else:
            parts = {destination: expr}
        for dest, e in parts.items():
            if e == 0 and dest != destination:
                continue                     # e.g. the imaginary part of a scalar diffusion operator: no empty routines
            self.residuals[dest] = self.residuals.get(dest, sp.Integer(0)) + e

    def residual_names(self) -> List[str]:
        return list(self.residuals.keys())

    # -- integral expressions (FiniteElementCode::_register_integral_function, src/codegen.cpp:6186-6212) -------------
    def add_integral_function(self, name: str, expr):
        """integrand INCLUDING its measure (the caller multiplies by dx, pyoomph/equations/generic.py:699); a vector integrand
        registers one expression per component, name_x, name_y, ... (src/codegen.cpp:6199-6208)"""
        if isinstance(expr, sp.MatrixBase):
            for i in range(expr.shape[0]):
                if expr[i, 0] != 0 or i < self.nodal_dim:
                    self.integral_expressions[name + "_" + (ex.DIRS[i] if i < self.nodal_dim else "phi")] = sp.sympify(expr[i, 0])
            return
        self.integral_expressions[name] = sp.sympify(expr)

    def integral_expression_names(self) -> List[str]:
        return list(self.integral_expressions.keys())

    INTEGRAL_GRADIENT_PREFIX = "d_integral_"

    def add_integral_gradient(self, name: str):
        """Register the residual contribution  "d_integral_<name>":  c_j = d/dU_j (integral expression <name>)  -- the Gateaux derivative
        of the integrand in the direction of the test functions.  Its residual vector (flag 0) is the dense ROW a global constraint
        g(U) = integral - target = 0 adds to the system, its Jacobian (flag 1) the constraint's second derivative, and the dense COLUMN of
        the multiplier is a parameter derivative dR/d(lambda) -- the bordered form of pyoomph's GlobalLagrangeMultiplier
        (pyoomph/generic/codegen.py:2927; SURVEY 8e: such rows are not sharded, their vectors are reduced over the ranks).  Fixed meshes,
        integrands without time derivatives."""
        if self.coordinates_as_dofs:
            raise NotImplementedError("gradient of an integral expression on a moving mesh")
        I = self.integral_expressions[name]
        eps = sp.Symbol("EPS__gateaux", real=True)
        repl = {}
        for f in self.nodal_fields():
            F = sp.Function("F__" + f.name, real=True)(*ex._args(self.nodal_dim))
            T = sp.Function("T__" + f.name, real=True)(*ex._args(self.nodal_dim))
            repl[F] = F + eps * T
        if any(isinstance(d, sp.Derivative) and ex.TIME in [v for v, _ in d.variable_count] for d in I.atoms(sp.Derivative)):
            raise NotImplementedError("gradient of an integral expression with time derivatives")
        g = sp.diff(I.subs(repl).doit(), eps).subs(eps, 0).doit()
        self.residuals[self.INTEGRAL_GRADIENT_PREFIX + name] = sp.expand(g)

    # -- local / extremum expressions and Z2 fluxes (src/codegen.cpp:4366-4453; pyoomph/generic/codegen.py:1213, :2126) -----------------
    def _register_components(self, dest: Dict[str, sp.Expr], name: str, expr):
        if isinstance(expr, sp.MatrixBase):
            if expr.shape[1] == 1:
                for i in range(expr.shape[0]):
                    if expr[i, 0] != 0 or i < self.nodal_dim:
                        dest[name + "_" + (ex.DIRS[i] if i < self.nodal_dim else "phi")] = sp.sympify(expr[i, 0])
            else:
                for i in range(self.nodal_dim):
                    for j in range(self.nodal_dim):
                        dest["%s_%s%s" % (name, ex.DIRS[i], ex.DIRS[j])] = sp.sympify(expr[i, j])
            return
        dest[name] = sp.sympify(expr)

    def add_local_function(self, name: str, expr):
        """node-wise output quantity (Equations.add_local_function, pyoomph/generic/codegen.py:1213): vectors / tensors register one
        expression per component"""
        self._register_components(self.local_expressions, name, expr() if callable(expr) else expr)

    def add_extremum_function(self, name: str, expr):
        self._register_components(self.extremum_expressions, name, expr() if callable(expr) else expr)

    def add_Z2_flux(self, expr):
        """flux terms of the Z2 error estimator (Equations.add_spatial_error_estimator -> _add_Z2_flux): every component of a
        scalar / vector / tensor expression is one flux term"""
        expr = expr() if callable(expr) else expr
        if isinstance(expr, sp.MatrixBase):
            self.Z2_fluxes += [sp.sympify(v) for v in expr]
        else:
            self.Z2_fluxes.append(sp.sympify(expr))

    def point_expression_names(self) -> List[Tuple[str, str]]:
        """(kind, name) of every expression of the point-evaluation routine, in its output order: local, extremum, Z2"""
        return [("local", n) for n in self.local_expressions] + [("extremum", n) for n in self.extremum_expressions] + \
               [("z2", "flux_%d" % i) for i in range(len(self.Z2_fluxes))]

    def point_form(self) -> ResidualForm:
        """EvalLocalExpression, EvalExtremumExpression and GetZ2Fluxes as ONE coefficient form without test functions and without
        measure: slot i carries expression i of point_expression_names()"""
        if "|points" in self._forms:
            return self._forms["|points"]
        exprs = list(self.local_expressions.values()) + list(self.extremum_expressions.values()) + list(self.Z2_fluxes)
        slots, R = [], []
        for i, e in enumerate(exprs):
            E = self.atomize(e)
            if any(s_ in self._test_syms for s_ in E.free_symbols):
                raise RuntimeError("Found test function in a custom integral/local expression")     # src/codegen.cpp:4145
            slots.append(TestSlot("__point_%d" % i, "d0"))
            R.append(E)
        used = set()
        for e in R:
            used |= {s_ for s_ in e.free_symbols if s_ in self._atom_syms}
        atoms = sorted((self._atom_syms[s_] for s_ in used), key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        allsyms = set().union(*[e.free_symbols for e in R]) if R else set()
        form = ResidualForm("|points", slots, R, {}, {}, atoms, uses_dx=ex.DX_EUL in allsyms, uses_dX=(ex.DX_LAG in allsyms or ex.ELEMSIZE_LAG in allsyms or ex.ELEMSIZE_LAG_CART in allsyms))
        self._forms["|points"] = form
        return form

    def integral_form(self) -> ResidualForm:
        """The integral expressions as a coefficient form without test functions: slot i carries integrand i (its "R"), so the
        emitters reuse their gather / geometry / interpolation code (write_code_integral_or_local_expressions,
        src/codegen.cpp:4125-4364 does the same with the residual machinery)."""
        if "|integrals" in self._forms:
            return self._forms["|integrals"]
        slots, R = [], []
        for i, (n, e) in enumerate(self.integral_expressions.items()):
            E = self.atomize(e)
            if any(s in self._test_syms for s in E.free_symbols):
                raise RuntimeError("Found test function in a custom integral/local expression")     # src/codegen.cpp:4145
            slots.append(TestSlot("__integral_%d" % i, "d0"))
            R.append(E)
        used = set()
        for e in R:
            used |= {s for s in e.free_symbols if s in self._atom_syms}
        atoms = sorted((self._atom_syms[s] for s in used), key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        allsyms = set().union(*[e.free_symbols for e in R]) if R else set()
        form = ResidualForm("|integrals", slots, R, {}, {}, atoms, uses_dx=ex.DX_EUL in allsyms, uses_dX=(ex.DX_LAG in allsyms or ex.ELEMSIZE_LAG in allsyms or ex.ELEMSIZE_LAG_CART in allsyms))
        self._forms["|integrals"] = form
        return form

    def _all_forms(self) -> List[ResidualForm]:
        return [self.derive(n) for n in self.residual_names()] + ([self.integral_form()] if self.integral_expressions else []) + \
               ([self.point_form()] if self.point_expression_names() else [])

    def space_nodes(self, space: str) -> Tuple[int, ...]:
        """Element-local node numbers of a space (src/elements.cpp:2870-2879)."""
        if space in ("C2", "Pos"):
            return tuple(range(self.etype.nnode))
        if space == "C1":
            return self.etype.c1_nodes
        raise KeyError(space)

    def dof_layout(self) -> List[Tuple[str, int]]:
        """Local dof order used by this engine: for each element node in local order, position
        dofs (if coordinates are dofs) then nodal values by index.  oomph's own local order is
        nodal values then solid positions (oomph-lib elements.cc:694-699); the generated code never
        depends on it, it only maps through *_local_eqn."""
        out: List[Tuple[str, int]] = []
        for l in range(self.etype.nnode):
            if self.coordinates_as_dofs:
                for d in ex.DIRS[:self.nodal_dim]:
                    out.append(("coordinate_" + d, l))
            for f in self.nodal_fields():
                nodes = self.space_nodes(f.space)
                if l in nodes:
                    out.append((f.name, nodes.index(l)))
        return out

    # -- atomisation ---------------------------------------------------------------------------
    def _time_scheme(self, dt_order: int) -> str:
        # get_default_timestepping_scheme (pyoomph/generic/problem.py:750-757); dt^2 -> Newmark2
        # (src/codegen.cpp:8245); first order default gets the _degr suffix (src/codegen.cpp:8116)
        if dt_order == 2:
            return "Newmark2"
        s = self.default_timestepping_scheme
        return s + "_degr" if s != "BDF1" else s

    def _atom(self, info: AtomInfo) -> sp.Symbol:
        if info not in self._atom_by_info:
            s = sp.Symbol("A__%s__dt%d%s__%s__p%d" % (info.field, info.dt_order, info.scheme, info.deriv, info.past), real=True)
            self._atom_by_info[info] = s
            self._atom_syms[s] = info
        return self._atom_by_info[info]

    def _test_atom(self, slot: TestSlot) -> sp.Symbol:
        s = sp.Symbol("Tst__%s__%s" % (slot.field, slot.deriv), real=True)
        self._test_syms[s] = slot
        return s

    def _classify(self, d):
        """Split a Derivative/AppliedUndef node into (kind, field, past, dt_order, eul_dirs, lag_dirs)."""
        if isinstance(d, sp.Derivative):
            base = d.expr
            counts = {v: n for v, n in d.variable_count}
        else:
            base, counts = d, {}
        nm = base.func.__name__
        past = 0
        if "__past" in nm:
            nm, p = nm.split("__past")
            past = int(p)
        kind, field = nm[0], nm[3:]
        dt = int(counts.get(ex.TIME, 0))
        eul = [i for i, c in enumerate(ex.EUL) for _ in range(int(counts.get(c, 0)))]
        lag = [i for i, c in enumerate(ex.LAG) for _ in range(int(counts.get(c, 0)))]
        return kind, field, past, dt, eul, lag

    def atomize(self, expr: sp.Expr) -> sp.Expr:
        expr = sp.sympify(expr)
        repl = {}
        nodes = list(expr.atoms(sp.Derivative)) + list(expr.atoms(AppliedUndef))
        for d in nodes:
            base = d.expr if isinstance(d, sp.Derivative) else d
            if not isinstance(base, AppliedUndef) or not base.func.__name__[:3] in ("F__", "T__"):
                raise RuntimeError("cannot atomise " + str(d))
            kind, field, past, dt, eul, lag = self._classify(d)
            if len(eul) + len(lag) > 1:
                raise RuntimeError("second spatial derivatives of C0 shape expansions are not available: " + str(d))
            deriv = "d0" if not (eul or lag) else ("dx%d" % eul[0] if eul else "dX%d" % lag[0])
            if kind == "T":
                if dt or past:
                    raise RuntimeError("time derivative of a test function")
                repl[d] = self._test_atom(TestSlot(field, deriv))
                continue
            # position-space rules (src/codegen.cpp:8220-8230, :8285-8324, :8376-8399)
            if field.startswith("coordinate_"):
                i = ex.DIRS.index(field[-1])
                if eul:
                    if dt:
                        raise RuntimeError("spatial derivative of the mesh velocity is not supported")
                    repl[d] = sp.Integer(1 if eul[0] == i else 0)
                    continue
            if field.startswith("lagrangian_"):
                i = ex.DIRS.index(field[-1])
                if dt:
                    repl[d] = sp.Integer(0)
                    continue
                if lag:
                    repl[d] = sp.Integer(1 if lag[0] == i else 0)
                    continue
                if eul:
                    raise RuntimeError("Eulerian derivative of Lagrangian coordinates is not supported")
            if dt > 2:
                raise RuntimeError("Too high dt order")
            scheme = self._time_scheme(dt) if dt else ""
            if field.endswith(ex.MODE_SUFFIX) and field not in self.fields and field[:-len(ex.MODE_SUFFIX)] not in self.fields:
                raise RuntimeError("mode copy of an unknown field: " + field)
            repl[d] = self._atom(AtomInfo(field, dt, scheme, deriv, past))
        return expr.xreplace(repl)

    def _mode_base(self, field: str) -> Optional[str]:
        """base field of a perturbation-mode copy that is NOT a field of its own (the azimuthal eigenvector lives on the dofs of the
        base fields); None for ordinary fields"""
        if field.endswith(ex.MODE_SUFFIX) and field not in self.fields:
            return field[:-len(ex.MODE_SUFFIX)]
        return None

    # -- derivation ----------------------------------------------------------------------------
    def unknown_field_names(self) -> List[str]:
        out = []
        if self.coordinates_as_dofs:
            out += ["coordinate_" + d for d in ex.DIRS[:self.nodal_dim]]
        out += [f.name for f in self.nodal_fields()]
        return out

    def derive(self, resname: str = "", parameter: Optional[str] = None) -> ResidualForm:
        key = resname + ("|dP_" + parameter if parameter else "")
        if key in self._forms:
            return self._forms[key]
        E = self.atomize(self.residuals[resname])
        if parameter is not None:
            E = sp.diff(E, self._param_syms[parameter])
        form = self._coefficient_form(E, key)
        self._forms[key] = form
        return form

    def _coefficient_form(self, E: sp.Expr, name: str) -> ResidualForm:
        if self.coordinates_as_dofs and (E.has(ex.ELEMSIZE_EUL) or E.has(ex.ELEMSIZE_EUL_CART)):
            raise NotImplementedError("element sizes on a moving mesh (elemsize_d_coords, src/jitbridge.h:180)")
        tests = sorted([s for s in E.free_symbols if s in self._test_syms], key=lambda s: s.name)
        slots: List[TestSlot] = []
        R: List[sp.Expr] = []
        rest = E
        for ts in tests:
            coeff = sp.diff(E, ts)
            if coeff.has(*tests):
                raise RuntimeError("residual is not linear in the test functions")
            slots.append(self._test_syms[ts])
            R.append(coeff)
            rest = rest - ts * coeff
        if sp.simplify(sp.expand(rest)) != 0:
            raise RuntimeError("residual has a part without test function: " + str(rest))
        unknowns = set(self.unknown_field_names())
        J: Dict[Tuple[int, str, str], sp.Expr] = {}
        M: Dict[Tuple[int, str, str], sp.Expr] = {}

        def add(dic, k, v):
            if v != 0:
                dic[k] = dic.get(k, sp.Integer(0)) + v

        def slot_index(slot: TestSlot) -> int:
            if slot not in slots:
                slots.append(slot)
                R.append(sp.Integer(0))
            return slots.index(slot)

        # Angular eigenproblem contributions (azimuthal stability): the residual is linear in the mode copies U__M1 of the fields; its
        # Jacobian and mass matrix are taken with respect to THEM (expansion_mode tags, src/codegen.cpp:7567, :8196) and land in the
        # columns of the base fields' dofs; afterwards the copies are evaluated at the base state.
        mode_atoms = {s: self._atom_syms[s] for s in E.free_symbols if s in self._atom_syms and self._mode_base(self._atom_syms[s].field)}
        nslot0 = len(slots)
        for si in range(nslot0):
            Rs = R[si]
            atoms = [s for s in Rs.free_symbols if s in self._atom_syms]
            for a in atoms:
                info = self._atom_syms[a]
                col_field = info.field
                if mode_atoms:
                    if a not in mode_atoms:
                        continue
                    col_field = self._mode_base(info.field)
                if info.past or col_field not in unknowns:
                    continue  # history values and non-dof data carry no Jacobian (src/codegen.cpp:8256)
                c = sp.diff(Rs, a)
                if info.dt_order == 0:
                    add(J, (si, col_field, info.deriv), c)
                else:
                    w = sp.Symbol("W__%s__%d" % (info.scheme, info.dt_order), real=True)
                    add(J, (si, col_field, info.deriv), w * c)
                    if info.dt_order == 1:
                        add(M, (si, col_field, info.deriv), c)  # __partial_t_mass_matrix (src/codegen.cpp:8260)
            if self.coordinates_as_dofs and self.etype.name.startswith("QuadFace"):
                raise NotImplementedError("bulk-face interface elements on a moving mesh (position derivatives of the face measure and normal)")
            if self.coordinates_as_dofs:
                RE = ex.DX_EUL * sp.diff(Rs, ex.DX_EUL)  # Eulerian-measure part
                slot = slots[si]
                for j, dj in enumerate(ex.DIRS[:self.nodal_dim]):
                    Xj = "coordinate_" + dj
                    # d(dx)/dX_j^l = dx * dpsi_l/dx_j   (int_pt_weights_d_coords, src/elements.cpp:3086)
                    add(J, (si, Xj, "dx%d" % j), RE)
                    # d(dpsi_m/dx_d)/dX_j^l = -dpsi_m/dx_j * dpsi_l/dx_d applied to every Eulerian gradient atom
                    for a in atoms:
                        info = self._atom_syms[a]
                        if info.deriv.startswith("dx") and not info.past:
                            aj = self._atom(dataclasses.replace(info, deriv="dx%d" % j))
                            add(J, (si, Xj, info.deriv), -sp.diff(Rs, a) * aj)
                    # ... and to the test-function gradient itself (src/codegen.cpp:8642)
                    if slot.deriv.startswith("dx"):
                        sj = slot_index(TestSlot(slot.field, "dx%d" % j))
                        add(J, (sj, Xj, slot.deriv), -Rs)
                    if self.etype.elem_dim < self.nodal_dim:
                        # Interface elements (a line in 2D): dpsi_m/dx_d = t_d psi'_m / g is the SURFACE gradient (g = t.t), so
                        #   d(dpsi_m/dx_d)/dX_j^l = -dpsi_m/dx_d dpsi_l/dx_j  +  n_d n_j (grad_S psi_m . grad_S psi_l)
                        # (the first part is the bulk identity: tau_d tau_j is symmetric), d(dx)/dX_j^l = dx dpsi_l/dx_j as in the bulk,
                        # and the unit normal n = (-t_y, t_x)/|t| moves with the nodes:
                        #   d n_i / dX_j^l = -tau_i n_j psi'_l/|t| = -tau_i n_j (tau . grad_S psi_l),   tau = (n_y, -n_x)
                        # (get_dnormal_dcoords_at_s, src/elements.cpp:1461-1490; the reference gets the gradient part from its general
                        # el_dim x nodal_dim tensors, src/elements.cpp:3051-3155).
                        if self.nodal_dim != 2:
                            raise NotImplementedError("moving interface elements: lines in 2D only")
                        nrm = ex.NORMAL[:2]
                        tau = (nrm[1], -nrm[0])
                        for a in atoms:
                            info = self._atom_syms[a]
                            if info.deriv.startswith("dx") and not info.past:
                                d = int(info.deriv[2:])
                                for b in range(2):
                                    ab = self._atom(dataclasses.replace(info, deriv="dx%d" % b))
                                    add(J, (si, Xj, "dx%d" % b), sp.diff(Rs, a) * nrm[d] * nrm[j] * ab)
                        if slot.deriv.startswith("dx"):
                            d = int(slot.deriv[2:])
                            for b in range(2):
                                sb = slot_index(TestSlot(slot.field, "dx%d" % b))
                                add(J, (sb, Xj, "dx%d" % b), Rs * nrm[d] * nrm[j])
                        for i in range(2):
                            dRn = sp.diff(Rs, nrm[i])
                            if dRn != 0:
                                for b in range(2):
                                    add(J, (si, Xj, "dx%d" % b), -dRn * tau[i] * nrm[j] * tau[b])
        if mode_atoms:
            if self.coordinates_as_dofs:
                raise NotImplementedError("mode expansion on a moving mesh")
            to_base = {s: self._atom(dataclasses.replace(i, field=self._mode_base(i.field))) for s, i in mode_atoms.items()}
            R = [e.xreplace(to_base) for e in R]
            J = {k: v.xreplace(to_base) for k, v in J.items()}
            M = {k: v.xreplace(to_base) for k, v in M.items()}
        used = set()
        for e in list(R) + list(J.values()) + list(M.values()):
            used |= {s for s in e.free_symbols if s in self._atom_syms}
        atoms = sorted((self._atom_syms[s] for s in used), key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        allsyms = set().union(*[e.free_symbols for e in list(R) + list(J.values()) + list(M.values())]) if R else set()
        return ResidualForm(name, slots, R, J, M, atoms, uses_dx=ex.DX_EUL in allsyms, uses_dX=(ex.DX_LAG in allsyms or ex.ELEMSIZE_LAG in allsyms or ex.ELEMSIZE_LAG_CART in allsyms))

    # -- Hessian (second derivatives), fixed meshes ------------------------------------------------
    def derive_hessian(self, resname: str = ""):
        """Coefficient form of HessianVectorProduct<i> (src/codegen.cpp:3646-3910) without the ndof^3 buffer.

        The reference fills H[i][j][k] = d/dU_k (dR_i/dU_j) (and the mass Hessian d/dU_k (dR_i/d(partial_t U_j)),
        src/codegen.cpp:1680-1843) and contracts the MIDDLE index with the vector (SET_DIRECTIONAL_SYMMETRIC_HESSIAN_FROM,
        src/jitbridge.h:663): N_ik = sum_j H_ijk Y_j = d((A.Y)_i)/dU_k for A = J or M.  With A_ij = T_b C_{s,(G,a)} S_a[l_j] this is

            N[(F,l_t),(H,l_k)] = T_b[l_t] * ( sum_{(G,a)} Yhat_{G,a} * dC_{s,(G,a)}/d atom(H,c) * fac ) * S_c[l_k]

        where Yhat_{G,a} is Y interpolated like field G (value or gradient).  Returns (slots, DJ, DM) with
        DJ[(slot, H, c)] = {(G, a): coefficient expression}."""
        if self.coordinates_as_dofs:
            raise RuntimeError("analytic Hessian with position dofs is outside the GPU path (second-order moving-mesh tensors)")
        form = self.derive(resname)
        unknowns = set(self.unknown_field_names())

        def second(coefs):
            out: Dict[Tuple[int, str, str], Dict[Tuple[str, str], sp.Expr]] = {}
            for (si, G, a), c in coefs.items():
                for sym in [x for x in c.free_symbols if x in self._atom_syms]:
                    info = self._atom_syms[sym]
                    if info.past or info.field not in unknowns:
                        continue
                    d = sp.diff(c, sym)
                    if d == 0:
                        continue
                    if info.dt_order:
                        d = d * sp.Symbol("W__%s__%d" % (info.scheme, info.dt_order), real=True)
                    dst = out.setdefault((si, info.field, info.deriv), {})
                    dst[(G, a)] = dst.get((G, a), sp.Integer(0)) + d
            return out
        return form, second(form.J), second(form.M)

    def derive_hessian_transposed(self, resname: str = ""):
        """Coefficient form of the TRANSPOSED contraction of HessianVectorProduct<i> (flags 4 and 5, src/jitbridge.h:637-691,
        src/codegen.cpp:3879-3905):  T_ik = sum_j H_jik Y_j = d((A^T.Y)_i)/dU_k  for A = J or M.  With A_ji = T_b[l_j] C_{s,(G,a)} S_a[l_i]
        the vector is contracted with the TEST side (Y interpolated like the tested field F of slot s with the slot's derivative b),
        the row is the former column (G, a), which becomes a test slot of the new form:

            T[(G,l_i),(H,l_k)] = S_a[l_i] * ( sum_{s=(F,b)} Yhat_{F,b} * dC_{s,(G,a)}/d atom(H,c) * fac ) * S_c[l_k]

        Returns (slots', TJ, TM) with TJ[(slot' index, H, c)] = {(F, b): coefficient expression}."""
        if self.coordinates_as_dofs:
            raise RuntimeError("analytic Hessian with position dofs is outside the GPU path (second-order moving-mesh tensors)")
        form = self.derive(resname)
        unknowns = set(self.unknown_field_names())
        slots_t: List[TestSlot] = []

        def slot_t(G, a):
            sl = TestSlot(G, a)
            if sl not in slots_t:
                slots_t.append(sl)
            return slots_t.index(sl)

        def second(coefs):
            out: Dict[Tuple[int, str, str], Dict[Tuple[str, str], sp.Expr]] = {}
            for (si, G, a), c in sorted(coefs.items()):
                F, b = form.slots[si].field, form.slots[si].deriv
                for sym in sorted((x for x in c.free_symbols if x in self._atom_syms), key=lambda x: x.name):
                    info = self._atom_syms[sym]
                    if info.past or info.field not in unknowns:
                        continue
                    d = sp.diff(c, sym)
                    if d == 0:
                        continue
                    if info.dt_order:
                        d = d * sp.Symbol("W__%s__%d" % (info.scheme, info.dt_order), real=True)
                    dst = out.setdefault((slot_t(G, a), info.field, info.deriv), {})
                    dst[(F, b)] = dst.get((F, b), sp.Integer(0)) + d
            return out
        TJ, TM = second(form.J), second(form.M)
        return slots_t, TJ, TM

    def hessian_form(self, resname: str = "", transposed: bool = False) -> ResidualForm:
        """d((J.Y))/dU and d((M.Y))/dU (transposed: d((J^T.Y))/dU, d((M^T.Y))/dU) as an ordinary coefficient form: the direction
        vector Y enters as auxiliary fields ``Y__<field>`` interpolated like <field>, so the batched R/J/M kernel skeleton
        assembles it unchanged (no residual)."""
        key = resname + ("|hessianT" if transposed else "|hessian")
        if key in self._forms:
            return self._forms[key]
        if self.coordinates_as_dofs:
            hf = self._hessian_form_moving_mesh(resname, key, transposed)
            self._forms[key] = hf
            return hf
        if transposed:
            slots_t, DJ, DM = self.derive_hessian_transposed(resname)
            form = ResidualForm(key, slots_t, [sp.Integer(0)] * len(slots_t), {}, {}, [])
        else:
            form, DJ, DM = self.derive_hessian(resname)

        def fold(D):
            out: Dict[Tuple[int, str, str], sp.Expr] = {}
            for k, terms in D.items():
                e = sp.Integer(0)
                for (G, a), c in terms.items():
                    yf = "Y__" + G
                    if yf not in self.fields:
                        self.fields[yf] = Field(yf, self.fields[G].space, -1, aux_of=G)
                    e = e + self._atom(AtomInfo(yf, 0, "", a, 0)) * c
                out[k] = e
            return out
        J, M = fold(DJ), fold(DM)
        used = set()
        for e in list(J.values()) + list(M.values()):
            used |= {s_ for s_ in e.free_symbols if s_ in self._atom_syms}
        atoms = sorted((self._atom_syms[s_] for s_ in used), key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        allsyms = set().union(*[e.free_symbols for e in list(J.values()) + list(M.values())]) if (J or M) else set()
        hf = ResidualForm(key, list(form.slots), [sp.Integer(0)] * len(form.slots), J, M, atoms,
                          uses_dx=ex.DX_EUL in allsyms, uses_dX=(ex.DX_LAG in allsyms or ex.ELEMSIZE_LAG in allsyms or ex.ELEMSIZE_LAG_CART in allsyms))
        self._forms[key] = hf
        return hf

    def _hessian_form_moving_mesh(self, resname: str, key: str, transposed: bool = False) -> ResidualForm:
        """d((J.Y))/dU and d((M.Y))/dU when the nodal positions are unknowns (the reference's second-order tensors d2_dx2_shape_dcoord,
        int_pt_weights_d2_coords, src/elements.cpp:3163-3217, src/codegen.cpp:1500-1881).  (A.Y)_i = sum_s T_s[l_i] sum_(G,a) C_{s,(G,a)} Yhat_{G,a}
        is itself a weak form that is linear in the test functions, with Y interpolated like the unknowns (values, gradients, and the
        POSITION columns of A contracted with the position part of Y).  Its derivative with respect to every unknown -- fields and
        positions, through the measure, the Eulerian gradients of fields, of Y and of the test functions, and the radius of an
        axisymmetric system -- is what `_coefficient_form` computes for any such form: the first-order identities applied once more.
        Transposed (flags 4 / 5): (A^T.Y)_i = sum_(s,G,a) S_a[l_i] C_{s,(G,a)} Yhat_{F_s,b_s} -- the former column (G, a) is the test slot, Y is
        interpolated like the tested field with the slot's derivative."""
        form = self.derive(resname)

        def contracted(coefs) -> sp.Expr:
            E = sp.Integer(0)
            for (si, G, a), c in coefs.items():
                yfield, yderiv, slot = (form.slots[si].field, form.slots[si].deriv, TestSlot(G, a)) if transposed else (G, a, form.slots[si])
                yf = "Y__" + yfield
                if yf not in self.fields:
                    self.fields[yf] = Field(yf, self.fields[yfield].space, -1, aux_of=yfield)
                E = E + self._test_atom(slot) * c * self._atom(AtomInfo(yf, 0, "", yderiv, 0))
            return E
        fJ = self._coefficient_form(contracted(form.J), key + "|J")
        slots = list(fJ.slots)
        J = dict(fJ.J)
        M: Dict[Tuple[int, str, str], sp.Expr] = {}
        if form.M:
            fM = self._coefficient_form(contracted(form.M), key + "|M")
            for (si, H, c), v in fM.J.items():
                sl = fM.slots[si]
                if sl not in slots:
                    slots.append(sl)
                M[(slots.index(sl), H, c)] = v
        used = set()
        for e in list(J.values()) + list(M.values()):
            used |= {s_ for s_ in e.free_symbols if s_ in self._atom_syms}
        atoms = sorted((self._atom_syms[s_] for s_ in used), key=lambda a: (a.field, a.dt_order, a.deriv, a.past))
        allsyms = set().union(*[e.free_symbols for e in list(J.values()) + list(M.values())]) if (J or M) else set()
        return ResidualForm(key, slots, [sp.Integer(0)] * len(slots), J, M, atoms, uses_dx=ex.DX_EUL in allsyms, uses_dX=(ex.DX_LAG in allsyms or ex.ELEMSIZE_LAG in allsyms or ex.ELEMSIZE_LAG_CART in allsyms))

    def atom_symbol(self, info: AtomInfo) -> sp.Symbol:
        return self._atom(info)

    def max_dt_order(self) -> int:
        m = 0
        for form in self._all_forms():
            for a in form.atoms:
                m = max(m, a.dt_order)
        return m

    def history_levels(self) -> int:
        """Number of nodal history values the routines read (T in SURVEY 8d)."""
        t = 1
        for form in self._all_forms():
            for a in form.atoms:
                if a.dt_order == 1:
                    t = max(t, 2 if a.scheme == "BDF1" else 3)
                if a.dt_order == 2:
                    t = max(t, 5)
                t = max(t, a.past + 1)
        return t


# ---------------------------------------------------------------------------------------------
# Equations front end (pyoomph/generic/codegen.py:1852 Equations)
# ---------------------------------------------------------------------------------------------
class Equations:
    def __init__(self):
        self._code: Optional[FiniteElementCode] = None
        self._children: List["Equations"] = []

    def get_current_code_generator(self) -> FiniteElementCode:
        assert self._code is not None
        return self._code

    def get_nodal_dimension(self) -> int:
        return self._code.nodal_dim

    def define_fields(self):
        pass

    def define_residuals(self):
        pass

    def define_additional_functions(self):
        """integral / local expressions (pyoomph/generic/codegen.py Equations.define_additional_functions)"""
        pass

    def add_integral_function(self, name: str, expr, with_gradient: bool = False):
        """pyoomph/generic/codegen.py:1251: the integrand carries its own measure (multiply by ``self.get_dx()``); with_gradient also
        registers the residual contribution "d_integral_<name>" (FiniteElementCode.add_integral_gradient)"""
        self._code.add_integral_function(name, expr)
        if with_gradient:
            self._code.add_integral_gradient(name)

    def add_local_function(self, name: str, expr):
        """pyoomph/generic/codegen.py:1213: quantity evaluated node-wise on output"""
        self._code.add_local_function(name, expr)

    def add_extremum_function(self, name: str, expr):
        self._code.add_extremum_function(name, expr)

    def add_spatial_error_estimator(self, expr):
        """pyoomph/generic/codegen.py:2126: flux terms of the Z2 error estimator"""
        self._code.add_Z2_flux(expr)

    def get_dx(self, lagrangian: bool = False, coordsys=None):
        """measure of the element's (or the given) coordinate system: dx, or 2 pi r dx when axisymmetric"""
        return (coordsys or self._code.coordinate_system).integral_dx(lagrangian)

    def define_scalar_field(self, name: str, space: str, **_scaling):
        self._code.define_scalar_field(name, space)

    def define_vector_field(self, name: str, space: str, dim: Optional[int] = None, **_scaling):
        self._code.define_vector_field(name, space, dim)

    def activate_coordinates_as_dofs(self, coordinate_space: Optional[str] = None):
        self._code.coordinates_as_dofs = True

    def add_residual(self, expr, destination: str = ""):
        self._code.add_residual(expr, destination)

    def get_global_parameter(self, name: str):
        return self._code._global_param_symbol(name)

    def __add__(self, other: "Equations") -> "Equations":
        return CombinedEquations([self, other])


class CombinedEquations(Equations):
    """``eqs_a + eqs_b`` (pyoomph/generic/codegen.py CombinedEquations)."""

    def __init__(self, parts: Sequence[Equations]):
        super().__init__()
        self.parts: List[Equations] = []
        for p in parts:
            self.parts += p.parts if isinstance(p, CombinedEquations) else [p]

    def define_fields(self):
        for p in self.parts:
            p._code = self._code
            p.define_fields()

    def define_residuals(self):
        for p in self.parts:
            p.define_residuals()

    def define_additional_functions(self):
        for p in self.parts:
            p.define_additional_functions()
