"""Symbolic weak-form vocabulary of the assembly hot path.

Mirrors the subset of pyoomph's expression API that the five BASELINE configs use
(/root/reference/pyoomph/expressions/generic.py: ``var`` :137, ``grad`` :355, ``weak`` :394,
``Weak`` :429, ``mesh_velocity`` :497, ``partial_t`` :507, ``material_derivative`` :564,
``testfunction`` :647, ``identity_matrix`` :826, ``vector`` :928, ``dot`` :982,
``transpose`` :1043, ``subexpression`` :1058).  In pyoomph these build GiNaC trees inside the C++
core (src/expressions.cpp); GiNaC is not available here, so sympy plays GiNaC's role: a field is
an applied undefined function of the Eulerian coordinates, the Lagrangian coordinates and time,
so that ``sympy.diff`` produces exactly the objects ``GiNaCShapeExpansion::derivative``
(src/codegen.cpp:8190) produces: spatial/temporal derivatives of the shape expansion.

Nothing in this module evaluates numbers; it only builds trees that ``codegen.FiniteElementCode``
turns into the coefficient form the CUDA kernels (and, independently, the CPU oracle) consume.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import sympy as sp
from sympy.core.function import AppliedUndef

# Independent symbols (src/expressions.cpp global symbols x,y,z,X,Y,Z,t)
EUL = sp.symbols("x y z", real=True)
LAG = sp.symbols("X Y Z", real=True)
TIME = sp.Symbol("t", real=True)
# measures: Eulerian dx / Lagrangian dX  (GiNaCSpatialIntegralSymbol, src/codegen.cpp:7562)
DX_EUL = sp.Symbol("M__dx", real=True)
DX_LAG = sp.Symbol("M__dX", real=True)
# unit normal of an interface element at the integration point (var("normal"); GiNaCNormalSymbol, src/codegen.cpp:7780-7823; printed as
# shapeinfo->normal[i] by the reference)
NORMAL = sp.symbols("NRM__0 NRM__1 NRM__2", real=True)
# element sizes (JITShapeInfo_t::elemsize_Eulerian / elemsize_Eulerian_cartesian, src/jitbridge.h:178): the integral of the (coordinate
# system's / Cartesian) measure over the element, one number per element
ELEMSIZE_EUL = sp.Symbol("ESZ__eulerian", positive=True)
ELEMSIZE_EUL_CART = sp.Symbol("ESZ__eulerian_cartesian", positive=True)
# ... and their Lagrangian siblings (src/jitbridge.h:179): integrated over the Lagrangian coordinates, so constants with respect to every dof
# (usable on moving meshes, where the Eulerian sizes would need elemsize_d_coords)
ELEMSIZE_LAG = sp.Symbol("ESZ__lagrangian", positive=True)
ELEMSIZE_LAG_CART = sp.Symbol("ESZ__lagrangian_cartesian", positive=True)

DIRS = ("x", "y", "z")

ExpressionOrNum = Union[sp.Expr, sp.MatrixBase, float, int]


class _Context:
    """The code generator whose residuals are currently being defined (pyoomph::__current_code)."""
    stack: List["object"] = []

    @classmethod
    def current(cls):
        if not cls.stack:
            raise RuntimeError("var()/testfunction() may only be used while an element code is being defined "
                               "(inside Equations.define_residuals)")
        return cls.stack[-1]


def _args(ndim: int):
    return tuple(EUL[:ndim]) + tuple(LAG[:ndim]) + (TIME,)


class FieldFunction(sp.Function):
    """Shape expansion  sum_l U^l psi_l  of a named field (ShapeExpansion, src/codegen.hpp)."""
    is_real = True


class TestFunctionSymbol(sp.Function):
    """Test function psi_{l_test} of a named field (TestFunction, src/codegen.hpp)."""
    is_real = True


MODE_SUFFIX = "__M1"      # name suffix of the perturbation-mode copy of a field (GiNaC_eval_at_expansion_mode(expr, 1))


def _field(name: str) -> sp.Expr:
    code = _Context.current()
    code._require_field(name)
    f = sp.Function("F__" + name, real=True)(*_args(code.nodal_dim))
    cs = code.coordinate_system
    if getattr(cs, "has_normal_mode_expansion", False) and cs.expands(code, name):
        # U = U_base + eps * U_mode * exp(i m phi)   (AxisymmetryBreakingCoordinateSystem.get_mode_expansion_of_var_or_test, coordsys.py:989-1016)
        f1 = sp.Function("F__" + name + MODE_SUFFIX, real=True)(*_args(code.nodal_dim))
        return f + cs.expansion_eps * f1 * cs.field_mode
    return f


def _test(name: str) -> sp.Expr:
    code = _Context.current()
    code._require_field(name)
    t = sp.Function("T__" + name, real=True)(*_args(code.nodal_dim))
    cs = code.coordinate_system
    if getattr(cs, "has_normal_mode_expansion", False) and cs.expands(code, name):
        return t * cs.test_mode              # test functions carry exp(-i m phi)
    return t


def _vector_components(code, name: str) -> Optional[List[str]]:
    if name in ("mesh", "coordinate"):
        return ["coordinate_" + d for d in DIRS[:code.nodal_dim]]
    if name == "lagrangian":
        return ["lagrangian_" + d for d in DIRS[:code.nodal_dim]]
    if name in code.vector_fields:
        return list(code.vector_fields[name])
    return None


_ALIASES = {"mesh_x": "coordinate_x", "mesh_y": "coordinate_y", "mesh_z": "coordinate_z"}


def _padding(code, comps) -> list:
    """zero components up to the vector dimension of the coordinate system (an axisymmetric (r,z) vector is (r,z,0))"""
    return [sp.Integer(0)] * max(0, code.coordinate_system.vector_dimension(code.nodal_dim) - len(comps))


def var(arg: Union[str, Sequence[str]]):
    """Field value(s) by name; vector fields expand to a column of components (generic.py:137)."""
    if not isinstance(arg, str):
        return tuple(var(a) for a in arg)
    code = _Context.current()
    if arg == "time":
        return TIME
    if arg in ("element_size_Eulerian", "cartesian_element_size_Eulerian", "element_length_h", "cartesian_element_length_h",
               "element_size_Lagrangian", "cartesian_element_size_Lagrangian"):
        # pyoomph/expressions/generic.py:174-180; stabilisation terms (SUPG / PSPG, artificial diffusion) are built on them
        if code.etype.elem_dim != code.nodal_dim or code.nodal_dim not in (2, 3):
            raise NotImplementedError("element sizes: bulk elements (two- or three-dimensional) only")
        if arg == "cartesian_element_size_Eulerian":
            return ELEMSIZE_EUL_CART
        if arg == "element_size_Lagrangian":
            return ELEMSIZE_LAG
        if arg == "cartesian_element_size_Lagrangian":
            return ELEMSIZE_LAG_CART
        if arg == "cartesian_element_length_h":
            return ELEMSIZE_EUL_CART ** sp.Rational(1, code.etype.elem_dim)
        return ELEMSIZE_EUL if arg == "element_size_Eulerian" else ELEMSIZE_EUL ** sp.Rational(1, code.etype.elem_dim)
    if arg == "normal":
        if code.etype.elem_dim >= code.nodal_dim and not code.etype.name.startswith("QuadFace"):
            raise RuntimeError("var(\"normal\") is defined on interface elements only")
        return sp.Matrix(list(NORMAL[:code.nodal_dim]) + _padding(code, NORMAL[:code.nodal_dim]))
    comps = _vector_components(code, arg)
    if comps is not None:
        return sp.Matrix([_field(c) for c in comps] + _padding(code, comps))
    return _field(_ALIASES.get(arg, arg))


def testfunction(arg: Union[str, Sequence[str]]):
    """Galerkin test function(s) of a field (generic.py:647)."""
    if not isinstance(arg, str):
        return tuple(testfunction(a) for a in arg)
    code = _Context.current()
    comps = _vector_components(code, arg)
    if comps is not None:
        return sp.Matrix([_test(c) for c in comps] + _padding(code, comps))
    return _test(_ALIASES.get(arg, arg))


def var_and_test(name: str):
    return var(name), testfunction(name)


def global_parameter(name: str) -> sp.Symbol:
    """Problem-level parameter, printed as (*(my_func_table->global_parameters[k])) (src/expressions.cpp:95)."""
    code = _Context.current()
    return code._global_param_symbol(name)


def _is_matrix(a) -> bool:
    return isinstance(a, sp.MatrixBase)


def _coords(lagrangian: bool):
    code = _Context.current()
    return (LAG if lagrangian else EUL)[:code.nodal_dim]


class BaseCoordinateSystem:
    """Differential operators and measure of a coordinate system (pyoomph/expressions/coordsys.py:37): the equation classes
    call the generic grad/div/weak below, which dispatch to the coordinate system of the element code (or to an explicit
    ``coordsys=`` argument, as pyoomph/equations/ALE.py:136-146 does for the mesh equations)."""

    def get_id_name(self) -> str:
        raise NotImplementedError

    def vector_dimension(self, nodal_dim: int) -> int:
        """components of a vector expression (get_actual_dimension, coordsys.py:394)"""
        return nodal_dim

    def integral_dx(self, lagrangian: bool):
        raise NotImplementedError

    def scalar_gradient(self, arg, lagrangian: bool):
        raise NotImplementedError

    def vector_gradient(self, arg, lagrangian: bool):
        raise NotImplementedError

    def vector_divergence(self, arg, lagrangian: bool):
        raise NotImplementedError

    def tensor_divergence(self, arg, lagrangian: bool):
        raise NotImplementedError


class CartesianCoordinateSystem(BaseCoordinateSystem):
    """coordsys.py:253.  Vectors that carry more components than the mesh has dimensions (the zero-padded vectors of an
    axisymmetric element code handed to a Cartesian operator) keep their size: the missing derivatives are zero."""

    def get_id_name(self) -> str:
        return "Cartesian"

    def integral_dx(self, lagrangian: bool):
        return DX_LAG if lagrangian else DX_EUL

    def scalar_gradient(self, arg, lagrangian: bool):
        return sp.Matrix([sp.diff(arg, c) for c in _coords(lagrangian)])

    def vector_gradient(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        n = arg.shape[0]
        return sp.Matrix(n, max(n, len(cs)), lambda i, j: sp.diff(arg[i, 0], cs[j]) if j < len(cs) else sp.Integer(0))

    def vector_divergence(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        return sum(sp.diff(arg[i, 0], cs[i]) for i in range(min(len(cs), arg.shape[0])))

    def tensor_divergence(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        return sp.Matrix([sum(sp.diff(arg[i, j], cs[j]) for j in range(min(len(cs), arg.shape[1]))) for i in range(arg.shape[0])])


class AxisymmetricCoordinateSystem(BaseCoordinateSystem):
    """(r, z) = (x, y), symmetry axis x = 0 (coordsys.py:386-560): vectors have three components (r, z, phi), the azimuthal one
    zero unless the field was defined with it (swirl); dx = 2 pi r dr dz (:403-416); grad of a scalar = (d/dr, d/dz, 0) (:418);
    grad of a vector = [[dr ur, dz ur, -uphi/r], [dr uz, dz uz, 0], [dr uphi, dz uphi, ur/r]] (:434-451 plus the swirl column);
    div u = dr ur + ur/r + dz uz (:475-483).  The radius is the position field coordinate_x, so on a moving mesh all of these
    terms get their position derivatives from the same symbolic differentiation as everything else."""

    def get_id_name(self) -> str:
        return "Axisymmetric"

    def vector_dimension(self, nodal_dim: int) -> int:
        if nodal_dim != 2:
            raise RuntimeError("the axisymmetric coordinate system is built for 2d meshes")
        return 3

    @staticmethod
    def _r(lagrangian: bool):
        return _field("lagrangian_x" if lagrangian else "coordinate_x")

    def integral_dx(self, lagrangian: bool):
        return 2 * pi * self._r(lagrangian) * (DX_LAG if lagrangian else DX_EUL)

    def scalar_gradient(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        return sp.Matrix([sp.diff(arg, cs[0]), sp.diff(arg, cs[1]), sp.Integer(0)])

    def vector_gradient(self, arg, lagrangian: bool):
        if arg.shape[0] != 3:
            raise RuntimeError("Cannot take a 2d axisymmetric vector gradient from a vector with dim!=3: " + str(arg))
        x, y = _coords(lagrangian)
        r = self._r(lagrangian)
        return sp.Matrix([[sp.diff(arg[0, 0], x), sp.diff(arg[0, 0], y), -arg[2, 0] / r],
                          [sp.diff(arg[1, 0], x), sp.diff(arg[1, 0], y), sp.Integer(0)],
                          [sp.diff(arg[2, 0], x), sp.diff(arg[2, 0], y), arg[0, 0] / r]])

    def vector_divergence(self, arg, lagrangian: bool):
        x, y = _coords(lagrangian)
        return sp.diff(arg[0, 0], x) + arg[0, 0] / self._r(lagrangian) + sp.diff(arg[1, 0], y)

    def tensor_divergence(self, T, lagrangian: bool):
        # coordsys.py:530-537 (first index contracted)
        x, y = _coords(lagrangian)
        r = self._r(lagrangian)
        return sp.Matrix([sp.diff(T[0, 0], x) + (T[0, 0] - T[2, 2]) / r + sp.diff(T[1, 0], y),
                          sp.diff(T[0, 1], x) + T[0, 1] / r + sp.diff(T[1, 1], y),
                          sp.diff(T[0, 2], x) + (T[0, 2] - T[2, 0]) / r + sp.diff(T[1, 2], y)])


class AxisymmetryBreakingCoordinateSystem(AxisymmetricCoordinateSystem):
    """Azimuthal normal-mode expansion about an axisymmetric base state (pyoomph/expressions/coordsys.py:967-1200; BASELINE config 5):
    every field is  U_base(r,z) + eps * U_mode(r,z) * exp(i m phi),  every test function carries  exp(-i m phi),  and the operators keep
    their phi-derivatives: grad s = (d_r s, d_z s, d_phi s / r), the third column of grad v gets (d_phi v_r - v_phi)/r, d_phi v_z / r,
    (d_phi v_phi + v_r)/r, div v gets d_phi v_phi / r.  A residual R then yields three contributions (`map_residual_*`):

      base           R at eps = 0, m = 0                                          -> the axisymmetric problem
      real_contrib_azimuthal_stability, imag_contrib_azimuthal_stability
                     Re / Im of dR/d(eps) at eps = 0 with m the azimuthal mode    -> their Jacobians / mass matrices with respect to the
                                                                                   MODE fields are Re/Im of the m-dependent linear operator

    The mode fields are not new nodal values: they stand on the dofs of their base fields (the eigenvector lives on the same equations),
    which is how codegen._coefficient_form differentiates them.  Fixed meshes only."""
    has_normal_mode_expansion = True
    real_contribution_name = "real_contrib_azimuthal_stability"      # pyoomph/generic/problem.py:101-102
    imag_contribution_name = "imag_contrib_azimuthal_stability"

    def __init__(self, angular_mode="azimuthal_m"):
        self.angular_mode = angular_mode if isinstance(angular_mode, str) else sp.sympify(angular_mode)
        self.expansion_eps = sp.Symbol("EPS__mode_expansion", real=True)
        self.m_angular_symbol = sp.Symbol("M__angular", real=True)
        self.phi = sp.Symbol("PHI__angular", real=True)
        self.field_mode = sp.exp(sp.I * self.m_angular_symbol * self.phi)
        self.test_mode = sp.exp(-sp.I * self.m_angular_symbol * self.phi)

    def expands(self, code, fieldname: str) -> bool:
        if fieldname.endswith(MODE_SUFFIX):
            return False
        if fieldname.startswith("lagrangian_") or fieldname.startswith("coordinate_"):
            if code.coordinates_as_dofs and fieldname.startswith("coordinate_"):
                raise NotImplementedError("azimuthal mode expansion on a moving mesh")
            return False
        return True

    def scalar_gradient(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        third = sp.Integer(0) if lagrangian else sp.diff(arg, self.phi) / self._r(lagrangian)
        return sp.Matrix([sp.diff(arg, cs[0]), sp.diff(arg, cs[1]), third])

    def vector_gradient(self, arg, lagrangian: bool):
        G = super().vector_gradient(arg, lagrangian)
        if not lagrangian:
            r = self._r(lagrangian)
            for i in range(3):
                G[i, 2] += sp.diff(arg[i, 0], self.phi) / r
        return G

    def vector_divergence(self, arg, lagrangian: bool):
        d = super().vector_divergence(arg, lagrangian)
        if not lagrangian:
            d += sp.diff(arg[2, 0], self.phi) / self._r(lagrangian)
        return d

    def tensor_divergence(self, T, lagrangian: bool):
        d = super().tensor_divergence(T, lagrangian)
        if not lagrangian:
            r = self._r(lagrangian)
            for j in range(3):
                d[j, 0] += sp.diff(T[2, j], self.phi) / r
        return d

    # ---- the three contributions of a residual (coordsys.py:1018-1041)
    def map_residual_on_base_mode(self, residual):
        return sp.sympify(residual).subs(self.expansion_eps, 0).subs(self.m_angular_symbol, 0).doit()

    def _first_order(self, residual):
        e = sp.diff(sp.expand(sp.sympify(residual)), self.expansion_eps).subs(self.expansion_eps, 0).doit()
        e = sp.expand(sp.powsimp(sp.expand(e)))               # exp(i m phi) exp(-i m phi) -> 1
        if e.has(self.phi):
            raise RuntimeError("the first-order mode expansion still depends on the azimuthal angle: " + str(e))
        m = self.angular_mode
        if isinstance(m, str):               # a global parameter by name (pyoomph: "azimuthal_m", problem.py:103)
            m = _Context.current()._global_param_symbol(m)
        return e.subs(self.m_angular_symbol, m)

    def map_residual_on_angular_eigenproblem_real(self, residual):
        e = self._first_order(residual)
        return sp.expand(e - sp.I * e.coeff(sp.I))

    def map_residual_on_angular_eigenproblem_imag(self, residual):
        return sp.expand(self._first_order(residual).coeff(sp.I))


class CartesianCoordinateSystemWithAdditionalNormalMode(CartesianCoordinateSystem):
    """Normal-mode expansion exp(i k z) in the direction a two-dimensional Cartesian domain does not resolve
    (pyoomph/expressions/coordsys.py:574-760; `Problem.setup_for_stability_analysis(additional_cartesian_mode=True)`): fields
    U_base(x, y) + eps * U_mode(x, y) * exp(i k z), tests with exp(-i k z), vectors with three components (x, y, z), the gradient's third
    entry d/dz, the divergence's d v_z / dz.  Contributions "real_contrib_normal_mode_stability" / "imag_contrib_normal_mode_stability"
    (pyoomph/generic/problem.py:108-109); k is the global parameter "normal_mode_k".  Same machinery as the azimuthal expansion."""
    has_normal_mode_expansion = True
    real_contribution_name = "real_contrib_normal_mode_stability"
    imag_contribution_name = "imag_contrib_normal_mode_stability"

    def __init__(self, normal_mode="normal_mode_k"):
        self.angular_mode = normal_mode if isinstance(normal_mode, str) else sp.sympify(normal_mode)
        self.expansion_eps = sp.Symbol("EPS__mode_expansion", real=True)
        self.m_angular_symbol = sp.Symbol("K__normal_mode", real=True)
        self.phi = sp.Symbol("XADD__normal_mode", real=True)          # the additional coordinate
        self.field_mode = sp.exp(sp.I * self.m_angular_symbol * self.phi)
        self.test_mode = sp.exp(-sp.I * self.m_angular_symbol * self.phi)

    expands = AxisymmetryBreakingCoordinateSystem.expands
    map_residual_on_base_mode = AxisymmetryBreakingCoordinateSystem.map_residual_on_base_mode
    _first_order = AxisymmetryBreakingCoordinateSystem._first_order
    map_residual_on_angular_eigenproblem_real = AxisymmetryBreakingCoordinateSystem.map_residual_on_angular_eigenproblem_real
    map_residual_on_angular_eigenproblem_imag = AxisymmetryBreakingCoordinateSystem.map_residual_on_angular_eigenproblem_imag

    def vector_dimension(self, nodal_dim: int) -> int:
        if nodal_dim != 2:
            raise RuntimeError("the additional normal mode is built for 2d meshes")
        return 3

    def scalar_gradient(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        return sp.Matrix([sp.diff(arg, cs[0]), sp.diff(arg, cs[1]), sp.Integer(0) if lagrangian else sp.diff(arg, self.phi)])

    def vector_gradient(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        n = arg.shape[0]
        return sp.Matrix(n, 3, lambda i, j: sp.diff(arg[i, 0], cs[j]) if j < 2 else (sp.Integer(0) if lagrangian else sp.diff(arg[i, 0], self.phi)))

    def vector_divergence(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        d = sp.diff(arg[0, 0], cs[0]) + sp.diff(arg[1, 0], cs[1])
        if arg.shape[0] > 2 and not lagrangian:
            d += sp.diff(arg[2, 0], self.phi)
        return d

    def tensor_divergence(self, arg, lagrangian: bool):
        cs = _coords(lagrangian)
        return sp.Matrix([sum(sp.diff(arg[i, j], cs[j]) for j in range(2)) + (sp.Integer(0) if (lagrangian or arg.shape[1] < 3) else sp.diff(arg[i, 2], self.phi))
                          for i in range(arg.shape[0])])


cartesian = CartesianCoordinateSystem()
axisymmetric = AxisymmetricCoordinateSystem()


def _coordsys(coordsys=None) -> BaseCoordinateSystem:
    return coordsys if coordsys is not None else _Context.current().coordinate_system


def grad(arg: ExpressionOrNum, lagrangian: bool = False, coordsys: Optional[BaseCoordinateSystem] = None):
    """Gradient: scalar -> column vector, vector -> matrix G[i,j]=d u_i / d x_j (generic.py:355), in the coordinate system
    of the element code unless ``coordsys`` is given."""
    cs = _coordsys(coordsys)
    if _is_matrix(arg):
        if arg.shape[1] != 1:
            raise RuntimeError("grad of a rank-2 tensor is not supported")
        return cs.vector_gradient(arg, lagrangian)
    return cs.scalar_gradient(sp.sympify(arg), lagrangian)


def div(arg, lagrangian: bool = False, coordsys: Optional[BaseCoordinateSystem] = None):
    cs = _coordsys(coordsys)
    if not _is_matrix(arg):
        raise RuntimeError("div needs a vector")
    if arg.shape[1] == 1:
        return cs.vector_divergence(arg, lagrangian)
    return cs.tensor_divergence(arg, lagrangian)


def dot(a, b):
    if _is_matrix(a) and _is_matrix(b):
        if a.shape[1] == 1 and b.shape[1] == 1:
            return sum(a[i, 0] * b[i, 0] for i in range(a.shape[0]))
        if a.shape[1] == 1:  # v . M
            return (a.T * b).T
        return a * b
    return a * b


def contract(a, b):
    """Full contraction of two equal-rank objects (generic.py:470)."""
    if _is_matrix(a) != _is_matrix(b):
        raise RuntimeError("cannot contract objects of different rank")
    if _is_matrix(a):
        if a.shape != b.shape:
            raise RuntimeError("shape mismatch in contract: %s vs %s" % (a.shape, b.shape))
        return sum(a[i, j] * b[i, j] for i in range(a.shape[0]) for j in range(a.shape[1]))
    return sp.sympify(a) * sp.sympify(b)


def double_dot(a, b):
    return contract(a, b)


def weak(a, b, *, lagrangian: bool = False, coordinate_system: Optional[BaseCoordinateSystem] = None):
    """(a,b) = integral of contract(a,b) over the element, Eulerian dx unless lagrangian (generic.py:394); the measure is the
    coordinate system's (2 pi r dx when axisymmetric)."""
    return contract(a, b) * _coordsys(coordinate_system).integral_dx(lagrangian)


def Weak(a, b, *, coordinate_system: Optional[BaseCoordinateSystem] = None):
    """Lagrangian weak form (generic.py:429)."""
    return weak(a, b, lagrangian=True, coordinate_system=coordinate_system)


def transpose(a):
    return a.T


def sym(a):
    return (a + a.T) / 2


def trace(a):
    return a.trace()


def identity_matrix(dim: int = -1):
    if dim < 0:
        code = _Context.current()
        dim = code.coordinate_system.vector_dimension(code.nodal_dim)
    return sp.eye(dim)


def vector(*args):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    return sp.Matrix([sp.sympify(a) for a in args])


def matproduct(a, b):
    return a * b


def dyadic(a, b):
    return a * b.T


def subexpression(what):
    """pyoomph wraps expensive terms for CSE (generic.py:1058); sympy.cse does that globally here."""
    return what


def rational_num(n, d=1):
    return sp.Rational(n, d)


def mesh_velocity():
    """partial_t(var("mesh"), ALE=False) (generic.py:497)."""
    return partial_t(var("mesh"), ALE=False)


def partial_t(f, order: int = 1, ALE: Union[str, bool] = "auto"):
    """Time derivative at fixed local coordinate, ALE-corrected on moving meshes (generic.py:507)."""
    if isinstance(f, str):
        f = var(f)
    if order == 0:
        return f
    code = _Context.current()
    d = (f.diff(TIME, order) if _is_matrix(f) else sp.diff(sp.sympify(f), TIME, order))
    use_ale = (ALE is True) or (ALE == "auto" and code.coordinates_as_dofs)
    if use_ale:
        if order != 1:
            raise ValueError("Currently, I can only take the first order time derivative with ALE")
        d = d - directional_derivative(f, mesh_velocity())
    return d


def directional_derivative(f, direction):
    """(direction . grad) f (generic.py:623)."""
    g = grad(f)
    if _is_matrix(f):
        return g * direction
    return dot(direction, g)


def material_derivative(f, velocity, ALE: Union[str, bool] = "auto", dt_factor=1, advection_factor=1):
    """dt_factor*partial_t(f) + advection_factor*(velocity.grad) f (generic.py:564)."""
    if isinstance(f, str):
        f = var(f)
    if isinstance(velocity, str):
        velocity = var(velocity)
    return dt_factor * partial_t(f, ALE=ALE) + advection_factor * directional_derivative(f, velocity)


def evaluate_in_past(expr, timestep_offset: int = 1):
    """Replace every field by its history value (generic.py:1125); only integer offsets."""
    expr = sp.sympify(expr) if not _is_matrix(expr) else expr
    repl = {}
    for f in expr.atoms(AppliedUndef):
        nm = f.func.__name__
        if nm.startswith("F__"):
            repl[f] = sp.Function(nm + "__past%d" % int(timestep_offset), real=True)(*f.args)
    return expr.xreplace(repl)


# math vocabulary (A.2 of SURVEY: pow, libm functions)
sqrt, exp, log, sin, cos, tan, tanh, atan2 = sp.sqrt, sp.exp, sp.log, sp.sin, sp.cos, sp.tan, sp.tanh, sp.atan2
pi = sp.Symbol("Pi", real=True)  # jitbridge_hang.h:355 defines Pi as 3.14159265359 (truncated); kept symbolic
