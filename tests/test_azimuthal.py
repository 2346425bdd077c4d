"""Azimuthal normal-mode expansion (BASELINE config 5; pyoomph/expressions/coordsys.py:967-1200): the symbolic side on the CPU."""
import numpy as np
import sympy as sp

from pyoomph_b200.codegen import Equations, FiniteElementCode
from pyoomph_b200.expressions import AxisymmetryBreakingCoordinateSystem, grad, partial_t, var_and_test, weak


class _Diffusion(Equations):
    def define_fields(self):
        self.define_scalar_field("c", "C2")

    def define_residuals(self):
        c, ct = var_and_test("c")
        self.add_residual(weak(partial_t(c), ct) + weak((1 + c * c) * grad(c), grad(ct)))


def _by_slot(form):
    return {(form.slots[k[0]].field, form.slots[k[0]].deriv, k[1], k[2]): v for k, v in form.J.items()}


def test_scalar_diffusion_gets_the_m_squared_over_r_squared_term_and_no_imaginary_part():
    """c = c0 + eps c1 exp(i m phi):  the eigenproblem operator of  d_t c - div(k(c) grad c)  is the axisymmetric one plus
    k(c0) m^2 / r^2 -- real; the linearisation of k about the base state appears in the value column"""
    m = 3
    code = FiniteElementCode("Quad2dC2", _Diffusion(), name="d", coordinate_system=AxisymmetryBreakingCoordinateSystem(m))
    assert code.residual_names() == ["", AxisymmetryBreakingCoordinateSystem.real_contribution_name]       # the imaginary part vanishes
    base, real = _by_slot(code.derive("")), _by_slot(code.derive(AxisymmetryBreakingCoordinateSystem.real_contribution_name))
    c0 = [a for a in code._atom_syms if code._atom_syms[a].field == "c" and code._atom_syms[a].deriv == "d0" and code._atom_syms[a].dt_order == 0][0]
    r = [a for a in code._atom_syms if code._atom_syms[a].field == "coordinate_x"][0]
    for k in set(base) | set(real):
        d = sp.simplify(real.get(k, 0) - base.get(k, 0))
        if k == ("c", "d0", "c", "d0"):
            dx = [s for s in d.free_symbols if s.name == "M__dx"][0]
            Pi = [s for s in d.free_symbols if s.name == "Pi"][0]          # the reference's truncated Pi is a symbol of its own
            assert sp.simplify(d - 2 * Pi * r * dx * (1 + c0 ** 2) * m ** 2 / r ** 2) == 0
        else:
            assert d == 0, (k, d)
    # the mode fields are gone from the emitted form: everything is evaluated at the base state
    assert all(not a.field.endswith("__M1") for a in code.derive(AxisymmetryBreakingCoordinateSystem.real_contribution_name).atoms)


def test_navier_stokes_with_swirl_mode_zero_is_the_axisymmetric_operator():
    """m = 0: the real contribution's Jacobian and mass matrix ARE the base ones, the imaginary contribution vanishes; m != 0: the
    imaginary part couples through i m / r (continuity: i m u_phi / r; azimuthal momentum: -i m p / r)"""
    from pyoomph_b200.equations import NavierStokesEquations
    cs = AxisymmetryBreakingCoordinateSystem("azimuthal_m")
    code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0, with_azimuthal_velocity=True), name="nsazi", coordinate_system=cs)
    assert code.residual_names() == ["", cs.real_contribution_name, cs.imag_contribution_name] and code.global_params == ["azimuthal_m"]
    mp = code._param_syms["azimuthal_m"]
    fb, fr, fi = code.derive(""), code.derive(cs.real_contribution_name), code.derive(cs.imag_contribution_name)
    b, r, i = _by_slot(fb), _by_slot(fr), _by_slot(fi)
    for k in set(b) | set(r):
        assert sp.simplify(r.get(k, 0).subs(mp, 0) - b.get(k, 0)) == 0, k
    assert all(sp.simplify(v.subs(mp, 0)) == 0 for v in i.values()) and len(i) > 5
    assert {(fr.slots[k[0]].field, k[1]) for k in fr.M} == {(fb.slots[k[0]].field, k[1]) for k in fb.M} and not fi.M
    Pi = [s for s in i[("pressure", "d0", "velocity_phi", "d0")].free_symbols if s.name == "Pi"][0]
    # continuity tested with q: (1/r) d_phi u_phi -> i m u_phi / r, times the measure 2 pi r dx (sign: the class's continuity residual)
    dxs = [s for s in i[("pressure", "d0", "velocity_phi", "d0")].free_symbols if s.name == "M__dx"][0]
    assert sp.simplify(i[("pressure", "d0", "velocity_phi", "d0")] ** 2 - (2 * Pi * dxs * mp) ** 2) == 0
    assert sp.simplify(i[("pressure", "d0", "velocity_phi", "d0")] - i[("velocity_phi", "d0", "pressure", "d0")]) == 0
    # Hessian routines of the contributions exist (the azimuthal Hopf / fold trackers contract them with the eigenvector)
    h = code.hessian_form(cs.real_contribution_name)
    assert len(h.J) > 0


def test_generic_second_derivative_route_reproduces_the_fixed_mesh_hessian_forms():
    """`_hessian_form_moving_mesh` (the contracted form (A.Y) / (A^T.Y) differentiated once more by `_coefficient_form`) against the
    dedicated fixed-mesh derivations `derive_hessian` / `derive_hessian_transposed`: identical coefficients in every field column"""
    from problems import make_problem
    c = make_problem("ns_unsteady", 2)["code"]
    for tr in (False, True):
        a = c.hessian_form("", transposed=tr)
        c.coordinates_as_dofs = True
        try:
            b = c._hessian_form_moving_mesh("", "|generic%d" % tr, tr)
        finally:
            c.coordinates_as_dofs = False
        for A, B in ((a.J, b.J), (a.M, b.M)):
            da = {(a.slots[k[0]], k[1], k[2]): v for k, v in A.items()}
            db = {(b.slots[k[0]], k[1], k[2]): v for k, v in B.items() if not k[1].startswith("coordinate")}
            assert set(da) == set(db) and len(da) == 8 * (A is a.J)
            for k in da:
                assert sp.simplify(da[k] - db[k]) == 0, k


def test_cartesian_normal_mode_expansion():
    """exp(i k z) normal to a 2D Cartesian domain (CartesianCoordinateSystemWithAdditionalNormalMode, coordsys.py:574): k = 0 gives the
    base operator and no imaginary part; the viscous term gains mu k^2; continuity couples to v_z through i k"""
    from problems import make_problem
    c = make_problem("ns_kz", 2)["code"]
    cs = c.coordinate_system
    assert c.residual_names() == ["", "real_contrib_normal_mode_stability", "imag_contrib_normal_mode_stability"] and c.global_params == ["normal_mode_k"]
    fb, fr, fi = c.derive(""), c.derive(cs.real_contribution_name), c.derive(cs.imag_contribution_name)
    k = c._param_syms["normal_mode_k"]
    b, r, i = _by_slot(fb), _by_slot(fr), _by_slot(fi)
    for key in set(b) | set(r):
        assert sp.simplify(r.get(key, 0).subs(k, 0) - b.get(key, 0)) == 0, key
    assert all(sp.simplify(v.subs(k, 0)) == 0 for v in i.values()) and len(i) > 5
    d = sp.factor(r[("velocity_x", "d0", "velocity_x", "d0")] - b[("velocity_x", "d0", "velocity_x", "d0")])
    dx = [s for s in d.free_symbols if s.name == "M__dx"][0]
    assert sp.simplify(d - sp.Float(0.01) * dx * k ** 2) == 0
    assert sp.simplify(i[("pressure", "d0", "velocity_z", "d0")] ** 2 - (dx * k) ** 2) == 0
