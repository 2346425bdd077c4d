"""Equation classes that define the BASELINE configs, written against the same front-end calls
as the reference's own classes so the definitions read alike:

* ``PoissonEquation``          /root/reference/pyoomph/equations/poisson.py:35-75
* ``StokesEquations`` / ``NavierStokesEquations`` (Taylor-Hood)  pyoomph/equations/navier_stokes.py:148-344, :422-503
* ``TransientHeatEquation``    partial_t(u) + Poisson part (config 3; cf. pyoomph/equations/advection_diffusion.py)
* ``PseudoElasticMesh``        pyoomph/equations/ALE.py:97-146

Scaling/non-dimensionalisation factors of the reference are all 1 here (no units in the configs).
"""
from __future__ import annotations

from typing import Optional

from .codegen import Equations
from .expressions import (Weak, cartesian, div, dot, grad, identity_matrix, material_derivative, mesh_velocity, partial_t, rational_num, sym, trace, var,
                          var_and_test, weak)


class PoissonEquation(Equations):
    """-div(coeff*grad(u)) = f  (poisson.py:35)."""

    def __init__(self, name: str = "u", *, space: str = "C2", source=None, coefficient=1):
        super().__init__()
        self.name, self.space, self.source, self.coefficient = name, space, source, coefficient

    def define_fields(self):
        self.define_scalar_field(self.name, self.space)

    def define_residuals(self):
        u, u_test = var_and_test(self.name)
        self.add_residual(weak(self.coefficient * grad(u), grad(u_test)))
        if self.source is not None:
            src = self.source() if callable(self.source) else self.source
            self.add_residual(-weak(src, u_test))


class TransientHeatEquation(PoissonEquation):
    """partial_t(u) - div(coeff*grad(u)) = f, default time scheme of the problem (BDF2 for config 3)."""

    def __init__(self, name: str = "u", *, space: str = "C2", source=None, coefficient=1, capacity=1):
        super().__init__(name, space=space, source=source, coefficient=coefficient)
        self.capacity = capacity

    def define_residuals(self):
        super().define_residuals()
        u, u_test = var_and_test(self.name)
        self.add_residual(weak(self.capacity * partial_t(u), u_test))


class StokesEquations(Equations):
    """Stokes flow, Taylor-Hood: velocity in C2, pressure in C1 (navier_stokes.py:148-344)."""

    def __init__(self, *, dynamic_viscosity=1.0, bulkforce=None, velocity_name="velocity", pressure_name="pressure",
                 pressure_sign_flip=False, pressure_factor=1, with_azimuthal_velocity=False):
        super().__init__()
        self.with_azimuthal_velocity = with_azimuthal_velocity      # axisymmetric flow with swirl: third velocity component
        self.dynamic_viscosity = dynamic_viscosity
        self.bulkforce = bulkforce
        self.velocity_name, self.pressure_name = velocity_name, pressure_name
        self.pressure_sign_flip, self.pressure_factor = pressure_sign_flip, pressure_factor

    def define_fields(self):
        self.define_vector_field(self.velocity_name, "C2", dim=3 if self.with_azimuthal_velocity else None)
        self.define_scalar_field(self.pressure_name, "C1")

    def define_stress_tensor(self):
        u, p = var(self.velocity_name), var(self.pressure_name)
        strain = sym(grad(u))
        return 2 * self.dynamic_viscosity * strain - identity_matrix() * self.pressure_factor * p * (-1 if self.pressure_sign_flip else 1)

    def define_residuals(self):
        u, u_test = var_and_test(self.velocity_name)
        p, p_test = var_and_test(self.pressure_name)
        stress_tensor = self.define_stress_tensor()
        Dv = grad(u_test)   # symmetric_test_function 'auto' resolves to the plain gradient for Cartesian TH
        self.add_residual(weak(stress_tensor, Dv))
        self.add_residual(weak(div(u), p_test))
        if self.bulkforce is not None:
            self.add_residual(-weak(self.bulkforce, u_test))


class NavierStokesEquations(StokesEquations):
    """Adds rho*(dt_factor*partial_t(u) + nonlinear_factor*(u.grad)u) (navier_stokes.py:422-503)."""

    def __init__(self, *, dynamic_viscosity=1.0, mass_density=1.0, bulkforce=None, dt_factor=1, nonlinear_factor=1, **kw):
        super().__init__(dynamic_viscosity=dynamic_viscosity, bulkforce=bulkforce, **kw)
        self.mass_density, self.dt_factor, self.nonlinear_factor = mass_density, dt_factor, nonlinear_factor

    def define_residuals(self):
        super().define_residuals()
        u, u_test = var_and_test(self.velocity_name)
        rho = self.mass_density
        self.add_residual(weak(rho * material_derivative(u, u, dt_factor=self.dt_factor, advection_factor=self.nonlinear_factor), u_test))


class PseudoElasticMesh(Equations):
    """Moving mesh as a linear-elastic pseudo solid in Lagrangian coordinates (ALE.py:97-146)."""

    def __init__(self, E=1, nu=rational_num(3, 10), coordsys=cartesian):
        super().__init__()
        self.E, self.nu, self.coordsys = E, nu, coordsys      # the mesh equations stay Cartesian by default (ALE.py:117)

    def define_fields(self):
        self.activate_coordinates_as_dofs()

    def define_residuals(self):
        E, nu = self.E, self.nu
        mu = E / 2 / (1 + nu)
        lmbda = E * nu / (1 + nu) / (1 - 2 * nu)
        lmbda = 2 * mu * lmbda / (lmbda + 2 * mu)
        eps = lambda v: sym(grad(v, lagrangian=True, coordsys=self.coordsys))
        sigma = lambda v: lmbda * trace(eps(v)) * identity_matrix() + 2 * mu * eps(v)
        x, x_test = var_and_test("mesh")
        X = var("lagrangian")
        self.add_residual(Weak(sigma(x - X), eps(x_test), coordinate_system=self.coordsys))


class NonlinearHeatEquation(Equations):
    """(1 + beta*u^2) partial_t(u) - div((1 + alpha*u) grad(u)) = 0: a solution-dependent mass matrix and conductivity, so that
    both Hessians d(J.Y)/dU and d(M.Y)/dU are non-trivial (test problem for the HessianVectorProduct path)."""

    def __init__(self, name: str = "u", alpha=0.5, beta=0.25):
        super().__init__()
        self.name, self.alpha, self.beta = name, alpha, beta

    def define_fields(self):
        self.define_scalar_field(self.name, "C2")

    def define_residuals(self):
        u, u_test = var_and_test(self.name)
        self.add_residual(weak((1 + self.beta * u ** 2) * partial_t(u), u_test) + weak((1 + self.alpha * u) * grad(u), grad(u_test)))


class IntegralObservables(Equations):
    """Named integrals over the domain, ``IntegralObservables(volume=1, kinetic_energy=lambda: dot(u, u) / 2)``
    (pyoomph/equations/generic.py:684-699): every integrand is multiplied by the measure of the coordinate system (or of
    ``_coordinate_system``; Lagrangian if ``_lagrangian``) and registered with ``add_integral_function``."""

    def __init__(self, _coordinate_system=None, _lagrangian: bool = False, _with_gradients: bool = False, **integral_observables):
        super().__init__()
        self._coordinate_system, self._lagrangian, self._with_gradients = _coordinate_system, _lagrangian, _with_gradients
        self.integral_observables = dict(integral_observables)

    def define_additional_functions(self):
        dx = self.get_dx(lagrangian=self._lagrangian, coordsys=self._coordinate_system)
        for k, v in self.integral_observables.items():
            v = v() if callable(v) else v
            self.add_integral_function(k, v * dx, with_gradient=self._with_gradients)


class LocalExpressions(Equations):
    """``LocalExpressions(vorticity=lambda: ..., speed=...)`` (pyoomph/equations/generic.py): quantities evaluated node-wise on output,
    registered with ``add_local_function``."""

    def __init__(self, **local_expressions):
        super().__init__()
        self.local_expressions = dict(local_expressions)

    def define_additional_functions(self):
        for k, v in self.local_expressions.items():
            self.add_local_function(k, v() if callable(v) else v)


class ExtremumObservables(Equations):
    """named expressions whose extremum over the mesh is sought (Mesh::evaluate_extremum, src/mesh.cpp:444)"""

    def __init__(self, **extremum_observables):
        super().__init__()
        self.extremum_observables = dict(extremum_observables)

    def define_additional_functions(self):
        for k, v in self.extremum_observables.items():
            self.add_extremum_function(k, v() if callable(v) else v)


class SpatialErrorEstimator(Equations):
    """``SpatialErrorEstimator(grad(var("u")))``: flux terms of the Z2 error estimator (pyoomph/generic/codegen.py:2126)"""

    def __init__(self, *fluxes):
        super().__init__()
        self.fluxes = list(fluxes)

    def define_additional_functions(self):
        for f in self.fluxes:
            self.add_spatial_error_estimator(f() if callable(f) else f)


# ---------------------------------------------------------------------------------------------------------------------------------
# Interface equations: element classes on the boundary edges of a bulk mesh (InterfaceEquations, pyoomph/generic/codegen.py;
# interface element classes src/elements.hpp:1435-2298).  They see the bulk fields at their nodes under the same names and nodal
# indices, add their own interface fields behind them, integrate with the surface measure, and grad / div are the surface operators.
# ---------------------------------------------------------------------------------------------------------------------------------
class DeclareFields(Equations):
    """fields that exist in the nodal record without a residual of their own on this element class: the bulk fields an interface class
    reads or merely skips over, or -- on the bulk side -- the slots of the interface fields (the reference resizes the nodal Data of
    the interface nodes instead, index_of_first_value_assigned_by_face_element, src/elements.cpp:12107-12193)"""

    def __init__(self, **fields):
        super().__init__()
        self.fields = dict(fields)          # name -> space; a name ending in "*" is a vector field

    def define_fields(self):
        for n, sp_ in self.fields.items():
            if n.endswith("*"):
                self.define_vector_field(n[:-1], sp_)
            else:
                self.define_scalar_field(n, sp_)


class RobinBC(Equations):
    """alpha*(u - u_ext) flux through the boundary (Neumann for alpha*u_ext alone): pyoomph/equations/poisson.py PoissonFarFieldMonopoleCondition
    / NeumannBC family; the interface elements of BASELINE config 1's tutorial script"""

    def __init__(self, name: str = "u", alpha=1, external=0, flux=0):
        super().__init__()
        self.name, self.alpha, self.external, self.flux = name, alpha, external, flux

    def define_fields(self):
        self.define_scalar_field(self.name, "C2")

    def define_residuals(self):
        u, u_test = var_and_test(self.name)
        ext = self.external() if callable(self.external) else self.external
        self.add_residual(weak(self.alpha * (u - ext) - self.flux, u_test))


class FreeSurfaceOnFixedMesh(Equations):
    """The interface terms of NavierStokesFreeSurface (pyoomph/equations/navier_stokes.py:520-676) on a mesh that does not move:
    surface tension  sigma * div_S(v)  on the velocity test functions, and the no-penetration condition  u.n = 0  imposed by a
    Lagrange multiplier field that lives on the interface nodes (the kinematic condition with the mesh velocity set to zero).
    Uses the surface divergence, the unit normal and an interface field: everything an interface element adds except the
    coordinate derivatives of the normal (moving meshes)."""

    def __init__(self, surface_tension=1.0, velocity_name: str = "velocity", pressure_name: str = "pressure", lagrange_name: str = "_kin_bc"):
        super().__init__()
        self.sigma, self.velocity_name, self.pressure_name, self.lagrange_name = surface_tension, velocity_name, pressure_name, lagrange_name

    def define_fields(self):
        self.define_vector_field(self.velocity_name, "C2")
        self.define_scalar_field(self.pressure_name, "C1")      # bulk field: only its slot in the nodal record matters here
        self.define_scalar_field(self.lagrange_name, "C2")

    def define_residuals(self):
        u, v = var_and_test(self.velocity_name)
        l, ltest = var_and_test(self.lagrange_name)
        n = var("normal")
        self.add_residual(weak(self.sigma, div(v)))
        self.add_residual(weak(dot(u, n), ltest) + weak(l, dot(n, v)))


class NavierStokesFreeSurface(Equations):
    """Kinematic and dynamic boundary condition of a free surface (pyoomph/equations/navier_stokes.py:520-676), the interface class of
    BASELINE config 4.  On a MOVING mesh (the bulk class solves for the nodal positions: ``static_interface=False`` / "auto" with
    coordinates as dofs) the Lagrange multiplier field ``_kin_bc`` enforces  (mesh_velocity - u).n = 0  and acts on the position
    equations:

        weak((mesh_velocity() - u).n, l_test)  -  weak(l, n.x_test)  +  weak(sigma, div_S(u_test))  [+ weak(traction, n.u_test)]

    (navier_stokes.py:641-655); on a static mesh it reduces to  -weak(u.n, l_test) + weak(l, n.u_test)  (:628-635).  The normal, the
    surface divergence and the line measure all depend on the nodal positions: their derivatives enter the Jacobian columns of the
    position dofs (codegen._coefficient_form, interface branch)."""

    def __init__(self, *, surface_tension=1, kinbc_name: str = "_kin_bc", static_interface="auto", additional_normal_traction=0,
                 velocity_name: str = "velocity", pressure_name: str = "pressure"):
        super().__init__()
        self.surface_tension, self.kinbc_name, self.static_interface = surface_tension, kinbc_name, static_interface
        self.additional_normal_traction, self.velocity_name, self.pressure_name = additional_normal_traction, velocity_name, pressure_name
        if static_interface not in ("auto", True, False):
            raise RuntimeError("property static_interface must be either 'auto', True or False")

    def define_fields(self):
        # the nodal record of the bulk class: velocity, pressure (only its slot matters here), the multiplier
        self.define_vector_field(self.velocity_name, "C2")
        self.define_scalar_field(self.pressure_name, "C1")
        self.define_scalar_field(self.kinbc_name, "C2")
        if self.static_interface is False:
            self.activate_coordinates_as_dofs()

    def define_residuals(self):
        n = var("normal")
        u, u_test = var_and_test(self.velocity_name)
        l, l_test = var_and_test(self.kinbc_name)
        static = self.static_interface
        if static == "auto":
            static = not self.get_current_code_generator().coordinates_as_dofs
        if static:
            self.add_residual(weak(-dot(u, n), l_test))
            self.add_residual(weak(l, dot(n, u_test)))
        else:
            _, R_test = var_and_test("mesh")
            self.add_residual(weak(dot(mesh_velocity() - u, n), l_test))
            self.add_residual(-weak(l, dot(n, R_test)))
        self.add_residual(weak(self.surface_tension, div(u_test)))
        if self.additional_normal_traction != 0:
            self.add_residual(weak(self.additional_normal_traction, dot(n, u_test)))


class NitscheDirichletBC(Equations):
    """Weakly imposed Dirichlet condition  u = u_D  of  -div(k(u) grad u) = f  on a boundary (Nitsche's method):

        - weak(k grad(u).n, v)  - weak(u - u_D, k grad(v).n)  + weak(gamma (u - u_D), v)

    The normal derivatives of the field AND of the test function are BULK quantities: the class lives on faces seen through their bulk
    elements (element type QuadFace2dC2; in the reference an interface element reaching into `bulk_eleminfo`, src/jitbridge.h:88-120 --
    the same access the evaporation flux of config 4 makes to the gas-side gradient through `opposite_eleminfo`)."""

    def __init__(self, name: str = "u", value=0, conductivity=None, penalty=100.0):
        super().__init__()
        self.name, self.value, self.conductivity, self.penalty = name, value, conductivity, penalty

    def define_fields(self):
        self.define_scalar_field(self.name, "C2")

    def define_residuals(self):
        u, v = var_and_test(self.name)
        n = var("normal")
        k = self.conductivity(u) if callable(self.conductivity) else (1 if self.conductivity is None else self.conductivity)
        uD = self.value() if callable(self.value) else self.value
        self.add_residual(-weak(k * dot(grad(u), n), v) - weak(u - uD, k * dot(grad(v), n)) + weak(self.penalty * (u - uD), v))


class StreamlineDiffusionAdvection(Equations):
    """Advection-diffusion with a streamline-upwind (SUPG) term scaled by the element length h = sqrt(element size)
    (pyoomph/equations/advection_diffusion.py; `var("element_length_h")`, pyoomph/expressions/generic.py:178):

        weak(partial_t(c) + w.grad(c), v) + weak(D grad(c), grad(v)) + weak(tau h (w.grad(c)), w.grad(v))

    One number per element -- the integral of the measure over ALL its integration points -- enters every point of the element."""

    def __init__(self, name: str = "c", wind=(1.0, 0.5), diffusivity=0.01, tau=0.5, cartesian_size: bool = False, lagrangian_size: bool = False):
        super().__init__()
        self.name, self.wind, self.D, self.tau, self.cartesian_size = name, wind, diffusivity, tau, cartesian_size
        self.lagrangian_size = lagrangian_size      # h from the Lagrangian element size: independent of the position dofs of a moving mesh

    def define_fields(self):
        self.define_scalar_field(self.name, "C2")

    def define_residuals(self):
        c, v = var_and_test(self.name)
        gc, gv = grad(c), grad(v)
        wc = sum(self.wind[i] * gc[i, 0] for i in range(len(self.wind)))
        wv = sum(self.wind[i] * gv[i, 0] for i in range(len(self.wind)))
        if self.lagrangian_size:
            import sympy as sp
            h = var("cartesian_element_size_Lagrangian" if self.cartesian_size else "element_size_Lagrangian") ** sp.Rational(1, len(self.wind))
        elif self.cartesian_size:
            import sympy as sp
            h = sp.sqrt(var("cartesian_element_size_Eulerian"))
        else:
            h = var("element_length_h")
        self.add_residual(weak(partial_t(c) + wc, v) + weak(self.D * gc, gv) + weak(self.tau * h * wc, wv))


class IntegralConstraint(Equations):
    """A global constraint  integral(expr) = target  enforced by ONE Lagrange multiplier that is a global parameter of the element class
    (the bordered form of pyoomph's GlobalLagrangeMultiplier + WeakContribution, pyoomph/generic/codegen.py:2927): the class gets

      * the residual term  lambda * d(integral)/dU  (written by the caller as  weak(lambda * d expr/d field, test)  for the field(s) the
        expression depends on: here the common case expr = field),
      * the integral expression itself with its gradient contribution (the dense row), and
      * the multiplier's dense column as the parameter derivative dR/d(lambda).

    The solver then works on [[J, b], [c^T, 0]] with b = dR/dlambda and c = integral_gradient: two vectors per Newton step, reduced over
    the ranks on several GPUs, no dense row inside the CSR matrix."""

    def __init__(self, field: str, name: Optional[str] = None, multiplier: Optional[str] = None):
        super().__init__()
        self.field, self.name = field, name or ("integral_" + field)
        self.multiplier = multiplier or ("lambda_" + field)

    def define_residuals(self):
        from .expressions import global_parameter
        _, test = var_and_test(self.field)
        self.add_residual(weak(global_parameter(self.multiplier), test))

    def define_additional_functions(self):
        self.add_integral_function(self.name, var(self.field) * self.get_dx(), with_gradient=True)
