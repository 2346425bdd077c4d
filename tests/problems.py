"""Shared problem definitions for the parity tests: the BASELINE configs at sizes the oracle finishes in seconds."""
import numpy as np

from pyoomph_b200.codegen import FiniteElementCode
from pyoomph_b200.equations import (NavierStokesEquations, NonlinearHeatEquation, PoissonEquation, PseudoElasticMesh,
                                    TransientHeatEquation)
from pyoomph_b200.expressions import exp, var, global_parameter
from pyoomph_b200.meshes import assign_equation_numbers


def smooth_field(pos, k, seed):
    """uniform(-1,1) * smooth envelope, fixed seed (SURVEY 8d value distributions)."""
    rng = np.random.default_rng(seed + 17 * k)
    ph = rng.uniform(0, 2 * np.pi, size=pos.shape[1] + 1)
    v = np.ones(pos.shape[0])
    for d in range(pos.shape[1]):
        v = v * np.sin((2 + k + d) * pos[:, d] + ph[d])
    return v + 0.1 * rng.uniform(-1, 1, size=pos.shape[0])


def poisson_source():
    # docs/source/tutorial/spatial/poisson/poisson_2d.py:35-41
    x, y = var("coordinate_x"), var("coordinate_y")
    return 100 * exp(-100 * ((x - 0.5) ** 2 + (y - 0.5) ** 2))


class UnstructuredView:
    """A structured mesh with its elements visited in random order and its nodes relabelled at random, WITHOUT the patch
    hint: what a mesh from an external generator looks like to the assembler (no lattice order to lean on)."""

    def __init__(self, mesh, seed):
        rng = np.random.default_rng(seed)
        eperm = rng.permutation(mesh.n_elem)
        new_of_old = rng.permutation(mesh.n_node)
        old_of_new = np.argsort(new_of_old)
        self.dim, self.N, self.element_type = mesh.dim, mesh.N, mesh.element_type
        self.elem_nodes = np.ascontiguousarray(new_of_old[mesh.elem_nodes[eperm]].astype(np.int32))
        self.node_pos = np.ascontiguousarray(mesh.node_pos[old_of_new])
        self.node_lattice = np.ascontiguousarray(mesh.node_lattice[old_of_new])
        self.boundaries = {k: np.sort(new_of_old[v]) for k, v in mesh.boundaries.items()}
        self.n_elem, self.n_node = mesh.n_elem, mesh.n_node

    def is_vertex(self):
        return np.all(self.node_lattice % 2 == 0, axis=1)


def distort(mesh, amplitude, seed):
    """smooth + random displacement of every node by up to `amplitude` element widths: non-affine elements, the mapping
    Jacobian differs at every Gauss point (mid nodes leave the straight line between their vertices)"""
    h = 1.0 / max(mesh.N)
    rng = np.random.default_rng(seed + 991)
    d = np.stack([smooth_field(mesh.node_pos, 20 + k, seed) for k in range(mesh.dim)], axis=1)
    mesh.node_pos = mesh.node_pos + amplitude * h * (0.6 * d / max(1e-300, np.abs(d).max()) + 0.4 * rng.uniform(-1, 1, size=d.shape))
    return mesh


def make_problem(kind: str, N: int, seed: int = 0, distortion: float = 0.0, unstructured: bool = False):
    """Returns dict(code, mesh, dofmap, vals[T,n_node,nval], pos_hist or None, unsteady(bool), params)."""
    params = {}
    variant = (distortion, unstructured)
    if distortion or unstructured:
        import pyoomph_b200.meshes as _m
        _mk = {"quad": _m.RectangularQuadMesh, "brick": _m.CuboidBrickMesh}

        def _variant(mesh):
            if unstructured:
                mesh = UnstructuredView(mesh, seed + 5)
            if distortion:
                mesh = distort(mesh, distortion, seed)
            return mesh

        def RectangularQuadMesh(n):
            return _variant(_mk["quad"](n))

        def CuboidBrickMesh(n):
            return _variant(_mk["brick"](n))
    else:
        from pyoomph_b200.meshes import CuboidBrickMesh, RectangularQuadMesh
    if kind == "nitsche_face":
        # interface class that reaches into its BULK element (normal derivatives of the field and of the test function on the face):
        # element type QuadFace2dC2 on two boundaries of a (distorted) Q9 mesh, Nitsche's method for a nonlinear diffusion problem
        import pyoomph_b200.meshes as _mm
        from pyoomph_b200.equations import NitscheDirichletBC
        from pyoomph_b200.expressions import var as _var
        bulk = _mm.RectangularQuadMesh(N)
        if distortion:
            bulk = distort(bulk, distortion, seed)
        mesh = _mm.boundary_face_mesh(bulk, ["right", "top", "bottom"])
        code = FiniteElementCode("QuadFace2dC2", NitscheDirichletBC("u", value=lambda: 0.3 + _var("coordinate_x") * _var("coordinate_y"),
                                                                     conductivity=lambda u: 1 + 0.5 * u * u, penalty=40.0), name="nitscheface")
        bulk_code = FiniteElementCode("Quad2dC2", PoissonEquation(source=poisson_source), name="poisson")
        dofmap = assign_equation_numbers(bulk, bulk_code, {"u": bulk.boundaries["left"]}, None)
        vals = np.zeros((1, bulk.n_node, 1))
        vals[0, :, 0] = smooth_field(bulk.node_pos, 0, seed)
        return dict(kind=kind, code=code, mesh=mesh, dofmap=dofmap, vals=vals, pos_hist=None, unsteady=False, params={}, bulk_mesh=bulk, bulk_code=bulk_code)
    if kind in ("robin_if", "robin_if_obs", "freesurf_if", "freesurf_mov_if", "freesurf_mov_axi_if"):
        # interface element classes (InterfaceElementLine1dC2) on boundary edges of a (distorted) Q9 mesh, on the bulk's nodes, nodal
        # values and equation numbers: a Robin condition for the Poisson field of config 1, and the free-surface terms of config 4
        # (surface tension, no-penetration through a Lagrange multiplier field on the interface) on a mesh that does not move
        import pyoomph_b200.meshes as _mm
        from pyoomph_b200.equations import DeclareFields, FreeSurfaceOnFixedMesh, NavierStokesFreeSurface, RobinBC
        from pyoomph_b200.expressions import var as _var
        bulk = _mm.RectangularQuadMesh(N)
        if distortion:
            bulk = distort(bulk, distortion, seed)
        mesh = _mm.boundary_line_mesh(bulk, ["right", "top"] if kind in ("robin_if", "robin_if_obs", "freesurf_mov_axi_if") else ["top", "left"])
        if kind == "robin_if_obs":
            # integral expressions over an INTERFACE (boundary length, mean value, the flux the Robin condition exchanges, the tangential
            # variation of the field, the outward normal integrated along the boundary)
            from pyoomph_b200.equations import IntegralObservables
            from pyoomph_b200.expressions import dot, grad
            obs = IntegralObservables(length=1, mean_u=lambda: _var("u"), exchange=lambda: 2.5 * (_var("u") - _var("coordinate_x") * _var("coordinate_y")),
                                      tangential_variation=lambda: dot(grad(_var("u")), grad(_var("u"))), normal=lambda: _var("normal"))
            code = FiniteElementCode("Line1dC2", RobinBC("u", alpha=2.5, external=lambda: _var("coordinate_x") * _var("coordinate_y"), flux=0.3) + obs, name="robinifobs")
            bulk_code = FiniteElementCode("Quad2dC2", PoissonEquation(source=poisson_source), name="poisson")
            pinned = {"u": bulk.boundaries["left"]}
        elif kind == "robin_if":
            code = FiniteElementCode("Line1dC2", RobinBC("u", alpha=2.5, external=lambda: _var("coordinate_x") * _var("coordinate_y"), flux=0.3), name="robinif")
            bulk_code = FiniteElementCode("Quad2dC2", PoissonEquation(source=poisson_source), name="poisson")
            pinned = {"u": bulk.boundaries["left"]}
        elif kind == "freesurf_mov_axi_if":
            # ... and AXISYMMETRIC, as BASELINE config 4 is (a droplet: r = x, the interface away from the axis): measure 2 pi r ds, surface
            # divergence with its v_r / r term, all of it depending on the position dofs
            code = FiniteElementCode("Line1dC2", NavierStokesFreeSurface(surface_tension=0.7, static_interface=False), name="freesurfmovaxi",
                                     coordinate_system="axisymmetric")
            bulk_code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + PseudoElasticMesh() +
                                          DeclareFields(_kin_bc="C2"), name="aleaxiif", coordinate_system="axisymmetric")
            wall = np.unique(np.concatenate([bulk.boundaries[b] for b in ("bottom", "left")]))
            off_interface = np.setdiff1d(np.arange(bulk.n_node), np.unique(mesh.elem_nodes))
            pinned = {"velocity_x": wall, "velocity_y": wall, "_kin_bc": off_interface}
        elif kind == "freesurf_mov_if":
            # config 4 as BASELINE names it: the free surface of a MOVING mesh (kinematic condition with the mesh velocity, the multiplier
            # acting on the position equations; normal, surface divergence and line measure depend on the position dofs)
            code = FiniteElementCode("Line1dC2", NavierStokesFreeSurface(surface_tension=0.7, static_interface=False), name="freesurfmov")
            bulk_code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + PseudoElasticMesh() +
                                          DeclareFields(_kin_bc="C2"), name="aleif")
            wall = np.unique(np.concatenate([bulk.boundaries[b] for b in ("bottom", "right")]))
            off_interface = np.setdiff1d(np.arange(bulk.n_node), np.unique(mesh.elem_nodes))
            pinned = {"velocity_x": wall, "velocity_y": wall, "_kin_bc": off_interface}
        else:
            code = FiniteElementCode("Line1dC2", FreeSurfaceOnFixedMesh(surface_tension=0.7), name="freesurfif")
            bulk_code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + DeclareFields(_kin_bc="C2"), name="nsif")
            wall = np.unique(np.concatenate([bulk.boundaries[b] for b in ("bottom", "right")]))
            off_interface = np.setdiff1d(np.arange(bulk.n_node), np.unique(mesh.elem_nodes))
            pinned = {"velocity_x": wall, "velocity_y": wall, "_kin_bc": off_interface}     # the multiplier exists on the interface nodes only
        unsteady = kind in ("freesurf_mov_if", "freesurf_mov_axi_if")
        pinned_pos = None
        if bulk_code.coordinates_as_dofs:
            pinned_pos = {"coordinate_x": bulk.boundaries["left" if kind == "freesurf_mov_axi_if" else "right"], "coordinate_y": bulk.boundaries["bottom"]}
        dofmap = assign_equation_numbers(bulk, bulk_code, pinned, pinned_pos)
        assert [f.name for f in code.nodal_fields()] == [f.name for f in bulk_code.nodal_fields()]
        assert code.coordinates_as_dofs == bulk_code.coordinates_as_dofs
        T, nval = max(code.history_levels(), bulk_code.history_levels()) if unsteady else code.history_levels(), code.n_nodal_values
        vals = np.zeros((T, bulk.n_node, nval))
        for t in range(T):
            for f in range(nval):
                vals[t, :, f] = smooth_field(bulk.node_pos, f + 3 * t, seed) * (1.0 - 0.05 * t)
        pos_hist = None
        if code.coordinates_as_dofs:
            pos_hist = np.stack([bulk.node_pos + 1e-3 * (1 + 0.3 * t) * np.stack(
                [smooth_field(bulk.node_pos, 10 + d + 2 * t, seed) for d in range(bulk.dim)], axis=1) for t in range(T)])
        return dict(kind=kind, code=code, mesh=mesh, dofmap=dofmap, vals=vals, pos_hist=pos_hist, unsteady=unsteady, params={}, bulk_mesh=bulk, bulk_code=bulk_code)
    if kind in ("poisson_tet", "heat3d_tet", "ns_tet"):
        # ten-node tetrahedra (BulkElementTetra3dC2 = TElement<3,3>, TGauss<3,3>): Poisson, transient heat, 3D Taylor-Hood P2/P1 NS
        import pyoomph_b200.meshes as _mm
        mesh = _mm.CuboidTetraMesh(N)
        if unstructured:
            raise NotImplementedError
        if distortion:
            mesh = distort(mesh, distortion, seed)
        params = {}
        if kind == "poisson_tet":
            code = FiniteElementCode("Tetra3dC2", PoissonEquation(source=poisson_source), name="poissontet")
            pinned = {"u": np.concatenate([mesh.boundaries["left"], mesh.boundaries["right"]])}
            unsteady = False
        elif kind == "heat3d_tet":
            code = FiniteElementCode("Tetra3dC2", TransientHeatEquation(), name="heat3dtet")
            pinned = {"u": mesh.boundaries["left"]}
            unsteady = True
        else:
            code = FiniteElementCode("Tetra3dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0), name="nstet")
            wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom", "back")]))
            pinned = {"velocity_x": wall, "velocity_y": wall, "velocity_z": wall}
            unsteady = True
        dofmap = assign_equation_numbers(mesh, code, pinned, None)
        T, nval = code.history_levels(), code.n_nodal_values
        vals = np.zeros((T, mesh.n_node, nval))
        for t in range(T):
            for f in range(nval):
                vals[t, :, f] = smooth_field(mesh.node_pos, f + 3 * t, seed) * (1.0 - 0.05 * t)
        return dict(kind=kind, code=code, mesh=mesh, dofmap=dofmap, vals=vals, pos_hist=None, unsteady=unsteady, params=params)
    if kind in ("poisson_tri", "ns_tri", "ale_tri", "supg_tri"):
        # six-node triangles (the element class of the reference's gmsh droplet meshes): Poisson, Taylor-Hood P2/P1 Navier-Stokes,
        # and NS on a pseudo-elastic moving mesh
        import pyoomph_b200.meshes as _mm
        mesh = _mm.RectangularTriangleMesh(N)
        if unstructured:
            raise NotImplementedError
        if distortion:
            mesh = distort(mesh, distortion, seed)
        if kind == "supg_tri":                    # element sizes on triangles
            from pyoomph_b200.equations import StreamlineDiffusionAdvection
            code = FiniteElementCode("Tri2dC2", StreamlineDiffusionAdvection(), name="supgtri")
            pinned = {"c": mesh.boundaries["left"]}
            unsteady = True
        elif kind == "poisson_tri":
            code = FiniteElementCode("Tri2dC2", PoissonEquation(source=poisson_source), name="poissontri")
            pinned = {"u": np.concatenate([mesh.boundaries["left"], mesh.boundaries["right"]])}
            unsteady = False
        else:
            eqs = NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0)
            if kind == "ale_tri":
                eqs = eqs + PseudoElasticMesh()
            code = FiniteElementCode("Tri2dC2", eqs, name=kind.replace("_", ""))
            wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom")]))
            pinned = {"velocity_x": wall, "velocity_y": wall}
            unsteady = True
    elif kind in ("poisson_hang", "ns_hang", "ns_unsteady_hang", "ale_hang"):
        # one level of quadtree refinement of some elements of a (distorted) Q9 mesh: hanging nodes on the edges between refined and
        # unrefined elements (a14; the same element classes -- hanging is a property of the mesh, the generated code is the same)
        import pyoomph_b200.meshes as _mm
        base = _mm.RectangularQuadMesh(N)
        if distortion:
            base = distort(base, distortion, seed)
        rng = np.random.default_rng(seed + 11)
        flags = np.zeros((N, N), dtype=bool)
        flags[N // 3: N // 3 + max(2, N // 3), N // 4: N // 4 + max(2, N // 2)] = True          # a block inside the mesh
        flags[0, 0] = flags[-1, -2] = True                                                       # corner / boundary elements
        flags |= rng.random((N, N)) < 0.08                                                       # isolated refined elements
        mesh = _mm.refine_quad_mesh(base, flags.ravel())
        if kind == "poisson_hang":
            code = FiniteElementCode("Quad2dC2", PoissonEquation(source=poisson_source), name="poisson")
            pinned = {"u": np.concatenate([mesh.boundaries["left"], mesh.boundaries["right"]])}
            unsteady = False
        elif kind == "ale_hang":
            # hanging nodes on a MOVING mesh: the positions of the hanging nodes hang on their masters' position dofs
            code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + PseudoElasticMesh(), name="ale")
            wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom")]))
            pinned = {"velocity_x": wall, "velocity_y": wall}
            unsteady = True
        else:
            code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0), name="ns")
            wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "right", "bottom", "top")]))
            pinned = {"velocity_x": wall, "velocity_y": wall, "pressure": np.array([0])}
            unsteady = kind == "ns_unsteady_hang"
    elif kind == "heat3d_hang":
        # one level of octree refinement of some Q27 elements: hanging nodes on the faces AND edges between refined and unrefined elements
        import pyoomph_b200.meshes as _mm
        base = _mm.CuboidBrickMesh(N)
        if distortion:
            base = distort(base, distortion, seed)
        flags = np.zeros((N, N, N), dtype=bool)
        flags[N // 2, N // 2, N // 2] = flags[0, 0, 0] = flags[-1, -1, 0] = True
        flags[N // 2, N // 2, min(N - 1, N // 2 + 1)] = True                       # two refined neighbours: a shared face without hanging nodes
        mesh = _mm.refine_brick_mesh(base, flags.ravel())
        code = FiniteElementCode("Brick3dC2", TransientHeatEquation(), name="heat3d")
        pinned = {"u": mesh.boundaries["left"]}
        unsteady = True
    elif kind == "ns_mean_pressure":
        # the pressure level fixed by the global constraint  integral(p) = 0  with a Lagrange multiplier (bordered system) instead of a
        # pinned pressure value
        from pyoomph_b200.equations import IntegralConstraint
        mesh = RectangularQuadMesh(N)
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + IntegralConstraint("pressure"),
                                 name="nsmeanp")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "right", "bottom", "top")]))
        pinned = {"velocity_x": wall, "velocity_y": wall}
        unsteady = False
        params = {"lambda_pressure": 0.0}
    elif kind == "ns_constraint":
        # integral expressions WITH their gradients with respect to the dofs (the dense rows of global constraints in bordered form):
        # mean pressure, kinetic energy, flux through the domain
        from pyoomph_b200.equations import IntegralObservables
        from pyoomph_b200.expressions import dot, grad, var
        mesh = RectangularQuadMesh(N)
        obs = IntegralObservables(_with_gradients=True, pressure_integral=lambda: var("pressure"),
                                  kinetic_energy=lambda: dot(var("velocity"), var("velocity")) / 2,
                                  enstrophy=lambda: (grad(var("velocity_y"))[0] - grad(var("velocity_x"))[1]) ** 2)
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + obs, name="nsconstraint")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom")]))
        pinned = {"velocity_x": wall, "velocity_y": wall}
        unsteady = False
    elif kind in ("supg", "supg_axi"):
        # element sizes (var("element_length_h"), "cartesian_element_size_Eulerian"): one number per element, the integral of the measure
        # over all of its integration points (with 2 pi r when axisymmetric), in a streamline-upwind term
        from pyoomph_b200.equations import StreamlineDiffusionAdvection
        mesh = RectangularQuadMesh(N)
        axi = kind == "supg_axi"
        code = FiniteElementCode("Quad2dC2", StreamlineDiffusionAdvection(cartesian_size=axi), name="supgaxi" if axi else "supg",
                                 coordinate_system="axisymmetric" if axi else "cartesian")
        pinned = {"c": mesh.boundaries["left"]}
        unsteady = True
    elif kind == "poisson":          # config 1
        mesh = RectangularQuadMesh(N)
        code = FiniteElementCode("Quad2dC2", PoissonEquation(source=poisson_source), name="poisson")
        pinned = {"u": np.concatenate([mesh.boundaries["left"], mesh.boundaries["right"]])}
        unsteady = False
    elif kind == "ns":             # config 2: lid-driven cavity, Taylor-Hood, Re=100
        mesh = RectangularQuadMesh(N)
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0), name="ns")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "right", "bottom", "top")]))
        pinned = {"velocity_x": wall, "velocity_y": wall, "pressure": np.array([0])}
        unsteady = False
    elif kind == "ns_unsteady":
        mesh = RectangularQuadMesh(N)
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0), name="ns")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "right", "bottom", "top")]))
        pinned = {"velocity_x": wall, "velocity_y": wall, "pressure": np.array([0])}
        unsteady = True
    elif kind == "ns_param":       # viscosity as global parameter -> dResidual/dParameter routines
        mesh = RectangularQuadMesh(N)

        class _NS(NavierStokesEquations):
            def define_residuals(self):
                self.dynamic_viscosity = global_parameter("mu")
                super().define_residuals()
        code = FiniteElementCode("Quad2dC2", _NS(mass_density=1.0), name="nsp")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "right", "bottom", "top")]))
        pinned = {"velocity_x": wall, "velocity_y": wall}
        unsteady = False
        params = {"mu": 0.013}
    elif kind == "nlheat":         # nonlinear mass matrix: exercises the mass Hessian
        mesh = RectangularQuadMesh(N)
        code = FiniteElementCode("Quad2dC2", NonlinearHeatEquation(), name="nlheat")
        pinned = {"u": mesh.boundaries["left"]}
        unsteady = True
    elif kind in ("supg3d", "supg_tet"):
        # element sizes in three dimensions (element_length_h = cube root of the element volume) on bricks and tetrahedra
        from pyoomph_b200.equations import StreamlineDiffusionAdvection
        import pyoomph_b200.meshes as _mm
        mesh = CuboidBrickMesh(N) if kind == "supg3d" else _mm.CuboidTetraMesh(N)
        code = FiniteElementCode("Brick3dC2" if kind == "supg3d" else "Tetra3dC2", StreamlineDiffusionAdvection(wind=(1.0, 0.5, -0.25)),
                                 name=kind.replace("_", ""))
        pinned = {"c": mesh.boundaries["left"]}
        unsteady = True
    elif kind == "heat3d":         # config 3
        mesh = CuboidBrickMesh(N)
        code = FiniteElementCode("Brick3dC2", TransientHeatEquation(), name="heat3d")
        pinned = {"u": mesh.boundaries["left"]}
        unsteady = True
    elif kind == "ale":            # config 4 bulk part: NS-TH on a pseudo-elastic moving mesh
        mesh = RectangularQuadMesh(N)
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + PseudoElasticMesh(), name="ale")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom")]))
        pinned = {"velocity_x": wall, "velocity_y": wall}
        unsteady = True
    elif kind in ("ns_axi", "ns_axi_swirl"):
        # configs 4/5, element class: axisymmetric Navier-Stokes (r = x, axis at the left boundary), Taylor-Hood; with swirl the
        # azimuthal velocity is a third C2 component (ndof_el = 31, the class the Hopf/azimuthal tracking of config 5 assembles)
        mesh = RectangularQuadMesh(N)
        swirl = kind.endswith("swirl")
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0, with_azimuthal_velocity=swirl),
                                 name="nsswirl" if swirl else "nsaxi", coordinate_system="axisymmetric")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("right", "bottom", "top")]))
        pinned = {"velocity_x": np.unique(np.concatenate([wall, mesh.boundaries["left"]])), "velocity_y": wall}
        if swirl:
            pinned["velocity_phi"] = np.unique(np.concatenate([mesh.boundaries["right"], mesh.boundaries["left"]]))
        unsteady = True
    elif kind in ("ns_obs", "ns_axi_obs", "ale_axi_obs", "heat3d_obs"):
        # integral expressions (IntegralObservables, pyoomph/equations/generic.py:684) next to the flow equations
        from pyoomph_b200.equations import IntegralObservables
        from pyoomph_b200.expressions import dot, grad, partial_t, var
        if kind == "heat3d_obs":
            mesh = CuboidBrickMesh(N)
            obs = IntegralObservables(volume=1, heat=lambda: var("u"), heating_rate=lambda: partial_t(var("u")),
                                      dirichlet_energy=lambda: dot(grad(var("u")), grad(var("u"))) / 2)
            code = FiniteElementCode("Brick3dC2", TransientHeatEquation() + obs, name="heat3dobs")
            pinned = {"u": mesh.boundaries["left"]}
        else:
            mesh = RectangularQuadMesh(N)
            obs = IntegralObservables(volume=1, kinetic_energy=lambda: dot(var("velocity"), var("velocity")) / 2, momentum=lambda: var("velocity"),
                                      pressure_integral=lambda: var("pressure"), dissipation=lambda: 0.01 * dot(grad(var("velocity_x")), grad(var("velocity_x"))))
            eqs = NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + obs
            if kind == "ale_axi_obs":
                eqs = eqs + PseudoElasticMesh()
            code = FiniteElementCode("Quad2dC2", eqs, name=kind.replace("_", ""), coordinate_system="axisymmetric" if "axi" in kind else "cartesian")
            wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom")]))
            pinned = {"velocity_x": wall, "velocity_y": wall}
        unsteady = True
    elif kind in ("ns_pts", "heat3d_pts", "ale_axi_pts"):
        # local expressions (node-wise output), extremum expressions and Z2 error-estimator fluxes next to the equations
        from pyoomph_b200.equations import ExtremumObservables, LocalExpressions, SpatialErrorEstimator
        from pyoomph_b200.expressions import dot, grad, partial_t, var
        if kind == "heat3d_pts":
            mesh = CuboidBrickMesh(N)
            extra = LocalExpressions(flux=lambda: -grad(var("u")), heating=lambda: partial_t(var("u"))) + \
                ExtremumObservables(hottest=lambda: var("u")) + SpatialErrorEstimator(lambda: grad(var("u")))
            code = FiniteElementCode("Brick3dC2", TransientHeatEquation() + extra, name="heat3dpts")
            pinned = {"u": mesh.boundaries["left"]}
        else:
            mesh = RectangularQuadMesh(N)
            extra = LocalExpressions(speed2=lambda: dot(var("velocity"), var("velocity")), p=lambda: var("pressure"),
                                     vorticity=lambda: grad(var("velocity_y"))[0] - grad(var("velocity_x"))[1], strain=lambda: grad(var("velocity"))) + \
                ExtremumObservables(pmax=lambda: var("pressure"), speed2=lambda: dot(var("velocity"), var("velocity"))) + \
                SpatialErrorEstimator(lambda: grad(var("velocity")))
            eqs = NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + extra
            if kind == "ale_axi_pts":
                eqs = eqs + PseudoElasticMesh()
            code = FiniteElementCode("Quad2dC2", eqs, name=kind.replace("_", ""), coordinate_system="axisymmetric" if "axi" in kind else "cartesian")
            wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom")]))
            pinned = {"velocity_x": wall, "velocity_y": wall}
        unsteady = True
    elif kind in ("ns_kz", "ns_kz_ext"):
        # the Cartesian sibling: normal mode exp(i k z) normal to a 2D Cartesian domain (three velocity components x, y, z)
        from pyoomph_b200.equations import DeclareFields
        from pyoomph_b200.expressions import MODE_SUFFIX, CartesianCoordinateSystemWithAdditionalNormalMode
        mesh = RectangularQuadMesh(N)
        eqs = NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0, with_azimuthal_velocity=True)
        base_names = ["velocity_x", "velocity_y", "velocity_z", "pressure"]
        if kind == "ns_kz_ext":
            eqs = eqs + DeclareFields(**{n + MODE_SUFFIX: ("C1" if n == "pressure" else "C2") for n in base_names})
        code = FiniteElementCode("Quad2dC2", eqs, name="nskz" if kind == "ns_kz" else "nskzext", coordinate_system=CartesianCoordinateSystemWithAdditionalNormalMode("normal_mode_k"))
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("right", "bottom", "top", "left")]))
        pinned = {"velocity_x": wall, "velocity_y": wall, "velocity_z": wall}
        if kind == "ns_kz_ext":
            pinned.update({n + MODE_SUFFIX: v for n, v in list(pinned.items())})
        unsteady = True
        params = {"normal_mode_k": 1.0}
    elif kind in ("ns_azi", "ns_azi_ext"):
        # config 5 as BASELINE names it: azimuthal normal-mode expansion exp(i m phi) about the axisymmetric NS base flow with swirl.
        # "ns_azi": the product's class -- base residual + real / imaginary contribution of the angular eigenproblem, the mode fields
        # standing on the dofs of the base fields.  "ns_azi_ext": the checker's class -- the SAME expanded residuals, but the mode fields
        # declared as nodal fields of their own, so that the ordinary Jacobian machinery of the oracle differentiates with respect to them.
        from pyoomph_b200.equations import DeclareFields
        from pyoomph_b200.expressions import MODE_SUFFIX, AxisymmetryBreakingCoordinateSystem
        mesh = RectangularQuadMesh(N)
        eqs = NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0, with_azimuthal_velocity=True)
        base_names = ["velocity_x", "velocity_y", "velocity_phi", "pressure"]
        if kind == "ns_azi_ext":
            eqs = eqs + DeclareFields(**{n + MODE_SUFFIX: ("C1" if n == "pressure" else "C2") for n in base_names})
        code = FiniteElementCode("Quad2dC2", eqs, name="nsazi" if kind == "ns_azi" else "nsaziext", coordinate_system=AxisymmetryBreakingCoordinateSystem("azimuthal_m"))
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("right", "bottom", "top")]))
        pinned = {"velocity_x": np.unique(np.concatenate([wall, mesh.boundaries["left"]])), "velocity_y": wall,
                  "velocity_phi": np.unique(np.concatenate([mesh.boundaries["right"], mesh.boundaries["left"]]))}
        if kind == "ns_azi_ext":
            pinned.update({n + MODE_SUFFIX: v for n, v in list(pinned.items())})
        unsteady = True
        params = {"azimuthal_m": 1.0}
    elif kind in ("supg_ale", "supg_ale_axi"):
        # Lagrangian element sizes on a MOVING mesh (element_size_Lagrangian: constants with respect to the position dofs); axisymmetric:
        # the coordinate system's 2 pi R at the Lagrangian position enters the size
        from pyoomph_b200.equations import StreamlineDiffusionAdvection
        mesh = RectangularQuadMesh(N)
        axi = kind == "supg_ale_axi"
        code = FiniteElementCode("Quad2dC2", StreamlineDiffusionAdvection(lagrangian_size=True, cartesian_size=not axi) + PseudoElasticMesh(),
                                 name=kind.replace("_", ""), coordinate_system="axisymmetric" if axi else "cartesian")
        pinned = {"c": mesh.boundaries["left"]}
        unsteady = True
    elif kind == "ale_axi":        # config 4 bulk part as BASELINE names it: axisymmetric NS-TH on a pseudo-elastic moving mesh
        mesh = RectangularQuadMesh(N)
        code = FiniteElementCode("Quad2dC2", NavierStokesEquations(dynamic_viscosity=0.01, mass_density=1.0) + PseudoElasticMesh(),
                                 name="aleaxi", coordinate_system="axisymmetric")
        wall = np.unique(np.concatenate([mesh.boundaries[b] for b in ("left", "bottom")]))
        pinned = {"velocity_x": wall, "velocity_y": wall}
        unsteady = True
    else:
        raise KeyError(kind)
    pinned_pos = None
    if code.coordinates_as_dofs:
        pinned_pos = {"coordinate_x": mesh.boundaries["left"], "coordinate_y": mesh.boundaries["bottom"]}
    dofmap = assign_equation_numbers(mesh, code, pinned, pinned_pos)
    T = code.history_levels()
    nval = code.n_nodal_values
    vals = np.zeros((T, mesh.n_node, nval))
    for t in range(T):
        for f in range(nval):
            vals[t, :, f] = smooth_field(mesh.node_pos, f + 3 * t, seed) * (1.0 - 0.05 * t)
    if kind in ("ns_azi_ext", "ns_kz_ext"):
        # base and mode fields carry the values the product's class has in its base fields (the contributions are evaluated at the base state)
        from pyoomph_b200.expressions import MODE_SUFFIX
        base = make_problem(kind[:-4], N, seed, distortion, unstructured)
        for f in base["code"].nodal_fields():
            vals[:, :, code.fields[f.name].index] = base["vals"][:, :, f.index]
            vals[:, :, code.fields[f.name + MODE_SUFFIX].index] = base["vals"][:, :, f.index]
    hanging = getattr(mesh, "hanging", None)
    if hanging is not None:
        # hanging values are the interpolation of their masters (Node::value on a hanging node), for every history level
        for f in code.nodal_fields():
            for n, (m, w) in hanging.of_space(f.space).items():
                vals[:, n, f.index] = vals[:, m, f.index] @ w
    pos_hist = None
    if code.coordinates_as_dofs:
        pos_hist = np.stack([mesh.node_pos + 1e-3 * (1 + 0.3 * t) * np.stack(
            [smooth_field(mesh.node_pos, 10 + d + 2 * t, seed) for d in range(mesh.dim)], axis=1) for t in range(T)])
        if hanging is not None:
            for n, (m, w) in hanging.C2.items():         # hanging nodes sit on their masters' interpolation, at every history level
                pos_hist[:, n, :] = np.einsum("k,tkd->td", w, pos_hist[:, m, :])
    return dict(kind=kind, code=code, mesh=mesh, dofmap=dofmap, vals=vals, pos_hist=pos_hist, unsteady=unsteady, params=params)


TIME = dict(t=0.3, dt=0.01, dtprev=0.012, unsteady_steps_done=2)


def make_oracle(pb, **kw):
    from oracle import OracleProblem
    op = OracleProblem(pb["code"], pb["mesh"], pb["dofmap"], pb["vals"], node_pos_hist=pb["pos_hist"], name=pb["code"].name, **kw)
    if pb["unsteady"]:
        op.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
    else:
        op.set_steady()
    if pb["params"]:
        op.set_params([pb["params"][n] for n in pb["code"].global_params])
    return op


def make_gpu(pb, **kw):
    from pyoomph_b200.assembly import B200Assembly
    asm = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, **kw)
    for t in range(pb["vals"].shape[0]):
        asm.set_nodal_values(t, pb["vals"][t])
    if pb["pos_hist"] is not None:
        for t in range(pb["pos_hist"].shape[0]):
            asm.set_nodal_positions(t, pb["pos_hist"][t])
    if pb["unsteady"]:
        asm.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
    else:
        asm.set_steady()
    if pb["params"]:
        asm.set_parameters(**pb["params"])
    return asm


def csr_to_sorted(n, rs, ci, va):
    """canonical form: scipy CSR with sorted indices (the reference's vectors_of_pairs columns are unsorted)."""
    from scipy.sparse import csr_matrix
    A = csr_matrix((va, ci, rs), shape=(n, n))
    A.sort_indices()
    return A


def compare_matrix(A_gpu, A_ref, tol=1e-12):
    """A_ref's pattern (exact zeros dropped, problem.cc:5524) must be a subset of the fixed GPU pattern; values agree
    to `tol` relative to the row scale; GPU-only entries must be numerically zero on that scale."""
    D = (A_gpu - A_ref).tocsr()
    rowmax = np.maximum(abs(A_ref).max(axis=1).toarray().ravel(), 1e-300)
    err = abs(D).max(axis=1).toarray().ravel() / rowmax
    # pattern subset: every reference entry position exists in the GPU pattern
    P = A_gpu.copy(); P.data[:] = 1.0
    Q = A_ref.copy(); Q.data[:] = 1.0
    missing = (Q - Q.multiply(P)).count_nonzero()
    return float(err.max()), int(missing)


def reference_pattern(indptr, indices, vals):
    """What the reference keeps of a matrix assembled on a fixed structural pattern: oomph's sparse assembly stores an entry only if
    fabs(value) > Numerical_zero_for_sparse_assembly = 0.0 (oomph-lib problem.cc:110, :5524; pyoomph's sorted variant
    src/problem.cpp:2146-2277 likewise), columns ascending as in the "maps" assembly (src/problem.cpp:2200, :2268-2274).
    Returns (row_ptr int32, col_idx int32, values) -- the CSR the reference would hand to its solver for these values."""
    indptr, indices, vals = np.asarray(indptr), np.asarray(indices), np.asarray(vals)
    n = indptr.size - 1
    keep = np.abs(vals) > 0.0
    rows = np.repeat(np.arange(n), np.diff(indptr))
    row_ptr = np.concatenate([[0], np.cumsum(np.bincount(rows[keep], minlength=n))]).astype(np.int32)
    return row_ptr, indices[keep].astype(np.int32), vals[keep]


def compare_csr_exact(indptr, indices, vals, ref_csr, tol=1e-12):
    """The north_star bar: after the reference's zero-drop rule the GPU matrix must have the reference's row_ptr / col_idx BIT-EXACT,
    and every value within `tol` RELATIVE TO THE ENTRY.  ref_csr = (row_start, col_index, values) of the oracle in any column order.

    Two counted allowances, both for entries that are sums with cancellation (their bits depend on summation order and FMA contraction
    -- the reference's own bits change with its compiler flags):
      * n_cancel_values: entries whose error exceeds tol*|entry| but not tol*(largest entry of the row);
      * n_cancel_pattern: entries present in only one of the two patterns whose magnitude is below tol*(largest entry of the row),
        i.e. an exact 0.0 on one side and rounding noise on the other.
    Anything else is reported in n_bad_values / n_bad_pattern and must be zero.  Returns a dict of these counts."""
    n = len(indptr) - 1
    rp, ci, va = reference_pattern(indptr, indices, vals)
    B = csr_to_sorted(n, *ref_csr)
    B.eliminate_zeros()
    out = dict(nnz_ref=int(B.nnz), nnz_gpu=int(va.size), n_cancel_values=0, n_cancel_pattern=0, n_bad_values=0, n_bad_pattern=0, worst_entry_rel=0.0)
    rowmax = np.maximum(abs(B).max(axis=1).toarray().ravel(), 1e-300) if B.nnz else np.full(n, 1e-300)
    out["pattern_equal"] = bool(np.array_equal(rp, B.indptr) and np.array_equal(ci, B.indices))
    if out["pattern_equal"]:
        a, b = va, B.data
        rows = np.repeat(np.arange(n), np.diff(rp))
    else:
        # union of the two patterns through (row, col) keys
        rows_a = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
        rows_b = np.repeat(np.arange(n, dtype=np.int64), np.diff(B.indptr))
        ka, kb = rows_a * n + ci, rows_b * n + B.indices          # ascending: rows ascending, columns ascending inside a row
        ku = np.union1d(ka, kb)
        ua, ub = np.zeros(ku.size), np.zeros(ku.size)
        in_a, in_b = np.zeros(ku.size, bool), np.zeros(ku.size, bool)
        ia, ib = np.searchsorted(ku, ka), np.searchsorted(ku, kb)
        ua[ia], ub[ib], in_a[ia], in_b[ib] = va, B.data, True, True
        rows_u = ku // n
        only = in_a != in_b
        mag = np.maximum(np.abs(ua[only]), np.abs(ub[only]))
        small = mag <= tol * rowmax[rows_u[only]]
        out["n_cancel_pattern"] = int(small.sum())
        out["n_bad_pattern"] = int((~small).sum())
        both = ~only
        a, b, rows = ua[both], ub[both], rows_u[both]
    d = np.abs(a - b)
    ok = d <= tol * np.abs(b)
    cancel = ~ok & (d <= tol * rowmax[rows])
    out["n_cancel_values"] = int(cancel.sum())
    out["n_bad_values"] = int((~ok & ~cancel).sum())
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(ok, d / np.maximum(np.abs(b), 1e-300), 0.0)
    out["worst_entry_rel"] = float(rel.max()) if rel.size else 0.0
    return out


def assert_csr_parity(indptr, indices, vals, ref_csr, tol=1e-12, label="", require_exact_pattern=False, max_cancel_fraction=0.30):
    """assert the north_star bar of compare_csr_exact; returns its statistics (printed by the tests so the driver log carries them).
    max_cancel_fraction bounds the counted allowances: on UNIFORM meshes many entries are analytic zeros by the symmetry of the
    reference element (e.g. the integral of psi^p_j d psi_i/dx for symmetric pairs) and come out as rounding noise in the reference and
    here alike (measured: 6 % of the Poisson entries, 15 % of the Taylor-Hood entries on uniform squares), hence the 30 % default;
    the distorted-mesh variants, where no such symmetry exists, are held to 1 %."""
    st = compare_csr_exact(indptr, indices, vals, ref_csr, tol)
    msg = "%s: %r" % (label, st)
    assert st["n_bad_pattern"] == 0 and st["n_bad_values"] == 0, msg
    if require_exact_pattern:
        assert st["pattern_equal"], msg
    assert st["n_cancel_pattern"] + st["n_cancel_values"] <= max_cancel_fraction * max(1, st["nnz_ref"]), msg
    return st
