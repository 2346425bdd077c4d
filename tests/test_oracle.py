"""CPU tests of the oracle itself: the algorithmic pins the reference offers for this path --
analytic-vs-finite-difference Jacobian (the reference's debug_analytical_jacobian, src/elements.cpp:5880), patch tests,
structure of the Poisson matrix, vectors_of_pairs CSR semantics, BDF weights -- plus the committed golden vectors."""
import json
import os

import numpy as np
import pytest

from problems import csr_to_sorted, make_oracle, make_problem

HERE = os.path.dirname(os.path.abspath(__file__))


def _fd_jacobian(op, pb, e, eps=1e-6):
    dm, vals = pb["dofmap"], pb["vals"]
    R, J, M, eq = op.element(e, flag=2)
    Jfd = np.zeros_like(J)
    pos0 = pb["pos_hist"][0] if pb["pos_hist"] is not None else None
    for j, g in enumerate(eq):
        w = np.argwhere(dm.node_eqn == g)
        res = {}
        if len(w):
            n, f = w[0]
            for sgn in (-1, 1):
                v = vals[0].copy(); v[n, f] += sgn * eps
                op.update_values(0, v)
                res[sgn] = op.element(e, flag=0)[0]
            op.update_values(0, vals[0])
        else:
            n, f = np.argwhere(dm.pos_eqn == g)[0]
            for sgn in (-1, 1):
                p = pos0.copy(); p[n, f] += sgn * eps
                op.update_values(0, None, p)
                res[sgn] = op.element(e, flag=0)[0]
            op.update_values(0, None, pos0)
        Jfd[:, j] = (res[1] - res[-1]) / (2 * eps)
    return J, Jfd


@pytest.mark.parametrize("kind,N", [("poisson", 4), ("ns", 4), ("ns_unsteady", 4), ("heat3d", 2), ("ale", 4), ("ns_param", 3)])
def test_analytic_jacobian_matches_finite_differences(kind, N):
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    J, Jfd = _fd_jacobian(op, pb, pb["mesh"].n_elem // 2)
    assert np.abs(J - Jfd).max() <= 2e-8 * np.abs(J).max()
    op.close()


@pytest.mark.parametrize("kind,N", [("ns", 4), ("heat3d", 2), ("ale", 3), ("ns_axi", 3), ("ns_axi_swirl", 3), ("ale_axi", 3)])
def test_analytic_jacobian_on_distorted_unstructured_meshes(kind, N):
    """non-affine elements (curved edges, mapping Jacobian varying over the Gauss points), random element and node order"""
    pb = make_problem(kind, N, distortion=0.12, unstructured=True)
    op = make_oracle(pb)
    for e in (0, pb["mesh"].n_elem - 1):
        J, Jfd = _fd_jacobian(op, pb, e)
        assert np.abs(J - Jfd).max() <= 5e-8 * np.abs(J).max()
    op.close()


def test_distorted_mesh_patch_test_and_relabelling_invariance():
    """(i) a linear field on a distorted Q9 mesh has an exactly constant gradient: the interior Laplace residual stays at the
    level the mistyped Gauss knots leave; the total area is the sum of the point weights.  (ii) relabelling nodes and
    elements permutes the matrix symmetrically and nothing else."""
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import PoissonEquation
    from scipy.sparse import csr_matrix
    pb = make_problem("poisson", 5, distortion=0.12)
    pb["code"] = FiniteElementCode("Quad2dC2", PoissonEquation(), name="laplace")
    x = pb["mesh"].node_pos
    pb["vals"][0][:, 0] = 0.5 - 1.5 * x[:, 0] + 2.5 * x[:, 1]
    op = make_oracle(pb)
    r, mats = op.assemble(flag=1)
    lat = pb["mesh"].node_lattice
    interior = np.all((lat > 2) & (lat < 2 * np.array(pb["mesh"].N) - 2), axis=1)
    assert np.abs(r[pb["dofmap"].node_eqn[interior, 0]]).max() <= 1e-7
    op.close()
    a = make_problem("ns_param", 4, distortion=0.1)          # (no pin by node number in this problem: pins follow the boundaries)
    b = make_problem("ns_param", 4, distortion=0.1, unstructured=True)
    # same geometry and values, different labels: node i of `a` is node q[i] of `b`
    key = lambda m: m.node_lattice[:, 0].astype(np.int64) * 1000 + m.node_lattice[:, 1]
    q = np.argsort(key(b["mesh"]))[np.argsort(np.argsort(key(a["mesh"])))]
    assert np.array_equal(b["mesh"].node_lattice[q], a["mesh"].node_lattice)
    b["mesh"].node_pos[q] = a["mesh"].node_pos
    for t in range(a["vals"].shape[0]):
        b["vals"][t][q] = a["vals"][t]
    oa, ob = make_oracle(a), make_oracle(b)
    ra, ma = oa.assemble(flag=1)
    rb, mb = ob.assemble(flag=1)
    n = a["dofmap"].n_dof
    ea, eb = a["dofmap"].node_eqn, b["dofmap"].node_eqn[q]
    m = ea >= 0
    assert np.array_equal(m, eb >= 0)
    perm = np.empty(n, dtype=np.int64); perm[ea[m]] = eb[m]          # equation of a -> equation of b
    A, B = csr_to_sorted(n, *ma[0]).tocoo(), csr_to_sorted(n, *mb[0])
    Ap = csr_matrix((A.data, (perm[A.row], perm[A.col])), shape=(n, n))
    assert abs(Ap - B).max() <= 1e-13 * abs(B).max()
    rp = np.empty(n); rp[perm] = ra
    assert np.abs(rp - rb).max() <= 1e-13 * np.abs(rb).max()
    oa.close(); ob.close()


def test_moving_mesh_tensors_match_finite_differences():
    """int_pt_weights_d_coords and d_dx_shape_dcoord (src/elements.cpp:3051-3155) against FD of the shape buffer."""
    pb = make_problem("ale", 3)
    op = make_oracle(pb)
    e, ipt, eps = 4, 5, 1e-6
    w, sh, dx, dX, wd, dd = op.point_shapes(e, ipt, 1)
    pos0 = pb["pos_hist"][0]
    nodes = pb["mesh"].elem_nodes[e]
    for l2 in (0, 4, 7):
        for i2 in range(2):
            out = []
            for sgn in (-1, 1):
                p = pos0.copy(); p[nodes[l2], i2] += sgn * eps
                op.update_values(0, None, p)
                out.append(op.point_shapes(e, ipt, 1))
            op.update_values(0, None, pos0)
            assert abs((out[1][0][0] - out[0][0][0]) / (2 * eps) - wd[i2, l2]) <= 1e-6 * np.abs(wd).max()
            assert np.abs((out[1][2] - out[0][2]) / (2 * eps) - dd[:, :, l2, i2]).max() <= 1e-6 * np.abs(dd).max()
    op.close()


def test_patch_test_and_poisson_structure():
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import PoissonEquation
    pb = make_problem("poisson", 6)
    pb["code"] = FiniteElementCode("Quad2dC2", PoissonEquation(), name="laplace")
    x = pb["mesh"].node_pos
    pb["vals"][0][:, 0] = 1.0 + 2.0 * x[:, 0] - 3.0 * x[:, 1]      # harmonic: interior residual vanishes
    op = make_oracle(pb)
    r, mats = op.assemble(flag=1)
    n = pb["dofmap"].n_dof
    A = csr_to_sorted(n, *mats[0])
    assert abs(A - A.T).max() <= 1e-13 * abs(A).max()
    lat = pb["mesh"].node_lattice
    interior = np.all((lat > 2) & (lat < 2 * np.array(pb["mesh"].N) - 2), axis=1)
    rows = pb["dofmap"].node_eqn[interior, 0]
    # With exact 3x3 Gauss points this residual would vanish to rounding.  The reference's mistyped Gauss<2,3> knots
    # (integral.cc:87-93, relative error 8e-9) make the rule inexact for the quadratic d(psi)/dx: the reference (and
    # therefore the oracle and the CUDA path) leaves a residual of order 1e-9 here.  Both bounds document that.
    assert 1e-11 <= np.abs(r[rows]).max() <= 1e-8
    assert np.abs(np.asarray(A.sum(axis=1)).ravel()[rows]).max() <= 1e-8
    assert np.linalg.eigvalsh(A.toarray()).min() > 0      # SPD on the free dofs
    area = sum(op.point_shapes(0, ipt, 0)[0][0] for ipt in range(9))
    assert abs(area - (1.0 / 6) ** 2) <= 1e-15
    op.close()


def test_vectors_of_pairs_semantics_and_threads():
    """problem.cc:5498-5572: first-touch column order, exact zeros dropped; the element-range split gives the same sums."""
    pb = make_problem("ns", 5)
    op = make_oracle(pb)
    r1, m1 = op.assemble(flag=1, nthreads=1)
    rs, ci, va = m1[0]
    assert np.all(va != 0.0)
    unsorted_rows = sum(1 for i in range(len(rs) - 1) if np.any(np.diff(ci[rs[i]:rs[i + 1]]) < 0))
    assert unsorted_rows > 0
    r4, m4 = op.assemble(flag=1, nthreads=4)
    n = pb["dofmap"].n_dof
    A1, A4 = csr_to_sorted(n, *m1[0]), csr_to_sorted(n, *m4[0])
    assert abs(A1 - A4).max() <= 1e-14 * abs(A1).max() and np.abs(r1 - r4).max() <= 1e-14 * np.abs(r1).max()
    op.close()


def test_bdf_weights_and_degraded_start():
    """src/timestepper.cpp:31-59 and the _degr rule src/elements.cpp:4611-4626: with zero unsteady steps done the BDF2
    default degrades to BDF1."""
    from pyoomph_b200.assembly import bdf_weights
    w1, w2 = bdf_weights(0.01, 0.012)
    assert w1[0] == 1.0 / 0.01 and w1[1] == -1.0 / 0.01
    assert abs(w2[:3].sum()) < 1e-9 and w2[0] == 1.0 / 0.01 + 1.0 / (0.01 + 0.012)
    pb = make_problem("heat3d", 2)
    op = make_oracle(pb)
    op.set_unsteady(0.3, 0.01, 0.012, 0)
    M_deg = op.element(0, flag=2)
    op.set_unsteady(0.3, 0.01, 0.012, 2)
    M_bdf2 = op.element(0, flag=2)
    assert np.array_equal(M_deg[2], M_bdf2[2])
    assert np.abs((M_deg[1] - M_bdf2[1]) - (w1[0] - w2[0]) * M_bdf2[2]).max() <= 1e-12 * np.abs(M_bdf2[1]).max()
    op.close()


@pytest.mark.parametrize("kind,N", [("poisson", 5), ("ns_unsteady", 4), ("heat3d", 2), ("ale", 4), ("ns_axi_swirl", 3), ("ale_axi_obs", 3), ("heat3d_obs", 2)])
def test_golden_vectors(kind, N):
    """Committed fixtures (tests/golden/make_golden.py): checksums of residual/Jacobian/mass matrix of the oracle.
    They were generated by the oracle itself (the reference cannot run here), so they pin regressions, not parity."""
    path = os.path.join(HERE, "golden", "%s_%d.json" % (kind, N))
    gold = json.load(open(path))
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    r, mats = op.assemble(flag=2)
    n = pb["dofmap"].n_dof
    assert n == gold["n_dof"]
    assert abs(np.abs(r).sum() - gold["res_l1"]) <= 1e-11 * gold["res_l1"]
    for m, key in zip(mats, ("jac", "mass")):
        A = csr_to_sorted(n, *m)
        assert A.nnz == gold[key + "_nnz"]
        assert abs(abs(A).sum() - gold[key + "_l1"]) <= 1e-11 * max(gold[key + "_l1"], 1e-300)
        v = np.cos(np.arange(n))
        assert abs(np.abs(A @ v).sum() - gold[key + "_matvec_l1"]) <= 1e-10 * max(gold[key + "_matvec_l1"], 1e-300)
    for k, val in gold.get("integrals", {}).items():
        assert abs(op.evaluate_integral_expressions()[k] - val) <= 1e-12 * max(abs(v_) for v_ in gold["integrals"].values())
    op.close()


@pytest.mark.parametrize("kind,N", [("ns_unsteady", 3), ("nlheat", 3), ("ns_axi_swirl", 3)])
def test_hessian_routine_matches_finite_differences_and_flags(kind, N):
    """the reference's debug_hessian idea (src/elements.cpp:5129): H.Y against finite differences of J and M; flag 0/1/3/4
    are consistent contractions of the same ndof^3 tensor"""
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    n = pb["dofmap"].n_dof
    rng = np.random.default_rng(2)
    Y = rng.uniform(-1, 1, (2, n))
    e = pb["mesh"].n_elem // 2
    P, Cs, eq = op.element_hessian(e, Y, flag=2)
    dm, vals, eps = pb["dofmap"], pb["vals"], 1e-6
    NJ, NM = np.zeros_like(P[0]), np.zeros_like(P[0])
    for k, g in enumerate(eq):
        nn, f = np.argwhere(dm.node_eqn == g)[0]
        out = []
        for sgn in (-1, 1):
            v = vals[0].copy(); v[nn, f] += sgn * eps
            op.update_values(0, v)
            out.append(op.element(e, flag=2))
        op.update_values(0, vals[0])
        NJ[:, k] = ((out[1][1] - out[0][1]) / (2 * eps)) @ Y[0][eq]
        NM[:, k] = ((out[1][2] - out[0][2]) / (2 * eps)) @ Y[0][eq]
    assert np.abs(P[0] - NJ).max() <= 1e-7 * max(np.abs(P[0]).max(), 1e-300)
    assert np.abs(Cs[0] - NM).max() <= 1e-7 * max(np.abs(NM).max(), np.abs(P[0]).max())
    if kind == "nlheat":
        assert np.abs(Cs[0]).max() > 1e-6          # the mass Hessian is really exercised
    # flag 3: raw tensors; flag 1 / 4 are its contractions over the middle / first index; flag 0 contracts both
    nd = len(eq)
    P1 = op.element_hessian(e, Y, flag=1)[0]
    P4 = op.element_hessian(e, Y, flag=4)[0]
    C = rng.uniform(-1, 1, (2, n))
    p0, _ = op.element_hessian(e, Y[0], C=C, flag=0)
    assert np.abs(p0 - np.stack([P1[0] @ C[v][eq] for v in range(2)])).max() <= 1e-12 * np.abs(p0).max()
    if kind == "ns_unsteady":                   # convective Hessian is symmetric in (j,k) but not in (i,j)
        assert np.abs(P1[0] - P4[0]).max() > 1e-8
    op.close()


def test_integral_expressions_analytic_pins():
    """EvalIntegralExpression (src/codegen.cpp:4125-4364) summed like Mesh::evaluate_integral_expression: volumes of the unit
    square / cube / cylinder (2 pi r dx with the truncated Pi of jitbridge_hang.h:355), exact integrals of interpolated
    polynomials, on uniform and on distorted meshes (the integral of 1 does not see interior distortion)."""
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import IntegralObservables, PoissonEquation
    from pyoomph_b200.expressions import var
    PI_T = 3.14159265359
    for kind, N, vol in (("ns_obs", 4, 1.0), ("ns_axi_obs", 4, PI_T), ("heat3d_obs", 2, 1.0)):
        pb = make_problem(kind, N)
        op = make_oracle(pb)
        obs = op.evaluate_integral_expressions()
        assert abs(obs["volume"] - vol) <= 2e-9 * vol              # mistyped Gauss<2,3> knots: exact only to ~1e-9 for r dx
        op.close()
    # interior distortion moves no boundary node here: the area stays 1 to the quadrature's accuracy for the curved interior edges
    pb = make_problem("ns_obs", 5)
    x = pb["mesh"].node_pos
    pb["vals"][0][:, pb["code"].fields["velocity_x"].index] = 1.0 + 2.0 * x[:, 0] - x[:, 1]           # Q9 reproduces it exactly
    pb["vals"][0][:, pb["code"].fields["velocity_y"].index] = x[:, 0] * x[:, 1]
    op = make_oracle(pb)
    obs = op.evaluate_integral_expressions()
    # the mistyped knots (+0.774596662941483 against -0.774596669241483) are not symmetric: even a linear integrand is off by ~1e-10
    assert 1e-12 <= abs(obs["momentum_x"] - 1.5) <= 1e-9 and abs(obs["momentum_y"] - 0.25) <= 1e-9
    assert list(obs) == ["volume", "kinetic_energy", "momentum_x", "momentum_y", "pressure_integral", "dissipation"]
    op.close()
    with pytest.raises(RuntimeError):
        from pyoomph_b200.expressions import testfunction
        FiniteElementCode("Quad2dC2", PoissonEquation() + IntegralObservables(bad=lambda: testfunction("u")), name="bad").integral_form()


def _cavity(N):
    """lid-driven cavity of BASELINE config 2 at Re = 100: u = 1 on the lid, no slip elsewhere, pressure pinned at node 0"""
    pb = make_problem("ns", N)
    mesh, code = pb["mesh"], pb["code"]
    vals = np.zeros_like(pb["vals"])
    vals[0][mesh.boundaries["top"], code.fields["velocity_x"].index] = 1.0
    pb["vals"] = vals
    return pb


def newton_cavity(pb, assemble, set_dofs, max_iter=12):
    """Problem.solve()'s Newton loop (oomph Problem::newton_solve): U <- U - J^-1 R with SuperLU, until max|R| < 1e-10"""
    from scipy.sparse.linalg import spsolve
    eq = pb["dofmap"].node_eqn
    m = eq >= 0
    U = np.zeros(pb["dofmap"].n_dof)
    U[eq[m]] = pb["vals"][0][m]
    history = []
    for _ in range(max_iter):
        set_dofs(U)
        r, A = assemble()
        history.append(float(np.abs(r).max()))
        if history[-1] < 1e-10:
            break
        U = U - spsolve(A.tocsc(), r)
    return U, history


def test_newton_solve_of_the_cavity_converges_quadratically():
    """end-to-end pin of residual AND Jacobian (sign conventions, consistency): Newton's method on the lid-driven cavity converges
    quadratically only with the exact Jacobian; the converged flow has the primary vortex (negative u on the lower half of the
    vertical centre line) and is divergence free in the weak sense (continuity rows of the residual vanish)."""
    N = 8
    pb = _cavity(N)
    op = make_oracle(pb)
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn
    m = eq >= 0

    def set_dofs(U):
        v = pb["vals"][0].copy()
        v[m] = U[eq[m]]
        op.update_values(0, v)

    def assemble():
        r, mats = op.assemble(flag=1)
        return r, csr_to_sorted(n, *mats[0])
    U, hist = newton_cavity(pb, assemble, set_dofs)
    assert hist[-1] < 1e-10 and len(hist) <= 9
    tail = [h for h in hist if h < 1e-2]
    assert len(tail) >= 2 and tail[1] <= 50 * tail[0] ** 2 + 1e-12       # quadratic once close
    ux = pb["vals"][0][:, pb["code"].fields["velocity_x"].index].copy()
    mx = m[:, pb["code"].fields["velocity_x"].index]
    ux[mx] = U[eq[mx, pb["code"].fields["velocity_x"].index]]
    lat = pb["mesh"].node_lattice
    centre_low = (lat[:, 0] == N) & (lat[:, 1] > 0) & (lat[:, 1] < N)
    assert ux[centre_low].min() < -0.05
    op.close()


@pytest.mark.parametrize("kind", ["ns_unsteady", "nlheat"])
def test_hessian_flags_of_the_oracle_are_consistent(kind):
    """SURVEY A.5 / src/jitbridge.h:624-691: flag 3 hands out the raw H_ijk (and the mass Hessian); flags 1/2 contract the MIDDLE index
    (sum_j H_ijk Y_j), flags 4/5 the FIRST (sum_j H_jik Y_j); flag 0 = Y_j H_ijk C_k.  All from one element of the oracle."""
    from problems import make_oracle, make_problem
    pb = make_problem(kind, 3)
    op = make_oracle(pb)
    rng = np.random.default_rng(2)
    n = pb["dofmap"].n_dof
    Y = rng.uniform(-1, 1, (2, n))
    e = pb["mesh"].n_elem // 2
    H, MH, eq = op.element_hessian(e, Y[:1], flag=3)
    Yl = Y[:, eq]
    P2, C2, _ = op.element_hessian(e, Y, flag=2)
    P5, C5, _ = op.element_hessian(e, Y, flag=5)
    for v in range(2):
        assert np.abs(P2[v] - np.einsum("ijk,j->ik", H, Yl[v])).max() <= 1e-13 * np.abs(P2).max()
        assert np.abs(C2[v] - np.einsum("ijk,j->ik", MH, Yl[v])).max() <= 1e-13 * max(np.abs(C2).max(), 1e-300)
        assert np.abs(P5[v] - np.einsum("jik,j->ik", H, Yl[v])).max() <= 1e-13 * np.abs(P5).max()
        assert np.abs(C5[v] - np.einsum("jik,j->ik", MH, Yl[v])).max() <= 1e-13 * max(np.abs(C5).max(), 1e-300)
    assert np.abs(H - H.transpose(0, 2, 1)).max() <= 1e-13 * np.abs(H).max()          # a Hessian: symmetric in (j, k)
    C = rng.uniform(-1, 1, (3, n))
    P0, _ = op.element_hessian(e, Y[:1], C, flag=0)
    assert np.abs(P0 - np.einsum("j,ijk,vk->vi", Yl[0], H, C[:, eq])).max() <= 1e-13 * np.abs(P0).max()
    op.close()


def test_point_expressions_of_the_oracle_linear_fields():
    """EvalLocalExpression / GetZ2Fluxes of the oracle on a distorted mesh: for linear velocity fields the vorticity, the strain
    components and the Z2 fluxes (the velocity gradient) are the analytic constants at every node and integration point; the
    extremum expression of the pressure equals the interpolated pressure."""
    from problems import make_oracle, make_problem
    pb = make_problem("ns_pts", 3, distortion=0.12)
    m = pb["mesh"]
    pb["vals"][0][:, 0] = 2 + 3 * m.node_pos[:, 0] - m.node_pos[:, 1]
    pb["vals"][0][:, 1] = -1 + 0.5 * m.node_pos[:, 0] + 4 * m.node_pos[:, 1]
    op = make_oracle(pb)
    names = [n for k, n in pb["code"].point_expression_names()]
    for pts in ("nodes", "gauss"):
        v = op.eval_point_expressions(pts)
        for nm, expect in (("vorticity", 1.5), ("strain_xx", 3.0), ("strain_xy", -1.0), ("strain_yx", 0.5), ("strain_yy", 4.0)):
            assert np.abs(v[:, :, names.index(nm)] - expect).max() <= 1e-11
        z = v[:, :, [i for i, n in enumerate(names) if n.startswith("flux_")]]
        assert np.abs(z - np.array([3.0, -1.0, 0.5, 4.0])).max() <= 1e-11
    vn = op.eval_point_expressions("nodes")
    ip = names.index("p")
    en = m.elem_nodes
    vert = [0, 2, 6, 8]
    assert np.abs(vn[:, vert, ip] - pb["vals"][0][en[:, vert], 2]).max() <= 1e-13       # C1 field at the vertex nodes: the nodal values
    op.close()


def test_triangle_elements_of_the_oracle():
    """BulkElementTri2dC2 in the oracle: symmetric Laplace matrix with zero row sums, patch test on a distorted triangle mesh, the
    element areas add up to the domain, analytic-vs-FD Jacobian of the Taylor-Hood P2/P1 Navier-Stokes class (src/elements.cpp:5880)."""
    from problems import csr_to_sorted, make_oracle, make_problem
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import PoissonEquation
    pb = make_problem("poisson_tri", 4, distortion=0.1)
    m = pb["mesh"]
    pb["code"] = FiniteElementCode("Tri2dC2", PoissonEquation(), name="laplacetri")
    pb["vals"][0][:, 0] = 1 + 2 * m.node_pos[:, 0] - 3 * m.node_pos[:, 1]
    op = make_oracle(pb)
    r, mats = op.assemble(flag=1)
    n = pb["dofmap"].n_dof
    A = csr_to_sorted(n, *mats[0])
    assert abs(A - A.T).max() <= 1e-13 * abs(A).max()
    lat = m.node_lattice
    interior = np.all((lat > 0) & (lat < 2 * np.array(m.N)), axis=1)
    rows = pb["dofmap"].node_eqn[interior, 0]
    rows = rows[rows >= 0]
    assert np.abs(r[rows]).max() <= 1e-12 * abs(A).max()              # a linear field is reproduced exactly: zero interior residual
    flat = make_problem("poisson_tri", 3)
    op2 = make_oracle(flat)
    area = sum(op2.point_shapes(e, ipt, flag=0)[0][0] for e in range(flat["mesh"].n_elem) for ipt in range(7))
    assert abs(area - 1.0) <= 1e-12
    op.close(); op2.close()
    pb = make_problem("ns_tri", 2, distortion=0.1)
    op = make_oracle(pb)
    _, mats = op.assemble(flag=1)
    n = pb["dofmap"].n_dof
    A = csr_to_sorted(n, *mats[0]).toarray()
    eq, eps = pb["dofmap"].node_eqn, 1e-6
    for node, f in ((5, 0), (7, 1), (0, 2), (12, 0)):
        g = eq[node, f]
        if g < 0:
            continue
        v = pb["vals"][0].copy()
        v[node, f] += eps
        op.update_values(0, v)
        rp, _ = op.assemble(flag=0)
        v[node, f] -= 2 * eps
        op.update_values(0, v)
        rm, _ = op.assemble(flag=0)
        op.update_values(0, pb["vals"][0])
        assert np.abs((rp - rm) / (2 * eps) - A[:, g]).max() <= 1e-7 * np.abs(A[:, g]).max()
    op.close()


def test_interface_elements_of_the_oracle():
    """InterfaceElementLine1dC2 in the oracle (el_dim 1 in nodal dimension 2: src/elements.cpp:3651-3672, normal :1730-1752): on a
    distorted Q9 mesh the edges' surface measure adds up to the length of the boundary polyline of parabolas, the unit normal is
    orthogonal to the tangent and LEAVES the bulk element the edge belongs to (what FaceElement::normal_sign achieves in the reference),
    the surface gradient of a linear field is its tangential projection, and the analytic Jacobians of the Robin and free-surface
    classes agree with finite differences of the residual."""
    from problems import csr_to_sorted, make_oracle, make_problem
    from pyoomph_b200.cuda_emitter import gauss_rule_1d
    pb = make_problem("freesurf_if", 4, distortion=0.15)
    im, bulk = pb["mesh"], pb["bulk_mesh"]
    a, b = np.array([0.7, -1.3]), 0.4
    pb["vals"][0][:, 0] = b + bulk.node_pos @ a                                        # velocity_x linear in (x, y)
    op = make_oracle(pb)
    kn, _ = gauss_rule_1d()
    for e in range(im.n_elem):
        xe = im.node_pos[im.elem_nodes[e]]
        centroid = bulk.node_pos[bulk.elem_nodes[im.bulk_element[e]]].mean(axis=0)
        for ipt in range(3):
            wts, sh, dx, dX, _, _ = op.point_shapes(e, ipt, flag=0)
            s = kn[ipt][0]
            t = np.array([s - 0.5, -2.0 * s, s + 0.5]) @ xe
            n = np.array([-t[1], t[0]]) / np.hypot(*t)
            assert abs(n @ t) <= 1e-14 and abs(np.hypot(*n) - 1.0) <= 1e-14
            assert n @ (sh @ xe - centroid) > 0.0                                      # outward
            assert abs(wts[0] - np.hypot(*t) * (5.0 / 9.0 if ipt != 1 else 8.0 / 9.0)) <= 1e-14
            grad_s = pb["vals"][0][im.elem_nodes[e], 0] @ dx                           # surface gradient of the linear field
            tau = t / np.hypot(*t)
            assert np.abs(grad_s - (a @ tau) * tau).max() <= 1e-12
    op.close()
    for kind in ("robin_if", "freesurf_if"):
        pb = make_problem(kind, 3, distortion=0.12)
        op = make_oracle(pb)
        _, mats = op.assemble(flag=1)
        n = pb["dofmap"].n_dof
        A = csr_to_sorted(n, *mats[0]).toarray()
        eq, eps = pb["dofmap"].node_eqn, 1e-6
        for node in np.unique(pb["mesh"].elem_nodes)[:6]:
            for f in range(eq.shape[1]):
                g = eq[node, f]
                if g < 0:
                    continue
                v = pb["vals"][0].copy()
                v[node, f] += eps
                op.update_values(0, v)
                rp, _ = op.assemble(flag=0)
                v[node, f] -= 2 * eps
                op.update_values(0, v)
                rm, _ = op.assemble(flag=0)
                op.update_values(0, pb["vals"][0])
                assert np.abs((rp - rm) / (2 * eps) - A[:, g]).max() <= 1e-8 * max(np.abs(A).max(), 1e-300)
        op.close()


def test_moving_free_surface_of_the_oracle_matches_finite_differences():
    """NavierStokesFreeSurface on a moving mesh (config 4's interface class): Jacobian columns of velocities, the multiplier AND the
    nodal positions -- derivative of the unit normal (d_normal_dcoord, src/elements.cpp:1461-1490), of the line measure and of the surface
    gradients of the test functions (the el_dim x nodal_dim tensors of src/elements.cpp:3051-3155) -- against central differences of the
    residual, BDF2 step with mesh velocity"""
    from problems import csr_to_sorted, make_oracle, make_problem
    _check_moving_free_surface("freesurf_mov_if")
    _check_moving_free_surface("freesurf_mov_axi_if")         # config 4 is axisymmetric: 2 pi r in the measure, v_r / r in the surface divergence


def _check_moving_free_surface(kind):
    from problems import csr_to_sorted, make_oracle, make_problem
    pb = make_problem(kind, 4, distortion=0.12)
    assert pb["code"].coordinates_as_dofs and pb["unsteady"]
    op = make_oracle(pb)
    _, mats = op.assemble(flag=1)
    n = pb["dofmap"].n_dof
    A = csr_to_sorted(n, *mats[0]).toarray()
    scale = np.abs(A).max()
    eq, peq, eps = pb["dofmap"].node_eqn, pb["dofmap"].pos_eqn, 1e-6
    nodes = np.unique(pb["mesh"].elem_nodes)
    checked_pos = 0
    for node in nodes[:7]:
        for f in range(eq.shape[1]):
            g = eq[node, f]
            if g < 0:
                continue
            v = pb["vals"][0].copy()
            v[node, f] += eps
            op.update_values(0, v)
            rp, _ = op.assemble(flag=0)
            v[node, f] -= 2 * eps
            op.update_values(0, v)
            rm, _ = op.assemble(flag=0)
            op.update_values(0, pb["vals"][0])
            assert np.abs((rp - rm) / (2 * eps) - A[:, g]).max() <= 1e-8 * scale, ("value", node, f)
        for d in range(2):
            g = peq[node, d]
            if g < 0:
                continue
            x = pb["pos_hist"][0].copy()
            x[node, d] += eps
            op.update_values(0, None, x)
            rp, _ = op.assemble(flag=0)
            x[node, d] -= 2 * eps
            op.update_values(0, None, x)
            rm, _ = op.assemble(flag=0)
            op.update_values(0, None, pb["pos_hist"][0])
            assert np.abs((rp - rm) / (2 * eps) - A[:, g]).max() <= 2e-8 * scale, ("position", node, d)
            checked_pos += 1
    assert checked_pos >= 8
    op.close()


def test_bulk_face_elements_of_the_oracle():
    """Faces seen through their bulk elements (QuadFace2dC2: the reference's bulk_eleminfo access of an interface element): on a distorted
    Q9 mesh the bulk gradient of a linear field is exact on the face, only the three face nodes carry values there, the measure adds up
    to the boundary length, the normal is unit, orthogonal to the face tangent and leaves the bulk element; the analytic Jacobian of
    Nitsche's method (normal derivatives of field and test function, nonlinear conductivity) agrees with finite differences."""
    from problems import csr_to_sorted, make_oracle, make_problem
    from pyoomph_b200.cuda_emitter import gauss_rule_1d
    pb = make_problem("nitsche_face", 4, distortion=0.12)
    im, bulk = pb["mesh"], pb["bulk_mesh"]
    op = make_oracle(pb)
    a = np.array([0.7, -1.3])
    kn, w1 = gauss_rule_1d()
    length = 0.0
    for e in range(im.n_elem):
        xe = bulk.node_pos[im.elem_nodes[e]]
        centroid = xe.mean(axis=0)
        for ipt in range(3):
            wts, sh, dx, _, _, _ = op.point_shapes(e, ipt, flag=0)
            assert np.abs((xe @ a) @ dx - a).max() <= 1e-12                      # bulk gradient, both components
            assert np.abs(sh[3:]).max() <= 1e-15 and abs(sh[:3].sum() - 1.0) <= 1e-15
            s = kn[ipt][0]
            t = np.array([s - 0.5, -2.0 * s, s + 0.5]) @ xe[:3]                   # tangent from the three face nodes
            assert abs(wts[0] - np.hypot(*t) * w1[ipt]) <= 1e-14
            length += wts[0]
            nrm = np.array([t[1], -t[0]]) / np.hypot(*t)
            assert nrm @ (sh @ xe - centroid) > 0.0                               # outward
    # boundary length of the three sides: sum of the parabola arcs ~ their chord polygon (loose), and equal to the line-element measure
    pbl = make_problem("robin_if", 4, distortion=0.12)
    assert length > 2.5
    n = pb["dofmap"].n_dof
    r, mats = op.assemble(flag=1)
    A = csr_to_sorted(n, *mats[0]).toarray()
    eq, eps = pb["dofmap"].node_eqn, 1e-6
    for node in np.unique(im.elem_nodes)[::2]:
        g = eq[node, 0]
        if g < 0:
            continue
        v = pb["vals"][0].copy()
        v[node, 0] += eps
        op.update_values(0, v)
        rp, _ = op.assemble(flag=0)
        v[node, 0] -= 2 * eps
        op.update_values(0, v)
        rm, _ = op.assemble(flag=0)
        op.update_values(0, pb["vals"][0])
        assert np.abs((rp - rm) / (2 * eps) - A[:, g]).max() <= 1e-8 * np.abs(A).max()
    # the Jacobian couples the face to the interior nodes of its bulk element (normal derivative): more than the 3x3 face block
    assert np.count_nonzero(A) > 9 * im.n_elem
    op.close()


def test_element_sizes_of_the_oracle():
    """var("element_length_h") / "cartesian_element_size_Eulerian" (fill_shape_info_element_sizes, src/elements.cpp:3527-3568): on a uniform
    N x N mesh the element length is 1/N, so the streamline-upwind residual equals the one written with that constant; on a distorted
    mesh (axisymmetric: Cartesian size, the measure's 2 pi r not included) the analytic Jacobian agrees with finite differences."""
    from problems import csr_to_sorted, make_oracle, make_problem
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import StreamlineDiffusionAdvection
    import pyoomph_b200.expressions as ex_
    N = 5
    pb = make_problem("supg", N)

    class _ConstH(StreamlineDiffusionAdvection):
        def define_residuals(self):
            real_var = ex_.var
            try:
                ex_.var = lambda a: (1.0 / N if a == "element_length_h" else real_var(a))
                import pyoomph_b200.equations as eqm
                saved = eqm.var
                eqm.var = ex_.var
                super().define_residuals()
            finally:
                ex_.var = real_var
                eqm.var = saved
    pbc = dict(pb, code=FiniteElementCode("Quad2dC2", _ConstH(), name="supgconst"))
    a, b = make_oracle(pb), make_oracle(pbc)
    ra, ma = a.assemble(flag=2)
    rb, mb = b.assemble(flag=2)
    assert np.abs(ra - rb).max() <= 1e-13 * np.abs(rb).max()
    n = pb["dofmap"].n_dof
    assert abs(csr_to_sorted(n, *ma[0]) - csr_to_sorted(n, *mb[0])).max() <= 1e-13 * abs(csr_to_sorted(n, *mb[0])).max()
    a.close(); b.close()
    # three dimensions: h = (element volume)^(1/3) = 1/N on uniform bricks
    pb3 = make_problem("supg3d", 3)

    class _ConstH3(_ConstH):
        pass
    N = 3           # read by _ConstH when the class is built
    pbc3 = dict(pb3, code=FiniteElementCode("Brick3dC2", _ConstH3(wind=(1.0, 0.5, -0.25)), name="supg3dconst"))
    a, b = make_oracle(pb3), make_oracle(pbc3)
    ra, _ = a.assemble(flag=0)
    rb, _ = b.assemble(flag=0)
    assert np.abs(ra - rb).max() <= 1e-13 * np.abs(rb).max()
    a.close(); b.close()
    for kind in ("supg", "supg_axi", "supg3d", "supg_tet", "supg_tri"):
        pb = make_problem(kind, 2 if kind == "supg_tet" else (3 if kind == "supg3d" else 4), distortion=0.12)
        op = make_oracle(pb)
        n = pb["dofmap"].n_dof
        _, mats = op.assemble(flag=1)
        A = csr_to_sorted(n, *mats[0]).toarray()
        eq, eps = pb["dofmap"].node_eqn, 1e-6
        for node in range(0, pb["mesh"].n_node, 4):
            g = eq[node, 0]
            if g < 0:
                continue
            v = pb["vals"][0].copy()
            v[node, 0] += eps
            op.update_values(0, v)
            rp, _ = op.assemble(flag=0)
            v[node, 0] -= 2 * eps
            op.update_values(0, v)
            rm, _ = op.assemble(flag=0)
            op.update_values(0, pb["vals"][0])
            assert np.abs((rp - rm) / (2 * eps) - A[:, g]).max() <= 1e-8 * np.abs(A).max()
        op.close()


def test_lagrangian_element_sizes_of_the_oracle():
    """var("element_size_Lagrangian") / "cartesian_element_size_Lagrangian" (src/elements.cpp:3546-3551, :3569-3574; jitbridge.h:179) on a
    MOVING mesh: on a uniform N x N Lagrangian mesh the Cartesian size is 1/N^2 whatever the current positions are, so the residual
    equals the one written with h = 1/N; the analytic Jacobian (value and position columns) agrees with finite differences on a
    distorted mesh, Cartesian and axisymmetric (2 pi R at the Lagrangian position)."""
    from problems import csr_to_sorted, make_oracle, make_problem
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import PseudoElasticMesh, StreamlineDiffusionAdvection
    import pyoomph_b200.expressions as ex_
    import pyoomph_b200.equations as eqm
    N = 4
    pb = make_problem("supg_ale", N)

    class _ConstH(StreamlineDiffusionAdvection):
        def define_residuals(self):
            saved = eqm.var
            eqm.var = lambda a: (1.0 / N ** 2 if a == "cartesian_element_size_Lagrangian" else saved(a))
            try:
                super().define_residuals()
            finally:
                eqm.var = saved
    pbc = dict(pb, code=FiniteElementCode("Quad2dC2", _ConstH(lagrangian_size=True, cartesian_size=True) + PseudoElasticMesh(), name="supgaleconst"))
    a, b = make_oracle(pb), make_oracle(pbc)
    ra, ma = a.assemble(flag=1)
    rb, mb = b.assemble(flag=1)
    assert np.abs(ra - rb).max() <= 1e-13 * np.abs(rb).max()
    n = pb["dofmap"].n_dof
    assert abs(csr_to_sorted(n, *ma[0]) - csr_to_sorted(n, *mb[0])).max() <= 1e-13 * abs(csr_to_sorted(n, *mb[0])).max()
    a.close(); b.close()
    for kind in ("supg_ale", "supg_ale_axi"):
        pb = make_problem(kind, 3, distortion=0.12)
        op = make_oracle(pb)
        n = pb["dofmap"].n_dof
        _, mats = op.assemble(flag=1)
        A = csr_to_sorted(n, *mats[0]).toarray()
        eps = 1e-6
        for node in range(0, pb["mesh"].n_node, 5):
            for (arr, eqs, key) in ((pb["vals"][0], pb["dofmap"].node_eqn, "node_val"), (pb["pos_hist"][0], pb["dofmap"].pos_eqn, "node_pos")):
                g = eqs[node, 0]
                if g < 0:
                    continue
                res = []
                for sgn in (+1, -1):
                    v = arr.copy()
                    v[node, 0] += sgn * eps
                    op.update_values(0, **{key: v})
                    res.append(op.assemble(flag=0)[0])
                op.update_values(0, **{key: arr})
                assert np.abs((res[0] - res[1]) / (2 * eps) - A[:, g]).max() <= 2e-7 * np.abs(A).max(), (kind, key, node)
        op.close()


def test_integral_gradient_contributions_match_finite_differences_of_the_integrals():
    """add_integral_function(..., with_gradient=True): the residual vector of "d_integral_<name>" is d(integral)/dU (checked against
    central differences of EvalIntegralExpression), a linear functional has a vanishing second derivative, a quadratic one a symmetric one"""
    from problems import csr_to_sorted, make_oracle, make_problem
    pb = make_problem("ns_constraint", 4, distortion=0.1)
    code = pb["code"]
    assert code.residual_names() == [""] + ["d_integral_" + k for k in code.integral_expression_names()]
    op = make_oracle(pb)
    n, eq = pb["dofmap"].n_dof, pb["dofmap"].node_eqn
    for iname in code.integral_expression_names():
        which = code.residual_names().index("d_integral_" + iname)
        c, mats = op.assemble(which=which, flag=1)
        H = csr_to_sorted(n, *mats[0])
        if iname == "pressure_integral":
            assert H.nnz == 0 or abs(H).max() == 0.0
        else:
            assert abs(H - H.T).max() <= 1e-13 * abs(H).max()
        eps = 1e-6
        for node in range(0, pb["mesh"].n_node, 7):
            for f in range(eq.shape[1]):
                g = eq[node, f]
                if g < 0:
                    continue
                v = pb["vals"][0].copy()
                v[node, f] += eps
                op.update_values(0, v)
                ip = op.evaluate_integral_expressions()[iname]
                v[node, f] -= 2 * eps
                op.update_values(0, v)
                im = op.evaluate_integral_expressions()[iname]
                op.update_values(0, pb["vals"][0])
                assert abs((ip - im) / (2 * eps) - c[g]) <= 1e-8 * np.abs(c).max()
    op.close()
