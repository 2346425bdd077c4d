"""GPU parity tests proper: CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Tolerance: fp64, 1e-12 relative (BASELINE north_star), measured against the row scale of the reference matrix
(entries that are sums with cancellation cannot be compared entry-relative); CSR pattern: the reference's
value-dependent pattern must be contained in the fixed structural pattern, and after sorting columns the
row_ptr/col_idx of both agree wherever the reference keeps the entry.
"""
import numpy as np
import pytest

from problems import assert_csr_parity, compare_matrix, csr_to_sorted, make_gpu, make_oracle, make_problem

TOL = 1e-12


def _record(kind, N, distortion, unstructured, st):
    """statistics of the bit-exact CSR comparison, printed (driver log) and appended to gpurun_out/ when that directory exists"""
    import json
    import os
    line = json.dumps(dict(kind=kind, N=N, distortion=distortion, unstructured=unstructured, **st))
    print("csr parity:", line)
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "csr_parity_stats.jsonl"), "a") as f:
            f.write(line + "\n")

CASES = [("poisson", 64), ("poisson", 5), ("ns", 12), ("ns_unsteady", 9), ("heat3d", 3), ("ale", 7), ("ns_param", 6),
         ("ns_axi", 9), ("ns_axi_swirl", 7), ("ale_axi", 6)]      # axisymmetric classes of configs 4 and 5


# (kind, N, distortion in element widths, unstructured = random element order + random node labels, no patch hint)
# six-node triangles (BulkElementTri2dC2 = TElement<2,3>, TGauss<2,3>): Poisson, Taylor-Hood P2/P1 NS, NS on a pseudo-elastic moving mesh
TRIANGLES = [("poisson_tri", 9, 0.0, False), ("poisson_tri", 8, 0.12, False), ("ns_tri", 7, 0.1, False), ("ale_tri", 5, 0.08, False)]

# ten-node tetrahedra (BulkElementTetra3dC2 = TElement<3,3>, TGauss<3,3> with its negative weight): Poisson, transient heat, 3D P2/P1 NS
TETRAHEDRA = [("poisson_tet", 3, 0.0, False), ("poisson_tet", 3, 0.1, False), ("heat3d_tet", 3, 0.1, False), ("ns_tet", 2, 0.1, False)]

# element sizes (one number per element from all of its integration points) in a streamline-upwind term; axisymmetric: the Cartesian size
ELEMENT_SIZES = [("supg", 9, 0.0, False), ("supg", 8, 0.12, False), ("supg_axi", 7, 0.1, False), ("supg3d", 3, 0.1, False), ("supg_tet", 2, 0.1, False), ("supg_tri", 6, 0.1, False),
                 ("supg_ale", 6, 0.1, False), ("supg_ale_axi", 5, 0.1, False)]   # Lagrangian sizes on moving meshes

VARIANTS = [("ns", 11, 0.12, False), ("ns_unsteady", 10, 0.1, True), ("heat3d", 3, 0.1, True), ("ale", 6, 0.08, True), ("poisson", 33, 0.15, True),
            ("ns_axi_swirl", 6, 0.1, True), ("ale_axi", 6, 0.08, True)]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,distortion,unstructured", [(k, n, 0.0, False) for k, n in CASES] + VARIANTS + TRIANGLES + TETRAHEDRA + ELEMENT_SIZES)
def test_residual_jacobian_mass_parity(kind, N, distortion, unstructured):
    pb = make_problem(kind, N, distortion=distortion, unstructured=unstructured)
    op = make_oracle(pb)
    asm = make_gpu(pb)
    n = pb["dofmap"].n_dof
    # flag 2: residual + Jacobian + mass matrix
    r_ref, mats = op.assemble(flag=2)
    asm.assemble(flag=2)
    r, jac, mass = asm.fetch(True, True)
    scale = np.abs(r_ref).max()
    assert np.abs(r - r_ref).max() <= TOL * scale
    for vals, (rs, ci, va) in zip((jac, mass), mats):
        A = csr_to_sorted(n, asm.indptr, asm.indices, vals)
        B = csr_to_sorted(n, rs, ci, va)
        if B.nnz == 0:
            assert np.abs(vals).max() == 0.0
            continue
        err, missing = compare_matrix(A, B)
        assert missing == 0
        assert err <= TOL, (kind, err)
        # the north_star bar: row_ptr / col_idx bit-exact after the reference's zero-drop rule, values 1e-12 relative to the ENTRY
        st = assert_csr_parity(asm.indptr, asm.indices, vals, (rs, ci, va), TOL, label="%s N=%d" % (kind, N),
                               # distorted meshes: no symmetric cancellations (measured <= 0.2 % on quads / bricks, 1.2 % on the
                               # P2/P1 triangles); uniform meshes: analytic zeros of the reference element come out as noise on both
                               # sides (6-15 % on squares, 45 % of the P2 stiffness entries on right triangles)
                               max_cancel_fraction=(0.02 if distortion > 0.0 else (0.60 if kind.endswith(("_tri", "_tet")) else 0.30)))
        _record(kind, N, distortion, unstructured, st)
    # flag 0 and flag 1 launches give the same numbers as the flag 2 launch (separate kernels)
    asm.assemble(flag=0)
    r0, _, _ = asm.fetch(False, False)
    assert np.abs(r0 - r).max() <= 1e-13 * scale      # separate kernels: different CSE, rounding-level differences only
    asm.assemble(flag=1)
    r1, j1, _ = asm.fetch(True, False)
    assert np.abs(r1 - r).max() <= 1e-13 * scale and np.abs(j1 - jac).max() <= 1e-13 * np.abs(jac).max()
    op.close()
    asm.close()


@pytest.mark.gpu
def test_parameter_derivative_parity():
    pb = make_problem("ns_param", 6)
    op = make_oracle(pb)
    asm = make_gpu(pb)
    n = pb["dofmap"].n_dof
    r_ref, mats = op.assemble(param=0, flag=1)
    asm.assemble(flag=1, parameter="mu")
    r, jac, _ = asm.fetch(True, False)
    assert np.abs(r - r_ref).max() <= TOL * np.abs(r_ref).max()
    err, missing = compare_matrix(csr_to_sorted(n, asm.indptr, asm.indices, jac), csr_to_sorted(n, *mats[0]))
    assert missing == 0 and err <= TOL
    _record("ns_param_dp", 6, 0.0, False, assert_csr_parity(asm.indptr, asm.indices, jac, mats[0], TOL, label="dJ/dmu"))


@pytest.mark.gpu
def test_deterministic_and_set_dofs_roundtrip():
    """Colouring instead of atomics: two assemblies are bit-identical; set_dofs(host vector) == set_nodal_values."""
    pb = make_problem("ns", 16)
    asm = make_gpu(pb)
    asm.assemble(flag=1)
    r1, j1, _ = asm.fetch()
    asm.assemble(flag=1)
    r2, j2, _ = asm.fetch()
    assert np.array_equal(r1, r2) and np.array_equal(j1, j2)
    eq = pb["dofmap"].node_eqn
    dofs = np.zeros(pb["dofmap"].n_dof)
    dofs[eq[eq >= 0]] = pb["vals"][0][eq >= 0]
    r3, j3, _ = asm.assemble_host(dofs, 1)
    assert np.array_equal(r1, r3) and np.array_equal(j1, j3)
    res, J = asm.get_residuals_and_jacobian(True)
    assert J.indptr.dtype == np.int32 and J.indices.dtype == np.int32 and J.data.dtype == np.float64
    assert np.array_equal(res, r1)
    asm.close()


@pytest.mark.gpu
def test_size_independent_properties_large():
    """BASELINE-size-style checks without the oracle: Poisson Jacobian is symmetric, rows of interior nodes sum to
    zero (constants are in the kernel of the Laplacian), and the residual of a linear field vanishes inside."""
    pb = make_problem("poisson", 192)
    mesh, dm = pb["mesh"], pb["dofmap"]
    pb["vals"][0][:, 0] = 2.0 + 3.0 * mesh.node_pos[:, 0] - 1.5 * mesh.node_pos[:, 1]
    from pyoomph_b200.codegen import FiniteElementCode
    from pyoomph_b200.equations import PoissonEquation
    pb["code"] = FiniteElementCode("Quad2dC2", PoissonEquation(), name="laplace")
    asm = make_gpu(pb)
    asm.assemble(flag=1)
    r, jac, _ = asm.fetch()
    from scipy.sparse import csr_matrix
    A = csr_matrix((jac, asm.indices, asm.indptr), shape=(dm.n_dof, dm.n_dof))
    scale = abs(A).max()
    assert abs(A - A.T).max() <= 1e-12 * scale
    lat = mesh.node_lattice
    interior = np.all((lat > 2) & (lat < 2 * np.array(mesh.N) - 2), axis=1)
    rows = dm.node_eqn[interior, 0]
    assert np.abs(np.asarray(A.sum(axis=1)).ravel()[rows]).max() <= 1e-11 * scale
    assert np.abs(r[rows]).max() <= 1e-10 * scale * np.abs(pb["vals"][0]).max()
    asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N", [("ns_unsteady", 8), ("nlheat", 7), ("ns", 6), ("ns_axi_swirl", 5)])
def test_hessian_vector_products_parity(kind, N):
    """HessianVectorProduct<i>: d(J.Y)/dU, d(M.Y)/dU (flags 1,2) and sum_jk Y_j H_ijk C_k (flag 0) against the oracle's
    ndof^3-buffer routine (which is itself checked against finite differences of the Jacobian in tests/test_oracle.py)."""
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    asm = make_gpu(pb)
    n = pb["dofmap"].n_dof
    rng = np.random.default_rng(1)
    Y = rng.uniform(-1, 1, (2, n))
    Jr, Mr = op.assemble_hessian(Y, flag=2)
    Jg, Mg = asm.assemble_hessian(Y, flag=2)
    from scipy.sparse import csr_matrix
    for ref, got in list(zip(Jr, Jg)) + list(zip(Mr, Mg)):
        A = csr_matrix((got, asm.indices, asm.indptr), shape=(n, n))
        scale = max(abs(ref).max(), 1e-300)
        assert abs(A - ref).max() <= 1e-12 * scale if ref.nnz else np.abs(got).max() == 0.0
    C = rng.uniform(-1, 1, (3, n))
    prod = asm.hessian_vector_products(Y[0], C)
    ref = np.stack([Jr[0] @ c for c in C])
    assert np.abs(prod - ref).max() <= 1e-12 * np.abs(ref).max()
    op.close(); asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,distortion,unstructured", [("ns_obs", 9, 0.0, False), ("ns_axi_obs", 8, 0.1, True), ("ale_axi_obs", 6, 0.08, False),
                                                          ("heat3d_obs", 3, 0.1, True), ("ns_obs", 100, 0.1, False),
                                                          ("robin_if_obs", 6, 0.1, False)])          # integrals over an interface (lines in 2D)
def test_integral_expressions_parity(kind, N, distortion, unstructured):
    """EvalIntegralExpression over the whole mesh (Mesh::evaluate_integral_expression): GPU kernel + fixed-order reduction against
    the oracle's element loop, 1e-12 relative to the integral of |integrand| scale (here: to the largest observable), twice
    (bit-reproducible), and next to a residual assembly on the same problem object."""
    pb = make_problem(kind, N, distortion=distortion, unstructured=unstructured)
    op = make_oracle(pb)
    asm = make_gpu(pb)
    ref = op.evaluate_integral_expressions()
    got = asm.evaluate_integral_expressions()
    assert asm.launch_count() == 2 and list(got) == list(ref) == asm.integral_names
    scale = max(abs(v) for v in ref.values())
    for k in ref:
        assert abs(got[k] - ref[k]) <= TOL * scale, (k, got[k], ref[k])
    asm.assemble(flag=1)
    again = asm.evaluate_integral_expressions()
    first = asm.integral_names[0]                          # "volume" for the bulk classes, "length" for the interface class
    assert again == got and asm.evaluate_observable(first) == got[first]
    op.close(); asm.close()


@pytest.mark.gpu
def test_edge_cases_single_element_everything_pinned_no_elements():
    """smallest and degenerate inputs: a one-element mesh; every dof of the class pinned (n_dof = 0, empty pattern: the launch must
    still run and write nothing); an assembler that was given no elements at all (a rank whose block is empty): zero matrix on the
    requested pattern, zero residual; bad arguments are reported, not computed around."""
    for kind in ("ns", "heat3d", "ale"):
        pb = make_problem(kind, 1)
        op, asm = make_oracle(pb), make_gpu(pb)
        n = pb["dofmap"].n_dof
        r_ref, mats = op.assemble(flag=1)
        asm.assemble(flag=1)
        r, jac, _ = asm.fetch()
        assert np.abs(r - r_ref).max() <= TOL * np.abs(r_ref).max()
        err, missing = compare_matrix(csr_to_sorted(n, asm.indptr, asm.indices, jac), csr_to_sorted(n, *mats[0]))
        assert missing == 0 and err <= TOL
        op.close(); asm.close()
    from pyoomph_b200.meshes import assign_equation_numbers
    pb = make_problem("ns", 3)
    everything = np.arange(pb["mesh"].n_node)
    pb["dofmap"] = assign_equation_numbers(pb["mesh"], pb["code"], {"velocity_x": everything, "velocity_y": everything, "pressure": everything})
    assert pb["dofmap"].n_dof == 0
    asm = make_gpu(pb)
    asm.assemble(flag=2)
    r, jac, mass = asm.fetch(True, True)
    assert asm.nnz == 0 and r.size == 0 and jac.size == 0 and mass.size == 0
    asm.close()
    pb = make_problem("ns", 3)
    from pyoomph_b200.assembly import B200Assembly
    asm = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, elements=np.zeros(0, dtype=np.int64),
                       extra_pattern=(np.array([0, 1], dtype=np.int32), np.array([1, 0], dtype=np.int32)))
    asm.assemble(flag=1)
    r, jac, _ = asm.fetch()
    assert asm.n_elem == 0 and asm.nnz == 2 and not jac.any()
    with pytest.raises(RuntimeError):
        asm.assemble(flag=3)
    with pytest.raises(ValueError):
        asm.assemble(flag=1, parameter="nope")
    asm.close()


@pytest.mark.gpu
def test_newton_solve_through_the_reference_facing_call():
    """Problem.solve()'s loop with the GPU assembler behind it: every iteration hands the host dof vector to assemble_host and gets the
    host residual and CSR values back (get_residuals_and_jacobian contract), SuperLU solves on the host.  Same iteration count and
    residual history as the oracle-driven loop, same converged flow."""
    from scipy.sparse import csr_matrix
    from test_oracle import _cavity, newton_cavity
    pb = _cavity(8)
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn
    m = eq >= 0
    op = make_oracle(pb)

    def set_dofs_cpu(U):
        v = pb["vals"][0].copy()
        v[m] = U[eq[m]]
        op.update_values(0, v)

    def assemble_cpu():
        r, mats = op.assemble(flag=1)
        return r, csr_to_sorted(n, *mats[0])
    U_ref, hist_ref = newton_cavity(pb, assemble_cpu, set_dofs_cpu)
    asm = make_gpu(pb)
    state = {}

    def assemble_gpu():
        r, jac, _ = asm.assemble_host(state["U"], 1)
        return r, csr_matrix((jac, asm.indices, asm.indptr), shape=(n, n))
    U, hist = newton_cavity(pb, assemble_gpu, lambda U_: state.__setitem__("U", U_))
    assert len(hist) == len(hist_ref) and hist[-1] < 1e-10
    assert np.abs(U - U_ref).max() <= 1e-9 * np.abs(U_ref).max()
    for a, b in zip(hist[:-1], hist_ref[:-1]):
        assert abs(a - b) <= 1e-6 * max(b, 1e-12) + 1e-12
    res, A = asm.get_residuals_and_jacobian(True)            # the CustomAssemblyBase entry point on the converged state
    assert np.abs(res).max() < 1e-10 and A.shape == (n, n) and A.indptr.dtype == np.int32 and A.indices.dtype == np.int32
    op.close(); asm.close()


@pytest.mark.gpu
def test_bdf2_time_stepping_of_config3_matches_the_oracle_driven_loop():
    """BASELINE config 3 as a time loop (Problem.run's unsteady_newton_solve): per step shift the history (device-side on the GPU), set the
    BDF weights (first step degraded to BDF1, src/elements.cpp:4611), one Newton solve of the linear problem; three steps with varying
    dt, GPU-assembled against oracle-assembled, same SuperLU: same trajectory."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.linalg import spsolve
    pb = make_problem("heat3d", 3)
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn
    m = eq >= 0
    start = pb["vals"][0].copy()
    pb["vals"] = np.stack([start, start, start])          # at rest before the first step
    op, asm = make_oracle(pb), make_gpu(pb)
    hist = [start.copy(), start.copy(), start.copy()]     # oracle side: nodal values per history level
    U_gpu = np.zeros(n); U_gpu[eq[m]] = start[m]
    t, dts = 0.0, [0.01, 0.012, 0.008]
    for step, dt in enumerate(dts):
        dtprev = dts[step - 1] if step else dt
        t += dt
        # --- oracle-driven step
        hist = [hist[0].copy(), hist[0], hist[1]]
        for lvl in range(3):
            op.update_values(lvl, hist[lvl])
        op.set_unsteady(t, dt, dtprev, step)
        r, mats = op.assemble(flag=1)
        U = np.zeros(n); U[eq[m]] = hist[0][m]
        U -= spsolve(csr_to_sorted(n, *mats[0]).tocsc(), r)
        hist[0][m] = U[eq[m]]
        # --- GPU-driven step: nothing but the dof vector crosses the host link
        asm.shift_time_values()
        asm.set_unsteady(t, dt, dtprev, step)
        rg, jac, _ = asm.assemble_host(U_gpu, 1)
        U_gpu = U_gpu - spsolve(csr_matrix((jac, asm.indices, asm.indptr), shape=(n, n)).tocsc(), rg)
        asm.set_dofs(U_gpu)
        assert np.abs(U_gpu - U).max() <= 1e-10 * np.abs(U).max(), step
    rg, _, _ = asm.assemble_host(None, 0)                 # linear problem: the step's residual is gone after one solve
    assert np.abs(rg).max() <= 1e-9 * np.abs(r).max()
    op.close(); asm.close()


@pytest.mark.gpu
def test_invalidate_rebuild_and_shift_time_values():
    """the packer's life cycle (SURVEY N-a): invalidate_cache() after a renumbering makes the next assembly fail loudly until
    rebuild(mesh, dofmap) re-packs (same compiled class, new pattern) -- results equal a fresh assembler's and the oracle's;
    shift_time_values() moves the history levels on the device like Problem::shift_time_values."""
    from pyoomph_b200.meshes import assign_equation_numbers
    pb = make_problem("ns_unsteady", 7)
    asm = make_gpu(pb)
    asm.assemble(flag=1)
    nnz0 = asm.nnz
    asm.actions_after_equation_numbering()
    with pytest.raises(RuntimeError):
        asm.assemble(flag=1)
    mesh = pb["mesh"]
    pb["dofmap"] = assign_equation_numbers(mesh, pb["code"], {"velocity_x": mesh.boundaries["left"], "velocity_y": mesh.boundaries["left"]})
    asm.rebuild(mesh, pb["dofmap"])
    assert asm.nnz != nnz0 and asm.n_dof == pb["dofmap"].n_dof
    for t in range(pb["vals"].shape[0]):
        asm.set_nodal_values(t, pb["vals"][t])
    op = make_oracle(pb)
    n = pb["dofmap"].n_dof
    r_ref, mats = op.assemble(flag=1)
    asm.assemble(flag=1)                         # time info survived the rebuild (BDF2 weights of make_gpu)
    r, jac, _ = asm.fetch()
    err, missing = compare_matrix(csr_to_sorted(n, asm.indptr, asm.indices, jac), csr_to_sorted(n, *mats[0]))
    assert missing == 0 and err <= TOL and np.abs(r - r_ref).max() <= TOL * np.abs(r_ref).max()
    op.close()
    # shift: level t <- level t-1, level 0 unchanged
    asm.shift_time_values()
    asm.assemble(flag=1)
    r_shift, j_shift, _ = asm.fetch()
    pb2 = dict(pb)
    pb2["vals"] = np.stack([pb["vals"][0], pb["vals"][0], pb["vals"][1]])
    ref = make_gpu(pb2)
    ref.assemble(flag=1)
    r2, j2, _ = ref.fetch()
    assert np.array_equal(r_shift, r2) and np.array_equal(j_shift, j2)
    asm.close(); ref.close()


@pytest.mark.gpu
def test_multi_assemble_request_matches_oracle():
    """MultiAssembleRequest (bifurcation_tools.py:449): R, J, M, dR/dp, dJ/dp, d(J.Y)/dU, d(M.Y)/dU from one request, in request
    order, against the oracle; the request needs 4 launches (one flag-2 launch, one parameter launch, one Hessian launch per vector)."""
    from scipy.sparse import csr_matrix
    from pyoomph_b200.multi_assembly import MultiAssembleRequest
    pb = make_problem("ns_param", 6)
    op = make_oracle(pb)
    asm = make_gpu(pb)
    n = pb["dofmap"].n_dof
    rng = np.random.default_rng(3)
    Y, Z = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    req = MultiAssembleRequest(asm).J().dJdU(Y).R().dRdp("mu").M().dMdU(Y).dJdp("mu").dJdU(Z)
    J, JY, R, Rp, M, MY, Jp, JZ = req.assemble()
    assert req.launches == 4
    r_ref, mats = op.assemble(flag=2)
    rp_ref, pmats = op.assemble(param=0, flag=1)
    Jh, Mh = op.assemble_hessian(np.stack([Y, Z]), flag=2)
    assert np.abs(R - r_ref).max() <= TOL * np.abs(r_ref).max() and np.abs(Rp - rp_ref).max() <= TOL * np.abs(rp_ref).max()
    for got, ref in ((J, csr_to_sorted(n, *mats[0])), (M, csr_to_sorted(n, *mats[1])), (Jp, csr_to_sorted(n, *pmats[0]))):
        err, missing = compare_matrix(got, ref)
        assert missing == 0 and err <= TOL
    for got, ref in ((JY, Jh[0]), (JZ, Jh[1]), (MY, Mh[0])):
        assert abs(got - ref).max() <= TOL * max(abs(ref).max(), 1e-300) if ref.nnz else abs(got).max() == 0.0
    # transposed contractions through the same request (hessian_vector_transposed of the reference's multi-assembly)
    JYt, MYt = MultiAssembleRequest(asm).dJdU(Y, transposed=True).dMdU(Y, transposed=True).assemble()
    Jt, Mt = op.assemble_hessian(Y[None, :], flag=5)
    for got, ref in ((JYt, Jt[0]), (MYt, Mt[0])):
        assert abs(got - ref).max() <= TOL * max(abs(ref).max(), 1e-300) if ref.nnz else abs(got).max() == 0.0
    # eigenproblem pair (Problem::assemble_eigenproblem_matrices): mass matrix and shifted Jacobian from one launch
    Me, Je = asm.assemble_eigenproblem_matrices(sigma_r=0.75)
    assert asm.launch_count() == 1 and abs(Me - M).max() == 0.0 and abs(Je - (J - 0.75 * M)).max() <= 1e-15 * abs(J).max()
    op.close(); asm.close()


@pytest.mark.gpu
def test_full_size_config2_windows_against_oracle():
    """BASELINE config 2 at FULL size (NS Taylor-Hood 1024x1024, 1.05 M elements, all 148 blocks, all tile gates): rows of the
    nodes interior to sampled 4x4-element windows -- domain corners and edges (pinned dofs), patch / unit / tile boundaries
    (multiples of 8 and 32 elements), random interior places -- against the CPU oracle assembled on the window alone; and the
    assembly is bit-reproducible at this size."""
    from windows import check_windows
    N = 1024
    pb = make_problem("ns", N)
    asm = make_gpu(pb)
    asm.assemble(flag=1)
    r, jac, _ = asm.fetch()
    rng = np.random.default_rng(5)
    windows = [(0, 0), (N - 4, N - 4), (0, N - 4), (N - 4, 0), (0, 510), (510, 0), (6, 6), (30, 30), (254, 510), (510, 766), (1018, 6)]
    windows += [tuple(int(x) for x in rng.integers(0, N - 4, size=2)) for _ in range(8)]
    worst = check_windows(pb, make_oracle, asm.indptr, asm.indices, jac, r, windows, w=4, tol=TOL)
    print("full-size window parity: worst row-scaled error %.2e over %d windows" % (worst, len(windows)))
    asm.assemble(flag=1)
    r2, jac2, _ = asm.fetch()
    assert np.array_equal(r, r2) and np.array_equal(jac, jac2)
    asm.close()


@pytest.mark.gpu
def test_full_size_config3_windows_against_oracle():
    """BASELINE config 3 at FULL size (transient heat, 126^3 = 2.0 M C2 bricks, BDF2): 2x2x2-element windows against the oracle."""
    from windows import check_windows
    N = 126
    pb = make_problem("heat3d", N)
    asm = make_gpu(pb)
    asm.assemble(flag=1)
    r, jac, _ = asm.fetch()
    rng = np.random.default_rng(7)
    windows = [(0, 0, 0), (N - 2, N - 2, N - 2), (0, 62, N - 2), (3, 3, 3), (62, 62, 62), (N - 2, 0, 63)]
    windows += [tuple(int(x) for x in rng.integers(0, N - 2, size=3)) for _ in range(4)]
    worst = check_windows(pb, make_oracle, asm.indptr, asm.indices, jac, r, windows, w=2, tol=TOL)
    print("full-size window parity (config 3): worst row-scaled error %.2e over %d windows" % (worst, len(windows)))
    asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N", [("ns_unsteady", 6), ("nlheat", 6), ("ns_axi_swirl", 4)])
def test_transposed_hessian_vector_products_parity(kind, N):
    """flags 4 and 5 of HessianVectorProduct<i> (src/jitbridge.h:637-691): T_ik = sum_j H_jik Y_j for the Jacobian and the mass
    Hessian, against the oracle's ndof^3-buffer routine; and against the non-transposed products through the symmetry
    T(Y) . C = (d(J.C)/dU)^T ... checked as a bilinear identity:  C^T T(Y) Z = Y^T N(C) Z  with N(C) = d(J.C)/dU."""
    pb = make_problem(kind, N)
    op = make_oracle(pb)
    asm = make_gpu(pb)
    n = pb["dofmap"].n_dof
    rng = np.random.default_rng(4)
    Y = rng.uniform(-1, 1, (2, n))
    Jr, Mr = op.assemble_hessian(Y, flag=5)
    Jg, Mg = asm.assemble_hessian(Y, flag=2, transposed=True)
    from scipy.sparse import csr_matrix
    for ref, got in list(zip(Jr, Jg)) + list(zip(Mr, Mg)):
        A = csr_matrix((got, asm.indices, asm.indptr), shape=(n, n))
        assert abs(A - ref).max() <= TOL * max(abs(ref).max(), 1e-300) if ref.nnz else np.abs(got).max() == 0.0
    C, Z = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    Tn, _ = asm.assemble_hessian(Y[:1], flag=1, transposed=True)
    Nc, _ = asm.assemble_hessian(C[None, :], flag=1)
    lhs = C @ (csr_matrix((Tn[0], asm.indices, asm.indptr), shape=(n, n)) @ Z)
    rhs = Y[0] @ (csr_matrix((Nc[0], asm.indices, asm.indptr), shape=(n, n)) @ Z)
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs), 1e-300)
    op.close(); asm.close()


@pytest.mark.gpu
def test_hessian_tensor_matches_the_reference_accumulation():
    """Problem::assemble_hessian_tensor (src/problem.cpp:1530-1560) on the GPU (one Hessian-vector launch per tensor slice) against the
    restated element loop over the oracle's flag-3 buffers, as SparseRank3Tensors: same (i, j, k) pattern up to exact cancellations,
    values 1e-12; and the tensor-vector product equals d(J.v)/dU."""
    from scipy.sparse import csr_matrix
    from pyoomph_b200.hessian_tensor import SparseRank3Tensor
    pb = make_problem("nlheat", 3)
    op, asm = make_oracle(pb), make_gpu(pb)
    n = pb["dofmap"].n_dof
    T = asm.assemble_hessian_tensor()
    ii, jj, kk, vv = op.assemble_hessian_tensor()
    R = SparseRank3Tensor(n)
    R.accumulate(ii, jj, kk, vv)
    dg = {(a, b, c): v for a, b, c, v in T.get_entries()}
    dr = {(a, b, c): v for a, b, c, v in R.get_entries()}
    scale = max(abs(v) for v in dr.values())
    for key in set(dg) | set(dr):
        assert abs(dg.get(key, 0.0) - dr.get(key, 0.0)) <= TOL * scale, key
    assert len(set(dg) ^ set(dr)) <= 0.02 * len(dr)
    v = np.random.default_rng(9).uniform(-1, 1, n)
    ci, rs = T.finalize_for_vector_product()
    M = csr_matrix((T.right_vector_mult(v), ci, rs), shape=(n, n))
    Nv, _ = asm.assemble_hessian(v[None, :], flag=1)
    assert abs(M - csr_matrix((Nv[0], asm.indices, asm.indptr), shape=(n, n))).max() <= 1e-12 * abs(M).max()
    op.close(); asm.close()


@pytest.mark.gpu
def test_newton_iteration_with_a_device_resident_solver():
    """SURVEY N-d: assemble on the device, solve on the device (registered DeviceLinearSystemSolver plugin), update the dofs on the
    device; the matrix never reaches the host.  One step of the (linear) transient heat problem ends where the host-side SuperLU step
    ends, and the solver registry behaves like pyoomph's."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.linalg import spsolve
    from pyoomph_b200.solvers import DeviceLinearSystemSolver, GenericLinearSystemSolver
    pb = make_problem("heat3d", 3)
    asm = make_gpu(pb)
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn
    m = eq >= 0
    U0 = np.zeros(n); U0[eq[m]] = pb["vals"][0][m]
    r, jac, _ = asm.assemble_host(U0, 1)
    U_ref = U0 - spsolve(csr_matrix((jac, asm.indices, asm.indptr), shape=(n, n)).tocsc(), r)
    solver = GenericLinearSystemSolver.factory_solver("torch_krylov")
    assert isinstance(solver, DeviceLinearSystemSolver)
    asm.set_dofs(U0)
    rmax, stats = asm.newton_step_on_device(solver)
    assert abs(rmax - np.abs(r).max()) <= 1e-12 * np.abs(r).max() and stats["relative_residual"] <= 1e-11
    U = asm.fetch_dofs()
    assert np.abs(U - U_ref).max() <= 1e-8 * np.abs(U_ref).max()
    rmax2, _ = asm.newton_step_on_device(solver)                # the problem is linear: the residual is gone after one step
    assert rmax2 <= 1e-8 * rmax
    with pytest.raises(RuntimeError):
        GenericLinearSystemSolver.factory_solver("no_such_solver")
    with pytest.raises(RuntimeError):
        GenericLinearSystemSolver.register_solver()(type("Dup", (GenericLinearSystemSolver,), {"idname": "torch_krylov"}))
    asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,distortion,unstructured", [("ns_pts", 7, 0.1, False), ("heat3d_pts", 3, 0.1, True), ("ale_axi_pts", 5, 0.08, True), ("ns_pts", 60, 0.0, False)])
def test_point_expressions_parity(kind, N, distortion, unstructured):
    """EvalLocalExpression / EvalExtremumExpression / GetZ2Fluxes (src/codegen.cpp:4366-4453) for every element at its nodes and at its
    integration points: one launch per point set against one reference-style call per (element, point, expression) of the oracle;
    Mesh::evaluate_extremum's scan (src/mesh.cpp:444-500) on top of it."""
    pb = make_problem(kind, N, distortion=distortion, unstructured=unstructured)
    asm = make_gpu(pb)
    op = make_oracle(pb)
    names = asm.point_names
    assert names == pb["code"].point_expression_names()
    small = pb["mesh"].n_elem <= 200
    for pts in ("nodes", "gauss"):
        got = asm.evaluate_point_expressions(pts)
        assert asm.launch_count() == 1
        if small:
            ref = op.eval_point_expressions(pts)
            assert got.shape == ref.shape
            for i in range(len(names)):
                scale = max(np.abs(ref[:, :, i]).max(), 1e-300)
                assert np.abs(got[:, :, i] - ref[:, :, i]).max() <= TOL * scale, (pts, names[i])
        again = asm.evaluate_point_expressions(pts)
        assert np.array_equal(got, again)
    loc = asm.evaluate_local_expressions_at_nodes()
    assert list(loc) == [n for k, n in names if k == "local"] and all(v.shape == (pb["mesh"].n_elem, pb["mesh"].elem_nodes.shape[1]) for v in loc.values())
    z2 = asm.get_Z2_fluxes()
    assert z2.shape[2] == sum(1 for k, _ in names if k == "z2")
    # extremum: the reference's sampling (integration points, then nodes, element by element, strict ">") on the GPU values
    xname = [n for k, n in names if k == "extremum"][0]
    xi = [i for i, (k, n) in enumerate(names) if k == "extremum"][0]
    for sign in (1, -1):
        val, elem, where = asm.evaluate_extremum(xname, sign)
        g, nd = asm.evaluate_point_expressions("gauss")[:, :, xi], asm.evaluate_point_expressions("nodes")[:, :, xi]
        best, be, bw = sign * nd[0, 0], 0, ("nodes", 0)
        for e in range(g.shape[0]):
            for q in range(g.shape[1]):
                if sign * g[e, q] > best:
                    best, be, bw = sign * g[e, q], e, ("gauss", q)
            for q in range(nd.shape[1]):
                if sign * nd[e, q] > best:
                    best, be, bw = sign * nd[e, q], e, ("nodes", q)
        assert (val, elem, where) == (sign * best, be, bw)
    with pytest.raises(RuntimeError):
        asm.evaluate_extremum("no_such_expression")
    op.close(); asm.close()


INTERFACES = [("robin_if", 6, 0.12), ("robin_if", 9, 0.0), ("freesurf_if", 5, 0.1), ("freesurf_if", 24, 0.05),
              # config 4's interface class on a MOVING mesh: kinematic condition with the mesh velocity, position dofs in the Jacobian
              ("freesurf_mov_if", 5, 0.1), ("freesurf_mov_if", 20, 0.06), ("freesurf_mov_axi_if", 6, 0.08),
              # faces seen through their bulk elements (bulk_eleminfo access): Nitsche's method with normal derivatives of field and test function
              ("nitsche_face", 5, 0.12), ("nitsche_face", 16, 0.0)]


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,distortion", INTERFACES)
def test_interface_element_classes_parity(kind, N, distortion):
    """Line elements in a 2D nodal space (InterfaceElementLine1dC2: tangent, outer normal and surface gradients from a 1x2 mapping)
    assembled on their own: residual and Jacobian against the oracle, CSR bit-exact after the zero-drop rule."""
    pb = make_problem(kind, N, distortion=distortion)
    op = make_oracle(pb)
    asm = make_gpu(pb)
    n = pb["dofmap"].n_dof
    r_ref, mats = op.assemble(flag=1)
    asm.assemble(flag=1)
    r, jac, _ = asm.fetch(True, False)
    scale = np.abs(r_ref).max()
    assert np.abs(r - r_ref).max() <= TOL * scale
    A = csr_to_sorted(n, asm.indptr, asm.indices, jac)
    B = csr_to_sorted(n, *mats[0])
    err, missing = compare_matrix(A, B)
    assert missing == 0 and err <= TOL, (kind, err, missing)
    # moving interface: sliding a node ALONG the line changes no integral, so the tangential position columns are analytic zeros that
    # come out as cancellation noise on both sides (measured 15 % of the entries); no value is outside the bar on the row scale
    st = assert_csr_parity(asm.indptr, asm.indices, jac, mats[0], TOL, label="%s N=%d" % (kind, N),
                           max_cancel_fraction=0.30 if (distortion == 0 or kind.startswith("freesurf_mov")) else 0.02)
    _record(kind, N, distortion, False, st)
    asm.assemble(flag=0)
    r0, _, _ = asm.fetch(False, False)
    assert np.abs(r0 - r).max() <= 1e-13 * scale
    op.close()
    asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,distortion", [("robin_if", 7, 0.1), ("freesurf_if", 6, 0.1), ("freesurf_if", 40, 0.0), ("freesurf_mov_if", 6, 0.08), ("nitsche_face", 7, 0.1), ("freesurf_mov_axi_if", 6, 0.06)])
def test_interface_class_assembled_into_the_matrix_of_its_bulk_class(kind, N, distortion):
    """A child problem (pb2_problem_create_child): the interface class scatters into the CSR matrix and residual its bulk class just
    wrote, on the device; the sum equals the oracle's two classes assembled into one matrix (oomph assembles all element classes of a
    Problem into one matrix, problem.cc:5332-5666)."""
    from pyoomph_b200.assembly import B200Assembly
    pb = make_problem(kind, N, distortion=distortion)
    bulk_pb = dict(pb, code=pb["bulk_code"], mesh=pb["bulk_mesh"])
    bulk = make_gpu(bulk_pb)
    child = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, parent=bulk)
    ob, oi = make_oracle(bulk_pb), make_oracle(pb)
    n = pb["dofmap"].n_dof
    rb, mb = ob.assemble(flag=1)
    ri, mi = oi.assemble(flag=1)
    J_ref = (csr_to_sorted(n, *mb[0]) + csr_to_sorted(n, *mi[0])).tocsr()
    J_ref.sort_indices()
    r_ref = rb + ri
    for rep in range(2):                       # the second pass starts from the first one's values: first-touch stores, then adds
        bulk.assemble(flag=1)
        r, jac, _ = bulk.fetch(True, False)
        assert np.abs(r - r_ref).max() <= TOL * np.abs(r_ref).max()
        err, missing = compare_matrix(csr_to_sorted(n, bulk.indptr, bulk.indices, jac), J_ref)
        assert missing == 0 and err <= TOL, (kind, rep, err, missing)
    assert bulk.launch_count() == 2            # one persistent launch per element class
    # the interface class really contributed
    assert np.abs(ri).max() > 1e-3 * np.abs(r_ref).max()
    # the bulk class alone, afterwards, gives the bulk matrix again (the child only ever adds to what the parent stored)
    child.close()
    bulk.children.clear()
    bulk.assemble(flag=1)
    r2, jac2, _ = bulk.fetch(True, False)
    assert np.abs(r2 - rb).max() <= TOL * np.abs(rb).max()
    for o in (ob, oi):
        o.close()
    bulk.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,distortion", [("poisson_hang", 8, 0.12), ("ns_hang", 6, 0.1), ("ns_unsteady_hang", 6, 0.08), ("ns_hang", 24, 0.05),
                                                ("heat3d_hang", 3, 0.08), ("heat3d_hang", 5, 0.0), ("ale_hang", 6, 0.08)])
def test_hanging_nodes_parity(kind, N, distortion):
    """a14: a mesh with hanging nodes (one quadtree level).  Oracle: the reference's hang macros inside the element routine with its
    local numbering of the master values.  Product: the unchanged element kernel over virtual equations for the hanging values, then
    P^T J_ext P on the device (fixed-order gathers).  Compared: residual, Jacobian, mass matrix of the real equations; CSR bit-exact after
    the zero-drop rule."""
    from problems import TIME
    from pyoomph_b200.hanging import HangingNodeAssembly
    pb = make_problem(kind, N, distortion=distortion)
    op = make_oracle(pb)
    n = pb["dofmap"].n_dof
    flag = 2 if pb["unsteady"] else 1
    r_ref, mats = op.assemble(flag=flag)
    asm = HangingNodeAssembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name)
    assert asm.n_dof == n and asm.n_ext > n
    for t in range(pb["vals"].shape[0]):
        asm.set_nodal_values(t, pb["vals"][t])
    if pb["pos_hist"] is not None:         # moving mesh: the positions of the hanging nodes hang on their masters' position dofs
        for t in range(pb["pos_hist"].shape[0]):
            asm.set_nodal_positions(t, pb["pos_hist"][t])
    if pb["unsteady"]:
        asm.set_unsteady(TIME["t"], TIME["dt"], TIME["dtprev"], TIME["unsteady_steps_done"])
    else:
        asm.set_steady()
    for rep in range(2):                   # the second assembly starts from the reduced matrix of the first
        asm.assemble(flag=flag)
        r, jac, mass = asm.fetch(True, flag == 2)
        assert np.abs(r - r_ref).max() <= TOL * np.abs(r_ref).max()
        for vals, ref in zip((jac, mass), mats):
            err, missing = compare_matrix(csr_to_sorted(n, asm.indptr, asm.indices, vals), csr_to_sorted(n, *ref))
            assert missing == 0 and err <= TOL, (kind, rep, err, missing)
    st = assert_csr_parity(asm.indptr, asm.indices, jac, mats[0], TOL, label="%s N=%d" % (kind, N), max_cancel_fraction=0.02 if distortion > 0 else 0.30)
    _record(kind, N, distortion, False, st)
    # the device holds J (+) I over the extended equations: virtual rows / columns cleared, unit diagonal, zero residual
    r_ext, j_ext, _ = asm.fetch_extended(True, False)
    A = csr_to_sorted(asm.n_ext, asm.asm.indptr, asm.asm.indices, j_ext)
    nv = asm.n_ext - n
    assert np.array_equal(A[n:, n:].toarray(), np.eye(nv)) and abs(A[n:, :n]).max() == 0.0 and abs(A[:n, n:]).max() == 0.0
    assert np.all(r_ext[n:] == 0.0)
    # a dof vector scattered on the device: hanging values follow their masters (pinned masters included)
    ne = pb["dofmap"].node_eqn
    u = np.zeros(n)
    u[ne[ne >= 0]] = pb["vals"][0][ne >= 0]
    if pb["pos_hist"] is not None:
        pe = pb["dofmap"].pos_eqn
        u[pe[pe >= 0]] = pb["pos_hist"][0][pe >= 0]
    asm.set_dofs(u)
    asm.assemble(flag=1)
    r2, jac2, _ = asm.fetch(True, False)
    assert np.abs(r2 - r_ref).max() <= TOL * np.abs(r_ref).max()
    # bit-reproducible (fixed-order gathers, no atomics in the reduction)
    asm.assemble(flag=1)
    r3, jac3, _ = asm.fetch(True, False)
    assert np.array_equal(r2, r3) and np.array_equal(jac2, jac3)
    op.close()
    asm.close()


def _azimuthal_maps(pb, pbx):
    """selection matrices between the product's numbering and the checker's extended class: rows = base-field equations, columns =
    equations of the mode copies (scipy CSR, n_ext x n)"""
    from scipy.sparse import csr_matrix
    from pyoomph_b200.expressions import MODE_SUFFIX
    n, nx = pb["dofmap"].n_dof, pbx["dofmap"].n_dof
    rr, rc, cr, cc = [], [], [], []
    for f in pb["code"].nodal_fields():
        g = pb["dofmap"].node_eqn[:, f.index]
        gb = pbx["dofmap"].node_eqn[:, pbx["code"].fields[f.name].index]
        gm = pbx["dofmap"].node_eqn[:, pbx["code"].fields[f.name + MODE_SUFFIX].index]
        live = g >= 0
        assert np.array_equal(live, gb >= 0) and np.array_equal(live, gm >= 0)
        rr += list(gb[live]); rc += list(g[live]); cr += list(gm[live]); cc += list(g[live])
    Sr = csr_matrix((np.ones(len(rr)), (rr, rc)), shape=(nx, n))
    Sc = csr_matrix((np.ones(len(cr)), (cr, cc)), shape=(nx, n))
    return Sr, Sc


@pytest.mark.gpu
@pytest.mark.parametrize("base,mode_param,N,distortion", [("ns_azi", "azimuthal_m", 4, 0.1), ("ns_azi", "azimuthal_m", 7, 0.0),
                                                          ("ns_kz", "normal_mode_k", 5, 0.1)])
def test_azimuthal_mode_contributions_parity(base, mode_param, N, distortion):
    """BASELINE config 5: real and imaginary contribution of the azimuthal (m = 1, then m = 2) eigenproblem of axisymmetric NS with swirl:
    Jacobian and mass matrix with respect to the mode fields, and the Hessian-vector products d(J.Y)/dU with respect to the BASE state
    that the azimuthal Hopf / fold trackers assemble.  The checker assembles an extended element class in which the mode fields are
    ordinary nodal fields and takes the (base rows, mode columns) block."""
    from problems import TIME
    pb = make_problem(base, N, distortion=distortion)                     # "ns_kz": the Cartesian sibling, normal mode exp(i k z)
    pbx = make_problem(base + "_ext", N, distortion=distortion)
    Sr, Sc = _azimuthal_maps(pb, pbx)
    n, nx = pb["dofmap"].n_dof, pbx["dofmap"].n_dof
    asm = make_gpu(pb)
    op = make_oracle(pbx)
    names = pb["code"].residual_names()
    assert names == pbx["code"].residual_names() and len(names) == 3
    rng = np.random.default_rng(5)
    for m in (1.0, 2.0):
        asm.set_parameters(**{mode_param: m})
        op.set_params([m])
        for which, rn in enumerate(names):
            r_x, mats = op.assemble(which=which, flag=2)
            asm.assemble(flag=2, residual=rn)
            r, jac, mass = asm.fetch(True, True)
            r_ref = Sr.T @ r_x
            assert np.abs(r - r_ref).max() <= TOL * max(np.abs(r_ref).max(), 1e-300), (rn, m)
            for vals, (rs, ci, va) in zip((jac, mass), mats):
                Jx = csr_to_sorted(nx, rs, ci, va)
                # base residual: its own Jacobian (base columns); contributions: the columns of the mode copies
                B = (Sr.T @ Jx @ (Sr if which == 0 else Sc)).tocsr()
                B.sort_indices()
                A = csr_to_sorted(n, asm.indptr, asm.indices, vals)
                if B.nnz == 0 or abs(B).max() == 0.0:
                    assert np.abs(vals).max() == 0.0
                    continue
                err, missing = compare_matrix(A, B)
                assert missing == 0 and err <= TOL, (rn, m, err, missing)
        # Hessian-vector products of the contributions: derivative with respect to the base state, contracted with a mode vector
        Y = rng.standard_normal(n)
        for which, rn in enumerate(names[1:], start=1):
            HJ, HM = asm.assemble_hessian(Y[None, :], flag=2, residual=rn)
            Jr, Mr = op.assemble_hessian((Sc @ Y)[None, :], flag=2, which=which)
            for vals, ref in ((HJ[0], Jr[0]), (HM[0], Mr[0])):
                B = (Sr.T @ ref @ Sr).tocsr()
                A = csr_to_sorted(n, asm.indptr, asm.indices, vals)
                if abs(B).max() == 0.0:
                    assert np.abs(vals).max() <= 1e-300
                    continue
                D = abs(A - B)
                assert D.max() <= 1e-11 * abs(B).max(), (rn, m, D.max(), abs(B).max())
    # the complex eigenproblem pair of mode m = 2 from the two contributions
    if base != "ns_azi":
        op.close()
        asm.close()
        return
    Mc, Jc = asm.assemble_azimuthal_eigenproblem_matrices(2.0)
    asm.assemble(flag=2, residual=names[1])
    _, jr, mr = asm.fetch(True, True)
    asm.assemble(flag=2, residual=names[2])
    _, ji, mi = asm.fetch(True, True)
    assert np.array_equal(Jc.data, jr + 1j * ji) and np.array_equal(Mc.data, mr + 1j * mi) and np.abs(ji).max() > 0
    op.close()
    asm.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kind,N,distortion", [("ale", 3, 0.08), ("ale_axi", 3, 0.08), ("ale_tri", 2, 0.08)])
def test_moving_mesh_hessian_vector_products(kind, N, distortion):
    """d(J.Y)/dU and d(M.Y)/dU with POSITION dofs among the unknowns (the reference's second-order moving-mesh tensors,
    src/elements.cpp:3163-3217): every column -- nodal values and nodal positions -- against central differences of the ORACLE's Jacobian
    and mass matrix (the reference checks its analytic derivatives the same way, src/elements.cpp:5880).  Tolerance 2e-6 of the largest
    entry: finite-difference accuracy, not round-off."""
    from scipy.sparse import csr_matrix
    pb = make_problem(kind, N, distortion=distortion)
    assert pb["code"].coordinates_as_dofs
    n = pb["dofmap"].n_dof
    asm = make_gpu(pb)
    op = make_oracle(pb)
    rng = np.random.default_rng(11)
    Y = rng.standard_normal(n)
    HJ, HM = asm.assemble_hessian(Y[None, :], flag=2)
    A = csr_matrix((HJ[0], asm.indices, asm.indptr), shape=(n, n)).toarray()
    AM = csr_matrix((HM[0], asm.indices, asm.indptr), shape=(n, n)).toarray()
    TJ, TM = asm.assemble_hessian(Y[None, :], flag=2, transposed=True)            # flags 4 / 5: d(J^T.Y)/dU, d(M^T.Y)/dU
    AT = csr_matrix((TJ[0], asm.indices, asm.indptr), shape=(n, n)).toarray()
    ATM = csr_matrix((TM[0], asm.indices, asm.indptr), shape=(n, n)).toarray()
    refT, refTM = np.zeros((n, n)), np.zeros((n, n))
    cols_T = {}

    def JY_MY():
        _, mats = op.assemble(flag=2)
        Jm, Mm = csr_to_sorted(n, *mats[0]), csr_to_sorted(n, *mats[1])
        cols_T["J"], cols_T["M"] = Jm.T @ Y, Mm.T @ Y
        return Jm @ Y, Mm @ Y
    eps = 1e-6
    ref, refM = np.zeros((n, n)), np.zeros((n, n))
    vals0, pos0 = pb["vals"][0], pb["pos_hist"][0]
    for node in range(pb["mesh"].n_node):
        for f in range(vals0.shape[1]):
            g = pb["dofmap"].node_eqn[node, f]
            if g < 0:
                continue
            v = vals0.copy()
            v[node, f] += eps
            op.update_values(0, v)
            jp, mp = JY_MY()
            tp = dict(cols_T)
            v[node, f] -= 2 * eps
            op.update_values(0, v)
            jm, mm = JY_MY()
            ref[:, g], refM[:, g] = (jp - jm) / (2 * eps), (mp - mm) / (2 * eps)
            refT[:, g], refTM[:, g] = (tp["J"] - cols_T["J"]) / (2 * eps), (tp["M"] - cols_T["M"]) / (2 * eps)
        op.update_values(0, vals0)
        for d in range(pos0.shape[1]):
            g = pb["dofmap"].pos_eqn[node, d]
            if g < 0:
                continue
            x = pos0.copy()
            x[node, d] += eps
            op.update_values(0, None, x)
            jp, mp = JY_MY()
            tp = dict(cols_T)
            x[node, d] -= 2 * eps
            op.update_values(0, None, x)
            jm, mm = JY_MY()
            ref[:, g], refM[:, g] = (jp - jm) / (2 * eps), (mp - mm) / (2 * eps)
            refT[:, g], refTM[:, g] = (tp["J"] - cols_T["J"]) / (2 * eps), (tp["M"] - cols_T["M"]) / (2 * eps)
        op.update_values(0, None, pos0)
    pos_cols = pb["dofmap"].pos_eqn[pb["dofmap"].pos_eqn >= 0]
    assert np.abs(ref[:, pos_cols]).max() > 1e-3 * np.abs(ref).max()          # the position columns are not a side show
    assert np.abs(A - ref).max() <= 2e-6 * np.abs(ref).max(), (np.abs(A - ref).max(), np.abs(ref).max())
    assert np.abs(AM - refM).max() <= 2e-6 * max(np.abs(refM).max(), 1e-300), (np.abs(AM - refM).max(), np.abs(refM).max())
    assert np.abs(AT - refT).max() <= 2e-6 * np.abs(refT).max(), (np.abs(AT - refT).max(), np.abs(refT).max())
    assert np.abs(ATM - refTM).max() <= 2e-6 * max(np.abs(refTM).max(), 1e-300), (np.abs(ATM - refTM).max(), np.abs(refTM).max())
    op.close()
    asm.close()


@pytest.mark.gpu
def test_newton_iterations_of_the_coupled_free_surface_problem():
    """BASELINE config 4 as a user runs it, in small: NS Taylor-Hood on a pseudo-elastic moving mesh (bulk class) + the free-surface class
    on two boundaries (kinematic condition with the mesh velocity, multiplier on the position equations), one BDF2 step solved by Newton's
    method.  GPU: parent + child problem assemble ONE matrix on the device per iteration; oracle: both classes assembled and summed.  Same
    SuperLU on both sides: same iterates (positions, velocities, pressure, multipliers), same residual history."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.linalg import splu
    from pyoomph_b200.assembly import B200Assembly
    pb = make_problem("freesurf_mov_if", 6, distortion=0.05)
    bulk_pb = dict(pb, code=pb["bulk_code"], mesh=pb["bulk_mesh"])
    n = pb["dofmap"].n_dof
    eq, peq = pb["dofmap"].node_eqn, pb["dofmap"].pos_eqn
    m, mp = eq >= 0, peq >= 0
    U0 = np.zeros(n)
    U0[eq[m]] = 0.05 * pb["vals"][0][m]                      # a gentle start: small velocities, slightly displaced mesh
    U0[peq[mp]] = pb["pos_hist"][0][mp]
    for p_ in (pb, bulk_pb):
        p_["vals"] = 0.05 * pb["vals"]

    def newton(assemble, set_dofs):
        U, hist = U0.copy(), []
        for it in range(8):
            set_dofs(U)
            r, J = assemble()
            hist.append(float(np.abs(r).max()))
            if hist[-1] < 1e-11:
                break
            U = U - splu(J.tocsc()).solve(r)
        return U, hist

    ob, oi = make_oracle(bulk_pb), make_oracle(pb)

    def set_cpu(U):
        v, x = bulk_pb["vals"][0].copy(), pb["pos_hist"][0].copy()
        v[m] = U[eq[m]]
        x[mp] = U[peq[mp]]
        for o in (ob, oi):
            o.update_values(0, v, x)

    def assemble_cpu():
        rb, mb = ob.assemble(flag=1)
        ri, mi = oi.assemble(flag=1)
        return rb + ri, (csr_to_sorted(n, *mb[0]) + csr_to_sorted(n, *mi[0])).tocsr()
    U_ref, hist_ref = newton(assemble_cpu, set_cpu)

    bulk = make_gpu(bulk_pb)
    child = B200Assembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, parent=bulk)

    def assemble_gpu():
        bulk.assemble(flag=1)
        r, jac, _ = bulk.fetch(True, False)
        return r, csr_matrix((jac, bulk.indices, bulk.indptr), shape=(n, n))
    U, hist = newton(assemble_gpu, bulk.set_dofs)
    assert hist_ref[-1] < 1e-11 and len(hist_ref) <= 7, hist_ref          # quadratic convergence of the coupled problem
    assert len(hist) == len(hist_ref)
    for a, b in zip(hist[:-1], hist_ref[:-1]):
        assert abs(a - b) <= 1e-6 * max(b, 1e-12) + 1e-12, (hist, hist_ref)
    assert np.abs(U - U_ref).max() <= 1e-9 * np.abs(U_ref).max()
    # the free surface moved and carries a multiplier: the interface class was part of the solve
    lam = pb["code"].fields["_kin_bc"].index
    assert np.abs(U[eq[:, lam][eq[:, lam] >= 0]]).max() > 1e-6 and np.abs(U[peq[mp]] - U0[peq[mp]]).max() > 1e-6
    assert child in bulk.children
    for o in (ob, oi):
        o.close()
    bulk.close()


@pytest.mark.gpu
def test_integral_gradients_for_global_constraints():
    """Dense rows of global constraints in bordered form (pyoomph's GlobalLagrangeMultiplier): c = d(integral expression)/dU as
    the residual vector of an automatically generated contribution, its second derivative as that contribution's Jacobian; against the
    oracle's routine and against central differences of the oracle's integral values."""
    pb = make_problem("ns_constraint", 7, distortion=0.1)
    asm = make_gpu(pb)
    op = make_oracle(pb)
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn
    names = pb["code"].residual_names()
    vals_gpu = asm.evaluate_integral_expressions()
    vals_ref = op.evaluate_integral_expressions()
    for iname in pb["code"].integral_expression_names():
        assert abs(vals_gpu[iname] - vals_ref[iname]) <= TOL * abs(vals_ref[iname])
        which = names.index("d_integral_" + iname)
        c = asm.integral_gradient(iname)
        c_ref, mats = op.assemble(which=which, flag=1)
        assert np.abs(c - c_ref).max() <= TOL * np.abs(c_ref).max()
        asm.assemble(flag=1, residual="d_integral_" + iname)
        _, H, _ = asm.fetch(True, False)
        B = csr_to_sorted(n, *mats[0])
        if abs(B).max() > 0:
            err, missing = compare_matrix(csr_to_sorted(n, asm.indptr, asm.indices, H), B)
            assert missing == 0 and err <= TOL
        else:
            assert np.abs(H).max() == 0.0                       # a linear functional (the pressure integral) has no second derivative
        eps = 1e-6
        for node in range(0, pb["mesh"].n_node, 23):
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
                assert abs((ip - im) / (2 * eps) - c[g]) <= 1e-7 * max(np.abs(c).max(), 1e-300)
    op.close()
    asm.close()


@pytest.mark.gpu
def test_newton_and_time_step_drivers():
    """pyoomph_b200.newton: Problem::newton_solve / unsteady_newton_solve on top of the assembly path.  (1) the lid-driven cavity: the
    same iterates as the oracle-driven loop with the same SuperLU, quadratic convergence; (2) the same solve with the matrix staying on
    the device (solver plugin); (3) one BDF2 step of config 3 against the oracle-assembled step; (4) a mesh with hanging nodes."""
    from scipy.sparse.linalg import splu
    from problems import TIME
    from test_oracle import _cavity, newton_cavity
    from pyoomph_b200.hanging import HangingNodeAssembly
    from pyoomph_b200.newton import newton_solve, unsteady_newton_solve
    from pyoomph_b200.solvers import GenericLinearSystemSolver
    pb = _cavity(8)
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn
    m = eq >= 0
    op = make_oracle(pb)

    def set_cpu(U):
        v = pb["vals"][0].copy()
        v[m] = U[eq[m]]
        op.update_values(0, v)

    def assemble_cpu():
        r, mats = op.assemble(flag=1)
        return r, csr_to_sorted(n, *mats[0])
    U_ref, hist_ref = newton_cavity(pb, assemble_cpu, set_cpu)
    asm = make_gpu(pb)
    U0 = np.zeros(n)
    U0[eq[m]] = pb["vals"][0][m]
    U, hist = newton_solve(asm, U0, tol=1e-10, max_iter=12)
    assert len(hist) == len(hist_ref) and np.abs(U - U_ref).max() <= 1e-9 * np.abs(U_ref).max()
    assert hist[-1] < 1e-10 and hist[-2] < 1e-4 and hist[-3] > hist[-2] ** 0.75          # quadratic tail
    Ud, histd = newton_solve(asm, U0, tol=1e-9, max_iter=12, device_solver=GenericLinearSystemSolver.factory_solver("torch_krylov"))
    assert np.abs(Ud - U_ref).max() <= 1e-7 * np.abs(U_ref).max() and histd[-1] < 1e-9
    op.close()
    asm.close()
    # (3) one implicit step of the Q27 heat equation
    pb = make_problem("heat3d", 3)
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn[:, 0]
    asm = make_gpu(pb)
    op = make_oracle(pb)
    U0 = np.zeros(n)
    U0[eq[eq >= 0]] = pb["vals"][0][eq >= 0, 0]
    dt, dtprev = 0.02, TIME["dt"]
    U1, h1 = unsteady_newton_solve(asm, U0, TIME["t"], dt, dtprev, 3, tol=1e-11)
    vals = pb["vals"].copy()
    vals[2], vals[1] = vals[1], vals[0]                       # shift_time_values
    for t_ in range(3):
        op.update_values(t_, vals[t_])
    op.set_unsteady(TIME["t"] + dt, dt, dtprev, 3)
    U = U0.copy()
    for _ in range(4):
        v = vals[0].copy()
        v[eq >= 0, 0] = U[eq[eq >= 0]]
        op.update_values(0, v)
        r, mats = op.assemble(flag=1)
        if np.abs(r).max() < 1e-11:
            break
        U = U - splu(csr_to_sorted(n, *mats[0]).tocsc()).solve(r)
    assert len(h1) == 2 and np.abs(U1 - U).max() <= 1e-10 * np.abs(U).max()          # a linear problem: one Newton step
    op.close()
    asm.close()
    # (4) hanging nodes: the driver works on the real equations, the virtual ones stay inside
    pb = make_problem("poisson_hang", 6, distortion=0.1)
    n = pb["dofmap"].n_dof
    h = HangingNodeAssembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name)
    h.set_nodal_values(0, pb["vals"][0])
    h.set_steady()
    Uh, hh = newton_solve(h, np.zeros(n), tol=1e-11)
    op = make_oracle(pb)
    ne = pb["dofmap"].node_eqn[:, 0]
    v = pb["vals"][0].copy()
    v[ne >= 0, 0] = Uh[ne[ne >= 0]]
    for nn_, (ms, w) in pb["mesh"].hanging.C2.items():
        v[nn_, 0] = v[ms, 0] @ w
    op.update_values(0, v)
    r, _ = op.assemble(flag=0)
    assert np.abs(r).max() < 1e-10 and len(hh) == 2
    op.close()
    h.close()


@pytest.mark.gpu
def test_global_constraint_with_a_lagrange_multiplier_in_bordered_form():
    """The pressure level of the lid-driven cavity fixed by  integral(p) = 0  through one global Lagrange multiplier (pyoomph's
    GlobalLagrangeMultiplier; SURVEY 8e: such dense rows are not part of the CSR matrix): Newton's method on [[J, b], [c^T, 0]] with
    b = dR/d(lambda) (parameter-derivative routine) and c = d(integral)/dU (gradient contribution), every piece assembled on the GPU,
    against the same loop on oracle-assembled pieces."""
    from scipy.sparse import bmat, csr_matrix
    from scipy.sparse.linalg import splu
    pb = make_problem("ns_mean_pressure", 6, distortion=0.05)
    code = pb["code"]
    n = pb["dofmap"].n_dof
    eq = pb["dofmap"].node_eqn
    m = eq >= 0
    pb["vals"][0][:] = 0.0
    pb["vals"][0][pb["mesh"].boundaries["top"], code.fields["velocity_x"].index] = 1.0          # the lid
    wg = code.residual_names().index("d_integral_integral_pressure")

    def bordered_newton(pieces):
        U, lam, hist = np.zeros(n), 0.0, []
        for _ in range(10):
            r, J, b, c, g = pieces(U, lam)
            hist.append(max(float(np.abs(r).max()), abs(g)))
            if hist[-1] < 1e-10:
                break
            K = bmat([[J, csr_matrix(b[:, None])], [csr_matrix(c[None, :]), None]]).tocsc()
            d = splu(K).solve(np.concatenate([r, [g]]))
            U, lam = U - d[:n], lam - d[n]
        return U, lam, hist

    op = make_oracle(pb)

    def cpu(U, lam):
        v = pb["vals"][0].copy()
        v[m] = U[eq[m]]
        op.update_values(0, v)
        op.set_params([lam])
        r, mats = op.assemble(flag=1)
        b, _ = op.assemble(which=0, param=0, flag=0)
        c, _ = op.assemble(which=wg, flag=0)
        return r, csr_to_sorted(n, *mats[0]), b, c, op.evaluate_integral_expressions()["integral_pressure"]
    U_ref, lam_ref, hist_ref = bordered_newton(cpu)

    asm = make_gpu(pb)

    def gpu(U, lam):
        asm.set_dofs(U)
        asm.set_parameters(lambda_pressure=lam)
        asm.assemble(flag=1)
        r, jac, _ = asm.fetch(True, False)
        asm.assemble(flag=0, parameter="lambda_pressure")
        b, _, _ = asm.fetch(False, False)
        c = asm.integral_gradient("integral_pressure")
        return r, csr_matrix((jac, asm.indices, asm.indptr), shape=(n, n)), b, c, asm.evaluate_integral_expressions()["integral_pressure"]
    U, lam, hist = bordered_newton(gpu)
    assert hist_ref[-1] < 1e-10 and len(hist) == len(hist_ref) <= 8
    assert np.abs(U - U_ref).max() <= 1e-9 * np.abs(U_ref).max() and abs(lam - lam_ref) <= 1e-9 * max(abs(lam_ref), 1e-3)
    r, J, b, c, g = gpu(U, lam)
    assert np.array_equal(b, c) or np.abs(b - c).max() <= 1e-15 * np.abs(c).max()             # dR/d(lambda) IS the constraint's gradient here
    assert abs(g) < 1e-12
    op.close()
    asm.close()
