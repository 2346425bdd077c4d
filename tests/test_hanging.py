"""Hanging nodes on the CPU: the refined mesh, the oracle's element-level treatment (the reference's hang macros and local
numbering) and the product's global form  J = P^T J_ext P  with its reduction lists -- two independent routes to the same matrix."""
import copy

import numpy as np
import pytest

from problems import compare_matrix, csr_to_sorted, make_oracle, make_problem
from pyoomph_b200.hanging import constraint_lists, extend_numbering, extra_pattern_for_constraints


def apply_constraints_numpy(lists, jac: np.ndarray, res: np.ndarray, diag_value: float = 1.0):
    """what the reduction kernels of pb2_core.cu do with the lists, restated in numpy (test infrastructure: compared with scipy's P^T J P)"""
    jac = jac.copy()
    for k, t in enumerate(lists["target_pos"]):
        a, b = lists["src_start"][k], lists["src_start"][k + 1]
        s = 0.0
        for i in range(a, b):
            s += lists["src_w"][i] * jac[lists["src_pos"][i]]
        jac[t] += s
    jac[lists["clear_pos"]] = 0.0
    jac[lists["diag_pos"]] = diag_value
    if res is not None:
        res = res.copy()
        for k, r in enumerate(lists["res_row"]):
            a, b = lists["res_start"][k], lists["res_start"][k + 1]
            res[r] += float(np.dot(lists["res_w"][a:b], res[lists["res_src"][a:b]]))
        res[lists["virt_rows"]] = 0.0
    return jac, res



def _vals_from_dofs(pb, u):
    """nodal values of a dof vector: equations scattered, pinned values kept, hanging values interpolated from their masters"""
    vals = pb["vals"].copy()
    ne = pb["dofmap"].node_eqn
    live = ne >= 0
    vals[0][live] = u[ne[live]]
    for f in pb["code"].nodal_fields():
        for n, (m, w) in pb["mesh"].hanging.of_space(f.space).items():
            vals[0, n, f.index] = vals[0, m, f.index] @ w
    return vals


def test_refined_mesh_is_conforming():
    pb = make_problem("ns_hang", 6, distortion=0.1)
    mesh = pb["mesh"]
    h = mesh.hanging
    assert len(h.C2) > 20 and len(h.C1) > 10
    for table in (h.C2, h.C1):
        for n, (m, w) in table.items():
            assert abs(w.sum() - 1.0) < 1e-15
            if table is h.C2:       # geometry is interpolated in the C2 space: the hanging node lies on the coarse edge's curve
                assert np.allclose(mesh.node_pos[n], w @ mesh.node_pos[m], atol=1e-15)
    # quadratic weights at the quarter points of a three-node edge
    ws = sorted(tuple(np.round(w, 12)) for _, w in h.C2.values())
    assert set(ws) == {(0.375, 0.75, -0.125), (-0.125, 0.75, 0.375)}
    # hanging values have no equation; every son is a valid element (positive area by the corner cross product)
    ne = pb["dofmap"].node_eqn
    fields = {f.name: f for f in pb["code"].nodal_fields()}
    assert np.all(ne[list(h.C2), fields["velocity_x"].index] < 0) and np.all(ne[list(h.C1), fields["pressure"].index] < 0)
    X = mesh.node_pos[mesh.elem_nodes]
    a, b = X[:, 2] - X[:, 0], X[:, 6] - X[:, 0]
    assert np.all(a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0] > 0)
    # total area = area of the unrefined (distorted) domain: no gaps, no overlaps -- integrate 1 with the oracle
    assert mesh.n_elem == 36 + 3 * int((mesh.parent_element[1:] == mesh.parent_element[:-1]).sum() // 3)


@pytest.mark.parametrize("kind,N,distortion", [("poisson_hang", 6, 0.12), ("ns_hang", 5, 0.1), ("heat3d_hang", 3, 0.08)])
def test_oracle_jacobian_with_hanging_nodes_matches_finite_differences(kind, N, distortion):
    """the reference's own check (src/elements.cpp:5880 analytic vs FD) with the hanging values following their masters"""
    pb = make_problem(kind, N, distortion=distortion)
    n = pb["dofmap"].n_dof
    ne = pb["dofmap"].node_eqn
    u0 = np.zeros(n)
    live = ne >= 0
    u0[ne[live]] = pb["vals"][0][live]
    op = make_oracle(pb)
    r0, mats = op.assemble(flag=1)
    J = csr_to_sorted(n, *mats[0]).toarray()
    rng = np.random.default_rng(3)
    cols = rng.choice(n, size=min(n, 40), replace=False)
    # dofs next to hanging nodes are the interesting ones: masters of hanging values first
    masters = []
    for f in pb["code"].nodal_fields():
        for _, (m, _) in pb["mesh"].hanging.of_space(f.space).items():
            masters += [g for g in ne[m, f.index] if g >= 0]
    cols = np.unique(np.concatenate([cols, np.array(masters[:60], dtype=np.int64)]))
    eps = 1e-6
    worst = 0.0
    for c in cols:
        up, um = u0.copy(), u0.copy()
        up[c] += eps
        um[c] -= eps
        op.update_values(0, _vals_from_dofs(pb, up)[0])
        rp, _ = op.assemble(flag=0)
        op.update_values(0, _vals_from_dofs(pb, um)[0])
        rm, _ = op.assemble(flag=0)
        fd = (rp - rm) / (2 * eps)
        worst = max(worst, np.abs(fd - J[:, c]).max() / max(1.0, np.abs(J[:, c]).max()))
    assert worst < 5e-8, worst
    op.close()


def test_patch_test_with_hanging_nodes():
    """a linear field is reproduced exactly by the constrained space: K u_lin vanishes in every row whose test function does not touch
    the boundary, and constants are in the kernel of every row (Poisson, no Dirichlet values)"""
    pb = make_problem("poisson_hang", 6)
    from pyoomph_b200.meshes import assign_equation_numbers
    pb["dofmap"] = assign_equation_numbers(pb["mesh"], pb["code"], {}, None)
    n = pb["dofmap"].n_dof
    ne = pb["dofmap"].node_eqn[:, 0]
    X = pb["mesh"].node_pos
    op = make_oracle(pb)
    _, mats = op.assemble(flag=1)
    K = csr_to_sorted(n, *mats[0])
    scale = abs(K).max()
    assert abs(K @ np.ones(n)).max() < 1e-13 * scale
    assert abs(K - K.T).max() < 1e-13 * scale
    u = np.zeros(n)
    u[ne[ne >= 0]] = (0.3 + 1.7 * X[:, 0] - 0.9 * X[:, 1])[ne >= 0]
    onb = np.zeros(pb["mesh"].n_node, dtype=bool)
    for b in pb["mesh"].boundaries.values():
        onb[b] = True
    # rows of nodes that share no element with a boundary node
    touches = np.zeros(pb["mesh"].n_node, dtype=bool)
    for en in pb["mesh"].elem_nodes:
        if onb[en].any():
            touches[en] = True
    # masters of hanging nodes inherit their supports
    for nn_, (m, _) in pb["mesh"].hanging.C2.items():
        if touches[nn_]:
            touches[m] = True
    interior = np.array([ne[i] for i in range(pb["mesh"].n_node) if ne[i] >= 0 and not touches[i]])
    assert interior.size > 20
    # 1e-8, not round-off: the reference's Gauss<2,3> knots carry a mistyped digit (relative 8e-9), so its quadrature is not
    # exact for the quadratic integrands; a wrong hanging weight would show at O(1)
    assert abs((K @ u)[interior]).max() < 1e-8 * scale
    op.close()


@pytest.mark.parametrize("kind,N,distortion", [("poisson_hang", 6, 0.12), ("ns_hang", 5, 0.1), ("ns_unsteady_hang", 5, 0.08), ("heat3d_hang", 3, 0.08), ("ale_hang", 4, 0.08)])
def test_global_reduction_equals_the_element_level_treatment(kind, N, distortion):
    """Route 1 (the reference's): hang macros inside the element routine.  Route 2 (the product's): virtual equations for the hanging
    values, the plain element routine, then P^T J_ext P -- by scipy and by the reduction lists the device kernels run."""
    from scipy.sparse import csr_matrix
    pb = make_problem(kind, N, distortion=distortion)
    n = pb["dofmap"].n_dof
    op = make_oracle(pb)
    flag = 2 if pb["unsteady"] else 1
    r_ref, mats_ref = op.assemble(flag=flag)
    op.close()
    ext = extend_numbering(pb["code"], pb["dofmap"], pb["mesh"].hanging)
    assert ext.n_ext > n
    mesh2 = copy.copy(pb["mesh"])
    mesh2.hanging = None                                     # the plain element routine over the extended numbering
    pb2 = dict(pb, mesh=mesh2, dofmap=ext.dofmap)
    op2 = make_oracle(pb2)
    r_ext, mats_ext = op2.assemble(flag=flag)
    op2.close()
    P = ext.prolongation()
    assert np.abs(P.T @ r_ext - r_ref).max() <= 1e-13 * np.abs(r_ref).max()
    for (rs, ci, va), ref in zip(mats_ext, mats_ref):
        Jx = csr_matrix((va, ci, rs), shape=(ext.n_ext, ext.n_ext))
        Jr = (P.T @ Jx @ P).tocsr()
        err, missing = compare_matrix(Jr, csr_to_sorted(n, *ref), tol=1e-13)
        assert err <= 1e-13, err
    # the reduction lists on the extended pattern (+ the extra entries of the master-master couplings)
    from pyoomph_b200.distributed import element_dof_table, structural_pattern
    ed = element_dof_table(pb["code"], mesh2, ext.dofmap, np.arange(mesh2.n_elem))
    ex_r, ex_c = extra_pattern_for_constraints(pb["code"], mesh2, ext)
    assert ex_r.size > 0
    ip, ix = structural_pattern(ed, ext.n_ext, extra=(ex_r, ex_c))
    lists = constraint_lists(ip, ix, ext)
    Jx = csr_matrix((mats_ext[0][2], mats_ext[0][1], mats_ext[0][0]), shape=(ext.n_ext, ext.n_ext))
    S = csr_matrix((np.arange(1, ix.size + 1, dtype=np.float64), ix, ip), shape=(ext.n_ext, ext.n_ext))
    vals = np.zeros(ix.size)
    Jc = Jx.tocoo()
    pos = np.asarray(S[Jc.row, Jc.col]).ravel().astype(np.int64) - 1
    assert np.all(pos >= 0)
    np.add.at(vals, pos, Jc.data)
    red, rres = apply_constraints_numpy(lists, vals, r_ext)
    R = csr_matrix((red, ix, ip), shape=(ext.n_ext, ext.n_ext))
    err, _ = compare_matrix(R[:n, :n].tocsr(), csr_to_sorted(n, *mats_ref[0]), tol=1e-13)
    assert err <= 1e-13
    assert np.abs(rres[:n] - r_ref).max() <= 1e-13 * np.abs(r_ref).max() and np.all(rres[n:] == 0.0)
    # virtual rows / columns: identity
    V = R[n:, :].toarray()
    assert np.array_equal(V[:, n:], np.eye(ext.n_ext - n)) and np.all(V[:, :n] == 0.0) and abs(R[:n, n:]).max() == 0.0


@pytest.mark.parametrize("kind,N", [("poisson_hang", 5), ("ns_hang", 4)])
def test_hanging_macros_of_the_reference_header_give_identical_matrices(kind, N):
    """the same generated C against /root/reference/src/jitbridge.h + jitbridge_hang.h (the real hang macros and HangInfo structs)"""
    import os
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference headers not present on this box")
    pb = make_problem(kind, N, distortion=0.1)
    a = make_oracle(pb)
    b = make_oracle(pb, reference_headers=True)
    ra, ma = a.assemble(flag=1)
    rb, mb = b.assemble(flag=1)
    assert np.array_equal(ra, rb)
    for x, y in zip(ma[0], mb[0]):
        assert np.array_equal(x, y)
    a.close()
    b.close()


def test_host_side_of_the_hanging_node_assembly():
    """HangingNodeAssembly without a GPU (pattern-only engine problem): the extended pattern holds every target of the reduction, the
    handed-out pattern is the n_dof x n_dof block, and it contains the oracle's (value-dependent) pattern"""
    from pyoomph_b200.hanging import HangingNodeAssembly
    pb = make_problem("ns_hang", 5, distortion=0.1)
    asm = HangingNodeAssembly(pb["code"], pb["mesh"], pb["dofmap"], name=pb["code"].name, device=-1)
    n = pb["dofmap"].n_dof
    assert asm.n_dof == n and asm.indptr.size == n + 1 and asm.indices.max() < n
    L = asm.lists
    rows_of_pos = np.repeat(np.arange(asm.n_ext), np.diff(asm.asm.indptr))
    assert np.all(rows_of_pos[L["target_pos"]] < n) and np.all(asm.asm.indices[L["target_pos"]] < n)          # targets are real entries
    src_rows, src_cols = rows_of_pos[L["src_pos"]], asm.asm.indices[L["src_pos"]]
    assert np.all((src_rows >= n) | (src_cols >= n))                                                        # sources are virtual: never a target
    assert np.all(np.diff(L["src_start"]) > 0) and np.all(np.diff(L["target_pos"]) > 0)
    op = make_oracle(pb)
    _, mats = op.assemble(flag=1)
    op.close()
    ones = csr_to_sorted(n, asm.indptr, asm.indices, np.ones(asm.nnz))
    _, missing = compare_matrix(ones, csr_to_sorted(n, *mats[0]))
    assert missing == 0
    asm.close()


def test_octree_refinement_is_conforming():
    """3D: faces AND edges between refined and unrefined elements are constrained (an edge is shared by up to four elements); masters never
    hang themselves; weights are the tensor products of the quadratic 1D weights with the zeros dropped (3 masters on edges and face
    mid-lines, 9 inside faces); hanging nodes lie on the coarse entity's surface; the shared face of two refined neighbours has none."""
    pb = make_problem("heat3d_hang", 3, distortion=0.08)
    mesh = pb["mesh"]
    h = mesh.hanging.C2
    assert len(h) > 100 and not mesh.hanging.C1
    masters_all = set(int(x) for m, _ in h.values() for x in m)
    sizes = set()
    for n, (m, w) in h.items():
        assert abs(w.sum() - 1.0) < 1e-14 and n not in masters_all
        assert np.allclose(mesh.node_pos[n], w @ mesh.node_pos[m], atol=1e-15)
        sizes.add(len(m))
    assert sizes == {3, 9}
    assert np.all(pb["dofmap"].node_eqn[list(h), 0] < 0)
    # the patch test in 3D: constants are in the kernel of the stiffness part (J - weight * M) of every row
    from pyoomph_b200.meshes import assign_equation_numbers
    pb["dofmap"] = assign_equation_numbers(mesh, pb["code"], {}, None)
    op = make_oracle(pb)
    n = pb["dofmap"].n_dof
    _, mats = op.assemble(flag=2)
    J, M = csr_to_sorted(n, *mats[0]), csr_to_sorted(n, *mats[1])
    one = np.ones(n)
    ratio = (J @ one) / (M @ one)                   # J 1 = w0 M 1 (the time weight) when K 1 = 0
    assert np.abs(ratio - ratio[0]).max() < 1e-9 * abs(ratio[0])
    op.close()


def test_oracle_jacobian_with_hanging_positions_matches_finite_differences():
    """hanging nodes on a MOVING mesh: the position dofs of the masters move the hanging nodes (hanginfo_Pos, src/elements.cpp:820-862);
    analytic Jacobian columns of master positions and of ordinary dofs against central differences with the hanging nodes re-placed"""
    pb = make_problem("ale_hang", 5, distortion=0.08)
    n = pb["dofmap"].n_dof
    eq, peq = pb["dofmap"].node_eqn, pb["dofmap"].pos_eqn
    hang = pb["mesh"].hanging
    assert np.all(peq[list(hang.C2)] < 0)                      # hanging positions are no dofs
    op = make_oracle(pb)
    _, mats = op.assemble(flag=1)
    A = csr_to_sorted(n, *mats[0]).toarray()

    def set_state(U):
        v, x = pb["vals"][0].copy(), pb["pos_hist"][0].copy()
        v[eq >= 0] = U[eq[eq >= 0]]
        x[peq >= 0] = U[peq[peq >= 0]]
        for f in pb["code"].nodal_fields():
            for nn_, (ms, w) in hang.of_space(f.space).items():
                v[nn_, f.index] = v[ms, f.index] @ w
        for nn_, (ms, w) in hang.C2.items():
            x[nn_] = w @ x[ms]
        op.update_values(0, v, x)
    U0 = np.zeros(n)
    U0[eq[eq >= 0]] = pb["vals"][0][eq >= 0]
    U0[peq[peq >= 0]] = pb["pos_hist"][0][peq >= 0]
    masters = [g for _, (ms, _) in hang.C2.items() for g in peq[ms].ravel() if g >= 0]
    cols = np.unique(np.array(masters[:40] + list(np.random.default_rng(0).choice(n, 20, replace=False))))
    eps, worst = 1e-6, 0.0
    for c in cols:
        up, um = U0.copy(), U0.copy()
        up[c] += eps
        um[c] -= eps
        set_state(up)
        rp, _ = op.assemble(flag=0)
        set_state(um)
        rm, _ = op.assemble(flag=0)
        worst = max(worst, np.abs((rp - rm) / (2 * eps) - A[:, c]).max() / np.abs(A).max())
    assert worst < 5e-8, worst
    op.close()
