"""Pins of the CPU oracle against the reference's own code (needs /root/reference; skipped on the GPU box).

1. Gauss tables and 1D Lagrange polynomials of oracle/driver.c and of the CUDA emitter == the oomph-lib sources compiled
   into oracle/_ref/libref_shim.so (bit-exact).
2. The generated-format plugin + driver compiled against the reference's jitbridge.h / jitbridge_hang.h gives bit-identical
   element matrices and assembled CSR as the build against the oracle's restated headers.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from problems import csr_to_sorted, make_oracle, make_problem

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE = os.path.join(os.path.dirname(HERE), "oracle")
HAVE_REF = os.path.isdir("/root/reference/src")
pytestmark = pytest.mark.skipif(not HAVE_REF, reason="reference tree not present")


@pytest.fixture(scope="module")
def shim():
    subprocess.run(["make", "-C", ORACLE], check=True, capture_output=True)
    return ctypes.CDLL(os.path.join(ORACLE, "_ref", "libref_shim.so"))


def test_gauss_tables_bit_exact(shim):
    from oracle import build_plugin
    from pyoomph_b200.cuda_emitter import gauss_rule
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    for dim, npt in ((2, 9), (3, 27)):
        kn, w = gauss_rule(dim)
        for ipt in range(npt):
            k_ref, w_ref = (ctypes.c_double * 3)(), ctypes.c_double()
            k_or, w_or = (ctypes.c_double * 3)(), ctypes.c_double()
            shim.ref_gauss(dim, ipt, k_ref, ctypes.byref(w_ref))
            drv.oracle_gauss(dim, ipt, k_or, ctypes.byref(w_or))
            assert list(k_ref)[:dim] == list(k_or)[:dim] == list(kn[ipt])
            assert w_ref.value == w_or.value == w[ipt]
    # the mistyped literal really is there
    k = (ctypes.c_double * 3)(); w_ = ctypes.c_double()
    shim.ref_gauss(2, 8, k, ctypes.byref(w_))
    assert k[0] == 0.774596662941483 and k[0] != 0.774596669241483


def test_lagrange_polynomials_bit_exact(shim):
    from oracle import build_plugin
    from pyoomph_b200.cuda_emitter import _lag
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    rng = np.random.default_rng(0)
    pts = list(rng.uniform(-1, 1, 20)) + [-0.774596669241483, 0.0, 0.774596662941483, 0.77459666924148]
    for order in (2, 3):
        for s in pts:
            p, d = (ctypes.c_double * 3)(), (ctypes.c_double * 3)()
            shim.ref_lagrange(order, ctypes.c_double(s), p, d)
            P, D = _lag(order, float(s))
            assert list(p)[:order] == P and list(d)[:order] == D
            # oracle: 2D tensor product at (s, s) contains the 1D values as products
            psi, dpsi = (ctypes.c_double * 9)(), (ctypes.c_double * 18)()
            sv = (ctypes.c_double * 2)(s, s)
            drv.oracle_dshape_local(2, order, sv, psi, dpsi)
            # (the driver is built like the reference's JIT code, -O3 -march=native, so gcc may contract 1.0-s*s into an
            #  FMA: the last bit can differ from the uncontracted shim; the reference's own bits depend on its build flags)
            for i in range(order):
                for j in range(order):
                    assert abs(psi[i * order + j] - p[i] * p[j]) <= 4e-16
                    assert abs(dpsi[(i * order + j) * 2 + 0] - p[i] * d[j]) <= 8e-16
                    assert abs(dpsi[(i * order + j) * 2 + 1] - d[i] * p[j]) <= 8e-16


@pytest.mark.parametrize("kind,N", [("ns_unsteady", 3), ("ale", 3), ("heat3d", 2), ("ale_axi_obs", 3)])
def test_reference_headers_give_identical_results(kind, N):
    pb = make_problem(kind, N)
    out = []
    for ref in (False, True):
        op = make_oracle(pb, reference_headers=ref)
        e = op.element(pb["mesh"].n_elem // 2, flag=2)
        r, m = op.assemble(flag=2)
        obs = (np.array(list(op.evaluate_integral_expressions().values())),) if pb["code"].integral_expressions else ()
        out.append(e + (r,) + tuple(x for mat in m for x in mat) + obs)     # EvalIntegralExpression through jitbridge.h:469 too
        op.close()
    for a, b in zip(*out):
        assert np.array_equal(a, b)


# ======================================================================================================================================
# Pins against the reference's COMPILED code: oracle/_ref/liboomph_ref.so = the vendored oomph-lib `generic` library (all 50 translation
# units of src/thirdparty/oomph-lib/include, no MPI) + pyoomph's src/timestepper.cpp and src/hessian_tensor.cpp (oracle/Makefile,
# oracle/ref_oomph.cpp).  These are the reference's own QElement / SolidNode / Mesh / Problem / MultiTimeStepper / SparseRank3Tensor.
# ======================================================================================================================================
ELEM_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double),
                           ctypes.POINTER(ctypes.c_double), ctypes.c_int)


@pytest.fixture(scope="module")
def oomph():
    subprocess.run(["make", "-C", ORACLE, "-j8"], check=True, capture_output=True)
    L = ctypes.CDLL(os.path.join(ORACLE, "_ref", "liboomph_ref.so"))
    L.ref_problem_create.restype = ctypes.c_void_p
    L.ref_problem_ndof.restype = ctypes.c_long
    L.ref_problem_assemble.restype = ctypes.c_long
    L.ref_element_geometry.restype = ctypes.c_double
    L.ref_rank3_product.restype = ctypes.c_long
    return L


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ref_problem(L, pb):
    """the reference's own Problem for `pb`: SolidNodes in mesh node order with the pins of the dof map, SolidQElement<dim,3> in element order"""
    mesh, dm = pb["mesh"], pb["dofmap"]
    val_pinned = np.ascontiguousarray((dm.node_eqn < 0).astype(np.uint8))
    pos_pinned = None if dm.pos_eqn is None else np.ascontiguousarray((dm.pos_eqn < 0).astype(np.uint8))
    pos = np.ascontiguousarray(mesh.node_pos, dtype=np.float64)
    en = np.ascontiguousarray(mesh.elem_nodes, dtype=np.int32)
    h = L.ref_problem_create(mesh.dim, ctypes.c_long(mesh.n_node), dm.node_eqn.shape[1], _dp(pos), ctypes.c_long(mesh.n_elem), _ip(en),
                             val_pinned.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                             None if pos_pinned is None else pos_pinned.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)))
    return ctypes.c_void_p(h), (pos, en, val_pinned, pos_pinned)


def test_qelement_shape_functions_bit_exact(oomph):
    """QElement<DIM,3>::dshape_local and QElement<DIM,2>::dshape_local (compiled Qelements.h) against the emitter's tables (bit-exact:
    same products in the same order) and the oracle's restated oracle_dshape_local, at every Gauss point and at random points; the
    reference's node order (first local coordinate fastest) is part of the comparison."""
    from oracle import build_plugin
    from pyoomph_b200.cuda_emitter import gauss_rule, shape_tables
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    rng = np.random.default_rng(3)
    for dim in (2, 3):
        kn, w = gauss_rule(dim)
        # default integration scheme of the compiled element = the tables of the emitter (incl. the mistyped Gauss<2,3> knots)
        k_ref, w_ref = (ctypes.c_double * 3)(), ctypes.c_double()
        assert oomph.ref_element_integral(dim, 0, k_ref, ctypes.byref(w_ref)) == len(kn)
        for ipt in range(len(kn)):
            oomph.ref_element_integral(dim, ipt, k_ref, ctypes.byref(w_ref))
            assert list(k_ref)[:dim] == list(kn[ipt]) and w_ref.value == w[ipt]
        for order in (3, 2):
            n = order ** dim
            psi_t, dpsi_t = shape_tables(dim, order, kn)
            pts = [np.array(k, dtype=np.float64) for k in kn] + [rng.uniform(-1, 1, dim) for _ in range(10)]
            for ip, s in enumerate(pts):
                psi, dpsi = np.zeros(n), np.zeros((n, dim))
                assert oomph.ref_qshape(dim, order, _dp(s), _dp(psi), _dp(dpsi)) == n
                if ip < len(kn):
                    assert np.array_equal(psi, np.array(psi_t[ip])) and np.array_equal(dpsi, np.array(dpsi_t[ip]))
                po, do = np.zeros(n), np.zeros((n, dim))
                drv.oracle_dshape_local(dim, order, _dp(s), _dp(po), _dp(do))
                # the driver is built with the reference's JIT flags (-O3 -march=native): FMA contraction may move the last bit
                assert np.abs(po - psi).max() <= 4e-16 and np.abs(do - dpsi).max() <= 1e-15
                assert abs(psi.sum() - 1.0) <= 1e-14


@pytest.mark.parametrize("kind,N,unstructured", [("ns", 5, False), ("ale", 4, False), ("heat3d", 2, False), ("ns_axi_swirl", 4, False),
                                                 ("poisson", 6, True), ("ale_axi", 4, True), ("ns_unsteady", 3, True)])
def test_equation_numbering_and_local_order_match_oomph(oomph, kind, N, unstructured):
    """Problem::assign_eqn_numbers of the compiled oomph-lib (Mesh::assign_global_eqn_numbers, mesh.cc:686-708; SolidNode positions
    first, nodes.cc:3652-3659) on the same nodes, pins and elements == meshes.assign_equation_numbers, array_equal; and the local
    equation order of every element (GeneralisedElement::assign_local_eqn_numbers: nodal values, then solid positions,
    elements.cc:694-699) == the order the oracle's fill_element_info restatement uses."""
    pb = make_problem(kind, N, unstructured=unstructured)
    mesh, dm = pb["mesh"], pb["dofmap"]
    h, keep = _ref_problem(oomph, pb)
    assert oomph.ref_problem_ndof(h) == dm.n_dof
    node_eqn = np.zeros_like(dm.node_eqn, dtype=np.int32)
    pos_eqn = np.zeros((mesh.n_node, mesh.dim), dtype=np.int32)
    oomph.ref_problem_numbering(h, _ip(node_eqn), _ip(pos_eqn))
    assert np.array_equal(node_eqn, dm.node_eqn)
    if dm.pos_eqn is not None:
        assert np.array_equal(pos_eqn, dm.pos_eqn)
    else:
        assert (pos_eqn == -1).all()
    op = make_oracle(pb)
    for e in range(mesh.n_elem):
        out = np.zeros(256, dtype=np.int32)
        n = oomph.ref_element_local_eqns(h, ctypes.c_long(e), _ip(out))
        _, _, _, eq = op.element(e, flag=0)
        assert n == eq.size and np.array_equal(out[:n], eq), (e, out[:n], eq)
    op.close()
    oomph.ref_problem_free(h)


@pytest.mark.parametrize("kind,N,distortion", [("ns", 3, 0.15), ("heat3d", 2, 0.1), ("ale", 3, 0.1)])
def test_eulerian_geometry_matches_oomph(oomph, kind, N, distortion):
    """FiniteElement::dshape_eulerian_at_knot of the compiled oomph-lib (Jacobian of the mapping, d psi/d x by its inverse, elements.cc)
    against the oracle's restated fill_shape_info_at_s (metric-tensor form, src/elements.cpp:3604-3703) on distorted meshes: the two
    are different formulas for the same numbers, so they agree to rounding (1e-13 of the row scale), not bitwise."""
    pb = make_problem(kind, N, distortion=distortion)
    mesh = pb["mesh"]
    h, keep = _ref_problem(oomph, pb)
    pb2 = dict(pb, pos_hist=None)                       # geometry at the mesh's own node positions, like the oomph nodes above
    op = make_oracle(pb2)
    nn, dim = mesh.elem_nodes.shape[1], mesh.dim
    nipt = 9 if dim == 2 else 27
    for e in range(0, mesh.n_elem, max(1, mesh.n_elem // 5)):
        for ipt in range(nipt):
            psi, dpsidx, w = np.zeros(nn), np.zeros((nn, dim)), ctypes.c_double()
            J = oomph.ref_element_geometry(h, ctypes.c_long(e), ipt, _dp(psi), _dp(dpsidx), ctypes.byref(w))
            wts, sh, dx, dX, _, _ = op.point_shapes(e, ipt, flag=0)
            assert abs(wts[0] - w.value * J) <= 1e-13 * abs(w.value * J)
            assert np.abs(sh - psi).max() <= 4e-16
            assert np.abs(dx - dpsidx).max() <= 1e-13 * np.abs(dpsidx).max()
    op.close()
    oomph.ref_problem_free(h)


def _assemble_through_oomph(L, h, op, n_dof, method, flag=1):
    """Problem::get_jacobian of the compiled oomph-lib with the oracle's generated routine as the element: returns
    (row_start, column_index, values, residuals) and whether every element's local order was the identity permutation"""
    identity = [True]

    def cb(ctx, e, n, geqn, R, J, fl):
        r, j, _, eq = op.element(e, flag=1 if fl else 0)
        g = np.ctypeslib.as_array(geqn, shape=(n,))
        if not (eq.size == n and np.array_equal(eq, g)):
            identity[0] = False
            return
        np.ctypeslib.as_array(R, shape=(n,))[:] = r
        if fl:
            np.ctypeslib.as_array(J, shape=(n * n,))[:] = j.ravel()
    cbf = ELEM_CB(cb)
    nnz = L.ref_problem_assemble(h, cbf, None, method)
    rs, ci, va, res = np.zeros(n_dof + 1, dtype=np.int32), np.zeros(nnz, dtype=np.int32), np.zeros(nnz), np.zeros(n_dof)
    L.ref_problem_result(h, _ip(rs), _ip(ci), _dp(va), _dp(res))
    return rs, ci, va, res, identity[0]


@pytest.mark.parametrize("kind,N,unstructured", [("ns", 4, False), ("ale", 3, False), ("heat3d", 2, False), ("poisson", 5, True), ("ns_unsteady", 4, True)])
def test_csr_assembly_byte_identical_with_oomph(oomph, kind, N, unstructured):
    """The oracle's element loop + vectors_of_pairs CSR build (oracle/driver.c) against oomph-lib's own
    Problem::sparse_assemble_row_or_column_compressed_with_vectors_of_pairs (problem.cc:5332-5666) fed with the SAME element matrices:
    row_start, column_index (unsorted, first-touch order), values and residuals must be BYTE-identical (same additions in the same
    order, same |v| > 0.0 drop rule).  The "maps" method (ascending columns; what pyoomph's multi-assembly and the GPU pattern use,
    src/problem.cpp:2200-2274) must equal the column-sorted oracle matrix, values byte-identical as well."""
    pb = make_problem(kind, N, unstructured=unstructured)
    n = pb["dofmap"].n_dof
    h, keep = _ref_problem(oomph, pb)
    op = make_oracle(pb)
    r_or, mats = op.assemble(flag=1)
    rs_o, ci_o, va_o = mats[0]
    rs, ci, va, res, ident = _assemble_through_oomph(oomph, h, op, n, 0)
    assert ident
    assert np.array_equal(rs, rs_o) and np.array_equal(ci, ci_o)
    assert va.tobytes() == va_o.tobytes() and res.tobytes() == r_or.tobytes()
    rs_m, ci_m, va_m, res_m, _ = _assemble_through_oomph(oomph, h, op, n, 2)      # maps: ascending columns
    B = csr_to_sorted(n, rs_o, ci_o, va_o)
    assert np.array_equal(rs_m, B.indptr) and np.array_equal(ci_m, B.indices)
    assert va_m.tobytes() == B.data.tobytes() and res_m.tobytes() == r_or.tobytes()
    op.close()
    oomph.ref_problem_free(h)


def test_timestepper_weights_bit_exact(oomph):
    """pyoomph::MultiTimeStepper::set_weights (compiled src/timestepper.cpp) against bdf_weights / newmark2_weights of the product
    host code and the oracle's C restatement: bit-exact for a range of (dt, dtprev)."""
    from oracle import build_plugin
    from pyoomph_b200.assembly import bdf_weights, newmark2_weights
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    rng = np.random.default_rng(11)
    for dt, dtp in [(0.01, 0.012), (1.0, 1.0), (3e-4, 7.1e-3)] + [tuple(rng.uniform(1e-4, 2.0, 2)) for _ in range(20)]:
        b1, b2, n1, n2 = np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(7)
        nt = oomph.ref_timestepper_weights(ctypes.c_double(dt), ctypes.c_double(dtp), _dp(b1), _dp(b2), _dp(n1), _dp(n2))
        assert nt == 5                                   # NSTEPS + 3 storage levels (non-adaptive), src/timestepper.hpp:49-50
        w1, w2 = bdf_weights(dt, dtp)
        m1, m2 = newmark2_weights(dt)
        assert np.array_equal(w1, b1) and np.array_equal(w2, b2) and np.array_equal(m1, n1) and np.array_equal(m2, n2)
        o1, o2, p1, p2 = np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(7)
        drv.oracle_bdf_weights(ctypes.c_double(dt), ctypes.c_double(dtp), _dp(o1), _dp(o2))
        drv.oracle_newmark2_weights(ctypes.c_double(dt), ctypes.c_double(0.5), ctypes.c_double(0.5), _dp(p1), _dp(p2))
        for a, b in ((o1, b1), (o2, b2), (p1, n1), (p2, n2)):
            assert np.abs(a - b).max() <= 4e-16 * max(1.0, np.abs(b).max())     # -O3 -march=native build of the driver: last-bit freedom


def test_sparse_rank3_tensor_matches_compiled_reference(oomph):
    """pyoomph_b200.hessian_tensor.SparseRank3Tensor against the compiled pyoomph::SparseRank3Tensor (src/hessian_tensor.cpp): same
    CSR pattern of M_ij = T_ijk v_k from finalize_for_vector_product, same product values; repeated (i, j, k) accumulate."""
    from pyoomph_b200.hessian_tensor import SparseRank3Tensor
    rng = np.random.default_rng(5)
    n, ne = 23, 900
    ii, jj, kk = (rng.integers(0, n, ne).astype(np.int32) for _ in range(3))
    ii[:50], jj[:50], kk[:50] = ii[50:100], jj[50:100], kk[50:100]          # repeated entries
    vv = rng.uniform(-1, 1, ne)
    vec = rng.uniform(-1, 1, n)
    T = SparseRank3Tensor(n, False)
    T.accumulate(ii, jj, kk, vv)
    ci, rs = T.finalize_for_vector_product()
    vals = T.right_vector_mult(vec)
    ci_r, rs_r, va_r = np.zeros(ne, dtype=np.int32), np.zeros(n + 1, dtype=np.int32), np.zeros(ne)
    nnz = oomph.ref_rank3_product(n, 0, ctypes.c_long(ne), _ip(ii), _ip(jj), _ip(kk), _dp(vv), _dp(vec), _ip(ci_r), _ip(rs_r), _dp(va_r))
    assert nnz == vals.size and np.array_equal(rs, rs_r) and np.array_equal(ci, ci_r[:nnz])
    assert np.abs(vals - va_r[:nnz]).max() <= 1e-14 * np.abs(va_r).max()
    ent = T.get_entries()
    assert len(ent) == len({(a, b, c) for a, b, c in zip(ii.tolist(), jj.tolist(), kk.tolist())}) and ent == sorted(ent, key=lambda t: t[:3])
    with pytest.raises(RuntimeError):
        SparseRank3Tensor(n).right_vector_mult(vec)


def test_triangle_shape_functions_and_tgauss_bit_exact(oomph):
    """TElement<2,3> / TElement<2,2>::dshape_local, local_coordinate_of_node and TGauss<2,3> of the compiled oomph-lib against the
    emitter's triangle tables (bit-exact) and the oracle's restatement."""
    from oracle import build_plugin
    from pyoomph_b200.cuda_emitter import TRIANGLE_NODE_COORDS, tgauss_rule, triangle_shape_tables
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    kn, w = tgauss_rule()
    k_ref, w_ref = (ctypes.c_double * 2)(), ctypes.c_double()
    assert oomph.ref_tgauss(0, k_ref, ctypes.byref(w_ref)) == 7 == len(kn)
    for ipt in range(7):
        oomph.ref_tgauss(ipt, k_ref, ctypes.byref(w_ref))
        assert list(k_ref) == list(kn[ipt]) and w_ref.value == w[ipt]
        ko, wo = (ctypes.c_double * 3)(), ctypes.c_double()
        drv.oracle_gauss_tri(ipt, ko, ctypes.byref(wo))
        assert list(ko)[:2] == list(kn[ipt]) and wo.value == w[ipt]
    rng = np.random.default_rng(8)
    pts = [np.array(k) for k in kn] + [np.array(c) for c in TRIANGLE_NODE_COORDS] + [rng.dirichlet((1, 1, 1))[:2] for _ in range(10)]
    for order, n in ((3, 6), (2, 3)):
        for s in pts:
            s = np.ascontiguousarray(s, dtype=np.float64)
            psi, dpsi, ns = np.zeros(n), np.zeros((n, 2)), np.zeros((n, 2))
            assert oomph.ref_tshape(order, _dp(s), _dp(psi), _dp(dpsi), _dp(ns)) == n
            pt, dt = triangle_shape_tables(order, [tuple(s)])
            assert np.array_equal(psi, np.array(pt[0])) and np.array_equal(dpsi, np.array(dt[0]))
            po, do = np.zeros(n), np.zeros((n, 2))
            drv.oracle_dshape_local_tri(order, _dp(s), _dp(po), _dp(do))
            assert np.abs(po - psi).max() <= 4e-16 and np.abs(do - dpsi).max() <= 1e-15
            if order == 3:
                assert np.array_equal(ns, np.array(TRIANGLE_NODE_COORDS))


def test_tetrahedron_shape_functions_and_tgauss_bit_exact(oomph):
    """TElement<3,3> / TElement<3,2>::dshape_local, local_coordinate_of_node and TGauss<3,3> (11 points, one negative weight) of the
    compiled oomph-lib against the emitter's tetrahedron tables (bit-exact) and the oracle's restatement."""
    from oracle import build_plugin
    from pyoomph_b200.cuda_emitter import TETRA_NODE_COORDS, tetra_shape_tables, tgauss3_rule
    pb = make_problem("poisson", 2)
    drv = ctypes.CDLL(build_plugin(pb["code"], pb["code"].name))
    kn, w = tgauss3_rule()
    k_ref, w_ref = (ctypes.c_double * 3)(), ctypes.c_double()
    assert oomph.ref_tgauss3(0, k_ref, ctypes.byref(w_ref)) == 11 == len(kn)
    for ipt in range(11):
        oomph.ref_tgauss3(ipt, k_ref, ctypes.byref(w_ref))
        assert list(k_ref) == list(kn[ipt]) and w_ref.value == w[ipt]
        ko, wo = (ctypes.c_double * 3)(), ctypes.c_double()
        drv.oracle_gauss_tet(ipt, ko, ctypes.byref(wo))
        assert list(ko) == list(kn[ipt]) and wo.value == w[ipt]
    rng = np.random.default_rng(9)
    pts = [np.array(k) for k in kn] + [np.array(c) for c in TETRA_NODE_COORDS] + [rng.dirichlet((1, 1, 1, 1))[:3] for _ in range(10)]
    for order, n in ((3, 10), (2, 4)):
        for s in pts:
            s = np.ascontiguousarray(s, dtype=np.float64)
            psi, dpsi, ns = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3))
            assert oomph.ref_tshape3(order, _dp(s), _dp(psi), _dp(dpsi), _dp(ns)) == n
            pt, dt = tetra_shape_tables(order, [tuple(s)])
            assert np.array_equal(psi, np.array(pt[0])) and np.array_equal(dpsi, np.array(dt[0]))
            po, do = np.zeros(n), np.zeros((n, 3))
            drv.oracle_dshape_local_tet(order, _dp(s), _dp(po), _dp(do))
            assert np.array_equal(po, psi) and np.array_equal(do, dpsi)
            if order == 3:
                assert np.array_equal(ns, np.array(TETRA_NODE_COORDS))
