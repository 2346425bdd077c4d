"""CPU tests of the host side: mesh/equation ordering rules, symbolic coefficient form, CUDA emission (cross-compiled
with nvcc for sm_100a, no GPU needed), compiler registry, and that the C-ABI library loads and exports every symbol
include/*.h declares (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import sympy as sp

from problems import make_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mesh_node_and_element_order_follow_reference_rules():
    from pyoomph_b200.meshes import CuboidBrickMesh, RectangularQuadMesh
    m = RectangularQuadMesh(2)
    # simplemeshes.py:224-240 + meshtemplate.cpp:401-409: vertices first, first-touch order; local layout
    # [n00,m_b,n10,m_l,c,m_r,n01,m_t,n11]
    assert m.elem_nodes[0].tolist() == [0, 9, 1, 10, 11, 12, 2, 13, 3]
    assert m.elem_nodes[1].tolist() == [2, 13, 3, 14, 15, 16, 4, 17, 5]      # iy inner: shares the top edge of element 0
    assert m.elem_nodes[2].tolist() == [1, 18, 6, 12, 19, 20, 3, 21, 7]      # ix outer: shares the right edge of element 0
    assert m.n_node == 25 and np.all(m.is_vertex()[:9]) and not np.any(m.is_vertex()[9:])
    np.testing.assert_allclose(m.node_pos[11], [0.25, 0.25])
    b = CuboidBrickMesh(2)
    assert b.n_node == 125 and b.elem_nodes.shape == (8, 27)
    # every element's nodes sit on its 3x3x3 lattice in oomph order (first coordinate fastest)
    for e in range(8):
        lat = b.node_lattice[b.elem_nodes[e]]
        lat = lat - lat[0]
        exp = np.array([[k % 3, (k // 3) % 3, k // 9] for k in range(27)])
        assert np.array_equal(lat, exp)
    # meshtemplate.cpp:583-617 creation order inside the first brick: edge mids 1,3 then face centre 4 ...
    assert b.elem_nodes[0][[1, 3, 4, 5, 7, 9]].tolist() == [27, 28, 29, 30, 31, 32]


def test_equation_numbering_order():
    from pyoomph_b200.meshes import assign_equation_numbers
    pb = make_problem("ale", 2)
    dm, code, mesh = pb["dofmap"], pb["code"], pb["mesh"]
    # node by node; positions first, then values by index; pinned skipped; C1 slots on non-vertex nodes are dummies
    flat = np.concatenate([dm.pos_eqn, dm.node_eqn], axis=1).ravel()
    nz = flat[flat >= 0]
    assert np.array_equal(nz, np.arange(dm.n_dof))
    p_idx = code.fields["pressure"].index
    assert np.all(dm.node_eqn[~mesh.is_vertex(), p_idx] == -1)
    assert np.all(dm.pos_eqn[mesh.boundaries["left"], 0] == -1)


def test_coefficient_form_matches_direct_differentiation():
    """J coefficients of the product path == derivative of the residual coefficient w.r.t. the interpolated atom."""
    pb = make_problem("ns_unsteady", 2)
    code = pb["code"]
    form = code.derive("")
    assert {s.field for s in form.slots} == {"velocity_x", "velocity_y", "pressure"}
    for (si, G, a), c in form.J.items():
        if a == "d0":
            at = [s for s, info in code._atom_syms.items() if info.field == G and info.deriv == "d0" and info.dt_order == 0]
            direct = sum(sp.diff(form.R[si], s) for s in at)
            dt = [s for s, info in code._atom_syms.items() if info.field == G and info.deriv == "d0" and info.dt_order == 1]
            w = sp.Symbol("W__BDF2_degr__1", real=True)
            direct += sum(w * sp.diff(form.R[si], s) for s in dt)
            assert sp.simplify(c - direct) == 0
    # mass matrix: only d/d(partial_t u) terms, velocity test x velocity unknown
    assert {(form.slots[si].field, G) for (si, G, a) in form.M} == {("velocity_x", "velocity_x"), ("velocity_y", "velocity_y")}
    assert code.history_levels() == 3 and code.max_dt_order() == 1


def test_moving_mesh_columns_present():
    code = make_problem("ale", 2)["code"]
    form = code.derive("")
    cols = {G for (_, G, _) in form.J}
    assert {"coordinate_x", "coordinate_y", "velocity_x", "velocity_y", "pressure"} <= cols
    assert form.uses_dX and form.uses_dx
    assert code.dof_layout()[:5] == [("coordinate_x", 0), ("coordinate_y", 0), ("velocity_x", 0), ("velocity_y", 0), ("pressure", 0)]


def test_compiler_registry_and_cross_compile():
    from pyoomph_b200.ccompiler import BaseCCompiler, CudaCCompiler, get_ccompiler
    assert "cuda" in BaseCCompiler._registry
    with pytest.raises(RuntimeError):
        BaseCCompiler.factory_compiler("tcc")          # no CPU compiler is registered: no CPU fallback
    cc = get_ccompiler("cuda")
    assert isinstance(cc, CudaCCompiler)
    from pyoomph_b200.cuda_emitter import CudaEmitter
    code = make_problem("poisson", 2)["code"]
    em = CudaEmitter(code, "poisson")
    src = em.emit()
    assert "JIT_ELEMENT_init_cuda" in src and "bar.sync" in src and "red.global.add.f64" in src
    so = cc.compile_code(src, "poisson")
    lib = ctypes.CDLL(so)
    assert hasattr(lib, "JIT_ELEMENT_init_cuda")
    # B_el of SURVEY 8(d): Poisson Q9 steady = 1296 B
    assert em.algorithmic_bytes(1) == 1296.0


def test_c_abi_exports_every_declared_symbol():
    from pyoomph_b200.ccompiler import build_core_library
    lib = ctypes.CDLL(build_core_library())
    hdr = open(os.path.join(ROOT, "include", "pyoomph_b200.h")).read()
    names = re.findall(r"\b(pb2_[a-z0-9_]+)\s*\(", hdr)
    assert len(set(names)) >= 25
    for n in set(names):
        assert hasattr(lib, n), n
    assert lib.pb2_version() == int(re.search(r"#define PB2_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "pb2_jit_cuda.h")).read()).group(1))


def test_struct_layouts_match_header():
    """ctypes mirrors of the C structs have the sizes the compiler gives them."""
    import subprocess
    import tempfile
    from pyoomph_b200.assembly import ClassInfo, MeshDesc, TimeInfo
    src = '#include <stdio.h>\n#include "pyoomph_b200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(pb2_class_info), sizeof(pb2_mesh_desc), sizeof(pb2_time_info));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], check=True)
        out = subprocess.run([os.path.join(d, "s")], capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(ClassInfo), ctypes.sizeof(MeshDesc), ctypes.sizeof(TimeInfo)]


def test_product_path_never_touches_the_oracle():
    """the oracle is test infrastructure: nothing under pyoomph_b200/ may import or execute it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pyoomph_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt, f


def test_multi_assemble_request_groups_launches_and_keeps_request_order():
    """host logic of MultiAssembleRequest (bifurcation_tools.py:449-531) against a recording stand-in for the assembler"""
    import numpy as np
    from pyoomph_b200.multi_assembly import MultiAssembleRequest

    class Fake:
        n_dof, nnz = 4, 4
        indptr, indices = np.arange(5, dtype=np.int32), np.arange(4, dtype=np.int32)
        residual_names, param_names = ["", "azi"], ["mu", "Re"]

        def __init__(self):
            self.calls = []

        def assemble(self, flag=1, residual="", parameter=None):
            self.calls.append(("rjm", residual, parameter, flag))
            self._tag = 100 * self.residual_names.index(residual) + 10 * (0 if parameter is None else 1 + self.param_names.index(parameter))

        def launch_count(self):
            return 1

        def fetch(self, want_jacobian=True, want_mass=False):
            t = self._tag
            return np.full(4, t + 0.0), np.full(4, t + 1.0) if want_jacobian else None, np.full(4, t + 2.0) if want_mass else None

        def assemble_hessian(self, Y, flag=2, residual="", transposed=False):
            self.calls.append(("hvp", residual, Y.shape[0], flag))
            return [np.full(4, 1000 + Y[v, 0]) for v in range(Y.shape[0])], [np.full(4, 2000 + Y[v, 0]) if flag >= 2 else None for v in range(Y.shape[0])]
    f = Fake()
    vs = [np.full(4, float(i)) for i in range(6)]
    req = MultiAssembleRequest(f).M().R().dJdp("Re").J("azi").dRdp("Re").dMdU(vs[5])
    for v in vs[:5]:
        req.dJdU(v)
    req.dJdU(vs[5])                    # same array object again: no new Hessian vector
    out = req.assemble()
    # one launch per (contribution, parameter) at the highest flag, Hessian vectors handed over in blocks of 4
    assert f.calls[:3] == [("rjm", "", None, 2), ("rjm", "", "Re", 1), ("rjm", "azi", None, 1)]
    assert sorted(c[2] for c in f.calls[3:]) == [2, 4] and req.launches == 5
    assert out[0].data[0] == 2.0 and out[1][0] == 0.0 and out[2].data[0] == 21.0 and out[3].data[0] == 101.0 and out[4][0] == 20.0
    assert out[5].data[0] == 2005.0 and [m.data[0] for m in out[6:]] == [1000.0, 1001.0, 1002.0, 1003.0, 1004.0, 1005.0]
    with pytest.raises(RuntimeError):
        MultiAssembleRequest(f).R("nope")
    with pytest.raises(RuntimeError):
        MultiAssembleRequest(f).dRdp("nu")


def test_axisymmetric_operators_match_the_reference_formulas():
    """AxisymmetricCoordinateSystem against pyoomph/expressions/coordsys.py:386-537 written out by hand: measure 2 pi r dx, scalar
    gradient (d/dr, d/dz, 0), vector gradient with the u_r/r (and swirl) entries, divergence d_r u_r + u_r/r + d_z u_z, tensor
    divergence; Cartesian operators on the zero-padded vectors of an axisymmetric code (the ALE mesh equations); coordsys overrides."""
    import sympy as sp
    from pyoomph_b200 import expressions as ex
    from pyoomph_b200.codegen import Equations, FiniteElementCode
    seen = {}

    class Probe(Equations):
        def define_fields(self):
            self.define_vector_field("velocity", "C2", dim=3)
            self.define_scalar_field("pressure", "C1")

        def define_residuals(self):
            u, p = ex.var("velocity"), ex.var("pressure")
            x, y = ex.EUL[:2]
            r = ex.var("coordinate_x")
            assert u.shape == (3, 1) and ex.var("mesh").shape == (3, 1) and ex.var("mesh")[2] == 0
            assert ex.identity_matrix().shape == (3, 3)
            G = ex.grad(u)
            assert G[0, 2] == -u[2] / r and G[2, 2] == u[0] / r and G[1, 2] == 0 and G[2, 0] == sp.diff(u[2], x) and G[0, 1] == sp.diff(u[0], y)
            assert ex.grad(p) == sp.Matrix([sp.diff(p, x), sp.diff(p, y), 0])
            assert sp.simplify(ex.div(u) - (sp.diff(u[0], x) + u[0] / r + sp.diff(u[1], y))) == 0
            T = ex.dyadic(u, u)
            dT = ex.div(T)
            assert sp.simplify(dT[0] - (sp.diff(T[0, 0], x) + (T[0, 0] - T[2, 2]) / r + sp.diff(T[1, 0], y))) == 0
            assert sp.simplify(dT[2] - (sp.diff(T[0, 2], x) + (T[0, 2] - T[2, 0]) / r + sp.diff(T[1, 2], y))) == 0
            assert ex.weak(p, p) == p * p * 2 * ex.pi * r * ex.DX_EUL
            assert ex.weak(p, p, coordinate_system=ex.cartesian) == p * p * ex.DX_EUL
            R = ex.var("lagrangian_x")
            assert ex.Weak(p, p) == p * p * 2 * ex.pi * R * ex.DX_LAG
            Gc = ex.grad(ex.var("mesh"), lagrangian=True, coordsys=ex.cartesian)        # 3-vector, 2 coordinates: zero padded 3x3
            assert Gc.shape == (3, 3) and Gc[2, :] == sp.zeros(1, 3) and Gc[:, 2] == sp.zeros(3, 1)
            ms = ex.var("mesh")
            assert ex.div(ms, coordsys=ex.cartesian) == sp.diff(ms[0], x) + sp.diff(ms[1], y)      # (atomised to 1 + 1 later)
            seen["ok"] = True
            self.add_residual(ex.weak(ex.div(u), ex.testfunction("pressure")) + ex.weak(ex.grad(u), ex.grad(ex.testfunction("velocity"))))
    code = FiniteElementCode("Quad2dC2", Probe(), name="probe", coordinate_system="axisymmetric")
    assert seen["ok"] and code.coordinate_system.get_id_name() == "Axisymmetric"
    assert [f for f, _ in code.dof_layout()].count("velocity_phi") == 9 and len(code.dof_layout()) == 31
    # Cartesian codes are untouched: 2-vectors, 2x2 identity, plain dx
    plain = make_problem("ns", 2)["code"]
    assert plain.coordinate_system.get_id_name() == "Cartesian" and len(plain.dof_layout()) == 22


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host alone (the one place besides tests/ where the oracle may execute): ONE JSON line
    with the keys the driver reads, the e2e object with zero copy bytes, a cpu_baseline describing the run; other ranks stay silent."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "poisson"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    quiet = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""


def test_spatial_patches_give_compact_patches_on_relabelled_meshes():
    """the locality hint for meshes without lattice order: Morton-ordered patches touch few other patches, so the greedy patch colouring
    of the schedule needs a handful of colours where consecutive-element patches of a shuffled mesh need one
    colour per patch"""
    from problems import UnstructuredView
    from pyoomph_b200.meshes import RectangularQuadMesh, CuboidBrickMesh, spatial_patches

    def patch_colours(elem_nodes, patch):
        npatch = patch.max() + 1
        node_patches = {}
        for e, nodes in enumerate(elem_nodes):
            for nd in nodes:
                node_patches.setdefault(int(nd), set()).add(int(patch[e]))
        adj = [set() for _ in range(npatch)]
        for ps in node_patches.values():
            for a in ps:
                adj[a] |= ps
        col = -np.ones(npatch, dtype=int)
        for p_ in range(npatch):                     # greedy in patch order, like pb2_problem_create
            used = {col[q] for q in adj[p_] if q != p_ and col[q] >= 0}
            c = 0
            while c in used:
                c += 1
            col[p_] = c
        return col.max() + 1
    for mesh, size in ((UnstructuredView(RectangularQuadMesh(32), 3), 64), (UnstructuredView(CuboidBrickMesh(12), 4), 64)):
        hint = spatial_patches(mesh.node_pos, mesh.elem_nodes, size)
        assert hint.dtype == np.int32 and hint.shape == (mesh.n_elem,) and np.bincount(hint).max() <= size
        naive = (np.arange(mesh.n_elem) // size).astype(np.int32)
        assert patch_colours(mesh.elem_nodes, hint) <= 14 and 2 * patch_colours(mesh.elem_nodes, hint) <= patch_colours(mesh.elem_nodes, naive)


def test_bench_kernels_fit_the_sm_without_spills():
    """resource check of the cross-compiled sm_100a kernels the bench lines are measured on (ptxas -v log kept next to the plugin):
    the flag-1 kernels of configs 2 and 3 use no local memory (a spill in the contraction costs 15 % of all instructions,
    profiles/r01_notes.md), stay within the 65 536-register file at their block size and within 227 KB of shared memory."""
    from pyoomph_b200.ccompiler import get_ccompiler
    from pyoomph_b200.cuda_emitter import CudaEmitter
    cc = get_ccompiler("cuda")
    for kind, kernel in (("ns", "pb2_ns_r0_f1"), ("heat3d", "pb2_heat3d_r0_f1")):
        code = make_problem(kind, 2)["code"]
        em = CudaEmitter(code, code.name)
        so = cc.compile_code(em.emit(), code.name)
        log = open(so[:-3] + ".log").read()
        m = re.search(r"Compiling entry function '%s'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads.*?Used (\d+) registers" % kernel, log, re.S)
        assert m, kernel
        stack, st, ld, regs = (int(x) for x in m.groups())
        epb, threads, smem = em._kernel_cfg[kernel]
        assert (stack, st, ld) == (0, 0, 0), (kernel, stack, st, ld)
        assert regs * threads <= 65536 and smem <= 227 * 1024 and threads % 32 == 0 and 2 <= epb <= 63
        assert "sm_100a" in log


def test_sum_factorised_columns_algebra():
    """index algebra of the flagged 3D sum-factorised contraction (cuda_emitter, PB2_SUMFAC) with the
    emitter's own tables and index expressions: contracting the third, second and first direction in turn reproduces the sum over
    all 27 Gauss points of W0 psi_c + Ws . dpsi_c for every column c."""
    from pyoomph_b200.cuda_emitter import _lag, gauss_rule, shape_tables
    kn, _ = gauss_rule(3)
    psi, dpsi = shape_tables(3, 3, kn)
    t1 = []
    for sk in kn:                                  # c_t1d[ipt*18 + d*6 + n] / [.. + 3 + n], as in CudaEmitter._emit_tables
        for d in range(3):
            P, D = _lag(3, sk[d])
            t1 += list(P) + list(D)
    t1 = np.array(t1)
    rng = np.random.default_rng(5)
    W0, Ws = rng.uniform(-1, 1, 27), rng.uniform(-1, 1, (27, 3))
    ref = np.array([sum(W0[i] * psi[i][c] + sum(Ws[i, b] * dpsi[i][c][b] for b in range(3)) for i in range(27)) for c in range(27)])
    acc = np.zeros(27)
    for sp in range(3):
        U03, U1, U2 = np.zeros(9), np.zeros(9), np.zeros(9)
        for sq in range(3):
            for sr in range(3):
                ipt = sp * 9 + sq * 3 + sr
                for c in range(3):
                    tL2, tD2 = t1[ipt * 18 + 12 + c], t1[ipt * 18 + 15 + c]
                    U03[sq * 3 + c] += W0[ipt] * tL2 + Ws[ipt, 2] * tD2
                    U1[sq * 3 + c] += Ws[ipt, 0] * tL2
                    U2[sq * 3 + c] += Ws[ipt, 1] * tL2
        for c in range(3):
            for b in range(3):
                M = sum(U03[q * 3 + c] * t1[q * 54 + 6 + b] + U2[q * 3 + c] * t1[q * 54 + 9 + b] for q in range(3))
                V = sum(U1[q * 3 + c] * t1[q * 54 + 6 + b] for q in range(3))
                for a in range(3):
                    acc[a + 3 * b + 9 * c] += M * t1[sp * 162 + a] + V * t1[sp * 162 + 3 + a]
    assert np.abs(acc - ref).max() <= 1e-14 * np.abs(ref).max()
