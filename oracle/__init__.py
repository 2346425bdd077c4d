"""ORACLE -- TEST INFRASTRUCTURE ONLY.

CPU restatement of pyoomph's assembly path (generated-format C plugin + restated host driver), used as the
checker in tests/, in __graft_entry__.smoke() and as bench.py's CPU baseline.  Nothing under pyoomph_b200/
imports this package; the product path never executes it.

Parity status: the reference cannot be built or imported in this environment (GiNaC/CLN are not
vendored) and ships no golden vectors for this path.  What IS pinned against the reference's
own code: Gauss tables and Lagrange shape functions (against oomph-lib sources compiled into oracle/_ref), the
plugin ABI headers and accumulate macros (same generated C compiled against /root/reference/src/jitbridge*.h,
bit-identical results).  The geometry/derivation logic is pinned only by patch, finite-difference and invariance
tests => "parity unpinned" at the Jacobian level.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from typing import Optional

import numpy as np

from .emit_c import emit_plugin_source

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
REF_BUILD = os.path.join(HERE, "_ref")
REFERENCE_SRC = "/root/reference/src"

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


def _dp(a):
    return a.ctypes.data_as(c_double_p) if a is not None else None


def _ip(a):
    return a.ctypes.data_as(c_int_p) if a is not None else None


def cpu_flags(fast_math: bool = False):
    """SystemCCompiler flags (/root/reference/pyoomph/generic/ccompiler.py:214-220, :233)."""
    fl = ["-O3", "-fPIC", "-march=native"]
    if fast_math:
        fl.append("-ffast-math")
    return fl


_CPU_TAG = None


def _cpu_tag() -> str:
    """fingerprint of the host CPU's instruction-set flags: the plugins are compiled with -march=native (the reference's SystemCCompiler
    flags), so a library built on one host must not be reused on a host with another instruction set (it is rebuilt there instead)"""
    global _CPU_TAG
    if _CPU_TAG is None:
        try:
            flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags"))
            _CPU_TAG = hashlib.sha1(" ".join(sorted(flags.split(":", 1)[1].split())).encode()).hexdigest()[:8]
        except Exception:
            _CPU_TAG = "unknown"
    return _CPU_TAG


def build_plugin(code, name: str, *, reference_headers: bool = False, fast_math: bool = False, force: bool = False) -> str:
    """Emit the generated-format C for `code`, compile it with the restated driver, return the .so path.

    reference_headers=True compiles the same sources against the reference's own jitbridge.h/jitbridge_hang.h
    where they lie (outputs only into oracle/_ref/)."""
    src = emit_plugin_source(code)
    outdir = REF_BUILD if reference_headers else BUILD
    os.makedirs(outdir, exist_ok=True)
    driver = open(os.path.join(HERE, "driver.c")).read()
    hdrs = open(os.path.join(HERE, "oracle_jit.h")).read() + open(os.path.join(HERE, "oracle_jit_hang.h")).read()
    tag = hashlib.sha1((src + driver + hdrs + str(fast_math) + _cpu_tag()).encode()).hexdigest()[:12]
    cfile = os.path.join(outdir, "%s_%s.c" % (name, tag))
    sofile = os.path.join(outdir, "%s_%s%s.so" % (name, tag, "_ref" if reference_headers else ""))
    if os.path.exists(sofile) and not force:
        return sofile
    # several processes (the ranks of a multi-process test) may build the same plugin at once: everyone compiles into files of its own
    # and renames the finished library into place, so that nobody ever loads a half-written one
    tmp = "%s.%d.tmp" % (sofile, os.getpid())
    ctmp = "%s.%d.c" % (cfile[:-2], os.getpid())
    with open(ctmp, "w") as f:
        f.write(src)
    cmd = ["gcc", "-std=gnu99"] + cpu_flags(fast_math) + ["-fopenmp", "-shared", "-I", HERE]
    if reference_headers:
        if not os.path.isdir(REFERENCE_SRC):
            raise RuntimeError("reference tree not present")
        cmd += ["-DORACLE_USE_REFERENCE_HEADERS", "-I", REFERENCE_SRC]
    cmd += [os.path.join(HERE, "driver.c"), ctmp, "-o", tmp, "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        for f_ in (tmp, ctmp):
            if os.path.exists(f_):
                os.remove(f_)
        raise RuntimeError("oracle plugin compilation failed:\n" + r.stderr[-4000:])
    os.replace(ctmp, cfile)
    os.replace(tmp, sofile)
    return sofile


class OracleProblem:
    """One mesh + one element class, assembled the reference's way on the CPU."""

    def __init__(self, code, mesh, dofmap, node_val: np.ndarray, *, node_pos_hist: Optional[np.ndarray] = None,
                 node_lagr: Optional[np.ndarray] = None, name: str = "plugin", reference_headers: bool = False,
                 fast_math: bool = False, so_path: Optional[str] = None):
        self.code, self.mesh, self.dofmap = code, mesh, dofmap
        so = so_path or build_plugin(code, name, reference_headers=reference_headers, fast_math=fast_math)
        self.lib = ctypes.CDLL(so)
        L = self.lib
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_create_typed.restype = ctypes.c_void_p
        L.oracle_assemble.restype = ctypes.c_double
        L.oracle_nnz.restype = ctypes.c_int64
        L.oracle_eval_integral.restype = ctypes.c_double
        L.oracle_eval_at_s.restype = ctypes.c_double
        L.oracle_integral_name.restype = ctypes.c_char_p
        L.oracle_element.restype = ctypes.c_int
        self.dim = mesh.dim
        self.elem_nodes = np.ascontiguousarray(mesh.elem_nodes, dtype=np.int32)
        node_val = np.ascontiguousarray(node_val, dtype=np.float64)
        if node_val.ndim == 2:
            node_val = node_val[None]
        self.T = node_val.shape[0]
        pos = mesh.node_pos[None] if node_pos_hist is None else node_pos_hist
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        lagr = np.ascontiguousarray(mesh.node_pos if node_lagr is None else node_lagr, dtype=np.float64)
        self.node_eqn = np.ascontiguousarray(dofmap.node_eqn, dtype=np.int32)
        self.pos_eqn = None if dofmap.pos_eqn is None else np.ascontiguousarray(dofmap.pos_eqn, dtype=np.int32)
        self.n_dof = dofmap.n_dof
        self.h = ctypes.c_void_p(L.oracle_create_typed(self.dim, int(mesh.elem_nodes.shape[1]), mesh.n_elem, _ip(self.elem_nodes), mesh.n_node,
                                                 node_val.shape[2], self.T, pos.shape[0], _dp(pos), _dp(lagr),
                                                 _dp(node_val), _ip(self.node_eqn), _ip(self.pos_eqn), self.n_dof))
        self.maxdof = 2 * mesh.elem_nodes.shape[1] * (self.dim + node_val.shape[2])     # x2: master values outside the element
        if code.etype.name.startswith("QuadFace"):
            L.oracle_set_face_mode(self.h)
        hanging = getattr(mesh, "hanging", None)
        if hanging is not None:
            for space, table in ((0, hanging.C2), (1, hanging.C1)):
                start = np.zeros(mesh.n_node + 1, dtype=np.int32)
                for n, (m, w) in table.items():
                    start[n + 1] = len(m)
                start = np.cumsum(start).astype(np.int32)
                masters = np.zeros(max(1, int(start[-1])), dtype=np.int32)
                weights = np.zeros(max(1, int(start[-1])))
                for n, (m, w) in table.items():
                    masters[start[n]:start[n + 1]] = m
                    weights[start[n]:start[n + 1]] = w
                L.oracle_set_hanging(self.h, space, _ip(start), _ip(masters), _dp(weights))

    def evaluate_integral_expressions(self):
        """{name: sum over elements of EvalIntegralExpression(index)} (Mesh::evaluate_integral_expression, src/mesh.cpp:536)"""
        n = self.lib.oracle_num_integrals(self.h)
        return {self.lib.oracle_integral_name(self.h, i).decode(): float(self.lib.oracle_eval_integral(self.h, i)) for i in range(n)}

    def eval_point_expressions(self, points: str = "nodes") -> np.ndarray:
        """every local expression, extremum expression and Z2 flux term (in that order, like the product's point_names) at the
        integration points or the nodes of every element, one reference-style call per point: [n_elem, n_points, n_expressions]"""
        dim = self.dim
        nn = self.mesh.elem_nodes.shape[1]
        tri = dim == 2 and nn == 6
        if points == "nodes":
            grid = (-1.0, 0.0, 1.0)
            if tri:
                pts = [np.array(c) for c in ((1.0, 0.0), (0.0, 1.0), (0.0, 0.0), (0.5, 0.5), (0.0, 0.5), (0.5, 0.0))]   # Telements.h:575-621
            else:
                pts = [np.array([grid[(l // 3 ** d) % 3] for d in range(dim)]) for l in range(nn)]    # local_coordinate_of_node
        else:
            pts = []
            for ipt in range(7 if tri else (9 if dim == 2 else 27)):
                k, w = (ctypes.c_double * 3)(), ctypes.c_double()
                if tri:
                    self.lib.oracle_gauss_tri(ipt, k, ctypes.byref(w))
                else:
                    self.lib.oracle_gauss(dim, ipt, k, ctypes.byref(w))
                pts.append(np.array(list(k)[:dim]))
        nl, nx, nz = (self.lib.oracle_num_point_exprs(self.h, k) for k in (0, 1, 2))
        out = np.zeros((self.mesh.elem_nodes.shape[0], len(pts), nl + nx + nz))
        zbuf = np.zeros(max(1, nz))
        for e in range(out.shape[0]):
            for ip, s in enumerate(pts):
                s = np.ascontiguousarray(s, dtype=np.float64)
                for i in range(nl):
                    out[e, ip, i] = self.lib.oracle_eval_at_s(self.h, 0, i, e, _dp(s), None)
                for i in range(nx):
                    out[e, ip, nl + i] = self.lib.oracle_eval_at_s(self.h, 1, i, e, _dp(s), None)
                if nz:
                    self.lib.oracle_eval_at_s(self.h, 2, 0, e, _dp(s), _dp(zbuf))
                    out[e, ip, nl + nx:] = zbuf[:nz]
        return out

    def set_params(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self.lib.oracle_set_params(self.h, _dp(v), len(v))

    def set_steady(self):
        z = np.zeros(7)
        self.lib.oracle_set_time(self.h, 1, 0, 0, _dp(z), _dp(z), _dp(z), _dp(z), _dp(z), _dp(z))

    def set_unsteady(self, t: float, dt: float, dtprev: float, unsteady_steps_done: int, ntstorage: Optional[int] = None):
        """MultiTimeStepper weights (src/timestepper.cpp:31-80); ntstorage defaults to the history levels handed to the constructor
        (3 for BDF2 problems, 5 = NSTEPS + 3 when the Newmark velocity / acceleration slots are stored)"""
        w1, w2, n1, n2 = np.zeros(7), np.zeros(7), np.zeros(7), np.zeros(7)
        self.lib.oracle_bdf_weights(ctypes.c_double(dt), ctypes.c_double(dtprev), _dp(w1), _dp(w2))
        self.lib.oracle_newmark2_weights(ctypes.c_double(dt), ctypes.c_double(0.5), ctypes.c_double(0.5), _dp(n1), _dp(n2))
        tt = np.zeros(7); tt[0] = t; tt[1] = t - dt; tt[2] = t - dt - dtprev
        dd = np.zeros(7); dd[0] = dt; dd[1] = dtprev
        if ntstorage is None:
            ntstorage = max(3, self.T)
        self.lib.oracle_set_time(self.h, 0, unsteady_steps_done, ntstorage, _dp(tt), _dp(dd), _dp(w1), _dp(w2), _dp(n1), _dp(n2))
        return w1, w2

    def update_values(self, t: int, node_val=None, node_pos=None):
        nv = None if node_val is None else np.ascontiguousarray(node_val, dtype=np.float64)
        npos = None if node_pos is None else np.ascontiguousarray(node_pos, dtype=np.float64)
        self.lib.oracle_update_values(self.h, t, _dp(nv), _dp(npos))

    def element(self, e: int, which: int = 0, param: int = -1, flag: int = 1):
        n = self.maxdof
        R, J, M = np.zeros(n), np.zeros(n * n), np.zeros(n * n)
        eq = np.zeros(n, dtype=np.int32)
        nd = self.lib.oracle_element(self.h, e, which, param, flag, _dp(R), _dp(J), _dp(M), _ip(eq))
        # the routine used jacobian_size = ndof
        return R[:nd].copy(), J[:nd * nd].reshape(nd, nd).copy(), M[:nd * nd].reshape(nd, nd).copy(), eq[:nd].copy()

    def assemble(self, which: int = 0, param: int = -1, flag: int = 1, nthreads: int = 1, fetch: bool = True):
        """Returns residual and (row_start, col_index, value) per matrix in the reference's vectors_of_pairs
        order (columns in first-touch order, exact zeros dropped).  fetch=False leaves the matrices in the C arrays the
        assembly produced (what the reference's timing of an assembly covers) and returns only the residual."""
        res = np.zeros(self.n_dof)
        self.lib.oracle_assemble(self.h, which, param, flag, _dp(res), nthreads)
        if not fetch:
            return res, []
        mats = []
        for m in range(0 if flag == 0 else (1 if flag == 1 else 2)):
            nnz = int(self.lib.oracle_nnz(self.h, m))
            rs = np.zeros(self.n_dof + 1, dtype=np.int32)
            ci = np.zeros(nnz, dtype=np.int32)
            va = np.zeros(nnz)
            self.lib.oracle_get_csr(self.h, m, _ip(rs), _ip(ci), _dp(va))
            mats.append((rs, ci, va))
        return res, mats

    def element_hessian(self, e: int, Y: np.ndarray, C: Optional[np.ndarray] = None, flag: int = 1, which: int = 0):
        """HessianVectorProduct<which> on one element; Y (and C for flag 0) are GLOBAL vectors [nvec][n_dof]."""
        Y = np.ascontiguousarray(np.atleast_2d(Y), dtype=np.float64)
        n = self.maxdof
        if flag == 0:
            C = np.ascontiguousarray(np.atleast_2d(C), dtype=np.float64)
            nvec = C.shape[0]
            prod, Cs = np.zeros(nvec * n), None
        elif flag == 3:
            nvec = 1
            prod, Cs = np.zeros(n * n * n), np.zeros(n * n * n)      # the raw H_ijk and mass Hessian, [i*n^2 + j*n + k]
        else:
            nvec = Y.shape[0]
            prod, Cs = np.zeros(nvec * n * n), np.zeros(nvec * n * n)
        eq = np.zeros(n, dtype=np.int32)
        nd = self.lib.oracle_element_hessian(self.h, e, which, _dp(Y), _dp(C) if flag == 0 else None, nvec, flag, _dp(prod), _dp(Cs), _ip(eq))
        if flag == 0:
            return prod[:nvec * nd].reshape(nvec, nd).copy(), eq[:nd].copy()
        if flag == 3:
            return prod[:nd ** 3].reshape(nd, nd, nd).copy(), Cs[:nd ** 3].reshape(nd, nd, nd).copy(), eq[:nd].copy()
        return prod[:nvec * nd * nd].reshape(nvec, nd, nd).copy(), Cs[:nvec * nd * nd].reshape(nvec, nd, nd).copy(), eq[:nd].copy()

    def assemble_hessian_tensor(self, which: int = 0):
        """Problem::assemble_hessian_tensor restated (src/problem.cpp:1530-1560): per element the flag-3 buffer of
        HessianVectorProduct, T(iG, jG, kG) += H_e[i][k][j] for |value| > 0; returns the list of (i, j, k, value) contributions in the
        reference's accumulation order"""
        ii, jj, kk, vv = [], [], [], []
        Y0 = np.zeros((1, self.n_dof))
        for e in range(self.mesh.elem_nodes.shape[0]):
            H, _, eq = self.element_hessian(e, Y0, flag=3, which=which)
            nd = eq.size
            for i in range(nd):
                for j in range(nd):
                    for k in range(nd):
                        hval = H[i, k, j]
                        if abs(hval) > 0.0:
                            ii.append(eq[i]); jj.append(eq[j]); kk.append(eq[k]); vv.append(hval)
        return np.array(ii, dtype=np.int32), np.array(jj, dtype=np.int32), np.array(kk, dtype=np.int32), np.array(vv)

    def assemble_hessian(self, Y: np.ndarray, flag: int = 2, which: int = 0):
        """global d(J.Y_v)/dU (and d(M.Y_v)/dU; flags 4 / 5: the transposed contractions) as scipy CSR matrices, element by element
        like get_multi_assembly"""
        from scipy.sparse import coo_matrix
        Y = np.ascontiguousarray(np.atleast_2d(Y), dtype=np.float64)
        nvec, n = Y.shape[0], self.n_dof
        rows, cols, vj, vm = [], [], [[] for _ in range(nvec)], [[] for _ in range(nvec)]
        for e in range(self.mesh.elem_nodes.shape[0]):
            P, Cs, eq = self.element_hessian(e, Y, flag=flag, which=which)
            r, c = np.meshgrid(eq, eq, indexing="ij")
            rows.append(r.ravel()); cols.append(c.ravel())
            for v in range(nvec):
                vj[v].append(P[v].ravel()); vm[v].append(Cs[v].ravel())
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        J = [coo_matrix((np.concatenate(vj[v]), (rows, cols)), shape=(n, n)).tocsr() for v in range(nvec)]
        M = [coo_matrix((np.concatenate(vm[v]), (rows, cols)), shape=(n, n)).tocsr() for v in range(nvec)]
        return J, M

    def point_shapes(self, e: int, ipt: int, flag: int = 1):
        nn, d = self.mesh.elem_nodes.shape[1], self.dim
        w = np.zeros(3); sh = np.zeros(nn); dx = np.zeros((nn, d)); dX = np.zeros((nn, d))
        wd = np.zeros((d, nn)); dd = np.zeros((nn, d, nn, d))
        self.lib.oracle_point_shapes(self.h, e, ipt, flag, _dp(w), _dp(sh), _dp(dx), _dp(dX), _dp(wd), _dp(dd))
        return w, sh, dx, dX, wd, dd

    def close(self):
        if self.h:
            self.lib.oracle_free(self.h)
            self.h = None
