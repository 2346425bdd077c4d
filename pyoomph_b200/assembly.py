"""Host-side assembler: the Python mirror of the reference interfaces for the assembly hot path.

``B200Assembly`` implements the contract of pyoomph's ``CustomAssemblyBase``
(/root/reference/pyoomph/generic/assembly.py:36-86): ``get_residuals_and_jacobian(require_jacobian, dparameter)``
returns a float64 residual and a CSR matrix with int32 ``indptr``/``indices`` and float64 ``data``, which is exactly
what ``Problem.get_custom_residuals_jacobian`` (pyoomph/generic/problem.py:1727-1745) hands to
``CRDoubleMatrix::build`` (src/problem.cpp:965); ``invalidate_cache`` etc. keep their meaning.  Everything numeric
happens in libpyoomph_b200.so through the C-ABI of include/pyoomph_b200.h; this module only marshals arrays.
The library and the CUDA plugin are mandatory: a missing build raises, nothing is computed on the CPU.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Tuple

import numpy as np

from .ccompiler import CudaCCompiler, get_ccompiler
from .codegen import FiniteElementCode
from .cuda_emitter import CudaEmitter

HERE = os.path.dirname(os.path.abspath(__file__))
PB2_NTW = 7
PB2_MAX_PARAMS = 16
PB2_MAX_FIELDS = 16

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)


class TimeInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double * PB2_NTW) for n in
                ("t", "dt", "w_dt_BDF1", "w_dt_BDF2", "w_dt_Newmark2", "w_d2t_Newmark2", "w_dt_BDF2_degr", "w_dt_Newmark2_degr")] + \
               [("ntstorage", ctypes.c_int), ("pad_", ctypes.c_int)]


class ClassInfo(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int), ("name", ctypes.c_char * 64),
        ("nodal_dim", ctypes.c_int), ("elem_dim", ctypes.c_int), ("nnode", ctypes.c_int), ("nnode_C1", ctypes.c_int),
        ("c1_nodes", ctypes.c_int * 8), ("n_int_pt", ctypes.c_int), ("nval", ctypes.c_int), ("n_fields", ctypes.c_int),
        ("field_names", (ctypes.c_char * 48) * PB2_MAX_FIELDS), ("field_space", ctypes.c_int * PB2_MAX_FIELDS),
        ("field_index", ctypes.c_int * PB2_MAX_FIELDS), ("moving_nodes", ctypes.c_int), ("ndof_el", ctypes.c_int),
        ("dof_node", ctypes.c_int * 160), ("dof_kind", ctypes.c_int * 160), ("dof_index", ctypes.c_int * 160),
        ("n_residuals", ctypes.c_int), ("residual_names", (ctypes.c_char * 48) * 8),
        ("n_params", ctypes.c_int), ("param_names", (ctypes.c_char * 48) * PB2_MAX_PARAMS),
        ("n_hist_val", ctypes.c_int), ("n_hist_pos", ctypes.c_int), ("max_dt_order", ctypes.c_int),
        ("elems_per_block", ctypes.c_int), ("threads_per_block", ctypes.c_int), ("smem_bytes", ctypes.c_int),
        ("hessian_generated", ctypes.c_int),
        ("alg_bytes_per_elem", ctypes.c_double * 3), ("flops_per_elem", ctypes.c_double * 3),
        ("alg_bytes_per_hist_level", ctypes.c_double),
        ("n_integrals", ctypes.c_int), ("pad_", ctypes.c_int), ("integral_names", (ctypes.c_char * 48) * 16),
        ("n_point_exprs", ctypes.c_int), ("pad2_", ctypes.c_int), ("point_names", (ctypes.c_char * 48) * 32), ("point_kind", ctypes.c_int * 32),
    ]


class MeshDesc(ctypes.Structure):
    _fields_ = [("n_elem", ctypes.c_longlong), ("n_node", ctypes.c_longlong), ("elem_nodes", c_int_p),
                ("node_eqn", c_int_p), ("pos_eqn", c_int_p), ("n_dof", ctypes.c_longlong), ("elem_patch", c_int_p),
                ("n_extra", ctypes.c_longlong), ("extra_rows", c_int_p), ("extra_cols", c_int_p)]


_LIB = None


def load_library() -> ctypes.CDLL:
    """Load libpyoomph_b200.so (built in-tree by __graft_entry__.build / ccompiler.build_core_library)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(HERE, "libpyoomph_b200.so")
    if not os.path.exists(path):
        raise RuntimeError("libpyoomph_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    L.pb2_last_error.restype = ctypes.c_char_p
    L.pb2_problem_launch_count.restype = ctypes.c_longlong
    for fn in ("pb2_class_load", "pb2_class_get_info", "pb2_problem_create", "pb2_problem_pattern", "pb2_problem_set_nodal_values",
               "pb2_problem_set_nodal_positions", "pb2_problem_set_lagrangian_positions", "pb2_problem_set_dofs",
               "pb2_problem_set_time", "pb2_problem_set_parameters", "pb2_problem_assemble", "pb2_problem_device_outputs",
               "pb2_problem_fetch", "pb2_problem_assemble_host", "pb2_problem_num_colours", "pb2_version",
               "pb2_problem_assemble_hessian", "pb2_problem_fetch_hessian", "pb2_problem_hessian_vector_products",
               "pb2_problem_pack_rows", "pb2_problem_unpack_add", "pb2_problem_eval_integrals", "pb2_problem_shift_time_values",
               "pb2_problem_set_history_dofs", "pb2_measure_fp64_peak", "pb2_problem_host_maps", "pb2_problem_device_pattern",
               "pb2_problem_device_dofs", "pb2_problem_update_dofs_device", "pb2_problem_eval_points", "pb2_problem_create_child", "pb2_problem_set_constraints"):
        getattr(L, fn).restype = ctypes.c_int
    L.pb2_problem_setup_seconds.restype = ctypes.c_double
    _LIB = L
    return L


def _check(rc: int):
    if rc != 0:
        raise RuntimeError("pyoomph_b200: " + load_library().pb2_last_error().decode())


def bdf_weights(dt: float, dtprev: float):
    """MultiTimeStepper::set_weights, BDF part (/root/reference/src/timestepper.cpp:31-59)."""
    w1, w2 = np.zeros(PB2_NTW), np.zeros(PB2_NTW)
    w2[0] = 1.0 / dt + 1.0 / (dt + dtprev)
    w2[1] = -(dt + dtprev) / (dt * dtprev)
    w2[2] = dt / ((dt + dtprev) * dtprev)
    w1[0] = 1.0 / dt
    w1[1] = -1.0 / dt
    return w1, w2


NSTEPS = 2          # MultiTimeStepper::NSTEPS (src/timestepper.hpp:49): history values 1..2, then the Newmark velocity / acceleration slots


def newmark2_weights(dt: float, beta1: float = 0.5, beta2: float = 0.5):
    """MultiTimeStepper::set_weights, Newmark2 part (/root/reference/src/timestepper.cpp:61-80; NewmarkBeta1 = NewmarkBeta2 = 0.5,
    src/timestepper.hpp:63).  Returns (first-derivative weights, second-derivative weights); the slots NSTEPS+1 and NSTEPS+2
    multiply the stored velocity and acceleration of the previous step."""
    w1, w2 = np.zeros(PB2_NTW), np.zeros(PB2_NTW)
    w2[0] = 2.0 / (beta2 * dt * dt)
    w2[1] = -2.0 / (beta2 * dt * dt)
    w2[NSTEPS + 1] = -2.0 / (dt * beta2)
    w2[NSTEPS + 2] = (beta2 - 1.0) / beta2
    w1[0] = beta1 * dt * w2[0]
    w1[1] = beta1 * dt * w2[1]
    w1[NSTEPS + 1] = 1.0 + beta1 * dt * w2[NSTEPS + 1]
    w1[NSTEPS + 2] = dt * (1.0 - beta1) + beta1 * dt * w2[NSTEPS + 2]
    return w1, w2


class CustomAssemblyBase:
    """Interface of pyoomph/generic/assembly.py:36 (hooks kept so that Problem.set_custom_assembler accepts it)."""

    def __init__(self) -> None:
        self.problem = None

    def _set_problem(self, problem):
        self.problem = problem

    def has_custom_solve_routine(self) -> bool:
        return False

    def invalidate_cache(self) -> None:
        pass

    def actions_after_adapt(self) -> None:
        self.invalidate_cache()

    def actions_after_remeshing(self) -> None:
        self.invalidate_cache()

    def actions_after_equation_numbering(self) -> None:
        self.invalidate_cache()

    def actions_after_setting_initial_condition(self) -> None:
        self.invalidate_cache()

    def actions_after_successful_newton_solve(self) -> None:
        pass

    def initialize(self) -> None:
        pass

    def finalize(self) -> None:
        pass

    def get_residuals_and_jacobian(self, require_jacobian: bool, dparameter: Optional[str] = None):
        raise RuntimeError("Must be implemented")


class B200Assembly(CustomAssemblyBase):
    """One element class on one mesh, assembled on one B200."""

    def __init__(self, code: FiniteElementCode, mesh, dofmap, *, name: str = "elem", device: int = 0,
                 compiler: Optional[CudaCCompiler] = None, emitter_options: Optional[dict] = None,
                 elements: Optional[np.ndarray] = None, extra_pattern: Optional[Tuple[np.ndarray, np.ndarray]] = None,
                 patch_hint=None, parent: Optional["B200Assembly"] = None):
        """parent: another assembly whose CSR matrix, residual and nodal data this element class scatters into (interface / boundary
        element classes on the parent's nodes and equations: pb2_problem_create_child); `mesh` then carries the child's elements
        (meshes.boundary_line_mesh) and `dofmap` is the parent's.  `parent.assemble()` assembles the parent and then its children.
        patch_hint: None = the mesh's own `element_patches()` if it has one, else consecutive elements; "spatial" = Morton-ordered
        patches from the element centroids (meshes.spatial_patches, for meshes without lattice order); or an int array [n_elem]."""
        super().__init__()
        self._patch_hint = patch_hint
        self.code, self.mesh, self.dofmap = code, mesh, dofmap
        self.lib = load_library()
        self.compiler = compiler or get_ccompiler("cuda")
        self.emitter = CudaEmitter(code, name, **(emitter_options or {}))
        self.plugin_path = self.compiler.compile_code(self.emitter.emit(), name)
        self.cls = ctypes.c_void_p()
        _check(self.lib.pb2_class_load(self.plugin_path.encode(), ctypes.byref(self.cls)))
        self.info = ClassInfo()
        _check(self.lib.pb2_class_get_info(self.cls, ctypes.byref(self.info)))
        self._device, self._elements, self._extra_pattern = device, elements, extra_pattern
        self.prob = None
        self.parent, self.children = parent, []
        if parent is not None:
            if elements is not None or extra_pattern is not None:
                raise ValueError("a child assembly has no element subset / pattern of its own")
            self._device = parent._device
        self._create_problem(mesh, dofmap)
        if parent is not None:
            parent.children.append(self)

    def _create_problem(self, mesh, dofmap):
        """pack mesh + numbering into the SoA buffers of a new pb2_problem (the element class and its kernels are kept)"""
        device, elements, extra_pattern = self._device, self._elements, self._extra_pattern
        self.mesh, self.dofmap = mesh, dofmap
        en = mesh.elem_nodes if elements is None else mesh.elem_nodes[elements]
        self._elem_nodes = np.ascontiguousarray(en, dtype=np.int32)
        self._node_eqn = np.ascontiguousarray(dofmap.node_eqn, dtype=np.int32)
        self._pos_eqn = None if dofmap.pos_eqn is None else np.ascontiguousarray(dofmap.pos_eqn, dtype=np.int32)
        if isinstance(self._patch_hint, str):
            if self._patch_hint != "spatial":
                raise ValueError("unknown patch hint '%s'" % self._patch_hint)
            from .meshes import spatial_patches
            patches = spatial_patches(mesh.node_pos, mesh.elem_nodes)
        elif self._patch_hint is not None:
            patches = np.asarray(self._patch_hint)
            if patches.shape != (mesh.elem_nodes.shape[0],):
                raise ValueError("patch hint must have one entry per element of the mesh")
        else:
            patches = mesh.element_patches() if hasattr(mesh, "element_patches") else None
        if patches is not None:
            patches = patches if elements is None else patches[elements]
            _, patches = np.unique(patches, return_inverse=True)       # dense ids, order preserved
            self._elem_patch = np.ascontiguousarray(patches, dtype=np.int32)
        else:
            self._elem_patch = None
        md = MeshDesc(self._elem_nodes.shape[0], mesh.n_node, self._elem_nodes.ctypes.data_as(c_int_p),
                      self._node_eqn.ctypes.data_as(c_int_p),
                      None if self._pos_eqn is None else self._pos_eqn.ctypes.data_as(c_int_p), dofmap.n_dof,
                      None if self._elem_patch is None else self._elem_patch.ctypes.data_as(c_int_p), 0, None, None)
        if extra_pattern is not None and len(extra_pattern[0]):
            self._ex_rows = np.ascontiguousarray(extra_pattern[0], dtype=np.int32)
            self._ex_cols = np.ascontiguousarray(extra_pattern[1], dtype=np.int32)
            md.n_extra = self._ex_rows.size
            md.extra_rows = self._ex_rows.ctypes.data_as(c_int_p)
            md.extra_cols = self._ex_cols.ctypes.data_as(c_int_p)
        self.prob = ctypes.c_void_p()
        if self.parent is not None:
            self.parent._fresh()
            _check(self.lib.pb2_problem_create_child(self.cls, self.parent.prob, ctypes.byref(md), ctypes.byref(self.prob)))
        else:
            _check(self.lib.pb2_problem_create(self.cls, device, ctypes.byref(md), ctypes.byref(self.prob)))
        rs, ci = c_int_p(), c_int_p()
        nnz, nrows = ctypes.c_longlong(), ctypes.c_longlong()
        _check(self.lib.pb2_problem_pattern(self.prob, ctypes.byref(rs), ctypes.byref(ci), ctypes.byref(nnz), ctypes.byref(nrows)))
        self.n_dof, self.nnz = int(nrows.value), int(nnz.value)
        # numpy-owned copies: the arrays end up inside scipy matrices and exchange lists that outlive rebuild() / close()
        self.indptr = np.ctypeslib.as_array(rs, shape=(self.n_dof + 1,)).copy()
        # an empty pattern (every dof pinned, or no element on this rank) has no column array at all
        self.indices = np.ctypeslib.as_array(ci, shape=(self.nnz,)).copy() if self.nnz > 0 else np.zeros(0, dtype=np.int32)
        self.setup_seconds = float(self.lib.pb2_problem_setup_seconds(self.prob))
        self.n_elem = self._elem_nodes.shape[0]
        self.param_names = [self.info.param_names[i].value.decode() for i in range(self.info.n_params)]
        self.residual_names = [self.info.residual_names[i].value.decode() for i in range(self.info.n_residuals)]
        self.integral_names = [self.info.integral_names[i].value.decode() for i in range(self.info.n_integrals)]
        self.point_names = [(("local", "extremum", "z2")[self.info.point_kind[i]], self.info.point_names[i].value.decode())
                            for i in range(self.info.n_point_exprs)]
        if not hasattr(self, "_params"):
            self._params = np.zeros(max(1, self.info.n_params))
        self._stale = False
        if device < 0:
            return                      # pattern-only problem: no device data to initialise
        if self.parent is not None:
            # nodal data are the parent's; only the class's own time info and parameters are set
            self.ti = self.parent.ti
            _check(self.lib.pb2_problem_set_time(self.prob, ctypes.byref(self.ti)))
            _check(self.lib.pb2_problem_set_parameters(self.prob, self._dp(self._params), self.info.n_params))
            self._last_flag = -1
            return
        self.set_nodal_positions(0, mesh.node_pos)
        for t in range(1, self.info.n_hist_pos):
            self.set_nodal_positions(t, mesh.node_pos)
        self.set_lagrangian_positions(mesh.node_pos)
        if hasattr(self, "ti"):
            _check(self.lib.pb2_problem_set_time(self.prob, ctypes.byref(self.ti)))     # rebuild: keep time info and parameters
            _check(self.lib.pb2_problem_set_parameters(self.prob, self._dp(self._params), self.info.n_params))
        else:
            self.set_steady()
        self._last_flag = -1

    # ---- invalidation (CustomAssemblyBase hooks, pyoomph/generic/assembly.py:52-65) --------------------------------------------
    def invalidate_cache(self) -> None:
        """equation numbering, adaptation or remeshing changed what was packed: the next assembly must not use it"""
        self._stale = True

    def rebuild(self, mesh=None, dofmap=None, elements=None, extra_pattern=None) -> None:
        """re-pack after renumbering / adaptation / remeshing: new pb2_problem (pattern, maps, schedule), same compiled class;
        nodal values and positions of the new mesh must be set again (set_nodal_values / set_dofs)"""
        if elements is not None or extra_pattern is not None:
            self._elements, self._extra_pattern = elements, extra_pattern
        for c in self.children:         # children alias this problem's buffers: released first, to be rebuilt by the caller afterwards
            c._release()
            c._stale = True
        if self.prob:
            self.lib.pb2_problem_free(self.prob)
            self.prob = None
        self._create_problem(mesh if mesh is not None else self.mesh, dofmap if dofmap is not None else self.dofmap)

    def _release(self):
        if self.prob:
            self.lib.pb2_problem_free(self.prob)
            self.prob = None

    def _fresh(self):
        if self._stale:
            raise RuntimeError("the packed problem was invalidated (renumbering / adaptation / remeshing): call rebuild(mesh, dofmap) first")

    def shift_time_values(self):
        """Problem::shift_time_values on the device-resident history levels"""
        _check(self.lib.pb2_problem_shift_time_values(self.prob))

    # ---- data ---------------------------------------------------------------------------------
    @staticmethod
    def _dp(a):
        return a.ctypes.data_as(c_double_p)

    def set_nodal_values(self, t: int, values: np.ndarray):
        v = np.ascontiguousarray(values, dtype=np.float64)
        assert v.shape == (self.mesh.n_node, self.info.nval)
        _check(self.lib.pb2_problem_set_nodal_values(self.prob, t, self._dp(v)))

    def set_nodal_positions(self, t: int, pos: np.ndarray):
        v = np.ascontiguousarray(pos, dtype=np.float64)
        _check(self.lib.pb2_problem_set_nodal_positions(self.prob, t, self._dp(v)))

    def set_lagrangian_positions(self, pos: np.ndarray):
        v = np.ascontiguousarray(pos, dtype=np.float64)
        _check(self.lib.pb2_problem_set_lagrangian_positions(self.prob, self._dp(v)))

    def set_dofs(self, dofs: np.ndarray, t: int = 0):
        """Problem.set_current_dofs (t = 0) / set_history_dofs(t) on the packed data: dof vector -> nodal values / positions of level t"""
        v = np.ascontiguousarray(dofs, dtype=np.float64)
        assert v.shape == (self.n_dof,)
        if t == 0:
            _check(self.lib.pb2_problem_set_dofs(self.prob, self._dp(v)))
        else:
            _check(self.lib.pb2_problem_set_history_dofs(self.prob, t, self._dp(v)))

    def set_parameters(self, **values: float):
        for k, v in values.items():
            self._params[self.param_names.index(k)] = v
        _check(self.lib.pb2_problem_set_parameters(self.prob, self._dp(self._params), self.info.n_params))

    def set_steady(self):
        """oomph steady solve: all time weights zero, ntstorage 0 (src/elements.cpp:4583-4596)."""
        self.ti = TimeInfo()
        _check(self.lib.pb2_problem_set_time(self.prob, ctypes.byref(self.ti)))
        for c in self.children:
            c.set_steady()

    def set_unsteady(self, t: float, dt: float, dtprev: float, unsteady_steps_done: int, ntstorage: Optional[int] = None):
        """MultiTimeStepper weights (BDF1, BDF2, Newmark2; src/timestepper.cpp:31-80) + the _degr selection rule of
        prepare_shape_buffer_for_integration (src/elements.cpp:4603-4626).  ntstorage: history levels the time sums run over; default =
        the levels the element class reads (3 for BDF2, NSTEPS+3 = 5 when a second time derivative brings the Newmark slots in)."""
        ti = TimeInfo()
        w1, w2 = bdf_weights(dt, dtprev)
        n1, n2 = newmark2_weights(dt)
        ti.t[0], ti.t[1], ti.t[2] = t, t - dt, t - dt - dtprev
        ti.dt[0], ti.dt[1] = dt, dtprev
        for i in range(PB2_NTW):
            ti.w_dt_BDF1[i], ti.w_dt_BDF2[i] = w1[i], w2[i]
            ti.w_dt_Newmark2[i], ti.w_d2t_Newmark2[i] = n1[i], n2[i]
            degr = w1[i] if unsteady_steps_done == 0 else w2[i]
            ti.w_dt_BDF2_degr[i] = degr
            ti.w_dt_Newmark2_degr[i] = degr if unsteady_steps_done <= 4 else n1[i]
        if ntstorage is None:
            ntstorage = max(3, int(self.info.n_hist_val))
        if ntstorage > max(int(self.info.n_hist_val), int(self.info.n_hist_pos)) and ntstorage > 3:
            raise ValueError("ntstorage %d exceeds the history levels the element class stores" % ntstorage)
        ti.ntstorage = ntstorage
        self.ti = ti
        _check(self.lib.pb2_problem_set_time(self.prob, ctypes.byref(ti)))
        for c in self.children:
            c.set_unsteady(t, dt, dtprev, unsteady_steps_done, ntstorage)

    # ---- assembly -----------------------------------------------------------------------------
    def assemble(self, flag: int = 1, residual: str = "", parameter: Optional[str] = None, stream: int = 0):
        """Device-resident assembly (results stay in HBM); use fetch() or device_outputs()."""
        self._fresh()
        ri = self.residual_names.index(residual)
        pi = -1 if parameter is None else self.param_names.index(parameter)
        _check(self.lib.pb2_problem_assemble(self.prob, ri, pi, flag, ctypes.c_void_p(stream)))
        self._last_flag = flag
        # the element classes attached to this one add to the matrix just written (same stream: ordered after it); a class that
        # does not know the residual / parameter contributes nothing
        for c in self.children:
            if residual in c.residual_names and (parameter is None or parameter in c.param_names):
                c.assemble(flag=flag, residual=residual, parameter=parameter, stream=stream)

    def fetch(self, want_jacobian: bool = True, want_mass: bool = False):
        res = np.empty(self.n_dof)
        jac = np.empty(self.nnz) if want_jacobian else None
        mass = np.empty(self.nnz) if want_mass else None
        _check(self.lib.pb2_problem_fetch(self.prob, self._dp(res), None if jac is None else self._dp(jac),
                                          None if mass is None else self._dp(mass)))
        return res, jac, mass

    def assemble_host(self, dofs: Optional[np.ndarray], flag: int = 1, residual: str = "", parameter: Optional[str] = None,
                      out: Optional[Tuple[np.ndarray, Optional[np.ndarray], Optional[np.ndarray]]] = None):
        """The reference-facing call: host dof vector in, host residual / CSR values out (copies included)."""
        self._fresh()
        ri = self.residual_names.index(residual)
        pi = -1 if parameter is None else self.param_names.index(parameter)
        if out is None:
            out = (np.empty(self.n_dof), np.empty(self.nnz) if flag >= 1 else None, np.empty(self.nnz) if flag >= 2 else None)
        res, jac, mass = out
        d = None if dofs is None else self._dp(np.ascontiguousarray(dofs, dtype=np.float64))
        _check(self.lib.pb2_problem_assemble_host(self.prob, ri, pi, flag, d, self._dp(res),
                                                  None if jac is None else self._dp(jac), None if mass is None else self._dp(mass)))
        return out

    # ---- Hessian-vector products (MultiAssembleRequest.dJdU / dMdU, pyoomph/generic/bifurcation_tools.py:465-531) -----
    def assemble_hessian(self, Y: np.ndarray, flag: int = 2, residual: str = "", transposed: bool = False):
        """d(J.Y_v)/dU (flag 1) and d(M.Y_v)/dU (flag 2) for every row Y_v of Y; transposed: d(J^T.Y_v)/dU and d(M^T.Y_v)/dU
        (the reference's flags 4 and 5, src/jitbridge.h:637-691); returns lists of CSR value arrays"""
        self._fresh()
        Y = np.ascontiguousarray(np.atleast_2d(Y), dtype=np.float64)
        assert Y.shape[1] == self.n_dof
        ri = self.residual_names.index(residual)
        _check(self.lib.pb2_problem_assemble_hessian(self.prob, ri, flag + (3 if transposed else 0), Y.shape[0], self._dp(Y), None))
        J, M = [], []
        for v in range(Y.shape[0]):
            jv = np.empty(self.nnz)
            mv = np.empty(self.nnz) if flag >= 2 else None
            _check(self.lib.pb2_problem_fetch_hessian(self.prob, v, self._dp(jv), None if mv is None else self._dp(mv)))
            J.append(jv); M.append(mv)
        return J, M

    def hessian_vector_products(self, Y: np.ndarray, C: np.ndarray, residual: str = "") -> np.ndarray:
        """flag 0 of HessianVectorProduct: out[v][i] = sum_jk Y_j H_ijk C_vk"""
        self._fresh()
        Y = np.ascontiguousarray(Y, dtype=np.float64).ravel()
        C = np.ascontiguousarray(np.atleast_2d(C), dtype=np.float64)
        out = np.empty_like(C)
        ri = self.residual_names.index(residual)
        _check(self.lib.pb2_problem_hessian_vector_products(self.prob, ri, self._dp(Y), self._dp(C), C.shape[0], self._dp(out)))
        return out

    def assemble_hessian_tensor(self, symmetric: bool = False, residual: str = ""):
        """Problem::assemble_hessian_tensor (src/problem.cpp:1530-1560): the global rank-3 tensor as a SparseRank3Tensor"""
        from .hessian_tensor import assemble_hessian_tensor
        return assemble_hessian_tensor(self, symmetric, residual)

    # ---- eigenproblem matrices (Problem::assemble_eigenproblem_matrices, src/problem.cpp:715 -> oomph EigenProblemHandler) --------
    def assemble_eigenproblem_matrices(self, sigma_r: float = 0.0, residual: str = ""):
        """(M, J - sigma_r*M) from ONE flag-2 launch, as scipy CSR matrices over the fixed pattern; the shift is applied to the
        assembled values (the reference shifts every element matrix, oomph-lib assembly_handler.cc:369-382: same sums)."""
        from scipy.sparse import csr_matrix
        self.assemble(flag=2, residual=residual)
        _, jac, mass = self.fetch(True, True)
        if sigma_r != 0.0:
            jac = jac - sigma_r * mass
        n = self.n_dof
        return csr_matrix((mass, self.indices, self.indptr), shape=(n, n)), csr_matrix((jac, self.indices, self.indptr), shape=(n, n))

    def integral_gradient(self, name: str) -> np.ndarray:
        """c_j = d(integral expression `name`)/dU_j at the current state: the residual vector of the contribution "d_integral_<name>"
        (registered with ``add_integral_function(..., with_gradient=True)``): the dense row of a global constraint in bordered form"""
        rn = self.code.INTEGRAL_GRADIENT_PREFIX + name
        if rn not in self.residual_names:
            raise RuntimeError("integral expression '%s' was registered without its gradient" % name)
        self.assemble(flag=0, residual=rn)
        r, _, _ = self.fetch(False, False)
        return r

    def assemble_azimuthal_eigenproblem_matrices(self, m: float, sigma_r: float = 0.0, residual: str = ""):
        """Complex (M, J - sigma_r M) of the azimuthal mode m about the current axisymmetric base state: two flag-2 launches (real and
        imaginary contribution, pyoomph/generic/problem.py:4543-4558 and the normal-mode eigensolve of problem.py) combined as
        J = J_re + i J_im, M = M_re + i M_im on the fixed pattern.  The element class must have been generated with an
        ``AxisymmetryBreakingCoordinateSystem`` whose mode is the global parameter ``azimuthal_m``."""
        from scipy.sparse import csr_matrix
        cs = self.code.coordinate_system
        if not getattr(cs, "has_normal_mode_expansion", False):
            raise RuntimeError("the element class has no azimuthal mode expansion")
        if "azimuthal_m" in self.param_names:
            self.set_parameters(azimuthal_m=float(m))
        parts = []
        for prefix in (cs.real_contribution_name, cs.imag_contribution_name):
            name = prefix + residual
            if name in self.residual_names:
                self.assemble(flag=2, residual=name)
                _, jac, mass = self.fetch(True, True)
            else:                       # a purely real operator has no imaginary contribution
                jac, mass = np.zeros(self.nnz), np.zeros(self.nnz)
            parts.append((jac, mass))
        jac = parts[0][0] + 1j * parts[1][0]
        mass = parts[0][1] + 1j * parts[1][1]
        if sigma_r != 0.0:
            jac = jac - sigma_r * mass
        n = self.n_dof
        return csr_matrix((mass, self.indices, self.indptr), shape=(n, n)), csr_matrix((jac, self.indices, self.indptr), shape=(n, n))

    # ---- integral expressions (Mesh.evaluate_observable -> BulkElementBase::eval_integral_expression, src/mesh.cpp:545) ---------
    def evaluate_integral_expressions(self) -> Dict[str, float]:
        """all integral expressions of the element class over all elements: one launch for the per-element values and one
        fixed-order reduction (bit-reproducible), {name: value}"""
        if not self.integral_names:
            raise RuntimeError("this element class defines no integral expressions")
        self._fresh()
        out = np.empty(len(self.integral_names))
        _check(self.lib.pb2_problem_eval_integrals(self.prob, self._dp(out), len(self.integral_names)))
        return {n: float(v) for n, v in zip(self.integral_names, out)}

    # ---- expressions at local coordinates: local expressions, extremum expressions, Z2 fluxes (src/codegen.cpp:4366-4453) -------------
    def evaluate_point_expressions(self, points: str = "nodes") -> np.ndarray:
        """all point expressions of the class at the integration points ("gauss") or the nodes ("nodes") of every element:
        array [n_elem, n_points, n_expressions] in the mesh's element order (expressions as in ``point_names``); one launch"""
        if not self.point_names:
            raise RuntimeError("this element class defines no local / extremum expressions or Z2 fluxes")
        self._fresh()
        ps = {"gauss": 0, "nodes": 1}[points]
        npts = int(self.info.n_int_pt if ps == 0 else self.info.nnode)
        out = np.empty((self.n_elem, npts, len(self.point_names)))
        _check(self.lib.pb2_problem_eval_points(self.prob, ps, self._dp(out), ctypes.c_longlong(out.size)))
        return out

    def evaluate_local_expressions_at_nodes(self) -> Dict[str, np.ndarray]:
        """BulkElementBase::eval_local_expression_at_node for every element and node (Mesh output): {name: [n_elem, nnode]}"""
        vals = self.evaluate_point_expressions("nodes")
        return {n: vals[:, :, i].copy() for i, (k, n) in enumerate(self.point_names) if k == "local"}

    def get_Z2_fluxes(self, points: str = "gauss") -> np.ndarray:
        """GetZ2Fluxes at the integration points (or nodes) of every element: [n_elem, n_points, num_Z2_flux_terms]"""
        vals = self.evaluate_point_expressions(points)
        idx = [i for i, (k, _) in enumerate(self.point_names) if k == "z2"]
        return vals[:, :, idx].copy()

    def evaluate_extremum(self, name: str, sign: int = 1):
        """Mesh::evaluate_extremum (src/mesh.cpp:444-500) without its final Newton refinement: the largest value of sign * expression
        over the integration points and the nodes of all elements, scanned in the reference's order (elements in mesh order, per element
        first the integration points, then the nodes; a later sample wins only if strictly larger).  Returns (value, element,
        ("gauss" | "nodes", point index))."""
        names = [n for k, n in self.point_names]
        kinds = [k for k, n in self.point_names]
        if name not in names or kinds[names.index(name)] != "extremum":
            raise RuntimeError("Extremum function " + name + " not defined on this mesh")
        i = names.index(name)
        g = sign * self.evaluate_point_expressions("gauss")[:, :, i]
        nd = sign * self.evaluate_point_expressions("nodes")[:, :, i]
        samples = np.concatenate([g, nd], axis=1)                       # per element: integration points, then nodes
        start = nd[0, 0]                                                  # the reference starts from node 0 of element 0
        flat = samples.ravel()
        k = int(np.argmax(flat))                                          # first occurrence of the maximum = the reference's strict ">" scan
        if flat[k] > start:
            e, q = divmod(k, samples.shape[1])
            where = ("gauss", q) if q < g.shape[1] else ("nodes", q - g.shape[1])
            return float(sign * flat[k]), e, where
        return float(sign * start), 0, ("nodes", 0)

    def evaluate_observable(self, name: str) -> float:
        return self.evaluate_integral_expressions()[name]

    def newton_step_on_device(self, solver, residual: str = ""):
        """one Newton iteration with a device-resident linear solver (pyoomph_b200.solvers.DeviceLinearSystemSolver): the matrix never
        crosses the host link (SURVEY N-d).  Returns (max |residual| before the step, solver statistics)."""
        from .solvers import newton_step_on_device
        self._fresh()
        return newton_step_on_device(self, solver, residual)

    def fetch_dofs(self) -> np.ndarray:
        """the engine's device dof vector (as set by set_dofs / updated by newton_step_on_device) back on the host"""
        import torch
        from .solvers import _DeviceArray
        d = ctypes.c_void_p()
        _check(self.lib.pb2_problem_device_dofs(self.prob, ctypes.byref(d)))
        t = torch.as_tensor(_DeviceArray(d.value, self.n_dof, "<f8"), device=torch.device("cuda", self._device))
        return t.cpu().numpy()

    def device_outputs(self):
        r, j, m = c_double_p(), c_double_p(), c_double_p()
        _check(self.lib.pb2_problem_device_outputs(self.prob, ctypes.byref(r), ctypes.byref(j), ctypes.byref(m)))
        return (ctypes.cast(r, ctypes.c_void_p).value, ctypes.cast(j, ctypes.c_void_p).value, ctypes.cast(m, ctypes.c_void_p).value)

    def pack_rows(self, rows_ptr: int, n_rows: int, pos_ptr: int, n_pos: int, flag: int, buf_ptr: int, stream=None):
        """interface exchange, sender side: residual[rows] | jac[pos] (| mass[pos]) -> buf (all device pointers)"""
        _check(self.lib.pb2_problem_pack_rows(self.prob, ctypes.c_void_p(rows_ptr), ctypes.c_longlong(n_rows), ctypes.c_void_p(pos_ptr),
                                              ctypes.c_longlong(n_pos), ctypes.c_uint(flag), ctypes.c_void_p(buf_ptr), ctypes.c_void_p(stream or 0)))

    def unpack_add(self, rows_ptr: int, n_rows: int, pos_ptr: int, n_pos: int, flag: int, buf_ptr: int, stream=None):
        """interface exchange, owner side: residual[rows] += ..., jac[pos] += ... from a received buffer"""
        _check(self.lib.pb2_problem_unpack_add(self.prob, ctypes.c_void_p(rows_ptr), ctypes.c_longlong(n_rows), ctypes.c_void_p(pos_ptr),
                                               ctypes.c_longlong(n_pos), ctypes.c_uint(flag), ctypes.c_void_p(buf_ptr), ctypes.c_void_p(stream or 0)))

    def host_maps(self):
        """pattern-only problems (device=-1): (perm, elem_rowstart[n_elem, ndof], elem_off[n_elem, ndof, ndof], elem_res[n_elem, ndof]) as the
        kernels read them, elements in scheduled order (perm[q] = mesh element)"""
        perm, rs, off, res, bits = c_int_p(), c_int_p(), ctypes.c_void_p(), c_int_p(), ctypes.c_int()
        _check(self.lib.pb2_problem_host_maps(self.prob, ctypes.byref(perm), ctypes.byref(rs), ctypes.byref(off), ctypes.byref(bits), ctypes.byref(res)))
        ne, nd = self.n_elem, int(self.info.ndof_el)
        if ne == 0:
            return np.zeros(0, np.int32), np.zeros((0, nd), np.int32), np.zeros((0, nd, nd), np.uint8), np.zeros((0, nd), np.int32)
        ty = ctypes.c_uint8 if bits.value == 8 else ctypes.c_uint16
        offa = np.ctypeslib.as_array(ctypes.cast(off, ctypes.POINTER(ty)), shape=(ne, nd, nd)).copy()
        return (np.ctypeslib.as_array(perm, shape=(ne,)).copy(), np.ctypeslib.as_array(rs, shape=(ne, nd)).copy(), offa,
                np.ctypeslib.as_array(res, shape=(ne, nd)).copy())

    def launch_count(self) -> int:
        return int(self.lib.pb2_problem_launch_count(self.prob)) + sum(c.launch_count() for c in self.children)

    def num_colours(self) -> int:
        return int(self.lib.pb2_problem_num_colours(self.prob))

    def num_launches(self) -> int:
        """kernel launches per assembly = element chunks x colours"""
        return int(self.lib.pb2_problem_num_launches(self.prob))

    # ---- CustomAssemblyBase contract -----------------------------------------------------------
    def sync_from_problem(self):
        """Bring the packed data up to date with the host problem this assembler was attached to by Problem.set_custom_assembler
        (pyoomph/generic/problem.py:1711-1721): current dofs and history dofs (Problem.get_current_dofs / get_history_dofs,
        src/pybind/problem.cpp:518-527), time-stepper state (Time::time/dt, MultiTimeStepper weights by their inputs dt, dtprev and
        get_num_unsteady_steps_done, src/timestepper.hpp:59) and global parameters (Problem.get_global_parameter(name).value).
        Called by get_residuals_and_jacobian before every assembly: Newton updates, time shifts and continuation steps of the host
        are what the kernels see.  Returns the current dof vector."""
        pr = self.problem
        dofs = np.asarray(pr.get_current_dofs()[0], dtype=np.float64)
        if dofs.shape != (self.n_dof,):
            raise RuntimeError("the problem has %d dofs, the packed assembler %d: renumbered without rebuild()?" % (dofs.size, self.n_dof))
        ts = pr.timestepper
        steady = bool(ts.is_steady())
        if steady:
            self.set_steady()
        else:
            tp = pr.time_pt()
            self.set_unsteady(float(tp.time()), float(tp.dt(0)), float(tp.dt(1)), int(ts.get_num_unsteady_steps_done()))
            for t in range(1, max(int(self.info.n_hist_val), int(self.info.n_hist_pos))):
                self.set_dofs(np.asarray(pr.get_history_dofs(t), dtype=np.float64), t)
        if self.param_names:
            self.set_parameters(**{n: float(pr.get_global_parameter(n).value) for n in self.param_names})
        return dofs

    def get_residuals_and_jacobian(self, require_jacobian: bool, dparameter: Optional[str] = None):
        """CustomAssemblyBase contract (pyoomph/generic/assembly.py:83, called from Problem.get_custom_residuals_jacobian,
        problem.py:1727-1745).  Attached to a problem, the current state of the problem is pulled first (sync_from_problem); detached
        (self.problem is None: tests, benchmarks), the packed data is used as the set_* calls left it."""
        dofs = self.sync_from_problem() if self.problem is not None else None
        if require_jacobian:
            if dparameter:
                raise RuntimeError("Cannot derive custom Jacobian with respect to a parameter yet")  # problem.py:1733
            res, jac, _ = self.assemble_host(dofs, 1)
            from scipy.sparse import csr_matrix
            return res, csr_matrix((jac, self.indices, self.indptr), shape=(self.n_dof, self.n_dof))
        res, _, _ = self.assemble_host(dofs, 0, parameter=dparameter)
        return res

    def close(self):
        for c in getattr(self, "children", []):       # they alias this problem's device buffers
            c._release()
        if getattr(self, "prob", None):
            self.lib.pb2_problem_free(self.prob)
            self.prob = None
        if getattr(self, "cls", None):
            self.lib.pb2_class_free(self.cls)
            self.cls = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
