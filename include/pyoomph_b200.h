/* pyoomph_b200.h -- C-ABI of the B200 element-assembly engine (libpyoomph_b200.so).
 *
 * Plain pointers and sizes only; no C++/torch types.  Each entry point cites the reference interface it
 * replaces (paths relative to /root/reference).  Every function returns 0 on success, non-zero on error with a
 * message retrievable through pb2_last_error(); nothing falls back to the CPU.
 *
 * Call sequence (mirrors Problem.initialise -> assign equation numbers -> get_jacobian):
 *   pb2_class_load        <- DynamicBulkElementCode ctor + CCompiler::get_init_func   src/problem.cpp:101-142, src/ccompiler.cpp:191-243
 *   pb2_problem_create    <- fill_element_info + assign_eqn_numbers consumers         src/elements.cpp:2713, oomph mesh.cc:686
 *   pb2_problem_set_*     <- nodal Data values / positions / history, Time, global parameters
 *   pb2_problem_assemble  <- Problem::get_jacobian / get_residuals                    src/problem.cpp:902-968 -> oomph problem.cc:5332-5666
 *   pb2_problem_fetch     <- the CSR triple handed to the linear solver               src/pybind/problem.cpp:572-602
 */
#ifndef PYOOMPH_B200_H
#define PYOOMPH_B200_H

#include <stddef.h>
#include "pb2_jit_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pb2_class pb2_class;
typedef struct pb2_problem pb2_problem;

/* mesh + numbering of one element class, host arrays (the data fill_element_info reads per element) */
typedef struct pb2_mesh_desc
{
  long long n_elem, n_node;
  const int *elem_nodes;   /* [n_elem][nnode] global node numbers in oomph local order */
  const int *node_eqn;     /* [n_node][nval] global equation per nodal value, <0 pinned (Data::eqn_number) */
  const int *pos_eqn;      /* [n_node][dim] equations of nodal positions (SolidNode), NULL if mesh is fixed */
  long long n_dof;         /* global number of equations */
  const int *elem_patch;   /* [n_elem] id of the compact patch (~64 neighbouring elements) an element belongs to, or NULL:
                              then consecutive elements in mesh order form the patches.  Only a locality hint. */
  /* multi-GPU: entries (extra_rows[i], extra_cols[i]) are added to the CSR pattern although no local element produces
   * them -- the columns that neighbouring ranks contribute to interface rows owned by this rank (problem.cc:6970-7121).
   * n_extra = 0 on a single GPU. */
  long long n_extra;
  const int *extra_rows, *extra_cols;
} pb2_mesh_desc;

int pb2_version(void);
const char *pb2_last_error(void);

/* load a compiled CUDA plugin (.so exporting JIT_ELEMENT_init_cuda).  Replaces dlopen+dlsym("JIT_ELEMENT_init")
 * of src/ccompiler.cpp:191-243. */
int pb2_class_load(const char *so_path, pb2_class **out);
int pb2_class_get_info(const pb2_class *cls, pb2_class_info *out);
void pb2_class_free(pb2_class *cls);

/* build colouring, the fixed CSR pattern (columns ascending per row, like the "maps" assembly of
 * src/problem.cpp:2200-2274) and the element->CSR position maps; upload everything to `device`. */
int pb2_problem_create(pb2_class *cls, int device, const pb2_mesh_desc *mesh, pb2_problem **out);
/* device < 0 creates a PATTERN-ONLY problem: schedule, CSR pattern and position maps on the host, no CUDA call at all; only
 * pb2_problem_pattern / num_colours / num_launches / setup_seconds / free accept it (everything that computes reports an error). */
void pb2_problem_free(pb2_problem *p);

/* pattern of the assembled matrix (host memory owned by the problem): CRDoubleMatrix row_start/column_index */
int pb2_problem_pattern(pb2_problem *p, const int **row_start, const int **column_index, long long *nnz, long long *n_rows);
/* pattern-only problems: the schedule (perm[q] = mesh element at scheduled position q) and the element -> CSR position maps in the
 * layout the kernels read (pb2_kernel_args: elem_rowstart, elem_off with map_bits, elem_res), host memory owned by the problem */
int pb2_problem_host_maps(pb2_problem *p, const int **perm, const int **elem_rowstart, const void **elem_off, int *map_bits, const int **elem_res);
int pb2_problem_num_colours(pb2_problem *p);
int pb2_problem_num_launches(pb2_problem *p); /* tiles (= patch colours, device-side gates) one persistent launch walks through */

/* nodal data, host -> device.  t = history index (0 current). */
int pb2_problem_set_nodal_values(pb2_problem *p, int t, const double *values /*[n_node][nval]*/);
int pb2_problem_set_nodal_positions(pb2_problem *p, int t, const double *pos /*[n_node][dim]*/);
int pb2_problem_set_lagrangian_positions(pb2_problem *p, const double *pos /*[n_node][dim]*/);
/* Problem::set_dofs equivalent: scatter a global dof vector (host) into current nodal values / positions on the device */
int pb2_problem_set_dofs(pb2_problem *p, const double *dofs /*[n_dof]*/);
/* Problem::set_history_dofs(t, dofs) equivalent (src/pybind/problem.cpp:540): the same scatter into history level t */
int pb2_problem_set_history_dofs(pb2_problem *p, int t, const double *dofs /*[n_dof]*/);
/* Problem::shift_time_values on the packed data (device to device): history level t <- level t-1 for the nodal values and, on
 * moving meshes, the nodal positions; level 0 keeps the current values.  Replaces a re-upload of every history level per time step. */
int pb2_problem_shift_time_values(pb2_problem *p);
int pb2_problem_set_time(pb2_problem *p, const pb2_time_info *ti);
int pb2_problem_set_parameters(pb2_problem *p, const double *values, int n);

/* one assembly on the device; results stay in HBM.  flag: 0 residual, 1 +Jacobian, 2 +mass matrix (jitbridge.h:285).
 * param_index >= 0 selects dResidual<i>dParameter_<p>.  `cuda_stream` may be NULL (default stream). */
int pb2_problem_assemble(pb2_problem *p, int residual_index, int param_index, unsigned flag, void *cuda_stream);
/* device pointers of the outputs (for GPU-side consumers) */
int pb2_problem_device_outputs(pb2_problem *p, double **residual, double **jac_vals, double **mass_vals);
/* Device-resident hand-off to a GPU linear solver (SURVEY N-d; the reference's solver plugins receive host arrays,
 * pyoomph/solvers/generic.py:64-118, src/pybind/solver.cpp:104-140): the fixed CSR pattern as device arrays (uploaded on first use)
 * next to pb2_problem_device_outputs' values and residual, the device copy of the dof vector, and the Newton update applied on the
 * device (dofs += alpha * delta, then the scatter of pb2_problem_set_dofs).  With these a Newton iteration moves no matrix over
 * the host link. */
int pb2_problem_device_pattern(pb2_problem *p, int **row_start, int **column_index);
int pb2_problem_device_dofs(pb2_problem *p, double **dofs);
int pb2_problem_update_dofs_device(pb2_problem *p, const double *d_delta, double alpha, void *cuda_stream);
/* multi-GPU interface exchange (oomph-lib ships off-rank row contributions to their owner, problem.cc:6970-7121): one kernel packs
 * residual[rows[i]] and, for flag >= 1, jac_vals[pos[j]] (flag 2: then mass_vals[pos[j]]) into `buf` (device, n_rows + flag * n_pos
 * doubles); the owner adds a received buffer with pb2_problem_unpack_add.  rows/pos are device arrays with unique entries, so the
 * add needs no atomics and the sum order is the call order. */
int pb2_problem_pack_rows(pb2_problem *p, const long long *rows, long long n_rows, const long long *pos, long long n_pos, unsigned flag,
                          double *buf, void *cuda_stream);
int pb2_problem_unpack_add(pb2_problem *p, const long long *rows, long long n_rows, const long long *pos, long long n_pos, unsigned flag,
                           const double *buf, void *cuda_stream);
/* copy results to host buffers (any may be NULL) */
int pb2_problem_fetch(pb2_problem *p, double *residual, double *jac_vals, double *mass_vals);
/* the reference-facing call: host dofs in, host residual/CSR values out, copies included */
int pb2_problem_assemble_host(pb2_problem *p, int residual_index, int param_index, unsigned flag, const double *dofs,
                              double *residual, double *jac_vals, double *mass_vals);
/* Hessian-vector assembly = HessianVectorProduct<i> (jitbridge.h:286) without the ndof^3 buffers, as used by
 * get_multi_assembly (src/elements.cpp:4983-4988) / the Hopf and azimuthal handlers (src/bifurcation.cpp):
 *   flag 1: for each of the n_vec vectors Y_v (host, [n_vec][n_dof]) the matrix  N_v = d(J.Y_v)/dU  (row i, column k:
 *           sum_j H_ijk Y_j) on the fixed CSR pattern;   flag 2: additionally  d(M.Y_v)/dU  for the mass matrix;
 *   flag 4 / 5: the transposed contractions  sum_j H_jik Y_j  = d(J^T.Y_v)/dU  (flag 5: and d(M^T.Y_v)/dU), jitbridge.h:637-691.
 * Results stay on the device; pb2_problem_fetch_hessian copies matrix v to the host (mass_vals may be NULL). */
int pb2_problem_assemble_hessian(pb2_problem *p, int residual_index, unsigned flag, int n_vec, const double *Y, void *cuda_stream);
int pb2_problem_fetch_hessian(pb2_problem *p, int v, double *jac_hessian_vals, double *mass_hessian_vals);
/* flag 0 of HessianVectorProduct: product_v[i] = sum_jk Y_j H_ijk C_vk  = (d(J.Y)/dU) C_v  for n_vec vectors C_v (host in, host out) */
int pb2_problem_hessian_vector_products(pb2_problem *p, int residual_index, const double *Y, const double *C, int n_vec, double *products);
/* Integral expressions of the element class over ALL elements of the problem: replaces the element loop of
 * Mesh::evaluate_integral_expression -> BulkElementBase::eval_integral_expression -> functable->EvalIntegralExpression
 * (src/mesh.cpp:536-560, src/elements.cpp:4648-4657, jitbridge.h:469).  out[k] = value of expression k (order of
 * pb2_class_info.integral_names), n_out = number of expressions.  One launch for the per-element values, one fixed-order
 * reduction: the result is bit-reproducible. */
int pb2_problem_eval_integrals(pb2_problem *p, double *out, int n_out);
/* Expressions evaluated at a local coordinate of every element: replaces the per-point calls of functable->EvalLocalExpression,
 * EvalExtremumExpression and GetZ2Fluxes (BulkElementBase::eval_local_expression_at_s / eval_extremum_expression_at_s / get_Z2_flux,
 * src/elements.cpp:4666-4704, :7285-7305) behind Mesh output, Mesh::evaluate_extremum (src/mesh.cpp:444-500) and the Z2 error
 * estimator.  point_set 0: the element's integration points, 1: its nodes.  out[element][point][expression] in the MESH's element
 * order, expressions in the order of pb2_class_info.point_names; n_out = n_elem * points * expressions.  One launch. */
int pb2_problem_eval_points(pb2_problem *p, int point_set, double *out, long long n_out);
/* number of kernel launches issued by the last assemble, and cumulative */
long long pb2_problem_launch_count(pb2_problem *p);

/* utilities for hosts without CUDA bindings of their own: pinned host buffers, event timing on a stream */
void *pb2_host_alloc(size_t nbytes);
void pb2_host_free(void *ptr);
int pb2_event_record(int idx, void *cuda_stream);
int pb2_event_elapsed_ms(int idx0, int idx1, float *ms);
int pb2_device_synchronize(void);
int pb2_device_count(int *n);
int pb2_flush_l2(int device);
/* wall time pb2_problem_create spent on colouring, pattern, position maps and upload (the reference's Jacobian_setup_time includes its
 * per-assembly pattern build, oomph linear_solver.cc:986-1005: reported next to the assembly time) */
double pb2_problem_setup_seconds(pb2_problem *p);
/* fp64 roofline denominator measured on this device: DFMA throughput of a dependent-chain-free kernel, in TFLOP/s */
int pb2_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif
