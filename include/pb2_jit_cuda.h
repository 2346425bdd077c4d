/* pb2_jit_cuda.h -- the GPU extension of pyoomph's generated-code plugin contract.
 *
 * pyoomph's CPU plugins export one symbol, JIT_ELEMENT_init(JITFuncSpec_Table_FiniteElement_t*)
 * (/root/reference/src/jitbridge.h:499, emitted at src/codegen.cpp:6401), which fills a table of
 * per-element function pointers (ResidualAndJacobian[i], ParameterDerivative[i][p], HessianVectorProduct[i],
 * jitbridge.h:412,446-451) called once per element by the host (src/elements.cpp:5112).
 *
 * A CUDA plugin (compiled by nvcc for sm_100a from the code pyoomph_b200.cuda_emitter writes) exports, next
 * to that, JIT_ELEMENT_init_cuda(pb2_cuda_table_t*): metadata of the element class plus *batched* launchers
 * that run one routine over a whole colour of elements.  The per-Gauss-point host callback
 * fill_shape_buffer_for_point (jitbridge.h:495) does not exist on this path: geometry is computed in-kernel.
 * Everything is plain C: pointers and sizes only.
 */
#ifndef PB2_JIT_CUDA_H
#define PB2_JIT_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define PB2_ABI_VERSION 11
#define PB2_MAX_PARAMS 16
#define PB2_NTW 7           /* time-stepper storage, MultiTimeStepper (src/timestepper.hpp:37-45) */
#define PB2_MAX_FIELDS 16
#define PB2_MAX_ROUTINES 64
#define PB2_MAX_HVEC 4      /* vectors per Hessian-vector launch */
#define PB2_MAX_INTEGRALS 16 /* integral expressions per element class */
#define PB2_MAX_POINT_EXPRS 32 /* local + extremum expressions + Z2 flux terms per element class */

/* what one launch writes (the `flag` of jitbridge.h:285 routines) */
#define PB2_FLAG_RESIDUAL 0u
#define PB2_FLAG_JACOBIAN 1u
#define PB2_FLAG_MASS 2u

/* residual position map encoding: >=0 accumulate at that position, <0 (but not SKIP) first touch: store at ~v */
#define PB2_MAP_SKIP (-2147483647 - 1)

/* error word written by the kernels (pb2_kernel_args.status) */
#define PB2_STATUS_GATE_TIMEOUT 1
#define PB2_GATE_TIMEOUT_NS 20000000000ull

/* time information = what prepare_shape_buffer_for_integration copies per element (src/elements.cpp:4577-4646);
 * here it is one constant block per launch.  The *_degr aliases are resolved by the host. */
typedef struct pb2_time_info
{
  double t[PB2_NTW], dt[PB2_NTW];
  double w_dt_BDF1[PB2_NTW], w_dt_BDF2[PB2_NTW], w_dt_Newmark2[PB2_NTW], w_d2t_Newmark2[PB2_NTW];
  double w_dt_BDF2_degr[PB2_NTW], w_dt_Newmark2_degr[PB2_NTW];
  int ntstorage; /* 0 when steady: history sums are skipped (src/elements.cpp:4585) */
  int pad_;
} pb2_time_info;

/* kernel argument block, passed by value (lands in the constant bank) */
typedef struct pb2_kernel_args
{
  int n_elem;                 /* elements of this launch (one colour) */
  int elem_begin;             /* first element of the launch in the colour-major element arrays */
  const int *elem_nodes;      /* [n_elem_total][nnode]            element -> node, oomph local order */
  const int *elem_eqn;        /* [n_elem_total][ndof_el]          local dof -> global equation, <0 pinned */
  const int *elem_rowstart;   /* [n_elem_total][ndof_el]          CSR position of the first entry of each local row, <0 pinned */
  const void *elem_off;       /* [n_elem_total][ndof_el*ndof_el]  (row,col) -> offset in that row | first-touch bit; all ones: skip */
  int map_bits, pad0_;        /* 8 or 16 bits per elem_off entry */
  const int *elem_res;        /* [n_elem_total][ndof_el]          residual position, same encoding */
  const double *node_pos;     /* [n_hist_pos][n_node][dim] */
  const double *node_lagr;    /* [n_node][dim] */
  const double *node_val;     /* [n_hist_val][n_node][nval] */
  long long n_node;
  int n_hist_val, n_hist_pos;
  double *residual;           /* [n_dof] */
  double *jac_vals;           /* [nnz] */
  double *mass_vals;          /* [nnz] */
  /* pipelined kernels: the whole assembly is one launch over batches ordered by tile = (chunk, colour) */
  const int *block_begin;     /* [grid+1]    range of each thread block in the batch lists below */
  const int *batch_elem;      /* [n_batches] first element of the batch (scheduled element order) */
  const int *batch_meta;      /* [n_batches] (tile << 7) | (wait-for-previous-batch << 6) | number of elements */
  const unsigned long long *batch_bar; /* [n_batches] bit i: the scatter warps synchronise before element i of the batch (it shares
                                          CSR entries with an element of the same batch scattered since the last barrier) */
  const int *tile_nbatch;     /* [n_tiles]   batches per tile */
  int *tile_done;             /* [n_tiles]   completion counters, zeroed by the host before the launch */
  int n_batches, n_tiles;
  unsigned long long *debug;  /* NULL, or [64] cycle counters filled by kernels built with PB2_TIMING=1 (development aid) */
  int *status;                /* [1] device-visible error word, 0 = fine.  PB2_STATUS_GATE_TIMEOUT: a tile gate of the persistent kernel was
                                 not opened within PB2_GATE_TIMEOUT_NS (blocks not co-resident / a lost block); the waiting warps give up,
                                 the kernel ends, results are invalid and the host reports the error (never a silent hang) */
  const double *hvec;         /* [n_hvec][n_dof]   Hessian-vector inputs (or NULL) */
  int n_hvec, pad_;
  double *integrals;          /* [n_elem_total][n_integrals]  per-element values of the integral expressions (kind 2 kernels) */
  pb2_time_info ti;
  double params[PB2_MAX_PARAMS]; /* global parameters (jitbridge.h:410) by value */
} pb2_kernel_args;

typedef struct pb2_class_info
{
  int abi_version;
  char name[64];
  int nodal_dim, elem_dim;
  int nnode, nnode_C1;
  int c1_nodes[8];
  int n_int_pt;
  int nval;                        /* nodal values per node (all continuous fields) */
  int n_fields;
  char field_names[PB2_MAX_FIELDS][48];
  int field_space[PB2_MAX_FIELDS]; /* 2: C2, 1: C1 */
  int field_index[PB2_MAX_FIELDS]; /* index into the nodal value record */
  int moving_nodes;                /* coordinates are dofs */
  int ndof_el;                     /* local dofs per element incl. pinned slots */
  /* local dof layout: for dof k: node (element-local), kind 0: position dim `index`, 1: nodal value `index` */
  int dof_node[160], dof_kind[160], dof_index[160];
  int n_residuals;
  char residual_names[8][48];
  int n_params;
  char param_names[PB2_MAX_PARAMS][48];
  int n_hist_val, n_hist_pos;      /* history levels the kernels read */
  int max_dt_order;
  int elems_per_block, threads_per_block, smem_bytes;
  int hessian_generated;
  double alg_bytes_per_elem[3];    /* algorithmic bytes per element for flag 0,1,2 */
  double flops_per_elem[3];        /* fp64 flops per element counted from the emitted code */
  double alg_bytes_per_hist_level; /* part of alg_bytes per history level beyond the current one (not read when steady) */
  /* integral expressions = numintegral_expressions / integral_expressions_names of jitbridge.h:417-418 */
  int n_integrals, pad_;
  char integral_names[PB2_MAX_INTEGRALS][48];
  /* expressions evaluated at a local coordinate of every element (kind 4 kernels): numlocal_expressions / numextremum_expressions /
   * num_Z2_flux_terms of jitbridge.h:420-424, :456 in one list; point_kind 0 local, 1 extremum, 2 Z2 flux */
  int n_point_exprs, pad2_;
  char point_names[PB2_MAX_POINT_EXPRS][48];
  int point_kind[PB2_MAX_POINT_EXPRS];
} pb2_class_info;

/* launch configuration of one generated routine */
typedef struct pb2_kernel_cfg
{
  int elems_per_batch;   /* elements one block processes per batch */
  int threads;           /* block size */
  int smem_bytes;        /* dynamic shared memory */
  int pipelined;         /* 1: persistent warp-specialised kernel, one launch per assembly (batch tables in args);
                            0: one launch per tile, args->elem_begin / n_elem select the tile */
  int blocks_per_sm;     /* resident blocks per SM (occupancy), for sizing the persistent grid */
  int pad_;
  const void *func;      /* opaque kernel handle for pb2_launch_fn */
} pb2_kernel_cfg;

/* routine = ResidualAndJacobian<residual_index> (param_index < 0) or dResidual<i>dParameter_<param_index>;
 * flag as in jitbridge.h:285.  kind 0: R/J/M routines, 1: Hessian-vector routines d(J.Y)/dU (flag 1) + d(M.Y)/dU (flag 2), 3: their
 * TRANSPOSED contractions d(J^T.Y)/dU, d(M^T.Y)/dU (flags 4 / 5 of HessianVectorProduct, jitbridge.h:637-691; same flag values 1 / 2
 * here), 2: EvalIntegralExpression for ALL integral
 * expressions at once (jitbridge.h:469; one launch per call, args->elem_begin / n_elem select the elements, args->integrals
 * receives the per-element values; residual_index, param_index and flag are ignored).  kind 4: EvalLocalExpression /
 * EvalExtremumExpression / GetZ2Fluxes (jitbridge.h:470-471, :458) for ALL point expressions of the class at the points of one set --
 * flag 0: the integration points, flag 1: the element's nodes -- args->integrals receives [element][point][expression]. */
typedef int (*pb2_query_fn)(int kind, int residual_index, int param_index, unsigned flag, pb2_kernel_cfg *out);
typedef int (*pb2_launch_fn)(const pb2_kernel_cfg *cfg, const pb2_kernel_args *args, int grid, void *cuda_stream);

typedef struct pb2_cuda_table
{
  pb2_class_info info;
  pb2_query_fn query;
  pb2_launch_fn launch;
} pb2_cuda_table_t;

typedef void (*JIT_ELEMENT_init_cuda_SPEC)(pb2_cuda_table_t *table);

#ifdef __cplusplus
}
#endif
#endif
