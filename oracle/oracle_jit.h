/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product path.
 *
 * CPU restatement of the *subset* of pyoomph's generated-code plugin ABI that the assembly hot
 * path touches (/root/reference/src/jitbridge.h:88-120 JITElementInfo_t, :153-165 hang info,
 * :169-283 JITShapeInfo_t, :285-286 routine signatures, :298-310 required shapes,
 * :329-497 function table).  Member names are the ABI the generated code is written against, so they
 * are kept; everything the hot path never reads is left out, hence the layout here is NOT the
 * reference's.  The generated plugins and oracle/driver.c compile against either this header or
 * (with -DORACLE_USE_REFERENCE_HEADERS -I/root/reference/src, only where the reference tree exists)
 * the reference's own jitbridge.h + jitbridge_hang.h; tests/test_oracle_ref.py checks that both
 * builds give bit-identical element matrices, which pins this restatement.
 */
#ifndef ORACLE_JIT_H
#define ORACLE_JIT_H

#ifdef ORACLE_USE_REFERENCE_HEADERS
#include "jitbridge.h"
#else

#include <math.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct JITElementInfo
{
  unsigned int nnode, nnode_C1, nnode_C2, nnode_C1TB, nnode_C2TB, nnode_DL;
  unsigned int nodal_dim;
  double ***nodal_coords; /* [node][x..,X..][t]      (jitbridge.h:99)  */
  double ***nodal_data;   /* [space-local node][field index][t]  (jitbridge.h:100) */
  int **nodal_local_eqn;  /* [space-local node][field index], <0: pinned (jitbridge.h:103) */
  int **pos_local_eqn;    /* [node][dim] (jitbridge.h:104) */
  unsigned int ndof;
  struct JITElementInfo *bulk_eleminfo;
  struct JITElementInfo *opposite_eleminfo;
} JITElementInfo_t;

#define DX_SHAPE_FUNCTION_DECL(what) double *const *const what /* jitbridge.h:149 */

typedef struct JITHangInfoEntry
{
  double weight;
  int *local_eqn;
} JITHangInfoEntry_t;

typedef struct JITHangInfo
{
  int nummaster; /* 0: not hanging */
  JITHangInfoEntry_t *masters;
} JITHangInfo_t;

typedef struct JITShapeInfo
{
  unsigned int n_int_pt;
  double int_pt_weight, int_pt_weight_Lagrangian, int_pt_weight_unity;
  double elemsize_Eulerian, elemsize_Eulerian_cartesian; /* jitbridge.h:178 */
  double elemsize_Lagrangian, elemsize_Lagrangian_cartesian; /* jitbridge.h:179 */
  double **int_pt_weights_d_coords;      /* [dim][node] */
  double ****int_pt_weights_d2_coords;   /* [dim][dim][node][node] */
  double *shape_C2, **dx_shape_C2, **dX_shape_C2, **dS_shape_C2;
  double ****d_dx_shape_dcoord_C2;       /* [node][dir][coord node][coord dir] */
  double *shape_C1, **dx_shape_C1, **dX_shape_C1, **dS_shape_C1;
  double ****d_dx_shape_dcoord_C1;
  double *shape_Pos, **dx_shape_Pos, **dX_shape_Pos, **dS_shape_Pos;
  double ****d_dx_shape_dcoord_Pos;
  unsigned int jacobian_size, mass_matrix_size;
  double *t, *dt;
  unsigned int timestepper_ntstorage;
  double *timestepper_weights_dt_BDF1, *timestepper_weights_dt_BDF2;
  double *timestepper_weights_dt_Newmark2, *timestepper_weights_d2t_Newmark2;
  double *timestepper_weights_dt_BDF2_degr, *timestepper_weights_dt_Newmark2_degr;
  JITHangInfo_t *hanginfo_C1, *hanginfo_C2, *hanginfo_Pos;
  struct JITShapeInfo *bulk_shapeinfo, *opposite_shapeinfo;
  double *normal;            /* [dim] unit normal of an interface element at the point (jitbridge.h:251) */
  double ***d_normal_dcoord; /* [dir][coord node][coord dir] (jitbridge.h:252) */
} JITShapeInfo_t;

typedef void (*JITFuncSpec_ResidualAndJacobian_FiniteElement)(const JITElementInfo_t *, const JITShapeInfo_t *, double *, double *, double *, unsigned);
typedef void (*JITFuncSpec_HessianVectorProduct_FiniteElement)(const JITElementInfo_t *, const JITShapeInfo_t *, const double *, double *, double *, unsigned, unsigned);
typedef double (*JITFuncSpec_EvalIntegralExpr_FiniteElement)(const JITElementInfo_t *, const JITShapeInfo_t *, unsigned); /* jitbridge.h:291 */

typedef struct JITFuncSpec_RequiredShapes_FiniteElement
{
  bool psi_C1, psi_C2, dx_psi_C1, dx_psi_C2, dX_psi_C1, dX_psi_C2;
  bool psi_Pos, dx_psi_Pos, dX_psi_Pos;
  bool elemsize_Eulerian_Pos, elemsize_Lagrangian_Pos, elemsize_Eulerian_cartesian_Pos, elemsize_Lagrangian_cartesian_Pos; /* jitbridge.h:305-306 */
} JITFuncSpec_RequiredShapes_FiniteElement_t;

typedef void (*JITFuncSpec_GetZ2Fluxes_FiniteElement)(const JITElementInfo_t *, const JITShapeInfo_t *, double *); /* jitbridge.h:287 */

typedef struct JITFuncSpec_Table_FiniteElement
{
  unsigned int nodal_dim, lagr_dim;
  unsigned int numfields_C1, numfields_C2, numfields_Pos;
  char **fieldnames_C1, **fieldnames_C2, **fieldnames_Pos;
  unsigned num_res_jacs;
  int current_res_jac;
  char **res_jac_names;
  JITFuncSpec_RequiredShapes_FiniteElement_t *shapes_required_ResJac;
  JITFuncSpec_RequiredShapes_FiniteElement_t *shapes_required_Hessian;
  unsigned numglobal_params;
  unsigned *global_paramindices;
  double **global_parameters;
  JITFuncSpec_ResidualAndJacobian_FiniteElement **ParameterDerivative;
  char *dominant_space;
  int max_dt_order;
  bool fd_jacobian, fd_position_jacobian, with_adaptivity;
  int integration_order;
  bool moving_nodes;
  JITFuncSpec_ResidualAndJacobian_FiniteElement *ResidualAndJacobian;
  JITFuncSpec_ResidualAndJacobian_FiniteElement *ResidualAndJacobianSteady;
  JITFuncSpec_HessianVectorProduct_FiniteElement *HessianVectorProduct;
  bool hessian_generated;
  unsigned numintegral_expressions;                                /* jitbridge.h:417-418 */
  char **integral_expressions_names;
  JITFuncSpec_EvalIntegralExpr_FiniteElement EvalIntegralExpression; /* jitbridge.h:469-470 */
  JITFuncSpec_RequiredShapes_FiniteElement_t shapes_required_IntegralExprs;
  unsigned numlocal_expressions;                                   /* jitbridge.h:420-424 */
  char **local_expressions_names;
  unsigned numextremum_expressions;
  char **extremum_expressions_names;
  JITFuncSpec_EvalIntegralExpr_FiniteElement EvalLocalExpression;  /* jitbridge.h:471-473 */
  JITFuncSpec_EvalIntegralExpr_FiniteElement EvalExtremumExpression;
  unsigned num_Z2_flux_terms;                                      /* jitbridge.h:456-458 */
  JITFuncSpec_GetZ2Fluxes_FiniteElement GetZ2Fluxes;
  char *domain_name;
  void (*check_compiler_size)(unsigned long long, unsigned long long, char *);
  void (*fill_shape_buffer_for_point)(unsigned, JITFuncSpec_RequiredShapes_FiniteElement_t *, int);
  double (*JacobianForElementSize)(const JITElementInfo_t *, const double *); /* jitbridge.h:484 */
  void (*clean_up)(struct JITFuncSpec_Table_FiniteElement *functable);
} JITFuncSpec_Table_FiniteElement_t;

#define JIT_API

#ifdef JIT_ELEMENT_SHARED_LIB
/* jitbridge.h:503-519 */
static double step(double x) { return x < 0 ? 0.0 : (x > 0 ? 1.0 : 0.5); }
static double signum(double x) { return x < 0 ? -1.0 : (x > 0 ? 1.0 : x); }

#define SET_INTERNAL_FIELD_NAME(tab, index, name) { tab[index] = strdup(name); }
#define SET_INTERNAL_NAME(var, name) { var = strdup(name); }
#define pyoomph_tested_free(x) if (x) free(x);
#endif

#endif /* ORACLE_USE_REFERENCE_HEADERS */
#endif
