/* ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * Restatement of the accumulate macros the generated routines are written with
 * (/root/reference/src/jitbridge_hang.h:104-184 residual/Jacobian/mass-matrix with hanging-node
 * redirection, :41-53 BEGIN/END_JACOBIAN, :355 the truncated Pi).  Semantics: a continuous-space dof
 * either is an ordinary node (one target equation, weight 1) or hangs on `nummaster` masters
 * (target equation and weight per master); contributions are evaluated once and added, weighted,
 * to every non-pinned target.  Pinned targets (equation < 0) are skipped.
 */
#ifndef ORACLE_JIT_HANG_H
#define ORACLE_JIT_HANG_H

#ifdef ORACLE_USE_REFERENCE_HEADERS
#include "jitbridge_hang.h"
#else

#include <assert.h>

static inline int oracle_n_targets(const JITHangInfo_t *h) { return h->nummaster ? h->nummaster : 1; }
static inline int oracle_target_eqn(const JITHangInfo_t *h, int m, int nodalind, int direct_eqn)
{
  return h->nummaster ? h->masters[m].local_eqn[nodalind] : direct_eqn;
}
static inline double oracle_target_weight(const JITHangInfo_t *h, int m) { return h->nummaster ? h->masters[m].weight : 1.0; }

#define BEGIN_RESIDUAL_CONTINUOUS_SPACE(EQN, CONTRIB, HANGINFO, NODALIND, LINDEX)   \
  nummaster = oracle_n_targets(&(HANGINFO)[LINDEX]);                                \
  _res_contrib = CONTRIB;                                                           \
  for (int m = 0; m < (int)nummaster; m++)                                          \
  {                                                                                 \
    local_eqn = oracle_target_eqn(&(HANGINFO)[LINDEX], m, NODALIND, EQN);           \
    hang_weight = oracle_target_weight(&(HANGINFO)[LINDEX], m);                     \
    if (local_eqn >= 0)                                                             \
    {
#define ADD_TO_RESIDUAL_CONTINUOUS_SPACE()   \
      assert(local_eqn < (int)eleminfo->ndof); \
      residuals[local_eqn] += hang_weight * _res_contrib;
#define END_RESIDUAL_CONTINUOUS_SPACE() \
    }                                   \
  }

#define BEGIN_JACOBIAN() \
  if (flag)              \
  {
#define END_JACOBIAN() }

#define BEGIN_JACOBIAN_HANG(EQN, CONTRIB, HANGINFO, NODALIND, LINDEX)               \
  nummaster2 = oracle_n_targets(&(HANGINFO)[LINDEX]);                               \
  _J_contrib = CONTRIB;                                                             \
  for (int m2 = 0; m2 < (int)nummaster2; m2++)                                      \
  {                                                                                 \
    local_unknown = oracle_target_eqn(&(HANGINFO)[LINDEX], m2, NODALIND, EQN);      \
    hang_weight2 = oracle_target_weight(&(HANGINFO)[LINDEX], m2);                   \
    if (local_unknown >= 0)                                                         \
    {
#define ADD_TO_JACOBIAN_HANG_HANG() \
      jacobian[local_eqn * shapeinfo->jacobian_size + local_unknown] += hang_weight * hang_weight2 * _J_contrib;
#define ADD_TO_MASS_MATRIX_HANG_HANG(MPART)                                                                         \
      if (flag == 2)                                                                                                \
      {                                                                                                             \
        mass_matrix[local_eqn * shapeinfo->mass_matrix_size + local_unknown] += hang_weight * hang_weight2 * (MPART); \
      }
#define END_JACOBIAN_HANG() \
    }                       \
  }

#define Pi 3.14159265359 /* sic: jitbridge_hang.h:355 */

#endif /* ORACLE_USE_REFERENCE_HEADERS */
#endif
