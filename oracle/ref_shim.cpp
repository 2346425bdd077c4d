// ORACLE -- TEST INFRASTRUCTURE ONLY.  Thin extern "C" window onto the REFERENCE's own sources, compiled where they lie
// (see oracle/Makefile; outputs only into oracle/_ref/).  Used by tests/test_oracle_ref.py to pin the restated tables of
// oracle/driver.c and pyoomph_b200/cuda_emitter.py:
//   Gauss<1,3>, Gauss<2,3> (with its mistyped knots), Gauss<3,3>   oomph-lib/include/integral.{h,cc}
//   OneDimLagrange::shape<2|3>, dshape<2|3>                         oomph-lib/include/shape.h:604-650
// QElement<DIM,3>::dshape_local itself cannot be linked without most of oomph-lib (matrices/linear_solver/problem), so
// the tensor-product ordering of Qelements.cc:348-377/:621-660 stays a restatement (pinned by interpolation tests).
#include "integral.h"
#include "shape.h"
using namespace oomph;
extern "C" {
void ref_gauss(int dim, int ipt, double *knot, double *w)
{
  if (dim == 1) { Gauss<1, 3> g; knot[0] = g.knot(ipt, 0); *w = g.weight(ipt); }
  else if (dim == 2) { Gauss<2, 3> g; knot[0] = g.knot(ipt, 0); knot[1] = g.knot(ipt, 1); *w = g.weight(ipt); }
  else { Gauss<3, 3> g; for (int i = 0; i < 3; i++) knot[i] = g.knot(ipt, i); *w = g.weight(ipt); }
}
void ref_lagrange(int order, double s, double *psi, double *dpsi)
{
  if (order == 3) { OneDimLagrange::shape<3>(s, psi); OneDimLagrange::dshape<3>(s, dpsi); }
  else { OneDimLagrange::shape<2>(s, psi); OneDimLagrange::dshape<2>(s, dpsi); }
}
}
