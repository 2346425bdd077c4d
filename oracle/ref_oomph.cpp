// ORACLE -- TEST INFRASTRUCTURE ONLY.  extern "C" window onto the REFERENCE's own compiled code: the vendored oomph-lib `generic`
// library (/root/reference/src/thirdparty/oomph-lib/include/*.cc, all 50 translation units, no MPI) and pyoomph's standalone
// sources src/timestepper.cpp and src/hessian_tensor.cpp, compiled where they lie by oracle/Makefile into oracle/_ref/liboomph_ref.so.
// Nothing of the reference is copied; this file only CALLS it.  tests/test_oracle_ref.py uses it to pin the restatements:
//   ref_qshape            QElement<DIM,NNODE_1D>::dshape_local                       (Qelements.h / Qelements.cc:348-377, :621-660)
//   ref_problem_*         SolidNode / SolidQElement<DIM,3> / Mesh / Problem:
//       numbering         Problem::assign_eqn_numbers -> Mesh::assign_global_eqn_numbers        (mesh.cc:686-708, nodes.cc:896-927, :3652-3659)
//       local order       GeneralisedElement::assign_local_eqn_numbers                           (elements.cc:694-699)
//       geometry          FiniteElement::dshape_eulerian_at_knot, J_eulerian_at_knot, Integral   (elements.cc, integral.h)
//       assembly          Problem::get_jacobian(DoubleVector&, CRDoubleMatrix&) -> sparse_assemble_row_or_column_compressed
//                         (problem.cc:4038, :4469-4560; vectors_of_pairs :5332-5666, maps, lists, two_vectors, two_arrays)
//   ref_timestepper_*     pyoomph::MultiTimeStepper::set_weights                      (src/timestepper.cpp:31-80)
//   ref_rank3_*           pyoomph::SparseRank3Tensor                                  (src/hessian_tensor.cpp:31-96)
// pyoomph's elements are QElement<DIM,3> + (Refineable)SolidQElement on SolidNodes (src/elements.hpp:831, src/nodes.hpp:46-51): the
// same oomph classes are instantiated here, with the element matrices supplied by a callback (the oracle's generated routine).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "Qelements.h"
#include "Telements.h"
#include "integral.h"
#include "mesh.h"
#include "nodes.h"
#include "problem.h"
#include "shape.h"
#include "timesteppers.h"

#include "hessian_tensor.hpp"
#include "timestepper.hpp"

using namespace oomph;

// element matrices come from the caller: R[n], J[n*n] (row-major) in the element's OOMPH local equation order; geqn[i] = global
// equation of local equation i
typedef void (*ref_elem_cb)(void *ctx, int e, int n, const int *geqn, double *R, double *J, int flag);

namespace
{
  struct Callback
  {
    ref_elem_cb fn = nullptr;
    void *ctx = nullptr;
  };

  template <unsigned DIM>
  class RefElement : public virtual SolidQElement<DIM, 3>
  {
  public:
    int id = 0;
    Callback *cb = nullptr;
    RefElement() : SolidQElement<DIM, 3>()
    {
      this->set_lagrangian_dimension(DIM);
      this->set_nnodal_lagrangian_type(1);
    }
    void call(Vector<double> &r, DenseMatrix<double> *j, int flag)
    {
      const unsigned n = this->ndof();
      std::vector<int> g(n);
      for (unsigned i = 0; i < n; i++) g[i] = (int)this->eqn_number(i);
      std::vector<double> R(n, 0.0), J(flag ? (size_t)n * n : 1, 0.0);
      cb->fn(cb->ctx, id, (int)n, g.data(), R.data(), J.data(), flag);
      for (unsigned i = 0; i < n; i++) r[i] += R[i];
      if (j)
        for (unsigned i = 0; i < n; i++)
          for (unsigned k = 0; k < n; k++) (*j)(i, k) += J[(size_t)i * n + k];
    }
    void fill_in_contribution_to_residuals(Vector<double> &r) { call(r, nullptr, 0); }
    void fill_in_contribution_to_jacobian(Vector<double> &r, DenseMatrix<double> &j) { call(r, &j, 1); }
    void output(std::ostream &) {}
  };

  class RefProblem : public Problem
  {
  public:
    RefProblem() {}
    void set_method(unsigned m) { Sparse_assembly_method = m; }
  };

  struct Handle
  {
    int dim = 2, nvalue = 0;
    RefProblem *problem = nullptr;
    Mesh *mesh = nullptr;
    std::vector<SolidNode *> nodes;
    std::vector<GeneralisedElement *> elems;
    Callback cb;
    DoubleVector *res = nullptr;
    CRDoubleMatrix *jac = nullptr;
    unsigned long ndof = 0;
  };
}

extern "C"
{
  // psi[n], dpsi[n][dim] of QElement<dim,nnode_1d> at local coordinate s (oomph's own node order)
  int ref_qshape(int dim, int nnode_1d, const double *s, double *psi, double *dpsi)
  {
    Vector<double> sv(dim);
    for (int i = 0; i < dim; i++) sv[i] = s[i];
    FiniteElement *el = nullptr;
    if (dim == 2 && nnode_1d == 3) el = new QElement<2, 3>;
    else if (dim == 2 && nnode_1d == 2) el = new QElement<2, 2>;
    else if (dim == 3 && nnode_1d == 3) el = new QElement<3, 3>;
    else if (dim == 3 && nnode_1d == 2) el = new QElement<3, 2>;
    else if (dim == 1 && nnode_1d == 3) el = new QElement<1, 3>;
    else if (dim == 1 && nnode_1d == 2) el = new QElement<1, 2>;
    else return -1;
    const unsigned n = el->nnode();
    Shape p(n);
    DShape dp(n, dim);
    el->dshape_local(sv, p, dp);
    for (unsigned l = 0; l < n; l++)
    {
      psi[l] = p[l];
      for (int i = 0; i < dim; i++) dpsi[l * dim + i] = dp(l, i);
    }
    delete el;
    return (int)n;
  }

  // TElement<2,nnode_1d>::dshape_local (Telements.h:519-545, :627-664) and local_coordinate_of_node; node_s may be NULL
  int ref_tshape(int nnode_1d, const double *s, double *psi, double *dpsi, double *node_s)
  {
    Vector<double> sv(2);
    sv[0] = s[0];
    sv[1] = s[1];
    FiniteElement *el = nnode_1d == 3 ? (FiniteElement *)new TElement<2, 3> : (FiniteElement *)new TElement<2, 2>;
    const unsigned n = el->nnode();
    Shape p(n);
    DShape dp(n, 2);
    el->dshape_local(sv, p, dp);
    for (unsigned l = 0; l < n; l++)
    {
      psi[l] = p[l];
      dpsi[l * 2] = dp(l, 0);
      dpsi[l * 2 + 1] = dp(l, 1);
      if (node_s)
      {
        Vector<double> ns;
        el->local_coordinate_of_node(l, ns);
        node_s[l * 2] = ns[0];
        node_s[l * 2 + 1] = ns[1];
      }
    }
    delete el;
    return (int)n;
  }

  // default integration scheme of TElement<2,3> (TGauss<2,3>, integral.cc:355-369); returns the number of points
  // TElement<3,3> / TElement<3,2>: shape functions, local derivatives and local node coordinates of the compiled oomph-lib
  int ref_tshape3(int nnode_1d, const double *s, double *psi, double *dpsi, double *node_s)
  {
    Vector<double> sv(3);
    for (int d = 0; d < 3; d++) sv[d] = s[d];
    FiniteElement *el = nnode_1d == 3 ? (FiniteElement *)new TElement<3, 3> : (FiniteElement *)new TElement<3, 2>;
    const unsigned n = el->nnode();
    Shape p(n);
    DShape dp(n, 3);
    el->dshape_local(sv, p, dp);
    for (unsigned l = 0; l < n; l++)
    {
      psi[l] = p[l];
      for (int d = 0; d < 3; d++) dpsi[l * 3 + d] = dp(l, d);
      if (node_s)
      {
        Vector<double> ns;
        el->local_coordinate_of_node(l, ns);
        for (int d = 0; d < 3; d++) node_s[l * 3 + d] = ns[d];
      }
    }
    delete el;
    return (int)n;
  }

  // default integration scheme of TElement<3,3> (TGauss<3,3>)
  int ref_tgauss3(int ipt, double *knot, double *w)
  {
    TElement<3, 3> el;
    const int n = (int)el.integral_pt()->nweight();
    if (ipt >= 0 && ipt < n)
    {
      for (int d = 0; d < 3; d++) knot[d] = el.integral_pt()->knot(ipt, d);
      *w = el.integral_pt()->weight(ipt);
    }
    return n;
  }

  int ref_tgauss(int ipt, double *knot, double *w)
  {
    TElement<2, 3> el;
    const int n = (int)el.integral_pt()->nweight();
    if (ipt >= 0 && ipt < n)
    {
      knot[0] = el.integral_pt()->knot(ipt, 0);
      knot[1] = el.integral_pt()->knot(ipt, 1);
      *w = el.integral_pt()->weight(ipt);
    }
    return n;
  }

  // default integration scheme of QElement<dim,3>: knot[dim] and weight of point ipt; returns the number of points
  int ref_element_integral(int dim, int ipt, double *knot, double *w)
  {
    FiniteElement *el = dim == 2 ? (FiniteElement *)new QElement<2, 3> : dim == 3 ? (FiniteElement *)new QElement<3, 3> : (FiniteElement *)new QElement<1, 3>;
    const int n = (int)el->integral_pt()->nweight();
    if (ipt >= 0 && ipt < n)
    {
      for (int i = 0; i < dim; i++) knot[i] = el->integral_pt()->knot(ipt, i);
      *w = el->integral_pt()->weight(ipt);
    }
    delete el;
    return n;
  }

  // mesh of SolidNodes (nvalue values each, Lagrangian = initial Eulerian position) and SolidQElement<dim,3>; val_pinned[n_node][nvalue],
  // pos_pinned[n_node][dim] (NULL: every position pinned = fixed mesh) as 0/1 bytes
  void *ref_problem_create(int dim, long n_node, int nvalue, const double *pos, long n_elem, const int *elem_nodes, const unsigned char *val_pinned,
                           const unsigned char *pos_pinned)
  {
    Handle *h = new Handle;
    h->dim = dim;
    h->nvalue = nvalue;
    h->problem = new RefProblem;
    h->mesh = new Mesh;
    h->nodes.resize(n_node);
    for (long n = 0; n < n_node; n++)
    {
      SolidNode *nd = new SolidNode(dim, 1, dim, 1, nvalue);
      for (int i = 0; i < dim; i++)
      {
        nd->x(i) = pos[n * dim + i];
        nd->xi(i) = pos[n * dim + i];
        if (!pos_pinned || pos_pinned[n * dim + i]) nd->pin_position(i);
      }
      for (int v = 0; v < nvalue; v++)
        if (val_pinned[n * nvalue + v]) nd->pin(v);
      h->nodes[n] = nd;
      h->mesh->add_node_pt(nd);
    }
    const int nn = dim == 2 ? 9 : 27;
    for (long e = 0; e < n_elem; e++)
    {
      GeneralisedElement *ge = nullptr;
      if (dim == 2)
      {
        RefElement<2> *el = new RefElement<2>;
        el->id = (int)e;
        el->cb = &h->cb;
        for (int l = 0; l < nn; l++) el->node_pt(l) = h->nodes[elem_nodes[e * nn + l]];
        ge = el;
      }
      else
      {
        RefElement<3> *el = new RefElement<3>;
        el->id = (int)e;
        el->cb = &h->cb;
        for (int l = 0; l < nn; l++) el->node_pt(l) = h->nodes[elem_nodes[e * nn + l]];
        ge = el;
      }
      h->elems.push_back(ge);
      h->mesh->add_element_pt(ge);
    }
    h->problem->mesh_pt() = h->mesh;
    h->ndof = h->problem->assign_eqn_numbers();
    return h;
  }

  long ref_problem_ndof(void *hp) { return (long)((Handle *)hp)->ndof; }

  // global equation numbers as oomph assigned them: node_eqn[n_node][nvalue], pos_eqn[n_node][dim]; pinned = -1
  void ref_problem_numbering(void *hp, int *node_eqn, int *pos_eqn)
  {
    Handle *h = (Handle *)hp;
    for (size_t n = 0; n < h->nodes.size(); n++)
    {
      for (int v = 0; v < h->nvalue; v++)
      {
        const long q = h->nodes[n]->eqn_number(v);
        node_eqn[n * h->nvalue + v] = q >= 0 ? (int)q : -1;
      }
      if (pos_eqn)
        for (int i = 0; i < h->dim; i++)
        {
          const long q = h->nodes[n]->variable_position_pt()->eqn_number(i);
          pos_eqn[n * h->dim + i] = q >= 0 ? (int)q : -1;
        }
    }
  }

  // local equation order of element e: out[i] = global equation of local equation i; returns ndof of the element
  int ref_element_local_eqns(void *hp, long e, int *out)
  {
    Handle *h = (Handle *)hp;
    GeneralisedElement *el = h->elems[e];
    const unsigned n = el->ndof();
    for (unsigned i = 0; i < n; i++) out[i] = (int)el->eqn_number(i);
    return (int)n;
  }

  // Eulerian geometry of element e at integration point ipt by oomph itself: J (returned), psi[nnode], dpsidx[nnode][dim], weight
  double ref_element_geometry(void *hp, long e, int ipt, double *psi, double *dpsidx, double *w)
  {
    Handle *h = (Handle *)hp;
    FiniteElement *el = dynamic_cast<FiniteElement *>(h->elems[e]);
    const unsigned n = el->nnode();
    Shape p(n);
    DShape dp(n, h->dim);
    const double J = el->dshape_eulerian_at_knot(ipt, p, dp);
    for (unsigned l = 0; l < n; l++)
    {
      psi[l] = p[l];
      for (int i = 0; i < h->dim; i++) dpsidx[l * h->dim + i] = dp(l, i);
    }
    *w = el->integral_pt()->weight(ipt);
    return J;
  }

  // one assembly by oomph's Problem::get_jacobian with the chosen Sparse_assembly_method (0 vectors_of_pairs, 1 two_vectors, 2 maps,
  // 3 lists, 4 two_arrays; problem.h:674-681); the result stays in the handle.  Returns nnz.
  long ref_problem_assemble(void *hp, ref_elem_cb fn, void *ctx, int method)
  {
    Handle *h = (Handle *)hp;
    h->cb.fn = fn;
    h->cb.ctx = ctx;
    h->problem->set_method((unsigned)method);
    delete h->res;
    delete h->jac;
    h->res = new DoubleVector;
    h->jac = new CRDoubleMatrix;
    h->problem->get_jacobian(*h->res, *h->jac);
    return (long)h->jac->nnz();
  }

  void ref_problem_result(void *hp, int *row_start, int *column_index, double *values, double *residuals)
  {
    Handle *h = (Handle *)hp;
    const unsigned long n = h->ndof, nnz = h->jac->nnz();
    memcpy(row_start, h->jac->row_start(), (n + 1) * sizeof(int));
    memcpy(column_index, h->jac->column_index(), nnz * sizeof(int));
    memcpy(values, h->jac->value(), nnz * sizeof(double));
    for (unsigned long i = 0; i < n; i++) residuals[i] = (*h->res)[i];
  }

  void ref_problem_free(void *hp)
  {
    Handle *h = (Handle *)hp;
    delete h->res;
    delete h->jac;
    delete h->problem; // deletes the mesh (nodes and elements) it owns
    delete h;
  }

  // pyoomph::MultiTimeStepper::set_weights for (dt, dtprev): out arrays of 7 (weights beyond the stepper's storage are 0);
  // returns ntstorage
  int ref_timestepper_weights(double dt, double dtprev, double *bdf1, double *bdf2, double *newmark2_dt, double *newmark2_d2t)
  {
    Time time(2);
    time.dt(0) = dt;
    time.dt(1) = dtprev;
    pyoomph::MultiTimeStepper ts(false);
    ts.time_pt() = &time;
    ts.set_weights();
    const int nt = (int)ts.ntstorage();
    for (int i = 0; i < 7; i++)
    {
      bdf1[i] = i < nt ? ts.weightBDF1(1, i) : 0.0;
      bdf2[i] = i < nt ? ts.weightBDF2(1, i) : 0.0;
      newmark2_dt[i] = i < nt ? ts.weightNewmark2(1, i) : 0.0;
      newmark2_d2t[i] = i < nt ? ts.weightNewmark2(2, i) : 0.0;
    }
    return nt;
  }

  // pyoomph::SparseRank3Tensor: accumulate n_ent entries (i,j,k,v), finalize_for_vector_product, right_vector_mult(vec);
  // out: col_index / row_start of the product matrix (caller sizes: col <= n_ent, row n+1) and its values; returns nnz
  long ref_rank3_product(int n, int symmetric, long n_ent, const int *ii, const int *jj, const int *kk, const double *vv, const double *vec,
                         int *col_index, int *row_start, double *values)
  {
    pyoomph::SparseRank3Tensor T((unsigned)n, symmetric != 0);
    for (long q = 0; q < n_ent; q++) T.accumulate(ii[q], jj[q], kk[q], vv[q]);
    auto pat = T.finalize_for_vector_product();
    const std::vector<int> &ci = std::get<0>(pat), &rs = std::get<1>(pat);
    std::vector<double> x(vec, vec + n);
    std::vector<double> val = T.right_vector_mult(x);
    for (size_t q = 0; q < ci.size(); q++) col_index[q] = ci[q];
    for (size_t q = 0; q < rs.size(); q++) row_start[q] = rs[q];
    for (size_t q = 0; q < val.size(); q++) values[q] = val[q];
    return (long)val.size();
  }
}
