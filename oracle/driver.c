/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed from the product path
 * (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it).
 *
 * CPU restatement of the host side of pyoomph's element assembly, i.e. everything around the generated
 * routine, following the reference file by file:
 *   - element-local data tables            /root/reference/src/elements.cpp:2713-2990 (fill_element_info)
 *   - per-assembly shape buffer prep       src/elements.cpp:4577-4646 (prepare_shape_buffer_for_integration),
 *                                          :4510-4562 (Pos space aliases the dominant space)
 *   - per-Gauss-point geometry and shapes  src/elements.cpp:4564-4575, :3593-4300 (fill_shape_info_at_s)
 *   - moving-mesh derivative tensors       src/elements.cpp:3051-3155 (fill_shape_info_at_s_dNodalPos_helper)
 *   - shape functions / node order         oomph-lib/include/shape.h:604-650, Qelements.cc:348-377, :621-660,
 *                                          src/elements.cpp:9274-9306, :11163-11210 (C1 on vertex nodes)
 *   - Gauss rules (literal tables, incl. the mistyped knots of Gauss<2,3>)  oomph-lib/include/integral.cc:84-102, :169-207
 *   - element wrapper                      src/elements.cpp:5054-5127 (fill_in_generic_residual_contribution_jit)
 *   - global loop + vectors_of_pairs CSR   oomph-lib/include/problem.cc:5459-5659, Numerical_zero 0.0 (:110)
 *   - local equation order                 oomph-lib/include/elements.cc:694-699 (nodal values, then solid positions)
 *   - time-stepper weight selection        src/elements.cpp:4583-4632 (_degr rule)
 * It is compiled together with one generated-format plugin (oracle/emit_c.py) into one shared object.
 * Parity status: the restated tables/shape functions are pinned against the reference's own oomph-lib
 * sources compiled into oracle/_ref (tests/test_oracle_ref.py); the geometry/assembly logic has no golden
 * vectors in the reference and is pinned by patch / finite-difference / invariance tests only.
 */
#define JIT_ELEMENT_SHARED_LIB
#include "oracle_jit.h"
#include <math.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

extern void JIT_ELEMENT_init(JITFuncSpec_Table_FiniteElement_t *functable);

/* ------------------------------------------------------------------ quadrature tables (verbatim literals) */
#define K3 0.774596669241483
#define K3T 0.774596662941483 /* sic, integral.cc:87-93 */
static const double Gauss23_knot[9][2] = {{-K3, -K3}, {-K3, 0.0}, {-K3, K3T}, {0.0, -K3}, {0.0, 0.0}, {0.0, K3T}, {K3T, -K3}, {K3T, 0.0}, {K3T, K3T}};
static const double Gauss23_weight[9] = {(25.0 / 81.0), (40.0 / 81.0), (25.0 / 81.0), (40.0 / 81.0), (64.0 / 81.0), (40.0 / 81.0), (25.0 / 81.0), (40.0 / 81.0), (25.0 / 81.0)};
#define K33 0.77459666924148
static double Gauss33_knot[27][3];
static const double Gauss33_weight[27] = {
    0.17146776406035, 0.27434842249657, 0.17146776406035, 0.27434842249657, 0.43895747599451, 0.27434842249657, 0.17146776406035,
    0.27434842249657, 0.17146776406035, 0.27434842249657, 0.43895747599451, 0.27434842249657, 0.43895747599451, 0.70233196159122,
    0.43895747599451, 0.27434842249657, 0.43895747599451, 0.27434842249657, 0.17146776406035, 0.27434842249657, 0.17146776406035,
    0.27434842249657, 0.43895747599451, 0.27434842249657, 0.17146776406035, 0.27434842249657, 0.17146776406035};
static void init_tables(void)
{
  static int done = 0;
  if (done) return;
  const double k[3] = {-K33, 0, K33};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int l = 0; l < 3; l++)
      {
        Gauss33_knot[9 * i + 3 * j + l][0] = k[i];
        Gauss33_knot[9 * i + 3 * j + l][1] = k[j];
        Gauss33_knot[9 * i + 3 * j + l][2] = k[l];
      }
  done = 1;
}

void oracle_gauss(int dim, int ipt, double *knot, double *weight)
{
  init_tables();
  if (dim == 2) { knot[0] = Gauss23_knot[ipt][0]; knot[1] = Gauss23_knot[ipt][1]; *weight = Gauss23_weight[ipt]; }
  else { for (int i = 0; i < 3; i++) knot[i] = Gauss33_knot[ipt][i]; *weight = Gauss33_weight[ipt]; }
}

/* Gauss<1,3> (oomph-lib integral.cc:50-53) */
void oracle_gauss_1d(int ipt, double *knot, double *weight)
{
  static const double k[3] = {-0.774596669241483, 0.0, 0.774596669241483}, w[3] = {(5.0 / 9.0), (8.0 / 9.0), (5.0 / 9.0)};
  knot[0] = k[ipt];
  *weight = w[ipt];
}

/* ------------------------------------------------------------------ shape functions */
static void lag3(double s, double *p) { p[0] = 0.5 * s * (s - 1.0); p[1] = 1.0 - s * s; p[2] = 0.5 * s * (s + 1.0); }
static void dlag3(double s, double *p) { p[0] = s - 0.5; p[1] = -2.0 * s; p[2] = s + 0.5; }
static void lag2(double s, double *p) { p[0] = 0.5 * (1.0 - s); p[1] = 0.5 * (1.0 + s); }
static void dlag2(double s, double *p) { (void)s; p[0] = -0.5; p[1] = 0.5; }

/* psi[n], dpsi[n][dim]; order = 3 (C2) or 2 (C1) */
void oracle_dshape_local(int dim, int order, const double *s, double *psi, double *dpsi)
{
  double P[3][3], D[3][3];
  for (int a = 0; a < dim; a++)
  {
    if (order == 3) { lag3(s[a], P[a]); dlag3(s[a], D[a]); }
    else { lag2(s[a], P[a]); dlag2(s[a], D[a]); }
  }
  int index = 0;
  if (dim == 2)
  {
    for (int i = 0; i < order; i++)
      for (int j = 0; j < order; j++)
      {
        dpsi[index * 2 + 0] = P[1][i] * D[0][j];
        dpsi[index * 2 + 1] = D[1][i] * P[0][j];
        psi[index] = P[1][i] * P[0][j];
        ++index;
      }
  }
  else
  {
    for (int i = 0; i < order; i++)
      for (int j = 0; j < order; j++)
        for (int k = 0; k < order; k++)
        {
          dpsi[index * 3 + 0] = P[2][i] * P[1][j] * D[0][k];
          dpsi[index * 3 + 1] = P[2][i] * D[1][j] * P[0][k];
          dpsi[index * 3 + 2] = D[2][i] * P[1][j] * P[0][k];
          psi[index] = P[2][i] * P[1][j] * P[0][k];
          ++index;
        }
  }
}

/* ------------------------------------------------------------------ triangles: TElement<2,3> / TElement<2,2> and TGauss<2,3>
 * (oomph-lib Telements.h:519-545, :627-664; integral.cc:355-369), literal tables and the same operation order */
static const double TGauss23_knot[7][2] = {{0.1012865073235, 0.1012865073235}, {0.7974269853531, 0.1012865073235}, {0.1012865073235, 0.7974269853531},
                                           {0.4701420641051, 0.0597158717898}, {0.4701420641051, 0.4701420641051}, {0.0597158717898, 0.4701420641051},
                                           {0.3333333333333, 0.3333333333333}};
static const double TGauss23_weight[7] = {0.5 * 0.1259391805448, 0.5 * 0.1259391805448, 0.5 * 0.1259391805448, 0.5 * 0.1323941527885,
                                          0.5 * 0.1323941527885, 0.5 * 0.1323941527885, 0.5 * 0.225};
void oracle_gauss_tri(int ipt, double *knot, double *weight)
{
  knot[0] = TGauss23_knot[ipt][0];
  knot[1] = TGauss23_knot[ipt][1];
  *weight = TGauss23_weight[ipt];
}
void oracle_dshape_local_tri(int order, const double *s, double *psi, double *dpsi)
{
  if (order == 3)
  {
    const double s_2 = 1.0 - s[0] - s[1];
    psi[0] = 2.0 * s[0] * (s[0] - 0.5);
    psi[1] = 2.0 * s[1] * (s[1] - 0.5);
    psi[2] = 2.0 * s_2 * (s_2 - 0.5);
    psi[3] = 4.0 * s[0] * s[1];
    psi[4] = 4.0 * s[1] * s_2;
    psi[5] = 4.0 * s_2 * s[0];
    dpsi[0] = 4.0 * s[0] - 1.0;
    dpsi[1] = 0.0;
    dpsi[2] = 0.0;
    dpsi[3] = 4.0 * s[1] - 1.0;
    dpsi[4] = 2.0 * (2.0 * s[0] - 1.5 + 2.0 * s[1]);
    dpsi[5] = 2.0 * (2.0 * s[0] - 1.5 + 2.0 * s[1]);
    dpsi[6] = 4.0 * s[1];
    dpsi[7] = 4.0 * s[0];
    dpsi[8] = -4.0 * s[1];
    dpsi[9] = 4.0 * (1.0 - s[0] - 2.0 * s[1]);
    dpsi[10] = 4.0 * (1.0 - 2.0 * s[0] - s[1]);
    dpsi[11] = -4.0 * s[0];
  }
  else
  {
    psi[0] = s[0];
    psi[1] = s[1];
    psi[2] = 1.0 - s[0] - s[1];
    dpsi[0] = 1.0; dpsi[1] = 0.0;
    dpsi[2] = 0.0; dpsi[3] = 1.0;
    dpsi[4] = -1.0; dpsi[5] = -1.0;
  }
}

/* ------------------------------------------------------------------ tetrahedra: TElement<3,3> / TElement<3,2> and TGauss<3,3>
 * (oomph-lib Telements.h:1978-2012, :2133-2197; integral.cc:752-776), literal tables and the same operation order */
static const double TGauss33_knot[11][3] = {{0.25, 0.25, 0.25},
                                            {0.785714285714286, 0.071428571428571, 0.071428571428571},
                                            {0.071428571428571, 0.071428571428571, 0.071428571428571},
                                            {0.071428571428571, 0.785714285714286, 0.071428571428571},
                                            {0.071428571428571, 0.071428571428571, 0.785714285714286},
                                            {0.399403576166799, 0.399403576166799, 0.100596423833201},
                                            {0.399403576166799, 0.100596423833201, 0.399403576166799},
                                            {0.100596423833201, 0.399403576166799, 0.399403576166799},
                                            {0.399403576166799, 0.100596423833201, 0.100596423833201},
                                            {0.100596423833201, 0.399403576166799, 0.100596423833201},
                                            {0.100596423833201, 0.100596423833201, 0.399403576166799}};
static const double TGauss33_weight[11] = {-0.01315555555556, 0.00762222222222, 0.00762222222222, 0.00762222222222, 0.00762222222222, 0.02488888888889,
                                           0.02488888888889,  0.02488888888889, 0.02488888888889, 0.02488888888889, 0.02488888888889};
void oracle_gauss_tet(int ipt, double *knot, double *weight)
{
  for (int d = 0; d < 3; d++) knot[d] = TGauss33_knot[ipt][d];
  *weight = TGauss33_weight[ipt];
}
void oracle_dshape_local_tet(int order, const double *s, double *psi, double *dpsi)
{
  const double s3 = 1.0 - s[0] - s[1] - s[2];
  if (order == 3)
  {
    psi[0] = (2.0 * s[0] - 1.0) * s[0];
    psi[1] = (2.0 * s[1] - 1.0) * s[1];
    psi[2] = (2.0 * s[2] - 1.0) * s[2];
    psi[3] = (2.0 * s3 - 1.0) * s3;
    psi[4] = 4.0 * s[0] * s[1];
    psi[5] = 4.0 * s[0] * s[2];
    psi[6] = 4.0 * s[0] * s3;
    psi[7] = 4.0 * s[1] * s[2];
    psi[8] = 4.0 * s[2] * s3;
    psi[9] = 4.0 * s[1] * s3;
    const double d[10][3] = {{4.0 * s[0] - 1.0, 0.0, 0.0}, {0.0, 4.0 * s[1] - 1.0, 0.0}, {0.0, 0.0, 4.0 * s[2] - 1.0},
                             {-4.0 * s3 + 1.0, -4.0 * s3 + 1.0, -4.0 * s3 + 1.0}, {4.0 * s[1], 4.0 * s[0], 0.0}, {4.0 * s[2], 0.0, 4.0 * s[0]},
                             {4.0 * (s3 - s[0]), -4.0 * s[0], -4.0 * s[0]}, {0.0, 4.0 * s[2], 4.0 * s[1]}, {-4.0 * s[2], -4.0 * s[2], 4.0 * (s3 - s[2])},
                             {-4.0 * s[1], 4.0 * (s3 - s[1]), -4.0 * s[1]}};
    for (int l = 0; l < 10; l++)
      for (int b = 0; b < 3; b++) dpsi[l * 3 + b] = d[l][b];
  }
  else
  {
    psi[0] = s[0];
    psi[1] = s[1];
    psi[2] = s[2];
    psi[3] = 1.0 - s[0] - s[1] - s[2];
    const double d[4][3] = {{1.0, 0.0, 0.0}, {0.0, 1.0, 0.0}, {0.0, 0.0, 1.0}, {-1.0, -1.0, -1.0}};
    for (int l = 0; l < 4; l++)
      for (int b = 0; b < 3; b++) dpsi[l * 3 + b] = d[l][b];
  }
}

/* ------------------------------------------------------------------ problem data */
#define MAXN 27
#define MAXD 3
#define NTW 7

typedef struct
{
  int dim, nnode, nnode_C1, n_int;
  int c1_nodes[8];
  int tri;  /* TElement<2,3> instead of QElement<2,3> */
  int edim; /* dimension of the element itself: dim for bulk elements, 1 for a line element in 2D (interface elements) */
  int face; /* a face (s1 = -1) of a Q9 bulk element seen through the bulk element: bulk shape functions and gradients on the face,
               Gauss<1,3> along it, measure |dx/ds0|, outer normal (the reference's bulk_eleminfo / opposite_eleminfo, jitbridge.h:88-120) */
} EType;

typedef struct { int col; double val; } Pair;
typedef struct { Pair *p; int n, cap; } Row;

typedef struct
{
  EType et;
  int n_elem, n_node, nval, T, n_dof, n_pos_hist;
  const int *elem_nodes; /* [n_elem][nnode] */
  double *pos;           /* [node][dim][T]   oomph Data layout: value-major, history contiguous */
  double *lagr;          /* [node][dim] */
  double *val;           /* [node][nval][T] */
  const int *node_eqn;   /* [node][nval] */
  const int *pos_eqn;    /* [node][dim] or NULL */
  JITFuncSpec_Table_FiniteElement_t *ft;
  double *params;
  double t[NTW], dt[NTW];
  double wBDF1[NTW], wBDF2[NTW], wNM2[NTW], wNM2_d2t[NTW];
  int steady, unsteady_steps_done, ntstorage;
  /* CSR result of the last assembly, per matrix (0: Jacobian, 1: mass matrix) */
  int *row_start[2], *col_index[2];
  double *value[2];
  int64_t nnz[2];
  /* hanging nodes per space (0: C2 / positions, 1: C1), CSR over the nodes: masters and weights (oomph-lib HangInfo) */
  int *hang_start[2], *hang_master[2];
  double *hang_weight[2];
} Oracle;

/* per-thread assembly state = the reference's global shape buffer + _currently_assembled_element
 * (src/elements.cpp:34,251) */
typedef struct
{
  Oracle *o;
  JITElementInfo_t ei;
  JITShapeInfo_t si;
  int elem;
  int node_of[MAXN];
  int eqn_of_local[MAXN * (MAXD + 16)];
  JITHangInfo_t nohang[MAXN];
#define MAXM 9
  JITHangInfo_t hang_C2[MAXN], hang_C1[8], hang_Pos[MAXN];
  JITHangInfoEntry_t hang_entries_C2[MAXN][MAXM], hang_entries_C1[8][MAXM], hang_entries_Pos[MAXN][MAXM];
  int hang_eqn_C2[MAXN][MAXM][32], hang_eqn_C1[8][MAXM][32], hang_eqn_Pos[MAXN][MAXM][MAXD];
  /* backing storage */
  double **coord_ptr[MAXN], **data_ptr[MAXN];
  double *coord_slots[MAXN][2 * MAXD], *data_slots[MAXN][32];
  int *eqn_ptr[MAXN], *poseqn_ptr[MAXN];
  int eqn_slots[MAXN][32], poseqn_slots[MAXN][MAXD];
} ThreadState;

static __thread ThreadState *TS = NULL;

static void *xcalloc(size_t n, size_t s)
{
  void *p = calloc(n ? n : 1, s);
  if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
  return p;
}

static double **alloc2(int a, int b)
{
  double **p = (double **)xcalloc(a, sizeof(double *));
  for (int i = 0; i < a; i++) p[i] = (double *)xcalloc(b, sizeof(double));
  return p;
}
static double ****alloc4(int a, int b, int c, int d)
{
  double ****p = (double ****)xcalloc(a, sizeof(double ***));
  for (int i = 0; i < a; i++)
  {
    p[i] = (double ***)xcalloc(b, sizeof(double **));
    for (int j = 0; j < b; j++) p[i][j] = alloc2(c, d);
  }
  return p;
}

static void check_size(unsigned long long a, unsigned long long b, char *what)
{
  if (a != b) { fprintf(stderr, "oracle: compiler size mismatch for %s\n", what); abort(); }
}

/* ------------------------------------------------------------------ geometry at one Gauss point
 * restates BulkElementBase::fill_shape_info_at_s (src/elements.cpp:3593-4268): tangents, metric, inverse metric, gab_gai and the
 * determinant for el_dim 1 (:3651-3672), 2 (:3677-3703), 3 (:3804-3836) in nodal dimension >= el_dim; the unit normal of a line
 * element in 2D (get_normal_at_s, :1730-1752) and its coordinate derivatives (get_dnormal_dcoords_at_s, :1461-1490) */
static void element_dshape_local(const Oracle *o, int order, const double *s, double *psi, double *dpsi)
{
  if (o->et.tri && o->et.dim == 3) oracle_dshape_local_tet(order, s, psi, dpsi);
  else if (o->et.tri) oracle_dshape_local_tri(order, s, psi, dpsi);
  else if (o->et.edim == 1)
  {
    if (order == 3) { lag3(s[0], psi); dlag3(s[0], dpsi); }
    else { lag2(s[0], psi); dlag2(s[0], dpsi); }
  }
  else oracle_dshape_local(o->et.dim, order, s, psi, dpsi);
}

static void fill_shape_info_at_s(ThreadState *ts, const double *s, double weight, unsigned flag,
                                 const JITFuncSpec_RequiredShapes_FiniteElement_t *req)
{
  (void)req; /* everything is filled; the reference skips unrequired spaces, values are identical */
  Oracle *o = ts->o;
  const int dim = o->et.dim, ed = o->et.edim, nn = o->et.nnode;
  JITShapeInfo_t *si = &ts->si;
  double psi[MAXN], dpsids[MAXN * MAXD];
  element_dshape_local(o, 3, s, psi, dpsids);
  double t[MAXD][MAXD], TL[MAXD][MAXD]; /* tangents t(a,i) Eulerian and Lagrangian */
  memset(t, 0, sizeof(t));
  memset(TL, 0, sizeof(TL));
  for (int l = 0; l < nn; l++)
  {
    for (int i = 0; i < dim; i++)
      for (int j = 0; j < ed; j++) t[j][i] += ts->ei.nodal_coords[l][i][0] * dpsids[l * ed + j];
    for (int i = 0; i < dim; i++)
      for (int j = 0; j < ed; j++) TL[j][i] += ts->ei.nodal_coords[l][dim + i][0] * dpsids[l * ed + j];
  }
  double gg[MAXD][MAXD], ggL[MAXD][MAXD], aup[MAXD][MAXD], detE = 0, detL = 0;
  const int require_dxdshape = (flag && o->ft->moving_nodes && !o->ft->fd_position_jacobian);
  for (int pass = 0; pass < 2; pass++)
  {
    double(*tt)[MAXD] = pass == 0 ? t : TL;
    double(*g)[MAXD] = pass == 0 ? gg : ggL;
    double amet[MAXD][MAXD], up[MAXD][MAXD], det_a;
    memset(up, 0, sizeof(up));
    for (int al = 0; al < ed; al++)
      for (int be = 0; be < ed; be++)
      {
        amet[al][be] = 0.0;
        for (int i = 0; i < dim; i++) amet[al][be] += tt[al][i] * tt[be][i];
      }
    if (ed == 1)
    {
      det_a = amet[0][0];
      up[0][0] = 1.0 / det_a;
      for (int i = 0; i < dim; i++) g[0][i] = tt[0][i] / det_a;
    }
    else if (ed == 2)
    {
      det_a = amet[0][0] * amet[1][1] - amet[0][1] * amet[1][0];
      up[0][0] = amet[1][1] / det_a;
      up[0][1] = -amet[0][1] / det_a;
      up[1][0] = -amet[1][0] / det_a;
      up[1][1] = amet[0][0] / det_a;
      for (int b = 0; b < 2; b++)
        for (int i = 0; i < dim; i++) g[b][i] = up[0][b] * tt[0][i] + up[1][b] * tt[1][i];
    }
    else
    {
      det_a = amet[0][0] * amet[1][1] * amet[2][2] + amet[0][1] * amet[1][2] * amet[2][0] + amet[0][2] * amet[1][0] * amet[2][1] - amet[0][0] * amet[1][2] * amet[2][1] - amet[0][1] * amet[1][0] * amet[2][2] - amet[0][2] * amet[1][1] * amet[2][0];
      up[0][0] = (amet[1][1] * amet[2][2] - amet[1][2] * amet[2][1]) / det_a;
      up[0][1] = -(amet[0][1] * amet[2][2] - amet[0][2] * amet[2][1]) / det_a;
      up[0][2] = (amet[0][1] * amet[1][2] - amet[0][2] * amet[1][1]) / det_a;
      up[1][0] = -(amet[1][0] * amet[2][2] - amet[1][2] * amet[2][0]) / det_a;
      up[1][1] = (amet[0][0] * amet[2][2] - amet[0][2] * amet[2][0]) / det_a;
      up[1][2] = -(amet[0][0] * amet[1][2] - amet[0][2] * amet[1][0]) / det_a;
      up[2][0] = (amet[1][0] * amet[2][1] - amet[1][1] * amet[2][0]) / det_a;
      up[2][1] = -(amet[0][0] * amet[2][1] - amet[0][1] * amet[2][0]) / det_a;
      up[2][2] = (amet[0][0] * amet[1][1] - amet[0][1] * amet[1][0]) / det_a;
      for (int b = 0; b < 3; b++)
        for (int i = 0; i < dim; i++) g[b][i] = up[0][b] * tt[0][i] + up[1][b] * tt[1][i] + up[2][b] * tt[2][i];
    }
    if (pass == 0)
    {
      detE = sqrt(o->et.face ? amet[0][0] : det_a);
      memcpy(aup, up, sizeof(up));
    }
    else
      detL = sqrt(o->et.face ? amet[0][0] : det_a);
  }

  /* moving-mesh helper (src/elements.cpp:3051-3155): int_pt_weights_d_coords and DXdshape_il_jb, el_dim x nodal_dim general */
  static __thread double DX[MAXD][MAXN][MAXD][MAXD];
  if (require_dxdshape)
  {
    double Tt[MAXN][MAXD][MAXD][MAXD], G[MAXN][MAXD][MAXD][MAXD];
    for (int l = 0; l < nn; l++)
      for (int i = 0; i < dim; i++)
      {
        double dshape_dx = 0.0;
        for (int a = 0; a < ed; a++)
          for (int b = 0; b < ed; b++) dshape_dx += aup[a][b] * dpsids[l * ed + b] * t[a][i];
        si->int_pt_weights_d_coords[i][l] = dshape_dx * detE * weight;
      }
    for (int l = 0; l < nn; l++)
      for (int c = 0; c < ed; c++)
        for (int d = 0; d < ed; d++)
          for (int j = 0; j < dim; j++) Tt[l][c][d][j] = dpsids[l * ed + c] * t[d][j] + dpsids[l * ed + d] * t[c][j];
    for (int l = 0; l < nn; l++)
      for (int a = 0; a < ed; a++)
        for (int b = 0; b < ed; b++)
          for (int j = 0; j < dim; j++)
          {
            double Gval = 0.0;
            for (int c = 0; c < ed; c++)
              for (int d = 0; d < ed; d++) Gval -= aup[a][c] * Tt[l][c][d][j] * aup[d][b];
            G[l][a][b][j] = Gval;
          }
    for (int i = 0; i < dim; i++)
      for (int l = 0; l < nn; l++)
        for (int j = 0; j < dim; j++)
          for (int b = 0; b < ed; b++)
          {
            double v = 0.0;
            for (int a = 0; a < ed; a++)
            {
              if (i == j) v += aup[a][b] * dpsids[l * ed + a];
              v += t[a][j] * G[l][a][b][i];
            }
            DX[i][l][j][b] = v;
          }
  }

  /* spaces: C2 (dominant, aliased by Pos) and C1 on the vertex nodes */
  for (int space = 0; space < 2; space++)
  {
    const int n = space == 0 ? nn : o->et.nnode_C1;
    double p1[8], d1[8 * MAXD];
    const double *P = psi, *D = dpsids;
    if (space == 1)
    {
      element_dshape_local(o, 2, s, p1, d1);
      P = p1;
      D = d1;
    }
    double *shape = space == 0 ? si->shape_C2 : si->shape_C1;
    double **dx = space == 0 ? si->dx_shape_C2 : si->dx_shape_C1;
    double **dX = space == 0 ? si->dX_shape_C2 : si->dX_shape_C1;
    double **dS = space == 0 ? si->dS_shape_C2 : si->dS_shape_C1;
    double ****dd = space == 0 ? si->d_dx_shape_dcoord_C2 : si->d_dx_shape_dcoord_C1;
    for (int l = 0; l < n; l++)
    {
      shape[l] = P[l];
      for (int i = 0; i < dim; i++)
      {
        dx[l][i] = 0.0;
        for (int b = 0; b < ed; b++) dx[l][i] += gg[b][i] * D[l * ed + b];
      }
      for (int i = 0; i < ed; i++) dS[l][i] = D[l * ed + i];
      for (int i = 0; i < dim; i++)
      {
        dX[l][i] = 0.0;
        for (int b = 0; b < ed; b++) dX[l][i] += ggL[b][i] * D[l * ed + b];
      }
      if (require_dxdshape)
        for (int i = 0; i < dim; i++)
          for (int l2 = 0; l2 < nn; l2++)
            for (int i2 = 0; i2 < dim; i2++)
            {
              dd[l][i][l2][i2] = 0.0;
              for (int b = 0; b < ed; b++) dd[l][i][l2][i2] += DX[i2][l2][i][b] * D[l * ed + b];
            }
    }
  }
  if (ed == 1 && dim == 2)
  {
    /* get_normal_at_s for a line element (src/elements.cpp:1730-1752): n = (-dxds_y, dxds_x) / |dxds| */
    double len = t[0][0] * t[0][0] + t[0][1] * t[0][1];
    if (len < 1e-20) len = 1;
    const double len_sqr = len;
    len = sqrt(len);
    si->normal[0] = -t[0][1] / len;
    si->normal[1] = t[0][0] / len;
    if (require_dxdshape)
    {
      /* get_dnormal_dcoords_at_s (src/elements.cpp:1461-1490) */
      (void)len_sqr;
      const double denom = 1 / (len * len * len);
      for (int i = 0; i < dim; i++)
        for (int l = 0; l < nn; l++)
          for (int k = 0; k < dim; k++) si->d_normal_dcoord[i][l][k] = dpsids[l] * denom * (k == 1 ? -1 : 1) * t[0][i] * t[0][1 - k];
    }
  }
  if (o->et.face)
  {
    /* outer unit normal of the face s1 = -1 of a counter-clockwise element: (t_y, -t_x) / |t|, t = dx/ds0 */
    const double len = sqrt(t[0][0] * t[0][0] + t[0][1] * t[0][1]);
    si->normal[0] = t[0][1] / len;
    si->normal[1] = -t[0][0] / len;
  }
  si->int_pt_weight_unity = weight;
  si->int_pt_weight = weight * detE;
  si->int_pt_weight_Lagrangian = weight * detL;
}

/* callback installed into the function table (src/elements.cpp:60-63 -> :4564) */
static void cb_fill_shape_buffer_for_point(unsigned ipt, JITFuncSpec_RequiredShapes_FiniteElement_t *req, int flag)
{
  ThreadState *ts = TS;
  double s[MAXD], w;
  if (ts->o->et.face)
  {
    oracle_gauss_1d((int)ipt, s, &w);
    s[1] = -1.0;
  }
  else if (ts->o->et.tri && ts->o->et.dim == 3) oracle_gauss_tet((int)ipt, s, &w);
  else if (ts->o->et.tri) oracle_gauss_tri((int)ipt, s, &w);
  else if (ts->o->et.edim == 1) oracle_gauss_1d((int)ipt, s, &w);
  else oracle_gauss(ts->o->et.dim, (int)ipt, s, &w);
  fill_shape_info_at_s(ts, s, w, (unsigned)flag, req);
}

/* ------------------------------------------------------------------ per-thread state */
static ThreadState *ts_create(Oracle *o)
{
  ThreadState *ts = (ThreadState *)xcalloc(1, sizeof(ThreadState));
  const int dim = o->et.dim, nn = o->et.nnode;
  ts->o = o;
  JITShapeInfo_t *si = &ts->si;
  si->int_pt_weights_d_coords = alloc2(dim, nn);
  si->shape_C2 = (double *)xcalloc(nn, sizeof(double));
  si->dx_shape_C2 = alloc2(nn, dim);
  si->dX_shape_C2 = alloc2(nn, dim);
  si->dS_shape_C2 = alloc2(nn, dim);
  si->d_dx_shape_dcoord_C2 = alloc4(nn, dim, nn, dim);
  si->shape_C1 = (double *)xcalloc(8, sizeof(double));
  si->dx_shape_C1 = alloc2(8, dim);
  si->dX_shape_C1 = alloc2(8, dim);
  si->dS_shape_C1 = alloc2(8, dim);
  si->d_dx_shape_dcoord_C1 = alloc4(8, dim, nn, dim);
  si->normal = (double *)xcalloc(MAXD, sizeof(double));
  si->d_normal_dcoord = (double ***)xcalloc(MAXD, sizeof(double **));
  for (int i = 0; i < MAXD; i++) si->d_normal_dcoord[i] = alloc2(nn, MAXD);
  /* set_remaining_shapes_appropriately: Pos aliases C2 (src/elements.cpp:4534-4542) */
  si->shape_Pos = si->shape_C2;
  si->dx_shape_Pos = si->dx_shape_C2;
  si->dX_shape_Pos = si->dX_shape_C2;
  si->dS_shape_Pos = si->dS_shape_C2;
  si->d_dx_shape_dcoord_Pos = si->d_dx_shape_dcoord_C2;
  si->t = (double *)xcalloc(NTW, sizeof(double));
  si->dt = (double *)xcalloc(NTW, sizeof(double));
  si->timestepper_weights_dt_BDF1 = (double *)xcalloc(NTW, sizeof(double));
  si->timestepper_weights_dt_BDF2 = (double *)xcalloc(NTW, sizeof(double));
  si->timestepper_weights_dt_Newmark2 = (double *)xcalloc(NTW, sizeof(double));
  si->timestepper_weights_d2t_Newmark2 = (double *)xcalloc(NTW, sizeof(double));
  si->hanginfo_C1 = ts->nohang;
  si->hanginfo_C2 = ts->nohang;
  si->hanginfo_Pos = ts->nohang;
  for (int l = 0; l < nn; l++)
  {
    ts->coord_ptr[l] = ts->coord_slots[l];
    ts->data_ptr[l] = ts->data_slots[l];
    ts->eqn_ptr[l] = ts->eqn_slots[l];
    ts->poseqn_ptr[l] = ts->poseqn_slots[l];
  }
  ts->ei.nodal_coords = ts->coord_ptr;
  ts->ei.nodal_data = ts->data_ptr;
  ts->ei.nodal_local_eqn = ts->eqn_ptr;
  ts->ei.pos_local_eqn = ts->poseqn_ptr;
  ts->ei.nnode = nn;
  ts->ei.nnode_C2 = nn;
  ts->ei.nnode_C1 = o->et.nnode_C1;
  ts->ei.nodal_dim = dim;
  return ts;
}

/* fill_element_info (src/elements.cpp:2713): pointer tables + local equation numbers.
 * Local order: nodal values node by node, then positions node by node (oomph elements.cc:694-699). */
static void bind_element(ThreadState *ts, int e)
{
  Oracle *o = ts->o;
  const int dim = o->et.dim, nn = o->et.nnode, nC2 = (int)o->ft->numfields_C2, nC1 = (int)o->ft->numfields_C1;
  const int *en = o->elem_nodes + (size_t)e * nn;
  ts->elem = e;
  int nloc = 0;
  for (int l = 0; l < nn; l++)
  {
    const int node = en[l];
    ts->node_of[l] = node;
    for (int i = 0; i < dim; i++)
    {
      ts->coord_slots[l][i] = o->pos + ((size_t)node * dim + i) * o->n_pos_hist;
      ts->coord_slots[l][dim + i] = o->lagr + (size_t)node * dim + i;
    }
    for (int f = 0; f < nC2; f++) ts->data_slots[l][f] = o->val + ((size_t)node * o->nval + f) * o->T;
  }
  for (int l = 0; l < o->et.nnode_C1; l++)
  {
    const int node = en[o->et.c1_nodes[l]];
    for (int f = nC2; f < nC2 + nC1; f++) ts->data_slots[l][f] = o->val + ((size_t)node * o->nval + f) * o->T;
  }
  /* local equations of nodal values in element-node order, value order */
  for (int l = 0; l < nn; l++)
    for (int f = 0; f < nC2 + nC1; f++) ts->eqn_slots[l][f] = -1;
  for (int l = 0; l < nn; l++)
  {
    const int node = en[l];
    int c1l = -1;
    for (int k = 0; k < o->et.nnode_C1; k++)
      if (o->et.c1_nodes[k] == l) c1l = k;
    for (int f = 0; f < nC2 + nC1; f++)
    {
      const int g = o->node_eqn[(size_t)node * o->nval + f];
      if (g < 0) continue;
      if (f < nC2)
      {
        ts->eqn_slots[l][f] = nloc;
        ts->eqn_of_local[nloc++] = g;
      }
      else if (c1l >= 0)
      {
        ts->eqn_slots[c1l][f] = nloc;
        ts->eqn_of_local[nloc++] = g;
      }
    }
  }
  for (int l = 0; l < nn; l++)
    for (int i = 0; i < dim; i++)
    {
      const int g = o->pos_eqn ? o->pos_eqn[(size_t)en[l] * dim + i] : -1;
      ts->poseqn_slots[l][i] = -1;
      if (g >= 0)
      {
        ts->poseqn_slots[l][i] = nloc;
        ts->eqn_of_local[nloc++] = g;
      }
    }
  /* hanging nodes: fill_hang_info_with_equations (src/elements.cpp:812-1160) on top of RefineableElement::assign_hanging_local_eqn_numbers
   * (oomph-lib refineable_elements.cc:312-470): nodes n, values j, masters m; a master value that is already a local dof keeps its
   * number, a new one is appended behind the element's own dofs; pinned master values give -1. */
  JITShapeInfo_t *si = &ts->si;
  si->hanginfo_C1 = ts->nohang;
  si->hanginfo_C2 = ts->nohang;
  si->hanginfo_Pos = ts->nohang;
  if (o->hang_start[0] || o->hang_start[1])
  {
    si->hanginfo_C2 = ts->hang_C2;
    si->hanginfo_C1 = ts->hang_C1;
    if (o->pos_eqn)
    {
      /* moving mesh: the POSITIONS of a (geometrically) hanging node hang on the positions of its masters (hanginfo_Pos,
       * src/elements.cpp:820-862; local_position_hang_eqn of oomph-lib's RefineableSolidElement) */
      si->hanginfo_Pos = ts->hang_Pos;
      for (int l = 0; l < nn; l++)
      {
        const int node = en[l];
        JITHangInfo_t *hi = &ts->hang_Pos[l];
        JITHangInfoEntry_t *ent = ts->hang_entries_Pos[l];
        hi->nummaster = 0;
        hi->masters = ent;
        if (!o->hang_start[0]) continue;
        const int h0 = o->hang_start[0][node], h1 = o->hang_start[0][node + 1];
        if (h1 == h0) continue;
        hi->nummaster = h1 - h0;
        for (int m = 0; m < h1 - h0; m++)
        {
          ent[m].weight = o->hang_weight[0][h0 + m];
          ent[m].local_eqn = ts->hang_eqn_Pos[l][m];
          for (int d = 0; d < dim; d++)
          {
            const int g = o->pos_eqn[(size_t)o->hang_master[0][h0 + m] * dim + d];
            int loc = -1;
            if (g >= 0)
            {
              for (int k = 0; k < nloc; k++)
                if (ts->eqn_of_local[k] == g) loc = k;
              if (loc < 0)
              {
                loc = nloc;
                ts->eqn_of_local[nloc++] = g;
              }
            }
            ent[m].local_eqn[d] = loc;
          }
        }
      }
    }
    for (int sp = 0; sp < 2; sp++)
    {
      const int nl = sp == 0 ? nn : o->et.nnode_C1, f0 = sp == 0 ? 0 : nC2, f1 = sp == 0 ? nC2 : nC2 + nC1;
      for (int l = 0; l < nl; l++)
      {
        const int node = sp == 0 ? en[l] : en[o->et.c1_nodes[l]];
        JITHangInfo_t *hi = sp == 0 ? &ts->hang_C2[l] : &ts->hang_C1[l];
        JITHangInfoEntry_t *ent = sp == 0 ? ts->hang_entries_C2[l] : ts->hang_entries_C1[l];
        hi->nummaster = 0;
        hi->masters = ent;
        if (!o->hang_start[sp]) continue;
        const int h0 = o->hang_start[sp][node], h1 = o->hang_start[sp][node + 1];
        if (h1 == h0) continue;
        if (h1 - h0 > MAXM)
        {
          fprintf(stderr, "oracle: more than %d masters\n", MAXM);
          abort();
        }
        hi->nummaster = h1 - h0;
        for (int m = 0; m < h1 - h0; m++)
        {
          ent[m].weight = o->hang_weight[sp][h0 + m];
          ent[m].local_eqn = sp == 0 ? ts->hang_eqn_C2[l][m] : ts->hang_eqn_C1[l][m];
          for (int f = 0; f < nC2 + nC1; f++) ent[m].local_eqn[f] = -1;
        }
        for (int f = f0; f < f1; f++)
          for (int m = 0; m < h1 - h0; m++)
          {
            const int g = o->node_eqn[(size_t)o->hang_master[sp][h0 + m] * o->nval + f];
            if (g < 0) continue;
            int loc = -1;
            for (int k = 0; k < nloc; k++)
              if (ts->eqn_of_local[k] == g) loc = k;
            if (loc < 0)
            {
              loc = nloc;
              ts->eqn_of_local[nloc++] = g;
            }
            ent[m].local_eqn[f] = loc;
          }
      }
    }
  }
  ts->ei.ndof = nloc;
}

/* the elements are FACES (s1 = -1) of Q9 bulk elements given by their nine (rotated) bulk nodes */
void oracle_set_face_mode(void *h)
{
  Oracle *o = (Oracle *)h;
  if (o->et.dim != 2 || o->et.nnode != 9 || o->pos_eqn) { fprintf(stderr, "oracle: face mode needs Q9 elements on a fixed mesh\n"); abort(); }
  o->et.face = 1;
  o->et.n_int = 3;
}

/* hanging nodes of one space (0: C2 and positions, 1: C1): CSR over the nodes */
void oracle_set_hanging(void *h, int space, const int *start, const int *masters, const double *weights)
{
  Oracle *o = (Oracle *)h;
  const int n = start[o->n_node];
  free(o->hang_start[space]);
  free(o->hang_master[space]);
  free(o->hang_weight[space]);
  o->hang_start[space] = (int *)xcalloc((size_t)o->n_node + 1, sizeof(int));
  o->hang_master[space] = (int *)xcalloc((size_t)n + 1, sizeof(int));
  o->hang_weight[space] = (double *)xcalloc((size_t)n + 1, sizeof(double));
  memcpy(o->hang_start[space], start, ((size_t)o->n_node + 1) * sizeof(int));
  memcpy(o->hang_master[space], masters, (size_t)n * sizeof(int));
  memcpy(o->hang_weight[space], weights, (size_t)n * sizeof(double));
}

/* prepare_shape_buffer_for_integration (src/elements.cpp:4577-4646) */
static void prepare_shape_buffer(ThreadState *ts)
{
  Oracle *o = ts->o;
  JITShapeInfo_t *si = &ts->si;
  si->n_int_pt = o->et.n_int;
  if (o->steady)
  {
    si->timestepper_ntstorage = 0;
    for (int i = 0; i < NTW; i++)
    {
      si->timestepper_weights_dt_BDF1[i] = 0;
      si->timestepper_weights_dt_BDF2[i] = 0;
      si->timestepper_weights_dt_Newmark2[i] = 0;
      si->timestepper_weights_d2t_Newmark2[i] = 0;
    }
    si->timestepper_weights_dt_BDF2_degr = si->timestepper_weights_dt_BDF2;
    si->timestepper_weights_dt_Newmark2_degr = si->timestepper_weights_dt_Newmark2;
  }
  else
  {
    si->timestepper_ntstorage = o->ntstorage;
    for (int i = 0; i < NTW; i++)
    {
      si->timestepper_weights_dt_BDF1[i] = o->wBDF1[i];
      si->timestepper_weights_dt_BDF2[i] = o->wBDF2[i];
      si->timestepper_weights_dt_Newmark2[i] = o->wNM2[i];
      si->timestepper_weights_d2t_Newmark2[i] = o->wNM2_d2t[i];
    }
    if (o->unsteady_steps_done == 0)
    {
      si->timestepper_weights_dt_BDF2_degr = si->timestepper_weights_dt_BDF1;
      si->timestepper_weights_dt_Newmark2_degr = si->timestepper_weights_dt_BDF1;
    }
    else if (o->unsteady_steps_done <= 4)
    {
      si->timestepper_weights_dt_BDF2_degr = si->timestepper_weights_dt_BDF2;
      si->timestepper_weights_dt_Newmark2_degr = si->timestepper_weights_dt_BDF2;
    }
    else
    {
      si->timestepper_weights_dt_BDF2_degr = si->timestepper_weights_dt_BDF2;
      si->timestepper_weights_dt_Newmark2_degr = si->timestepper_weights_dt_Newmark2;
    }
  }
  for (int i = 0; i < NTW; i++)
  {
    si->t[i] = o->t[i];
    si->dt[i] = o->dt[i];
  }
}

/* fill_in_generic_residual_contribution_jit (src/elements.cpp:5054-5127): R,J,M are caller-zeroed [ndof],[ndof^2] */
static void element_rjm_bound(ThreadState *ts, int which, int param, unsigned flag, double *R, double *J, double *M)
{
  Oracle *o = ts->o;
  TS = ts;
  prepare_shape_buffer(ts);
  ts->si.jacobian_size = ts->ei.ndof;
  ts->si.mass_matrix_size = ts->ei.ndof;
  /* fill_shape_info_element_sizes (src/elements.cpp:3527-3568): sum over all integration points of w * J, times the coordinate
   * system's JacobianForElementSize at the point for the non-Cartesian size */
  const JITFuncSpec_RequiredShapes_FiniteElement_t *rq = &o->ft->shapes_required_ResJac[which];
  if (rq->elemsize_Eulerian_Pos || rq->elemsize_Eulerian_cartesian_Pos || rq->elemsize_Lagrangian_Pos || rq->elemsize_Lagrangian_cartesian_Pos)
  {
    double esz = 0.0, esz_cart = 0.0, eszL = 0.0, eszL_cart = 0.0;
    for (int q = 0; q < o->et.n_int; q++)
    {
      double s[MAXD], w;
      if (o->et.tri && o->et.dim == 3) oracle_gauss_tet(q, s, &w);
      else if (o->et.tri) oracle_gauss_tri(q, s, &w);
      else if (o->et.edim == 1) oracle_gauss_1d(q, s, &w);
      else oracle_gauss(o->et.dim, q, s, &w);
      fill_shape_info_at_s(ts, s, w, 0u, NULL);
      double x[MAXD] = {0, 0, 0};
      for (int l = 0; l < o->et.nnode; l++)
        for (int i = 0; i < o->et.dim; i++) x[i] += ts->ei.nodal_coords[l][i][0] * ts->si.shape_C2[l];
      esz_cart += ts->si.int_pt_weight;
      esz += ts->si.int_pt_weight * o->ft->JacobianForElementSize(&ts->ei, x);
      /* the Lagrangian sizes: interpolated_xi and J_lagrangian_at_knot (src/elements.cpp:3546-3551, :3569-3574) */
      double xi[MAXD] = {0, 0, 0};
      for (int l = 0; l < o->et.nnode; l++)
        for (int i = 0; i < o->et.dim; i++) xi[i] += ts->ei.nodal_coords[l][o->et.dim + i][0] * ts->si.shape_C2[l];
      eszL_cart += ts->si.int_pt_weight_Lagrangian;
      eszL += ts->si.int_pt_weight_Lagrangian * o->ft->JacobianForElementSize(&ts->ei, xi);
    }
    ts->si.elemsize_Eulerian = esz;
    ts->si.elemsize_Eulerian_cartesian = esz_cart;
    ts->si.elemsize_Lagrangian = eszL;
    ts->si.elemsize_Lagrangian_cartesian = eszL_cart;
  }
  JITFuncSpec_ResidualAndJacobian_FiniteElement func;
  if (param >= 0)
    func = o->ft->ParameterDerivative[which][param];
  else if (o->steady && o->ft->ResidualAndJacobianSteady && o->ft->ResidualAndJacobianSteady[which])
    func = o->ft->ResidualAndJacobianSteady[which];
  else
    func = o->ft->ResidualAndJacobian[which];
  func(&ts->ei, &ts->si, R, J, M, flag);
}

static void element_rjm(ThreadState *ts, int e, int which, int param, unsigned flag, double *R, double *J, double *M)
{
  bind_element(ts, e);
  element_rjm_bound(ts, which, param, flag, R, J, M);
}

/* ------------------------------------------------------------------ public C entry points (ctypes) */
void *oracle_create_typed(int dim, int nnode, int n_elem, const int *elem_nodes, int n_node, int nval, int T, int n_pos_hist,
                          const double *node_pos, const double *node_lagr, const double *node_val, const int *node_eqn,
                          const int *pos_eqn, int n_dof)
{
  init_tables();
  Oracle *o = (Oracle *)xcalloc(1, sizeof(Oracle));
  static const int c1q[4] = {0, 2, 6, 8}, c1b[8] = {0, 2, 6, 8, 18, 20, 24, 26}, c1t[3] = {0, 1, 2};
  static const int c1l[2] = {0, 2};
  o->et.dim = dim;
  o->et.edim = dim;
  o->et.tri = (dim == 2 && nnode == 6) || (dim == 3 && nnode == 10); /* simplex elements: TElement<2,3>, TElement<3,3> */
  if (dim == 2 && nnode == 3)
  {
    /* InterfaceElementLine1dC2: QElement<1,3> in a 2D space, C1 on the end nodes, Gauss<1,3> */
    o->et.edim = 1;
    o->et.nnode = 3;
    o->et.nnode_C1 = 2;
    o->et.n_int = 3;
    memcpy(o->et.c1_nodes, c1l, sizeof(c1l));
  }
  else if (o->et.tri && dim == 3)
  {
    /* BulkElementTetra3dC2 (src/elements.cpp:11397-11409): 10 nodes, C1 on the vertices, TGauss<3,3> */
    static const int c1tet[4] = {0, 1, 2, 3};
    o->et.nnode = 10;
    o->et.nnode_C1 = 4;
    o->et.n_int = 11;
    memcpy(o->et.c1_nodes, c1tet, sizeof(c1tet));
  }
  else if (o->et.tri)
  {
    /* BulkElementTri2dC2 (src/elements.cpp:9844-9856): 6 nodes, C1 on the vertices, TGauss<2,3> */
    o->et.nnode = 6;
    o->et.nnode_C1 = 3;
    o->et.n_int = 7;
    memcpy(o->et.c1_nodes, c1t, sizeof(c1t));
  }
  else
  {
    o->et.nnode = dim == 2 ? 9 : 27;
    o->et.nnode_C1 = dim == 2 ? 4 : 8;
    o->et.n_int = dim == 2 ? 9 : 27;
    memcpy(o->et.c1_nodes, dim == 2 ? c1q : c1b, sizeof(int) * o->et.nnode_C1);
  }
  if (nnode != o->et.nnode) { fprintf(stderr, "oracle: unsupported element (dim %d, %d nodes)\n", dim, nnode); abort(); }
  o->n_elem = n_elem;
  o->n_node = n_node;
  o->nval = nval;
  o->T = T;
  o->n_pos_hist = n_pos_hist;
  o->n_dof = n_dof;
  o->elem_nodes = elem_nodes;
  o->node_eqn = node_eqn;
  o->pos_eqn = pos_eqn;
  /* transpose [t][node][k] -> [node][k][t] (oomph Data keeps a value's history contiguous) */
  o->pos = (double *)xcalloc((size_t)n_node * dim * n_pos_hist, sizeof(double));
  for (int t = 0; t < n_pos_hist; t++)
    for (size_t n = 0; n < (size_t)n_node; n++)
      for (int i = 0; i < dim; i++) o->pos[(n * dim + i) * n_pos_hist + t] = node_pos[((size_t)t * n_node + n) * dim + i];
  o->lagr = (double *)xcalloc((size_t)n_node * dim, sizeof(double));
  memcpy(o->lagr, node_lagr, sizeof(double) * (size_t)n_node * dim);
  o->val = (double *)xcalloc((size_t)n_node * nval * T, sizeof(double));
  for (int t = 0; t < T; t++)
    for (size_t n = 0; n < (size_t)n_node; n++)
      for (int f = 0; f < nval; f++) o->val[(n * nval + f) * T + t] = node_val[((size_t)t * n_node + n) * nval + f];
  o->ft = (JITFuncSpec_Table_FiniteElement_t *)xcalloc(1, sizeof(JITFuncSpec_Table_FiniteElement_t));
  o->ft->check_compiler_size = check_size;
  JIT_ELEMENT_init(o->ft);
  o->ft->fill_shape_buffer_for_point = cb_fill_shape_buffer_for_point;
  o->params = (double *)xcalloc(o->ft->numglobal_params + 1, sizeof(double));
  for (unsigned k = 0; k < o->ft->numglobal_params; k++) o->ft->global_parameters[k] = &o->params[k];
  o->steady = 1;
  if ((int)o->ft->nodal_dim != dim) { fprintf(stderr, "oracle: plugin dimension mismatch\n"); abort(); }
  return o;
}

void *oracle_create(int dim, int n_elem, const int *elem_nodes, int n_node, int nval, int T, int n_pos_hist,
                    const double *node_pos, const double *node_lagr, const double *node_val, const int *node_eqn,
                    const int *pos_eqn, int n_dof)
{
  return oracle_create_typed(dim, dim == 2 ? 9 : 27, n_elem, elem_nodes, n_node, nval, T, n_pos_hist, node_pos, node_lagr, node_val, node_eqn, pos_eqn, n_dof);
}

int oracle_num_params(void *h) { return (int)((Oracle *)h)->ft->numglobal_params; }
int oracle_num_residuals(void *h) { return (int)((Oracle *)h)->ft->num_res_jacs; }
int oracle_moving_nodes(void *h) { return (int)((Oracle *)h)->ft->moving_nodes; }

void oracle_set_params(void *h, const double *p, int n)
{
  Oracle *o = (Oracle *)h;
  for (int k = 0; k < n && k < (int)o->ft->numglobal_params; k++) o->params[k] = p[k];
}

void oracle_set_time(void *h, int steady, int unsteady_steps_done, int ntstorage, const double *t, const double *dt,
                     const double *wBDF1, const double *wBDF2, const double *wNM2, const double *wNM2_d2t)
{
  Oracle *o = (Oracle *)h;
  o->steady = steady;
  o->unsteady_steps_done = unsteady_steps_done;
  o->ntstorage = ntstorage;
  for (int i = 0; i < NTW; i++)
  {
    o->t[i] = t[i];
    o->dt[i] = dt[i];
    o->wBDF1[i] = wBDF1[i];
    o->wBDF2[i] = wBDF2[i];
    o->wNM2[i] = wNM2[i];
    o->wNM2_d2t[i] = wNM2_d2t[i];
  }
}

/* MultiTimeStepper::set_weights (src/timestepper.cpp:31-72), BDF part */
void oracle_bdf_weights(double dt, double dtprev, double *wBDF1, double *wBDF2)
{
  for (int i = 0; i < NTW; i++) wBDF1[i] = wBDF2[i] = 0.0;
  wBDF2[0] = 1.0 / dt + 1.0 / (dt + dtprev);
  wBDF2[1] = -(dt + dtprev) / (dt * dtprev);
  wBDF2[2] = dt / ((dt + dtprev) * dtprev);
  wBDF1[0] = 1.0 / dt;
  wBDF1[1] = -1.0 / dt;
}

/* MultiTimeStepper::set_weights (src/timestepper.cpp:61-80), Newmark2 part: NSTEPS = 2 (src/timestepper.hpp:49), NewmarkBeta1 =
 * NewmarkBeta2 = 0.5 unless setNewmark2Coeffs was called (src/timestepper.hpp:63, :109) */
void oracle_newmark2_weights(double dt, double beta1, double beta2, double *w_dt, double *w_d2t)
{
  const int NSTEPS = 2;
  for (int i = 0; i < NTW; i++) w_dt[i] = w_d2t[i] = 0.0;
  w_d2t[0] = 2.0 / (beta2 * dt * dt);
  w_d2t[1] = -2.0 / (beta2 * dt * dt);
  w_d2t[NSTEPS + 1] = -2.0 / (dt * beta2);
  w_d2t[NSTEPS + 2] = (beta2 - 1.0) / beta2;
  w_dt[0] = beta1 * dt * w_d2t[0];
  w_dt[1] = beta1 * dt * w_d2t[1];
  w_dt[NSTEPS + 1] = 1.0 + beta1 * dt * w_d2t[NSTEPS + 1];
  w_dt[NSTEPS + 2] = dt * (1.0 - beta1) + beta1 * dt * w_d2t[NSTEPS + 2];
}

/* update nodal values/positions at history level t from [node][k] arrays */
void oracle_update_values(void *h, int t, const double *node_val, const double *node_pos)
{
  Oracle *o = (Oracle *)h;
  if (node_val)
    for (size_t n = 0; n < (size_t)o->n_node; n++)
      for (int f = 0; f < o->nval; f++) o->val[(n * o->nval + f) * o->T + t] = node_val[n * o->nval + f];
  if (node_pos)
    for (size_t n = 0; n < (size_t)o->n_node; n++)
      for (int i = 0; i < o->et.dim; i++) o->pos[(n * o->et.dim + i) * o->n_pos_hist + t] = node_pos[n * o->et.dim + i];
}

/* one element: dense R[ndof], J[ndof^2], M[ndof^2] in the oracle's (= oomph's) local order + global eqn map */
int oracle_element(void *h, int e, int which, int param, unsigned flag, double *R, double *J, double *M, int *eqns)
{
  Oracle *o = (Oracle *)h;
  ThreadState *ts = ts_create(o);
  element_rjm(ts, e, which, param, flag, R, J, M);
  const int n = (int)ts->ei.ndof;
  for (int i = 0; i < n; i++) eqns[i] = ts->eqn_of_local[i];
  free(ts); /* small leak of the shape tables: test-only entry point */
  return n;
}

/* shape buffer of one Gauss point of one element, for the geometry tests */
void oracle_point_shapes(void *h, int e, int ipt, unsigned flag, double *weights /*3*/, double *shape, double *dx_shape,
                         double *dX_shape, double *w_dcoords, double *d_dx_dcoord)
{
  Oracle *o = (Oracle *)h;
  ThreadState *ts = ts_create(o);
  const int dim = o->et.dim, nn = o->et.nnode;
  TS = ts;
  bind_element(ts, e);
  prepare_shape_buffer(ts);
  cb_fill_shape_buffer_for_point((unsigned)ipt, NULL, (int)flag);
  weights[0] = ts->si.int_pt_weight;
  weights[1] = ts->si.int_pt_weight_Lagrangian;
  weights[2] = ts->si.int_pt_weight_unity;
  for (int l = 0; l < nn; l++)
  {
    shape[l] = ts->si.shape_C2[l];
    for (int i = 0; i < dim; i++)
    {
      dx_shape[l * dim + i] = ts->si.dx_shape_C2[l][i];
      dX_shape[l * dim + i] = ts->si.dX_shape_C2[l][i];
      if (w_dcoords) w_dcoords[i * nn + l] = ts->si.int_pt_weights_d_coords[i][l];
      if (d_dx_dcoord)
        for (int l2 = 0; l2 < nn; l2++)
          for (int i2 = 0; i2 < dim; i2++) d_dx_dcoord[((l * dim + i) * nn + l2) * dim + i2] = ts->si.d_dx_shape_dcoord_C2[l][i][l2][i2];
    }
  }
  free(ts);
}

static void row_add(Row *r, int col, double v)
{
  for (int k = 0; k < r->n; k++)
    if (r->p[k].col == col)
    {
      r->p[k].val += v;
      return;
    }
  if (r->n == r->cap)
  {
    r->cap = r->cap ? 2 * r->cap : 32;
    r->p = (Pair *)realloc(r->p, sizeof(Pair) * r->cap);
  }
  r->p[r->n].col = col;
  r->p[r->n].val = v;
  r->n++;
}

/* sparse_assemble_row_or_column_compressed_with_vectors_of_pairs (problem.cc:5332-5666).  flag: 0 residual only,
 * 1 +Jacobian, 2 +mass matrix.  nthreads>1 splits the element range statically like First_el_for_assembly
 * (problem.cc:5354) with one private matrix per range, merged in range order afterwards. */
double oracle_assemble(void *h, int which, int param, unsigned flag, double *residuals, int nthreads)
{
  Oracle *o = (Oracle *)h;
  const int nmat = flag == 0 ? 0 : (flag == 1 ? 1 : 2);
  if (nthreads < 1) nthreads = 1;
  for (int m = 0; m < 2; m++)
  {
    free(o->row_start[m]); free(o->col_index[m]); free(o->value[m]);
    o->row_start[m] = o->col_index[m] = NULL; o->value[m] = NULL; o->nnz[m] = 0;
  }
  Row **rows = (Row **)xcalloc((size_t)nthreads * 2, sizeof(Row *));
  double **res_t = (double **)xcalloc(nthreads, sizeof(double *));
  for (int th = 0; th < nthreads; th++)
  {
    for (int m = 0; m < nmat; m++) rows[th * 2 + m] = (Row *)xcalloc(o->n_dof, sizeof(Row));
    res_t[th] = th == 0 ? residuals : (double *)xcalloc(o->n_dof, sizeof(double));
  }
  memset(residuals, 0, sizeof(double) * o->n_dof);
  const int timing = getenv("ORACLE_TIMING") != NULL;
  double tm0 = 0.0, tm1 = 0.0, tm2 = 0.0;
#ifdef _OPENMP
  tm0 = omp_get_wtime();
#pragma omp parallel num_threads(nthreads)
#endif
  {
#ifdef _OPENMP
    const int th = omp_get_thread_num();
#else
    const int th = 0;
#endif
    ThreadState *ts = ts_create(o);
    const int maxdof = 2 * o->et.nnode * (o->et.dim + o->nval); /* x2: master values outside the element (hanging nodes) */
    double *R = (double *)xcalloc(maxdof, sizeof(double));
    double *J = (double *)xcalloc((size_t)maxdof * maxdof, sizeof(double));
    double *M = (double *)xcalloc((size_t)maxdof * maxdof, sizeof(double));
    const int lo = (int)((int64_t)o->n_elem * th / nthreads), hi = (int)((int64_t)o->n_elem * (th + 1) / nthreads);
    for (int e = lo; e < hi; e++)
    {
      bind_element(ts, e);
      const int n = (int)ts->ei.ndof;
      memset(R, 0, sizeof(double) * n);
      if (nmat > 0) memset(J, 0, sizeof(double) * n * n);
      if (nmat > 1) memset(M, 0, sizeof(double) * n * n);
      element_rjm_bound(ts, which, param, flag, R, nmat > 0 ? J : NULL, nmat > 1 ? M : NULL);
      for (int i = 0; i < n; i++)
      {
        const int eqn = ts->eqn_of_local[i];
        res_t[th][eqn] += R[i];
        for (int j = 0; j < n; j++)
        {
          const int unknown = ts->eqn_of_local[j];
          for (int m = 0; m < nmat; m++)
          {
            const double value = (m == 0 ? J : M)[i * n + j];
            if (fabs(value) > 0.0) row_add(&rows[th * 2 + m][eqn], unknown, value);
          }
        }
      }
    }
    free(R); free(J); free(M); free(ts);
  }
  /* merge per-range results in range order (single range: no-op).  Rows are independent, so the merge runs over row blocks in
   * parallel: the order of the contributions to one row is still the range order (same result as a serial merge); this is what
   * the reference gets from MPI ranks that each assemble their own rows (problem.cc:6543), without a serial bottleneck that
   * would undersell the CPU baseline */
#ifdef _OPENMP
  tm1 = omp_get_wtime();
#endif
  if (nthreads > 1)
  {
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int i = 0; i < o->n_dof; i++)
      for (int th = 1; th < nthreads; th++)
      {
        residuals[i] += res_t[th][i];
        for (int m = 0; m < nmat; m++)
        {
          Row *r = &rows[th * 2 + m][i], *dst = &rows[m][i];
          if (r->n == 0) { free(r->p); continue; }
          if (dst->n == 0)
          {
            /* the row has been touched by no earlier range (true for all rows but those of the interface nodes between two element
             * ranges): adopt it as it is -- same entries, same order as merging it entry by entry */
            free(dst->p);
            *dst = *r;
            continue;
          }
          for (int k = 0; k < r->n; k++) row_add(dst, r->p[k].col, r->p[k].val);
          free(r->p);
        }
      }
    for (int th = 1; th < nthreads; th++)
    {
      free(res_t[th]);
      for (int m = 0; m < nmat; m++) free(rows[th * 2 + m]);
    }
  }
#ifdef _OPENMP
  tm2 = omp_get_wtime();
#endif
  for (int m = 0; m < nmat; m++)
  {
    Row *rw = rows[m];
    o->row_start[m] = (int *)xcalloc(o->n_dof + 1, sizeof(int));
    for (int i = 0; i < o->n_dof; i++) o->row_start[m][i + 1] = o->row_start[m][i] + rw[i].n;
    const int entries = o->row_start[m][o->n_dof];
    o->nnz[m] = entries;
    o->col_index[m] = (int *)malloc(sizeof(int) * (size_t)(entries > 0 ? entries : 1));
    o->value[m] = (double *)malloc(sizeof(double) * (size_t)(entries > 0 ? entries : 1));
    /* rows are independent: the copy into the CSR arrays runs over row blocks (same arrays as the serial loop of problem.cc:5600-5659) */
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (int i = 0; i < o->n_dof; i++)
    {
      int p = 0;
      for (int j = o->row_start[m][i]; j < o->row_start[m][i + 1]; j++, p++)
      {
        o->col_index[m][j] = rw[i].p[p].col;
        o->value[m][j] = rw[i].p[p].val;
      }
      free(rw[i].p);
    }
    free(rw);
  }
  free(rows);
  free(res_t);
#ifdef _OPENMP
  if (timing) fprintf(stderr, "[oracle timing] %d threads: element loop %.1f ms, merge %.1f ms, CSR build %.1f ms\n", nthreads, (tm1 - tm0) * 1e3, (tm2 - tm1) * 1e3, (omp_get_wtime() - tm2) * 1e3);
#endif
  (void)timing; (void)tm0; (void)tm1; (void)tm2;
  return (double)o->nnz[0];
}

/* fill_in_generic_hessian / get_multi_assembly Hessian part (src/elements.cpp:5512, :4983-4988): local slices of the global
 * vectors are handed to HessianVectorProduct<i>; outputs are dense per element.  flag as in SURVEY A.5.
 *   flag 0: Yg = one vector [n_dof], Cg = nvec vectors [nvec][n_dof]; product[nvec][ndof]
 *   flag 1,2,4,5: Yg = nvec vectors; product (and Cs for 2,5) [nvec][ndof][ndof] */
int oracle_element_hessian(void *h, int e, int which, const double *Yg, const double *Cg, int nvec, unsigned flag, double *product, double *Cs, int *eqns)
{
  Oracle *o = (Oracle *)h;
  ThreadState *ts = ts_create(o);
  TS = ts;
  bind_element(ts, e);
  prepare_shape_buffer(ts);
  const int n = (int)ts->ei.ndof;
  ts->si.jacobian_size = n;
  ts->si.mass_matrix_size = n;
  const int ny = flag == 0 ? 1 : nvec;
  double *Yl = (double *)xcalloc((size_t)ny * n, sizeof(double));
  for (int v = 0; v < ny; v++)
    for (int j = 0; j < n; j++) Yl[v * n + j] = Yg[(size_t)v * o->n_dof + ts->eqn_of_local[j]];
  double *Cl = Cs;
  if (flag == 0)
  {
    Cl = (double *)xcalloc((size_t)nvec * n, sizeof(double));
    for (int v = 0; v < nvec; v++)
      for (int j = 0; j < n; j++) Cl[v * n + j] = Cg[(size_t)v * o->n_dof + ts->eqn_of_local[j]];
  }
  o->ft->HessianVectorProduct[which](&ts->ei, &ts->si, Yl, Cl, product, (unsigned)nvec, flag);
  for (int i = 0; i < n; i++) eqns[i] = ts->eqn_of_local[i];
  free(Yl);
  if (flag == 0) free(Cl);
  free(ts);
  return n;
}

/* Mesh::evaluate_integral_expression (src/mesh.cpp:536-548): sum over the elements, in mesh order, of
 * BulkElementBase::eval_integral_expression (src/elements.cpp:4648-4657: prepare the shape buffer, call the generated routine) */
int oracle_num_integrals(void *h) { return (int)((Oracle *)h)->ft->numintegral_expressions; }
const char *oracle_integral_name(void *h, int i) { return ((Oracle *)h)->ft->integral_expressions_names[i]; }
double oracle_eval_integral(void *h, int index)
{
  Oracle *o = (Oracle *)h;
  if (!o->ft->EvalIntegralExpression || index < 0 || (unsigned)index >= o->ft->numintegral_expressions) return NAN;
  ThreadState *ts = ts_create(o);
  TS = ts;
  double res = 0.0;
  for (int e = 0; e < o->n_elem; e++)
  {
    bind_element(ts, e);
    prepare_shape_buffer(ts);
    res += o->ft->EvalIntegralExpression(&ts->ei, &ts->si, (unsigned)index);
  }
  free(ts);
  return res;
}

/* expressions at a local coordinate of one element: BulkElementBase::eval_local_expression_at_s / eval_extremum_expression_at_s /
 * get_Z2_flux (src/elements.cpp:4666-4704, :7285-7305): fill the shape buffer at s (weight 0: these carry no measure), prepare the
 * time weights, call the generated routine.  kind 0 local (returns the value of expression `index`), 1 extremum, 2 Z2 fluxes (all
 * flux terms into out[], returns their number). */
int oracle_num_point_exprs(void *h, int kind)
{
  const JITFuncSpec_Table_FiniteElement_t *ft = ((Oracle *)h)->ft;
  return kind == 0 ? (int)ft->numlocal_expressions : kind == 1 ? (int)ft->numextremum_expressions : (int)ft->num_Z2_flux_terms;
}
double oracle_eval_at_s(void *h, int kind, int index, int e, const double *s, double *out)
{
  Oracle *o = (Oracle *)h;
  ThreadState *ts = ts_create(o);
  TS = ts;
  bind_element(ts, e);
  fill_shape_info_at_s(ts, s, 0.0, 0u, NULL);
  prepare_shape_buffer(ts);
  double res = 0.0;
  if (kind == 0 && o->ft->EvalLocalExpression) res = o->ft->EvalLocalExpression(&ts->ei, &ts->si, (unsigned)index);
  else if (kind == 1 && o->ft->EvalExtremumExpression) res = o->ft->EvalExtremumExpression(&ts->ei, &ts->si, (unsigned)index);
  else if (kind == 2 && o->ft->GetZ2Fluxes)
  {
    o->ft->GetZ2Fluxes(&ts->ei, &ts->si, out);
    res = (double)o->ft->num_Z2_flux_terms;
  }
  else
    res = NAN;
  free(ts);
  return res;
}

int64_t oracle_nnz(void *h, int m) { return ((Oracle *)h)->nnz[m]; }
void oracle_get_csr(void *h, int m, int *row_start, int *col_index, double *value)
{
  Oracle *o = (Oracle *)h;
  memcpy(row_start, o->row_start[m], sizeof(int) * (o->n_dof + 1));
  memcpy(col_index, o->col_index[m], sizeof(int) * o->nnz[m]);
  memcpy(value, o->value[m], sizeof(double) * o->nnz[m]);
}

void oracle_free(void *h)
{
  Oracle *o = (Oracle *)h;
  for (int m = 0; m < 2; m++) { free(o->row_start[m]); free(o->col_index[m]); free(o->value[m]); }
  if (o->ft->clean_up) o->ft->clean_up(o->ft);
  free(o->ft); free(o->pos); free(o->lagr); free(o->val); free(o->params);
  for (int sp = 0; sp < 2; sp++) { free(o->hang_start[sp]); free(o->hang_master[sp]); free(o->hang_weight[sp]); }
  free(o);
}
