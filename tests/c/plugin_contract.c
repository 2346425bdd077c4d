/* Test program for the plugin contract (tests/test_plugin_contract.py builds and runs it): loads a CUDA plugin the way the reference's
 * host does -- zero-filled JITFuncSpec_Table_FiniteElement_t, check_compiler_size set first, JIT_ELEMENT_init through dlsym
 * (/root/reference/src/problem.cpp:101-142, src/ccompiler.cpp:191-243), host callbacks installed afterwards, clean_up at the end
 * (src/problem.cpp:155) -- and next to it JIT_ELEMENT_init_cuda; compares the two tables' metadata; with a problem file as third
 * argument it also runs ONE assembly through the engine's C-ABI (pb2_problem_assemble_host) and writes residual and CSR values.
 * Compiled against the reference's own jitbridge.h.  usage: plugin_contract <plugin.so> <libpyoomph_b200.so> [problem.bin out.bin] */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jitbridge.h"
#include "pyoomph_b200.h"

static int size_mismatches = 0, size_checks = 0;
static void check_compiler_size(unsigned long long a, unsigned long long b, char *name)
{
  size_checks++;
  if (a != b)
  {
    size_mismatches++;
    fprintf(stderr, "size mismatch for %s: %llu vs %llu\n", name, a, b);
  }
}
static double dummy_get_element_size(void *p) { (void)p; return 0.0; }

#define CHECK(cond)                                                         \
  do                                                                        \
  {                                                                         \
    if (!(cond))                                                            \
    {                                                                       \
      fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #cond);             \
      return 1;                                                             \
    }                                                                       \
  } while (0)

int main(int argc, char **argv)
{
  if (argc < 3) return 2;
  void *h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 1; }
  JIT_ELEMENT_init_SPEC init = (JIT_ELEMENT_init_SPEC)dlsym(h, "JIT_ELEMENT_init");
  JIT_ELEMENT_init_cuda_SPEC init_cuda = (JIT_ELEMENT_init_cuda_SPEC)dlsym(h, "JIT_ELEMENT_init_cuda");
  CHECK(init && init_cuda);
  /* exactly the host's sequence */
  JITFuncSpec_Table_FiniteElement_t *ft = (JITFuncSpec_Table_FiniteElement_t *)malloc(sizeof(JITFuncSpec_Table_FiniteElement_t));
  memset(ft, 0, sizeof(JITFuncSpec_Table_FiniteElement_t));
  ft->check_compiler_size = check_compiler_size;
  init(ft);
  CHECK(size_checks >= 16 && size_mismatches == 0);
  ft->handle = h;
  ft->get_element_size = dummy_get_element_size;
  pb2_cuda_table_t ct;
  init_cuda(&ct);
  const pb2_class_info *ci = &ct.info;
  CHECK(ci->abi_version == PB2_ABI_VERSION);
  CHECK((int)ft->nodal_dim == ci->nodal_dim && (int)ft->lagr_dim == ci->nodal_dim);
  CHECK((int)(ft->numfields_C2 + ft->numfields_C1) == ci->n_fields && (int)(ft->numfields_C2 + ft->numfields_C1) == ci->nval);
  CHECK(ft->numfields_C2 == ft->numfields_C2_bulk && ft->numfields_C2 == ft->numfields_C2_basebulk && ft->numfields_C2 == ft->numfields_C2_new);
  CHECK(ft->nodal_offset_C2_basebulk == 0 && ft->nodal_offset_C1_basebulk == ft->numfields_C2 && ft->buffer_offset_C1_basebulk == ft->numfields_C2);
  for (int i = 0; i < ci->n_fields; i++)
  {
    /* nodal value index -> (space, index in space): C2 fields first, then C1 (src/codegen.cpp:2795-2813) */
    const int idx = ci->field_index[i];
    const char *nm = ci->field_space[i] == 2 ? ft->fieldnames_C2[idx - (int)ft->nodal_offset_C2_basebulk] : ft->fieldnames_C1[idx - (int)ft->nodal_offset_C1_basebulk];
    CHECK(strcmp(nm, ci->field_names[i]) == 0);
  }
  CHECK((int)ft->numfields_Pos == 2 * ci->nodal_dim && strcmp(ft->fieldnames_Pos[0], "coordinate_x") == 0);
  CHECK((int)ft->num_res_jacs == ci->n_residuals && ft->current_res_jac == 0);
  for (int i = 0; i < ci->n_residuals; i++) CHECK(strcmp(ft->res_jac_names[i], ci->residual_names[i]) == 0);
  CHECK((int)ft->numglobal_params == ci->n_params);
  for (int i = 0; i < ci->n_params; i++) CHECK(ft->global_paramindices[i] == (unsigned)i && ft->global_parameters[i] == NULL);
  CHECK((ft->moving_nodes ? 1 : 0) == ci->moving_nodes && ft->max_dt_order == ci->max_dt_order && ft->integration_order == 0);
  CHECK((ft->hessian_generated ? 1 : 0) == ci->hessian_generated);
  CHECK((int)ft->numintegral_expressions == ci->n_integrals);
  for (int i = 0; i < ci->n_integrals; i++) CHECK(strcmp(ft->integral_expressions_names[i], ci->integral_names[i]) == 0);
  CHECK(strcmp(ft->dominant_space, "C2") == 0 && ft->domain_name && strcmp(ft->domain_name, ci->name) == 0);
  /* what the host dereferences right after init (src/problem.cpp:117-128) exists for every residual */
  for (unsigned i = 0; i < ft->num_res_jacs; i++)
  {
    CHECK(ft->shapes_required_ResJac[i].psi_Pos && ft->shapes_required_Hessian[i].psi_Pos);
    CHECK(ft->ResidualAndJacobian[i] && ft->ResidualAndJacobianSteady[i] && ft->ResidualAndJacobian_NoHang[i] && ft->HessianVectorProduct[i]);
    CHECK(ft->ParameterDerivative[i] != NULL);
  }
  CHECK(ft->clean_up != NULL && ct.query != NULL && ct.launch != NULL);
  printf("contract ok: %s dim %d, %u C2 + %u C1 fields, %u residuals, %u parameters, %d size checks\n", ci->name, ci->nodal_dim, ft->numfields_C2,
         ft->numfields_C1, ft->num_res_jacs, ft->numglobal_params, size_checks);

  if (argc >= 5)
  {
    /* one assembly through the engine's C-ABI, the engine library loaded like a host application would */
    void *eng = dlopen(argv[2], RTLD_NOW | RTLD_GLOBAL);
    if (!eng) { fprintf(stderr, "dlopen engine: %s\n", dlerror()); return 1; }
    int (*class_load)(const char *, pb2_class **) = (int (*)(const char *, pb2_class **))dlsym(eng, "pb2_class_load");
    int (*problem_create)(pb2_class *, int, const pb2_mesh_desc *, pb2_problem **) = (int (*)(pb2_class *, int, const pb2_mesh_desc *, pb2_problem **))dlsym(eng, "pb2_problem_create");
    int (*pattern)(pb2_problem *, const int **, const int **, long long *, long long *) = (int (*)(pb2_problem *, const int **, const int **, long long *, long long *))dlsym(eng, "pb2_problem_pattern");
    int (*set_pos)(pb2_problem *, int, const double *) = (int (*)(pb2_problem *, int, const double *))dlsym(eng, "pb2_problem_set_nodal_positions");
    int (*set_lagr)(pb2_problem *, const double *) = (int (*)(pb2_problem *, const double *))dlsym(eng, "pb2_problem_set_lagrangian_positions");
    int (*set_val)(pb2_problem *, int, const double *) = (int (*)(pb2_problem *, int, const double *))dlsym(eng, "pb2_problem_set_nodal_values");
    int (*assemble_host)(pb2_problem *, int, int, unsigned, const double *, double *, double *, double *) =
        (int (*)(pb2_problem *, int, int, unsigned, const double *, double *, double *, double *))dlsym(eng, "pb2_problem_assemble_host");
    const char *(*last_error)(void) = (const char *(*)(void))dlsym(eng, "pb2_last_error");
    void (*problem_free)(pb2_problem *) = (void (*)(pb2_problem *))dlsym(eng, "pb2_problem_free");
    void (*class_free)(pb2_class *) = (void (*)(pb2_class *))dlsym(eng, "pb2_class_free");
    CHECK(class_load && problem_create && pattern && set_pos && set_lagr && set_val && assemble_host && last_error && problem_free && class_free);
    FILE *f = fopen(argv[3], "rb");
    CHECK(f != NULL);
    long long hdr[4]; /* n_elem, n_node, n_dof, nnode */
    CHECK(fread(hdr, sizeof(long long), 4, f) == 4);
    const long long ne = hdr[0], nn = hdr[1], ndof = hdr[2], nnode = hdr[3];
    int *en = (int *)malloc(sizeof(int) * ne * nnode), *eq = (int *)malloc(sizeof(int) * nn * ci->nval);
    double *pos = (double *)malloc(sizeof(double) * nn * ci->nodal_dim), *val = (double *)malloc(sizeof(double) * nn * ci->nval);
    CHECK(fread(en, sizeof(int), ne * nnode, f) == (size_t)(ne * nnode) && fread(eq, sizeof(int), nn * ci->nval, f) == (size_t)(nn * ci->nval));
    CHECK(fread(pos, sizeof(double), nn * ci->nodal_dim, f) == (size_t)(nn * ci->nodal_dim) && fread(val, sizeof(double), nn * ci->nval, f) == (size_t)(nn * ci->nval));
    fclose(f);
    pb2_class *cls = NULL;
    pb2_problem *pr = NULL;
    if (class_load(argv[1], &cls)) { fprintf(stderr, "%s\n", last_error()); return 1; }
    pb2_mesh_desc md;
    memset(&md, 0, sizeof(md));
    md.n_elem = ne; md.n_node = nn; md.elem_nodes = en; md.node_eqn = eq; md.n_dof = ndof;
    if (problem_create(cls, 0, &md, &pr)) { fprintf(stderr, "%s\n", last_error()); return 1; }
    const int *rs, *cidx;
    long long nnz, nrows;
    CHECK(pattern(pr, &rs, &cidx, &nnz, &nrows) == 0 && nrows == ndof);
    CHECK(set_pos(pr, 0, pos) == 0 && set_lagr(pr, pos) == 0 && set_val(pr, 0, val) == 0);
    double *res = (double *)malloc(sizeof(double) * (ndof + 1)), *jac = (double *)malloc(sizeof(double) * (nnz + 1));
    if (assemble_host(pr, 0, -1, 1u, NULL, res, jac, NULL)) { fprintf(stderr, "%s\n", last_error()); return 1; }
    f = fopen(argv[4], "wb");
    CHECK(f != NULL);
    fwrite(&nnz, sizeof(long long), 1, f);
    fwrite(rs, sizeof(int), ndof + 1, f);
    fwrite(cidx, sizeof(int), nnz, f);
    fwrite(res, sizeof(double), ndof, f);
    fwrite(jac, sizeof(double), nnz, f);
    fclose(f);
    problem_free(pr);
    class_free(cls);
    printf("assembly ok: %lld elements, %lld dofs, %lld nnz\n", ne, ndof, nnz);
  }
  ft->clean_up(ft);
  CHECK(ft->fieldnames_C2 == NULL && ft->res_jac_names == NULL && ft->ResidualAndJacobian == NULL);
  free(ft);
  dlclose(h);
  return 0;
}
