// pb2_core.cu -- host runtime of the B200 element-assembly engine behind include/pyoomph_b200.h.
//
// What the reference does per assembly on the host (oomph-lib problem.cc:5332-5666: element loop, dense element
// matrices, linear search into per-row vectors, CSR conversion) is split here into a one-off setup
// (colouring, fixed CSR pattern, element->CSR position maps with first-touch flags, device upload) and a per-assembly
// sequence of kernel launches, one per colour, of the generated batched routine.  No CPU fallback exists: every
// error is reported, nothing is computed on the host.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <omp.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "pyoomph_b200.h"

#define PB2_DOF_NO_TARGET ((long long)0x8000000000000000ULL)
static thread_local std::string g_err;
static int fail(const std::string &m)
{
  g_err = m;
  return 1;
}
#define CUDA_OK(call)                                                                               \
  do                                                                                                \
  {                                                                                                 \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e_));         \
  } while (0)

#define NEED_DEVICE(p)                                                                                          \
  do                                                                                                            \
  {                                                                                                             \
    if ((p)->device < 0) return fail("pattern-only problem (created with device < 0): it holds no device data"); \
    CUDA_OK(cudaSetDevice((p)->device));                                                                        \
  } while (0)

// Large host arrays of the setup (CSR columns, position maps: hundreds of MB at the BASELINE sizes): std::vector would value-initialise
// and page-fault them on one thread; here the memory is first touched by the OpenMP threads that fill it.
template <class T>
struct RawVec
{
  T *d = nullptr;
  size_t n = 0;
  RawVec() {}
  RawVec(const RawVec &) = delete;
  RawVec &operator=(const RawVec &) = delete;
  ~RawVec() { free(d); }
  void resize_uninit(size_t m)
  {
    free(d);
    d = (T *)malloc(std::max<size_t>(1, m) * sizeof(T));
    n = d ? m : 0;
  }
  void assign(size_t m, T v)
  {
    resize_uninit(m);
    T *q = d;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < (long long)m; i++) q[i] = v;
  }
  void swap(RawVec &o)
  {
    std::swap(d, o.d);
    std::swap(n, o.n);
  }
  void clear()
  {
    free(d);
    d = nullptr;
    n = 0;
  }
  T *data() { return d; }
  const T *data() const { return d; }
  T *begin() { return d; }
  T *end() { return d + n; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T &operator[](size_t i) { return d[i]; }
  const T &operator[](size_t i) const { return d[i]; }
};

struct pb2_class
{
  void *handle = nullptr;
  pb2_cuda_table_t table;
};

struct pb2_problem
{
  pb2_class *cls = nullptr;
  int device = 0;
  long long n_elem = 0, n_node = 0, n_dof = 0, nnz = 0;
  int ndof_el = 0, nnode = 0, dim = 0, nval = 0, T_val = 1, T_pos = 1;
  std::vector<int> colour_begin; // [(nunit*ncol)+1] ranges of the (unit, colour) groups in the permuted element order
  std::vector<int> unit_tile;    // [nunit] tile (= patch colour) of each unit, units sorted by tile
  int n_colours = 0, n_tiles = 0;
  std::vector<int> perm;         // permuted position -> original element
  bool local_order = true;       // elements of a unit in patch order (chunks of a few elements, colour-sorted inside a chunk)
  std::vector<int> unit_begin;   // [nunit+1] element range of each unit in the permuted order
  std::vector<int> h_elem_nodes; // permuted element -> nodes (host copy, for the barrier masks of the batch tables)
  std::vector<int> row_start;
  RawVec<int> col_index;
  // device
  int *d_elem_nodes = nullptr, *d_elem_eqn = nullptr, *d_elem_rowstart = nullptr, *d_elem_res = nullptr;
  void *d_elem_off = nullptr;
  int map_bits = 8;
  long long *d_dof_target = nullptr;
  double *d_node_pos = nullptr, *d_node_lagr = nullptr, *d_node_val = nullptr;
  double *d_residual = nullptr, *d_jac = nullptr, *d_mass = nullptr, *d_dofs = nullptr;
  pb2_time_info ti;
  double params[PB2_MAX_PARAMS];
  long long launches_last = 0, launches_total = 0;
  int n_sms = 148;
  // batch tables of the pipelined kernels, per elements-per-batch value
  struct BatchTables
  {
    int epb = 0, grid = 0, n_batches = 0, n_tiles = 0;
    int *d_batch_elem = nullptr, *d_batch_meta = nullptr, *d_tile_nbatch = nullptr, *d_tile_done = nullptr, *d_block_begin = nullptr;
    unsigned long long *d_batch_bar = nullptr;
  };
  std::vector<BatchTables> batch_tables;
  cudaStream_t copy_stream = nullptr;
  unsigned long long *d_debug = nullptr;
  double *d_hvec = nullptr, *d_hess = nullptr, *d_hessM = nullptr;
  int hess_nvec = 0;
  int *d_row_start = nullptr, *d_col_index = nullptr;
  double *d_integrals = nullptr;   // [n_elem][n_integrals] per-element integral expressions, then [n_integrals] sums
  int *d_untouched = nullptr;      // CSR positions no local element writes (pattern entries owned for other ranks' contributions)
  long long n_untouched = 0;
  int *h_status = nullptr, *d_status = nullptr; // error word of the kernels: mapped pinned host memory, read without a copy
  cudaEvent_t ev_inputs = nullptr; // recorded on the legacy stream after every input update; assemblies on other streams wait for it
  double setup_seconds = 0.0;      // wall time of pb2_problem_create (colouring, pattern, maps, upload)
  // pattern-only problems (device < 0) keep the maps on the host for inspection (pb2_problem_host_maps)
  // hanging-node constraints (pb2_problem_set_constraints): reduction lists of P^T J P on the extended system
  long long c_ntarget = 0, c_nres = 0, c_nclear = 0, c_nvirt = 0;
  int *d_c_target = nullptr, *d_c_src_start = nullptr, *d_c_src_pos = nullptr, *d_c_res_row = nullptr, *d_c_res_start = nullptr,
      *d_c_res_src = nullptr, *d_c_clear = nullptr, *d_c_virt_rows = nullptr, *d_c_diag = nullptr;
  double *d_c_src_w = nullptr, *d_c_res_w = nullptr;
  int n_children = 0;              // child problems alive (they alias this problem's device buffers)
  pb2_problem *parent = nullptr;   // child problem: another element class scattering into the parent's matrix, residual and nodal data
  int *d_untouched_rows = nullptr; // residual rows no element of this problem writes (rows only a child class contributes to)
  long long n_untouched_rows = 0;
  RawVec<int> h_elem_rowstart, h_elem_res;
  RawVec<uint8_t> h_off8;
  RawVec<uint16_t> h_off16;
};

static int inputs_changed(pb2_problem *p);
static int check_status(pb2_problem *p);

extern "C" int pb2_version(void) { return PB2_ABI_VERSION; }
extern "C" const char *pb2_last_error(void) { return g_err.c_str(); }

extern "C" int pb2_class_load(const char *so_path, pb2_class **out)
{
  *out = nullptr;
  void *h = dlopen(so_path, RTLD_NOW | RTLD_LOCAL);
  if (!h) return fail(std::string("dlopen failed: ") + dlerror());
  auto init = (JIT_ELEMENT_init_cuda_SPEC)dlsym(h, "JIT_ELEMENT_init_cuda");
  if (!init)
  {
    dlclose(h);
    return fail("plugin does not export JIT_ELEMENT_init_cuda");
  }
  pb2_class *c = new pb2_class;
  c->handle = h;
  init(&c->table);
  if (c->table.info.abi_version != PB2_ABI_VERSION)
  {
    delete c;
    dlclose(h);
    return fail("plugin ABI version mismatch");
  }
  *out = c;
  return 0;
}

extern "C" int pb2_class_get_info(const pb2_class *cls, pb2_class_info *out)
{
  *out = cls->table.info;
  return 0;
}

extern "C" void pb2_class_free(pb2_class *cls)
{
  if (!cls) return;
  if (cls->handle) dlclose(cls->handle);
  delete cls;
}

template <class T>
static int upload(T **dptr, const std::vector<T> &v)
{
  CUDA_OK(cudaMalloc((void **)dptr, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) CUDA_OK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

// out = sorted union without duplicates of two sorted lists (out must hold na + nb entries); returns its length
static inline int merge_unique(const int *a, int na, const int *b, int nb, int *out)
{
  int i = 0, j = 0, n = 0;
  while (i < na && j < nb)
  {
    const int x = a[i], y = b[j];
    const int v = x < y ? x : y;
    i += x <= y;
    j += y <= x;
    if (n == 0 || out[n - 1] != v) out[n++] = v;
  }
  for (; i < na; i++)
    if (n == 0 || out[n - 1] != a[i]) out[n++] = a[i];
  for (; j < nb; j++)
    if (n == 0 || out[n - 1] != b[j]) out[n++] = b[j];
  return n;
}

template <class T>
static int upload(T **dptr, const RawVec<T> &v)
{
  CUDA_OK(cudaMalloc((void **)dptr, std::max<size_t>(1, v.size()) * sizeof(T)));
  if (!v.empty()) CUDA_OK(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

static int create_impl(pb2_class *cls, int device, const pb2_mesh_desc *m, pb2_problem *parent, pb2_problem **out);

extern "C" int pb2_problem_create(pb2_class *cls, int device, const pb2_mesh_desc *m, pb2_problem **out)
{
  return create_impl(cls, device, m, nullptr, out);
}

// A CHILD problem: elements of another class (interface elements on the bulk's boundary, src/elements.hpp:1435-2298) that live on the
// parent's nodes and equations and scatter into the parent's CSR matrix and residual (the reference assembles all element classes of a
// problem into one matrix, oomph-lib problem.cc:5332-5666).  The parent's pattern must contain every entry the child produces: create
// the parent with them as extra pattern entries (pb2_mesh_desc.extra_*).  The child owns its schedule and position maps only; it
// always ADDS (no first-touch stores): assemble the parent first, then its children, on the same stream.
extern "C" int pb2_problem_create_child(pb2_class *cls, pb2_problem *parent, const pb2_mesh_desc *m, pb2_problem **out)
{
  *out = nullptr;
  if (!parent) return fail("child problem needs a parent");
  if (parent->parent) return fail("the parent of a child problem must be a root problem");
  if (m->n_node != parent->n_node || m->n_dof != parent->n_dof) return fail("a child problem lives on the parent's nodes and equations: n_node / n_dof differ");
  if (cls->table.info.nval != parent->nval || cls->table.info.nodal_dim != parent->dim) return fail("child and parent element classes must share the nodal record (fields, dimension)");
  if (cls->table.info.n_hist_val > parent->T_val || cls->table.info.n_hist_pos > parent->T_pos) return fail("the child class reads more history levels than the parent stores");
  if (m->n_extra != 0) return fail("a child problem has no pattern of its own");
  return create_impl(cls, parent->device, m, parent, out);
}

static int create_impl(pb2_class *cls, int device, const pb2_mesh_desc *m, pb2_problem *parent, pb2_problem **out)
{
  *out = nullptr;
  const pb2_class_info &ci = cls->table.info;
  // device < 0: PATTERN-ONLY problem -- colouring, schedule, CSR pattern and position maps are built on the host and nothing touches
  // CUDA (for hosts that need row_start / column_index before a device is chosen, and for testing the pattern without a GPU);
  // every entry point that would compute refuses such a problem
  const bool dev = device >= 0;
  if (dev)
  {
    CUDA_OK(cudaSetDevice(device));
    CUDA_OK(cudaFree(0)); // the context exists from here on: its creation (~1 s, once per process) is not part of setup_seconds
  }
  const double t_create0 = omp_get_wtime();
  pb2_problem *p = new pb2_problem;
  if (dev) CUDA_OK(cudaDeviceGetAttribute(&p->n_sms, cudaDevAttrMultiProcessorCount, device));
  const bool setup_timing = getenv("PB2_SETUP_TIMING") != nullptr;
  double t_phase = omp_get_wtime();
  auto phase = [&](const char *what) {
    if (setup_timing)
    {
      const double now = omp_get_wtime();
      fprintf(stderr, "[pb2 setup] %-28s %.3f s\n", what, now - t_phase);
      t_phase = now;
    }
  };
  p->cls = cls;
  p->device = device;
  p->parent = parent;
  if (parent) parent->n_children++;
  p->n_elem = m->n_elem;
  p->n_node = m->n_node;
  p->n_dof = m->n_dof;
  p->ndof_el = ci.ndof_el;
  p->nnode = ci.nnode;
  p->dim = ci.nodal_dim;
  p->nval = ci.nval;
  p->T_val = ci.n_hist_val;
  p->T_pos = ci.n_hist_pos;
  memset(&p->ti, 0, sizeof(p->ti));
  memset(p->params, 0, sizeof(p->params));
  if (ci.moving_nodes && !m->pos_eqn)
  {
    pb2_problem_free(p); // releases whatever has been uploaded so far
    return fail("element class has position dofs but the mesh gives no pos_eqn");
  }
  if (m->n_elem * (long long)ci.ndof_el * ci.ndof_el > 0x7fffffffLL * 2)
  {
    // maps are indexed with long long on the device; only nnz must fit int32
  }
  const long long ne = m->n_elem;
  const int nn = ci.nnode, nd = ci.ndof_el;

  // ---- colouring: greedy over elements in mesh order, conflicts = shared nodes (all dofs live on nodes)
  std::vector<uint64_t> node_mask(m->n_node, 0);
  std::vector<int> colour(ne);
  int ncol = 0;
  for (long long e = 0; e < ne; e++)
  {
    uint64_t used = 0;
    for (int l = 0; l < nn; l++) used |= node_mask[m->elem_nodes[e * nn + l]];
    int c = 0;
    while (c < 64 && (used >> c & 1)) c++;
    if (c == 64)
    {
      pb2_problem_free(p); // releases whatever has been uploaded so far
      return fail("more than 64 colours needed");
    }
    colour[e] = c;
    ncol = std::max(ncol, c + 1);
    for (int l = 0; l < nn; l++) node_mask[m->elem_nodes[e * nn + l]] |= (uint64_t)1 << c;
  }
  std::vector<uint64_t>().swap(node_mask);
  phase("element colouring");
  // ---- schedule (DESIGN.md "Schedule"): elements are grouped into compact patches; patches are coloured (two patches
  // sharing a node get different colours) and a patch colour is a TILE: all patches of a tile are independent and
  // are processed concurrently by different thread blocks, each block running the element colours of its patches
  // back to back (so the CSR rows interior to a patch are completed while they sit in L2), and the tiles follow
  // each other behind a device-side gate (ncol_patch global synchronisations per assembly instead of one per colour
  // and chunk).  Several patches of a tile form a UNIT, the piece of work a block takes at a time.
  p->n_colours = ncol;
  std::vector<int> patch_of(ne);
  int npatch = 0;
  if (m->elem_patch)
  {
    for (long long e = 0; e < ne; e++)
    {
      patch_of[e] = m->elem_patch[e];
      if (patch_of[e] < 0)
      {
        pb2_problem_free(p); // releases whatever has been uploaded so far
        return fail("negative patch id");
      }
      npatch = std::max(npatch, patch_of[e] + 1);
    }
  }
  else
  {
    long long psize = 64;
    if (const char *cs = getenv("PB2_PATCH_ELEMS")) psize = std::max(1LL, atoll(cs));
    for (long long e = 0; e < ne; e++) patch_of[e] = (int)(e / psize);
    npatch = (int)((ne + psize - 1) / psize);
  }
  // patch colouring through the nodes: greedy in patch order
  std::vector<int> pcolour(npatch, -1);
  int npcol = 0;
  {
    // node -> patches (each node is touched by few patches)
    std::vector<int> np_start(m->n_node + 1, 0);
    std::vector<std::pair<int, int>> node_patch; // (node, patch) unique pairs
    node_patch.reserve((size_t)ne * nn / 4);
    {
      std::vector<int> last_patch_of_node(m->n_node, -1);
      // elements sorted by patch so that duplicates of (node, patch) are adjacent in time
      std::vector<int> order(ne);
      for (long long e = 0; e < ne; e++) order[e] = (int)e;
      std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return patch_of[x] < patch_of[y]; });
      for (long long i = 0; i < ne; i++)
      {
        const int e = order[i], pa = patch_of[e];
        for (int l = 0; l < nn; l++)
        {
          const int node = m->elem_nodes[(long long)e * nn + l];
          if (last_patch_of_node[node] != pa)
          {
            last_patch_of_node[node] = pa;
            node_patch.push_back({node, pa});
          }
        }
      }
    }
    for (auto &x : node_patch) np_start[x.first + 1]++;
    for (long long n = 0; n < m->n_node; n++) np_start[n + 1] += np_start[n];
    std::vector<int> np(node_patch.size());
    {
      std::vector<int> fp(np_start.begin(), np_start.end() - 1);
      for (auto &x : node_patch) np[fp[x.first]++] = x.second;
    }
    // patch -> nodes
    std::vector<int> pn_start(npatch + 1, 0);
    for (auto &x : node_patch) pn_start[x.second + 1]++;
    for (int q = 0; q < npatch; q++) pn_start[q + 1] += pn_start[q];
    std::vector<int> pn(node_patch.size());
    {
      std::vector<int> fp(pn_start.begin(), pn_start.end() - 1);
      for (auto &x : node_patch) pn[fp[x.second]++] = x.first;
    }
    for (int q = 0; q < npatch; q++)
    {
      uint64_t used = 0;
      for (int i = pn_start[q]; i < pn_start[q + 1]; i++)
      {
        const int node = pn[i];
        for (int j = np_start[node]; j < np_start[node + 1]; j++)
        {
          const int c = pcolour[np[j]];
          if (c >= 0) used |= (uint64_t)1 << c;
        }
      }
      int c = 0;
      while (c < 64 && (used >> c & 1)) c++;
      if (c == 64)
      {
        pb2_problem_free(p); // releases whatever has been uploaded so far
        return fail("more than 64 patch colours needed");
      }
      pcolour[q] = c;
      npcol = std::max(npcol, c + 1);
    }
  }
  phase("patch colouring");
  // units: consecutive patches of one tile.  A tile ends at a device-wide gate, so its units should split evenly over the persistent
  // grid (one block per SM): among 4, 2, 1 patches per unit take the one with the least idle time ceil(U / grid) * grid / U at the gates,
  // larger units preferred unless a smaller one saves more than 3 % (measured on one B200, profiles/r02_notes.md: 262 k elements 4 -> 1
  // patches per unit 0.792 -> 0.700 ms, 524 k elements 1.562 -> 1.391 ms; the 1 M-element mesh stays at 4)
  int unit_patches = 4;
  {
    std::vector<long long> per_tile(std::max(1, npcol), 0);
    for (int q = 0; q < npatch; q++) per_tile[pcolour[q]]++;
    const double grid_blocks = (double)std::max(1, p->n_sms);
    double best = 1e300;
    for (int cand : {4, 2, 1})
    {
      double idle = 0.0;
      for (int t = 0; t < npcol; t++)
      {
        const double U = (double)((per_tile[t] + cand - 1) / cand);
        if (U > 0) idle += std::ceil(U / grid_blocks) * grid_blocks / U;
      }
      if (idle < best * 0.97)
      {
        best = idle;
        unit_patches = cand;
      }
    }
  }
  if (const char *cs = getenv("PB2_UNIT_PATCHES")) unit_patches = std::max(1, atoi(cs));
  std::vector<int> unit_of_patch(npatch), unit_tile;
  {
    std::vector<int> cnt(npcol, 0), cur(npcol, -1);
    for (int q = 0; q < npatch; q++)
    {
      const int t = pcolour[q];
      if (cur[t] < 0 || cnt[t] == unit_patches)
      {
        cur[t] = (int)unit_tile.size();
        unit_tile.push_back(t);
        cnt[t] = 0;
      }
      unit_of_patch[q] = cur[t];
      cnt[t]++;
    }
  }
  // order units by tile, elements by (unit, colour, patch, mesh order)
  const int nunit = (int)unit_tile.size();
  std::vector<int> unit_rank(nunit);
  {
    std::vector<int> ord(nunit);
    for (int u = 0; u < nunit; u++) ord[u] = u;
    std::stable_sort(ord.begin(), ord.end(), [&](int x, int y) { return unit_tile[x] < unit_tile[y]; });
    for (int r = 0; r < nunit; r++) unit_rank[ord[r]] = r;
    p->unit_tile.assign(nunit, 0);
    for (int u = 0; u < nunit; u++) p->unit_tile[unit_rank[u]] = unit_tile[u];
  }
  p->n_tiles = npcol;
  p->perm.resize(ne);
  for (long long e = 0; e < ne; e++) p->perm[e] = (int)e;
  if (const char *cs = getenv("PB2_ORDER")) p->local_order = strcmp(cs, "colour") != 0;
  if (p->local_order)
  {
    // LOCAL order (default): inside a unit the elements follow patch after patch; a patch is cut into chunks of a few
    // consecutive elements (mesh order) and only inside a chunk the elements are sorted by colour.  Elements that share
    // CSR entries are then scattered within a few microseconds of each other by the SAME thread block (in program order,
    // separated by block barriers where they conflict, see batch_bar), so the entry is still in L2 when it is completed.
    int chunk = 16;
    if (const char *cs = getenv("PB2_CHUNK")) chunk = std::max(1, atoi(cs));
    std::vector<int> pos_in_patch(ne), cnt(npatch, 0);
    for (long long e = 0; e < ne; e++) pos_in_patch[e] = cnt[patch_of[e]]++;
    std::stable_sort(p->perm.begin(), p->perm.end(), [&](int x, int y) {
      const int ux = unit_rank[unit_of_patch[patch_of[x]]], uy = unit_rank[unit_of_patch[patch_of[y]]];
      if (ux != uy) return ux < uy;
      if (patch_of[x] != patch_of[y]) return patch_of[x] < patch_of[y];
      const int cx = pos_in_patch[x] / chunk, cy = pos_in_patch[y] / chunk;
      if (cx != cy) return cx < cy;
      return colour[x] < colour[y];
    });
  }
  else
    std::stable_sort(p->perm.begin(), p->perm.end(), [&](int x, int y) {
      const int ux = unit_rank[unit_of_patch[patch_of[x]]], uy = unit_rank[unit_of_patch[patch_of[y]]];
      if (ux != uy) return ux < uy;
      if (colour[x] != colour[y]) return colour[x] < colour[y];
      return patch_of[x] < patch_of[y];
    });
  // colour_begin: ranges of (unit, colour) groups in the permuted order (colour-major order only)
  p->colour_begin.assign((size_t)nunit * ncol + 1, 0);
  for (long long e = 0; e < ne; e++) p->colour_begin[(size_t)unit_rank[unit_of_patch[patch_of[e]]] * ncol + colour[e] + 1]++;
  for (size_t c = 0; c + 1 < p->colour_begin.size(); c++) p->colour_begin[c + 1] += p->colour_begin[c];
  p->unit_begin.assign((size_t)nunit + 1, 0);
  for (long long e = 0; e < ne; e++) p->unit_begin[(size_t)unit_rank[unit_of_patch[patch_of[e]]] + 1]++;
  for (int u = 0; u < nunit; u++) p->unit_begin[u + 1] += p->unit_begin[u];

  phase("schedule order");
  // ---- permuted element tables
  std::vector<int> elem_nodes((size_t)ne * nn), elem_eqn((size_t)ne * nd);
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < ne; q++)
  {
    const long long e = p->perm[q];
    for (int l = 0; l < nn; l++) elem_nodes[q * nn + l] = m->elem_nodes[e * nn + l];
    for (int k = 0; k < nd; k++)
    {
      const long long node = m->elem_nodes[e * nn + ci.dof_node[k]];
      elem_eqn[q * nd + k] = ci.dof_kind[k] == 0 ? m->pos_eqn[node * ci.nodal_dim + ci.dof_index[k]]
                                                 : m->node_eqn[node * ci.nval + ci.dof_index[k]];
    }
  }

  p->h_elem_nodes = elem_nodes;

  phase("element tables");
  // ---- dof -> elements adjacency (permuted element ids)
  const long long nrow = m->n_dof;
  std::vector<int> adj_start(nrow + 1, 0);
  for (size_t i = 0; i < elem_eqn.size(); i++)
    if (elem_eqn[i] >= 0) adj_start[elem_eqn[i] + 1]++;
  for (long long r = 0; r < nrow; r++) adj_start[r + 1] += adj_start[r];
  std::vector<int> adj(adj_start[nrow]);
  {
    std::vector<int> fp(adj_start.begin(), adj_start.end() - 1);
    for (long long q = 0; q < ne; q++)
      for (int k = 0; k < nd; k++)
      {
        const int g = elem_eqn[q * nd + k];
        if (g >= 0) adj[fp[g]++] = (int)q;
      }
  }

  // extra pattern entries (interface columns contributed by other ranks), bucketed by row
  std::vector<int> ex_start(nrow + 1, 0), ex_col(std::max<long long>(0, m->n_extra));
  for (long long i = 0; i < m->n_extra; i++)
  {
    if (m->extra_rows[i] < 0 || m->extra_rows[i] >= nrow || m->extra_cols[i] < 0 || m->extra_cols[i] >= nrow)
    {
      pb2_problem_free(p); // releases whatever has been uploaded so far
      return fail("extra pattern entry out of range");
    }
    ex_start[m->extra_rows[i] + 1]++;
  }
  for (long long r = 0; r < nrow; r++) ex_start[r + 1] += ex_start[r];
  {
    std::vector<int> fp(ex_start.begin(), ex_start.end() - 1);
    for (long long i = 0; i < m->n_extra; i++) ex_col[fp[m->extra_rows[i]]++] = m->extra_cols[i];
  }
  phase("adjacency");
  // dofs of every element sorted by equation, with their local index
  RawVec<int> sorted_eq, sorted_k;
  sorted_eq.resize_uninit((size_t)ne * nd);
  sorted_k.resize_uninit((size_t)ne * nd);
  std::vector<int> n_sorted(ne);
#pragma omp parallel
  {
    std::vector<std::pair<int, int>> srt(nd);
#pragma omp for schedule(static)
    for (long long q = 0; q < ne; q++)
    {
      const int *eq = &elem_eqn[(size_t)q * nd];
      int ns = 0;
      for (int k = 0; k < nd; k++)
        if (eq[k] >= 0) srt[ns++] = {eq[k], k};
      std::sort(srt.begin(), srt.begin() + ns);
      n_sorted[q] = ns;
      for (int s_ = 0; s_ < ns; s_++)
      {
        sorted_eq[(size_t)q * nd + s_] = srt[s_].first;
        sorted_k[(size_t)q * nd + s_] = srt[s_].second;
      }
    }
  }
  phase("sorted element dof lists");
  // ---- CSR pattern, ascending columns.  One pass: every thread takes a contiguous block of rows, sorts the dofs of the row's elements
  // once and keeps the columns in a private buffer; rows whose element list equals the previous row's (the dofs of one node) reuse its
  // columns; then prefix sum and parallel copy.
  p->row_start.assign(nrow + 1, 0);
  if (parent)
  {
    p->row_start = parent->row_start; // the child scatters into the parent's matrix
    p->nnz = parent->nnz;
  }
  else
  {
    const int nth = omp_get_max_threads();
    std::vector<std::vector<int>> tcols(nth);
    std::vector<long long> tbeg(nth + 1, 0);
#pragma omp parallel num_threads(nth)
    {
      const int t = omp_get_thread_num();
      const long long r0 = nrow * t / nth, r1 = nrow * (t + 1) / nth;
      std::vector<int> &buf = tcols[t];
      buf.reserve((size_t)((double)(r1 - r0) * ((double)ne * nd * nd / std::max<long long>(1, nrow)) * 0.3) + 1024);   // ~ nnz per row block, over-reserved
      std::vector<int> tmp(256), tmp2(256);
      size_t prev_off = 0;
      int prev_n = -1;
      for (long long r = r0; r < r1; r++)
      {
        const int na = adj_start[r + 1] - adj_start[r];
        const bool same = r > r0 && prev_n >= 0 && ex_start[r + 1] == ex_start[r] && ex_start[r] == ex_start[r - 1] &&
                          na == adj_start[r] - adj_start[r - 1] && std::equal(adj.begin() + adj_start[r], adj.begin() + adj_start[r + 1], adj.begin() + adj_start[r - 1]);
        if (same)
        {
          const size_t o = buf.size();
          buf.resize(o + prev_n);
          std::copy(buf.begin() + prev_off, buf.begin() + prev_off + prev_n, buf.begin() + o);
          prev_off = o;
        }
        else
        {
          // the dof lists of the row's elements are sorted already: merge them list by list (dropping duplicates) instead of sorting
          // their union
          int n_acc = 0;
          for (int a_ = adj_start[r]; a_ < adj_start[r + 1]; a_++)
          {
            const long long q = adj[a_];
            const int nb = n_sorted[q];
            if ((int)tmp.size() < n_acc + nb) { tmp.resize(2 * (n_acc + nb) + 64); tmp2.resize(tmp.size()); }
            n_acc = merge_unique(tmp.data(), n_acc, sorted_eq.data() + (size_t)q * nd, nb, tmp2.data());
            tmp.swap(tmp2);
          }
          if (ex_start[r + 1] > ex_start[r])
          {
            std::vector<int> ex(ex_col.begin() + ex_start[r], ex_col.begin() + ex_start[r + 1]);
            std::sort(ex.begin(), ex.end());
            if ((int)tmp.size() < n_acc + (int)ex.size()) { tmp.resize(2 * (n_acc + ex.size()) + 64); tmp2.resize(tmp.size()); }
            n_acc = merge_unique(tmp.data(), n_acc, ex.data(), (int)ex.size(), tmp2.data());
            tmp.swap(tmp2);
          }
          prev_n = n_acc;
          prev_off = buf.size();
          buf.insert(buf.end(), tmp.begin(), tmp.begin() + prev_n);
        }
        p->row_start[r + 1] = prev_n;
      }
    }
    phase("  pattern: row sorts");
    long long tot = 0;
    for (int t = 0; t < nth; t++)
    {
      tbeg[t] = tot;
      tot += (long long)tcols[t].size();
    }
    if (tot >= 0x7fffffffLL)
    {
      pb2_problem_free(p); // releases whatever has been uploaded so far
      return fail("nnz exceeds int32 CSR indexing");
    }
    for (long long r = 0; r < nrow; r++) p->row_start[r + 1] += p->row_start[r];
    p->nnz = tot;
    p->col_index.resize_uninit(p->nnz);
    phase("  pattern: allocate");
#pragma omp parallel num_threads(nth)
    {
      const int t = omp_get_thread_num();
      std::copy(tcols[t].begin(), tcols[t].end(), p->col_index.data() + tbeg[t]);
      std::vector<int>().swap(tcols[t]);
    }
  }
  phase("CSR pattern");

  // ---- element -> CSR position maps, first-touch flags and their compressed form, in ONE pass over the matrix rows.
  // Per local row of an element: its CSR row start (int32) and per (row, col) the offset inside the row plus a first-touch bit, 8 bits
  // if every row is shorter than 127 entries, else 16 (4*ndof^2 -> ndof^2 bytes).  First touch = the element with the smallest
  // scheduled index q among those reaching an entry stores, later ones add: no zero-fill of the outputs, fixed sum order.
  // A row walks its elements in ascending q (adj is sorted), merges each element's sorted dof list with the row's columns
  // (both ascending: one linear walk) and writes that element's slice of the map; rows with the element list of the previous row
  // (dofs of one node) copy the previous row's slices.
  int maxlen = 0;
  for (long long r = 0; r < nrow; r++) maxlen = std::max(maxlen, p->row_start[r + 1] - p->row_start[r]);
  p->map_bits = maxlen < 127 ? 8 : 16;
  if (maxlen >= 32767)
  {
    pb2_problem_free(p); // releases whatever has been uploaded so far
    return fail("CSR rows longer than 32766 entries are not supported by the position map");
  }
  RawVec<int> elem_rowstart, elem_res;
  elem_rowstart.assign((size_t)ne * nd, -1);
  elem_res.assign((size_t)ne * nd, PB2_MAP_SKIP);
  RawVec<uint8_t> off8;
  RawVec<uint16_t> off16;
  if (p->map_bits == 8)
    off8.assign((size_t)ne * nd * nd + 16, (uint8_t)0xFF);   // + slack: the kernels prefetch the 4-byte words that COVER a batch's slice
  else
    off16.assign((size_t)ne * nd * nd + 8, (uint16_t)0xFFFF);
  const unsigned FIRST = p->map_bits == 8 ? 0x80u : 0x8000u;
  std::vector<std::vector<int>> t_untouched(omp_get_max_threads()), t_untouched_rows(omp_get_max_threads());
  const int *const col_base = parent ? parent->col_index.data() : p->col_index.data();
  const bool child = parent != nullptr;
  int missing_entries = 0;
#pragma omp parallel
  {
    std::vector<int> firstq(maxlen + 1), loc_k(nd), loc_off(nd), prev_row_k;
    std::vector<int> &unt = t_untouched[omp_get_thread_num()];
    std::vector<int> &untr = t_untouched_rows[omp_get_thread_num()];
    prev_row_k.reserve(64);
#pragma omp for schedule(static, 4096)
    for (long long r = 0; r < nrow; r++)
    {
      const int rb = p->row_start[r], len = p->row_start[r + 1] - rb;
      const int *cols = col_base + rb;
      const int a0 = adj_start[r], a1 = adj_start[r + 1];
      if (a0 == a1)
      {
        if (m->n_extra > 0) untr.push_back((int)r); // a row only other classes / ranks contribute to: its residual restarts from zero
        if (m->n_extra > 0)
          for (int i = 0; i < len; i++) unt.push_back(rb + i);
        continue;
      }
      // dofs of one node: same elements and same columns as the previous row => the same offsets and first touches; only the row
      // start differs.  (Not across the first row of a thread's block, and not when extra pattern entries are involved.)
      const bool same = !child && r > 0 && (r % 4096) != 0 && m->n_extra == 0 && len == p->row_start[r] - p->row_start[r - 1] && a1 - a0 == a0 - adj_start[r - 1] &&
                        std::equal(adj.begin() + a0, adj.begin() + a1, adj.begin() + adj_start[r - 1]);
      if (same)
      {
        for (int a_ = a0; a_ < a1; a_++)
        {
          const long long q = adj[a_];
          const int *se = &sorted_eq[(size_t)q * nd], *sk = &sorted_k[(size_t)q * nd];
          const int ns = n_sorted[q];
          const int s_ = (int)(std::lower_bound(se, se + ns, (int)r) - se); // r is a dof of q
          const int ir = sk[s_], ip = sk[s_ - 1];                            // ... and r - 1 the one before it in q's sorted list
          const size_t base = ((size_t)q * nd + ir) * nd, basep = ((size_t)q * nd + ip) * nd;
          if (p->map_bits == 8)
            memcpy(&off8[base], &off8[basep], nd);
          else
            memcpy(&off16[base], &off16[basep], nd * sizeof(uint16_t));
          elem_rowstart[(size_t)q * nd + ir] = rb;
          elem_res[(size_t)q * nd + ir] = a_ == a0 ? ~(int)r : (int)r;
        }
        continue;
      }
      for (int i = 0; i < len; i++) firstq[i] = -1;
      for (int a_ = a0; a_ < a1; a_++)
      {
        const long long q = adj[a_];
        const int *se = &sorted_eq[(size_t)q * nd], *sk = &sorted_k[(size_t)q * nd];
        const int ns = n_sorted[q];
        int ir = -1, pos = 0;
        for (int s_ = 0; s_ < ns; s_++)
        {
          if (se[s_] == (int)r) ir = sk[s_];
          while (pos < len && cols[pos] < se[s_]) pos++; // every dof of an adjacent element is a column of this row
          if (pos >= len || cols[pos] != se[s_])
          {
            // only possible for a child: the parent's pattern lacks an entry this class produces
#pragma omp atomic
            missing_entries++;
            pos = 0;
            loc_k[s_] = -1;
            continue;
          }
          loc_k[s_] = sk[s_];
          loc_off[s_] = pos;
        }
        // ir >= 0: r is a dof of q (that is what adjacency means)
        const size_t base = ((size_t)q * nd + ir) * nd;
        for (int s_ = 0; s_ < ns; s_++)
        {
          if (loc_k[s_] < 0) continue;
          const int off = loc_off[s_];
          unsigned code = (unsigned)off;
          if (firstq[off] < 0)
          {
            firstq[off] = (int)q;
            if (!child) code |= FIRST; // a child adds to what the parent (and earlier children) wrote
          }
          if (p->map_bits == 8)
            off8[base + loc_k[s_]] = (uint8_t)code;
          else
            off16[base + loc_k[s_]] = (uint16_t)code;
        }
        elem_rowstart[(size_t)q * nd + ir] = rb;
        elem_res[(size_t)q * nd + ir] = (a_ == a0 && !child) ? ~(int)r : (int)r;
      }
      if (m->n_extra > 0)
        for (int i = 0; i < len; i++)
          if (firstq[i] < 0) unt.push_back(rb + i);
    }
  }
  sorted_eq.clear();
  sorted_k.clear();
  std::vector<int>().swap(adj);
  std::vector<int>().swap(adj_start);
  if (missing_entries > 0)
  {
    pb2_problem_free(p);
    return fail("the parent's CSR pattern lacks " + std::to_string(missing_entries) + " entries this element class produces: create the parent with them as extra pattern entries");
  }
  if (m->n_extra > 0)
  {
    std::vector<int> untr;
    for (auto &v : t_untouched_rows) untr.insert(untr.end(), v.begin(), v.end());
    std::sort(untr.begin(), untr.end());
    p->n_untouched_rows = (long long)untr.size();
    if (dev && upload(&p->d_untouched_rows, untr))
    {
      pb2_problem_free(p); // releases whatever has been uploaded so far
      return 1;
    }
  }
  if (m->n_extra > 0)
  {
    // CSR positions no local element writes (extra pattern entries): re-zeroed before every assembly
    std::vector<int> unt;
    for (auto &v : t_untouched) unt.insert(unt.end(), v.begin(), v.end());
    std::sort(unt.begin(), unt.end());
    p->n_untouched = (long long)unt.size();
    if (dev && upload(&p->d_untouched, unt))
    {
      pb2_problem_free(p); // releases whatever has been uploaded so far
      return 1;
    }
  }
  phase("position maps + first touch");

  // ---- dof -> nodal storage target for set_dofs: >=0 index into node_val (t=0), <0: ~index into node_pos (t=0)
  std::vector<long long> dof_target(nrow, PB2_DOF_NO_TARGET);
  for (long long n = 0; n < m->n_node; n++)
  {
    for (int f = 0; f < ci.nval; f++)
    {
      const int g = m->node_eqn[n * ci.nval + f];
      if (g >= 0) dof_target[g] = n * ci.nval + f;
    }
    if (m->pos_eqn)
      for (int d = 0; d < ci.nodal_dim; d++)
      {
        const int g = m->pos_eqn[n * ci.nodal_dim + d];
        if (g >= 0) dof_target[g] = ~(n * ci.nodal_dim + d);
      }
  }

  phase("dof targets");
  p->setup_seconds = omp_get_wtime() - t_create0;
  if (!dev)
  {
    p->h_elem_rowstart.swap(elem_rowstart);
    p->h_elem_res.swap(elem_res);
    p->h_off8.swap(off8);
    p->h_off16.swap(off16);
    *out = p;
    return 0;
  }
  // ---- upload
  if (upload(&p->d_elem_nodes, elem_nodes) || upload(&p->d_elem_eqn, elem_eqn) || upload(&p->d_elem_rowstart, elem_rowstart) ||
      upload(&p->d_elem_res, elem_res) || upload(&p->d_dof_target, dof_target))
  {
    pb2_problem_free(p);
    return 1;
  }
  if (p->map_bits == 8 ? upload((uint8_t **)&p->d_elem_off, off8) : upload((uint16_t **)&p->d_elem_off, off16))
  {
    pb2_problem_free(p);
    return 1;
  }
  if (parent)
  {
    // nodal data, dof vector and outputs are the parent's (history strides included: the child reads the parent's levels)
    p->T_val = parent->T_val;
    p->T_pos = parent->T_pos;
    p->d_node_pos = parent->d_node_pos;
    p->d_node_lagr = parent->d_node_lagr;
    p->d_node_val = parent->d_node_val;
    p->d_residual = parent->d_residual;
    p->d_dofs = parent->d_dofs;
    p->d_jac = parent->d_jac;
    CUDA_OK(cudaHostAlloc((void **)&p->h_status, sizeof(int), cudaHostAllocMapped));
    *p->h_status = 0;
    CUDA_OK(cudaHostGetDevicePointer((void **)&p->d_status, p->h_status, 0));
    CUDA_OK(cudaEventCreateWithFlags(&p->ev_inputs, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(p->ev_inputs, 0));
    p->setup_seconds = omp_get_wtime() - t_create0;
    *out = p;
    return 0;
  }
  const size_t npos = (size_t)p->T_pos * m->n_node * ci.nodal_dim, nlag = (size_t)m->n_node * ci.nodal_dim,
               nvals = (size_t)p->T_val * m->n_node * std::max(1, ci.nval);
  CUDA_OK(cudaMalloc((void **)&p->d_node_pos, npos * sizeof(double)));
  CUDA_OK(cudaMalloc((void **)&p->d_node_lagr, nlag * sizeof(double)));
  CUDA_OK(cudaMalloc((void **)&p->d_node_val, nvals * sizeof(double)));
  CUDA_OK(cudaMemset(p->d_node_pos, 0, npos * sizeof(double)));
  CUDA_OK(cudaMemset(p->d_node_lagr, 0, nlag * sizeof(double)));
  CUDA_OK(cudaMemset(p->d_node_val, 0, nvals * sizeof(double)));
  CUDA_OK(cudaMalloc((void **)&p->d_residual, std::max<size_t>(1, nrow) * sizeof(double)));
  CUDA_OK(cudaMalloc((void **)&p->d_dofs, std::max<size_t>(1, nrow) * sizeof(double)));
  CUDA_OK(cudaMalloc((void **)&p->d_jac, std::max<size_t>(1, p->nnz) * sizeof(double)));
  // rows / entries no local element touches (ghost rows, element subsets) read as zero, never as uninitialised memory
  CUDA_OK(cudaMemset(p->d_residual, 0, std::max<size_t>(1, nrow) * sizeof(double)));
  CUDA_OK(cudaMemset(p->d_dofs, 0, std::max<size_t>(1, nrow) * sizeof(double)));
  CUDA_OK(cudaMemset(p->d_jac, 0, std::max<size_t>(1, p->nnz) * sizeof(double)));
  CUDA_OK(cudaHostAlloc((void **)&p->h_status, sizeof(int), cudaHostAllocMapped));
  *p->h_status = 0;
  CUDA_OK(cudaHostGetDevicePointer((void **)&p->d_status, p->h_status, 0));
  CUDA_OK(cudaEventCreateWithFlags(&p->ev_inputs, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(p->ev_inputs, 0));
  phase("upload + device buffers");
  p->setup_seconds = omp_get_wtime() - t_create0;
  *out = p;
  return 0;
}

extern "C" void pb2_problem_free(pb2_problem *p)
{
  if (!p) return;
  if (p->n_children > 0)
  {
    // child problems alias this problem's device buffers: releasing it now would leave them dangling.  Keep it (the caller releases the
    // children first and calls again); pb2_last_error says why nothing happened.
    fail("pb2_problem_free: " + std::to_string(p->n_children) + " child problem(s) still alive; release them first");
    return;
  }
  if (p->parent) p->parent->n_children--; // children are released before their parent
  if (p->device < 0)
  {
    delete p;
    return;
  }
  cudaSetDevice(p->device);
  if (p->parent)
  {
    // aliases of the parent's buffers are not this problem's to release
    p->d_node_pos = p->d_node_lagr = p->d_node_val = p->d_residual = p->d_jac = p->d_dofs = nullptr;
  }
  cudaFree(p->d_untouched_rows);
  cudaFree(p->d_c_target); cudaFree(p->d_c_src_start); cudaFree(p->d_c_src_pos); cudaFree(p->d_c_src_w);
  cudaFree(p->d_c_res_row); cudaFree(p->d_c_res_start); cudaFree(p->d_c_res_src); cudaFree(p->d_c_res_w);
  cudaFree(p->d_c_clear); cudaFree(p->d_c_virt_rows); cudaFree(p->d_c_diag);
  cudaFree(p->d_elem_nodes);
  cudaFree(p->d_elem_eqn);
  cudaFree(p->d_elem_rowstart);
  cudaFree(p->d_elem_off);
  cudaFree(p->d_integrals);
  cudaFree(p->d_elem_res);
  cudaFree(p->d_dof_target);
  cudaFree(p->d_node_pos);
  cudaFree(p->d_node_lagr);
  cudaFree(p->d_node_val);
  cudaFree(p->d_residual);
  cudaFree(p->d_jac);
  cudaFree(p->d_mass);
  cudaFree(p->d_dofs);
  cudaFree(p->d_untouched);
  cudaFree(p->d_hvec);
  cudaFree(p->d_hess);
  cudaFree(p->d_hessM);
  cudaFree(p->d_row_start);
  cudaFree(p->d_col_index);
  cudaFree(p->d_debug);
  if (p->h_status) cudaFreeHost(p->h_status);
  if (p->ev_inputs) cudaEventDestroy(p->ev_inputs);
  for (auto &b : p->batch_tables)
  {
    cudaFree(b.d_batch_elem);
    cudaFree(b.d_batch_bar);
    cudaFree(b.d_batch_meta);
    cudaFree(b.d_tile_nbatch);
    cudaFree(b.d_tile_done);
    cudaFree(b.d_block_begin);
  }
  delete p;
}

extern "C" int pb2_problem_pattern(pb2_problem *p, const int **row_start, const int **column_index, long long *nnz, long long *n_rows)
{
  if (row_start) *row_start = p->row_start.data();
  if (column_index) *column_index = p->parent ? p->parent->col_index.data() : p->col_index.data();
  if (nnz) *nnz = p->nnz;
  if (n_rows) *n_rows = p->n_dof;
  return 0;
}

extern "C" int pb2_problem_host_maps(pb2_problem *p, const int **perm, const int **elem_rowstart, const void **elem_off, int *map_bits, const int **elem_res)
{
  if (p->device >= 0) return fail("the position maps of a device problem live in HBM; create a pattern-only problem (device < 0) to inspect them");
  if (perm) *perm = p->perm.data();
  if (elem_rowstart) *elem_rowstart = p->h_elem_rowstart.data();
  if (elem_off) *elem_off = p->map_bits == 8 ? (const void *)p->h_off8.data() : (const void *)p->h_off16.data();
  if (map_bits) *map_bits = p->map_bits;
  if (elem_res) *elem_res = p->h_elem_res.data();
  return 0;
}

extern "C" int pb2_problem_num_colours(pb2_problem *p) { return p->n_colours; }
extern "C" int pb2_problem_num_launches(pb2_problem *p) { return p->n_tiles; }

extern "C" int pb2_problem_set_nodal_values(pb2_problem *p, int t, const double *values)
{
  if (t < 0 || t >= p->T_val) return fail("history index out of range");
  NEED_DEVICE(p);
  const size_t n = (size_t)p->n_node * p->nval;
  CUDA_OK(cudaMemcpy(p->d_node_val + (size_t)t * n, values, n * sizeof(double), cudaMemcpyHostToDevice));
  return inputs_changed(p);
}

extern "C" int pb2_problem_set_nodal_positions(pb2_problem *p, int t, const double *pos)
{
  if (t < 0 || t >= p->T_pos) return fail("position history index out of range");
  NEED_DEVICE(p);
  const size_t n = (size_t)p->n_node * p->dim;
  CUDA_OK(cudaMemcpy(p->d_node_pos + (size_t)t * n, pos, n * sizeof(double), cudaMemcpyHostToDevice));
  return inputs_changed(p);
}

extern "C" int pb2_problem_set_lagrangian_positions(pb2_problem *p, const double *pos)
{
  NEED_DEVICE(p);
  CUDA_OK(cudaMemcpy(p->d_node_lagr, pos, (size_t)p->n_node * p->dim * sizeof(double), cudaMemcpyHostToDevice));
  return inputs_changed(p);
}

// every update of the packed inputs runs on the legacy default stream; an assembly on another (possibly non-blocking) stream is
// ordered behind it through this event
static int inputs_changed(pb2_problem *p)
{
  CUDA_OK(cudaEventRecord(p->ev_inputs, 0));
  return 0;
}

static __global__ void pb2_zero_positions(double *__restrict__ a, double *__restrict__ b, const int *__restrict__ pos, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  a[pos[i]] = 0.0;
  if (b) b[pos[i]] = 0.0;
}

static __global__ void pb2_scatter_dofs(const double *__restrict__ dofs, const long long *__restrict__ target, long long n,
                                        double *__restrict__ node_val, double *__restrict__ node_pos)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long t = target[i];
  if (t == PB2_DOF_NO_TARGET) return; // a column-only (ghost) dof of a row-block partition: no nodal storage on this GPU
  if (t >= 0)
    node_val[t] = dofs[i];
  else
    node_pos[~t] = dofs[i];
}

static int scatter_dofs_from_device(pb2_problem *p, cudaStream_t s, int t = 0)
{
  const int bs = 256;
  const long long nb = (p->n_dof + bs - 1) / bs;
  if (nb > 0)
  {
    // history level t of the nodal values / positions (levels beyond what a buffer stores are not written: the pointer stays at level 0
    // only when that buffer has a single level and the class never reads its history)
    double *val = p->d_node_val + (size_t)std::min(t, p->T_val - 1) * p->n_node * std::max(1, p->nval);
    double *pos = p->d_node_pos + (size_t)std::min(t, p->T_pos - 1) * p->n_node * p->dim;
    pb2_scatter_dofs<<<(unsigned)nb, bs, 0, s>>>(p->d_dofs, p->d_dof_target, p->n_dof, val, pos);
    p->launches_last++;
    p->launches_total++;
    CUDA_OK(cudaGetLastError());
  }
  return inputs_changed(p);
}

extern "C" int pb2_problem_set_dofs(pb2_problem *p, const double *dofs)
{
  NEED_DEVICE(p);
  CUDA_OK(cudaMemcpy(p->d_dofs, dofs, (size_t)p->n_dof * sizeof(double), cudaMemcpyHostToDevice));
  return scatter_dofs_from_device(p, 0);
}

// Problem::set_history_dofs(t, ...) (src/pybind/problem.cpp:540): a dof vector of history level t -> nodal values / positions of that level
extern "C" int pb2_problem_set_history_dofs(pb2_problem *p, int t, const double *dofs)
{
  if (t < 0 || t >= std::max(p->T_val, p->T_pos)) return fail("history index out of range");
  NEED_DEVICE(p);
  if (t > 0 && p->T_pos <= t && p->cls->table.info.moving_nodes) return fail("position history level not stored");
  CUDA_OK(cudaMemcpy(p->d_dofs, dofs, (size_t)p->n_dof * sizeof(double), cudaMemcpyHostToDevice));
  return scatter_dofs_from_device(p, 0, t);
}

// oomph's Problem::shift_time_values (TimeStepper::shift_time_values of every Data, timesteppers.h): history level t takes the
// values of level t-1, level 0 keeps the current values as the initial guess of the new step; device to device, no host copy
extern "C" int pb2_problem_shift_time_values(pb2_problem *p)
{
  NEED_DEVICE(p);
  const pb2_class_info &ci = p->cls->table.info;
  const size_t nv = (size_t)p->n_node * std::max(1, ci.nval) * sizeof(double), np_ = (size_t)p->n_node * ci.nodal_dim * sizeof(double);
  for (int t = p->T_val - 1; t >= 1; t--)
    CUDA_OK(cudaMemcpyAsync((char *)p->d_node_val + (size_t)t * nv, (char *)p->d_node_val + (size_t)(t - 1) * nv, nv, cudaMemcpyDeviceToDevice, 0));
  for (int t = p->T_pos - 1; t >= 1; t--)
    CUDA_OK(cudaMemcpyAsync((char *)p->d_node_pos + (size_t)t * np_, (char *)p->d_node_pos + (size_t)(t - 1) * np_, np_, cudaMemcpyDeviceToDevice, 0));
  return inputs_changed(p);
}

extern "C" int pb2_problem_set_time(pb2_problem *p, const pb2_time_info *ti)
{
  p->ti = *ti;
  return 0;
}

extern "C" int pb2_problem_set_parameters(pb2_problem *p, const double *values, int n)
{
  if (n > PB2_MAX_PARAMS) return fail("too many global parameters");
  for (int i = 0; i < n; i++) p->params[i] = values[i];
  return 0;
}

// error word of the kernels (tile gate timed out): reported once, by the next call that could hand out results
static int check_status(pb2_problem *p)
{
  if (p->h_status && *(volatile int *)p->h_status != 0)
  {
    const int st = *(volatile int *)p->h_status;
    *(volatile int *)p->h_status = 0;
    return fail(st == PB2_STATUS_GATE_TIMEOUT ? "a tile gate of the persistent assembly kernel timed out (blocks not co-resident); the results of that assembly are invalid"
                                              : "the assembly kernel reported error " + std::to_string(st));
  }
  return 0;
}

static void fill_common_args(pb2_problem *p, pb2_kernel_args &a)
{
  memset(&a, 0, sizeof(a));
  a.status = p->d_status;
  a.elem_nodes = p->d_elem_nodes;
  a.elem_eqn = p->d_elem_eqn;
  a.elem_rowstart = p->d_elem_rowstart;
  a.elem_off = p->d_elem_off;
  a.map_bits = p->map_bits;
  a.elem_res = p->d_elem_res;
  a.node_pos = p->d_node_pos;
  a.node_lagr = p->d_node_lagr;
  a.node_val = p->d_node_val;
  a.n_node = p->n_node;
  a.n_hist_val = p->T_val;
  a.n_hist_pos = p->T_pos;
  a.residual = p->d_residual;
  a.ti = p->ti;
  memcpy(a.params, p->params, sizeof(a.params));
}

static int run_routine(pb2_problem *p, int kind, int residual_index, int param_index, unsigned flag, double *out_jac, double *out_mass, const double *hvec, void *cuda_stream)
{
  const pb2_class_info &ci = p->cls->table.info;
  if (residual_index < 0 || residual_index >= ci.n_residuals) return fail("residual index out of range");
  if (param_index >= ci.n_params) return fail("parameter index out of range");
  if (check_status(p)) return 1;
  if (cuda_stream) CUDA_OK(cudaStreamWaitEvent((cudaStream_t)cuda_stream, p->ev_inputs, 0));
  pb2_kernel_args a;
  fill_common_args(p, a);
  a.jac_vals = out_jac;
  a.mass_vals = out_mass;
  a.hvec = hvec;
  p->launches_last = 0;
  if (p->n_untouched > 0 && flag >= 1u)
  {
    // entries that only other ranks contribute to start from zero in every assembly
    const int bs = 256;
    pb2_zero_positions<<<(unsigned)((p->n_untouched + bs - 1) / bs), bs, 0, (cudaStream_t)cuda_stream>>>(out_jac, flag >= 2u ? out_mass : nullptr, p->d_untouched, p->n_untouched);
    CUDA_OK(cudaGetLastError());
    p->launches_last++;
    p->launches_total++;
  }
  if (p->n_untouched_rows > 0 && p->n_children > 0 && kind == 0)
  {
    // residual rows only a child class contributes to restart from zero with the parent's assembly
    const int bs = 256;
    pb2_zero_positions<<<(unsigned)((p->n_untouched_rows + bs - 1) / bs), bs, 0, (cudaStream_t)cuda_stream>>>(a.residual, nullptr, p->d_untouched_rows, p->n_untouched_rows);
    CUDA_OK(cudaGetLastError());
    p->launches_last++;
    p->launches_total++;
  }
  pb2_kernel_cfg cfg;
  int rc = p->cls->table.query(kind, residual_index, param_index, flag, &cfg);
  if (rc != 0)
    return fail("plugin has no kernel for this routine (rc " + std::to_string(rc) + (rc >= 100 ? std::string(": ") + cudaGetErrorString((cudaError_t)(rc - 100)) : "") + ")");
  const int ntile = (int)p->colour_begin.size() - 1;
  if (!cfg.pipelined && ntile > 4096) return fail("the phase-synchronous kernels (PB2_PIPELINE=0) need PB2_UNIT_PATCHES large enough to keep the launch count reasonable");
  if (cfg.pipelined)
  {
    // one persistent launch; every block walks its own list of batches (units of one tile after the other)
    const int ncol = p->n_colours, nunit = (int)p->unit_tile.size();
    long long nb_total = 0;
    if (p->local_order)
      for (int u = 0; u < nunit; u++) nb_total += (p->unit_begin[u + 1] - p->unit_begin[u] + cfg.elems_per_batch - 1) / cfg.elems_per_batch;
    else
      for (int u = 0; u < nunit; u++)
        for (int c = 0; c < ncol; c++)
        {
          const int n = p->colour_begin[(size_t)u * ncol + c + 1] - p->colour_begin[(size_t)u * ncol + c];
          nb_total += (n + cfg.elems_per_batch - 1) / cfg.elems_per_batch;
        }
    // every block must be resident (the tile gate spins): grid <= SMs x occupancy
    const int grid = (int)std::min<long long>(std::max<long long>(1, nb_total), (long long)p->n_sms * cfg.blocks_per_sm);
    pb2_problem::BatchTables *bt = nullptr;
    for (auto &b : p->batch_tables)
      if (b.epb == cfg.elems_per_batch && b.grid == grid) bt = &b;
    if (!bt)
    {
      pb2_problem::BatchTables nb;
      nb.epb = cfg.elems_per_batch;
      nb.grid = grid;
      std::vector<std::vector<int>> be(grid), bm(grid);
      std::vector<std::vector<unsigned long long>> bbar(grid);
      std::vector<int> tn(std::max(1, p->n_tiles), 0);
      std::vector<int> stamp;
      int epoch = 0;
      if (p->local_order) stamp.assign((size_t)p->n_node, -1);
      int u = 0;
      for (int t = 0; t < p->n_tiles; t++)
      {
        int k = 0;
        for (; u < nunit && p->unit_tile[u] == t; u++, k++)
        {
          const int blk = (k + t * 37) % grid; // round robin, rotated per tile so that remainders do not pile up
          if (p->local_order)
          {
            // batches = consecutive elements of the unit; bit i of the barrier mask: element i of the batch shares a node
            // with an element scattered since the last barrier, so the scatter warps synchronise before it (the order
            // of the contributions to every CSR entry is then the element order: deterministic, no lost first touch)
            for (int e = p->unit_begin[u]; e < p->unit_begin[u + 1]; e += nb.epb)
            {
              const int nel = std::min(nb.epb, p->unit_begin[u + 1] - e);
              unsigned long long mask = 0;
              ++epoch;
              for (int i = 0; i < nel; i++)
              {
                const int *en = &p->h_elem_nodes[(size_t)(e + i) * p->nnode];
                bool conflict = false;
                for (int l = 0; l < p->nnode; l++) conflict |= stamp[en[l]] == epoch;
                if (conflict)
                {
                  mask |= 1ull << i;
                  ++epoch;
                }
                for (int l = 0; l < p->nnode; l++) stamp[en[l]] = epoch;
              }
              be[blk].push_back(e);
              bm[blk].push_back((t << 7) | nel);
              bbar[blk].push_back(mask);
              tn[t]++;
            }
            continue;
          }
          bool first_of_unit = true;
          for (int c = 0; c < ncol; c++)
          {
            const int b0 = p->colour_begin[(size_t)u * ncol + c], b1 = p->colour_begin[(size_t)u * ncol + c + 1];
            bool first_of_colour = true;
            for (int e = b0; e < b1; e += nb.epb)
            {
              // bit 6: this batch must wait for the previous batch of the same unit (a new element colour starts)
              const int fence = (!first_of_unit && first_of_colour) ? 64 : 0;
              be[blk].push_back(e);
              bm[blk].push_back((t << 7) | fence | std::min(nb.epb, b1 - e));
              bbar[blk].push_back(0ull);
              tn[t]++;
              first_of_colour = false;
              first_of_unit = false;
            }
          }
        }
      }
      std::vector<int> fe, fm, bb(grid + 1, 0);
      std::vector<unsigned long long> fb;
      for (int b = 0; b < grid; b++)
      {
        bb[b + 1] = bb[b] + (int)be[b].size();
        fe.insert(fe.end(), be[b].begin(), be[b].end());
        fm.insert(fm.end(), bm[b].begin(), bm[b].end());
        fb.insert(fb.end(), bbar[b].begin(), bbar[b].end());
      }
      nb.n_batches = (int)fe.size();
      nb.n_tiles = p->n_tiles;
      if (upload(&nb.d_batch_elem, fe) || upload(&nb.d_batch_meta, fm) || upload(&nb.d_batch_bar, fb) || upload(&nb.d_tile_nbatch, tn) || upload(&nb.d_block_begin, bb)) return 1;
      CUDA_OK(cudaMalloc((void **)&nb.d_tile_done, std::max(1, nb.n_tiles) * sizeof(int)));
      p->batch_tables.push_back(nb);
      bt = &p->batch_tables.back();
    }
    CUDA_OK(cudaMemsetAsync(bt->d_tile_done, 0, std::max(1, bt->n_tiles) * sizeof(int), (cudaStream_t)cuda_stream));
    a.batch_elem = bt->d_batch_elem;
    a.batch_meta = bt->d_batch_meta;
    a.batch_bar = bt->d_batch_bar;
    a.tile_nbatch = bt->d_tile_nbatch;
    a.tile_done = bt->d_tile_done;
    a.block_begin = bt->d_block_begin;
    a.n_batches = bt->n_batches;
    a.n_tiles = bt->n_tiles;
    a.n_elem = (int)p->n_elem;
    if (getenv("PB2_TIMING"))
    {
      if (!p->d_debug) CUDA_OK(cudaMalloc((void **)&p->d_debug, 64 * sizeof(unsigned long long)));
      CUDA_OK(cudaMemsetAsync(p->d_debug, 0, 64 * sizeof(unsigned long long), (cudaStream_t)cuda_stream));
      a.debug = p->d_debug;
    }
    rc = p->cls->table.launch(&cfg, &a, grid, cuda_stream);
    if (rc != 0) return fail("kernel launch failed (plugin rc " + std::to_string(rc) + (rc >= 100 ? std::string(": ") + cudaGetErrorString((cudaError_t)(rc - 100)) : "") + ")");
    p->launches_last++;
    p->launches_total++;
    if (a.debug)
    {
      unsigned long long h[64];
      CUDA_OK(cudaMemcpy(h, p->d_debug, sizeof(h), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[pb2 timing] grid %d batches %d:", grid, bt->n_batches);
      for (int i = 0; i < 24; i++) fprintf(stderr, " %.0f", (double)h[i] / grid);
      fprintf(stderr, "\n");
    }
    return 0;
  }
  if (p->local_order) return fail("the phase-synchronous kernels (PB2_PIPELINE=0) need the colour-major element order (PB2_ORDER=colour)");
  for (int c = 0; c < ntile; c++)
  {
    a.elem_begin = p->colour_begin[c];
    a.n_elem = p->colour_begin[c + 1] - p->colour_begin[c];
    if (a.n_elem == 0) continue;
    const int nbatch = (a.n_elem + cfg.elems_per_batch - 1) / cfg.elems_per_batch;
    rc = p->cls->table.launch(&cfg, &a, std::min(nbatch, p->n_sms * cfg.blocks_per_sm), cuda_stream);
    if (rc != 0)
      return fail("kernel launch failed (plugin rc " + std::to_string(rc) + (rc >= 100 ? std::string(": ") + cudaGetErrorString((cudaError_t)(rc - 100)) : "") + ")");
    p->launches_last++;
    p->launches_total++;
  }
  return 0;
}

// ---- hanging-node constraints: J = P^T J_ext P, R = P^T R_ext on the assembled extended system (pyoomph_b200/hanging.py).
// One thread per target: a fixed-order weighted sum of its sources (entries of virtual rows / columns, which no target is) -- no atomics.
static __global__ void pb2_constraint_gather(double *__restrict__ a, double *__restrict__ b, const int *__restrict__ target, const int *__restrict__ start,
                                             const int *__restrict__ src, const double *__restrict__ w, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = target[i];
  double sa = 0.0, sb = 0.0;
  for (int k = start[i]; k < start[i + 1]; k++)
  {
    sa += w[k] * a[src[k]];
    if (b) sb += w[k] * b[src[k]];
  }
  a[t] += sa;
  if (b) b[t] += sb;
}

// virtual rows and columns: cleared, unit diagonal in the Jacobian (zero in a mass matrix or a parameter derivative), residual zero
// (the diagonal entries are among the cleared ones: they are written by the launch after this one)
static __global__ void pb2_constraint_clear(double *__restrict__ jac, double *__restrict__ mass, double *__restrict__ res, const int *__restrict__ clear, long long nclear,
                                            const int *__restrict__ rows, long long nvirt)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nclear)
  {
    if (jac) jac[clear[i]] = 0.0;
    if (mass) mass[clear[i]] = 0.0;
  }
  if (i < nvirt && res) res[rows[i]] = 0.0;
}

static __global__ void pb2_constraint_diag(double *__restrict__ jac, const int *__restrict__ diag, long long nvirt, double diag_value)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nvirt) jac[diag[i]] = diag_value;
}

extern "C" int pb2_problem_set_constraints(pb2_problem *p, long long n_target, const int *target_pos, const int *src_start, const int *src_pos, const double *src_w,
                                           long long n_res, const int *res_row, const int *res_start, const int *res_src, const double *res_w,
                                           long long n_clear, const int *clear_pos, long long n_virtual, const int *virtual_rows, const int *virtual_diag_pos)
{
  NEED_DEVICE(p);
  if (p->parent || p->n_children) return fail("constraints on parent / child problems are not supported");
  for (long long i = 0; i < n_target; i++)
    if (target_pos[i] < 0 || target_pos[i] >= p->nnz) return fail("constraint target outside the matrix");
  const long long ns = n_target ? src_start[n_target] : 0, nrs = n_res ? res_start[n_res] : 0;
  for (long long i = 0; i < ns; i++)
    if (src_pos[i] < 0 || src_pos[i] >= p->nnz) return fail("constraint source outside the matrix");
  for (long long i = 0; i < n_virtual; i++)
    if (virtual_rows[i] < 0 || virtual_rows[i] >= p->n_dof || virtual_diag_pos[i] < 0 || virtual_diag_pos[i] >= p->nnz) return fail("virtual equation outside the system");
  for (int **d : {&p->d_c_target, &p->d_c_src_start, &p->d_c_src_pos, &p->d_c_res_row, &p->d_c_res_start, &p->d_c_res_src, &p->d_c_clear, &p->d_c_virt_rows, &p->d_c_diag})
  {
    cudaFree(*d); // replacing earlier lists
    *d = nullptr;
  }
  cudaFree(p->d_c_src_w); cudaFree(p->d_c_res_w);
  p->d_c_src_w = p->d_c_res_w = nullptr;
  p->c_nvirt = 0;
  std::vector<int> v;
  auto up_i = [&](int **d, const int *h, long long n) { v.assign(h, h + n); return upload(d, v); };
  std::vector<double> vd;
  auto up_d = [&](double **d, const double *h, long long n) { vd.assign(h, h + n); return upload(d, vd); };
  if (up_i(&p->d_c_target, target_pos, n_target) || up_i(&p->d_c_src_start, src_start, n_target + 1) || up_i(&p->d_c_src_pos, src_pos, ns) || up_d(&p->d_c_src_w, src_w, ns) ||
      up_i(&p->d_c_res_row, res_row, n_res) || up_i(&p->d_c_res_start, res_start, n_res + 1) || up_i(&p->d_c_res_src, res_src, nrs) || up_d(&p->d_c_res_w, res_w, nrs) ||
      up_i(&p->d_c_clear, clear_pos, n_clear) || up_i(&p->d_c_virt_rows, virtual_rows, n_virtual) || up_i(&p->d_c_diag, virtual_diag_pos, n_virtual))
    return 1;
  p->c_ntarget = n_target;
  p->c_nres = n_res;
  p->c_nclear = n_clear;
  p->c_nvirt = n_virtual;
  return 0;
}

// after an R / J / M assembly (or a parameter derivative: zero instead of unit diagonal) of a problem with constraints
static int apply_constraints(pb2_problem *p, unsigned flag, bool parameter_derivative, double *jac, double *mass, void *cuda_stream)
{
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const int bs = 256;
  auto grid = [&](long long n) { return (unsigned)std::max<long long>(1, (n + bs - 1) / bs); };
  if (flag >= 1u && p->c_ntarget > 0)
  {
    pb2_constraint_gather<<<grid(p->c_ntarget), bs, 0, st>>>(jac, flag >= 2u ? mass : nullptr, p->d_c_target, p->d_c_src_start, p->d_c_src_pos, p->d_c_src_w, p->c_ntarget);
    p->launches_last++;
  }
  if (p->c_nres > 0)
  {
    pb2_constraint_gather<<<grid(p->c_nres), bs, 0, st>>>(p->d_residual, nullptr, p->d_c_res_row, p->d_c_res_start, p->d_c_res_src, p->d_c_res_w, p->c_nres);
    p->launches_last++;
  }
  pb2_constraint_clear<<<grid(std::max(p->c_nclear, p->c_nvirt)), bs, 0, st>>>(flag >= 1u ? jac : nullptr, flag >= 2u ? mass : nullptr, p->d_residual, p->d_c_clear, p->c_nclear,
                                                                             p->d_c_virt_rows, p->c_nvirt);
  p->launches_last++;
  if (flag >= 1u && !parameter_derivative && p->c_nvirt > 0)
  {
    pb2_constraint_diag<<<grid(p->c_nvirt), bs, 0, st>>>(jac, p->d_c_diag, p->c_nvirt, 1.0);
    p->launches_last++;
  }
  CUDA_OK(cudaGetLastError());
  p->launches_total += 4;
  return 0;
}

extern "C" int pb2_problem_assemble(pb2_problem *p, int residual_index, int param_index, unsigned flag, void *cuda_stream)
{
  NEED_DEVICE(p);
  if (flag > 2u) return fail("flag must be 0, 1 or 2");
  if (p->parent)
  {
    // a child adds to the matrices its parent assembled (flag 2: the parent's mass matrix must exist)
    if (flag == 2u && !p->parent->d_mass) return fail("assemble the parent with flag 2 before its children: no mass matrix yet");
    return run_routine(p, 0, residual_index, param_index, flag, p->parent->d_jac, p->parent->d_mass, nullptr, cuda_stream);
  }
  if (flag == 2u && !p->d_mass) CUDA_OK(cudaMalloc((void **)&p->d_mass, std::max<size_t>(1, p->nnz) * sizeof(double)));
  const int rc = run_routine(p, 0, residual_index, param_index, flag, p->d_jac, p->d_mass, nullptr, cuda_stream);
  if (rc == 0 && p->c_nvirt > 0) return apply_constraints(p, flag, param_index >= 0, p->d_jac, p->d_mass, cuda_stream);
  return rc;
}

extern "C" int pb2_problem_assemble_hessian(pb2_problem *p, int residual_index, unsigned flag, int n_vec, const double *Y, void *cuda_stream)
{
  NEED_DEVICE(p);
  if (!p->cls->table.info.hessian_generated) return fail("this element class was generated without Hessian routines");
  if (p->c_nvirt > 0) return fail("Hessian routines of a problem with hanging-node constraints are not supported");
  if (flag != 1u && flag != 2u && flag != 4u && flag != 5u)
    return fail("Hessian assembly: flag must be 1 (d(J.Y)/dU), 2 (+ d(M.Y)/dU), 4 (d(J^T.Y)/dU) or 5 (+ d(M^T.Y)/dU)");
  const int hkind = flag >= 4u ? 3 : 1;     // transposed contraction: plugin kind 3
  if (flag >= 4u) flag -= 3u;
  if (n_vec < 1) return fail("n_vec must be positive");
  if (n_vec > p->hess_nvec)
  {
    cudaFree(p->d_hvec); cudaFree(p->d_hess); cudaFree(p->d_hessM);
    p->d_hvec = p->d_hess = p->d_hessM = nullptr;
    CUDA_OK(cudaMalloc((void **)&p->d_hvec, (size_t)n_vec * std::max<long long>(1, p->n_dof) * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&p->d_hess, (size_t)n_vec * std::max<long long>(1, p->nnz) * sizeof(double)));
    CUDA_OK(cudaMalloc((void **)&p->d_hessM, (size_t)n_vec * std::max<long long>(1, p->nnz) * sizeof(double)));
    p->hess_nvec = n_vec;
  }
  CUDA_OK(cudaMemcpyAsync(p->d_hvec, Y, (size_t)n_vec * p->n_dof * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)cuda_stream));
  long long launches = 0;
  if (p->parent)
  {
    // a child has no first-touch stores: its Hessian contribution (kept in its own buffers, to be added to the parent's) starts from zero
    CUDA_OK(cudaMemsetAsync(p->d_hess, 0, (size_t)n_vec * std::max<long long>(1, p->nnz) * sizeof(double), (cudaStream_t)cuda_stream));
    if (flag == 2u) CUDA_OK(cudaMemsetAsync(p->d_hessM, 0, (size_t)n_vec * std::max<long long>(1, p->nnz) * sizeof(double), (cudaStream_t)cuda_stream));
  }
  for (int v = 0; v < n_vec; v++)
  {
    const int rc = run_routine(p, hkind, residual_index, -1, flag, p->d_hess + (size_t)v * p->nnz, p->d_hessM + (size_t)v * p->nnz, p->d_hvec + (size_t)v * p->n_dof, cuda_stream);
    if (rc) return rc;
    launches += p->launches_last;
  }
  p->launches_last = launches;
  return 0;
}

extern "C" int pb2_problem_fetch_hessian(pb2_problem *p, int v, double *jac_hessian_vals, double *mass_hessian_vals)
{
  NEED_DEVICE(p);
  if (v < 0 || v >= p->hess_nvec) return fail("Hessian vector index out of range");
  CUDA_OK(cudaDeviceSynchronize());
  if (check_status(p)) return 1;
  if (jac_hessian_vals) CUDA_OK(cudaMemcpy(jac_hessian_vals, p->d_hess + (size_t)v * p->nnz, (size_t)p->nnz * sizeof(double), cudaMemcpyDeviceToHost));
  if (mass_hessian_vals) CUDA_OK(cudaMemcpy(mass_hessian_vals, p->d_hessM + (size_t)v * p->nnz, (size_t)p->nnz * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// one warp per CSR row: y = A x
static __global__ void pb2_spmv(const int *__restrict__ row_start, const int *__restrict__ col, const double *__restrict__ val,
                                const double *__restrict__ x, double *__restrict__ y, long long n_rows)
{
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n_rows) return;
  double s = 0.0;
  for (int i = row_start[row] + lane; i < row_start[row + 1]; i += 32) s += val[i] * x[col[i]];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) y[row] = s;
}

extern "C" int pb2_problem_hessian_vector_products(pb2_problem *p, int residual_index, const double *Y, const double *C, int n_vec, double *products)
{
  int rc = pb2_problem_assemble_hessian(p, residual_index, 1u, 1, Y, nullptr);
  if (rc) return rc;
  if (!p->d_row_start)
  {
    if (upload(&p->d_row_start, p->row_start) || upload(&p->d_col_index, p->parent ? p->parent->col_index : p->col_index)) return 1;
  }
  double *d_x = nullptr, *d_y = nullptr;
  CUDA_OK(cudaMalloc((void **)&d_x, std::max<long long>(1, p->n_dof) * sizeof(double)));
  CUDA_OK(cudaMalloc((void **)&d_y, std::max<long long>(1, p->n_dof) * sizeof(double)));
  for (int v = 0; v < n_vec; v++)
  {
    CUDA_OK(cudaMemcpy(d_x, C + (size_t)v * p->n_dof, (size_t)p->n_dof * sizeof(double), cudaMemcpyHostToDevice));
    const int bs = 256;
    const long long nb = (p->n_dof * 32 + bs - 1) / bs;
    pb2_spmv<<<(unsigned)nb, bs>>>(p->d_row_start, p->d_col_index, p->d_hess, d_x, d_y, p->n_dof);
    CUDA_OK(cudaGetLastError());
    p->launches_last++;
    p->launches_total++;
    CUDA_OK(cudaMemcpy(products + (size_t)v * p->n_dof, d_y, (size_t)p->n_dof * sizeof(double), cudaMemcpyDeviceToHost));
  }
  cudaFree(d_x);
  cudaFree(d_y);
  return 0;
}

// ---- integral expressions: per-element values by the generated kernel, then one fixed-order reduction per expression
static __global__ void __launch_bounds__(1024) pb2_reduce_integrals(const double *__restrict__ per_elem, long long n_elem, int n_int, double *__restrict__ out)
{
  __shared__ double s[1024];
  const int k = blockIdx.x;
  double acc = 0.0;
  for (long long e = threadIdx.x; e < n_elem; e += 1024) acc += per_elem[e * n_int + k]; // same elements, same order, every time
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1)
  {
    if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[k] = s[0];
}

extern "C" int pb2_problem_eval_integrals(pb2_problem *p, double *out, int n_out)
{
  NEED_DEVICE(p);
  const pb2_class_info &ci = p->cls->table.info;
  if (ci.n_integrals < 1) return fail("this element class defines no integral expressions");
  if (n_out != ci.n_integrals) return fail("n_out must equal the number of integral expressions of the class");
  pb2_kernel_cfg cfg;
  int rc = p->cls->table.query(2, 0, -1, 0u, &cfg);
  if (rc != 0) return fail("plugin has no integral-expression kernel (rc " + std::to_string(rc) + ")");
  if (!p->d_integrals)
    CUDA_OK(cudaMalloc((void **)&p->d_integrals, ((size_t)std::max<long long>(1, p->n_elem) + 1) * ci.n_integrals * sizeof(double)));
  pb2_kernel_args a;
  fill_common_args(p, a);
  a.elem_begin = 0;
  a.n_elem = (int)p->n_elem;
  a.integrals = p->d_integrals;
  p->launches_last = 0;
  double *d_sum = p->d_integrals + (size_t)std::max<long long>(1, p->n_elem) * ci.n_integrals;
  if (p->n_elem > 0)
  {
    const long long nbatch = (p->n_elem + cfg.elems_per_batch - 1) / cfg.elems_per_batch;
    rc = p->cls->table.launch(&cfg, &a, (int)std::min<long long>(nbatch, (long long)p->n_sms * cfg.blocks_per_sm), nullptr);
    if (rc != 0) return fail("integral kernel launch failed (plugin rc " + std::to_string(rc) + (rc >= 100 ? std::string(": ") + cudaGetErrorString((cudaError_t)(rc - 100)) : "") + ")");
    p->launches_last++;
  }
  pb2_reduce_integrals<<<ci.n_integrals, 1024>>>(p->d_integrals, p->n_elem, ci.n_integrals, d_sum);
  CUDA_OK(cudaGetLastError());
  p->launches_last++;
  p->launches_total += p->launches_last;
  CUDA_OK(cudaMemcpy(out, d_sum, (size_t)ci.n_integrals * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- point expressions: EvalLocalExpression / EvalExtremumExpression / GetZ2Fluxes of every element at one point set
extern "C" int pb2_problem_eval_points(pb2_problem *p, int point_set, double *out, long long n_out)
{
  NEED_DEVICE(p);
  const pb2_class_info &ci = p->cls->table.info;
  if (ci.n_point_exprs < 1) return fail("this element class defines no local / extremum expressions or Z2 fluxes");
  if (point_set != 0 && point_set != 1) return fail("point set must be 0 (integration points) or 1 (nodes)");
  const int npts = point_set == 0 ? ci.n_int_pt : ci.nnode;
  const long long per_elem = (long long)npts * ci.n_point_exprs, total = p->n_elem * per_elem;
  if (n_out != total) return fail("n_out must be n_elem * points * expressions");
  if (check_status(p)) return 1;
  pb2_kernel_cfg cfg;
  int rc = p->cls->table.query(4, 0, -1, (unsigned)point_set, &cfg);
  if (rc != 0) return fail("plugin has no point-expression kernel (rc " + std::to_string(rc) + ")");
  p->launches_last = 0;
  if (p->n_elem == 0) return 0;
  double *d_buf = nullptr;
  CUDA_OK(cudaMalloc((void **)&d_buf, (size_t)total * sizeof(double)));
  pb2_kernel_args a;
  fill_common_args(p, a);
  a.elem_begin = 0;
  a.n_elem = (int)p->n_elem;
  a.integrals = d_buf;
  const long long nbatch = (p->n_elem + cfg.elems_per_batch - 1) / cfg.elems_per_batch;
  rc = p->cls->table.launch(&cfg, &a, (int)std::min<long long>(nbatch, (long long)p->n_sms * std::max(1, cfg.blocks_per_sm)), nullptr);
  if (rc != 0)
  {
    cudaFree(d_buf);
    return fail("point-expression kernel launch failed (plugin rc " + std::to_string(rc) + (rc >= 100 ? std::string(": ") + cudaGetErrorString((cudaError_t)(rc - 100)) : "") + ")");
  }
  p->launches_last++;
  p->launches_total++;
  std::vector<double> tmp((size_t)total);
  const cudaError_t e = cudaMemcpy(tmp.data(), d_buf, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_buf);
  if (e != cudaSuccess) return fail(std::string("cudaMemcpy: ") + cudaGetErrorString(e));
  // the kernel works in schedule order: hand the values back in the mesh's element order
#pragma omp parallel for schedule(static)
  for (long long q = 0; q < p->n_elem; q++) memcpy(out + (size_t)p->perm[q] * per_elem, tmp.data() + (size_t)q * per_elem, (size_t)per_elem * sizeof(double));
  return 0;
}

extern "C" int pb2_problem_device_outputs(pb2_problem *p, double **residual, double **jac_vals, double **mass_vals)
{
  if (residual) *residual = p->d_residual;
  if (jac_vals) *jac_vals = p->d_jac;
  if (mass_vals) *mass_vals = p->d_mass;
  return 0;
}

// ---- device-resident hand-off to a GPU linear solver (SURVEY N-d): the CSR pattern on the device next to the values, and the Newton
// update applied where the dofs live, so that per iteration only what the host really needs crosses the host link
extern "C" int pb2_problem_device_pattern(pb2_problem *p, int **row_start, int **column_index)
{
  NEED_DEVICE(p);
  if (!p->d_row_start)
  {
    if (upload(&p->d_row_start, p->row_start) || upload(&p->d_col_index, p->parent ? p->parent->col_index : p->col_index)) return 1;
  }
  if (row_start) *row_start = p->d_row_start;
  if (column_index) *column_index = p->d_col_index;
  return 0;
}

extern "C" int pb2_problem_device_dofs(pb2_problem *p, double **dofs)
{
  NEED_DEVICE(p);
  *dofs = p->d_dofs;
  return 0;
}

static __global__ void pb2_axpy_kernel(double *__restrict__ y, const double *__restrict__ x, double alpha, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += alpha * x[i];
}

// dofs += alpha * delta (both on the device), then the same scatter into the nodal storage as pb2_problem_set_dofs: the Newton update
// x -= dx of Problem::newton_solve (oomph-lib problem.cc) without a host round trip.  `cuda_stream`: the stream the solver produced
// delta on (NULL: default stream).
extern "C" int pb2_problem_update_dofs_device(pb2_problem *p, const double *d_delta, double alpha, void *cuda_stream)
{
  NEED_DEVICE(p);
  const int bs = 256;
  const long long nb = (p->n_dof + bs - 1) / bs;
  if (nb > 0)
  {
    pb2_axpy_kernel<<<(unsigned)nb, bs, 0, (cudaStream_t)cuda_stream>>>(p->d_dofs, d_delta, alpha, p->n_dof);
    CUDA_OK(cudaGetLastError());
    if (cuda_stream)
    {
      // the scatter and every later assembly are ordered behind the update through the inputs event
      pb2_scatter_dofs<<<(unsigned)nb, bs, 0, (cudaStream_t)cuda_stream>>>(p->d_dofs, p->d_dof_target, p->n_dof, p->d_node_val, p->d_node_pos);
      CUDA_OK(cudaGetLastError());
      CUDA_OK(cudaEventRecord(p->ev_inputs, (cudaStream_t)cuda_stream));
      CUDA_OK(cudaStreamWaitEvent(0, p->ev_inputs, 0));
      p->launches_total += 2;
      return 0;
    }
    p->launches_total++;
  }
  return scatter_dofs_from_device(p, 0);
}

// ---- interface exchange: pack / unpack-add of halo rows (one launch each; SM-count-multiple grids, 16-byte friendly index loads)
static __global__ void pb2_pack_kernel(const double *__restrict__ res, const double *__restrict__ jac, const double *__restrict__ mass,
                                       const long long *__restrict__ rows, long long n_rows, const long long *__restrict__ pos, long long n_pos,
                                       int nmat, double *__restrict__ buf)
{
  const long long total = n_rows + (long long)nmat * n_pos;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
  {
    if (i < n_rows)
      buf[i] = res[rows[i]];
    else
    {
      const long long j = i - n_rows;
      buf[i] = j < n_pos ? jac[pos[j]] : mass[pos[j - n_pos]];
    }
  }
}

static __global__ void pb2_unpack_add_kernel(double *__restrict__ res, double *__restrict__ jac, double *__restrict__ mass,
                                             const long long *__restrict__ rows, long long n_rows, const long long *__restrict__ pos, long long n_pos,
                                             int nmat, const double *__restrict__ buf)
{
  const long long total = n_rows + (long long)nmat * n_pos;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
  {
    if (i < n_rows)
      res[rows[i]] += buf[i];
    else
    {
      const long long j = i - n_rows;
      if (j < n_pos)
        jac[pos[j]] += buf[i];
      else
        mass[pos[j - n_pos]] += buf[i];
    }
  }
}

static int exchange_grid(pb2_problem *p, long long total) { return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)p->n_sms * 8)); }

extern "C" int pb2_problem_pack_rows(pb2_problem *p, const long long *rows, long long n_rows, const long long *pos, long long n_pos, unsigned flag,
                                     double *buf, void *cuda_stream)
{
  NEED_DEVICE(p);
  if (flag > 2u) return fail("flag must be 0, 1 or 2");
  if (flag == 2u && !p->d_mass) return fail("no mass matrix has been assembled");
  const long long total = n_rows + (long long)flag * n_pos;
  if (total <= 0) return 0;
  pb2_pack_kernel<<<exchange_grid(p, total), 256, 0, (cudaStream_t)cuda_stream>>>(p->d_residual, p->d_jac, p->d_mass, rows, n_rows, pos, n_pos, (int)flag, buf);
  CUDA_OK(cudaGetLastError());
  p->launches_total++;
  return 0;
}

extern "C" int pb2_problem_unpack_add(pb2_problem *p, const long long *rows, long long n_rows, const long long *pos, long long n_pos, unsigned flag,
                                      const double *buf, void *cuda_stream)
{
  NEED_DEVICE(p);
  if (flag > 2u) return fail("flag must be 0, 1 or 2");
  if (flag == 2u && !p->d_mass) return fail("no mass matrix has been assembled");
  const long long total = n_rows + (long long)flag * n_pos;
  if (total <= 0) return 0;
  pb2_unpack_add_kernel<<<exchange_grid(p, total), 256, 0, (cudaStream_t)cuda_stream>>>(p->d_residual, p->d_jac, p->d_mass, rows, n_rows, pos, n_pos, (int)flag, buf);
  CUDA_OK(cudaGetLastError());
  p->launches_total++;
  return 0;
}

extern "C" int pb2_problem_fetch(pb2_problem *p, double *residual, double *jac_vals, double *mass_vals)
{
  NEED_DEVICE(p);
  CUDA_OK(cudaDeviceSynchronize());
  if (check_status(p)) return 1;
  if (residual) CUDA_OK(cudaMemcpy(residual, p->d_residual, (size_t)p->n_dof * sizeof(double), cudaMemcpyDeviceToHost));
  if (jac_vals) CUDA_OK(cudaMemcpy(jac_vals, p->d_jac, (size_t)p->nnz * sizeof(double), cudaMemcpyDeviceToHost));
  if (mass_vals)
  {
    if (!p->d_mass) return fail("no mass matrix has been assembled");
    CUDA_OK(cudaMemcpy(mass_vals, p->d_mass, (size_t)p->nnz * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return 0;
}

extern "C" int pb2_problem_assemble_host(pb2_problem *p, int residual_index, int param_index, unsigned flag, const double *dofs,
                                         double *residual, double *jac_vals, double *mass_vals)
{
  if (dofs)
  {
    const int rc = pb2_problem_set_dofs(p, dofs);
    if (rc) return rc;
  }
  int rc = pb2_problem_assemble(p, residual_index, param_index, flag, nullptr);
  if (rc) return rc;
  if (dofs) p->launches_last += 1; // the dof scatter kernel
  return pb2_problem_fetch(p, residual, flag >= 1 ? jac_vals : nullptr, flag >= 2 ? mass_vals : nullptr);
}

extern "C" long long pb2_problem_launch_count(pb2_problem *p) { return p->launches_last; }

// ---- small utilities for hosts without their own CUDA bindings: pinned buffers and stream-0 event timing
extern "C" void *pb2_host_alloc(size_t nbytes)
{
  void *ptr = nullptr;
  if (cudaHostAlloc(&ptr, nbytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return ptr;
}
extern "C" void pb2_host_free(void *ptr) { cudaFreeHost(ptr); }

static cudaEvent_t g_events[64];
static int g_event_device[64]; // device + 1 the slot's event belongs to, 0 = not created
extern "C" int pb2_event_record(int idx, void *cuda_stream)
{
  if (idx < 0 || idx >= 64) return fail("event index out of range");
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  if (g_event_device[idx] != dev + 1)
  {
    // an event belongs to the device that was current when it was created: a slot reused on another device gets a new event
    if (g_event_device[idx] != 0) cudaEventDestroy(g_events[idx]);
    CUDA_OK(cudaEventCreate(&g_events[idx]));
    g_event_device[idx] = dev + 1;
  }
  CUDA_OK(cudaEventRecord(g_events[idx], (cudaStream_t)cuda_stream));
  return 0;
}
extern "C" int pb2_event_elapsed_ms(int i0, int i1, float *ms)
{
  if (i0 < 0 || i0 >= 64 || i1 < 0 || i1 >= 64 || !g_event_device[i0] || g_event_device[i0] != g_event_device[i1])
    return fail("events were not recorded on the same device");
  CUDA_OK(cudaEventSynchronize(g_events[i1]));
  CUDA_OK(cudaEventElapsedTime(ms, g_events[i0], g_events[i1]));
  return 0;
}
extern "C" int pb2_device_synchronize(void)
{
  CUDA_OK(cudaDeviceSynchronize());
  return 0;
}
extern "C" int pb2_device_count(int *n)
{
  CUDA_OK(cudaGetDeviceCount(n));
  return 0;
}
extern "C" double pb2_problem_setup_seconds(pb2_problem *p) { return p->setup_seconds; }

// ---- fp64 roofline denominator: dependent-chain-free DFMA throughput of the device ("P64 measured on the box by an FMA
// microbenchmark").  16 independent accumulators per thread, 4 blocks of 256 threads per SM: enough warps and ILP to keep the fp64
// pipe of every sub-partition issuing back to back; 2 flops per DFMA.
static __global__ void __launch_bounds__(256) pb2_fp64_fma_kernel(double *out, int iters, double x0)
{
  double a[16];
  const double m = 1.0 + 1e-9 * x0, c = 1e-9 * (threadIdx.x + 1);
#pragma unroll
  for (int i = 0; i < 16; i++) a[i] = x0 + i;
  for (int it = 0; it < iters; it++)
  {
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = fma(a[i], m, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += a[i];
  if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s; // never true: keeps the chain alive without a store
}

extern "C" int pb2_measure_fp64_peak(int device, double *tflops)
{
  CUDA_OK(cudaSetDevice(device));
  int n_sms = 0;
  CUDA_OK(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
  double *d_out = nullptr;
  const int grid = n_sms * 8, bs = 256, iters = 4096;
  CUDA_OK(cudaMalloc((void **)&d_out, (size_t)grid * bs * sizeof(double)));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++)
  {
    CUDA_OK(cudaEventRecord(e0, 0));
    pb2_fp64_fma_kernel<<<grid, bs>>>(d_out, iters, 1.0);
    CUDA_OK(cudaEventRecord(e1, 0));
    CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    const double tf = 2.0 * 16.0 * iters * (double)grid * bs / (ms * 1e-3) / 1e12;
    if (rep >= 1) best = std::max(best, tf); // first repetition warms up
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  CUDA_OK(cudaGetLastError());
  *tflops = best;
  return 0;
}

extern "C" int pb2_flush_l2(int device)
{
  // overwrite a buffer larger than the 126 MB L2
  static double *buf = nullptr;
  const size_t n = (size_t)256 << 20;
  CUDA_OK(cudaSetDevice(device));
  if (!buf) CUDA_OK(cudaMalloc((void **)&buf, n));
  CUDA_OK(cudaMemsetAsync(buf, 1, n, 0));
  return 0;
}
