// fb_api.cu — context lifetime, the step, and the extern "C" surface declared in
// include/fembrain_b200.h.  Reference interfaces replaced are cited in that header.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <set>
#include <string>
#include <thread>
#include <vector>

#include "fb_internal.h"

// ---- errors ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void fb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char *fb_last_error_string(void) { return g_err; }
extern "C" int fb_abi_version(void) { return FB_ABI_VERSION; }
extern "C" const char *fb_status_string(int s) {
  switch (s) {
    case FB_OK: return "ok";
    case FB_ERR_INVALID_ARGUMENT: return "invalid argument";
    case FB_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (there is no CPU fallback)";
    case FB_ERR_CUDA: return "CUDA runtime error";
    case FB_ERR_OUT_OF_MEMORY: return "out of device memory";
    case FB_ERR_SOLVER_NOT_CONVERGED: return "PCG did not converge within cg_max_iterations";
    case FB_ERR_BAD_MESH: return "bad mesh";
    case FB_ERR_COMM: return "communication error";
    case FB_ERR_NOT_SUPPORTED: return "not supported";
    default: return "unknown status";
  }
}

extern "C" void fb_default_params(fb_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->youngs_modulus = 1e7;
  p->poisson_ratio = 0.46;
  p->density = 1000.0;
  p->timestep = 0.0333;
  p->damping_mass = 0.0;
  p->damping_stiffness = 0.01;
  p->cg_epsilon = 1e-6;
  p->cg_max_iterations = 10000;
  p->polar_tolerance = 1e-6;
  p->internal_force_scaling = 1.0;
  p->device = 0;
  p->keep_raw_stiffness = 0;
  p->solver_variant = FB_SOLVER_JACOBI_PCG;
  p->warm_start = 0;
}

namespace {

#define CHECK_CTX(c)                         \
  do {                                       \
    if (!(c)) {                              \
      fb_set_error("NULL context");          \
      return FB_ERR_INVALID_ARGUMENT;        \
    }                                        \
    cudaError_t e_ = cudaSetDevice((c)->device); \
    if (e_ != cudaSuccess) {                 \
      fb_set_error("cudaSetDevice(%d): %s", (c)->device, cudaGetErrorString(e_)); \
      return FB_ERR_CUDA;                    \
    }                                        \
  } while (0)

__global__ void k_mark_fixed(int n, const int *__restrict__ dofs, unsigned char *__restrict__ fixed) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fixed[dofs[i]] = 1;
}
__global__ void k_axpy(int n, double a, const double *__restrict__ x, double *__restrict__ y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += a * x[i];
}
__global__ void k_count_isolated(int nV, const int *__restrict__ diag, int *__restrict__ out) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < nV && diag[v] < 0) atomicMin(out, v);
}
inline unsigned gridFor(size_t n, int tb) { return (unsigned)((n + tb - 1) / tb); }

int upload(fb_context *c, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return FB_OK;
  FB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}
int download(fb_context *c, void *dst, const void *src, size_t bytes) {
  if (bytes == 0) return FB_OK;
  FB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

void free_all(fb_context *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  fb_mg_release(c);
  fb_pcg_release(c);
  fb_dist_destroy(c);
  fb_batch_destroy(c);
  fb_sym_release(c);
  fb_tma_release(c);
  void *ptrs[] = {c->x0, c->tets, c->edata, c->bp, c->bc, c->brow, c->diag, c->seg, c->src, c->colIdx, c->mblk,
                  c->fixed, c->cdofs, c->T, c->Keff, c->Kraw, c->scrK, c->scrF, c->q, c->qvel, c->qaccel, c->fext,
                  c->fint, c->qres, c->rhs, c->x, c->res, c->dir, c->Ad, c->invD, c->tmp, c->sc, c->partials,
                  c->contact_dev, c->haptic_idx_dev, c->haptic_f_dev, c->edges_dev, c->edge_degree_dev, c->haptic_stamp, c->haptic_listA, c->haptic_listB, c->ctaRows, c->pers_prof, c->ga_incp, c->ga_ctaV, c->ga_lists, c->ga_erec, c->ga_xu};
  for (void *p : ptrs)
    if (p) fb_dev_free(p);
  if (c->sc_host) cudaFreeHost(c->sc_host);
  if (c->fext_host) cudaFreeHost(c->fext_host);
  for (auto &e : c->ev)
    if (e) cudaEventDestroy(e);
  for (auto &e : c->evChunk)
    if (e) cudaEventDestroy(e);
  for (auto &e : c->evProf)
    if (e) cudaEventDestroy(e);
  if (c->stream && !c->stream_borrowed) cudaStreamDestroy(c->stream);
  free(c->cdofs_host);
  free(c->haptic_idx_host);
  free(c->haptic_f_host);
  free(c->adj_host_bp);
  free(c->adj_host_bc);
  free(c->edges_host);
  free(c->edge_degree_host);
  delete c;
}

}  // namespace
int fb_apply_constraints(fb_context *c, int nC, const int *cdofs_sorted) {
  // validates like SparseMatrix::BuildRenumberingVector (sparseMatrix.cpp:896-938): in range, strictly ascending
  for (int i = 0; i < nC; i++) {
    if (cdofs_sorted[i] < 0 || cdofs_sorted[i] >= c->r) {
      fb_set_error("constrained DOF %d out of range [0, %d)", cdofs_sorted[i], c->r);
      return FB_ERR_INVALID_ARGUMENT;
    }
    if (i && cdofs_sorted[i] <= cdofs_sorted[i - 1]) {
      fb_set_error("constrained DOFs must be strictly ascending (duplicate or unsorted at position %d)", i);
      return FB_ERR_INVALID_ARGUMENT;
    }
  }
  if (c->cdofs) { fb_dev_free(c->cdofs); c->cdofs = nullptr; }
  free(c->cdofs_host);
  c->cdofs_host = (int *)malloc(sizeof(int) * (size_t)(nC ? nC : 1));
  memcpy(c->cdofs_host, cdofs_sorted, sizeof(int) * (size_t)nC);
  c->nC = nC;
  fb_mg_invalidate(c);   // coarse levels of a multigrid hierarchy carry the constraints too
  FB_TRY(fb_dev_alloc(c, &c->cdofs, (size_t)nC));
  FB_CUDA(cudaMemsetAsync(c->fixed, 0, (size_t)(c->r ? c->r : 1), c->stream));
  if (nC) {
    FB_TRY(upload(c, c->cdofs, c->cdofs_host, sizeof(int) * (size_t)nC));
    k_mark_fixed<<<gridFor(nC, 256), 256, 0, c->stream>>>(nC, c->cdofs, c->fixed);
    c->launches++;
  }
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

namespace {
int fixed_vertices_to_dofs(int nV, int nFixed, const int *fv, std::vector<int> &dofs) {
  // Deformable::FixedVerticesToFixedDOF (DEF/Deformable.cpp:294-314): sort, then 3v, 3v+1, 3v+2
  if (nFixed < 0 || (nFixed > 0 && !fv)) {
    fb_set_error("bad fixed vertex list");
    return FB_ERR_INVALID_ARGUMENT;
  }
  std::vector<int> s(fv, fv + nFixed);
  std::sort(s.begin(), s.end());
  dofs.resize(3 * (size_t)nFixed);
  for (int i = 0; i < nFixed; i++) {
    if (s[i] < 0 || s[i] >= nV) {
      fb_set_error("fixed vertex %d out of range [0, %d)", s[i], nV);
      return FB_ERR_INVALID_ARGUMENT;
    }
    if (i && s[i] == s[i - 1]) {
      fb_set_error("fixed vertex %d listed twice", s[i]);
      return FB_ERR_INVALID_ARGUMENT;
    }
    dofs[3 * (size_t)i] = 3 * s[i];
    dofs[3 * (size_t)i + 1] = 3 * s[i] + 1;
    dofs[3 * (size_t)i + 2] = 3 * s[i] + 2;
  }
  return FB_OK;
}

}  // namespace
int fb_create_local(fb_context **out, int nV, const double *x0, int nT, const int *tets, int nC, const int *cdofs,
                    const double *E, const double *nu, const double *rho, const fb_params *prm) {
  if (!out) { fb_set_error("out is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  *out = nullptr;
  if (nV < 0 || nT < 0 || (nV > 0 && !x0) || (nT > 0 && !tets)) {
    fb_set_error("bad mesh arguments (nV=%d, nT=%d)", nV, nT);
    return FB_ERR_INVALID_ARGUMENT;
  }
  // contribution offsets (seg, incidence prefix sums) are 32-bit signed: 16 nT must fit (same bound as fb_create_batch)
  if ((long long)nT * 16 > 0x7fffffffll || (long long)nV * 3 > 0x7fffffffll) {
    fb_set_error("mesh too large for 32-bit element/DOF ids; partition it");
    return FB_ERR_INVALID_ARGUMENT;
  }
  fb_params p;
  if (prm) p = *prm; else fb_default_params(&p);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    fb_set_error("no CUDA device visible; libfembrain_b200 has no CPU fallback");
    return FB_ERR_NO_DEVICE;
  }
  if (p.device < 0 || p.device >= ndev) { fb_set_error("device %d not in [0, %d)", p.device, ndev); return FB_ERR_NO_DEVICE; }
  cudaDeviceProp dp;
  FB_CUDA(cudaGetDeviceProperties(&dp, p.device));
  if (dp.major != 10) {
    fb_set_error("device %d is sm_%d%d; this library contains sm_100a code only", p.device, dp.major, dp.minor);
    return FB_ERR_NO_DEVICE;
  }
  FB_CUDA(cudaSetDevice(p.device));

  fb_context *c = new (std::nothrow) fb_context();
  if (!c) return FB_ERR_OUT_OF_MEMORY;
  memset(c, 0, sizeof(*c));
  c->device = p.device;
  c->sm_count = dp.multiProcessorCount;
  c->prm = p;
  c->nV = nV; c->nT = nT; c->r = 3 * nV;
  c->haptic_rings = 5;  // m_hapticForceNeighorhoodSize, DEF/Deformable.cpp ctor
  c->warp = 1;          // CorotationalLinearFEMForceModel(fem) — warp defaults to 1, Deformable.cpp:186
  c->uniform_material = (!E && !nu && !rho) ? 1 : 0;
  c->row_lo = 0; c->row_hi = nV;
  int st = FB_OK;
#define CR(call) do { st = (call); if (st != FB_OK) { free_all(c); return st; } } while (0)
#define CRC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fb_set_error("%s -> %s", #call, cudaGetErrorString(e__)); free_all(c); return e__ == cudaErrorMemoryAllocation ? FB_ERR_OUT_OF_MEMORY : FB_ERR_CUDA; } } while (0)
  CRC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {  // keep freed device memory in the default pool (see fb_dev_alloc); fb_trim_memory() releases it
    cudaMemPool_t pool;
    unsigned long long keep = ~0ull;
    if (cudaDeviceGetDefaultMemPool(&pool, p.device) != cudaSuccess ||
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) != cudaSuccess) cudaGetLastError();
  }
  for (auto &e : c->ev) CRC(cudaEventCreate(&e));
  for (auto &e : c->evChunk) CRC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CR(fb_dev_alloc(c, &c->x0, 3 * (size_t)nV));
  CR(fb_dev_alloc(c, &c->tets, 4 * (size_t)nT));
  CR(upload(c, c->x0, x0, sizeof(double) * 3 * (size_t)nV));
  CR(upload(c, c->tets, tets, sizeof(int) * 4 * (size_t)nT));
  CR(fb_build_topology(c));
  // a vertex in no tetrahedron has an empty matrix row: the reference reads diagonal index -1
  {
    int *flag = nullptr, h = 0x7fffffff;
    CRC(cudaMalloc(&flag, sizeof(int)));
    CRC(cudaMemcpyAsync(flag, &h, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (nV) { k_count_isolated<<<gridFor(nV, 256), 256, 0, c->stream>>>(nV, c->diag, flag); c->launches++; }
    CRC(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CRC(cudaStreamSynchronize(c->stream));
    fb_dev_free(flag);
    if (h != 0x7fffffff) {
      fb_set_error("vertex %d belongs to no tetrahedron", h);
      free_all(c);
      return FB_ERR_BAD_MESH;
    }
  }
  CR(fb_dev_alloc(c, &c->edata, 16 * (size_t)nT));
  double *dE = nullptr, *dnu = nullptr, *drho = nullptr;
  if (E) { CRC(cudaMalloc(&dE, sizeof(double) * (size_t)(nT ? nT : 1))); CR(upload(c, dE, E, sizeof(double) * (size_t)nT)); }
  if (nu) { CRC(cudaMalloc(&dnu, sizeof(double) * (size_t)(nT ? nT : 1))); CR(upload(c, dnu, nu, sizeof(double) * (size_t)nT)); }
  if (rho) { CRC(cudaMalloc(&drho, sizeof(double) * (size_t)(nT ? nT : 1))); CR(upload(c, drho, rho, sizeof(double) * (size_t)nT)); }
  st = fb_launch_element_data(c, dE, dnu, drho);
  cudaStreamSynchronize(c->stream);
  fb_dev_free(dE); fb_dev_free(dnu); fb_dev_free(drho);
  if (st != FB_OK) { free_all(c); return st; }
  CR(fb_dev_alloc(c, &c->mblk, (size_t)c->nB));
  CR(fb_launch_mass(c));
  CR(fb_dev_alloc(c, &c->fixed, (size_t)c->r));
  c->rowmask = c->fixed;
  CR(fb_apply_constraints(c, nC, cdofs));
  // T = hK + D is allocated on first use by the two-phase path only (fb_launch_assembly): the gather path never stores it
  CR(fb_dev_alloc(c, &c->Keff, (size_t)c->nnzK + 4));  // + slack: fb_tma.cu copies whole 16-byte lines around a row tile
  if (p.keep_raw_stiffness) CR(fb_dev_alloc(c, &c->Kraw, (size_t)c->nnzK));
  CR(fb_build_gather_plan(c));  // two-phase scratch (1248 B/tet) is allocated on first use, only if this plan is off
  double **vecs[] = {&c->q, &c->qvel, &c->qaccel, &c->fext, &c->fint, &c->qres, &c->rhs, &c->x, &c->res, &c->dir, &c->Ad, &c->invD, &c->tmp};
  for (double **v : vecs) {
    // the search direction is the one vector a neighbour rank writes into (halo push over CUDA IPC): not from the pool
    if (v == &c->dir) CR(fb_dev_alloc_plain(c, v, (size_t)c->r));
    else CR(fb_dev_alloc(c, v, (size_t)c->r));
    CRC(cudaMemsetAsync(*v, 0, sizeof(double) * (size_t)(c->r ? c->r : 1), c->stream));
  }
  CR(fb_dev_alloc(c, &c->sc, 1));
  CRC(cudaMemsetAsync(c->sc, 0, sizeof(FbScalars), c->stream));
  CR(fb_dev_alloc(c, &c->partials, 4 * (size_t)FB_MAX_PARTIALS));
  CR(fb_dev_alloc(c, &c->contact_dev, 1));
  CRC(cudaMallocHost(&c->sc_host, sizeof(FbScalars) * 4));
  memset(c->sc_host, 0, sizeof(FbScalars) * 4);
  // lanes per block row for the SpMV: rows hold 3*nb scalars
  double avg3 = nV ? 3.0 * (double)c->nB / (double)nV : 0.0;
  c->spmv_group = avg3 <= 12.0 ? 8 : (avg3 <= 72.0 ? 16 : 32);
  CR(fb_spmv_plan(c));
  // Optional L2 policy for the solver's matrix stream (FEMBRAIN_B200_L2PIN=f, default off): a fraction f of the
  // persisting carve-out is given to a window over Keff (hit = persisting, miss = streaming).  Measured on B200
  // (profiles/r01_l2pin.txt): no gain at f = 0.5 and a LOSS at f = 1 (PCG iteration 55.8 -> 68 us at 1M tets,
  // 523 -> 620 us at 10M) because the vectors lose their share of L2 — so it stays off.
  {
    const char *env = getenv("FEMBRAIN_B200_L2PIN");
    const double frac = env ? atof(env) : 0.0;
    if (frac > 0.0 && c->nnzK > 0 && dp.persistingL2CacheMaxSize > 0 && dp.accessPolicyMaxWindowSize > 0) {
      size_t persist = (size_t)(frac * (double)dp.persistingL2CacheMaxSize);
      if (persist > (size_t)dp.persistingL2CacheMaxSize) persist = (size_t)dp.persistingL2CacheMaxSize;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, persist) == cudaSuccess) {
        size_t win = sizeof(double) * (size_t)c->nnzK;
        if (win > (size_t)dp.accessPolicyMaxWindowSize) win = (size_t)dp.accessPolicyMaxWindowSize;
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = c->Keff;
        attr.accessPolicyWindow.num_bytes = win;
        double ratio = 0.95 * (double)persist / (double)win;
        attr.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess) c->l2_pinned_bytes = (size_t)(attr.accessPolicyWindow.hitRatio * (double)win);
      }
      cudaGetLastError();
    }
  }
  CRC(cudaStreamSynchronize(c->stream));
  CRC(cudaGetLastError());
  // the one variant that needs nothing but the matrix can be asked for at creation; FB_SOLVER_MG_PCG needs fb_set_grid first
  if (p.solver_variant == FB_SOLVER_BLOCK_JACOBI_PCG) CR(fb_set_solver(c, p.solver_variant, p.warm_start));
#undef CR
#undef CRC
  *out = c;
  return FB_OK;
}

// host copies of the block structure (inspection + haptic ring spreading)
int fb_fetch_structure(fb_context *c, std::vector<int> &bp, std::vector<int> &bc) {
  bp.resize((size_t)c->nV + 1);
  bc.resize((size_t)c->nB);
  FB_TRY(download(c, bp.data(), c->bp, sizeof(int) * bp.size()));
  FB_TRY(download(c, bc.data(), c->bc, sizeof(int) * bc.size()));
  return FB_OK;
}

namespace {
int fetch_structure(fb_context *c, std::vector<int> &bp, std::vector<int> &bc) { return fb_fetch_structure(c, bp, bc); }

// old DOF -> constrained DOF (or -1), the map of SparseMatrix::RemoveRowsColumns (sparseMatrix.cpp:1296-1322)
void old_to_new(const fb_context *c, std::vector<int> &m) {
  m.assign((size_t)c->r, 0);
  int dof = 0, cnt = 0;
  for (int i = 0; i < c->nC; i++) {
    while (dof < c->cdofs_host[i]) m[dof++] = cnt++;
    m[dof++] = -1;
  }
  while (dof < c->r) m[dof++] = cnt++;
}

// InsertRows / RemoveRows (VEGA/insertRows/insertRows.cpp:29-109) between host vectors
void expand_constrained(const fb_context *c, const double *xc, std::vector<double> &full) {
  full.assign((size_t)c->r, 0.0);
  int dst = 0, src = 0;
  for (int i = 0; i < c->nC; i++) {
    while (dst < c->cdofs_host[i]) full[dst++] = xc[src++];
    full[dst++] = 0.0;
  }
  while (dst < c->r) full[dst++] = xc[src++];
}
void compress_constrained(const fb_context *c, const double *full, double *xc) {
  int n = 0, row = 0;
  for (int i = 0; i < c->nC; i++) {
    while (row < c->cdofs_host[i]) xc[n++] = full[row++];
    row++;
  }
  while (row < c->r) xc[n++] = full[row++];
}

}  // namespace
int fb_do_step(fb_context *c) {
  cudaStream_t st = c->stream;
  FB_CUDA(cudaEventRecord(c->ev[0], st));
  // forceModel->GetForceAndMatrix(q, internalForces, tangentStiffnessMatrix) + Keff formation
  // ... and qresidual = (h K + D) qvel, in the reference's summation order; rhs = -h (qres + fint - fext)
  FB_TRY(fb_launch_assembly(c, c->q, c->prm.keep_raw_stiffness ? c->Kraw : nullptr, true, true));
  FB_CUDA(cudaEventRecord(c->ev[1], st));
  FB_CUDA(cudaEventRecord(c->ev[2], st));
  if (fb_mg_active(c)) FB_TRY(fb_mg_prepare(c));   // variants: coarse operators / FP32 copies of this step's Keff (timed with the solve)
  FB_TRY(fb_pcg_solve(c, c->prm.cg_epsilon, c->prm.cg_max_iterations));
  FB_CUDA(cudaEventRecord(c->ev[3], st));
  const bool failed = c->last_iters < 0;
  if (!failed) FB_TRY(fb_launch_state_update(c));
  FB_CUDA(cudaEventRecord(c->ev[4], st));
  FB_CUDA(cudaStreamSynchronize(st));
  FB_CUDA(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->ms_assembly = ms;
  cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->ms_solve = ms;
  cudaEventElapsedTime(&ms, c->ev[0], c->ev[4]); c->ms_step = ms;
  if (failed) {
    fb_set_error("PCG sparse solver returned non-zero exit status %d", c->last_iters);
    return FB_ERR_SOLVER_NOT_CONVERGED;
  }
  return FB_OK;
}

// =====================================================================================================
extern "C" {

int fb_create(fb_context **out, int nV, const double *x0, int nT, const int *tets, int nFixed, const int *fixedVerts,
              const fb_params *prm) {
  return fb_create_with_materials(out, nV, x0, nT, tets, nFixed, fixedVerts, nullptr, nullptr, nullptr, prm);
}

int fb_create_with_materials(fb_context **out, int nV, const double *x0, int nT, const int *tets, int nFixed,
                             const int *fixedVerts, const double *E, const double *nu, const double *rho,
                             const fb_params *prm) {
  std::vector<int> dofs;
  FB_TRY(fixed_vertices_to_dofs(nV, nFixed, fixedVerts, dofs));
  return fb_create_local(out, nV, x0, nT, tets, (int)dofs.size(), dofs.data(), E, nu, rho, prm);
}

int fb_create_with_constrained_dofs(fb_context **out, int nV, const double *x0, int nT, const int *tets, int nC,
                                    const int *cdofs, const fb_params *prm) {
  if (nC < 0 || (nC > 0 && !cdofs)) { fb_set_error("bad constrained DOF list"); return FB_ERR_INVALID_ARGUMENT; }
  return fb_create_local(out, nV, x0, nT, tets, nC, cdofs, nullptr, nullptr, nullptr, prm);
}

void fb_destroy(fb_context *c) { free_all(c); }

// Deformable::syncForceModel after a topology change (cutCompleted -> syncForceModel, main.cpp:614-617,
// DEF/Deformable.cpp:127-220): the reference deletes TetMesh, CorotationalLinearFEM, force model, mass matrix and integrator
// and rebuilds them all from the edited VolMesh on the same Deformable object; state restarts at rest (a new integrator).
// Same here on the same context handle: parameters, timestep / damping / CG settings and the Deformable-level options are
// kept, the mesh-sized buffers come back from the memory pool.  The solver variant falls back to the reference's solver
// unless it is the mesh-independent one (a cut mesh is no tensor grid any more).
int fb_sync_force_model(fb_context *c, int nV, const double *x0, int nT, const int *tets, int nFixed, const int *fixedVerts) {
  CHECK_CTX(c);
  if (c->dist || c->batch) { fb_set_error("fb_sync_force_model: recreate partitioned / batch contexts"); return FB_ERR_NOT_SUPPORTED; }
  std::vector<int> dofs;
  FB_TRY(fixed_vertices_to_dofs(nV, nFixed, fixedVerts, dofs));
  fb_params prm = c->prm;
  const int variant = fb_mg_active(c);
  prm.solver_variant = FB_SOLVER_JACOBI_PCG;
  fb_context *n = nullptr;
  FB_TRY(fb_create_local(&n, nV, x0, nT, tets, (int)dofs.size(), dofs.data(), nullptr, nullptr, nullptr, &prm));
  // Deformable-level options survive the re-setup (they are members of Deformable, not of the integrator)
  n->gravity = c->gravity; n->floor_enabled = c->floor_enabled; n->floor_y = c->floor_y; n->haptic_rings = c->haptic_rings;
  n->profiling = 0;
  n->launches += c->launches;
  // swap the guts so that the caller's handle stays valid, then release the old mesh's buffers (back to the pool)
  fb_context tmp = *c;
  *c = *n;
  *n = tmp;
  free_all(n);
  if (variant == FB_SOLVER_BLOCK_JACOBI_PCG) FB_TRY(fb_set_solver(c, variant, 0));
  return FB_OK;
}

int fb_set_fixed_vertices(fb_context *c, int nFixed, const int *fv) {
  CHECK_CTX(c);
  std::vector<int> dofs;
  if (c->dist) { fb_set_error("fb_set_fixed_vertices on a partitioned context: recreate it"); return FB_ERR_NOT_SUPPORTED; }
  FB_TRY(fixed_vertices_to_dofs(c->nV, nFixed, fv, dofs));
  return fb_apply_constraints(c, (int)dofs.size(), dofs.data());
}

int fb_num_vertices(const fb_context *c) {
  int nV = 0, nT = 0;
  if (c && fb_dist_global_sizes(c, &nV, &nT) == FB_OK) return nV;
  return c ? c->nV : 0;
}
int fb_num_tets(const fb_context *c) {
  int nV = 0, nT = 0;
  if (c && fb_dist_global_sizes(c, &nV, &nT) == FB_OK) return nT;
  return c ? c->nT : 0;
}
int fb_num_dofs(const fb_context *c) {
  int nV = 0, nT = 0;
  if (c && fb_dist_global_sizes(c, &nV, &nT) == FB_OK) return 3 * nV;
  return c ? c->r : 0;
}
int fb_num_constrained_dofs(const fb_context *c) { return c ? c->nC : 0; }
int fb_num_local_dofs(const fb_context *c) { return c ? c->r : 0; }
long long fb_nnz_stiffness(const fb_context *c) { return c ? c->nnzK : 0; }
long long fb_nnz_mass(const fb_context *c) { return c ? 3ll * c->nB : 0; }

long long fb_nnz_system(const fb_context *cc) {
  fb_context *c = const_cast<fb_context *>(cc);
  if (!c) return 0;
  if (cudaSetDevice(c->device) != cudaSuccess) return -1;
  std::vector<int> bp, bc, m;
  if (fetch_structure(c, bp, bc) != FB_OK) return -1;
  old_to_new(c, m);
  long long nnz = 0;
  for (int v = 0; v < c->nV; v++) {
    int rows = 0;
    for (int k = 0; k < 3; k++) rows += m[3 * (size_t)v + k] >= 0;
    if (!rows) continue;
    long long cols = 0;
    for (int p = bp[v]; p < bp[v + 1]; p++)
      for (int l = 0; l < 3; l++) cols += m[3 * (size_t)bc[p] + l] >= 0;
    nnz += rows * cols;
  }
  return nnz;
}

// ---- forces / state -------------------------------------------------------------------------------
int fb_set_external_forces(fb_context *c, const double *f) {
  CHECK_CTX(c);
  if (!f) { fb_set_error("f is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  if (c->dist) return fb_dist_upload_global(c, f, c->fext);
  return upload(c, c->fext, f, sizeof(double) * (size_t)c->r);
}
int fb_set_external_forces_dev(fb_context *c, const double *f) {
  CHECK_CTX(c);
  if (!f) { fb_set_error("f is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  FB_CUDA(cudaMemcpyAsync(c->fext, f, sizeof(double) * (size_t)c->r, cudaMemcpyDeviceToDevice, c->stream));
  return FB_OK;
}
int fb_add_external_forces(fb_context *c, const double *f) {
  CHECK_CTX(c);
  if (!f) { fb_set_error("f is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  if (c->dist) FB_TRY(fb_dist_upload_global(c, f, c->tmp));
  else FB_TRY(upload(c, c->tmp, f, sizeof(double) * (size_t)c->r));
  if (c->r) { k_axpy<<<gridFor(c->r, 256), 256, 0, c->stream>>>(c->r, 1.0, c->tmp, c->fext); c->launches++; }
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}
int fb_set_external_forces_to_zero(fb_context *c) {
  CHECK_CTX(c);
  FB_CUDA(cudaMemsetAsync(c->fext, 0, sizeof(double) * (size_t)(c->r ? c->r : 1), c->stream));
  return FB_OK;
}
int fb_get_external_forces(fb_context *c, double *f) {
  CHECK_CTX(c);
  if (!f) { fb_set_error("f is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  if (c->dist) return fb_dist_download_owned(c, c->fext, f);
  return download(c, f, c->fext, sizeof(double) * (size_t)c->r);
}
int fb_set_state(fb_context *c, const double *q, const double *qvel, const double *qaccel) {
  CHECK_CTX(c);
  if (!q) { fb_set_error("q is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  if (c->dist) {
    FB_TRY(fb_dist_upload_global(c, q, c->q));
    if (qvel) FB_TRY(fb_dist_upload_global(c, qvel, c->qvel));
    if (qaccel) FB_TRY(fb_dist_upload_global(c, qaccel, c->qaccel));
    return FB_OK;
  }
  size_t bytes = sizeof(double) * (size_t)c->r;
  FB_TRY(upload(c, c->q, q, bytes));
  if (qvel) FB_TRY(upload(c, c->qvel, qvel, bytes));
  if (qaccel) FB_TRY(upload(c, c->qaccel, qaccel, bytes));
  return FB_OK;
}
int fb_get_state(fb_context *c, double *q, double *qvel, double *qaccel) {
  CHECK_CTX(c);
  if (c->dist) {
    // global-length outputs: this rank's owned entries, zeros elsewhere (sum over ranks = the full vector)
    if (q) FB_TRY(fb_dist_download_owned(c, c->q, q));
    if (qvel) FB_TRY(fb_dist_download_owned(c, c->qvel, qvel));
    if (qaccel) FB_TRY(fb_dist_download_owned(c, c->qaccel, qaccel));
    return FB_OK;
  }
  size_t bytes = sizeof(double) * (size_t)c->r;
  if (bytes == 0) return FB_OK;
  if (q) FB_CUDA(cudaMemcpyAsync(q, c->q, bytes, cudaMemcpyDeviceToHost, c->stream));
  if (qvel) FB_CUDA(cudaMemcpyAsync(qvel, c->qvel, bytes, cudaMemcpyDeviceToHost, c->stream));
  if (qaccel) FB_CUDA(cudaMemcpyAsync(qaccel, c->qaccel, bytes, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}
int fb_get_state_dev(fb_context *c, double *q, double *qvel, double *qaccel) {
  CHECK_CTX(c);
  size_t bytes = sizeof(double) * (size_t)c->r;
  if (bytes == 0) return FB_OK;
  if (q) FB_CUDA(cudaMemcpyAsync(q, c->q, bytes, cudaMemcpyDeviceToDevice, c->stream));
  if (qvel) FB_CUDA(cudaMemcpyAsync(qvel, c->qvel, bytes, cudaMemcpyDeviceToDevice, c->stream));
  if (qaccel) FB_CUDA(cudaMemcpyAsync(qaccel, c->qaccel, bytes, cudaMemcpyDeviceToDevice, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}
const double *fb_displacements_dev(const fb_context *c) { return c ? c->q : nullptr; }
int fb_reset_to_rest(fb_context *c) {
  CHECK_CTX(c);
  size_t bytes = sizeof(double) * (size_t)(c->r ? c->r : 1);
  FB_CUDA(cudaMemsetAsync(c->q, 0, bytes, c->stream));
  FB_CUDA(cudaMemsetAsync(c->qvel, 0, bytes, c->stream));
  FB_CUDA(cudaMemsetAsync(c->qaccel, 0, bytes, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}
int fb_set_timestep(fb_context *c, double h) { if (!c) return FB_ERR_INVALID_ARGUMENT; c->prm.timestep = h; return FB_OK; }
int fb_set_damping(fb_context *c, double dm, double dk) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  c->prm.damping_mass = dm; c->prm.damping_stiffness = dk;
  return FB_OK;
}
int fb_set_warp(fb_context *c, int warp) {
  if (!c || warp < 0 || warp > 2) { fb_set_error("fb_set_warp: warp must be 0 (linear), 1 (corotational) or 2 (exact tangent)"); return FB_ERR_INVALID_ARGUMENT; }
  if (warp != 1 && c->mg && fb_mg_active(c)) { fb_set_error("fb_set_warp: the solver variants re-assemble their coarse levels with warp = 1"); return FB_ERR_NOT_SUPPORTED; }
  c->warp = warp;
  return FB_OK;
}
int fb_get_warp(const fb_context *c) { return c ? c->warp : -1; }
int fb_set_internal_force_scaling(fb_context *c, double s) { if (!c) return FB_ERR_INVALID_ARGUMENT; c->prm.internal_force_scaling = s; return FB_OK; }
int fb_set_cg(fb_context *c, double eps, int maxIt) {
  if (!c || maxIt < 0) return FB_ERR_INVALID_ARGUMENT;
  c->prm.cg_epsilon = eps; c->prm.cg_max_iterations = maxIt;
  return FB_OK;
}

// ---- the step --------------------------------------------------------------------------------------
int fb_step(fb_context *c) {
  CHECK_CTX(c);
  return fb_do_step(c);
}

// Many independent contexts, one step each, from a small pool of host threads.  Every context has its own stream, and fb_step
// blocks its caller until the step is done: stepped one after the other, the latency-bound phases of small meshes (a multigrid
// cycle on 200k tets is ~60 kernels of a few microseconds) leave the GPU mostly idle — 270 mesh-steps/s for 32 meshes; from 8
// threads their kernels interleave on the device: 715.  Results are those of n sequential fb_step calls (contexts share nothing).
int fb_step_many(fb_context *const *ctxs, int n, int host_threads, int *status) {
  if (n < 0 || (n > 0 && !ctxs)) { fb_set_error("fb_step_many: bad arguments"); return FB_ERR_INVALID_ARGUMENT; }
  for (int i = 0; i < n; i++)
    if (!ctxs[i]) { fb_set_error("fb_step_many: context %d is NULL", i); return FB_ERR_INVALID_ARGUMENT; }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < i; j++)
      if (ctxs[i] == ctxs[j]) { fb_set_error("fb_step_many: context %d listed twice", i); return FB_ERR_INVALID_ARGUMENT; }
  const int T = std::max(1, std::min(host_threads, n));
  std::vector<int> st((size_t)n, FB_OK);
  std::vector<std::string> msg((size_t)n);
  auto work = [&](int t) {
    for (int i = t; i < n; i += T) {
      st[(size_t)i] = fb_step(ctxs[i]);
      if (st[(size_t)i] != FB_OK) msg[(size_t)i] = fb_last_error_string();   // (the message buffer is per thread)
    }
  };
  if (T == 1) {
    work(0);
  } else {
    std::vector<std::thread> pool;
    for (int t = 1; t < T; t++) pool.emplace_back(work, t);
    work(0);
    for (std::thread &th : pool) th.join();
  }
  int first = FB_OK;
  for (int i = 0; i < n; i++) {
    if (status) status[i] = st[(size_t)i];
    if (first == FB_OK && st[(size_t)i] != FB_OK) { first = st[(size_t)i]; fb_set_error("context %d: %s", i, msg[(size_t)i].c_str()); }
  }
  return first;
}

// ---- statistics --------------------------------------------------------------------------------------
double fb_force_assembly_seconds(const fb_context *c) { return c ? 1e-3 * c->ms_assembly : 0.0; }
double fb_system_solve_seconds(const fb_context *c) { return c ? 1e-3 * c->ms_solve : 0.0; }
double fb_step_seconds(const fb_context *c) { return c ? 1e-3 * c->ms_step : 0.0; }
int fb_last_cg_iterations(const fb_context *c) { return c ? c->last_iters : 0; }
double fb_last_cg_residual_ratio(const fb_context *c) { return c ? c->last_ratio : 0.0; }
long long fb_kernel_launches(const fb_context *c) { return c ? c->launches : 0; }
size_t fb_device_bytes(const fb_context *c) { return c ? c->bytes : 0; }

// ---- inspection ----------------------------------------------------------------------------------------
int fb_get_stiffness_csr(fb_context *c, int *ia, int *ja) {
  CHECK_CTX(c);
  std::vector<int> bp, bc;
  FB_TRY(fetch_structure(c, bp, bc));
  for (int v = 0; v < c->nV; v++) {
    const int nb = bp[v + 1] - bp[v];
    for (int k = 0; k < 3; k++) {
      const long long start = 9ll * bp[v] + 3ll * nb * k;
      if (ia) ia[3 * (size_t)v + k] = (int)start;
      if (ja)
        for (int j = 0; j < nb; j++)
          for (int l = 0; l < 3; l++) ja[start + 3 * j + l] = 3 * bc[bp[v] + j] + l;
    }
  }
  if (ia) ia[c->r] = (int)c->nnzK;
  return FB_OK;
}

int fb_get_mass_csr(fb_context *c, int *ia, int *ja, double *a) {
  CHECK_CTX(c);
  std::vector<int> bp, bc;
  FB_TRY(fetch_structure(c, bp, bc));
  std::vector<double> mb((size_t)c->nB);
  FB_TRY(download(c, mb.data(), c->mblk, sizeof(double) * mb.size()));
  for (int v = 0; v < c->nV; v++) {
    const int nb = bp[v + 1] - bp[v];
    for (int k = 0; k < 3; k++) {
      const long long start = 3ll * bp[v] + (long long)nb * k;
      if (ia) ia[3 * (size_t)v + k] = (int)start;
      for (int j = 0; j < nb; j++) {
        if (ja) ja[start + j] = 3 * bc[bp[v] + j] + k;
        if (a) a[start + j] = mb[(size_t)bp[v] + j];
      }
    }
  }
  if (ia) ia[c->r] = 3 * c->nB;
  return FB_OK;
}

int fb_get_submatrix_map(fb_context *c, int *idx) {
  CHECK_CTX(c);
  if (!idx) return FB_ERR_INVALID_ARGUMENT;
  std::vector<int> bp, bc;
  FB_TRY(fetch_structure(c, bp, bc));
  size_t n = 0;
  for (int v = 0; v < c->nV; v++) {
    const int nb = bp[v + 1] - bp[v];
    for (int k = 0; k < 3; k++)
      for (int j = 0; j < nb; j++) idx[n++] = 3 * j + k;
  }
  return FB_OK;
}

static int system_structure(fb_context *c, int *ia, int *ja, double *a, int *superRows, int *superIdx) {
  std::vector<int> bp, bc, m;
  FB_TRY(fetch_structure(c, bp, bc));
  old_to_new(c, m);
  std::vector<double> ke;
  if (a) {
    ke.resize((size_t)c->nnzK);
    FB_TRY(download(c, ke.data(), c->Keff, sizeof(double) * ke.size()));
  }
  long long nnz = 0;
  int row = 0;
  for (int v = 0; v < c->nV; v++) {
    const int nb = bp[v + 1] - bp[v];
    for (int k = 0; k < 3; k++) {
      const int i = 3 * v + k;
      if (m[i] < 0) continue;
      if (ia) ia[row] = (int)nnz;
      if (superRows) superRows[row] = i;
      const long long start = 9ll * bp[v] + 3ll * nb * k;
      for (int j = 0; j < nb; j++)
        for (int l = 0; l < 3; l++) {
          const int nc = m[3 * (size_t)bc[bp[v] + j] + l];
          if (nc < 0) continue;
          if (ja) ja[nnz] = nc;
          if (a) a[nnz] = ke[start + 3 * j + l];
          if (superIdx) superIdx[nnz] = 3 * j + l;
          nnz++;
        }
      row++;
    }
  }
  if (ia) ia[row] = (int)nnz;
  return FB_OK;
}

int fb_get_system_csr(fb_context *c, int *ia, int *ja, double *a) {
  CHECK_CTX(c);
  return system_structure(c, ia, ja, a, nullptr, nullptr);
}
int fb_get_super_maps(fb_context *c, int *superRows, int *superIdx) {
  CHECK_CTX(c);
  return system_structure(c, nullptr, nullptr, nullptr, superRows, superIdx);
}
int fb_get_constrained_dofs(fb_context *c, int *dofs) {
  if (!c || (!dofs && c->nC > 0)) return FB_ERR_INVALID_ARGUMENT;   // an empty list needs no buffer (a rank without fixed vertices)
  if (c->nC == 0) return FB_OK;
  memcpy(dofs, c->cdofs_host, sizeof(int) * (size_t)c->nC);
  return FB_OK;
}

int fb_get_element_maps(fb_context *c, int *row4, int *col16) {
  CHECK_CTX(c);
  if (row4) FB_TRY(download(c, row4, c->tets, sizeof(int) * 4 * (size_t)c->nT));
  if (col16) FB_TRY(download(c, col16, c->colIdx, sizeof(int) * 16 * (size_t)c->nT));
  return FB_OK;
}

int fb_get_element_data(fb_context *c, double *minv16, double *k0) {
  CHECK_CTX(c);
  const int CH = 1 << 16;
  double *dm = nullptr, *dk = nullptr;
  if (minv16) FB_CUDA(cudaMalloc(&dm, sizeof(double) * 16 * (size_t)CH));
  if (k0) FB_CUDA(cudaMalloc(&dk, sizeof(double) * 144 * (size_t)CH));
  int st = FB_OK;
  for (int el0 = 0; el0 < c->nT && st == FB_OK; el0 += CH) {
    const int n = std::min(CH, c->nT - el0);
    st = fb_launch_expand_element(c, dm, dk, el0, n);
    if (st == FB_OK && minv16) st = download(c, minv16 + 16 * (size_t)el0, dm, sizeof(double) * 16 * (size_t)n);
    if (st == FB_OK && k0) st = download(c, k0 + 144 * (size_t)el0, dk, sizeof(double) * 144 * (size_t)n);
  }
  fb_dev_free(dm); fb_dev_free(dk);
  return st;
}

int fb_compute_force_and_matrix(fb_context *c, const double *u, double *f, double *Ka) {
  CHECK_CTX(c);
  if (!u) { fb_set_error("u is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  if (!c->Kraw) FB_TRY(fb_dev_alloc(c, &c->Kraw, (size_t)c->nnzK));
  FB_TRY(upload(c, c->Ad, u, sizeof(double) * (size_t)c->r));
  // internal forces land in c->fint like in the reference integrator's buffer; save and restore it so that
  // the call does not disturb the state of the last step
  FB_CUDA(cudaMemcpyAsync(c->tmp, c->fint, sizeof(double) * (size_t)c->r, cudaMemcpyDeviceToDevice, c->stream));
  FB_TRY(fb_launch_assembly(c, c->Ad, c->Kraw, false));
  if (f) FB_TRY(download(c, f, c->fint, sizeof(double) * (size_t)c->r));
  FB_CUDA(cudaMemcpyAsync(c->fint, c->tmp, sizeof(double) * (size_t)c->r, cudaMemcpyDeviceToDevice, c->stream));
  if (Ka) FB_TRY(download(c, Ka, c->Kraw, sizeof(double) * (size_t)c->nnzK));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

int fb_get_effective_stiffness_values(fb_context *c, double *a) {
  CHECK_CTX(c);
  if (!a) return FB_ERR_INVALID_ARGUMENT;
  return download(c, a, c->Keff, sizeof(double) * (size_t)c->nnzK);
}
int fb_get_rhs(fb_context *c, double *bc_) {
  CHECK_CTX(c);
  if (!bc_) return FB_ERR_INVALID_ARGUMENT;
  std::vector<double> full((size_t)c->r);
  FB_TRY(download(c, full.data(), c->rhs, sizeof(double) * full.size()));
  compress_constrained(c, full.data(), bc_);
  return FB_OK;
}
int fb_get_internal_forces(fb_context *c, double *f) {
  CHECK_CTX(c);
  if (!f) return FB_ERR_INVALID_ARGUMENT;
  return download(c, f, c->fint, sizeof(double) * (size_t)c->r);
}
int fb_get_qdelta(fb_context *c, double *d) {
  CHECK_CTX(c);
  if (!d) return FB_ERR_INVALID_ARGUMENT;
  return download(c, d, c->x, sizeof(double) * (size_t)c->r);  // InsertRows(buffer -> qdelta): zeros at constrained DOFs
}

int fb_solve(fb_context *c, const double *b, double *x, double eps, int maxIt, int *iterations) {
  CHECK_CTX(c);
  if (!x || maxIt < 0) return FB_ERR_INVALID_ARGUMENT;
  const size_t bytes = sizeof(double) * (size_t)c->r;
  if (b) {
    std::vector<double> full;
    expand_constrained(c, b, full);
    FB_CUDA(cudaMemcpyAsync(c->tmp, c->rhs, bytes, cudaMemcpyDeviceToDevice, c->stream));  // keep the step's rhs
    FB_TRY(upload(c, c->rhs, full.data(), bytes));
  }
  int st = fb_pcg_solve(c, eps, maxIt);
  if (b) cudaMemcpyAsync(c->rhs, c->tmp, bytes, cudaMemcpyDeviceToDevice, c->stream);
  if (st != FB_OK) return st;
  std::vector<double> full((size_t)c->r);
  FB_TRY(download(c, full.data(), c->x, bytes));
  compress_constrained(c, full.data(), x);
  if (iterations) *iterations = c->last_iters;
  return FB_OK;
}

int fb_system_multiply(fb_context *c, const double *x, double *y) {
  CHECK_CTX(c);
  if (!x || !y) return FB_ERR_INVALID_ARGUMENT;
  std::vector<double> full;
  expand_constrained(c, x, full);
  FB_TRY(upload(c, c->dir, full.data(), sizeof(double) * full.size()));
  FB_TRY(fb_launch_spmv(c, c->Keff, c->dir, c->Ad, false));
  FB_TRY(download(c, full.data(), c->Ad, sizeof(double) * full.size()));
  compress_constrained(c, full.data(), y);
  return FB_OK;
}

// ---- timers / profiling ------------------------------------------------------------------------------------
int fb_timer_start(fb_context *c) {
  CHECK_CTX(c);
  FB_CUDA(cudaEventRecord(c->ev[5], c->stream));
  return FB_OK;
}
int fb_timer_stop(fb_context *c, double *seconds) {
  CHECK_CTX(c);
  if (!seconds) return FB_ERR_INVALID_ARGUMENT;
  FB_CUDA(cudaEventRecord(c->ev[6], c->stream));
  FB_CUDA(cudaEventSynchronize(c->ev[6]));
  float ms = 0;
  FB_CUDA(cudaEventElapsedTime(&ms, c->ev[5], c->ev[6]));
  *seconds = 1e-3 * ms;
  return FB_OK;
}
int fb_set_profiling(fb_context *c, int enabled) {
  CHECK_CTX(c);
  if (enabled && !c->evProf[0])
    for (auto &e : c->evProf) FB_CUDA(cudaEventCreate(&e));
  c->profiling = enabled != 0;
  c->prof_sum_s = 0.0;
  c->prof_samples = 0;
  if (c->pers_prof) FB_CUDA(cudaMemsetAsync(c->pers_prof, 0, 2 * sizeof(unsigned long long), c->stream));
  return FB_OK;
}
int fb_get_spmv_profile(fb_context *c, double *mean, int *samples, double *bytes) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  if (c->pers_grid > 0 && c->pers_prof && !c->dist) {
    // persistent kernel: %globaltimer around the SpMV phase including its grid barrier, sampled by CTA 0
    unsigned long long h[2] = {0, 0};
    cudaSetDevice(c->device);
    FB_CUDA(cudaMemcpyAsync(h, c->pers_prof, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    if (mean) *mean = h[1] ? 1e-9 * (double)h[0] / (double)h[1] : 0.0;
    if (samples) *samples = (int)h[1];
  } else {
    if (mean) *mean = c->prof_samples ? c->prof_sum_s / c->prof_samples : 0.0;
    if (samples) *samples = c->prof_samples;
  }
  // 8 B value per scalar nonzero + 4 B block column per 3x3 block + per block row: 4 B row pointer,
  // 24 B of x (compulsory read), 24 B of y (write) [+ 3 B mask + 24 B d re-read for the fused dot]
  if (bytes) *bytes = 8.0 * (double)c->nnzK + 4.0 * (double)c->nB + 52.0 * (double)c->nV;
  if (bytes && c->sym_want && c->sym) *bytes = (double)fb_sym_bytes_per_product(c);  // upper triangle + lower-block records
  return FB_OK;
}

// ---- micro-benchmarks -------------------------------------------------------------------------------------
int fb_bench_spmv(fb_context *c, int repeats, double *sec) {
  CHECK_CTX(c);
  if (!sec || repeats <= 0) return FB_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < 3; i++) FB_TRY(fb_launch_spmv(c, c->Keff, c->dir, c->Ad, false));
  FB_CUDA(cudaEventRecord(c->ev[3], c->stream));
  for (int i = 0; i < repeats; i++) FB_TRY(fb_launch_spmv(c, c->Keff, c->dir, c->Ad, false));
  FB_CUDA(cudaEventRecord(c->ev[7], c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  FB_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[7]));
  *sec = 1e-3 * ms / repeats;
  return FB_OK;
}
int fb_bench_assembly(fb_context *c, int repeats, double *sec) {
  CHECK_CTX(c);
  if (!sec || repeats <= 0) return FB_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < 2; i++) FB_TRY(fb_launch_assembly(c, c->q, nullptr, true));
  FB_CUDA(cudaEventRecord(c->ev[3], c->stream));
  for (int i = 0; i < repeats; i++) FB_TRY(fb_launch_assembly(c, c->q, nullptr, true));
  FB_CUDA(cudaEventRecord(c->ev[7], c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  float ms = 0;
  FB_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[7]));
  *sec = 1e-3 * ms / repeats;
  return FB_OK;
}
}  // extern "C"
// ---- guard bands -------------------------------------------------------------------------------------------------
#include <mutex>
#include <unordered_map>
namespace {
constexpr size_t FB_GUARD = 256;
struct GuardRec { unsigned char *base; size_t bytes; };
std::mutex g_guard_mu;
std::unordered_map<void *, GuardRec> g_guards;
}  // namespace
bool fb_guard_enabled() {
  static const bool on = getenv("FEMBRAIN_B200_GUARD") && atoi(getenv("FEMBRAIN_B200_GUARD")) != 0;
  return on;
}
cudaError_t fb_guard_alloc(void **p, size_t bytes, cudaStream_t st) {
  unsigned char *base = nullptr;
  const size_t padded = ((bytes + 255) / 256) * 256;   // the band after the buffer starts at the next 256-byte boundary
  cudaError_t e = cudaMallocAsync((void **)&base, padded + 2 * FB_GUARD, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(base, 0xA5, padded + 2 * FB_GUARD, st);
  if (e != cudaSuccess) return e;
  *p = base + FB_GUARD;
  std::lock_guard<std::mutex> lk(g_guard_mu);
  g_guards[*p] = GuardRec{base, bytes};
  return cudaSuccess;
}
void fb_dev_free(void *p) {
  if (!p) return;
  if (fb_guard_enabled()) {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    auto it = g_guards.find(p);
    if (it != g_guards.end()) {
      cudaFree(it->second.base);
      g_guards.erase(it);
      return;
    }
  }
  cudaFree(p);
}
extern "C" int fb_check_guards(long long *checked, long long *corrupted) {
  // bytes [bytes, padded) behind a buffer are slack the kernels may legitimately touch (e.g. whole 16-byte lines around a
  // row tile); the bands proper must still hold the fill pattern
  if (!checked || !corrupted) return FB_ERR_INVALID_ARGUMENT;
  *checked = *corrupted = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) { cudaGetLastError(); return FB_ERR_CUDA; }
  std::lock_guard<std::mutex> lk(g_guard_mu);
  unsigned char host[2 * FB_GUARD];
  for (auto &kv : g_guards) {
    const GuardRec &g = kv.second;
    const size_t padded = ((g.bytes + 255) / 256) * 256;
    if (cudaMemcpy(host, g.base, FB_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(host + FB_GUARD, g.base + FB_GUARD + padded, FB_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess) {
      cudaGetLastError();   // an allocation that lives on another device of the process: skipped
      continue;
    }
    bool bad = false;
    for (size_t i = 0; i < 2 * FB_GUARD; i++) bad |= host[i] != 0xA5;
    (*checked)++;
    if (bad) {
      (*corrupted)++;
      fb_set_error("guard band of a %zu-byte device allocation was overwritten", g.bytes);
    }
  }
  return FB_OK;
}
bool fb_use_pool() {
  static const bool on = !(getenv("FEMBRAIN_B200_POOL") && atoi(getenv("FEMBRAIN_B200_POOL")) == 0);
  return on;
}
extern "C" {
int fb_trim_memory(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); return FB_ERR_NO_DEVICE; }
  for (int d = 0; d < ndev; d++) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  }
  cudaGetLastError();
  return FB_OK;
}
int fb_bench_cg_iteration(fb_context *c, int repeats, double *sec) {
  CHECK_CTX(c);
  if (!sec || repeats <= 0) return FB_ERR_INVALID_ARGUMENT;
  return fb_pcg_bench_iteration(c, repeats, sec);
}

}  // extern "C"
