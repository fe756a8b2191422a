// fb_internal.h — context layout and kernel-launcher prototypes shared by the translation units of
// libfembrain_b200.so.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/fembrain_b200.h"

// ---- error plumbing -------------------------------------------------------------------------
void fb_set_error(const char *fmt, ...);
#define FB_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      fb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (e__ == cudaErrorMemoryAllocation) ? FB_ERR_OUT_OF_MEMORY : FB_ERR_CUDA;      \
    }                                                                                      \
  } while (0)
#define FB_TRY(call)             \
  do {                           \
    int s__ = (call);            \
    if (s__ != FB_OK) return s__; \
  } while (0)

// ---- device-resident scalars of the PCG loop --------------------------------------------------
// rho[it & 1] holds sum r_i^2 / diag_i after iteration `it` (it = 0: initial); the loop reads
// rho[(it-1)&1] and writes rho[it&1], so no kernel both reads and writes one slot.
struct FbScalars {
  double rho[2];
  double rho0;
  double dq;        // d . (A d) of the current iteration
  double rq, qq;    // (r, q)_D and (q, q)_D of the current iteration (fused schedule)
  double eps2;      // epsilon^2
  int max_it;
  int iters;        // iterations completed
  int done;         // loop condition of CGSolver.cpp:150 is false
  unsigned int ticket_a;  // last-block tickets (reset by the last block)
  unsigned int ticket_b;
  int pad[3];
  // partitioned contexts: this rank's partial sums; ncclAllReduce(part -> dq / rho[it&1]).  Kept apart from the
  // reduced values so that the all-reduces that still run after `done` cannot compound stale numbers.
  double dq_part, rho_part;
  int comm_error;   // a bounded peer wait ran out (partitioned contexts, peer-memory exchange)
  int pad2;
};

#define FB_MAX_PARTIALS 4096

struct FbDist;   // multi-GPU state (fb_dist.cu)
struct FbBatch;  // a batch of independent meshes in one context (fb_batch.cu)
struct FbSym;    // upper-triangle storage for the solver's products (fb_sym.cu)
struct FbTma;    // row tiles for the bulk-copy staged products (fb_tma.cu)
struct FbMg;     // labelled solver variants: block-Jacobi / multigrid preconditioned CG (fb_mg.cu)


struct fb_context {
  int device;
  int sm_count;
  cudaStream_t stream;
  fb_params prm;
  double lambda_uniform, mu_uniform;

  // sizes
  int nV, nT, r;
  int nB;         // 3x3 blocks of K (vertex-pair adjacency incl. self)
  int nC;         // constrained DOFs
  long long nnzK;  // 9 * nB
  size_t bytes;   // device bytes held
  size_t l2_pinned_bytes;  // bytes of Keff covered by the persisting-L2 access policy window

  // mesh
  double *x0;  // [3 nV] rest positions
  int *tets;   // [4 nT]
  // per-element data, structure-of-arrays planes of nT doubles each:
  // 0..11 G (upper 4x3 of MInverse), 12 volume, 13 lambda, 14 mu, 15 density
  double *edata;

  // block structure of K: row pointer, block columns, owning row of each block, diagonal block of each row
  int *bp;    // [nV+1]
  int *bc;    // [nB]
  int *brow;  // [nB]
  int *diag;  // [nV]
  // contributions: for block b, entries seg[b]..seg[b+1] of src, each el*16 + 4*i + j, ascending el
  int *seg;            // [nB+1]
  unsigned int *src;   // [16 nT]
  int *colIdx;         // [16 nT] element -> block position inside row (reference's columnIndices)
  double *mblk;        // [nB] mass scalar of each block (M = mblk (x) I3)
  unsigned char *fixed;    // [r] 1 = constrained DOF (state is pinned to zero there)
  unsigned char *rowmask;  // [r] rows the solver skips: == fixed, or fixed + ghost rows in partitioned contexts
  int *cdofs;            // [nC] sorted constrained DOFs (device)
  int *cdofs_host;

  // matrices in the reference's CSR value order: idx = 9*bp[v] + k*3*nb(v) + 3*jpos + l
  double *T;     // h*K + D   (the matrix DoTimestep multiplies qvel with)
  double *Keff;  // M + h*D + h^2*K
  double *Kraw;  // raw K, only when prm.keep_raw_stiffness or inspection asked for it
  // per-element scratch of the two-phase deterministic assembly
  double *scrK;  // [16][nT][9]
  double *scrF;  // [12][nT]
  // row-gather assembly (fb_assembly.cu; default): per-CTA index lists, element and vertex records.
  // ga_ctas == 0 -> two-phase path (scrK/scrF, allocated on first use).
  int ga_ctas, ga_cfg;
  int warp;                  // CorotationalLinearFEM's warp argument: 1 (default; the only one the gather path serves), 0 linear, 2 exact tangent
  int *ga_incp;              // [nV+1] incidence (vertex, element, i) counts, prefix sum
  int *ga_ctaV;              // [ga_ctas+1] first vertex of every CTA
  unsigned char *ga_lists;   // [ga_ctas] GaLists blobs
  double *ga_erec;           // [nT][24] element records: R[9] (rewritten every assembly), G[12], volume, lambda, mu
  double *ga_xu;             // [nV][6] rest position and displacement side by side (rewritten every assembly)

  // integrator state and work vectors, all [r]
  double *q, *qvel, *qaccel, *fext, *fint, *qres, *rhs, *x, *res, *dir, *Ad, *invD, *tmp;
  FbScalars *sc;        // device
  FbScalars *sc_host;   // pinned, 4 slots
  double *partials;     // [2][FB_MAX_PARTIALS]
  cudaEvent_t ev[8];
  cudaEvent_t evChunk[4];

  // Deformable-level options
  int gravity, floor_enabled, haptic_in_progress, haptic_rings, contact_count;
  double floor_y;
  int nHaptic;
  int *haptic_idx_host;
  double *haptic_f_host;
  int *adj_host_bp, *adj_host_bc;  // (unused since the ring spreading moved to the device; kept zero)
  int *haptic_idx_dev, *edges_dev, *edge_degree_dev;   // device copies for k_haptic_spread
  double *haptic_f_dev;
  unsigned int *haptic_stamp, haptic_stamp_base;        // visit stamps of the ring walk, never cleared between frames
  int *haptic_listA, *haptic_listB;
  int nEdges, *edges_host;         // optional reference edge array (VolMesh::m_vEdges order), 2 ints per edge
  int *edge_degree_host;           // incident-edge count per vertex of that array
  int haptic_quirk;                // replicate VolMesh::get_node_neighbors (DEF/VolMesh.cpp:1346-1363) exactly
  double *fext_host;
  int *contact_dev;

  // stats
  float ms_assembly, ms_solve, ms_step;
  int last_iters;        // signed like the reference's return value
  double last_ratio;
  long long launches;
  int row_lo, row_hi;    // block rows the solver's products visit: all rows, or the OWNED rows of a partitioned context
  int spmv_group;        // lanes per block row chosen at setup
  int use_rows3;         // 1: k_spmv_rows3 (16 lanes per row, all loads of a row in flight), 0: k_spmv<G>
  int rows3_minb;        // resident CTAs/SM requested for k_spmv_rows3 modes 0-2 (4 or 5)
  int grid_spmv[4], grid_vec;  // one-resident-wave launch shapes per SpMV mode and for the vector kernels
  int pdl;                     // launch the PCG kernels with programmatic dependent launch
  int pcg_fused, pcg_graph;    // two-kernel schedule / CUDA-graph replay of a 30-iteration period
  void *graph_exec;            // cudaGraphExec_t
  int graph_kernels, graph_failed;
  // persistent cooperative PCG kernel (fb_pcg_persistent.cu): contiguous row range per CTA, equal block counts
  int *ctaRows;          // [pers_grid + 1] device
  int pers_grid;         // 0 = use the three-kernels-per-iteration path
  unsigned long long *pers_prof;  // device [2]: summed ns of sampled SpMV phases, sample count
  // optional sampling of SpMV launch durations inside fb_step
  int profiling, nprof;
  cudaEvent_t evProf[128];
  double prof_sum_s;
  int prof_samples;

  int comm_poisoned;     // last solve returned FB_ERR_COMM: reset the last-block tickets before the next one
  FbDist *dist;
  FbMg *mg;
  int stream_borrowed;   // the stream belongs to another context (coarse levels of a multigrid hierarchy)
  int uniform_material;  // created without per-element E / nu / density arrays
  int have_solution;     // c->x holds the solution of a converged solve (warm start of the solver variants)
  FbBatch *batch;
  FbSym *sym;
  FbTma *tma;
  int l2_evict;  // FEMBRAIN_B200_L2EVICT=1
  int tma_want;  // FEMBRAIN_B200_SPMV=tma: products with the matrix staged through shared memory by cp.async.bulk (experimental)
  int sym_want;  // FEMBRAIN_B200_SPMV=sym: products of the three-kernel schedule from the block-upper triangle (plan built at the first solve)
};

// ---- fb_api.cu -----------------------------------------------------------------------------------
int fb_do_step(fb_context *c);
int fb_create_local(fb_context **out, int nV, const double *x0, int nT, const int *tets, int nC, const int *cdofs,
                    const double *E, const double *nu, const double *rho, const fb_params *prm);
int fb_apply_constraints(fb_context *c, int nC, const int *cdofs_sorted);
#ifdef __cplusplus
#include <vector>
int fb_fetch_structure(fb_context *c, std::vector<int> &bp, std::vector<int> &bc);
#endif
// ---- fb_setup.cu ---------------------------------------------------------------------------------
int fb_build_topology(fb_context *c);
// ---- fb_fem.cu (compiled with -fmad=false) ---------------------------------------------------------
int fb_launch_element_data(fb_context *c, const double *E, const double *nu, const double *rho);
int fb_launch_mass(fb_context *c);
int fb_launch_assembly(fb_context *c, const double *u, double *Kraw, bool effective, bool rhs = false);
int fb_launch_rhs(fb_context *c);
int fb_launch_spmv_exact(fb_context *c, const double *A, const double *x, double *y);
int fb_launch_state_update(fb_context *c);
int fb_launch_expand_element(fb_context *c, double *minv16_dev, double *k0_dev, int el0, int n);
// ---- fb_assembly.cu (compiled with -fmad=false) -----------------------------------------------------
int fb_build_gather_plan(fb_context *c);  // after fb_build_topology
int fb_launch_assembly_gather(fb_context *c, const double *u, double *Kraw, bool effective, bool rhs);
// ---- fb_pcg.cu -------------------------------------------------------------------------------------
int fb_pcg_solve(fb_context *c, double eps, int max_it);  // solves Keff x = rhs (masked), x0 = 0
int fb_launch_spmv(fb_context *c, const double *A, const double *x, double *y, bool masked);
int fb_pcg_bench_iteration(fb_context *c, int repeats, double *sec);
int fb_spmv_plan(fb_context *c);  // call once after the block structure exists
void fb_pcg_release(fb_context *c);
int fb_pcg_launch_product_dq(fb_context *c, const double *d, double *q, int *nSlots);  // q = mask(Keff d), d.q partials in c->partials[0..nSlots)
int fb_pcg_launch_residual(fb_context *c, const double *x, double *r);               // r = mask(rhs - Keff x)
// ---- fb_mg.cu (labelled solver variants) ----------------------------------------------------------------------------------
int fb_mg_active(const fb_context *c);   // 0, or the FB_SOLVER_* variant in use
int fb_mg_prepare(fb_context *c);        // per step, after the assembly: coarse operators, FP32 copies, block inverses
int fb_mg_pcg_solve(fb_context *c, double eps, int max_it);
void fb_mg_invalidate(fb_context *c);    // constraints changed: the hierarchy is rebuilt at the next step
void fb_mg_release(fb_context *c);
// (fb_experiments_built() of the public header: 1 when csrc/experiments/ was compiled in (fb_experiments_on.cu), 0 for the default library)
// ---- experiments/fb_pcg_persistent.cu (stubs in fb_experiments_off.cu by default) ------------------
int fb_pcg_plan_persistent(fb_context *c);
int fb_pcg_launch_persistent(fb_context *c);
// ---- fb_batch.cu -----------------------------------------------------------------------------------
int fb_batch_pcg_solve(fb_context *c, double eps, int max_it);  // every mesh of the batch, own scalars and stopping rule each
void fb_batch_destroy(fb_context *c);
// ---- experiments/fb_sym.cu -------------------------------------------------------------------------------------
int fb_sym_plan(fb_context *c);   // FB_OK with c->sym == nullptr: not applicable, keep the full-matrix kernels
int fb_sym_pack(fb_context *c);   // U <- upper(Keff), start of every solve
void fb_sym_launch(fb_context *c, int mode, const double *x, double *y, const double *b, double *slots);
int fb_sym_grid(const fb_context *c, int mode);
size_t fb_sym_bytes_per_product(const fb_context *c);
void fb_sym_release(fb_context *c);
// ---- experiments/fb_tma.cu -------------------------------------------------------------------------------------
int fb_tma_plan(fb_context *c);   // FB_OK with c->tma == nullptr: not applicable, keep the default kernels
void fb_tma_launch(fb_context *c, int mode, const double *x, double *y, const double *b, double *slots);
int fb_tma_grid(const fb_context *c);
int fb_tma_failed(fb_context *c);
void fb_tma_release(fb_context *c);
// ---- fb_dist.cu ------------------------------------------------------------------------------------
int fb_dist_halo_exchange(fb_context *c, double *vec);
int fb_dist_allreduce_scalar(fb_context *c, const double *dev_part, double *dev_total);
void fb_dist_destroy(fb_context *c);
int fb_dist_refresh_rowmask(fb_context *c);  // rowmask = fixed + ghost rows (after constraints change)
// global-length host vector <-> this rank's local device vector
int fb_dist_upload_global(fb_context *c, const double *global_host, double *local_dev);
int fb_dist_download_owned(fb_context *c, const double *local_dev, double *global_host);
int fb_dist_global_sizes(const fb_context *c, int *nV, int *nT);
// peer-memory (CUDA IPC over NVLink) exchange: 1 when the context exchanges scalars and halos through peer stores
int fb_dist_p2p(const fb_context *c);
struct FbPeerArgs;
void fb_dist_peer_args(fb_context *c, FbPeerArgs *out);        // rank/world/comm pointers, epochs zeroed
unsigned long long fb_dist_epoch(fb_context *c, int it, int family);  // epoch of (current solve, iteration, family)
void fb_dist_next_solve(fb_context *c);
int fb_dist_reset_tickets(fb_context *c);
unsigned int fb_dist_halo_mask(const fb_context *c);
int fb_dist_halo_push(fb_context *c, const double *vec, unsigned long long epoch);
struct FbPushArgs;
void fb_dist_push_args(fb_context *c, FbPushArgs *out, unsigned long long epoch);  // halo push fused into k_direction

// Device memory comes from the device's default stream-ordered pool (cudaMallocAsync on the context's stream); the
// pool's release threshold is raised at the first fb_create, so destroy -> create cycles (re-setup after a cut,
// DEF/Deformable.cpp:127-220 via cutCompleted) reuse memory instead of paying cudaMalloc/cudaFree again
// (fb_create at 1M tets: 75-200 ms with cudaMalloc, profiles/r01_setup_time.txt).  fb_trim_memory() gives it back.
// Buffers exported with CUDA IPC (peer-memory exchange) cannot live in a pool: fb_dev_alloc_plain.
bool fb_use_pool();  // false with FEMBRAIN_B200_POOL=0: plain cudaMalloc for everything
// Guard bands (FEMBRAIN_B200_GUARD=1; tests/test_guard_gpu.py): every pool allocation of a context is wrapped in two
// 256-byte bands filled with 0xA5, and fb_check_guards() reports allocations whose bands were written to.  This is the
// library's own out-of-bounds-write detector: compute-sanitizer is closed on the GPU pool this was developed on.
bool fb_guard_enabled();
cudaError_t fb_guard_alloc(void **p, size_t bytes, cudaStream_t st);
void fb_dev_free(void *p);  // cudaFree, or the guarded allocation's base when p came from fb_guard_alloc
template <typename T>
static inline int fb_dev_alloc_plain(fb_context *c, T **p, size_t n) {
  size_t bytes = (n ? n : 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void **)p, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fb_set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    *p = nullptr;
    return FB_ERR_OUT_OF_MEMORY;
  }
  c->bytes += bytes;
  return FB_OK;
}
template <typename T>
static inline int fb_dev_alloc(fb_context *c, T **p, size_t n) {
  size_t bytes = (n ? n : 1) * sizeof(T);
  if (!fb_use_pool()) return fb_dev_alloc_plain(c, p, n);
  cudaError_t e = fb_guard_enabled() ? fb_guard_alloc((void **)p, bytes, c->stream) : cudaMallocAsync((void **)p, bytes, c->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fb_set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    *p = nullptr;
    return FB_ERR_OUT_OF_MEMORY;
  }
  c->bytes += bytes;
  return FB_OK;
}
// temporaries of the setup code: stream-ordered, freed with fb_tmp_free on the same stream
template <typename T>
static inline cudaError_t fb_tmp_alloc(cudaStream_t st, T **p, size_t bytes) { return cudaMallocAsync((void **)p, bytes ? bytes : 1, st); }
static inline void fb_tmp_free(cudaStream_t st, void *p) { if (p) cudaFreeAsync(p, st); }
