// fb_batch.cu — a batch of independent meshes in ONE context (BASELINE.json config 4: many small brain meshes per GPU).
//
// Reference: every mesh is its own Deformable with its own integrator and CGSolver (src/deformable/Deformable.cpp:127-220,
// src/3rdparty/vegafem/sparseSolver/CGSolver.cpp:129-190); meshes never interact.  Stepping them as separate contexts
// from several host threads leaves a B200 mostly idle: a 200k-tet mesh's kernels run 10 us each and an iteration is
// three dependent launches, so 32 meshes on 8 streams reach 81 mesh-steps/s against 76 for a single stream
// (profiles/r01_batch_graph.txt).  Here the batch is ONE block-diagonal system:
//   * setup, element kernels, assembly, rhs and state update are the ordinary single-mesh code on the concatenated mesh
//     (disconnected components change nothing there; every value stays bit-identical to the per-mesh context);
//   * PCG keeps one set of scalars PER MESH (rho, rho0, alpha, beta, iteration count, loop flag): each mesh runs the
//     reference's recurrences, refresh period and stopping rule on its own and stops on its own; the kernels just cover
//     all meshes that are still iterating in one launch.
// Rows are cut into chunks of BT_ROWS block rows that never straddle two meshes; one CTA per chunk in every kernel.  Dot
// products: per-chunk sums in fixed slots, added per mesh in slot order by every consumer CTA of that mesh — deterministic,
// no atomics on floating-point data.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fb_internal.h"
#include "fb_pcg_common.cuh"

struct FbBatch {
  int count;                 // meshes
  std::vector<int> vtx;      // [count+1] first vertex of every mesh in the concatenated numbering
  std::vector<int> tet;      // [count+1] first tet
  int nChunks;
  // device
  int *chunkRow;             // [nChunks+1] block-row boundaries
  int *chunkMesh;            // [nChunks]
  int *meshChunk;            // [count+1] chunk range of every mesh
  double *rho;               // [2][count]
  double *rho0;              // [count]
  int *iters;                // [count] iterations completed
  int *done;                 // [count]
  unsigned int *ticket;      // [count] chunks of the mesh that have finished the direction update
  int *active;               // [1] meshes still iterating
  double *slotsA, *slotsB;   // [nChunks] per-chunk sums (d.q / sum r^2 invD)
  std::vector<int> itersHost;
  std::vector<double> ratioHost;
};

namespace {

constexpr int BT_TB = 256;
constexpr int BT_ROWS = 128;  // block rows per chunk: 8 rows per 16-lane group

struct BatchArgs {
  const int *chunkRow, *chunkMesh, *meshChunk;
  double *rho, *rho0;
  int *iters, *done, *active;
  unsigned int *ticket;
  double *slotsA, *slotsB;
  int count, maxIt;
  double eps2;
};

// rows [r0, r1) of y = mask(A x) (MODE 1, returns sum x.y) or y = mask(b - A x) (MODE 2, returns sum y^2 invD); 16 lanes per
// block row, all loads of a row in flight (the loop body of k_spmv_rows3)
template <int MODE>
__device__ __forceinline__ double batch_rows(int r0, int r1, const int *__restrict__ bp, const int *__restrict__ bc,
                                             const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
                                             const unsigned char *__restrict__ mask, const double *__restrict__ b,
                                             const double *__restrict__ invD) {
  const int lane = threadIdx.x & (TILE_G - 1);
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int groups = BT_TB / TILE_G;
  double part = 0.0;
  for (int v = r0 + threadIdx.x / TILE_G; v < r1; v += groups) {
    const int rs = __ldg(bp + v), re = __ldg(bp + v + 1);
    const int n3 = 3 * (re - rs);
    const size_t row = 3 * (size_t)v + (lane < 3 ? lane : 0);
    double xr = 0.0, br = 0.0, wr = 0.0;
    unsigned char mk = 0;
    if (lane < 3) {
      mk = __ldg(mask + row);
      if (MODE == 1) xr = __ldg(x + row);
      if (MODE == 2) { br = __ldg(b + row); wr = __ldg(invD + row); }
    }
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int base = 0; base < n3; base += TILE_CHUNK) {
      RowVals val;
      int col[3];
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int t = base + lane + TILE_G * p;
        col[p] = (t < n3) ? __ldg(bc + rs + t / 3) : -1;
      }
      load_row_chunk(A, rs, n3, base, lane, val);
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int t = base + lane + TILE_G * p;
        const double xv = (col[p] >= 0) ? __ldg(x + 3 * (size_t)col[p] + (t % 3)) : 0.0;
        acc0 = fma(val.v[p][0], xv, acc0); acc1 = fma(val.v[p][1], xv, acc1); acc2 = fma(val.v[p][2], xv, acc2);
      }
    }
#pragma unroll
    for (int o = TILE_G / 2; o > 0; o >>= 1) {
      acc0 += __shfl_xor_sync(gmask, acc0, o, TILE_G);
      acc1 += __shfl_xor_sync(gmask, acc1, o, TILE_G);
      acc2 += __shfl_xor_sync(gmask, acc2, o, TILE_G);
    }
    if (lane < 3) {
      double s = (lane == 0) ? acc0 : ((lane == 1) ? acc1 : acc2);
      if (MODE == 1) {
        if (mk) s = 0.0;
        y[row] = s;
        part = fma(xr, s, part);
      } else {
        const double rres = mk ? 0.0 : (br - s);
        y[row] = rres;
        part += (rres * rres) * wr;
      }
    }
  }
  return part;
}

// q = A d with per-chunk d.q (MODE 1) / r = b - A x with per-chunk sum r^2 invD (MODE 2)
template <int MODE>
__global__ void __launch_bounds__(BT_TB) kb_spmv(BatchArgs a, const int *__restrict__ bp, const int *__restrict__ bc,
                                                 const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
                                                 const unsigned char *__restrict__ mask, const double *__restrict__ b,
                                                 const double *__restrict__ invD) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x;
  if (a.done[a.chunkMesh[ch]]) return;
  const double part = batch_rows<MODE>(a.chunkRow[ch], a.chunkRow[ch + 1], bp, bc, A, x, y, mask, b, invD);
  block_reduce_to_slot<BT_TB>(part, MODE == 1 ? a.slotsA : a.slotsB);  // slot [blockIdx.x] = [ch]
}

// r = b (x0 = 0), d = invD r, x = 0, q = 0, per-chunk sum r^2 invD                                  (CGSolver.cpp:139-147)
__global__ void __launch_bounds__(BT_TB) kb_init(BatchArgs a, const double *__restrict__ b, const double *__restrict__ invD,
                                                 double *__restrict__ x, double *__restrict__ r, double *__restrict__ d,
                                                 double *__restrict__ q) {
  const int ch = blockIdx.x;
  double part = 0.0;
  for (size_t i = 3 * (size_t)a.chunkRow[ch] + threadIdx.x; i < 3 * (size_t)a.chunkRow[ch + 1]; i += BT_TB) {
    const double bi = b[i], di = invD[i];
    x[i] = 0.0; r[i] = bi; q[i] = 0.0; d[i] = di * bi;
    part += (bi * bi) * di;
  }
  block_reduce_to_slot<BT_TB>(part, a.slotsB);
}

// per mesh: rho0, loop condition at iteration 1                                                     (CGSolver.cpp:147-150)
__global__ void kb_begin(BatchArgs a) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= a.count) return;
  double total = 0.0;
  for (int s = a.meshChunk[m]; s < a.meshChunk[m + 1]; s++) total += a.slotsB[s];
  a.rho[m] = total;       // rho[0][m]
  a.rho0[m] = total;
  a.iters[m] = 0;
  a.ticket[m] = 0u;
  const int fin = !((total > a.eps2 * total) && (1 <= a.maxIt));
  a.done[m] = fin;
  if (fin) atomicSub(a.active, 1);
}

// x += alpha d; REFRESH ? nothing more : (r -= alpha q; per-chunk sum r^2 invD)                     (CGSolver.cpp:155-174)
template <bool REFRESH>
__global__ void __launch_bounds__(BT_TB) kb_update(BatchArgs a, const double *__restrict__ d, const double *__restrict__ q,
                                                   const double *__restrict__ invD, double *__restrict__ x,
                                                   double *__restrict__ r) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x, m = a.chunkMesh[ch];
  if (a.done[m]) return;
  const int it = a.iters[m] + 1;
  const double dq = cta_sum_slots<BT_TB>(a.slotsA + a.meshChunk[m], a.meshChunk[m + 1] - a.meshChunk[m]);
  const double alpha = a.rho[((it - 1) & 1) * a.count + m] / dq;
  double part = 0.0;
  for (size_t i = 3 * (size_t)a.chunkRow[ch] + threadIdx.x; i < 3 * (size_t)a.chunkRow[ch + 1]; i += BT_TB) {
    x[i] = fma(alpha, d[i], x[i]);
    if (!REFRESH) {
      const double ri = fma(-alpha, q[i], r[i]);
      r[i] = ri;
      part += (ri * ri) * invD[i];
    }
  }
  if (!REFRESH) block_reduce_to_slot<BT_TB>(part, a.slotsB);
}

// beta = rho'/rho; d = invD r + beta d; per mesh: iteration++ and the loop condition                (CGSolver.cpp:176-183, 150)
__global__ void __launch_bounds__(BT_TB) kb_direction(BatchArgs a, const double *__restrict__ r, const double *__restrict__ invD,
                                                      double *__restrict__ d) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x, m = a.chunkMesh[ch];
  if (a.done[m]) return;
  const int it = a.iters[m] + 1;
  const int nch = a.meshChunk[m + 1] - a.meshChunk[m];
  const double rhoNew = cta_sum_slots<BT_TB>(a.slotsB + a.meshChunk[m], nch);
  const double rhoOld = a.rho[((it - 1) & 1) * a.count + m];
  const double beta = rhoNew / rhoOld;
  for (size_t i = 3 * (size_t)a.chunkRow[ch] + threadIdx.x; i < 3 * (size_t)a.chunkRow[ch + 1]; i += BT_TB)
    d[i] = fma(invD[i], r[i], beta * d[i]);
  // bookkeeping by the mesh's last chunk to finish, so that no chunk of this launch can still be reading iters / done
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(&a.ticket[m], 1u) == (unsigned)nch - 1u);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    a.ticket[m] = 0u;
    a.rho[(it & 1) * a.count + m] = rhoNew;
    a.iters[m] = it;
    if (!((rhoNew > a.eps2 * a.rho0[m]) && (it + 1 <= a.maxIt))) {
      a.done[m] = 1;
      atomicSub(a.active, 1);
    }
  }
}

void batch_args(const fb_context *c, double eps, int maxIt, BatchArgs *a) {
  const FbBatch *b = c->batch;
  a->chunkRow = b->chunkRow; a->chunkMesh = b->chunkMesh; a->meshChunk = b->meshChunk;
  a->rho = b->rho; a->rho0 = b->rho0; a->iters = b->iters; a->done = b->done; a->active = b->active; a->ticket = b->ticket;
  a->slotsA = b->slotsA; a->slotsB = b->slotsB;
  a->count = b->count; a->maxIt = maxIt; a->eps2 = eps * eps;
}

}  // namespace

void fb_batch_destroy(fb_context *c) {
  FbBatch *b = c->batch;
  if (!b) return;
  void *ptrs[] = {b->chunkRow, b->chunkMesh, b->meshChunk, b->rho, b->rho0, b->iters, b->done, b->ticket, b->active, b->slotsA, b->slotsB};
  for (void *p : ptrs)
    if (p) cudaFree(p);
  delete b;
  c->batch = nullptr;
}

// every mesh's Jacobi-PCG, x0 = 0, all meshes in every launch; c->last_iters = the largest iteration count, negative if
// some mesh did not converge (per mesh: fb_batch_last_cg_iterations)
int fb_batch_pcg_solve(fb_context *c, double eps, int maxIt) {
  FbBatch *b = c->batch;
  cudaStream_t st = c->stream;
  if (c->r == 0) { c->last_iters = 0; c->last_ratio = 0.0; return FB_OK; }
  BatchArgs a;
  batch_args(c, eps, maxIt, &a);
  const int grid = b->nChunks;
  FB_CUDA(cudaMemcpyAsync(b->active, &b->count, sizeof(int), cudaMemcpyHostToDevice, st));
  kb_init<<<grid, BT_TB, 0, st>>>(a, c->rhs, c->invD, c->x, c->res, c->dir, c->Ad);
  kb_begin<<<(b->count + 127) / 128, 128, 0, st>>>(a);
  c->launches += 2;
  const int CH = 30;
  int *activeHost[2] = {reinterpret_cast<int *>(&c->sc_host[0]), reinterpret_cast<int *>(&c->sc_host[1])};  // pinned
  int it = 1, slot = 0, pending = 0;
  bool finished = false;
  while (!finished && it <= maxIt) {
    const int end = (it + CH - 1 < maxIt) ? it + CH - 1 : maxIt;
    for (; it <= end; it++) {
      // a mesh that has stopped is skipped by every kernel; meshes that stop at different iterations keep their own
      // refresh phase because `it` is common to all meshes that are still iterating (all started at 1)
      fb_launch(c->pdl, st, kb_spmv<1>, grid, BT_TB, a, c->bp, c->bc, c->Keff, c->dir, c->Ad, c->rowmask, c->rhs, c->invD);
      if (it % 30 == 0) {
        fb_launch(c->pdl, st, kb_update<true>, grid, BT_TB, a, c->dir, c->Ad, c->invD, c->x, c->res);
        fb_launch(c->pdl, st, kb_spmv<2>, grid, BT_TB, a, c->bp, c->bc, c->Keff, c->x, c->res, c->rowmask, c->rhs, c->invD);
        c->launches++;
      } else {
        fb_launch(c->pdl, st, kb_update<false>, grid, BT_TB, a, c->dir, c->Ad, c->invD, c->x, c->res);
      }
      fb_launch(c->pdl, st, kb_direction, grid, BT_TB, a, c->res, c->invD, c->dir);
      c->launches += 3;
    }
    FB_CUDA(cudaMemcpyAsync(activeHost[slot], b->active, sizeof(int), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaEventRecord(c->evChunk[slot], st));
    pending++;
    if (pending == 2) {
      const int prev = slot ^ 1;
      FB_CUDA(cudaEventSynchronize(c->evChunk[prev]));
      if (*activeHost[prev] <= 0) finished = true;
      pending--;
    }
    slot ^= 1;
  }
  const int M = b->count;
  std::vector<double> rho(2 * (size_t)M), rho0((size_t)M);
  b->itersHost.resize((size_t)M);
  b->ratioHost.resize((size_t)M);
  FB_CUDA(cudaMemcpyAsync(rho.data(), b->rho, sizeof(double) * 2 * (size_t)M, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaMemcpyAsync(rho0.data(), b->rho0, sizeof(double) * (size_t)M, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaMemcpyAsync(b->itersHost.data(), b->iters, sizeof(int) * (size_t)M, cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  FB_CUDA(cudaGetLastError());
  int worst = 0;
  bool anyFailed = false;
  double worstRatio = 0.0;
  for (int m = 0; m < M; m++) {
    const int its = b->itersHost[m];
    const double rf = rho[(size_t)(its & 1) * M + m];
    const bool notConverged = rf > eps * eps * rho0[m];
    b->ratioHost[m] = rho0[m] != 0.0 ? rf / rho0[m] : 0.0;
    if (notConverged) { b->itersHost[m] = -its; anyFailed = true; }  // the reference's return value, per mesh (CGSolver.cpp:189)
    worst = std::max(worst, its);
    worstRatio = std::max(worstRatio, b->ratioHost[m]);
  }
  c->last_iters = anyFailed ? -worst : worst;
  c->last_ratio = worstRatio;
  return FB_OK;
}

extern "C" {

int fb_create_batch(fb_context **out, int count, const int *numVertices, const double *restPositions, const int *numTets,
                    const int *tets, const int *numFixed, const int *fixedVertices, const fb_params *prm) {
  if (!out) return FB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (count < 1 || !numVertices || !numTets || !numFixed) { fb_set_error("bad arguments to fb_create_batch"); return FB_ERR_INVALID_ARGUMENT; }
  FbBatch *b = new FbBatch();
  b->count = count;
  b->vtx.assign((size_t)count + 1, 0);
  b->tet.assign((size_t)count + 1, 0);
  long long nFix = 0;
  for (int m = 0; m < count; m++) {
    if (numVertices[m] < 1 || numTets[m] < 1 || numFixed[m] < 0) { delete b; fb_set_error("mesh %d of the batch is empty or has a negative count", m); return FB_ERR_INVALID_ARGUMENT; }
    const long long nv = (long long)b->vtx[m] + numVertices[m], nt = (long long)b->tet[m] + numTets[m];
    if (nv > 0x7fffffff / 3 || nt > 0x7fffffff / 16) { delete b; fb_set_error("batch too large for 32-bit ids; use several batches"); return FB_ERR_INVALID_ARGUMENT; }
    b->vtx[m + 1] = (int)nv;
    b->tet[m + 1] = (int)nt;
    nFix += numFixed[m];
  }
  const int nV = b->vtx[count], nT = b->tet[count];
  if (!restPositions || !tets || (nFix > 0 && !fixedVertices)) { delete b; fb_set_error("bad arguments to fb_create_batch"); return FB_ERR_INVALID_ARGUMENT; }
  // concatenated mesh: tets and fixed vertices shifted by the mesh's first vertex
  std::vector<int> ct(4 * (size_t)nT), cd;
  cd.reserve(3 * (size_t)nFix);
  size_t fo = 0;
  for (int m = 0; m < count; m++) {
    const int off = b->vtx[m], nv = numVertices[m];
    for (size_t i = 4 * (size_t)b->tet[m]; i < 4 * (size_t)b->tet[m + 1]; i++) {
      if (tets[i] < 0 || tets[i] >= nv) { delete b; fb_set_error("mesh %d: tetrahedron %zu references a vertex outside [0, %d)", m, i / 4 - (size_t)b->tet[m], nv); return FB_ERR_BAD_MESH; }
      ct[i] = tets[i] + off;
    }
    std::vector<int> fv(fixedVertices + fo, fixedVertices + fo + numFixed[m]);
    fo += (size_t)numFixed[m];
    std::sort(fv.begin(), fv.end());  // Deformable::FixedVerticesToFixedDOF (DEF/Deformable.cpp:294-314), per mesh
    for (size_t i = 0; i < fv.size(); i++) {
      if (fv[i] < 0 || fv[i] >= nv) { delete b; fb_set_error("mesh %d: fixed vertex %d out of range [0, %d)", m, fv[i], nv); return FB_ERR_INVALID_ARGUMENT; }
      if (i && fv[i] == fv[i - 1]) { delete b; fb_set_error("mesh %d: fixed vertex %d listed twice", m, fv[i]); return FB_ERR_INVALID_ARGUMENT; }
      for (int k = 0; k < 3; k++) cd.push_back(3 * (fv[i] + off) + k);
    }
  }
  fb_context *c = nullptr;
  int st = fb_create_local(&c, nV, restPositions, nT, ct.data(), (int)cd.size(), cd.data(), nullptr, nullptr, nullptr, prm);
  if (st != FB_OK) { delete b; return st; }
  c->batch = b;
  // chunks of BT_ROWS block rows, never across two meshes
  std::vector<int> chunkRow, chunkMesh, meshChunk((size_t)count + 1, 0);
  for (int m = 0; m < count; m++) {
    meshChunk[m] = (int)chunkMesh.size();
    for (int r0 = b->vtx[m]; r0 < b->vtx[m + 1]; r0 += BT_ROWS) { chunkRow.push_back(r0); chunkMesh.push_back(m); }
  }
  meshChunk[count] = (int)chunkMesh.size();
  chunkRow.push_back(nV);
  b->nChunks = (int)chunkMesh.size();
  // (a chunk ends where the next one starts: the last chunk of a mesh at the first row of the next mesh)
  b->chunkRow = b->chunkMesh = b->meshChunk = b->iters = b->done = b->active = nullptr;
  b->rho = b->rho0 = b->slotsA = b->slotsB = nullptr;
  b->ticket = nullptr;
#define BCHK(call) do { st = (call); if (st != FB_OK) { fb_destroy(c); return st; } } while (0)
#define BCUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fb_set_error("%s -> %s", #call, cudaGetErrorString(e__)); fb_destroy(c); return FB_ERR_CUDA; } } while (0)
  BCHK(fb_dev_alloc(c, &b->chunkRow, chunkRow.size()));
  BCHK(fb_dev_alloc(c, &b->chunkMesh, chunkMesh.size()));
  BCHK(fb_dev_alloc(c, &b->meshChunk, meshChunk.size()));
  BCHK(fb_dev_alloc(c, &b->rho, 2 * (size_t)count));
  BCHK(fb_dev_alloc(c, &b->rho0, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->iters, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->done, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->ticket, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->active, 1));
  BCHK(fb_dev_alloc(c, &b->slotsA, (size_t)b->nChunks));
  BCHK(fb_dev_alloc(c, &b->slotsB, (size_t)b->nChunks));
  BCUDA(cudaMemcpyAsync(b->chunkRow, chunkRow.data(), sizeof(int) * chunkRow.size(), cudaMemcpyHostToDevice, c->stream));
  BCUDA(cudaMemcpyAsync(b->chunkMesh, chunkMesh.data(), sizeof(int) * chunkMesh.size(), cudaMemcpyHostToDevice, c->stream));
  BCUDA(cudaMemcpyAsync(b->meshChunk, meshChunk.data(), sizeof(int) * meshChunk.size(), cudaMemcpyHostToDevice, c->stream));
  BCUDA(cudaMemsetAsync(b->iters, 0, sizeof(int) * (size_t)count, c->stream));
  BCUDA(cudaMemsetAsync(b->done, 0, sizeof(int) * (size_t)count, c->stream));
  BCUDA(cudaMemsetAsync(b->ticket, 0, sizeof(unsigned int) * (size_t)count, c->stream));
  BCUDA(cudaStreamSynchronize(c->stream));
#undef BCHK
#undef BCUDA
  b->itersHost.assign((size_t)count, 0);
  b->ratioHost.assign((size_t)count, 0.0);
  *out = c;
  return FB_OK;
}

int fb_batch_count(const fb_context *c) { return (c && c->batch) ? c->batch->count : 0; }

int fb_batch_offsets(const fb_context *c, int *vertexOffsets, int *tetOffsets) {
  if (!c || !c->batch) { fb_set_error("not a batch context"); return FB_ERR_INVALID_ARGUMENT; }
  if (vertexOffsets) memcpy(vertexOffsets, c->batch->vtx.data(), sizeof(int) * c->batch->vtx.size());
  if (tetOffsets) memcpy(tetOffsets, c->batch->tet.data(), sizeof(int) * c->batch->tet.size());
  return FB_OK;
}

int fb_batch_last_cg_iterations(const fb_context *c, int *iterations, double *residualRatios) {
  if (!c || !c->batch) { fb_set_error("not a batch context"); return FB_ERR_INVALID_ARGUMENT; }
  if (iterations) memcpy(iterations, c->batch->itersHost.data(), sizeof(int) * (size_t)c->batch->count);
  if (residualRatios) memcpy(residualRatios, c->batch->ratioHost.data(), sizeof(double) * (size_t)c->batch->count);
  return FB_OK;
}

}  // extern "C"
