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
// Work units never straddle two meshes.  Products (q = A d, r = b - A x): units of 32 block rows, one WARP per unit, one resident
// wave of CTAs whose warps walk the units without any block barrier.  Vector kernels: units of 256 block rows, every CTA takes
// a contiguous run of them, so it adds a mesh's slots (alpha, beta) once per mesh it touches.  Dot products: per-unit / per-warp
// sums in fixed slots, added per mesh in slot order by every consumer CTA of that mesh — deterministic, no atomics on
// floating-point data.  Measured (B200, 32 meshes of 196,608 tets, profiles/r01_batch_context.txt): 109 mesh-steps/s against
// 79.6 for a context per mesh on 8 streams; the product kernel reads 1.30 GB per launch at 5.2 TB/s (ncu).
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
  int nW, nU;                // product units (one warp each) and vector units (one CTA trip each)
  // device
  int4 *wmeta;               // [nW] {first block row, end block row, mesh, 0} of every product unit
  int4 *umeta;               // [nU] the same for the vector units
  int *meshW, *meshU;        // [count+1] unit ranges of every mesh
  double *rho;               // [2][count]
  double *rho0;              // [count]
  int *iters;                // [count] iterations completed
  int *done;                 // [count]
  unsigned int *ticket;      // [count] vector units of the mesh that have finished the direction update
  int *active;               // [1] meshes still iterating
  int *wnext;                // [1] next product unit to hand out (dynamic distribution)
  double *slotsA;            // [nW] per-unit sums of the products (d.q; after a refresh sum r^2 invD)
  double *slotsB;            // [nU * warps per CTA] per-warp sums of the vector kernels (sum r^2 invD)
  std::vector<int> itersHost;
  std::vector<double> ratioHost;
};

namespace {

constexpr int BT_TB = 256;
constexpr int BT_WARPS = BT_TB / 32;
constexpr int BT_WROWS = 32;   // block rows per product unit: 16 rows for each 16-lane half of the warp
constexpr int BT_UROWS = 256;  // block rows per vector unit: 768 scalars, 3 per thread, all requested at once
constexpr int BT_UITEMS = 3 * BT_UROWS / BT_TB;
constexpr int BT_SD = 2048;    // meshes whose `done` flag is staged in shared memory by the product kernel

struct BatchArgs {
  const int4 *wmeta, *umeta;
  const int *meshW, *meshU;
  double *rho, *rho0;
  int *iters, *done, *active, *wnext;
  unsigned int *ticket;
  double *slotsA, *slotsB;
  int count, maxIt, nW, nU;
  int wnext0;  // > 0: product units beyond the first two per warp are handed out through *wnext, which starts here
  double eps2;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// q = A d with per-unit d.q (MODE 1) / r = b - A x with per-unit sum r^2 invD (MODE 2).  One resident wave of CTAs; every
// WARP walks product units w, w + nWarps, ... on its own (no block barrier in the loop): 16 lanes per block row, all loads of a
// row in flight, the next row's pointers — and the next unit's — requested one step ahead (the loop body of k_spmv_rows3).
template <int MODE>
__global__ void __launch_bounds__(BT_TB, 4) kb_spmv(BatchArgs a, const int *__restrict__ bp, const int *__restrict__ bc,
                                                    const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
                                                    const unsigned char *__restrict__ mask, const double *__restrict__ b,
                                                    const double *__restrict__ invD) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned char sdone[BT_SD];
  for (int i = threadIdx.x; i < a.count && i < BT_SD; i += BT_TB) sdone[i] = (unsigned char)a.done[i];
  __syncthreads();
  const int lane = threadIdx.x & (TILE_G - 1), half = (threadIdx.x >> 4) & 1;
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int nWarps = gridDim.x * BT_WARPS;
  // Units of a warp: its own index, that + nWarps, then (wnext0 > 0) whatever the shared counter hands out next, so that a warp
  // which drew short or skipped units takes more of them.  The counter is read two units ahead (lane 0, broadcast at the end
  // of the unit in between) and the unit record one unit ahead: neither latency is ever waited for.
  const bool dyn = a.wnext0 > 0;
  int w = blockIdx.x * BT_WARPS + (threadIdx.x >> 5);
  int wn = w + nWarps;
  int4 meta = make_int4(0, 0, 0, 0);
  int rs = 0, re = 0;
  if (w < a.nW) {
    meta = __ldg(a.wmeta + w);
    if (meta.x + half < meta.y) { rs = __ldg(bp + meta.x + half); re = __ldg(bp + meta.x + half + 1); }
  }
  while (w < a.nW) {
    int4 metaN = make_int4(0, 0, 0, 0);
    if (wn < a.nW) metaN = __ldg(a.wmeta + wn);
    int wnn = wn + nWarps;
    if (dyn && (threadIdx.x & 31) == 0) wnn = atomicAdd(a.wnext, 1);
    const int m = meta.z;
    const bool fin = (m < BT_SD) ? (sdone[m] != 0) : (a.done[m] != 0);
    if (!fin) {
      double part = 0.0;
      const int r1 = meta.y;
      int v = meta.x + half;
      while (v < r1) {
        const int vn = v + 2;
        int rsn = 0, ren = 0;
        if (vn < r1) { rsn = __ldg(bp + vn); ren = __ldg(bp + vn + 1); }
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
        v = vn; rs = rsn; re = ren;
      }
      // the first row of the warp's next unit, requested before the (short) reduction below
      if (wn < a.nW && metaN.x + half < metaN.y) { rs = __ldg(bp + metaN.x + half); re = __ldg(bp + metaN.x + half + 1); }
      part = warp_sum(part);  // fixed tree: lanes 0-2 of both halves hold the partials, the rest zeros
      if ((threadIdx.x & 31) == 0) a.slotsA[w] = part;
    } else if (wn < a.nW && metaN.x + half < metaN.y) {
      rs = __ldg(bp + metaN.x + half); re = __ldg(bp + metaN.x + half + 1);
    }
    if (dyn) wnn = __shfl_sync(0xffffffffu, wnn, 0);
    w = wn; wn = wnn; meta = metaN;
  }
}

// vector units [u0, u1) of this CTA: contiguous, so a CTA changes mesh (and has to add that mesh's slots) rarely
__device__ __forceinline__ void unit_range(const BatchArgs &a, int *u0, int *u1) {
  *u0 = (int)((long long)blockIdx.x * a.nU / gridDim.x);
  *u1 = (int)((long long)(blockIdx.x + 1) * a.nU / gridDim.x);
}

// r = b (x0 = 0), d = invD r, x = 0, q = 0, per-warp sum r^2 invD                                   (CGSolver.cpp:139-147)
__global__ void __launch_bounds__(BT_TB, 4) kb_init(BatchArgs a, const double *__restrict__ b, const double *__restrict__ invD,
                                                    double *__restrict__ x, double *__restrict__ r, double *__restrict__ d,
                                                    double *__restrict__ q) {
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.wnext = a.wnext0;
  int u0, u1;
  unit_range(a, &u0, &u1);
  for (int u = u0; u < u1; u++) {
    const int4 meta = __ldg(a.umeta + u);
    const int i0 = 3 * meta.x + (int)threadIdx.x, iend = 3 * meta.y;  // 3 nV fits an int (checked by fb_create_batch)
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < BT_UITEMS; k++) {
      const int i = i0 + k * BT_TB;
      if (i < iend) {
        const double bi = b[i], di = invD[i];
        x[i] = 0.0; r[i] = bi; q[i] = 0.0; d[i] = di * bi;
        part += (bi * bi) * di;
      }
    }
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) a.slotsB[(size_t)u * BT_WARPS + (threadIdx.x >> 5)] = part;
  }
}

// per mesh (one CTA each): rho0, loop condition at iteration 1                                      (CGSolver.cpp:147-150)
__global__ void __launch_bounds__(BT_TB) kb_begin(BatchArgs a) {
  const int m = blockIdx.x;
  const double total = cta_sum_slots<BT_TB>(a.slotsB + (size_t)a.meshU[m] * BT_WARPS, (a.meshU[m + 1] - a.meshU[m]) * BT_WARPS);
  if (threadIdx.x != 0) return;
  a.rho[m] = total;       // rho[0][m]
  a.rho0[m] = total;
  a.iters[m] = 0;
  a.ticket[m] = 0u;
  const int fin = !((total > a.eps2 * total) && (1 <= a.maxIt));
  a.done[m] = fin;
  if (fin) atomicSub(a.active, 1);
}

// x += alpha d; REFRESH ? nothing more : (r -= alpha q; per-warp sum r^2 invD)                      (CGSolver.cpp:155-174)
// A unit's operands are requested BEFORE the mesh's scalars are known, so their latency overlaps the slot sum.
template <bool REFRESH>
__global__ void __launch_bounds__(BT_TB, 4) kb_update(BatchArgs a, const double *__restrict__ d, const double *__restrict__ q,
                                                      const double *__restrict__ invD, double *__restrict__ x,
                                                      double *__restrict__ r) {
  pdl_wait();
  pdl_trigger();
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.wnext = a.wnext0;  // for the next product (this kernel runs between two of them)
  int u0, u1;
  unit_range(a, &u0, &u1);
  int cur = -1;
  bool fin = true;
  double alpha = 0.0;
  for (int u = u0; u < u1; u++) {
    const int4 meta = __ldg(a.umeta + u);
    if (meta.z == cur && fin) continue;  // further units of a mesh that has stopped
    const int i0 = 3 * meta.x + (int)threadIdx.x, iend = 3 * meta.y;  // 3 nV fits an int (checked by fb_create_batch)
    double dv[BT_UITEMS], xv[BT_UITEMS], qv[BT_UITEMS], wv[BT_UITEMS], rv[BT_UITEMS];
#pragma unroll
    for (int k = 0; k < BT_UITEMS; k++) {
      const int i = i0 + k * BT_TB;
      dv[k] = xv[k] = qv[k] = wv[k] = rv[k] = 0.0;
      if (i < iend) {
        dv[k] = d[i]; xv[k] = x[i];
        if (!REFRESH) { qv[k] = q[i]; wv[k] = invD[i]; rv[k] = r[i]; }
      }
    }
    if (meta.z != cur) {  // uniform over the CTA
      cur = meta.z;
      fin = a.done[cur] != 0;
      if (!fin) {
        const int it = a.iters[cur] + 1;
        const double dq = cta_sum_slots<BT_TB>(a.slotsA + a.meshW[cur], a.meshW[cur + 1] - a.meshW[cur]);
        alpha = a.rho[((it - 1) & 1) * a.count + cur] / dq;
      }
    }
    if (fin) continue;
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < BT_UITEMS; k++) {
      const int i = i0 + k * BT_TB;
      if (i < iend) {
        x[i] = fma(alpha, dv[k], xv[k]);
        if (!REFRESH) {
          const double ri = fma(-alpha, qv[k], rv[k]);
          r[i] = ri;
          part += (ri * ri) * wv[k];
        }
      }
    }
    if (!REFRESH) {
      part = warp_sum(part);
      if ((threadIdx.x & 31) == 0) a.slotsB[(size_t)u * BT_WARPS + (threadIdx.x >> 5)] = part;
    }
  }
}

// beta = rho'/rho; d = invD r + beta d; per mesh: iteration++ and the loop condition                (CGSolver.cpp:176-183, 150)
// fromProduct: rho' was summed by the refresh product (slotsA, product units) instead of kb_update (slotsB)
__global__ void __launch_bounds__(BT_TB, 4) kb_direction(BatchArgs a, const double *__restrict__ r, const double *__restrict__ invD,
                                                         double *__restrict__ d, int fromProduct) {
  pdl_wait();
  pdl_trigger();
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.wnext = a.wnext0;
  int u0, u1;
  unit_range(a, &u0, &u1);
  int cur = -1, it = 0;
  unsigned int mine = 0u;  // units of mesh `cur` finished by this CTA
  bool fin = true;
  double beta = 0.0, rhoNew = 0.0;
  for (int u = u0; u <= u1; u++) {
    int4 meta = make_int4(0, 0, -1, 0);  // u == u1: only hands in the tickets of the last mesh
    if (u < u1) meta = __ldg(a.umeta + u);
    if (meta.z == cur && fin) continue;
    const int i0 = 3 * meta.x + (int)threadIdx.x, iend = 3 * meta.y;  // 3 nV fits an int (checked by fb_create_batch)
    double rv[BT_UITEMS], wv[BT_UITEMS], dv[BT_UITEMS];
#pragma unroll
    for (int k = 0; k < BT_UITEMS; k++) {
      const int i = i0 + k * BT_TB;
      rv[k] = wv[k] = dv[k] = 0.0;
      if (i < iend) { rv[k] = r[i]; wv[k] = invD[i]; dv[k] = d[i]; }
    }
    if (meta.z != cur) {  // uniform over the CTA
      // Bookkeeping by the CTA that hands in the mesh's last units: every CTA reads iters / rho / done of a mesh before it
      // hands in its tickets for that mesh, so nobody can still be reading them.
      __syncthreads();
      if (threadIdx.x == 0 && cur >= 0 && !fin && mine) {
        __threadfence();
        const unsigned int total = (unsigned int)(a.meshU[cur + 1] - a.meshU[cur]);
        if (atomicAdd(&a.ticket[cur], mine) + mine == total) {
          a.ticket[cur] = 0u;
          a.rho[(it & 1) * a.count + cur] = rhoNew;
          a.iters[cur] = it;
          if (!((rhoNew > a.eps2 * a.rho0[cur]) && (it + 1 <= a.maxIt))) {
            a.done[cur] = 1;
            atomicSub(a.active, 1);
          }
        }
      }
      cur = meta.z;
      mine = 0u;
      fin = true;
      if (cur >= 0) {
        fin = a.done[cur] != 0;
        if (!fin) {
          it = a.iters[cur] + 1;
          rhoNew = fromProduct ? cta_sum_slots<BT_TB>(a.slotsA + a.meshW[cur], a.meshW[cur + 1] - a.meshW[cur])
                               : cta_sum_slots<BT_TB>(a.slotsB + (size_t)a.meshU[cur] * BT_WARPS, (a.meshU[cur + 1] - a.meshU[cur]) * BT_WARPS);
          beta = rhoNew / a.rho[((it - 1) & 1) * a.count + cur];
        }
      }
    }
    if (fin) continue;
#pragma unroll
    for (int k = 0; k < BT_UITEMS; k++) {
      const int i = i0 + k * BT_TB;
      if (i < iend) d[i] = fma(wv[k], rv[k], beta * dv[k]);
    }
    mine++;
  }
}

void batch_args(const fb_context *c, double eps, int maxIt, BatchArgs *a) {
  const FbBatch *b = c->batch;
  a->wmeta = b->wmeta; a->umeta = b->umeta; a->meshW = b->meshW; a->meshU = b->meshU;
  a->rho = b->rho; a->rho0 = b->rho0; a->iters = b->iters; a->done = b->done; a->active = b->active; a->wnext = b->wnext; a->ticket = b->ticket;
  a->slotsA = b->slotsA; a->slotsB = b->slotsB;
  a->count = b->count; a->maxIt = maxIt; a->nW = b->nW; a->nU = b->nU; a->eps2 = eps * eps;
}

}  // namespace

void fb_batch_destroy(fb_context *c) {
  FbBatch *b = c->batch;
  if (!b) return;
  void *ptrs[] = {b->wmeta, b->umeta, b->meshW, b->meshU, b->rho, b->rho0, b->iters, b->done, b->ticket, b->active, b->wnext, b->slotsA, b->slotsB};
  for (void *p : ptrs)
    if (p) fb_dev_free(p);
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
  const int wave = 4 * c->sm_count;  // one resident wave: 4 CTAs of 256 threads per SM
  const int gridP = std::max(1, std::min(wave, (b->nW + BT_WARPS - 1) / BT_WARPS));
  const int gridV = std::max(1, std::min(wave, b->nU));
  // hand-out through the counter: measured no gain over the static walk (108.5 vs 109.1 mesh-steps/s), opt-in
  static const bool dynamicUnits = getenv("FEMBRAIN_B200_BATCH_DYNAMIC") && atoi(getenv("FEMBRAIN_B200_BATCH_DYNAMIC")) != 0;
  a.wnext0 = dynamicUnits ? 2 * gridP * BT_WARPS : 0;
  FB_CUDA(cudaMemcpyAsync(b->active, &b->count, sizeof(int), cudaMemcpyHostToDevice, st));
  kb_init<<<gridV, BT_TB, 0, st>>>(a, c->rhs, c->invD, c->x, c->res, c->dir, c->Ad);
  kb_begin<<<b->count, BT_TB, 0, st>>>(a);
  c->launches += 2;
  const int CH = 30;
  int *activeHost[2] = {reinterpret_cast<int *>(&c->sc_host[0]), reinterpret_cast<int *>(&c->sc_host[1])};  // pinned
  int it = 1, slot = 0, pending = 0;
  int activeSeen = b->count;  // as of the last poll (one chunk behind)
  bool finished = false;
  c->nprof = 0;
  while (!finished && it <= maxIt) {
    const int end = (it + CH - 1 < maxIt) ? it + CH - 1 : maxIt;
    for (; it <= end; it++) {
      // a mesh that has stopped is skipped by every kernel; meshes that stop at different iterations keep their own
      // refresh phase because `it` is common to all meshes that are still iterating (all started at 1)
      // fb_set_profiling: event pairs around every 16th product while (as far as the host knows) all meshes still iterate
      const bool sample = c->profiling && (it % 16 == 1) && c->nprof < 64 && activeSeen == b->count;
      if (sample) cudaEventRecord(c->evProf[2 * c->nprof], st);
      fb_launch(c->pdl, st, kb_spmv<1>, gridP, BT_TB, a, c->bp, c->bc, c->Keff, c->dir, c->Ad, c->rowmask, c->rhs, c->invD);
      if (sample) { cudaEventRecord(c->evProf[2 * c->nprof + 1], st); c->nprof++; }
      if (it % 30 == 0) {
        fb_launch(c->pdl, st, kb_update<true>, gridV, BT_TB, a, c->dir, c->Ad, c->invD, c->x, c->res);
        fb_launch(c->pdl, st, kb_spmv<2>, gridP, BT_TB, a, c->bp, c->bc, c->Keff, c->x, c->res, c->rowmask, c->rhs, c->invD);
        fb_launch(c->pdl, st, kb_direction, gridV, BT_TB, a, c->res, c->invD, c->dir, 1);
        c->launches++;
      } else {
        fb_launch(c->pdl, st, kb_update<false>, gridV, BT_TB, a, c->dir, c->Ad, c->invD, c->x, c->res);
        fb_launch(c->pdl, st, kb_direction, gridV, BT_TB, a, c->res, c->invD, c->dir, 0);
      }
      c->launches += 3;
    }
    FB_CUDA(cudaMemcpyAsync(activeHost[slot], b->active, sizeof(int), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaEventRecord(c->evChunk[slot], st));
    pending++;
    if (pending == 2) {
      const int prev = slot ^ 1;
      FB_CUDA(cudaEventSynchronize(c->evChunk[prev]));
      activeSeen = *activeHost[prev];
      if (activeSeen <= 0) finished = true;
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
  for (int i = 0; i < c->nprof; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->evProf[2 * i], c->evProf[2 * i + 1]) == cudaSuccess) { c->prof_sum_s += 1e-3 * ms; c->prof_samples++; }
  }
  c->nprof = 0;
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
      if (tets[i] < 0 || tets[i] >= nv) {
        fb_set_error("mesh %d: tetrahedron %zu references a vertex outside [0, %d)", m, i / 4 - (size_t)b->tet[m], nv);
        delete b;
        return FB_ERR_BAD_MESH;
      }
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
  // product units of BT_WROWS block rows and vector units of BT_UROWS, never across two meshes
  std::vector<int4> wmeta, umeta;
  std::vector<int> meshW((size_t)count + 1, 0), meshU((size_t)count + 1, 0);
  int wrows = BT_WROWS;
  if (const char *e = getenv("FEMBRAIN_B200_BATCH_WROWS")) wrows = std::max(2, atoi(e));  // experiments
  for (int m = 0; m < count; m++) {
    meshW[m] = (int)wmeta.size();
    meshU[m] = (int)umeta.size();
    const int e = b->vtx[m + 1];
    for (int r0 = b->vtx[m]; r0 < e; r0 += wrows) wmeta.push_back(make_int4(r0, std::min(r0 + wrows, e), m, 0));
    for (int r0 = b->vtx[m]; r0 < e; r0 += BT_UROWS) umeta.push_back(make_int4(r0, std::min(r0 + BT_UROWS, e), m, 0));
  }
  meshW[count] = (int)wmeta.size();
  meshU[count] = (int)umeta.size();
  b->nW = (int)wmeta.size();
  b->nU = (int)umeta.size();
  b->wmeta = b->umeta = nullptr;
  b->meshW = b->meshU = b->iters = b->done = b->active = b->wnext = nullptr;
  b->rho = b->rho0 = b->slotsA = b->slotsB = nullptr;
  b->ticket = nullptr;
#define BCHK(call) do { st = (call); if (st != FB_OK) { fb_destroy(c); return st; } } while (0)
#define BCUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fb_set_error("%s -> %s", #call, cudaGetErrorString(e__)); fb_destroy(c); return FB_ERR_CUDA; } } while (0)
  BCHK(fb_dev_alloc(c, &b->wmeta, wmeta.size()));
  BCHK(fb_dev_alloc(c, &b->umeta, umeta.size()));
  BCHK(fb_dev_alloc(c, &b->meshW, meshW.size()));
  BCHK(fb_dev_alloc(c, &b->meshU, meshU.size()));
  BCHK(fb_dev_alloc(c, &b->rho, 2 * (size_t)count));
  BCHK(fb_dev_alloc(c, &b->rho0, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->iters, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->done, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->ticket, (size_t)count));
  BCHK(fb_dev_alloc(c, &b->active, 1));
  BCHK(fb_dev_alloc(c, &b->wnext, 1));
  BCHK(fb_dev_alloc(c, &b->slotsA, (size_t)b->nW));
  BCHK(fb_dev_alloc(c, &b->slotsB, (size_t)b->nU * BT_WARPS));
  BCUDA(cudaMemcpyAsync(b->wmeta, wmeta.data(), sizeof(int4) * wmeta.size(), cudaMemcpyHostToDevice, c->stream));
  BCUDA(cudaMemcpyAsync(b->umeta, umeta.data(), sizeof(int4) * umeta.size(), cudaMemcpyHostToDevice, c->stream));
  BCUDA(cudaMemcpyAsync(b->meshW, meshW.data(), sizeof(int) * meshW.size(), cudaMemcpyHostToDevice, c->stream));
  BCUDA(cudaMemcpyAsync(b->meshU, meshU.data(), sizeof(int) * meshU.size(), cudaMemcpyHostToDevice, c->stream));
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
