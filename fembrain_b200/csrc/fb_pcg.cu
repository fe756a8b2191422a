// fb_pcg.cu — Jacobi-preconditioned conjugate gradients on the constrained effective matrix.
//
// Reference: CGSolver::SolveLinearSystemWithJacobiPreconditioner
// (src/3rdparty/vegafem/sparseSolver/CGSolver.cpp:129-190; SpMV sparseMatrix/sparseMatrix.cpp:405-413):
//   invD = 1/diag(A); r = b - A x0 (x0 = 0 => r = b); d = invD r; rho = sum r^2 invD; rho0 = rho
//   while rho > eps^2 rho0 and it <= maxIt:
//     q = A d; alpha = rho / (d.q); x += alpha d
//     it % 30 == 0 ? r = b - A x : r -= alpha q
//     rho' = sum r^2 invD; beta = rho'/rho; d = invD r + beta d
// Same recurrences for x, r, d, same refresh period, same stopping rule (tested on the directly summed
// rho'); sums are parallel reductions with a FIXED association order (deterministic run to run), so iterates
// agree with the reference's sequential sums to rounding, not bitwise.
//
// The constrained system (fixed rows/columns removed, CGSolver on systemMatrix) is solved IN PLACE on
// full-length vectors: rows of constrained DOFs produce 0 and their entries of d, r, x stay 0, which
// contributes exact zeros to every sum — arithmetically the compacted system, without the per-step gather of
// AssignSuperMatrix (sparseMatrix.cpp:993-1002).
//
// Two schedules, both with alpha/beta/rho and the loop flag resident on the device:
//  * fused (default on one GPU), TWO kernels per iteration:
//      k_spmv_rows3<3>   q = A d, and the three sums d.q, (r,q)_D, (q,q)_D  ((u,v)_D = sum u v / diag)
//      k_fused_update    alpha = rho/d.q; x += alpha d; r -= alpha q; d = invD r + beta d; rho' = sum r^2 invD
//    beta needs rho' before r is updated; it is taken from the identity rho' = rho - 2 alpha (r,q)_D +
//    alpha^2 (q,q)_D (exact in exact arithmetic, differs from the direct sum by rounding), while the rho used
//    for the next alpha and for the stopping test is the DIRECT sum accumulated in the same pass.  This removes one
//    launch and 24 B/row of vector traffic per iteration.  Every 30th iteration runs the reference's refresh
//    unfused (x update, r = b - A x with the direct rho', direction update).
//    A CUDA graph replays one 30-iteration period (58 + 4 kernels) per launch.
//  * kernels (partitioned contexts; FEMBRAIN_B200_PCG=kernels), THREE kernels per iteration in the reference's
//    literal order: k_spmv<1> (q, d.q), k_update (x, r, rho'), k_direction (beta, d), with NCCL calls between.
// Per-CTA partial sums go to a fixed slot; the last CTA to finish (integer ticket) adds the slots in index order.
//
// Matrix layout: the reference's CSR value order with 3x3-block-compressed column indices: block row v owns
// 9*nb doubles at 9*bp[v]: three scalar rows of 3*nb values each; column of entry t of a scalar row is
// 3*bc[bp[v] + t/3] + t%3.  8.44 bytes per nonzero instead of CSR's 12.
#include <cstdlib>
#include <cstring>

#include "fb_internal.h"
#include "fb_pcg_common.cuh"

namespace {

constexpr int SPMV_TB = 256;
constexpr int VEC_TB = 256;

// Deterministic block reduction of N values followed by the "last block adds all slots in order" pattern.
// Value k of CTA b goes to slots[k * FB_MAX_PARTIALS + b].  Returns true in every thread of the last block;
// total[] is then valid in thread 0.
template <int TB, int N>
__device__ __forceinline__ bool block_reduce_to_total(double (&v)[N], double *slots, unsigned int *ticket, double (&total)[N]) {
  __shared__ double wsum[N][TB / 32];
  __shared__ bool isLast;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) wsum[k][warp] = s;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      double s = (lane < TB / 32) ? wsum[k][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if (lane == 0) slots[(size_t)k * FB_MAX_PARTIALS + blockIdx.x] = s;
    }
    if (lane == 0) {
      __threadfence();
      unsigned int tk = atomicAdd(ticket, 1u);
      isLast = (tk == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!isLast) return false;
  __threadfence();
  // fixed-order final sum: thread t adds slots t, t+TB, ...; then the same block tree
#pragma unroll
  for (int k = 0; k < N; k++) {
    double s = 0.0;
    const volatile double *sl = slots + (size_t)k * FB_MAX_PARTIALS;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += TB) s += sl[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __syncthreads();
    if (lane == 0) wsum[k][warp] = s;
    __syncthreads();
    if (warp == 0) {
      double z = (lane < TB / 32) ? wsum[k][lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) z += __shfl_down_sync(0xffffffffu, z, o);
      if (lane == 0) total[k] = z;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;
  return true;
}

// what the lanes holding a finished row do with it, per MODE:
// 0: y = A x                                   generic product (no mask, no sums)
// 1: y = mask(A x); sum x.y                    q = A d with d.q                 (kernels schedule)
// 2: y = mask(b - A x); sum y^2 invD           exact-residual refresh
// 3: y = mask(A x); sums x.y, (rv,y)_D, (y,y)_D   q = A d with the three sums   (fused schedule; rv = r)
template <int MODE>
__device__ __forceinline__ void finish_row(double s, size_t row, const double *__restrict__ x, double *__restrict__ y,
                                           const unsigned char *__restrict__ mask, const double *__restrict__ b,
                                           const double *__restrict__ invD, double (&part)[3]) {
  if (MODE == 0) {
    y[row] = s;
  } else if (MODE == 1) {
    if (mask[row]) s = 0.0;
    y[row] = s;
    part[0] = fma(x[row], s, part[0]);
  } else if (MODE == 2) {
    const double rres = mask[row] ? 0.0 : (b[row] - s);
    y[row] = rres;
    part[0] += (rres * rres) * invD[row];
  } else {
    if (mask[row]) s = 0.0;
    y[row] = s;
    const double wi = invD[row];
    part[0] = fma(x[row], s, part[0]);
    part[1] += (b[row] * s) * wi;  // b carries r in this mode
    part[2] += (s * s) * wi;
  }
}

template <int MODE>
__device__ __forceinline__ void finish_kernel(double (&part)[3], FbScalars *sc, double *slots, double *outp, const FbPeerArgs &pa) {
  if (MODE == 1 || MODE == 2) {
    if (pa.enabled) {  // partitioned, peer-memory exchange: this rank's total goes to every rank's comm block
      double p1[1] = {part[0]}, t1[1];
      if (block_reduce_to_total<SPMV_TB, 1>(p1, slots, &sc->ticket_a, t1) && threadIdx.x == 0)
        peer_publish(pa, MODE == 1 ? FB_COMM_DQ : FB_COMM_RHO, t1[0]);
      return;
    }
    if (outp == nullptr) {  // deferred: the consumer kernel adds the slots
      block_reduce_to_slot<SPMV_TB>(part[0], slots);
      return;
    }
    double p1[1] = {part[0]}, t1[1];
    if (block_reduce_to_total<SPMV_TB, 1>(p1, slots, &sc->ticket_a, t1) && threadIdx.x == 0) *outp = t1[0];
  } else if (MODE == 3) {
    double t3[3];
    if (block_reduce_to_total<SPMV_TB, 3>(part, slots, &sc->ticket_a, t3) && threadIdx.x == 0) {
      sc->dq = t3[0];
      sc->rq = t3[1];
      sc->qq = t3[2];
    }
  }
}

// ---- generic row-per-G-lanes SpMV (G = 8 or 32: meshes with very short or very long rows) -------------------
template <int G, int MODE>
__global__ void __launch_bounds__(SPMV_TB) k_spmv(int rowBeg, int rowEnd, const int *__restrict__ bp, const int *__restrict__ bc,
                                                  const double *__restrict__ A, const double *__restrict__ x,
                                                  double *__restrict__ y, const unsigned char *__restrict__ mask,
                                                  const double *__restrict__ b, const double *__restrict__ invD,
                                                  FbScalars *sc, double *slots, double *outp, FbPeerArgs pa) {
  pdl_wait();
  pdl_trigger();
  if (MODE != 0) {
    if (sc->done) return;
  }
  const int lane = threadIdx.x & (G - 1);
  // the G lanes of a group always take the same trips through the row loop; other groups of the warp may
  // not, so shuffles name only the group's own lanes
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  const int groupsPerBlock = SPMV_TB / G;
  const int group = blockIdx.x * groupsPerBlock + threadIdx.x / G;
  const int nGroups = gridDim.x * groupsPerBlock;
  double part[3] = {0.0, 0.0, 0.0};
  for (int v = rowBeg + group; v < rowEnd; v += nGroups) {
    const int rs = __ldg(bp + v), re = __ldg(bp + v + 1);
    const int n3 = 3 * (re - rs);
    const double *a0 = A + 9 * (size_t)rs;
    const double *a1 = a0 + n3;
    const double *a2 = a1 + n3;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    int t = lane;
    for (; t + G < n3; t += 2 * G) {  // two passes per trip: six independent streaming loads in flight per lane
      const int jb0 = t / 3, l0 = t - 3 * jb0;
      const int t1 = t + G;
      const int jb1 = t1 / 3, l1 = t1 - 3 * jb1;
      const int c0 = __ldg(bc + rs + jb0), c1 = __ldg(bc + rs + jb1);
      const double v00 = ld_stream(a0 + t), v01 = ld_stream(a1 + t), v02 = ld_stream(a2 + t);
      const double v10 = ld_stream(a0 + t1), v11 = ld_stream(a1 + t1), v12 = ld_stream(a2 + t1);
      const double x0 = __ldg(x + 3 * (size_t)c0 + l0), x1 = __ldg(x + 3 * (size_t)c1 + l1);
      acc0 = fma(v00, x0, acc0); acc1 = fma(v01, x0, acc1); acc2 = fma(v02, x0, acc2);
      acc0 = fma(v10, x1, acc0); acc1 = fma(v11, x1, acc1); acc2 = fma(v12, x1, acc2);
    }
    if (t < n3) {
      const int jb0 = t / 3, l0 = t - 3 * jb0;
      const int c0 = __ldg(bc + rs + jb0);
      const double v00 = ld_stream(a0 + t), v01 = ld_stream(a1 + t), v02 = ld_stream(a2 + t);
      const double x0 = __ldg(x + 3 * (size_t)c0 + l0);
      acc0 = fma(v00, x0, acc0); acc1 = fma(v01, x0, acc1); acc2 = fma(v02, x0, acc2);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) {
      acc0 += __shfl_xor_sync(gmask, acc0, o, G);
      acc1 += __shfl_xor_sync(gmask, acc1, o, G);
      acc2 += __shfl_xor_sync(gmask, acc2, o, G);
    }
    if (lane < 3) finish_row<MODE>((lane == 0) ? acc0 : ((lane == 1) ? acc1 : acc2), 3 * (size_t)v + lane, x, y, mask, b, invD, part);
  }
  finish_kernel<MODE>(part, sc, slots, outp, pa);
}

// ---- row-per-16-lanes SpMV with every load of a row in flight at once (no shared memory, no block syncs) ----
// The three 16-wide passes over a row are fully unrolled and predicated, so a lane has 9 streaming value loads +
// 3 column loads outstanding before the first multiply, and the row pointers of the group's next row are fetched
// one row ahead.  MINB = resident CTAs per SM requested from the compiler (4 -> <= 64 registers, no spills).
//
// Block rows [rowBeg, nV): all rows on one GPU, the owned rows of a partitioned context (ghost rows are not visited).
// With peer-memory exchange the wait for the neighbours' halo of d sits at the top.  Tried and dropped
// (profiles/r01_spmv_segments_ab.txt): (a) one kernel walking three row segments — rows that read no ghost column, the
// wait, then the rows next to the cuts: sharing that body cost the single-GPU kernel 10 % (38.9 vs 35.2 us at 1M tets);
// (b) two launches, interior rows first and the rows next to the cuts (which wait, add both sums and publish) second:
// 279 vs 268 ms per step on 2 GPUs at 10M tets — the halo has already landed when the product starts, so there is no
// wait to hide, and the second launch's own latency chain is added to every iteration.
template <int MODE, int MINB, bool EVICT = false>
__global__ void __launch_bounds__(SPMV_TB, MINB) k_spmv_rows3(int rowBeg, int nV, const int *__restrict__ bp, const int *__restrict__ bc,
                                                              const double *__restrict__ A, const double *__restrict__ x,
                                                              double *__restrict__ y, const unsigned char *__restrict__ mask,
                                                              const double *__restrict__ b, const double *__restrict__ invD,
                                                              FbScalars *sc, double *slots, double *outp, FbPeerArgs pa) {
  pdl_wait();
  pdl_trigger();
  if (MODE != 0) {
    if (sc->done) return;
    if (pa.enabled && pa.haloMask) peer_wait_halo(pa, sc);  // the neighbours' d has landed in my ghost entries
  }
  const unsigned long long policy = EVICT ? l2_policy_evict_first() : 0ull;
  const int lane = threadIdx.x & (TILE_G - 1);
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int groupsPerBlock = SPMV_TB / TILE_G;
  const int group = blockIdx.x * groupsPerBlock + threadIdx.x / TILE_G;
  const int nGroups = gridDim.x * groupsPerBlock;
  double part[3] = {0.0, 0.0, 0.0};
  int v = rowBeg + group;
  int rs = 0, re = 0;
  if (v < nV) { rs = __ldg(bp + v); re = __ldg(bp + v + 1); }
  while (v < nV) {
    const int vn = v + nGroups;
    int rsn = 0, ren = 0;
    if (vn < nV) { rsn = __ldg(bp + vn); ren = __ldg(bp + vn + 1); }
    const int n3 = 3 * (re - rs);
    // the lanes that will own the finished row fetch its vector entries now, so the loads fly with the row's values
    const size_t row = 3 * (size_t)v + (lane < 3 ? lane : 0);
    double xr = 0.0, br = 0.0, wr = 0.0;
    unsigned char mk = 0;
    if (lane < 3) {
      if (MODE != 0) mk = __ldg(mask + row);
      if (MODE == 1 || MODE == 3) xr = __ldg(x + row);
      if (MODE == 2 || MODE == 3) { br = __ldg(b + row); wr = __ldg(invD + row); }
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
      if (EVICT) load_row_chunk_hint(A, rs, n3, base, lane, val, policy);
      else load_row_chunk(A, rs, n3, base, lane, val);
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
      if (MODE == 0) {
        y[row] = s;
      } else if (MODE == 1) {
        if (mk) s = 0.0;
        y[row] = s;
        part[0] = fma(xr, s, part[0]);
      } else if (MODE == 2) {
        const double rres = mk ? 0.0 : (br - s);
        y[row] = rres;
        part[0] += (rres * rres) * wr;
      } else {
        if (mk) s = 0.0;
        y[row] = s;
        part[0] = fma(xr, s, part[0]);
        part[1] += (br * s) * wr;  // b carries r in this mode
        part[2] += (s * s) * wr;
      }
    }
    v = vn; rs = rsn; re = ren;
  }
  finish_kernel<MODE>(part, sc, slots, outp, pa);
}

// r = b (x0 = 0), d = invD r, x = 0, rho0 = sum r^2 invD           (CGSolver.cpp:139-147)
__global__ void __launch_bounds__(VEC_TB) k_cg_init(int n, const double *__restrict__ b, const double *__restrict__ invD,
                                                    double *__restrict__ x, double *__restrict__ r, double *__restrict__ d,
                                                    double *__restrict__ q,
                                                    FbScalars *sc, double *slots, double *outp, FbPeerArgs pa,
                                                    const unsigned char *__restrict__ skipMask) {
  double part[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * VEC_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * VEC_TB) {
    const double bi = b[i], di = invD[i];
    x[i] = 0.0;
    r[i] = bi;
    q[i] = 0.0;  // rows the products skip (ghost rows of a partitioned context) must read as zero in the updates
    // peer-memory exchange: ghost entries of d belong to the neighbour's push (constrained entries stay 0 for ever)
    if (!(skipMask && skipMask[i])) d[i] = di * bi;
    part[0] += (bi * bi) * di;
  }
  double total[1];
  if (block_reduce_to_total<VEC_TB, 1>(part, slots, &sc->ticket_b, total) && threadIdx.x == 0) {
    if (pa.enabled) peer_publish(pa, FB_COMM_RHO, total[0]);
    else *outp = total[0];
  }
}

// after rho[0] is final (all-reduced in partitioned contexts): initial residual, loop condition at iteration 1
__global__ void __launch_bounds__(32) k_cg_begin(FbScalars *sc, double eps, int maxIt, FbPeerArgs pa) {
  double total;
  if (pa.enabled) {
    sc->comm_error = 0;
    total = peer_collect<32>(pa, FB_COMM_RHO, sc);
    if (threadIdx.x != 0) return;
    sc->rho[0] = total;
  } else {
    if (threadIdx.x != 0) return;
    total = sc->rho[0];
  }
  sc->rho0 = total;
  sc->eps2 = eps * eps;
  sc->max_it = maxIt;
  sc->iters = 0;
  sc->dq = sc->rq = sc->qq = 0.0;
  // while ((residualNorm2 > eps*eps*initialResidualNorm2) && (iteration <= maxIterations)), iteration = 1
  sc->done = (!((total > eps * eps * total) && (1 <= maxIt))) || (pa.enabled && sc->comm_error);
}

// x += alpha d; REFRESH ? nothing more : (r -= alpha q; rho' = sum r^2 invD)     (CGSolver.cpp:155-174)
// itArg > 0: iteration number from the host; itArg <= 0: from the device counter (graph replay)
template <bool REFRESH>
__global__ void __launch_bounds__(VEC_TB) k_update(int n, const double *__restrict__ d, const double *__restrict__ q,
                                                   const double *__restrict__ invD, double *__restrict__ x,
                                                   double *__restrict__ r, FbScalars *sc, double *slots, int itArg, double *outp,
                                                   const double *__restrict__ dqSlots, int nDqSlots, FbPeerArgs pa) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const int it = itArg > 0 ? itArg : sc->iters + 1;
  // two doubles per thread per trip (128-bit loads/stores); element n-1 of an odd-length vector is handled last.
  // The first item's operands are requested BEFORE alpha is known, so their latency overlaps the slot sum / peer wait.
  const size_t n2 = (size_t)n >> 1;
  const double2 *d2 = reinterpret_cast<const double2 *>(d), *q2 = reinterpret_cast<const double2 *>(q);
  const double2 *w2 = reinterpret_cast<const double2 *>(invD);
  double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
  const size_t stride = (size_t)gridDim.x * VEC_TB;
  size_t i = (size_t)blockIdx.x * VEC_TB + threadIdx.x;
  double2 dv = make_double2(0.0, 0.0), xv = dv, qv = dv, wv = dv, rv = dv;
  if (i < n2) {
    dv = d2[i]; xv = x2[i];
    if (!REFRESH) { qv = q2[i]; wv = w2[i]; rv = r2[i]; }
  }
  const double dq = pa.enabled ? peer_collect<VEC_TB>(pa, FB_COMM_DQ, sc)
                               : (dqSlots ? cta_sum_slots<VEC_TB>(dqSlots, nDqSlots) : sc->dq);
  const double alpha = sc->rho[(it - 1) & 1] / dq;
  double part[1] = {0.0};
  while (i < n2) {
    xv.x = fma(alpha, dv.x, xv.x); xv.y = fma(alpha, dv.y, xv.y);
    x2[i] = xv;
    if (!REFRESH) {
      rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
      r2[i] = rv;
      part[0] += (rv.x * rv.x) * wv.x;
      part[0] += (rv.y * rv.y) * wv.y;
    }
    i += stride;
    if (i < n2) {
      dv = d2[i]; xv = x2[i];
      if (!REFRESH) { qv = q2[i]; wv = w2[i]; rv = r2[i]; }
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const size_t i = (size_t)n - 1;
    x[i] = fma(alpha, d[i], x[i]);
    if (!REFRESH) {
      const double ri = fma(-alpha, q[i], r[i]);
      r[i] = ri;
      part[0] += (ri * ri) * invD[i];
    }
  }
  if (!REFRESH) {
    if (pa.enabled) {
      double total[1];
      if (block_reduce_to_total<VEC_TB, 1>(part, slots, &sc->ticket_b, total) && threadIdx.x == 0) peer_publish(pa, FB_COMM_RHO, total[0]);
    } else if (outp == nullptr) {
      block_reduce_to_slot<VEC_TB>(part[0], slots);
    } else {
      double total[1];
      if (block_reduce_to_total<VEC_TB, 1>(part, slots, &sc->ticket_b, total) && threadIdx.x == 0) *outp = total[0];
    }
  }
}

// beta = rho'/rho; d = invD r + beta d; iteration++ and loop condition            (CGSolver.cpp:176-183, 150)
__global__ void __launch_bounds__(VEC_TB) k_direction(int n, const double *__restrict__ r, const double *__restrict__ invD,
                                                      double *__restrict__ d, FbScalars *sc, int itArg,
                                                      const double *__restrict__ rhoSlots, int nRhoSlots, FbPeerArgs pa,
                                                      const unsigned char *__restrict__ skipMask, FbPushArgs push) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const int it = itArg > 0 ? itArg : sc->iters + 1;
  const size_t n2 = (size_t)n >> 1;
  const double2 *r2 = reinterpret_cast<const double2 *>(r), *w2 = reinterpret_cast<const double2 *>(invD);
  const uchar2 *m2 = reinterpret_cast<const uchar2 *>(skipMask);
  double2 *d2 = reinterpret_cast<double2 *>(d);
  const size_t stride = (size_t)gridDim.x * VEC_TB;
  size_t i = (size_t)blockIdx.x * VEC_TB + threadIdx.x;
  // first item requested before beta is known: the loads overlap the slot sum / peer wait
  double2 rv = make_double2(0.0, 0.0), wv = rv, dv = rv;
  uchar2 mk = make_uchar2(0, 0);
  if (i < n2) {
    rv = r2[i]; wv = w2[i]; dv = d2[i];
    if (skipMask) mk = m2[i];
  }
  const double rhoNew = pa.enabled ? peer_collect<VEC_TB>(pa, FB_COMM_RHO, sc)
                                   : (rhoSlots ? cta_sum_slots<VEC_TB>(rhoSlots, nRhoSlots) : sc->rho[it & 1]);
  const double rhoOld = sc->rho[(it - 1) & 1];
  const double eps2 = sc->eps2, rho0 = sc->rho0;
  const int maxIt = sc->max_it;
  const double beta = rhoNew / rhoOld;
  while (i < n2) {
    dv.x = fma(wv.x, rv.x, beta * dv.x); dv.y = fma(wv.y, rv.y, beta * dv.y);
    // peer-memory exchange: ghost entries of d (masked) are written by the neighbour's push, never here
    if (!(mk.x | mk.y)) {
      d2[i] = dv;
    } else {
      if (!mk.x) d[2 * i] = dv.x;
      if (!mk.y) d[2 * i + 1] = dv.y;
    }
    if (push.nNbr) {  // owned entries next to a cut go straight into the neighbours' ghost entries (NVLink stores)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const size_t dof = 2 * i + h;
        const int v = (int)(dof / 3), k = (int)(dof - 3 * (size_t)v);
        if (push.pushFlag[v]) {
          const double val = h ? dv.y : dv.x;
          for (int e = push.pushPtr[v]; e < push.pushPtr[v + 1]; e++) {
            const int2 t = push.pushEnt[e];
            push.peerVec[t.x][3 * (size_t)t.y + k] = val;
          }
        }
      }
    }
    i += stride;
    if (i < n2) {
      rv = r2[i]; wv = w2[i]; dv = d2[i];
      if (skipMask) mk = m2[i];
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0 && !(skipMask && skipMask[n - 1])) {
    const double val = fma(invD[n - 1], r[n - 1], beta * d[n - 1]);
    d[n - 1] = val;
    if (push.nNbr) {
      const int v = (n - 1) / 3, k = (n - 1) - 3 * v;
      for (int e = push.pushPtr[v]; e < push.pushPtr[v + 1]; e++) push.peerVec[push.pushEnt[e].x][3 * (size_t)push.pushEnt[e].y + k] = val;
    }
  }
  // bookkeeping by the last CTA to finish, so that no CTA of this launch can still be reading sc->iters / done
  __shared__ bool last;
  if (push.nNbr) __threadfence_system();  // this thread's peer stores are out before the CTA's ticket
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = (atomicAdd(&sc->ticket_b, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    if (push.nNbr) {  // every CTA's stores are out: raise this rank's halo flag in each neighbour's comm block
      __threadfence_system();
      for (int j = 0; j < push.nNbr; j++)
        ((volatile unsigned long long *)pa.comm[push.nbrRank[j]])[FB_COMM_FLAG(pa.parity, FB_COMM_HALO, pa.rank)] = push.epoch;
    }
    sc->ticket_b = 0u;
    if (rhoSlots || pa.enabled) sc->rho[it & 1] = rhoNew;
    sc->iters = it;
    if (!((rhoNew > eps2 * rho0) && (it + 1 <= maxIt))) sc->done = 1;
  }
}

// The fused vector kernel of the two-kernel schedule (see the header comment).
__global__ void __launch_bounds__(VEC_TB) k_fused_update(int n, const double *__restrict__ q, const double *__restrict__ invD,
                                                         double *__restrict__ x, double *__restrict__ r, double *__restrict__ d,
                                                         FbScalars *sc, double *slots) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const int it = sc->iters + 1;
  const double rho = sc->rho[(it - 1) & 1];
  const double alpha = rho / sc->dq;
  // rho' = (r - alpha q, r - alpha q)_D = rho - 2 alpha (r,q)_D + alpha^2 (q,q)_D
  const double rhoF = fma(alpha, fma(alpha, sc->qq, -2.0 * sc->rq), rho);
  const double beta = rhoF / rho;
  const double eps2 = sc->eps2, rho0 = sc->rho0;
  const int maxIt = sc->max_it;
  double part[1] = {0.0};
  const size_t n2 = (size_t)n >> 1;
  const double2 *q2 = reinterpret_cast<const double2 *>(q), *w2 = reinterpret_cast<const double2 *>(invD);
  double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r), *d2 = reinterpret_cast<double2 *>(d);
  for (size_t i = (size_t)blockIdx.x * VEC_TB + threadIdx.x; i < n2; i += (size_t)gridDim.x * VEC_TB) {
    const double2 qv = q2[i], wv = w2[i];
    double2 dv = d2[i], xv = x2[i], rv = r2[i];
    xv.x = fma(alpha, dv.x, xv.x); xv.y = fma(alpha, dv.y, xv.y);
    rv.x = fma(-alpha, qv.x, rv.x); rv.y = fma(-alpha, qv.y, rv.y);
    dv.x = fma(wv.x, rv.x, beta * dv.x); dv.y = fma(wv.y, rv.y, beta * dv.y);
    x2[i] = xv; r2[i] = rv; d2[i] = dv;
    part[0] += (rv.x * rv.x) * wv.x;
    part[0] += (rv.y * rv.y) * wv.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const size_t i = (size_t)n - 1;
    const double di = d[i], wi = invD[i];
    x[i] = fma(alpha, di, x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    d[i] = fma(wi, ri, beta * di);
    part[0] += (ri * ri) * wi;
  }
  double total[1];
  if (block_reduce_to_total<VEC_TB, 1>(part, slots, &sc->ticket_b, total) && threadIdx.x == 0) {
    sc->rho[it & 1] = total[0];  // the directly summed rho': next alpha and the stopping rule use this one
    sc->iters = it;
    if (!((total[0] > eps2 * rho0) && (it + 1 <= maxIt))) sc->done = 1;
  }
}

// ---- launch shapes: ONE resident wave per kernel (sm_count x CTAs/SM from the occupancy API), so grid-stride
// loops load every SM equally (no partial second wave) and the number of per-CTA partial sums stays small.
template <typename K>
int one_wave(const fb_context *c, K kernel, int tb, size_t want) {
  int perSM = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, tb, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
  size_t cap = (size_t)c->sm_count * (size_t)perSM;
  // (Sizing the grids of a batch's contexts for 1/N of the GPU so that N of them run side by side was tried for config 4:
  //  32 meshes of 196,608 tets on one GPU, 8 host threads: 79.6 / 81.3 / 72.3 / 65.6 mesh-steps/s for N = 1 / 2 / 4 / 8 —
  //  no gain, dropped; profiles/r01_batch_graph.txt.)
  if (cap > FB_MAX_PARTIALS) cap = FB_MAX_PARTIALS;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <int G>
void plan_generic(fb_context *c) {
  const size_t gpb = SPMV_TB / G;
  const size_t want = ((size_t)c->nV + gpb - 1) / gpb;
  c->grid_spmv[0] = one_wave(c, k_spmv<G, 0>, SPMV_TB, want);
  c->grid_spmv[1] = one_wave(c, k_spmv<G, 1>, SPMV_TB, want);
  c->grid_spmv[2] = one_wave(c, k_spmv<G, 2>, SPMV_TB, want);
  c->grid_spmv[3] = 0;
}

template <int MODE>
void launch_spmv_mode(fb_context *c, const double *A, const double *x, double *y, const double *b, double *outp,
                      const FbPeerArgs *peer = nullptr) {
  FbPeerArgs pa;
  if (peer) pa = *peer; else memset(&pa, 0, sizeof(pa));
  const int grid = c->grid_spmv[MODE];
  // MODE 0 is the plain product of the inspection calls: all rows.  The solver's products visit owned rows only.
  const int lo = (MODE == 0) ? 0 : c->row_lo, hi = (MODE == 0) ? c->nV : c->row_hi;
  if (c->use_rows3) {
    if (c->l2_evict && MODE != 0)
      fb_launch(c->pdl != 0, c->stream, k_spmv_rows3<MODE, 4, true>, grid, SPMV_TB, lo, hi, c->bp, c->bc, A, x, y, c->rowmask, b, c->invD, c->sc, c->partials, outp, pa);
    else
      fb_launch(c->pdl && MODE != 0, c->stream, k_spmv_rows3<MODE, 4>, grid, SPMV_TB, lo, hi, c->bp, c->bc, A, x, y, c->rowmask, b, c->invD, c->sc, c->partials, outp, pa);
  } else if (MODE != 3) {
    constexpr int M = MODE == 3 ? 1 : MODE;
    switch (c->spmv_group) {
      case 8: k_spmv<8, M><<<grid, SPMV_TB, 0, c->stream>>>(lo, hi, c->bp, c->bc, A, x, y, c->rowmask, b, c->invD, c->sc, c->partials, outp, pa); break;
      case 32: k_spmv<32, M><<<grid, SPMV_TB, 0, c->stream>>>(lo, hi, c->bp, c->bc, A, x, y, c->rowmask, b, c->invD, c->sc, c->partials, outp, pa); break;
      default: k_spmv<16, M><<<grid, SPMV_TB, 0, c->stream>>>(lo, hi, c->bp, c->bc, A, x, y, c->rowmask, b, c->invD, c->sc, c->partials, outp, pa); break;
    }
  }
  c->launches++;
}

// partitioned context with peer-memory exchange: same three kernels, no NCCL call and no extra launch for the sums —
// producers store their total into every rank's comm block, consumers wait on their own flags (fb_pcg_common.cuh);
// the halo of d is pushed into the neighbours' ghost entries by one small kernel and awaited at the top of the SpMV.
void enqueue_iteration_p2p(fb_context *c, int it) {
  const int n = c->r, vg = c->grid_vec;
  double *slotsV = c->partials + 3 * (size_t)FB_MAX_PARTIALS;
  FbPeerArgs base, pa;
  fb_dist_peer_args(c, &base);
  // q = A d: waits for halo(it-1), publishes d.q(it)
  pa = base;
  pa.haloMask = fb_dist_halo_mask(c);
  pa.epochWait = fb_dist_epoch(c, it - 1, FB_COMM_HALO);
  pa.epoch = fb_dist_epoch(c, it, FB_COMM_DQ);
  const bool sample = c->profiling && (it % 16 == 1) && c->nprof < 64;  // includes the wait for the neighbours' halo
  if (sample) cudaEventRecord(c->evProf[2 * c->nprof], c->stream);
  launch_spmv_mode<1>(c, c->Keff, c->dir, c->Ad, c->rhs, nullptr, &pa);
  if (sample) { cudaEventRecord(c->evProf[2 * c->nprof + 1], c->stream); c->nprof++; }
  // x, r update: collects d.q(it), publishes rho'(it) (or, on refresh iterations, the SpMV after it does)
  pa = base;
  pa.epochWait = fb_dist_epoch(c, it, FB_COMM_DQ);
  pa.epoch = fb_dist_epoch(c, it, FB_COMM_RHO);
  if (it % 30 == 0) {
    fb_launch(c->pdl, c->stream, k_update<true>, vg, VEC_TB, n, c->dir, c->Ad, c->invD, c->x, c->res, c->sc, slotsV, it, nullptr, nullptr, 0, pa);
    c->launches++;
    FbPeerArgs pr = base;
    pr.epoch = fb_dist_epoch(c, it, FB_COMM_RHO);
    launch_spmv_mode<2>(c, c->Keff, c->x, c->res, c->rhs, nullptr, &pr);
  } else {
    fb_launch(c->pdl, c->stream, k_update<false>, vg, VEC_TB, n, c->dir, c->Ad, c->invD, c->x, c->res, c->sc, slotsV, it, nullptr, nullptr, 0, pa);
    c->launches++;
  }
  // direction: collects rho'(it); ghost entries are left to the neighbours
  pa = base;
  pa.epochWait = fb_dist_epoch(c, it, FB_COMM_RHO);
  FbPushArgs push;
  fb_dist_push_args(c, &push, fb_dist_epoch(c, it, FB_COMM_HALO));  // the halo of the new d leaves from this kernel
  fb_launch(c->pdl, c->stream, k_direction, vg, VEC_TB, n, c->res, c->invD, c->dir, c->sc, it, nullptr, 0, pa, c->rowmask, push);
  c->launches++;
}

// the reference's literal order, three kernels (+ NCCL in partitioned contexts without peer mapping)
// fromCounter: the kernels take the iteration number from the device counter instead of an argument, so that a captured
// 30-iteration period can be replayed (`it` then only places the refresh)
int enqueue_iteration_kernels(fb_context *c, int it, bool fromCounter = false) {
  if (fb_dist_p2p(c)) { enqueue_iteration_p2p(c, it); return FB_OK; }
  const int itArg = fromCounter ? 0 : it;
  const int n = c->r, vg = c->grid_vec;
  double *slotsV = c->partials + 3 * (size_t)FB_MAX_PARTIALS;
  FbPeerArgs nopeer;
  memset(&nopeer, 0, sizeof(nopeer));
  FbPushArgs nopush;
  memset(&nopush, 0, sizeof(nopush));
  const bool sample = c->profiling && (it % 16 == 1) && c->nprof < 64;
  if (sample) cudaEventRecord(c->evProf[2 * c->nprof], c->stream);
  // one GPU: per-CTA sums stay in their slots and the next kernel adds them (no ticket / last-CTA pass on the SpMV tail);
  // partitioned: the sums must become one scalar for ncclAllReduce, so the last-CTA pass is kept
  const bool defer = !c->dist;
  double *dqOut = c->dist ? &c->sc->dq_part : nullptr;
  double *rhoOut = c->dist ? &c->sc->rho_part : nullptr;
  const double *dqSlots = defer ? c->partials : nullptr;
  const double *rhoSlots = nullptr;
  int nRhoSlots = 0;
  const bool sym = c->sym_want && c->sym, tma = c->tma_want && c->tma;
  const int nDq = sym ? fb_sym_grid(c, 1) : (tma ? fb_tma_grid(c) : c->grid_spmv[1]);
  if (sym) fb_sym_launch(c, 1, c->dir, c->Ad, c->rhs, c->partials);
  else if (tma) fb_tma_launch(c, 1, c->dir, c->Ad, c->rhs, c->partials);
  else launch_spmv_mode<1>(c, c->Keff, c->dir, c->Ad, c->rhs, dqOut);
  if (sample) { cudaEventRecord(c->evProf[2 * c->nprof + 1], c->stream); c->nprof++; }
  if (c->dist) FB_TRY(fb_dist_allreduce_scalar(c, &c->sc->dq_part, &c->sc->dq));
  if (it % 30 == 0) {
    fb_launch(c->pdl, c->stream, k_update<true>, vg, VEC_TB, n, c->dir, c->Ad, c->invD, c->x, c->res, c->sc, slotsV, itArg, rhoOut, dqSlots, nDq, nopeer);
    c->launches++;
    if (sym) fb_sym_launch(c, 2, c->x, c->res, c->rhs, c->partials);
    else if (tma) fb_tma_launch(c, 2, c->x, c->res, c->rhs, c->partials);
    else launch_spmv_mode<2>(c, c->Keff, c->x, c->res, c->rhs, rhoOut);
    if (defer) { rhoSlots = c->partials; nRhoSlots = sym ? fb_sym_grid(c, 2) : (tma ? fb_tma_grid(c) : c->grid_spmv[2]); }
  } else {
    fb_launch(c->pdl, c->stream, k_update<false>, vg, VEC_TB, n, c->dir, c->Ad, c->invD, c->x, c->res, c->sc, slotsV, itArg, rhoOut, dqSlots, nDq, nopeer);
    c->launches++;
    if (defer) { rhoSlots = slotsV; nRhoSlots = vg; }
  }
  if (c->dist) FB_TRY(fb_dist_allreduce_scalar(c, &c->sc->rho_part, &c->sc->rho[it & 1]));
  fb_launch(c->pdl, c->stream, k_direction, vg, VEC_TB, n, c->res, c->invD, c->dir, c->sc, itArg, rhoSlots, nRhoSlots, nopeer, nullptr, nopush);
  c->launches++;
  if (c->dist) FB_TRY(fb_dist_halo_exchange(c, c->dir));
  return FB_OK;
}

// two kernels; `it` is only used to place the refresh (the kernels read the iteration from the device counter,
// so a captured period can be replayed).  Returns the number of kernels enqueued.
int enqueue_iteration_fused(fb_context *c, int it, bool allowSample) {
  const int n = c->r, vg = c->grid_vec;
  FbPeerArgs nopeer;
  memset(&nopeer, 0, sizeof(nopeer));
  FbPushArgs nopush;
  memset(&nopush, 0, sizeof(nopush));
  double *slotsV = c->partials + 3 * (size_t)FB_MAX_PARTIALS;
  const bool sample = allowSample && c->profiling && (it % 16 == 1) && c->nprof < 64;
  if (sample) cudaEventRecord(c->evProf[2 * c->nprof], c->stream);
  launch_spmv_mode<3>(c, c->Keff, c->dir, c->Ad, c->res, nullptr);
  if (sample) { cudaEventRecord(c->evProf[2 * c->nprof + 1], c->stream); c->nprof++; }
  if (it % 30 == 0) {
    fb_launch(c->pdl, c->stream, k_update<true>, vg, VEC_TB, n, c->dir, c->Ad, c->invD, c->x, c->res, c->sc, slotsV, 0, nullptr, nullptr, 0, nopeer);
    c->launches++;
    launch_spmv_mode<2>(c, c->Keff, c->x, c->res, c->rhs, &c->sc->rho[0]);  // it is even: rho[it & 1] = rho[0]
    fb_launch(c->pdl, c->stream, k_direction, vg, VEC_TB, n, c->res, c->invD, c->dir, c->sc, 0, nullptr, 0, nopeer, nullptr, nopush);
    c->launches++;
    return 4;
  }
  fb_launch(c->pdl, c->stream, k_fused_update, vg, VEC_TB, n, c->Ad, c->invD, c->x, c->res, c->dir, c->sc, slotsV);
  c->launches++;
  return 2;
}

int finish_solve(fb_context *c) {
  cudaStream_t st = c->stream;
  FB_CUDA(cudaMemcpyAsync(&c->sc_host[2], c->sc, sizeof(FbScalars), cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  FB_CUDA(cudaGetLastError());
  for (int i = 0; i < c->nprof; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, c->evProf[2 * i], c->evProf[2 * i + 1]) == cudaSuccess) { c->prof_sum_s += 1e-3 * ms; c->prof_samples++; }
  }
  c->nprof = 0;
  const FbScalars &s = c->sc_host[2];
  const double rhoFinal = s.rho[s.iters & 1];
  const bool notConverged = rhoFinal > s.eps2 * s.rho0;
  c->last_iters = s.iters * (notConverged ? -1 : 1);
  c->last_ratio = (s.rho0 != 0.0) ? rhoFinal / s.rho0 : 0.0;
  if (c->tma_want && c->tma && fb_tma_failed(c)) {
    fb_set_error("bulk-copy staged SpMV: an mbarrier wait ran out (FEMBRAIN_B200_SPMV=tma is experimental)");
    return FB_ERR_CUDA;
  }
  if (c->dist && s.comm_error) {
    // CTAs that saw `done` raised in mid-kernel left without handing in their tickets: reset them before the next solve
    c->comm_poisoned = 1;
    fb_set_error(s.comm_error == 2 ? "peer-memory exchange overrun (a rank published a later value before this one was collected)"
                                   : "peer-memory exchange timed out (a rank stopped publishing)");
    return FB_ERR_COMM;
  }
  return FB_OK;
}

int start_solve(fb_context *c, double eps, int maxIt) {
  cudaStream_t st = c->stream;
  FbPeerArgs pa;
  memset(&pa, 0, sizeof(pa));
  const bool p2p = fb_dist_p2p(c) != 0;
  if (c->dist) fb_dist_next_solve(c);
  if (c->comm_poisoned) {  // the previous solve ended in FB_ERR_COMM: last-block tickets may be half counted
    FB_CUDA(cudaMemsetAsync(&c->sc->ticket_a, 0, 2 * sizeof(unsigned int), st));
    FB_TRY(fb_dist_reset_tickets(c));
    c->comm_poisoned = 0;
  }
  // products from the block-upper triangle (fb_sym.cu): single-GPU three-kernel schedule only
  if (c->sym_want && (c->dist || c->batch || c->pers_grid > 0 || c->pcg_fused || c->pcg_graph)) c->sym_want = 0;
  if (c->sym_want && !c->sym) {
    FB_TRY(fb_sym_plan(c));
    if (!c->sym) c->sym_want = 0;
  }
  if (c->sym_want) FB_TRY(fb_sym_pack(c));
  // products with the matrix staged through shared memory by the bulk-copy engine (fb_tma.cu, experimental): same scope
  if (c->tma_want && (c->dist || c->batch || c->pers_grid > 0 || c->pcg_fused || c->pcg_graph)) c->tma_want = 0;
  if (c->tma_want && !c->tma) {
    FB_TRY(fb_tma_plan(c));
    if (!c->tma) c->tma_want = 0;
  }
  if (p2p) {
    fb_dist_peer_args(c, &pa);
    pa.epoch = fb_dist_epoch(c, 0, FB_COMM_RHO);
  }
  k_cg_init<<<c->grid_vec, VEC_TB, 0, st>>>(c->r, c->rhs, c->invD, c->x, c->res, c->dir, c->Ad, c->sc, c->partials + 3 * (size_t)FB_MAX_PARTIALS,
                                            c->dist ? &c->sc->rho_part : &c->sc->rho[0], pa, p2p ? c->rowmask : nullptr);
  if (c->dist && !p2p) FB_TRY(fb_dist_allreduce_scalar(c, &c->sc->rho_part, &c->sc->rho[0]));  // rho0 is a global sum
  if (p2p) { pa.epoch = 0; pa.epochWait = fb_dist_epoch(c, 0, FB_COMM_RHO); }
  k_cg_begin<<<1, 32, 0, st>>>(c->sc, eps, maxIt, pa);
  c->launches += 2;
  // ghost entries of d = invD r live on the neighbours
  if (p2p) FB_TRY(fb_dist_halo_push(c, c->dir, fb_dist_epoch(c, 0, FB_COMM_HALO)));
  else if (c->dist) FB_TRY(fb_dist_halo_exchange(c, c->dir));
  return FB_OK;
}

// one 30-iteration period of the fused schedule as a CUDA graph (built lazily, once per context)
int ensure_period_graph(fb_context *c) {
  if (c->graph_exec || c->graph_failed) return FB_OK;
  cudaStream_t st = c->stream;
  cudaGraph_t graph = nullptr;
  const long long before = c->launches;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); c->graph_failed = 1; return FB_OK; }
  int kernels = 0;
  if (c->pcg_fused) {
    for (int k = 1; k <= 30; k++) kernels += enqueue_iteration_fused(c, k, false);
  } else {
    for (int k = 1; k <= 30; k++) (void)enqueue_iteration_kernels(c, k, true);  // no NCCL in a captured period (single GPU only)
    kernels = (int)(c->launches - before);
  }
  cudaError_t e = cudaStreamEndCapture(st, &graph);
  c->launches = before;  // captured, not launched
  if (e != cudaSuccess || !graph) { cudaGetLastError(); c->graph_failed = 1; return FB_OK; }
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) { cudaGetLastError(); c->graph_failed = 1; return FB_OK; }
  c->graph_exec = exec;
  c->graph_kernels = kernels;
  return FB_OK;
}

}  // namespace

// Chooses the SpMV variant and the launch shapes for this mesh.  Measured on B200, 998,250-tet cube
// (profiles/r01_spmv_variants.txt): k_spmv<16> 45.0 us in step / 36.8 us isolated; k_spmv_rows3 at 48 registers
// (5 CTAs/SM) 42.8 / 36.6 us; forcing 40 or 32 registers spills and is slower (54 / 66 us); a shared-memory-tiled
// variant with x staged per 64-row tile was slower too (60 us: three block-wide barriers per tile, 3 CTAs/SM) and
// was removed.
int fb_spmv_plan(fb_context *c) {
  const char *env = getenv("FEMBRAIN_B200_SPMV");
  c->use_rows3 = (c->spmv_group == 16) && !(env && !strcmp(env, "rows"));
  c->sym_want = env && !strcmp(env, "sym");
  c->tma_want = env && !strcmp(env, "tma");
  const char *ev = getenv("FEMBRAIN_B200_L2EVICT");
  c->l2_evict = ev && atoi(ev) != 0;  // matrix loads of k_spmv_rows3 with an L2 evict-first hint (not measured yet)
  const size_t n = (size_t)c->r;
  // vector kernels: at most one resident wave, one 16-byte item per thread on small meshes.  Fatter CTAs (2/4/8 items per
  // thread, fewer CTAs adding the producer's slots) were measured slower: 51.3 / 52.8 / 55.7 us per iteration at 1M tets,
  // 22.0 / 24.5 / 28.8 at 200k (profiles/r01_vec_items.txt) — these kernels are latency bound, more CTAs hide more of it
  const char *vi = getenv("FEMBRAIN_B200_VEC_ITEMS");
  const size_t items = (vi && atoi(vi) > 0) ? (size_t)atoi(vi) : 1;
  const size_t wantV = ((n >> 1) + VEC_TB * items - 1) / (VEC_TB * items);
  int gv = one_wave(c, k_fused_update, VEC_TB, wantV);
  gv = min(gv, one_wave(c, k_update<false>, VEC_TB, wantV));
  gv = min(gv, one_wave(c, k_direction, VEC_TB, wantV));
  gv = min(gv, one_wave(c, k_cg_init, VEC_TB, wantV));
  c->grid_vec = gv;
  if (c->use_rows3) {
    const size_t gpb = SPMV_TB / TILE_G;
    const size_t want = ((size_t)c->nV + gpb - 1) / gpb;
    // 4 CTAs/SM (<= 64 registers, no spills) measured best; 5 CTAs/SM (48 registers) spills (profiles/r01_pcg_schedules.txt)
    c->rows3_minb = 4;
    c->grid_spmv[0] = one_wave(c, k_spmv_rows3<0, 4>, SPMV_TB, want);
    c->grid_spmv[1] = one_wave(c, k_spmv_rows3<1, 4>, SPMV_TB, want);
    c->grid_spmv[2] = one_wave(c, k_spmv_rows3<2, 4>, SPMV_TB, want);
    c->grid_spmv[3] = one_wave(c, k_spmv_rows3<3, 4>, SPMV_TB, want);
  } else if (c->spmv_group == 8) {
    plan_generic<8>(c);
  } else if (c->spmv_group == 32) {
    plan_generic<32>(c);
  } else {
    plan_generic<16>(c);
  }
  const char *mode = getenv("FEMBRAIN_B200_PCG");
  // Default: the reference's literal order (three kernels).  Measured on B200 after the row-owner prefetch
  // (profiles/r01_pcg_schedules.txt): kernels vs fused, us per iteration in step: 26.7 / 26.3 at 200k tets, 56.4 / 56.0
  // at 1M, 497.6 / 525.8 at 10M — the extra reads and the three-sum epilogue of k_spmv_rows3<3> cost what the saved
  // launch gains, and graph replay changes nothing (the gaps are device-side dependencies, not host launch cost).
  c->pcg_fused = c->use_rows3 && mode && (!strcmp(mode, "fused") || !strcmp(mode, "fused_nograph"));
  // CUDA-graph replay of a 30-iteration period of the three-kernel schedule (3 x 30 + 1 kernels per cudaGraphLaunch),
  // opt-in with FEMBRAIN_B200_PCG_GRAPH=1.  Measured on B200 (profiles/r01_batch_graph.txt): the isolated iteration loop
  // gains (19.4 -> 17.9 us at 200 k tets, 50.1 -> 48.7 at 1M) but a step does not (13.05 vs 13.07 ms at 200 k tets: the
  // gaps inside a step are device-side dependencies), nor does the batch of 32 concurrent meshes (81.3 vs 78.9 mesh-steps/s),
  // which therefore is not bound by the host's launch rate.
  const char *gr = getenv("FEMBRAIN_B200_PCG_GRAPH");
  const bool graphKernels = gr && atoi(gr) != 0;
  c->pcg_graph = c->pcg_fused ? !(mode && !strcmp(mode, "fused_nograph")) : graphKernels;
  const char *pdl = getenv("FEMBRAIN_B200_PDL");
  c->pdl = !(pdl && atoi(pdl) == 0);
  return fb_pcg_plan_persistent(c);
}

// hooks for the solver variants (fb_mg.cu): the FP64 products of this file with their per-CTA sums left in c->partials
int fb_pcg_launch_product_dq(fb_context *c, const double *d, double *q, int *nSlots) {
  launch_spmv_mode<1>(c, c->Keff, d, q, c->rhs, nullptr);
  *nSlots = c->grid_spmv[1];
  return FB_OK;
}
int fb_pcg_launch_residual(fb_context *c, const double *x, double *r) {
  launch_spmv_mode<2>(c, c->Keff, x, r, c->rhs, nullptr);
  return FB_OK;
}

int fb_launch_spmv(fb_context *c, const double *A, const double *x, double *y, bool masked) {
  (void)masked;
  if (c->nV == 0) return FB_OK;
  launch_spmv_mode<0>(c, A, x, y, c->rhs, nullptr);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

void fb_pcg_release(fb_context *c) {
  if (c->graph_exec) { cudaGraphExecDestroy((cudaGraphExec_t)c->graph_exec); c->graph_exec = nullptr; }
}

// Solves Keff x = rhs on the constrained DOFs, x0 = 0.  On return c->last_iters holds the
// reference's return value: +iterations if converged, -iterations otherwise (CGSolver.cpp:189).
int fb_pcg_solve(fb_context *c, double eps, int maxIt) {
  if (c->batch) return fb_batch_pcg_solve(c, eps, maxIt);
  if (fb_mg_active(c)) return fb_mg_pcg_solve(c, eps, maxIt);   // labelled variants: same system, same stopping rule
  cudaStream_t st = c->stream;
  if (c->r == 0) { c->last_iters = 0; c->last_ratio = 0.0; return FB_OK; }
  c->nprof = 0;
  FB_TRY(start_solve(c, eps, maxIt));
  if (c->pers_grid > 0 && !c->dist) {  // opt-in: one cooperative kernel runs the whole loop (fb_pcg_persistent.cu)
    FB_TRY(fb_pcg_launch_persistent(c));
    return finish_solve(c);
  }
  const bool fused = c->pcg_fused && !c->dist;
  const bool useGraph = c->pcg_graph && !c->dist && !c->profiling;
  if (useGraph) FB_TRY(ensure_period_graph(c));
  // Iterations are enqueued in chunks of one refresh period; the loop condition lives on the device (kernels turn
  // into no-ops once `done` is set).  The host looks at the flag of chunk k-1 while chunk k runs.
  const int CH = 30;
  int it = 1, slot = 0, pending = 0;
  bool finished = false;
  while (!finished && it <= maxIt) {
    const int end = (it + CH - 1 < maxIt) ? it + CH - 1 : maxIt;
    if (useGraph && c->graph_exec && end - it + 1 == CH) {
      FB_CUDA(cudaGraphLaunch((cudaGraphExec_t)c->graph_exec, st));
      c->launches += c->graph_kernels;
      it += CH;
    } else {
      for (; it <= end; it++) {
        if (fused) enqueue_iteration_fused(c, it, true);
        else FB_TRY(enqueue_iteration_kernels(c, it));
      }
    }
    FB_CUDA(cudaMemcpyAsync(&c->sc_host[slot], c->sc, sizeof(FbScalars), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaEventRecord(c->evChunk[slot], st));
    pending++;
    if (pending == 2) {
      const int prev = slot ^ 1;
      FB_CUDA(cudaEventSynchronize(c->evChunk[prev]));
      if (c->sc_host[prev].done) finished = true;
      pending--;
    }
    slot ^= 1;
  }
  return finish_solve(c);
}

int fb_pcg_bench_iteration(fb_context *c, int repeats, double *sec) {
  if (c->batch) { fb_set_error("fb_bench_cg_iteration on a batch context"); return FB_ERR_NOT_SUPPORTED; }
  // time `repeats` full CG iterations on the current system with the stopping rule disabled (eps = 0)
  if (c->r == 0 || repeats <= 0) { *sec = 0.0; return FB_OK; }
  cudaStream_t st = c->stream;
  const int warm = 30;
  repeats = ((repeats + 29) / 30) * 30;  // whole refresh periods
  FB_TRY(start_solve(c, 0.0, (c->pers_grid > 0 && !c->dist) ? repeats : (1 << 30)));
  const bool fused = c->pcg_fused && !c->dist;
  if (c->pers_grid > 0 && !c->dist) {
    FB_CUDA(cudaEventRecord(c->ev[3], st));
    FB_TRY(fb_pcg_launch_persistent(c));
  } else {
    const bool useGraph = c->pcg_graph && !c->dist;
    if (useGraph) FB_TRY(ensure_period_graph(c));
    int it = 1;
    auto run = [&](int count) -> int {
      const int end = it + count - 1;
      while (it <= end) {
        if (useGraph && c->graph_exec && (it - 1) % 30 == 0 && end - it + 1 >= 30) {
          FB_CUDA(cudaGraphLaunch((cudaGraphExec_t)c->graph_exec, st));
          c->launches += c->graph_kernels;
          it += 30;
        } else {
          if (fused) enqueue_iteration_fused(c, it, false);
          else FB_TRY(enqueue_iteration_kernels(c, it));
          it++;
        }
      }
      return FB_OK;
    };
    FB_TRY(run(warm));
    FB_CUDA(cudaEventRecord(c->ev[3], st));
    FB_TRY(run(repeats));
  }
  FB_CUDA(cudaEventRecord(c->ev[7], st));
  FB_CUDA(cudaStreamSynchronize(st));
  FB_CUDA(cudaGetLastError());
  float ms = 0;
  FB_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[7]));
  *sec = 1e-3 * ms / repeats;
  return FB_OK;
}
