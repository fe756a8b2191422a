// fb_pcg.cu — Jacobi-preconditioned conjugate gradients on the constrained effective matrix.
//
// Reference: CGSolver::SolveLinearSystemWithJacobiPreconditioner
// (src/3rdparty/vegafem/sparseSolver/CGSolver.cpp:129-190; SpMV sparseMatrix/sparseMatrix.cpp:405-413):
//   invD = 1/diag(A); r = b - A x0 (x0 = 0 => r = b); d = invD r; rho = sum r^2 invD; rho0 = rho
//   while rho > eps^2 rho0 and it <= maxIt:
//     q = A d; alpha = rho / (d.q); x += alpha d
//     it % 30 == 0 ? r = b - A x : r -= alpha q
//     rho' = sum r^2 invD; beta = rho'/rho; d = invD r + beta d
// Same recurrences, same refresh period, same stopping rule; sums are parallel reductions with a
// FIXED association order (deterministic run to run), so iterates agree with the reference's
// sequential sums to rounding, not bitwise.
//
// The constrained system (fixed rows/columns removed, CGSolver on systemMatrix) is solved IN PLACE
// on full-length vectors: rows of constrained DOFs produce 0 and their entries of d, r, x stay 0,
// which contributes exact zeros to every sum — arithmetically the compacted system, without the
// per-step gather of AssignSuperMatrix (sparseMatrix.cpp:993-1002).
//
// Three kernels per iteration, all HBM-streaming, scalars (alpha, beta, rho, loop condition)
// stay on the device:
//   k_spmv_cg     q = A d  (+ d.q)            A values + block columns streamed once, d gathered
//   k_update      x += alpha d; r -= alpha q  (+ rho')
//   k_direction   d = invD r + beta d         (+ loop bookkeeping)
// Block-level partial sums go to a fixed slot per CTA; the last CTA to finish (integer ticket)
// adds the slots in index order.
//
// Matrix layout: the reference's CSR value order with 3x3-block-compressed column indices:
// block row v owns 9*nb doubles at 9*bp[v]: three scalar rows of 3*nb values each; column of entry
// t of a scalar row is 3*bc[bp[v] + t/3] + t%3.  8.44 bytes per nonzero instead of CSR's 12.
#include <cstdlib>
#include <cstring>

#include "fb_internal.h"
#include "fb_pcg_common.cuh"

namespace {

constexpr int SPMV_TB = 256;
constexpr int VEC_TB = 256;

// Deterministic block reduction followed by the "last block adds all slots in order" pattern.
// Returns true in every thread of the last block; *total is then valid in thread 0.
template <int TB>
__device__ __forceinline__ bool block_reduce_to_total(double v, double *slots, unsigned int *ticket, double *total) {
  __shared__ double wsum[TB / 32];
  __shared__ bool isLast;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double s = (lane < TB / 32) ? wsum[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) {
      slots[blockIdx.x] = s;
      __threadfence();
      unsigned int tk = atomicAdd(ticket, 1u);
      isLast = (tk == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!isLast) return false;
  __threadfence();
  // fixed-order final sum: thread t adds slots t, t+TB, ...; then the same block tree
  double s = 0.0;
  for (unsigned int i = threadIdx.x; i < gridDim.x; i += TB) s += ((volatile double *)slots)[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  __syncthreads();
  if (lane == 0) wsum[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double z = (lane < TB / 32) ? wsum[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_down_sync(0xffffffffu, z, o);
    if (lane == 0) {
      *total = z;
      *ticket = 0u;
    }
  }
  return true;
}

// MODE 0: y = A x                     (no mask, no reduction)       — generic product
// MODE 1: y = mask(A x), sum x.y      (x = d, y = q)                — CG iteration
// MODE 2: y = mask(b - A x), sum y^2 invD   (x = x, y = r)          — exact-residual refresh
template <int G, int MODE>
__global__ void __launch_bounds__(SPMV_TB) k_spmv(int nV, const int *__restrict__ bp, const int *__restrict__ bc,
                                                  const double *__restrict__ A, const double *__restrict__ x,
                                                  double *__restrict__ y, const unsigned char *__restrict__ fixed,
                                                  const double *__restrict__ b, const double *__restrict__ invD,
                                                  FbScalars *sc, double *slots, double *outp) {
  if (MODE != 0) {
    if (sc->done) return;
  }
  const int lane = threadIdx.x & (G - 1);
  // the G lanes of a group always take the same trips through the row loop; other groups of the warp may
  // not, so shuffles name only the group's own lanes
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  const int groupsPerBlock = SPMV_TB / G;
  const int group = blockIdx.x * groupsPerBlock + threadIdx.x / G;
  const int nGroups = gridDim.x * groupsPerBlock;
  double part = 0.0;
  for (int v = group; v < nV; v += nGroups) {
    const int rs = __ldg(bp + v), re = __ldg(bp + v + 1);
    const int n3 = 3 * (re - rs);
    const double *a0 = A + 9 * (size_t)rs;
    const double *a1 = a0 + n3;
    const double *a2 = a1 + n3;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    int t = lane;
    // two passes per trip: six independent streaming loads in flight per lane
    for (; t + G < n3; t += 2 * G) {
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
    if (lane < 3) {
      double s = (lane == 0) ? acc0 : ((lane == 1) ? acc1 : acc2);
      const size_t row = 3 * (size_t)v + lane;
      if (MODE == 0) {
        y[row] = s;
      } else if (MODE == 1) {
        if (fixed[row]) s = 0.0;
        y[row] = s;
        part = fma(x[row], s, part);
      } else {
        double rr = fixed[row] ? 0.0 : (b[row] - s);
        y[row] = rr;
        part += (rr * rr) * invD[row];
      }
    }
  }
  if (MODE == 1) {
    double total;
    if (block_reduce_to_total<SPMV_TB>(part, slots, &sc->ticket_a, &total) && threadIdx.x == 0) *outp = total;
  } else if (MODE == 2) {
    double total;
    if (block_reduce_to_total<SPMV_TB>(part, slots, &sc->ticket_a, &total) && threadIdx.x == 0) *outp = total;
  }
}

// ---- row-per-16-lanes SpMV with every load of a row in flight at once (no shared memory, no block syncs) ----
// Same mapping as k_spmv<16,*>, but the three 16-wide passes over a row are fully unrolled and predicated, so a
// lane has 9 streaming value loads + 3 column loads outstanding before the first multiply, and the row pointers
// of the group's next row are fetched one row ahead.  MINB = resident CTAs per SM requested from the compiler.
template <int MODE, int MINB>
__global__ void __launch_bounds__(SPMV_TB, MINB) k_spmv_rows3(int nV, const int *__restrict__ bp, const int *__restrict__ bc,
                                                              const double *__restrict__ A, const double *__restrict__ x,
                                                              double *__restrict__ y, const unsigned char *__restrict__ fixed,
                                                              const double *__restrict__ b, const double *__restrict__ invD,
                                                              FbScalars *sc, double *slots, double *outp) {
  if (MODE != 0) {
    if (sc->done) return;
  }
  const int lane = threadIdx.x & (TILE_G - 1);
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int groupsPerBlock = SPMV_TB / TILE_G;
  const int group = blockIdx.x * groupsPerBlock + threadIdx.x / TILE_G;
  const int nGroups = gridDim.x * groupsPerBlock;
  double part = 0.0;
  int v = group;
  int rs = 0, re = 0;
  if (v < nV) { rs = __ldg(bp + v); re = __ldg(bp + v + 1); }
  while (v < nV) {
    const int vn = v + nGroups;
    int rsn = 0, ren = 0;
    if (vn < nV) { rsn = __ldg(bp + vn); ren = __ldg(bp + vn + 1); }
    const int n3 = 3 * (re - rs);
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
      const size_t row = 3 * (size_t)v + lane;
      if (MODE == 0) {
        y[row] = s;
      } else if (MODE == 1) {
        if (fixed[row]) s = 0.0;
        y[row] = s;
        part = fma(x[row], s, part);
      } else {
        const double rres = fixed[row] ? 0.0 : (b[row] - s);
        y[row] = rres;
        part += (rres * rres) * invD[row];
      }
    }
    v = vn; rs = rsn; re = ren;
  }
  if (MODE == 1) {
    double total;
    if (block_reduce_to_total<SPMV_TB>(part, slots, &sc->ticket_a, &total) && threadIdx.x == 0) *outp = total;
  } else if (MODE == 2) {
    double total;
    if (block_reduce_to_total<SPMV_TB>(part, slots, &sc->ticket_a, &total) && threadIdx.x == 0) *outp = total;
  }
}

template <int MODE, int MINB>
void launch_rows3(fb_context *c, const double *A, const double *x, double *y, double *outp) {
  static int perSM = 0;
  if (!perSM) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_rows3<MODE, MINB>, SPMV_TB, 0) != cudaSuccess || perSM < 1) perSM = 1;
  }
  const size_t groupsPerBlock = SPMV_TB / TILE_G;
  size_t want = ((size_t)c->nV + groupsPerBlock - 1) / groupsPerBlock;
  size_t cap = (size_t)c->sm_count * (size_t)perSM;
  if (cap > FB_MAX_PARTIALS) cap = FB_MAX_PARTIALS;
  const int grid = (int)(want < cap ? (want ? want : 1) : cap);
  k_spmv_rows3<MODE, MINB><<<grid, SPMV_TB, 0, c->stream>>>(c->nV, c->bp, c->bc, A, x, y, c->rowmask, c->rhs, c->invD, c->sc, c->partials, outp);
  c->launches++;
}

// r = b (x0 = 0), d = invD r, x = 0, rho0 = sum r^2 invD           (CGSolver.cpp:139-147)
__global__ void __launch_bounds__(VEC_TB) k_cg_init(int n, const double *__restrict__ b, const double *__restrict__ invD,
                                                    double *__restrict__ x, double *__restrict__ r, double *__restrict__ d,
                                                    FbScalars *sc, double *slots, double *outp) {
  double part = 0.0;
  for (size_t i = (size_t)blockIdx.x * VEC_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * VEC_TB) {
    const double bi = b[i], di = invD[i];
    x[i] = 0.0;
    r[i] = bi;
    d[i] = di * bi;
    part += (bi * bi) * di;
  }
  double total;
  if (block_reduce_to_total<VEC_TB>(part, slots, &sc->ticket_b, &total) && threadIdx.x == 0) *outp = total;
}

// after rho[0] is final (all-reduced in partitioned contexts): initial residual, loop condition at iteration 1
__global__ void k_cg_begin(FbScalars *sc, double eps, int maxIt) {
  const double total = sc->rho[0];
  sc->rho0 = total;
  sc->eps2 = eps * eps;
  sc->max_it = maxIt;
  sc->iters = 0;
  sc->dq = 0.0;
  // while ((residualNorm2 > eps*eps*initialResidualNorm2) && (iteration <= maxIterations)), iteration = 1
  sc->done = !((total > eps * eps * total) && (1 <= maxIt));
}

// x += alpha d; REFRESH ? nothing more : (r -= alpha q; rho' = sum r^2 invD)     (CGSolver.cpp:155-174)
template <bool REFRESH>
__global__ void __launch_bounds__(VEC_TB) k_update(int n, const double *__restrict__ d, const double *__restrict__ q,
                                                   const double *__restrict__ invD, double *__restrict__ x,
                                                   double *__restrict__ r, FbScalars *sc, double *slots, int it, double *outp) {
  if (sc->done) return;
  const double alpha = sc->rho[(it - 1) & 1] / sc->dq;
  double part = 0.0;
  for (size_t i = (size_t)blockIdx.x * VEC_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * VEC_TB) {
    const double di = d[i];
    x[i] = fma(alpha, di, x[i]);
    if (!REFRESH) {
      const double ri = fma(-alpha, q[i], r[i]);
      r[i] = ri;
      part += (ri * ri) * invD[i];
    }
  }
  if (!REFRESH) {
    double total;
    if (block_reduce_to_total<VEC_TB>(part, slots, &sc->ticket_b, &total) && threadIdx.x == 0) *outp = total;
  }
}

// beta = rho'/rho; d = invD r + beta d; iteration++ and loop condition            (CGSolver.cpp:176-183, 150)
__global__ void __launch_bounds__(VEC_TB) k_direction(int n, const double *__restrict__ r, const double *__restrict__ invD,
                                                      double *__restrict__ d, FbScalars *sc, int it) {
  if (sc->done) return;
  const double rhoNew = sc->rho[it & 1], rhoOld = sc->rho[(it - 1) & 1];
  const double beta = rhoNew / rhoOld;
  for (size_t i = (size_t)blockIdx.x * VEC_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * VEC_TB)
    d[i] = fma(invD[i], r[i], beta * d[i]);
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    sc->iters = it;
    // `done` is read at kernel entry by this kernel's other blocks too; a block that sees the new value
    // early only skips a direction update nobody will use.
    if (!((rhoNew > sc->eps2 * sc->rho0) && (it + 1 <= sc->max_it))) sc->done = 1;
  }
}

// Grids are sized to ONE resident wave: sm_count x (blocks of this kernel that fit on an SM), so the
// grid-stride loops see every SM equally loaded (no partial second wave) and the number of per-CTA
// partial sums stays small and fixed.
int vec_grid(const fb_context *c, size_t n) {
  static int perSM = 0;
  if (!perSM) {
    int a = 1, b = 1, d = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_update<false>, VEC_TB, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_direction, VEC_TB, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&d, k_cg_init, VEC_TB, 0);
    perSM = a < b ? a : b;
    if (d < perSM) perSM = d;
    if (perSM < 1) perSM = 1;
  }
  size_t want = (n + VEC_TB - 1) / VEC_TB;
  size_t cap = (size_t)c->sm_count * (size_t)perSM;
  if (cap > FB_MAX_PARTIALS) cap = FB_MAX_PARTIALS;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

template <int G, int MODE>
void launch_spmv_g(fb_context *c, const double *A, const double *x, double *y, double *outp) {
  static int perSM = 0;  // resident CTAs of this instantiation per SM (same for every B200)
  if (!perSM) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv<G, MODE>, SPMV_TB, 0) != cudaSuccess || perSM < 1) perSM = 1;
  }
  const size_t groupsPerBlock = SPMV_TB / G;
  size_t want = ((size_t)c->nV + groupsPerBlock - 1) / groupsPerBlock;
  size_t cap = (size_t)c->sm_count * (size_t)perSM;
  if (cap > FB_MAX_PARTIALS) cap = FB_MAX_PARTIALS;
  const int grid = (int)(want < cap ? (want ? want : 1) : cap);
  k_spmv<G, MODE><<<grid, SPMV_TB, 0, c->stream>>>(c->nV, c->bp, c->bc, A, x, y, c->rowmask, c->rhs, c->invD, c->sc,
                                                           c->partials, outp);
  c->launches++;
}

template <int MODE>
void launch_spmv_mode(fb_context *c, const double *A, const double *x, double *y, double *outp) {
  if (c->use_tiled) { launch_rows3<MODE, 5>(c, A, x, y, outp); return; }
  switch (c->spmv_group) {
    case 8: launch_spmv_g<8, MODE>(c, A, x, y, outp); break;
    case 32: launch_spmv_g<32, MODE>(c, A, x, y, outp); break;
    default: launch_spmv_g<16, MODE>(c, A, x, y, outp); break;
  }
}

void enqueue_iteration(fb_context *c, int it) {
  const int n = c->r;
  const int vg = vec_grid(c, (size_t)n);
  double *slotsB = c->partials + FB_MAX_PARTIALS;
  const bool sample = c->profiling && (it % 16 == 1) && c->nprof < 64;
  if (sample) cudaEventRecord(c->evProf[2 * c->nprof], c->stream);
  double *dqOut = c->dist ? &c->sc->dq_part : &c->sc->dq;
  double *rhoOut = c->dist ? &c->sc->rho_part : &c->sc->rho[it & 1];
  launch_spmv_mode<1>(c, c->Keff, c->dir, c->Ad, dqOut);
  if (sample) { cudaEventRecord(c->evProf[2 * c->nprof + 1], c->stream); c->nprof++; }
  if (c->dist) fb_dist_allreduce_scalar(c, &c->sc->dq_part, &c->sc->dq);
  if (it % 30 == 0) {
    k_update<true><<<vg, VEC_TB, 0, c->stream>>>(n, c->dir, c->Ad, c->invD, c->x, c->res, c->sc, slotsB, it, rhoOut);
    c->launches++;
    launch_spmv_mode<2>(c, c->Keff, c->x, c->res, rhoOut);
  } else {
    k_update<false><<<vg, VEC_TB, 0, c->stream>>>(n, c->dir, c->Ad, c->invD, c->x, c->res, c->sc, slotsB, it, rhoOut);
    c->launches++;
  }
  if (c->dist) fb_dist_allreduce_scalar(c, &c->sc->rho_part, &c->sc->rho[it & 1]);
  k_direction<<<vg, VEC_TB, 0, c->stream>>>(n, c->res, c->invD, c->dir, c->sc, it);
  c->launches++;
  if (c->dist) fb_dist_halo_exchange(c, c->dir);
}

}  // namespace

// Chooses the SpMV variant for this mesh.  Measured on B200, 998,250-tet cube (profiles/r01_spmv_variants.txt):
// k_spmv<16> 45.0 us in step / 36.8 us isolated; k_spmv_rows3 at 48 registers (5 CTAs/SM) 42.8 / 36.6 us; forcing
// 40 or 32 registers spills and is slower (54 / 66 us); a shared-memory-tiled variant with x staged per 64-row tile
// was slower too (60 us: three block-wide barriers per tile, 3 CTAs/SM) and was removed.
int fb_spmv_plan(fb_context *c) {
  const char *env = getenv("FEMBRAIN_B200_SPMV");
  c->use_tiled = (c->spmv_group == 16) && !(env && !strcmp(env, "rows"));
  return fb_pcg_plan_persistent(c);
}

int fb_launch_spmv(fb_context *c, const double *A, const double *x, double *y, bool masked) {
  (void)masked;
  if (c->nV == 0) return FB_OK;
  launch_spmv_mode<0>(c, A, x, y, nullptr);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

// Solves Keff x = rhs on the constrained DOFs, x0 = 0.  On return c->last_iters holds the
// reference's return value: +iterations if converged, -iterations otherwise (CGSolver.cpp:189).
int fb_pcg_solve(fb_context *c, double eps, int maxIt) {
  const int n = c->r;
  cudaStream_t st = c->stream;
  if (n == 0) { c->last_iters = 0; c->last_ratio = 0.0; return FB_OK; }
  const int vg = vec_grid(c, (size_t)n);
  if (c->pers_grid > 0 && !c->dist) {
    // one cooperative kernel runs the whole loop (fb_pcg_persistent.cu)
    k_cg_init<<<vg, VEC_TB, 0, st>>>(n, c->rhs, c->invD, c->x, c->res, c->dir, c->sc, c->partials + FB_MAX_PARTIALS, &c->sc->rho[0]);
    k_cg_begin<<<1, 1, 0, st>>>(c->sc, eps, maxIt);
    c->launches += 2;
    FB_TRY(fb_pcg_launch_persistent(c));
    FB_CUDA(cudaMemcpyAsync(&c->sc_host[2], c->sc, sizeof(FbScalars), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    FB_CUDA(cudaGetLastError());
    const FbScalars &s = c->sc_host[2];
    const double rhoFinal = s.rho[s.iters & 1];
    const bool notConverged = rhoFinal > s.eps2 * s.rho0;
    c->last_iters = s.iters * (notConverged ? -1 : 1);
    c->last_ratio = (s.rho0 != 0.0) ? rhoFinal / s.rho0 : 0.0;
    return FB_OK;
  }
  k_cg_init<<<vg, VEC_TB, 0, st>>>(n, c->rhs, c->invD, c->x, c->res, c->dir, c->sc, c->partials + FB_MAX_PARTIALS,
                                   c->dist ? &c->sc->rho_part : &c->sc->rho[0]);
  if (c->dist) FB_TRY(fb_dist_allreduce_scalar(c, &c->sc->rho_part, &c->sc->rho[0]));  // rho0 is a global sum
  k_cg_begin<<<1, 1, 0, st>>>(c->sc, eps, maxIt);
  c->launches += 2;
  if (c->dist) FB_TRY(fb_dist_halo_exchange(c, c->dir));  // ghost entries of d = invD r live on the neighbours
  // Iterations are enqueued in chunks; the loop condition lives on the device (kernels turn into
  // no-ops once `done` is set).  The host looks at the flag of chunk k-1 while chunk k runs.
  const int CH = 32;
  c->nprof = 0;
  int it = 1, slot = 0, pending = 0;
  bool finished = false;
  while (!finished && it <= maxIt) {
    const int end = (it + CH - 1 < maxIt) ? it + CH - 1 : maxIt;
    for (; it <= end; it++) enqueue_iteration(c, it);
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
  return FB_OK;
}

int fb_pcg_bench_iteration(fb_context *c, int repeats, double *sec) {
  // time `repeats` full CG iterations on the current system without the stopping rule
  const int n = c->r;
  if (n == 0 || repeats <= 0) { *sec = 0.0; return FB_OK; }
  cudaStream_t st = c->stream;
  const int vg = vec_grid(c, (size_t)n);
  if (c->pers_grid > 0 && !c->dist) {
    k_cg_init<<<vg, VEC_TB, 0, st>>>(n, c->rhs, c->invD, c->x, c->res, c->dir, c->sc, c->partials + FB_MAX_PARTIALS, &c->sc->rho[0]);
    k_cg_begin<<<1, 1, 0, st>>>(c->sc, 0.0, repeats);
    c->launches += 2;
    FB_CUDA(cudaEventRecord(c->ev[3], st));
    FB_TRY(fb_pcg_launch_persistent(c));
    FB_CUDA(cudaEventRecord(c->ev[7], st));
    FB_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    FB_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[7]));
    *sec = 1e-3 * ms / repeats;
    return FB_OK;
  }
  k_cg_init<<<vg, VEC_TB, 0, st>>>(n, c->rhs, c->invD, c->x, c->res, c->dir, c->sc, c->partials + FB_MAX_PARTIALS,
                                   c->dist ? &c->sc->rho_part : &c->sc->rho[0]);
  if (c->dist) FB_TRY(fb_dist_allreduce_scalar(c, &c->sc->rho_part, &c->sc->rho[0]));
  k_cg_begin<<<1, 1, 0, st>>>(c->sc, 0.0, 1 << 30);
  c->launches += 2;
  if (c->dist) FB_TRY(fb_dist_halo_exchange(c, c->dir));
  for (int it = 1; it <= 3; it++) enqueue_iteration(c, it);
  FB_CUDA(cudaEventRecord(c->ev[3], st));
  for (int it = 4; it < 4 + repeats; it++) enqueue_iteration(c, it);
  FB_CUDA(cudaEventRecord(c->ev[7], st));
  FB_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  FB_CUDA(cudaEventElapsedTime(&ms, c->ev[3], c->ev[7]));
  *sec = 1e-3 * ms / repeats;
  return FB_OK;
}
