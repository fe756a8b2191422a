// fb_mg.cu — LABELLED SOLVER VARIANTS: faster-converging preconditioners for the implicit-Euler solve.
//
// Not the reference's algorithm.  The reference solves Keff dv = rhs with Jacobi-preconditioned CG
// (CGSolver::SolveLinearSystemWithJacobiPreconditioner, src/3rdparty/vegafem/sparseSolver/CGSolver.cpp:129-190), which
// needs 750-1100 iterations per step on the 10M-tet benchmark mesh — >= 0.36 s per step even at 100 % of the HBM roofline.
// That path (fb_pcg.cu) stays the default and the parity path.  The variants below solve THE SAME linear system (the
// bit-identical Keff and rhs of the assembly) to THE SAME stopping rule — the Jacobi-weighted residual
// sum r_i^2 / diag_i <= eps^2 * sum b_i^2 / diag_i of CGSolver.cpp:147-150 — with a better preconditioner:
//
//   FB_SOLVER_BLOCK_JACOBI_PCG  z = B^-1 r with B the 3x3 diagonal blocks of Keff (any mesh)
//   FB_SOLVER_MG_PCG            z = one multigrid V(1,1) cycle (meshes whose vertices form a tensor grid, fb_set_grid):
//       levels     the grid coarsened 2:1 per axis (every other node plane, plus the last one) down to <= 3 nodes per axis;
//       operators  every coarse level is an ordinary context of the coarse TruthCube-split mesh (VolMeshSamples.cpp:76-116
//                  pattern on the coarse node coordinates) whose Keff is RE-ASSEMBLED every step by the same assembly kernels
//                  at the injected displacement (corotational rotations included) — no sparse triple products;
//       transfer   trilinear interpolation P (3x3 identity blocks) and restriction P^T, constrained DOFs masked on both sides;
//       smoother   damped 3x3-block Jacobi, omega = 1.4 / lambda_max(B^-1 A) (power iteration, refreshed every 32 solves);
//       coarsest   explicit dense inverse (<= 81 unknowns), rebuilt every step;
//       precision  the whole cycle runs in FP32 on an FP32 copy of each level's Keff (4.4 instead of 8.4 bytes per
//                  nonzero streamed); the outer CG — A d, the dot products, x, r, d — stays FP64 on the FP64 Keff.
//   The cycle is a fixed symmetric positive definite linear operator (same pre/post smoother, R = P^T), so plain PCG applies.
//   Iteration counts are mesh independent (CPU prototype on the reference's matrices: 34 at 16^3, 24^3 and 32^3 nodes
//   where Jacobi-PCG needs 558 / 691 / 728).
//   warm start (optional, both variants): x0 = the previous step's solution instead of 0.
//
// Everything is deterministic: sums are per-CTA slots added in index order, no floating-point atomics.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fb_internal.h"
#include "fb_pcg_common.cuh"

namespace {

constexpr int MG_TB = 256;
constexpr int MG_MAX_LEVELS = 12;
constexpr int MG_MAX_DENSE = 96;   // unknowns of the coarsest level's dense inverse
constexpr int MG_SLOTS = FB_MAX_PARTIALS;

struct MgLevel {
  fb_context *ctx;
  int n[3], nV, r;
  float *A32, *Binv, *b, *x, *xn, *res, *pv;
  double *u;                 // levels > 0: displacement injected from the finest level
  int nc[3];                 // node counts of the next coarser level (0 = this is the coarsest)
  int *cidx[3];              // device [nc]: fine index of every coarse node, per axis
  int *fa[3];                // device [n]: per fine index the coarse node at or below it ...
  float *fw[3];              // ... and the weight of the NEXT coarse node (0 when the fine node is a coarse node)
  int *twin;                 // device [nV of the coarser level]: fine vertex coinciding with each coarse vertex
  float lmax;
  int grid_spmv, grid_vec;
};

}  // namespace

struct FbMg {
  int variant, warm;
  int grid[3];               // tensor-grid dimensions given with fb_set_grid (0 = none)
  int nLevels;
  MgLevel L[MG_MAX_LEVELS];
  double *dense;             // [n*n] coarsest-level matrix, inverted in place every step
  float *denseInv;           // [n*n] fp32 copy of the inverse
  int nDense;
  double *slotsM, *slotsZ;   // per-CTA partial sums: weighted residual, r.z
  int solves;                // since the last lambda_max estimate
  int prepared;
  long long vcycles;
};

namespace {

// ---- FP32 copies of a level's operator --------------------------------------------------------------------------------
__global__ void k_mg_convert(size_t n, const double *__restrict__ a, float *__restrict__ o) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = (float)a[i];
}

// inverse of the 3x3 diagonal block of every vertex, constrained DOFs decoupled (their rows and columns of the inverse are 0)
__global__ void k_mg_block_inverse(int nV, const int *__restrict__ bp, const int *__restrict__ diag, const double *__restrict__ A,
                                   const unsigned char *__restrict__ mask, float *__restrict__ Binv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const int rs = bp[v], nb = bp[v + 1] - rs, jpos = diag[v] - rs;
  double m[3][3];
  bool fx[3];
  for (int k = 0; k < 3; k++) fx[k] = mask[3 * (size_t)v + k] != 0;
  for (int k = 0; k < 3; k++)
    for (int l = 0; l < 3; l++) {
      double val = A[9 * (size_t)rs + (size_t)k * 3 * nb + 3 * jpos + l];
      if (fx[k] || fx[l]) val = (k == l) ? 1.0 : 0.0;
      m[k][l] = val;
    }
  const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1], c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2], c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  const double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
  const double id = (det != 0.0) ? 1.0 / det : 0.0;
  double inv[3][3];
  inv[0][0] = c00 * id; inv[1][0] = c01 * id; inv[2][0] = c02 * id;
  inv[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  inv[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  inv[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  inv[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  inv[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  inv[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  for (int k = 0; k < 3; k++)
    for (int l = 0; l < 3; l++) Binv[9 * (size_t)v + 3 * k + l] = (fx[k] || fx[l]) ? 0.0f : (float)inv[k][l];
}

// ---- the FP32 product: 16 lanes per block row, three unrolled passes (rows of the tet stencil hold <= 16 blocks) --------
// MODE 0: out = mask(b - A x)
// MODE 1: out = x + omega * Binv (b - A x)      [DOT: also sum b . out into slots]
// MODE 2: out = Binv (A x)                       (power iteration)
template <int MODE, bool DOT>
__global__ void __launch_bounds__(MG_TB, 4) k_mg_spmv(int nV, const int *__restrict__ bp, const int *__restrict__ bc,
                                                     const float *__restrict__ A, const float *__restrict__ x,
                                                     const float *__restrict__ b, const float *__restrict__ Binv,
                                                     const unsigned char *__restrict__ mask, float omega, float *__restrict__ out,
                                                     double *slots, const FbScalars *sc) {
  pdl_wait();
  pdl_trigger();
  if (sc && sc->done) return;
  const int lane = threadIdx.x & 15;
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int group = blockIdx.x * (MG_TB / 16) + threadIdx.x / 16;
  const int nGroups = gridDim.x * (MG_TB / 16);
  double part = 0.0;
  // TWO block rows per trip: a 4-byte load moves half of what the FP64 product's loads move, so twice as many have to be in
  // flight per lane to keep the same number of bytes in flight per SM (one row per trip ran at 4.1 TB/s, profiles/r02_mg_*)
  for (int v0 = 2 * group; v0 < nV; v0 += 2 * nGroups) {
    int rs[2], n3[2];
    float acc[2][3];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int v = v0 + h;
      rs[h] = 0; n3[h] = 0;
      if (v < nV) { rs[h] = __ldg(bp + v); n3[h] = 3 * (__ldg(bp + v + 1) - rs[h]); }
      acc[h][0] = acc[h][1] = acc[h][2] = 0.f;
    }
    const int nmax = max(n3[0], n3[1]);
    for (int base = 0; base < nmax; base += 48) {
      float val[2][3][3];
      int col[2][3];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const float *a0 = A + 9 * (size_t)rs[h];
#pragma unroll
        for (int p = 0; p < 3; p++) {
          const int t = base + lane + 16 * p;
          const bool ok = t < n3[h];
          col[h][p] = ok ? __ldg(bc + rs[h] + t / 3) : -1;
          val[h][p][0] = ok ? __ldcs(a0 + t) : 0.f;
          val[h][p][1] = ok ? __ldcs(a0 + n3[h] + t) : 0.f;
          val[h][p][2] = ok ? __ldcs(a0 + 2 * (size_t)n3[h] + t) : 0.f;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; h++)
#pragma unroll
        for (int p = 0; p < 3; p++) {
          const int t = base + lane + 16 * p;
          const float xv = (col[h][p] >= 0) ? __ldg(x + 3 * (size_t)col[h][p] + (t % 3)) : 0.f;
          acc[h][0] = fmaf(val[h][p][0], xv, acc[h][0]);
          acc[h][1] = fmaf(val[h][p][1], xv, acc[h][1]);
          acc[h][2] = fmaf(val[h][p][2], xv, acc[h][2]);
        }
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
#pragma unroll
      for (int h = 0; h < 2; h++) {
        acc[h][0] += __shfl_xor_sync(gmask, acc[h][0], o, 16);
        acc[h][1] += __shfl_xor_sync(gmask, acc[h][1], o, 16);
        acc[h][2] += __shfl_xor_sync(gmask, acc[h][2], o, 16);
      }
    // lanes 0-2 finish row v0, lanes 8-10 row v0 + 1
    const int h = lane >> 3, k = lane & 7;
    const int v = v0 + h;
    if (k < 3 && v < nV) {
      const float a0v = h ? acc[1][0] : acc[0][0], a1v = h ? acc[1][1] : acc[0][1], a2v = h ? acc[1][2] : acc[0][2];
      const size_t row = 3 * (size_t)v + k;
      if (MODE == 0) {
        const float ax = k == 0 ? a0v : (k == 1 ? a1v : a2v);
        out[row] = mask[row] ? 0.f : (b[row] - ax);
      } else {
        float r0, r1, r2;
        if (MODE == 1) {
          r0 = b[3 * (size_t)v] - a0v; r1 = b[3 * (size_t)v + 1] - a1v; r2 = b[3 * (size_t)v + 2] - a2v;
        } else {
          r0 = a0v; r1 = a1v; r2 = a2v;
        }
        const float *bi = Binv + 9 * (size_t)v + 3 * k;
        const float corr = fmaf(bi[0], r0, fmaf(bi[1], r1, bi[2] * r2));   // zero rows / columns at constrained DOFs
        const float val = (MODE == 1) ? fmaf(omega, corr, x[row]) : corr;
        out[row] = val;
        if (DOT) part = fma((double)b[row], (double)val, part);
      }
    }
  }
  if (DOT) block_reduce_to_slot<MG_TB>(part, slots);
}

// x = omega Binv b  (a smoothing step from the zero initial guess)   [DOT: also sum b . x, for the one-level variant]
template <bool DOT>
__global__ void __launch_bounds__(MG_TB) k_mg_presmooth(int nV, const float *__restrict__ Binv, const float *__restrict__ b, float omega,
                                                        float *__restrict__ x, double *slots, const FbScalars *sc) {
  pdl_wait();
  pdl_trigger();
  if (sc && sc->done) return;
  double part = 0.0;
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < 3 * (size_t)nV; i += (size_t)gridDim.x * MG_TB) {
    const size_t v = i / 3;
    const int k = (int)(i - 3 * v);
    const float *bi = Binv + 9 * v + 3 * k;
    const float val = omega * fmaf(bi[0], b[3 * v], fmaf(bi[1], b[3 * v + 1], bi[2] * b[3 * v + 2]));
    x[i] = val;
    if (DOT) part = fma((double)b[i], (double)val, part);
  }
  if (DOT) block_reduce_to_slot<MG_TB>(part, slots);
}

struct GridMaps {
  int nf[3], nc[3];
  const int *cidx[3];
  const int *fa[3];
  const float *fw[3];
};

// b_c = P^T res_f : one thread per coarse DOF gathers the fine nodes between its neighbouring coarse nodes (fixed order)
__global__ void __launch_bounds__(MG_TB) k_mg_restrict(GridMaps g, const float *__restrict__ rf, const unsigned char *__restrict__ maskC,
                                                       float *__restrict__ bc, float *__restrict__ xc_zero, const FbScalars *sc) {
  if (sc && sc->done) return;
  const size_t nC = (size_t)g.nc[0] * g.nc[1] * g.nc[2];
  for (size_t t = (size_t)blockIdx.x * MG_TB + threadIdx.x; t < 3 * nC; t += (size_t)gridDim.x * MG_TB) {
    const size_t V = t / 3;
    const int k = (int)(t - 3 * V);
    const int K = (int)(V % g.nc[2]), J = (int)((V / g.nc[2]) % g.nc[1]), I = (int)(V / ((size_t)g.nc[2] * g.nc[1]));
    const int C[3] = {I, J, K};
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = C[d] > 0 ? g.cidx[d][C[d] - 1] + 1 : g.cidx[d][C[d]];
      hi[d] = C[d] + 1 < g.nc[d] ? g.cidx[d][C[d] + 1] - 1 : g.cidx[d][C[d]];
    }
    float s = 0.f;
    for (int i = lo[0]; i <= hi[0]; i++) {
      const float wi = (g.fa[0][i] == I) ? 1.f - g.fw[0][i] : g.fw[0][i];
      for (int j = lo[1]; j <= hi[1]; j++) {
        const float wj = (g.fa[1][j] == J) ? 1.f - g.fw[1][j] : g.fw[1][j];
        for (int kk = lo[2]; kk <= hi[2]; kk++) {
          const float wk = (g.fa[2][kk] == K) ? 1.f - g.fw[2][kk] : g.fw[2][kk];
          const size_t f = ((size_t)i * g.nf[1] + j) * g.nf[2] + kk;
          s = fmaf(wi * wj * wk, rf[3 * f + k], s);
        }
      }
    }
    bc[t] = maskC[t] ? 0.f : s;
    if (xc_zero) xc_zero[t] = 0.f;
  }
}

// x_f += P x_c : one thread per fine DOF, trilinear weights; constrained fine DOFs stay untouched (zero)
__global__ void __launch_bounds__(MG_TB) k_mg_prolong_add(GridMaps g, const float *__restrict__ xc, const unsigned char *__restrict__ maskF,
                                                          float *__restrict__ xf, const FbScalars *sc) {
  if (sc && sc->done) return;
  const size_t nF = (size_t)g.nf[0] * g.nf[1] * g.nf[2];
  for (size_t t = (size_t)blockIdx.x * MG_TB + threadIdx.x; t < 3 * nF; t += (size_t)gridDim.x * MG_TB) {
    if (maskF[t]) continue;
    const size_t v = t / 3;
    const int k = (int)(t - 3 * v);
    const int kk = (int)(v % g.nf[2]), j = (int)((v / g.nf[2]) % g.nf[1]), i = (int)(v / ((size_t)g.nf[2] * g.nf[1]));
    const int a[3] = {g.fa[0][i], g.fa[1][j], g.fa[2][kk]};
    const float w[3] = {g.fw[0][i], g.fw[1][j], g.fw[2][kk]};
    float s = 0.f;
    for (int di = 0; di < 2; di++) {
      const float wi = di ? w[0] : 1.f - w[0];
      if (wi == 0.f) continue;
      for (int dj = 0; dj < 2; dj++) {
        const float wj = dj ? w[1] : 1.f - w[1];
        if (wj == 0.f) continue;
        for (int dk = 0; dk < 2; dk++) {
          const float wk = dk ? w[2] : 1.f - w[2];
          if (wk == 0.f) continue;
          const size_t V = ((size_t)(a[0] + di) * g.nc[1] + (a[1] + dj)) * g.nc[2] + (a[2] + dk);
          s = fmaf(wi * wj * wk, xc[3 * V + k], s);
        }
      }
    }
    xf[t] += s;
  }
}

// displacement of the coarse level = the finest level's displacement at the coinciding vertices (twin maps compose)
__global__ void k_mg_inject(int nVc, const int *__restrict__ twin, const double *__restrict__ uf, double *__restrict__ uc) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * nVc) return;
  const int V = t / 3, k = t - 3 * V;
  uc[t] = uf[3 * (size_t)twin[V] + k];
}

// ---- coarsest level: dense matrix, explicit inverse by Gauss-Jordan (SPD: no pivoting), one CTA ------------------------
__global__ void k_mg_dense_build(int nV, const int *__restrict__ bp, const int *__restrict__ bc, const double *__restrict__ A,
                                 const unsigned char *__restrict__ mask, double *__restrict__ D) {
  const int n = 3 * nV;
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) D[i] = 0.0;
  __syncthreads();
  for (int row = threadIdx.x; row < n; row += blockDim.x) {
    const int v = row / 3, k = row - 3 * v;
    if (mask[row]) { D[(size_t)row * n + row] = 1.0; continue; }
    const int rs = bp[v], nb = bp[v + 1] - rs;
    for (int j = 0; j < nb; j++)
      for (int l = 0; l < 3; l++) {
        const int col = 3 * bc[rs + j] + l;
        if (!mask[col]) D[(size_t)row * n + col] = A[9 * (size_t)rs + (size_t)k * 3 * nb + 3 * j + l];
      }
  }
}
__global__ void __launch_bounds__(256) k_mg_dense_invert(int n, double *__restrict__ D, const unsigned char *__restrict__ mask,
                                                         float *__restrict__ inv) {
  // Gauss-Jordan in place in global memory (n <= 96: the matrix lives in L1/L2); column p eliminated per trip
  __shared__ double prow[MG_MAX_DENSE], pcol[MG_MAX_DENSE];
  __shared__ double pivInv;
  for (int p = 0; p < n; p++) {
    if (threadIdx.x == 0) pivInv = 1.0 / D[(size_t)p * n + p];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      prow[j] = D[(size_t)p * n + j] * pivInv;
      pcol[j] = D[(size_t)j * n + p];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
      const int i = t / n, j = t - i * n;
      double val;
      if (i == p) val = (j == p) ? pivInv : prow[j];
      else if (j == p) val = -pcol[i] * pivInv;
      else val = D[t] - pcol[i] * prow[j];
      D[t] = val;
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
    const int i = t / n, j = t - i * n;
    inv[t] = (mask[i] || mask[j]) ? 0.f : (float)D[t];
  }
}
__global__ void __launch_bounds__(128) k_mg_dense_apply(int n, const float *__restrict__ inv, const float *__restrict__ b,
                                                        float *__restrict__ x, const FbScalars *sc) {
  if (sc && sc->done) return;
  __shared__ float sb[MG_MAX_DENSE];
  for (int j = threadIdx.x; j < n; j += blockDim.x) sb[j] = b[j];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < n; j++) s = fmaf(inv[(size_t)i * n + j], sb[j], s);
    x[i] = s;
  }
}

// ---- reductions used off the hot path (power iteration) ------------------------------------------------------------------
__global__ void __launch_bounds__(MG_TB) k_mg_norm2(size_t n, const float *__restrict__ a, double *slots) {
  double part = 0.0;
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < n; i += (size_t)gridDim.x * MG_TB) part = fma((double)a[i], (double)a[i], part);
  block_reduce_to_slot<MG_TB>(part, slots);
}
__global__ void k_mg_scale_copy(size_t n, float s, const float *__restrict__ a, float *__restrict__ o) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = s * a[i];
}
__global__ void k_mg_fill_pattern(size_t n, const unsigned char *__restrict__ mask, float *__restrict__ o) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    o[i] = mask[i] ? 0.f : 1.0f + 0.37f * (float)((i * 2654435761ull) % 1000ull) * 1e-3f;
}

// ---- outer PCG, FP64 ----------------------------------------------------------------------------------------------------
// FbScalars use in this solver: rho[it & 1] = r.z after iteration it, rho0 = m0 = sum b^2 / diag, rq = m (latest weighted
// residual), dq slots from the FP64 product as in fb_pcg.cu.
// r = b - (warm ? A x0 : 0), x = x0, m0 and (no warm start) m partials; the V-cycle input is r in FP32
__global__ void __launch_bounds__(MG_TB) k_mgcg_init(int n, const double *__restrict__ b, const double *__restrict__ invD,
                                                     double *__restrict__ x, double *__restrict__ r, float *__restrict__ r32,
                                                     int warm, double *slotsM0, double *slotsM, FbScalars *sc) {
  // warm: r already holds b - A x0 (product kernel, mode 2) and x holds x0
  // the first preconditioner application runs before k_mgcg_begin: it must not see the previous solve's `done`
  if (blockIdx.x == 0 && threadIdx.x == 0) sc->done = 0;
  double p0 = 0.0, p1 = 0.0;
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * MG_TB) {
    const double bi = b[i], wi = invD[i];
    double ri;
    if (warm) ri = r[i];
    else { ri = bi; r[i] = bi; x[i] = 0.0; }
    r32[i] = (float)ri;
    p0 = fma(bi * bi, wi, p0);
    p1 = fma(ri * ri, wi, p1);
  }
  block_reduce_to_slot<MG_TB>(p0, slotsM0);
  __syncthreads();
  block_reduce_to_slot<MG_TB>(p1, slotsM);
}

// after the first preconditioner application: scalars of iteration 0, d = z
__global__ void __launch_bounds__(MG_TB) k_mgcg_begin(int n, const float *__restrict__ z, double *__restrict__ d, FbScalars *sc,
                                                      const double *slotsM0, const double *slotsM, const double *slotsZ, int nSlotsV,
                                                      int nSlotsZ, double eps, int maxIt) {
  const double m0 = cta_sum_slots<MG_TB>(slotsM0, nSlotsV);
  const double m = cta_sum_slots<MG_TB>(slotsM, nSlotsV);
  const double rz = cta_sum_slots<MG_TB>(slotsZ, nSlotsZ);
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * MG_TB) d[i] = (double)z[i];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    sc->rho0 = m0;
    sc->rq = m;
    sc->rho[0] = rz;
    sc->eps2 = eps * eps;
    sc->max_it = maxIt;
    sc->iters = 0;
    sc->comm_error = 0;
    // the reference's loop condition at iteration 1 (CGSolver.cpp:150) on the same weighted residual
    sc->done = !((m > eps * eps * m0) && (1 <= maxIt));
  }
}

// alpha = r.z / d.q; x += alpha d; r -= alpha q; r32 = r; m partial
__global__ void __launch_bounds__(MG_TB) k_mgcg_update(int n, const double *__restrict__ d, const double *__restrict__ q,
                                                       const double *__restrict__ invD, double *__restrict__ x, double *__restrict__ r,
                                                       float *__restrict__ r32, const FbScalars *sc, int it, const double *dqSlots,
                                                       int nDq, double *slotsM) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const double dq = cta_sum_slots<MG_TB>(dqSlots, nDq);
  const double alpha = sc->rho[(it - 1) & 1] / dq;
  double part = 0.0;
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * MG_TB) {
    x[i] = fma(alpha, d[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    r32[i] = (float)ri;
    part = fma(ri * ri, invD[i], part);
  }
  block_reduce_to_slot<MG_TB>(part, slotsM);
}

// one CTA: the stopping rule, decided BEFORE the next preconditioner application so that a converged solve does not pay
// for one more cycle
__global__ void __launch_bounds__(MG_TB) k_mgcg_check(FbScalars *sc, int it, const double *slotsM, int nSlots) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const double m = cta_sum_slots<MG_TB>(slotsM, nSlots);
  if (threadIdx.x == 0) {
    sc->rq = m;
    sc->iters = it;
    if (!((m > sc->eps2 * sc->rho0) && (it + 1 <= sc->max_it))) sc->done = 1;
  }
}

// beta = r.z' / r.z; d = z + beta d
__global__ void __launch_bounds__(MG_TB) k_mgcg_direction(int n, const float *__restrict__ z, double *__restrict__ d, FbScalars *sc, int it,
                                                          const double *slotsZ, int nSlotsZ) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const double rzNew = cta_sum_slots<MG_TB>(slotsZ, nSlotsZ);
  const double beta = rzNew / sc->rho[(it - 1) & 1];
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * MG_TB)
    d[i] = fma(beta, d[i], (double)z[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) sc->rho[it & 1] = rzNew;   // read next by k_mgcg_update(it + 1): other slot than the one read here
}

int grid_for_n(const fb_context *c, size_t n, int perThread = 1) {
  size_t want = (n + (size_t)MG_TB * perThread - 1) / ((size_t)MG_TB * perThread);
  size_t cap = (size_t)c->sm_count * 8;
  if (cap > MG_SLOTS) cap = MG_SLOTS;
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

GridMaps maps_of(const MgLevel &L) {
  GridMaps g;
  for (int d = 0; d < 3; d++) {
    g.nf[d] = L.n[d]; g.nc[d] = L.nc[d];
    g.cidx[d] = L.cidx[d]; g.fa[d] = L.fa[d]; g.fw[d] = L.fw[d];
  }
  return g;
}

template <typename T>
int mg_upload(fb_context *c, T **dev, const std::vector<T> &h) {
  FB_TRY(fb_dev_alloc(c, dev, h.size()));
  if (!h.empty()) FB_CUDA(cudaMemcpyAsync(*dev, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

// TruthCube split (VolMeshSamples.cpp:76-116: six tets per cell, listed corner order) on a tensor grid of node coordinates
void grid_mesh(const std::vector<double> ax[3], std::vector<double> &verts, std::vector<int> &tets) {
  const int nx = (int)ax[0].size(), ny = (int)ax[1].size(), nz = (int)ax[2].size();
  verts.resize(3 * (size_t)nx * ny * nz);
  for (int i = 0; i < nx; i++)
    for (int j = 0; j < ny; j++)
      for (int k = 0; k < nz; k++) {
        const size_t v = ((size_t)i * ny + j) * nz + k;
        verts[3 * v] = ax[0][i]; verts[3 * v + 1] = ax[1][j]; verts[3 * v + 2] = ax[2][k];
      }
  // corner ids LBN, LBF, LTN, LTF, RBN, RBF, RTN, RTF = (i, j, k) offsets with x = R, y = T, z = F
  static const int T6[6][4] = {{0, 2, 4, 1}, {6, 2, 1, 4}, {6, 2, 3, 1}, {6, 4, 1, 5}, {6, 1, 3, 5}, {6, 3, 7, 5}};
  tets.clear();
  tets.reserve(24 * (size_t)(nx - 1) * (ny - 1) * (nz - 1));
  for (int i = 0; i + 1 < nx; i++)
    for (int j = 0; j + 1 < ny; j++)
      for (int k = 0; k + 1 < nz; k++) {
        const int base = (i * ny + j) * nz + k;
        const int corner[8] = {base, base + 1, base + nz, base + nz + 1, base + ny * nz, base + ny * nz + 1, base + ny * nz + nz,
                               base + ny * nz + nz + 1};
        for (int t = 0; t < 6; t++)
          for (int a = 0; a < 4; a++) tets.push_back(corner[T6[t][a]]);
      }
}

void free_level(fb_context *owner, MgLevel &L, int li) {
  void *ptrs[] = {L.A32, L.Binv, L.b, L.x, L.xn, L.res, L.pv, L.u, L.twin, L.cidx[0], L.cidx[1], L.cidx[2], L.fa[0], L.fa[1], L.fa[2],
                  L.fw[0], L.fw[1], L.fw[2]};
  for (void *p : ptrs)
    if (p) fb_dev_free(p);
  if (li > 0 && L.ctx) fb_destroy(L.ctx);
  (void)owner;
  memset(&L, 0, sizeof(L));
}

int alloc_level_vectors(fb_context *c, MgLevel &L) {
  fb_context *lc = L.ctx;
  L.nV = lc->nV; L.r = lc->r;
  FB_TRY(fb_dev_alloc(c, &L.A32, (size_t)lc->nnzK + 4));
  FB_TRY(fb_dev_alloc(c, &L.Binv, 9 * (size_t)lc->nV));
  float **vecs[] = {&L.b, &L.x, &L.xn, &L.res, &L.pv};
  for (float **v : vecs) {
    FB_TRY(fb_dev_alloc(c, v, (size_t)lc->r + 4));
    FB_CUDA(cudaMemsetAsync(*v, 0, sizeof(float) * ((size_t)lc->r + 4), c->stream));
  }
  int perSM = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_mg_spmv<1, true>, MG_TB, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
  size_t want = ((size_t)lc->nV + 2 * (MG_TB / 16) - 1) / (2 * (MG_TB / 16));   // two block rows per 16-lane group and trip
  size_t cap = (size_t)c->sm_count * perSM;
  if (cap > MG_SLOTS) cap = MG_SLOTS;
  L.grid_spmv = (int)std::max<size_t>(1, std::min(want, cap));
  L.grid_vec = grid_for_n(c, (size_t)lc->r);
  L.lmax = 0.f;
  return FB_OK;
}

// lambda_max(Binv A) of one level by power iteration (host reads the norms: only at setup and every 32 solves)
int estimate_lmax(fb_context *c, FbMg *mg, MgLevel &L, int its) {
  cudaStream_t st = c->stream;
  fb_context *lc = L.ctx;
  if (L.r == 0) { L.lmax = 1.f; return FB_OK; }
  const bool cold = L.lmax == 0.f;
  if (cold) k_mg_fill_pattern<<<L.grid_vec, MG_TB, 0, st>>>((size_t)L.r, lc->rowmask, L.pv);
  double lam = L.lmax;
  std::vector<double> h((size_t)L.grid_vec);
  for (int k = 0; k < its; k++) {
    k_mg_spmv<2, false><<<L.grid_spmv, MG_TB, 0, st>>>(L.nV, lc->bp, lc->bc, L.A32, L.pv, nullptr, L.Binv, lc->rowmask, 0.f, L.res, nullptr, nullptr);
    k_mg_norm2<<<L.grid_vec, MG_TB, 0, st>>>((size_t)L.r, L.res, mg->slotsM);
    k_mg_norm2<<<L.grid_vec, MG_TB, 0, st>>>((size_t)L.r, L.pv, mg->slotsZ);
    double ny = 0.0, nx = 0.0;
    FB_CUDA(cudaMemcpyAsync(h.data(), mg->slotsM, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    for (double v : h) ny += v;
    FB_CUDA(cudaMemcpyAsync(h.data(), mg->slotsZ, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    for (double v : h) nx += v;
    c->launches += 3;
    if (!(nx > 0.0) || !(ny > 0.0) || !std::isfinite(ny)) break;
    lam = std::sqrt(ny / nx);
    k_mg_scale_copy<<<L.grid_vec, MG_TB, 0, st>>>((size_t)L.r, (float)(1.0 / std::sqrt(ny)), L.res, L.pv);
    c->launches++;
  }
  L.lmax = (float)((lam > 0.0 && std::isfinite(lam)) ? lam : 3.0);
  return FB_OK;
}

// z = V(b) on level li, result in L.xn (li = 0: with the dot b.z into slotsZ)
void vcycle(fb_context *c, FbMg *mg, int li) {
  cudaStream_t st = c->stream;
  MgLevel &L = mg->L[li];
  fb_context *lc = L.ctx;
  const FbScalars *sc = c->sc;
  if (li == mg->nLevels - 1) {   // coarsest: dense inverse
    k_mg_dense_apply<<<1, 128, 0, st>>>(L.r, mg->denseInv, L.b, L.xn, sc);
    c->launches++;
    return;
  }
  MgLevel &C = mg->L[li + 1];
  const float omega = 1.4f / L.lmax;
  const GridMaps g = maps_of(L);
  fb_launch(true, st, k_mg_presmooth<false>, L.grid_vec, MG_TB, L.nV, L.Binv, L.b, omega, L.x, (double *)nullptr, sc);
  fb_launch(true, st, k_mg_spmv<0, false>, L.grid_spmv, MG_TB, L.nV, lc->bp, lc->bc, L.A32, L.x, L.b, L.Binv, lc->rowmask, omega, L.res,
            (double *)nullptr, sc);
  k_mg_restrict<<<C.grid_vec, MG_TB, 0, st>>>(g, L.res, C.ctx->rowmask, C.b, (float *)nullptr, sc);
  c->launches += 3;
  vcycle(c, mg, li + 1);
  k_mg_prolong_add<<<L.grid_vec, MG_TB, 0, st>>>(g, C.xn, lc->rowmask, L.x, sc);
  if (li == 0)
    fb_launch(true, st, k_mg_spmv<1, true>, L.grid_spmv, MG_TB, L.nV, lc->bp, lc->bc, L.A32, L.x, L.b, L.Binv, lc->rowmask, omega, L.xn,
              mg->slotsZ, sc);
  else
    fb_launch(true, st, k_mg_spmv<1, false>, L.grid_spmv, MG_TB, L.nV, lc->bp, lc->bc, L.A32, L.x, L.b, L.Binv, lc->rowmask, omega, L.xn,
              (double *)nullptr, sc);
  c->launches += 2;
}

// z = preconditioner(r32) on the finest level: result in L[0].xn, r.z partials in slotsZ (nSlotsZ returned)
int apply_preconditioner(fb_context *c, FbMg *mg) {
  MgLevel &L = mg->L[0];
  if (mg->variant == FB_SOLVER_MG_PCG && mg->nLevels > 1) {
    vcycle(c, mg, 0);
    mg->vcycles++;
    return L.grid_spmv;
  }
  fb_launch(true, c->stream, k_mg_presmooth<true>, L.grid_vec, MG_TB, L.nV, L.Binv, L.b, 1.0f, L.xn, mg->slotsZ, (const FbScalars *)c->sc);
  c->launches++;
  return L.grid_vec;
}

}  // namespace

// ---- setup -----------------------------------------------------------------------------------------------------------
void fb_mg_release(fb_context *c) {
  FbMg *mg = c->mg;
  if (!mg) return;
  for (int li = mg->nLevels - 1; li >= 0; li--) free_level(c, mg->L[li], li);
  if (mg->dense) fb_dev_free(mg->dense);
  if (mg->denseInv) fb_dev_free(mg->denseInv);
  if (mg->slotsM) fb_dev_free(mg->slotsM);
  if (mg->slotsZ) fb_dev_free(mg->slotsZ);
  delete mg;
  c->mg = nullptr;
}

static int mg_ensure(fb_context *c) {
  if (c->mg) return FB_OK;
  FbMg *mg = new FbMg();
  memset(mg, 0, sizeof(*mg));
  c->mg = mg;
  FB_TRY(fb_dev_alloc(c, &mg->slotsM, 2 * (size_t)MG_SLOTS));
  FB_TRY(fb_dev_alloc(c, &mg->slotsZ, (size_t)MG_SLOTS));
  return FB_OK;
}

// builds (or rebuilds) the level hierarchy for the current variant; level 0 always exists for variants != 0
static int mg_build(fb_context *c) {
  FbMg *mg = c->mg;
  for (int li = mg->nLevels - 1; li >= 0; li--) free_level(c, mg->L[li], li);
  mg->nLevels = 0;
  mg->prepared = 0;
  if (mg->dense) { fb_dev_free(mg->dense); mg->dense = nullptr; }
  if (mg->denseInv) { fb_dev_free(mg->denseInv); mg->denseInv = nullptr; }
  if (mg->variant == FB_SOLVER_JACOBI_PCG) return FB_OK;
  MgLevel &L0 = mg->L[0];
  memset(&L0, 0, sizeof(L0));
  L0.ctx = c;
  FB_TRY(alloc_level_vectors(c, L0));
  mg->nLevels = 1;
  if (mg->variant != FB_SOLVER_MG_PCG) return FB_OK;
  const int *n0 = mg->grid;
  // axis coordinates from the rest positions (vertex (i, j, k) has index (i ny + j) nz + k)
  std::vector<double> x0(3 * (size_t)c->nV);
  FB_CUDA(cudaMemcpyAsync(x0.data(), c->x0, sizeof(double) * x0.size(), cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  std::vector<double> ax[3];
  ax[0].resize(n0[0]); ax[1].resize(n0[1]); ax[2].resize(n0[2]);
  for (int i = 0; i < n0[0]; i++) ax[0][i] = x0[3 * ((size_t)i * n0[1] * n0[2])];
  for (int j = 0; j < n0[1]; j++) ax[1][j] = x0[3 * ((size_t)j * n0[2]) + 1];
  for (int k = 0; k < n0[2]; k++) ax[2][k] = x0[3 * (size_t)k + 2];
  for (int i = 0; i < n0[0]; i += std::max(1, n0[0] / 7))       // spot check: the vertices really are that tensor grid
    for (int j = 0; j < n0[1]; j += std::max(1, n0[1] / 7))
      for (int k = 0; k < n0[2]; k += std::max(1, n0[2] / 7)) {
        const size_t v = ((size_t)i * n0[1] + j) * n0[2] + k;
        if (x0[3 * v] != ax[0][i] || x0[3 * v + 1] != ax[1][j] || x0[3 * v + 2] != ax[2][k]) {
          fb_set_error("fb_set_grid: vertex %zu is not node (%d,%d,%d) of a tensor grid in i-major numbering", v, i, j, k);
          return FB_ERR_INVALID_ARGUMENT;
        }
      }
  for (int d = 0; d < 3; d++) L0.n[d] = n0[d];
  std::vector<int> twinFine((size_t)c->nV);   // finest-level vertex of every vertex of the current level
  for (int v = 0; v < c->nV; v++) twinFine[v] = v;
  std::vector<unsigned char> fixedFine((size_t)c->r, 0);
  for (int i = 0; i < c->nC; i++) fixedFine[c->cdofs_host[i]] = 1;
  int li = 0;
  while (true) {
    MgLevel &L = mg->L[li];
    if (L.n[0] <= 3 && L.n[1] <= 3 && L.n[2] <= 3) break;
    if (li + 1 >= MG_MAX_LEVELS) break;
    // 2:1 coarsening per axis: every other node plus the last one; axes with <= 3 nodes are kept
    std::vector<int> cidx[3], fa[3];
    std::vector<float> fw[3];
    std::vector<double> cax[3];
    for (int d = 0; d < 3; d++) {
      const int n = L.n[d];
      if (n <= 3) { for (int i = 0; i < n; i++) cidx[d].push_back(i); }
      else {
        for (int i = 0; i < n; i += 2) cidx[d].push_back(i);
        if (cidx[d].back() != n - 1) cidx[d].push_back(n - 1);
      }
      // per fine node i: a = the coarse node at or below it, w = weight of coarse node a + 1 (0 when i IS coarse node a;
      // the kernels never touch a + 1 when w == 0, so the last node needs no special case)
      fa[d].resize(n); fw[d].resize(n);
      int a = 0;
      for (int i = 0; i < n; i++) {
        while (a + 1 < (int)cidx[d].size() && cidx[d][a + 1] <= i) a++;
        fa[d][i] = a;
        fw[d][i] = (cidx[d][a] == i) ? 0.f : (float)((ax[d][i] - ax[d][cidx[d][a]]) / (ax[d][cidx[d][a + 1]] - ax[d][cidx[d][a]]));
      }
      for (size_t q = 0; q < cidx[d].size(); q++) cax[d].push_back(ax[d][cidx[d][q]]);
      L.nc[d] = (int)cidx[d].size();
    }
    for (int d = 0; d < 3; d++) {
      FB_TRY(mg_upload(c, &L.cidx[d], cidx[d]));
      FB_TRY(mg_upload(c, &L.fa[d], fa[d]));
      FB_TRY(mg_upload(c, &L.fw[d], fw[d]));
    }
    // the coarse level: mesh, constrained DOFs (a coarse DOF is constrained iff its finest-level twin is), context
    std::vector<double> cv;
    std::vector<int> ct;
    grid_mesh(cax, cv, ct);
    const int nVc = L.nc[0] * L.nc[1] * L.nc[2];
    std::vector<int> twin((size_t)nVc), twinF((size_t)nVc), cd;
    for (int I = 0; I < L.nc[0]; I++)
      for (int J = 0; J < L.nc[1]; J++)
        for (int K = 0; K < L.nc[2]; K++) {
          const int V = (I * L.nc[1] + J) * L.nc[2] + K;
          const int f = (cidx[0][I] * L.n[1] + cidx[1][J]) * L.n[2] + cidx[2][K];
          twin[V] = f;
          twinF[V] = twinFine[f];
        }
    for (int V = 0; V < nVc; V++)
      for (int k = 0; k < 3; k++)
        if (fixedFine[3 * (size_t)twinF[V] + k]) cd.push_back(3 * V + k);
    MgLevel &C = mg->L[li + 1];
    memset(&C, 0, sizeof(C));
    fb_context *lc = nullptr;
    fb_params prm = c->prm;
    prm.solver_variant = 0;
    FB_TRY(fb_create_local(&lc, nVc, cv.data(), (int)(ct.size() / 4), ct.data(), (int)cd.size(), cd.data(), nullptr, nullptr, nullptr, &prm));
    // the level works on the owner's stream so that its kernels are ordered with the solve
    cudaStreamSynchronize(lc->stream);
    cudaStreamDestroy(lc->stream);
    lc->stream = c->stream;
    lc->stream_borrowed = 1;
    C.ctx = lc;
    for (int d = 0; d < 3; d++) C.n[d] = L.nc[d];
    mg->nLevels = li + 2;
    FB_TRY(alloc_level_vectors(c, C));
    FB_TRY(fb_dev_alloc(c, &C.u, (size_t)lc->r));
    FB_TRY(mg_upload(c, &C.twin, twinF));   // straight from the FINEST level: injection needs no intermediate vectors
    c->bytes += lc->bytes;
    twinFine.swap(twinF);
    for (int d = 0; d < 3; d++) ax[d].swap(cax[d]);
    li++;
  }
  MgLevel &Lc = mg->L[mg->nLevels - 1];
  if (mg->nLevels > 1) {
    if (Lc.r > MG_MAX_DENSE) { fb_set_error("multigrid: coarsest level has %d unknowns (> %d)", Lc.r, MG_MAX_DENSE); return FB_ERR_NOT_SUPPORTED; }
    mg->nDense = Lc.r;
    FB_TRY(fb_dev_alloc(c, &mg->dense, (size_t)Lc.r * Lc.r + 1));
    FB_TRY(fb_dev_alloc(c, &mg->denseInv, (size_t)Lc.r * Lc.r + 1));
  }
  return FB_OK;
}

// Per step, after the assembly of the finest level: coarse operators at the injected displacement, FP32 copies, block
// inverses, coarsest inverse, lambda_max (first solve, then every 32).
int fb_mg_prepare(fb_context *c) {
  FbMg *mg = c->mg;
  if (!mg || mg->variant == FB_SOLVER_JACOBI_PCG) return FB_OK;
  cudaStream_t st = c->stream;
  if (mg->nLevels == 0) FB_TRY(mg_build(c));
  for (int li = 0; li < mg->nLevels; li++) {
    MgLevel &L = mg->L[li];
    fb_context *lc = L.ctx;
    if (li > 0) {
      lc->prm.timestep = c->prm.timestep;
      lc->prm.damping_mass = c->prm.damping_mass;
      lc->prm.damping_stiffness = c->prm.damping_stiffness;
      lc->prm.internal_force_scaling = c->prm.internal_force_scaling;
      k_mg_inject<<<(3 * L.nV + 255) / 256, 256, 0, st>>>(L.nV, L.twin, c->q, L.u);
      c->launches++;
      const long long before = lc->launches;
      FB_TRY(fb_launch_assembly(lc, L.u, nullptr, true));
      c->launches += lc->launches - before;
    }
    k_mg_convert<<<grid_for_n(c, (size_t)lc->nnzK, 4), MG_TB, 0, st>>>((size_t)lc->nnzK, lc->Keff, L.A32);
    k_mg_block_inverse<<<(L.nV + 127) / 128, 128, 0, st>>>(L.nV, lc->bp, lc->diag, lc->Keff, lc->rowmask, L.Binv);
    c->launches += 2;
  }
  if (mg->nLevels > 1) {
    MgLevel &Lc = mg->L[mg->nLevels - 1];
    k_mg_dense_build<<<1, 256, 0, st>>>(Lc.nV, Lc.ctx->bp, Lc.ctx->bc, Lc.ctx->Keff, Lc.ctx->rowmask, mg->dense);
    k_mg_dense_invert<<<1, 256, 0, st>>>(Lc.r, mg->dense, Lc.ctx->rowmask, mg->denseInv);
    c->launches += 2;
    const bool first = mg->L[0].lmax == 0.f;
    if (first || mg->solves >= 32) {
      for (int li = 0; li + 1 < mg->nLevels; li++) FB_TRY(estimate_lmax(c, mg, mg->L[li], first ? 12 : 3));
      mg->solves = 0;
    }
  }
  FB_CUDA(cudaGetLastError());
  mg->prepared = 1;
  return FB_OK;
}

void fb_mg_invalidate(fb_context *c) {
  FbMg *mg = c->mg;
  if (!mg) return;
  for (int li = mg->nLevels - 1; li >= 0; li--) free_level(c, mg->L[li], li);
  mg->nLevels = 0;
  mg->prepared = 0;
}

int fb_mg_active(const fb_context *c) { return (c->mg && c->mg->variant != FB_SOLVER_JACOBI_PCG) ? c->mg->variant : 0; }
int fb_mg_warm(const fb_context *c) { return c->mg ? c->mg->warm : 0; }

// PCG with the variant's preconditioner on Keff x = rhs (masked).  Same contract as fb_pcg_solve.
int fb_mg_pcg_solve(fb_context *c, double eps, int maxIt) {
  FbMg *mg = c->mg;
  cudaStream_t st = c->stream;
  if (c->r == 0) { c->last_iters = 0; c->last_ratio = 0.0; return FB_OK; }
  if (!mg->prepared) FB_TRY(fb_mg_prepare(c));   // fb_solve before any step: operators of the current Keff
  MgLevel &L = mg->L[0];
  const int n = c->r, gv = L.grid_vec;
  double *slotsM0 = mg->slotsM + MG_SLOTS, *slotsM = mg->slotsM;
  const int warm = mg->warm && c->have_solution;
  c->nprof = 0;
  if (warm) {   // r = mask(b - A x0); the product kernel is a no-op while the previous solve's `done` is still set
    FB_CUDA(cudaMemsetAsync(&c->sc->done, 0, sizeof(int), st));
    FB_TRY(fb_pcg_launch_residual(c, c->x, c->res));
  }
  k_mgcg_init<<<gv, MG_TB, 0, st>>>(n, c->rhs, c->invD, c->x, c->res, L.b, warm, slotsM0, slotsM, c->sc);
  c->launches++;
  int nZ = apply_preconditioner(c, mg);
  k_mgcg_begin<<<gv, MG_TB, 0, st>>>(n, L.xn, c->dir, c->sc, slotsM0, slotsM, mg->slotsZ, gv, nZ, eps, maxIt);
  c->launches++;
  const int CH = 6;
  static const bool trace = getenv("FEMBRAIN_B200_MG_TRACE") && atoi(getenv("FEMBRAIN_B200_MG_TRACE")) != 0;
  int it = 1, slot = 0, pending = 0;
  bool finished = false;
  while (!finished && it <= maxIt) {
    const int end = (it + CH - 1 < maxIt) ? it + CH - 1 : maxIt;
    for (; it <= end; it++) {
      const bool sample = c->profiling && (it % 4 == 1) && c->nprof < 64;
      if (sample) cudaEventRecord(c->evProf[2 * c->nprof], st);
      int nDq = 0;
      FB_TRY(fb_pcg_launch_product_dq(c, c->dir, c->Ad, &nDq));    // q = A d, d.q partials in c->partials
      if (sample) { cudaEventRecord(c->evProf[2 * c->nprof + 1], st); c->nprof++; }
      fb_launch(true, st, k_mgcg_update, gv, MG_TB, n, c->dir, c->Ad, c->invD, c->x, c->res, L.b, (const FbScalars *)c->sc, it,
                (const double *)c->partials, nDq, slotsM);
      fb_launch(true, st, k_mgcg_check, 1, MG_TB, c->sc, it, (const double *)slotsM, gv);
      c->launches += 2;
      nZ = apply_preconditioner(c, mg);
      fb_launch(true, st, k_mgcg_direction, gv, MG_TB, n, (const float *)L.xn, c->dir, c->sc, it, (const double *)mg->slotsZ, nZ);
      c->launches++;
    }
    FB_CUDA(cudaMemcpyAsync(&c->sc_host[slot], c->sc, sizeof(FbScalars), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaEventRecord(c->evChunk[slot], st));
    pending++;
    if (pending == 2) {
      const int prev = slot ^ 1;
      FB_CUDA(cudaEventSynchronize(c->evChunk[prev]));
      if (c->sc_host[prev].done) finished = true;
      if (trace) fprintf(stderr, "[mg] it %d  m/m0 %.3e  r.z %.3e\n", c->sc_host[prev].iters, c->sc_host[prev].rq / c->sc_host[prev].rho0,
                         c->sc_host[prev].rho[c->sc_host[prev].iters & 1]);
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
  const bool notConverged = s.rq > s.eps2 * s.rho0;
  c->last_iters = s.iters * (notConverged ? -1 : 1);
  c->last_ratio = (s.rho0 != 0.0) ? s.rq / s.rho0 : 0.0;
  mg->solves++;
  c->have_solution = !notConverged;
  return FB_OK;
}

// =====================================================================================================================
extern "C" {

int fb_set_grid(fb_context *c, int nx, int ny, int nz) {
  if (!c) { fb_set_error("NULL context"); return FB_ERR_INVALID_ARGUMENT; }
  if (cudaSetDevice(c->device) != cudaSuccess) { cudaGetLastError(); return FB_ERR_CUDA; }
  if (nx < 2 || ny < 2 || nz < 2 || (long long)nx * ny * nz != (long long)c->nV) {
    fb_set_error("fb_set_grid: %d x %d x %d nodes do not match the mesh's %d vertices", nx, ny, nz, c->nV);
    return FB_ERR_INVALID_ARGUMENT;
  }
  if (c->dist || c->batch) { fb_set_error("fb_set_grid: partitioned and batch contexts use the reference's solver"); return FB_ERR_NOT_SUPPORTED; }
  FB_TRY(mg_ensure(c));
  c->mg->grid[0] = nx; c->mg->grid[1] = ny; c->mg->grid[2] = nz;
  if (c->mg->variant == FB_SOLVER_MG_PCG) return mg_build(c);
  return FB_OK;
}

int fb_set_solver(fb_context *c, int variant, int warm_start) {
  if (!c) { fb_set_error("NULL context"); return FB_ERR_INVALID_ARGUMENT; }
  if (cudaSetDevice(c->device) != cudaSuccess) { cudaGetLastError(); return FB_ERR_CUDA; }
  if (variant < FB_SOLVER_JACOBI_PCG || variant > FB_SOLVER_MG_PCG) { fb_set_error("unknown solver variant %d", variant); return FB_ERR_INVALID_ARGUMENT; }
  if (variant != FB_SOLVER_JACOBI_PCG && (c->dist || c->batch)) {
    fb_set_error("solver variants are single-mesh, single-GPU: partitioned and batch contexts use the reference's solver");
    return FB_ERR_NOT_SUPPORTED;
  }
  FB_TRY(mg_ensure(c));
  if (variant == FB_SOLVER_MG_PCG) {
    if (c->mg->grid[0] == 0) { fb_set_error("FB_SOLVER_MG_PCG needs the tensor grid of the mesh: call fb_set_grid first"); return FB_ERR_INVALID_ARGUMENT; }
    if (!c->uniform_material) { fb_set_error("FB_SOLVER_MG_PCG: per-element materials are not carried to the coarse levels"); return FB_ERR_NOT_SUPPORTED; }
  }
  c->mg->warm = warm_start != 0;
  if (variant == c->mg->variant && (variant == FB_SOLVER_JACOBI_PCG || c->mg->nLevels > 0)) return FB_OK;
  c->mg->variant = variant;
  c->prm.solver_variant = variant;
  return mg_build(c);
}

int fb_get_solver(const fb_context *c, int *variant, int *warm_start, int *levels) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  if (variant) *variant = c->mg ? c->mg->variant : FB_SOLVER_JACOBI_PCG;
  if (warm_start) *warm_start = c->mg ? c->mg->warm : 0;
  if (levels) *levels = c->mg ? c->mg->nLevels : 0;
  return FB_OK;
}

const char *fb_solver_name(int variant) {
  switch (variant) {
    case FB_SOLVER_JACOBI_PCG: return "jacobi_pcg (the reference's algorithm, CGSolver.cpp:129-190)";
    case FB_SOLVER_BLOCK_JACOBI_PCG: return "block_jacobi_pcg (variant: 3x3 block-diagonal preconditioner, FP32 apply)";
    case FB_SOLVER_MG_PCG: return "mg_pcg (variant: geometric multigrid V(1,1) preconditioner on re-assembled coarse levels, FP32 cycle, FP64 CG)";
    default: return "unknown";
  }
}

}  // extern "C"
