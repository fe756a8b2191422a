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
//   FB_SOLVER_MG_PCG            z = one multigrid V(3,3) cycle (meshes whose vertices form a tensor grid, fb_set_grid):
//       levels     the grid coarsened 2:1 per axis (every other node plane, plus the last one) down to <= 3 nodes per axis;
//       operators  every coarse level is an ordinary context of the coarse TruthCube-split mesh (VolMeshSamples.cpp:76-116
//                  pattern on the coarse node coordinates) whose Keff is RE-ASSEMBLED every step by the same assembly kernels
//                  at the injected displacement (corotational rotations included) — no sparse triple products;
//       transfer   trilinear interpolation P (3x3 identity blocks) and restriction P^T, constrained DOFs masked on both sides;
//       smoother   three sweeps of a Chebyshev iteration preconditioned by the 3x3 block diagonal B, on the interval
//                  [1.1 lambda_max / 20, 1.1 lambda_max] of B^-1 A (lambda_max by power iteration: 30 iterations at build, 3 more
//                  every 128 solves; damped block Jacobi with FEMBRAIN_B200_MG_SMOOTHER=jacobi);
//       coarsest   explicit dense inverse (<= 81 unknowns), rebuilt every step in shared memory;
//       precision  the whole cycle runs in FP32 arithmetic on a reduced-precision copy of each level's Keff: FP16 values
//                  (default; scaled by a power of two per level so that the largest diagonal entry sits near 2^13) or FP32
//                  (FEMBRAIN_B200_MG_PREC=fp32), 3x3 blocks padded to 3x4 so that one lane loads one block with three
//                  8-byte (16-byte) loads and one float4 of x — 28 (52) instead of 76 bytes per block streamed; tensor-grid
//                  levels of >= 200,000 vertices store the FP16 blocks slot-major instead (k_mg_spmv_ell: one thread per row,
//                  24 coalesced bytes per block, x staged in shared memory: 0.87 of the HBM copy peak);
//                  the outer CG — A d, the dot products, x, r, d — stays FP64 on the FP64 Keff.
//   The cycle is a fixed symmetric positive definite linear operator (same pre/post smoother, R = P^T), so plain PCG applies.
//   Iteration counts are mesh independent (CPU prototype on the reference's matrices: 34 at 16^3, 24^3 and 32^3 nodes
//   where Jacobi-PCG needs 558 / 691 / 728).
//   warm start (optional, both variants): x0 = the previous step's solution instead of 0.
//
// Everything is deterministic: sums are per-CTA slots added in index order, no floating-point atomics.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_fp16.h>

#include "fb_internal.h"
#include "fb_pcg_common.cuh"

namespace {

// developer aid (FEMBRAIN_B200_MG_TIMING=1): CUDA events between the phases of ONE iteration per solve, printed to stderr
struct MgTiming {
  bool on, armed;
  cudaEvent_t ev[16];
  int n;
  const char *name[16];
};
MgTiming g_tm = {false, false, {}, 0, {}};
void tm_mark(cudaStream_t st, const char *what) {
  if (!g_tm.armed || g_tm.n >= 16) return;
  if (!g_tm.ev[g_tm.n]) cudaEventCreate(&g_tm.ev[g_tm.n]);
  cudaEventRecord(g_tm.ev[g_tm.n], st);
  g_tm.name[g_tm.n++] = what;
}

constexpr int MG_TB = 256;
constexpr int VS = 4;   // floats per vertex in the FP32 level vectors: (x, y, z, 0) so that a lane fetches a vertex with one float4 load
constexpr int MG_MAX_LEVELS = 12;
constexpr int MG_MAX_DENSE = 96;   // unknowns of the coarsest level's dense inverse
constexpr int MG_SLOTS = FB_MAX_PARTIALS;

struct MgLevel {
  fb_context *ctx;
  int n[3], nV, r;
  void *AB;                  // [nB][3][4] blocks of Keff, __half or float (padded rows), scaled by scale[0]
  float *scale;              // device [2]: power-of-two scale of AB and its inverse; [2] = bits of the running min of 1/diag
  float *Binv, *b, *x, *xn, *res, *pv;
  double *u;                 // levels > 0: displacement injected from the finest level
  int nc[3];                 // node counts of the next coarser level (0 = this is the coarsest)
  int *cidx[3];              // device [nc]: fine index of every coarse node, per axis
  int *fa[3];                // device [n]: per fine index the coarse node at or below it ...
  float *fw[3];              // ... and the weight of the NEXT coarse node (0 when the fine node is a coarse node)
  int *twin;                 // device [nV of the coarser level]: fine vertex coinciding with each coarse vertex
  float lmax;
  int grid_spmv, grid_vec;
  // structured storage of the level's matrix (tensor-grid levels: every row's blocks sit at a subset of ONE set of <= 16 grid
  // offsets): AE[slot][vertex][3][4] __half, empty slots zero; see k_mg_spmv_ell
  __half *AE;
  int nSlots, grid_ell, grid_pack;
  int slotOff[16];           // vertex-index offset of every slot
  signed char slotOf[27];    // (di+1)*9 + (dj+1)*3 + (dk+1) -> slot, -1 = not in the stencil
};

}  // namespace

struct FbMg {
  int variant, warm;
  int grid[3];               // tensor-grid dimensions given with fb_set_grid (0 = none)
  int nLevels;
  MgLevel L[MG_MAX_LEVELS];
  float *denseInv;           // [n*n] fp32 copy of the inverse
  int nDense;
  double *slotsM, *slotsZ;   // per-CTA partial sums: weighted residual, r.z
  int nu;                    // smoothing sweeps before and after the coarse correction
  int cheb;                  // nu >= 2: the sweeps are a Chebyshev iteration on [hi lambda_max / alpha, hi lambda_max] instead of nu damped steps
  float chebAlpha, chebHi;
  int nuCoarse;              // sweeps on levels >= 1 (0: the same as nu): they cost launch latency, not bandwidth
  int fallbacks;             // solves repeated after a lambda_max re-estimate (see fb_mg_pcg_solve)
  float coarseScale;         // the prolongated correction is added times this factor (1: plain; the cycle stays symmetric for any > 0)
  int useEll;                // structured slot-major storage + k_mg_spmv_ell on tensor-grid levels (FP16 storage only)
  int ellMinV;               // ... on levels with at least this many vertices (a row per thread needs that many rows to fill the GPU)
  signed char *slotOfDev;    // device copies of every level's slotOf table, 27 bytes per level
  int useGraph, capturing, subFailed, subKernels;   // levels >= 1 replayed as one CUDA graph
  void *subGraph;            // cudaGraphExec_t
  float *subResult;
  int preDone;               // the finest level's first smoothing sweep was fused into k_mgcg_update for this cycle
  int half;                  // 1: FP16 storage of the levels' matrices (default), 0: FP32
  int solves;                // since the last lambda_max estimate
  int prepared;
  long long vcycles;
};

namespace {

// ---- reduced-precision copy of a level's operator ----------------------------------------------------------------------
// scale[0] = 2^e with max diag * 2^e in [2^13, 2^14) (FP16 storage; 1 for FP32), scale[1] = 2^-e; the minimum of 1/diag over
// the unconstrained DOFs is taken with an integer atomicMin on the bits of the positive floats (order independent)
__global__ void __launch_bounds__(MG_TB) k_mg_scale_min(size_t n, const double *__restrict__ invD, unsigned int *bits) {
  float m = __int_as_float(0x7f800000);
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < n; i += (size_t)gridDim.x * MG_TB) {
    const float w = (float)invD[i];
    if (w > 0.f) m = fminf(m, w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMin(bits, __float_as_uint(m));
}
__global__ void k_mg_scale_final(float *scale, int half) {
  unsigned int *bits = reinterpret_cast<unsigned int *>(scale + 2);
  const float minInv = __uint_as_float(*bits);
  float s = 1.f;
  if (half && minInv > 0.f && minInv < __int_as_float(0x7f800000)) {
    int e;
    frexpf(1.f / minInv, &e);        // max diag = f * 2^e, f in [0.5, 1)
    s = ldexpf(1.f, 14 - e);         // max diag * s in [2^13, 2^14)
  }
  scale[0] = s;
  scale[1] = 1.f / s;
  *bits = 0x7f800000u;
}

template <typename T> __device__ __forceinline__ T mg_store(float v);
template <> __device__ __forceinline__ float mg_store<float>(float v) { return v; }
template <> __device__ __forceinline__ __half mg_store<__half>(float v) { return __float2half_rn(v); }

// Row v with nb blocks owns 12 nb entries at 12 bp[v]: THREE PLANES of nb x 4 entries, plane k = row k of every block
// (entry (j, k, l) at 12 bp[v] + k * 4 nb + 4 j + l, l = 3 is padding).  The 16 lanes of a row then read 16 x 8 (16) contiguous
// bytes per load instruction — one or two L1 wavefronts instead of the four of a block-major layout.  One thread per block.
template <typename T>
__global__ void __launch_bounds__(MG_TB) k_mg_pack(int nB, const int *__restrict__ bp, const int *__restrict__ brow, const double *__restrict__ A,
                                                   const float *__restrict__ scale, T *__restrict__ AB) {
  const float s = scale[0];
  for (size_t b = (size_t)blockIdx.x * MG_TB + threadIdx.x; b < (size_t)nB; b += (size_t)gridDim.x * MG_TB) {
    const int v = brow[b], rs = bp[v], nb = bp[v + 1] - rs, j = (int)b - rs;
    const double *a = A + 9 * (size_t)rs + 3 * j;
    T *o = AB + 12 * (size_t)rs + 4 * j;
#pragma unroll
    for (int k = 0; k < 3; k++) {
#pragma unroll
      for (int l = 0; l < 3; l++) o[(size_t)k * 4 * nb + l] = mg_store<T>((float)(a[(size_t)k * 3 * nb + l] * (double)s));
      o[(size_t)k * 4 * nb + 3] = mg_store<T>(0.f);
    }
  }
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

// one 3x3 block = one 8-byte (__half) or 16-byte (float) word from each of the three planes of its block row
template <typename T> struct MgRaw;
template <> struct MgRaw<__half> { typedef uint2 type; };
template <> struct MgRaw<float> { typedef float4 type; };
__device__ __forceinline__ void mg_unpack(const uint2 &w, float &a0, float &a1, float &a2) {
  const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&w.x));
  const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&w.y));
  a0 = lo.x; a1 = lo.y; a2 = hi.x;
}
__device__ __forceinline__ void mg_unpack(const float4 &w, float &a0, float &a1, float &a2) { a0 = w.x; a1 = w.y; a2 = w.z; }

// ---- the product of the cycle: 16 lanes per block row, ONE LANE PER 3x3 BLOCK, two block rows per trip, software pipelined --
// A lane fetches its block (three 8/16-byte words), the block's column and that vertex of x (one float4).  A trip's loads form
// a chain of three dependent round trips (row pointer -> column -> x); un-pipelined, that chain — not DRAM, not the L1
// wavefront rate — set the pace: 270-300 us per product at 10M tets for 0.73 GB (FP16), as slow as the 2 GB FP64 product
// (profiles/r02_mg_iteration_phases.txt).  Here the row pointers are fetched two trips ahead and columns + blocks one trip
// ahead, so a trip waits for one round trip (the x gather) only.
// MODE 0: out = mask(b - A x)
// MODE 1: out = x + omega * Binv (b - A x)      [DOT: also sum b . out into slots]
// MODE 2: out = Binv (A x)                       (power iteration)
template <typename T, int MODE, bool DOT>
__global__ void __launch_bounds__(MG_TB, 4) k_mg_spmv(int nV, const int *__restrict__ bp, const int *__restrict__ bc,
                                                     const T *__restrict__ AB, const float *__restrict__ scale,
                                                     const float *__restrict__ x, const float *__restrict__ b,
                                                     const float *__restrict__ Binv, const unsigned char *__restrict__ mask, float omega,
                                                     float gamma, int usePrev, float *__restrict__ out, double *slots, const FbScalars *sc) {
  typedef typename MgRaw<T>::type Raw;
  pdl_wait();
  pdl_trigger();
  if (sc && sc->done) return;
  const float inv = __ldg(scale + 1);
  const int lane = threadIdx.x & 15;
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int group = blockIdx.x * (MG_TB / 16) + threadIdx.x / 16;
  const int stride = 2 * gridDim.x * (MG_TB / 16);
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  const Raw *AR = reinterpret_cast<const Raw *>(AB);
  double part = 0.0;
  // pipeline registers: [0] = this trip, [1] = next trip, per row h of the pair
  int rs[2][2], nb[2][2], col[2];
  Raw raw[2][3];
  auto load_ptrs = [&](int v0, int slot) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int v = v0 + h;
      rs[slot][h] = 0; nb[slot][h] = 0;
      if (v < nV) { rs[slot][h] = __ldg(bp + v); nb[slot][h] = __ldg(bp + v + 1) - rs[slot][h]; }
    }
  };
  auto load_blocks = [&](int slot) {   // first 16 blocks of both rows of the pair in `slot`
#pragma unroll
    for (int h = 0; h < 2; h++) {
      col[h] = -1;
      if (lane < nb[slot][h]) {
        col[h] = __ldg(bc + rs[slot][h] + lane);
        const Raw *p = AR + 3 * (size_t)rs[slot][h] + lane;
#pragma unroll
        for (int k = 0; k < 3; k++) raw[h][k] = __ldcs(p + (size_t)k * nb[slot][h]);
      }
    }
  };
  int v0 = 2 * group;
  load_ptrs(v0, 0);
  load_ptrs(v0 + stride, 1);
  load_blocks(0);
  for (; v0 < nV; v0 += stride) {
    // this trip's operands are in flight or have arrived; start the gather of x, which depends on col only
    float4 xv[2];
    float a[2][3][3];
    int nbc[2], rsc[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      rsc[h] = rs[0][h]; nbc[h] = nb[0][h];
      xv[h] = (col[h] >= 0) ? __ldg(x4 + col[h]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        a[h][k][0] = a[h][k][1] = a[h][k][2] = 0.f;
        if (col[h] >= 0) mg_unpack(raw[h][k], a[h][k][0], a[h][k][1], a[h][k][2]);
      }
    }
    // the finishing lanes' own operands (independent of everything above)
    const int hh = lane >> 3, kk = lane & 7;
    const int vf = v0 + hh;
    const bool fin = kk < 3 && vf < nV;
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    float bi0 = 0.f, bi1 = 0.f, bi2 = 0.f, xr = 0.f;
    unsigned char mk = 0;
    if (fin) {
      if (MODE != 2) bv = *reinterpret_cast<const float4 *>(b + VS * (size_t)vf);
      if (MODE == 0) mk = mask[3 * (size_t)vf + kk];
      if (MODE != 0) {
        const float *bi = Binv + 9 * (size_t)vf + 3 * kk;
        bi0 = bi[0]; bi1 = bi[1]; bi2 = bi[2];
      }
      if (MODE == 1) xr = x[VS * (size_t)vf + kk];
    }
    // next trip: columns + blocks (its row pointers arrived a trip ago); the trip after: row pointers
    rs[0][0] = rs[1][0]; rs[0][1] = rs[1][1]; nb[0][0] = nb[1][0]; nb[0][1] = nb[1][1];
    load_blocks(0);
    load_ptrs(v0 + 2 * stride, 1);
    float acc[2][3];
#pragma unroll
    for (int h = 0; h < 2; h++)
#pragma unroll
      for (int k = 0; k < 3; k++) acc[h][k] = fmaf(a[h][k][0], xv[h].x, fmaf(a[h][k][1], xv[h].y, a[h][k][2] * xv[h].z));
    // rows with more than 16 blocks (never on the tet stencil of a tensor grid): the rest, un-pipelined
#pragma unroll
    for (int h = 0; h < 2; h++)
      for (int j = 16 + lane; j < nbc[h]; j += 16) {
        const int c2 = __ldg(bc + rsc[h] + j);
        const float4 x2 = __ldg(x4 + c2);
        const Raw *p = AR + 3 * (size_t)rsc[h] + j;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          float a0, a1, a2;
          mg_unpack(__ldcs(p + (size_t)k * nbc[h]), a0, a1, a2);
          acc[h][k] = fmaf(a0, x2.x, fmaf(a1, x2.y, fmaf(a2, x2.z, acc[h][k])));
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
    if (fin) {
      const float a0v = inv * (hh ? acc[1][0] : acc[0][0]), a1v = inv * (hh ? acc[1][1] : acc[0][1]), a2v = inv * (hh ? acc[1][2] : acc[0][2]);
      const size_t row = VS * (size_t)vf + kk;
      if (MODE == 0) {
        const float ax = kk == 0 ? a0v : (kk == 1 ? a1v : a2v);
        const float bk = kk == 0 ? bv.x : (kk == 1 ? bv.y : bv.z);
        out[row] = mk ? 0.f : (bk - ax);
      } else {
        const float r0 = (MODE == 1) ? bv.x - a0v : a0v, r1 = (MODE == 1) ? bv.y - a1v : a1v, r2 = (MODE == 1) ? bv.z - a2v : a2v;
        const float corr = fmaf(bi0, r0, fmaf(bi1, r1, bi2 * r2));   // zero rows / columns at constrained DOFs
        float val = (MODE == 1) ? fmaf(omega, corr, xr) : corr;
        if (MODE == 1 && gamma != 0.f) val = fmaf(gamma, xr - (usePrev ? out[row] : 0.f), val);   // Chebyshev momentum: out holds x_(k-1)
        out[row] = val;
        if (DOT) {
          const float bk = kk == 0 ? bv.x : (kk == 1 ? bv.y : bv.z);
          part = fma((double)bk, (double)val, part);
        }
      }
    }
  }
  if (DOT) block_reduce_to_slot<MG_TB>(part, slots);
}

// ---- structured product for tensor-grid levels -----------------------------------------------------------------------
// On a tensor grid the tet stencil is the same set of grid offsets (di, dj, dk) for every vertex (15 of the 27 for the
// TruthCube split; boundary vertices use a subset).  Storing the blocks slot-major, AE[slot][vertex], turns the product into
// a streaming kernel: ONE THREAD PER ROW, matrix loads coalesced across the threads of a warp (24 contiguous bytes per
// thread and slot), x staged through shared memory as nine runs of consecutive vertices (one per (di, dj)), no shuffles.
// That removes the scattered float4 gathers whose L1 wavefronts bounded k_mg_spmv (77 % of the LSU data pipe at 4.5 TB/s,
// profiles/r02_mg_spmv16_ncu_details.txt).
// Tried and dropped: HALF storage (diagonal + blocks towards higher vertices; the block towards v - off read as the transpose
// of AE[slot][v - off], a coalesced second read meant to come out of L2): correct, but 187 us per product instead of 142 at
// 1.77M vertices — the same 15 x 24 B per row still cross L2 -> SM, plus the transposed FMAs, at lower occupancy.
struct EllArgs {
  int nV, nSlots, nz, nynz;
  int slotOff[16];        // di * ny*nz + dj * nz + dk
  signed char slotSeg[16];   // (di+1)*3 + (dj+1): which of the nine staged runs
  signed char slotDk[16];
};

constexpr int ELL_T = 128;   // rows (threads) per tile

// FP64 block rows -> slot-major FP16.  A tile of 32 consecutive vertices' rows is one contiguous range of Keff: read coalesced
// into shared memory, then one lane per (block position j, vertex) so that a warp writes 32 consecutive 24-byte entries of one
// slot.  (One thread per block: 24-byte writes nV entries apart; one thread per vertex: 24-byte reads a row apart — 1.25 ms at
// 1.77M vertices for 2.5 GB of traffic, profiles/r02_mg_prepare_phases.txt.)
constexpr int PK_V = 32, PK_T = 128;
constexpr size_t PK_SMEM = sizeof(double) * PK_V * 16 * 9 + sizeof(int) * (PK_V * 16 + PK_V + 1);
__global__ void __launch_bounds__(PK_T) k_mg_pack_ell(int nV, int ny, int nz, const int *__restrict__ bp, const int *__restrict__ bc,
                                                      const double *__restrict__ A, const float *__restrict__ scale,
                                                      const signed char *__restrict__ slotOf27, __half *__restrict__ AE) {
  extern __shared__ double pk_sm[];
  double *sa = pk_sm;
  int *scol = reinterpret_cast<int *>(pk_sm + PK_V * 16 * 9);
  int *sbp = scol + PK_V * 16;
  const double s = (double)scale[0];
  const int nTiles = (nV + PK_V - 1) / PK_V;
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int v0 = tile * PK_V, nVt = min(PK_V, nV - v0);
    if (threadIdx.x <= nVt) sbp[threadIdx.x] = bp[v0 + threadIdx.x];
    __syncthreads();
    const int b0 = sbp[0], nBt = sbp[nVt] - b0;
    if (nBt <= PK_V * 16) {
      const double *src = A + 9 * (size_t)b0;
      for (int i = threadIdx.x; i < 9 * nBt; i += PK_T) sa[i] = __ldcs(src + i);
      for (int i = threadIdx.x; i < nBt; i += PK_T) scol[i] = bc[b0 + i];
    }
    __syncthreads();
    for (int item = threadIdx.x; item < 16 * PK_V; item += PK_T) {
      const int j = item / PK_V, vl = item - j * PK_V;
      if (vl >= nVt || nBt > PK_V * 16) continue;
      const int rs = sbp[vl] - b0, nb = sbp[vl + 1] - sbp[vl];
      if (j >= nb) continue;
      const int v = v0 + vl, c = scol[rs + j];
      const int vk = v % nz, vj = (v / nz) % ny, vi = v / (nz * ny);
      const int ck = c % nz, cj = (c / nz) % ny, ci = c / (nz * ny);
      const int slot = slotOf27[(ci - vi + 1) * 9 + (cj - vj + 1) * 3 + (ck - vk + 1)];
      const double *a = sa + 9 * rs + 3 * j;
      uint2 *o = reinterpret_cast<uint2 *>(AE) + 3 * ((size_t)slot * nV + v);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const __half2 lo = __floats2half2_rn((float)(a[k * 3 * nb] * s), (float)(a[k * 3 * nb + 1] * s));
        const __half2 hi = __floats2half2_rn((float)(a[k * 3 * nb + 2] * s), 0.f);
        uint2 w;
        w.x = *reinterpret_cast<const unsigned int *>(&lo);
        w.y = *reinterpret_cast<const unsigned int *>(&hi);
        o[k] = w;
      }
    }
    __syncthreads();
  }
}

// MODE 0: out = mask(b - A x);  MODE 1: out = x + omega Binv (b - A x) [DOT: sum b . out];  MODE 2: out = Binv (A x)
template <int MODE, bool DOT>
__global__ void __launch_bounds__(ELL_T, 8) k_mg_spmv_ell(EllArgs g, const __half *__restrict__ AE, const float *__restrict__ scale,
                                                         const float *__restrict__ x, const float *__restrict__ b,
                                                         const float *__restrict__ Binv, const unsigned char *__restrict__ mask, float omega,
                                                         float gamma, int usePrev, float *__restrict__ out, double *slots, const FbScalars *sc) {
  pdl_wait();
  pdl_trigger();
  if (sc && sc->done) return;
  __shared__ float4 xs[9][ELL_T + 2];
  const float inv = __ldg(scale + 1);
  const float4 *x4 = reinterpret_cast<const float4 *>(x);
  const uint2 *AR = reinterpret_cast<const uint2 *>(AE);
  double part = 0.0;
  const int nTiles = (g.nV + ELL_T - 1) / ELL_T;
  for (int tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const int v0 = tile * ELL_T;
    // the nine runs of x this tile's rows can touch: vertices v0 + (di ny nz + dj nz) - 1 ... + ELL_T (clamped: out-of-range
    // and wrapped positions only ever meet structurally zero blocks)
    for (int e = threadIdx.x; e < 9 * (ELL_T + 2); e += ELL_T) {
      const int seg = e / (ELL_T + 2), t = e - seg * (ELL_T + 2);
      long long idx = (long long)v0 + (long long)(seg / 3 - 1) * g.nynz + (long long)(seg % 3 - 1) * g.nz - 1 + t;
      idx = idx < 0 ? 0 : (idx >= g.nV ? g.nV - 1 : idx);
      xs[seg][t] = __ldg(x4 + idx);
    }
    __syncthreads();
    const int v = v0 + threadIdx.x;
    if (v < g.nV) {
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
      for (int s0 = 0; s0 < g.nSlots; s0 += 5) {
        uint2 w[5][3];
#pragma unroll
        for (int q = 0; q < 5; q++) {
          const int sl = s0 + q;
          if (sl < g.nSlots) {
            const uint2 *p = AR + 3 * ((size_t)sl * g.nV + v);
            w[q][0] = __ldcs(p); w[q][1] = __ldcs(p + 1); w[q][2] = __ldcs(p + 2);
          }
        }
#pragma unroll
        for (int q = 0; q < 5; q++) {
          const int sl = s0 + q;
          if (sl < g.nSlots) {
            const float4 xv = xs[g.slotSeg[sl]][threadIdx.x + 1 + g.slotDk[sl]];
            float a0, a1, a2;
            mg_unpack(w[q][0], a0, a1, a2); acc0 = fmaf(a0, xv.x, fmaf(a1, xv.y, fmaf(a2, xv.z, acc0)));
            mg_unpack(w[q][1], a0, a1, a2); acc1 = fmaf(a0, xv.x, fmaf(a1, xv.y, fmaf(a2, xv.z, acc1)));
            mg_unpack(w[q][2], a0, a1, a2); acc2 = fmaf(a0, xv.x, fmaf(a1, xv.y, fmaf(a2, xv.z, acc2)));
          }
        }
      }
      acc0 *= inv; acc1 *= inv; acc2 *= inv;
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (MODE == 0) {
        const float4 bv = *reinterpret_cast<const float4 *>(b + VS * (size_t)v);
        o.x = mask[3 * (size_t)v] ? 0.f : bv.x - acc0;
        o.y = mask[3 * (size_t)v + 1] ? 0.f : bv.y - acc1;
        o.z = mask[3 * (size_t)v + 2] ? 0.f : bv.z - acc2;
      } else {
        float r0 = acc0, r1 = acc1, r2 = acc2;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == 1) {
          bv = *reinterpret_cast<const float4 *>(b + VS * (size_t)v);
          r0 = bv.x - acc0; r1 = bv.y - acc1; r2 = bv.z - acc2;
        }
        const float *bi = Binv + 9 * (size_t)v;
        o.x = fmaf(bi[0], r0, fmaf(bi[1], r1, bi[2] * r2));
        o.y = fmaf(bi[3], r0, fmaf(bi[4], r1, bi[5] * r2));
        o.z = fmaf(bi[6], r0, fmaf(bi[7], r1, bi[8] * r2));
        if (MODE == 1) {
          const float4 xc = xs[4][threadIdx.x + 1];   // the row's own vertex: run (di, dj) = (0, 0), dk = 0
          o.x = fmaf(omega, o.x, xc.x); o.y = fmaf(omega, o.y, xc.y); o.z = fmaf(omega, o.z, xc.z);
          if (gamma != 0.f) {   // Chebyshev momentum: out holds x_(k-1)
            float4 pv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (usePrev) pv = reinterpret_cast<const float4 *>(out)[v];
            o.x = fmaf(gamma, xc.x - pv.x, o.x); o.y = fmaf(gamma, xc.y - pv.y, o.y); o.z = fmaf(gamma, xc.z - pv.z, o.z);
          }
          if (DOT) part = fma((double)bv.x, (double)o.x, fma((double)bv.y, (double)o.y, fma((double)bv.z, (double)o.z, part)));
        }
      }
      reinterpret_cast<float4 *>(out)[v] = o;
    }
    __syncthreads();
  }
  if (DOT) block_reduce_to_slot<ELL_T>(part, slots);
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
    const float val = omega * fmaf(bi[0], b[VS * v], fmaf(bi[1], b[VS * v + 1], bi[2] * b[VS * v + 2]));
    x[VS * v + k] = val;
    if (DOT) part = fma((double)b[VS * v + k], (double)val, part);
  }
  if (DOT) block_reduce_to_slot<MG_TB>(part, slots);
}

struct GridMaps {
  int nf[3], nc[3];
  const int *cidx[3];
  const int *fa[3];
  const float *fw[3];
};

// b_c = P^T res_f : one thread per coarse VERTEX gathers the fine vertices between its neighbouring coarse nodes (fixed order),
// one float4 load per fine vertex
__global__ void __launch_bounds__(MG_TB) k_mg_restrict(GridMaps g, const float *__restrict__ rf, const unsigned char *__restrict__ maskC,
                                                       float *__restrict__ bc, float *__restrict__ xc_zero, const FbScalars *sc) {
  if (sc && sc->done) return;
  (void)xc_zero;
  const size_t nC = (size_t)g.nc[0] * g.nc[1] * g.nc[2];
  const float4 *rf4 = reinterpret_cast<const float4 *>(rf);
  for (size_t V = (size_t)blockIdx.x * MG_TB + threadIdx.x; V < nC; V += (size_t)gridDim.x * MG_TB) {
    const int K = (int)(V % g.nc[2]), J = (int)((V / g.nc[2]) % g.nc[1]), I = (int)(V / ((size_t)g.nc[2] * g.nc[1]));
    const int C[3] = {I, J, K};
    int lo[3], hi[3];
    for (int d = 0; d < 3; d++) {
      lo[d] = C[d] > 0 ? g.cidx[d][C[d] - 1] + 1 : g.cidx[d][C[d]];
      hi[d] = C[d] + 1 < g.nc[d] ? g.cidx[d][C[d] + 1] - 1 : g.cidx[d][C[d]];
    }
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = lo[0]; i <= hi[0]; i++) {
      const float wi = (g.fa[0][i] == I) ? 1.f - g.fw[0][i] : g.fw[0][i];
      for (int j = lo[1]; j <= hi[1]; j++) {
        const float wj = (g.fa[1][j] == J) ? 1.f - g.fw[1][j] : g.fw[1][j];
        for (int kk = lo[2]; kk <= hi[2]; kk++) {
          const float wk = (g.fa[2][kk] == K) ? 1.f - g.fw[2][kk] : g.fw[2][kk];
          const size_t f = ((size_t)i * g.nf[1] + j) * g.nf[2] + kk;
          const float w = wi * wj * wk;
          const float4 r = __ldg(rf4 + f);
          sx = fmaf(w, r.x, sx); sy = fmaf(w, r.y, sy); sz = fmaf(w, r.z, sz);
        }
      }
    }
    float4 o;
    o.x = maskC[3 * V] ? 0.f : sx; o.y = maskC[3 * V + 1] ? 0.f : sy; o.z = maskC[3 * V + 2] ? 0.f : sz; o.w = 0.f;
    reinterpret_cast<float4 *>(bc)[V] = o;
  }
}

// x_f += P x_c : one thread per fine VERTEX, trilinear weights, float4 loads; constrained fine DOFs stay untouched (zero)
__global__ void __launch_bounds__(MG_TB) k_mg_prolong_add(GridMaps g, const float *__restrict__ xc, const unsigned char *__restrict__ maskF,
                                                          float *__restrict__ xf, float scale, const FbScalars *sc) {
  if (sc && sc->done) return;
  const size_t nF = (size_t)g.nf[0] * g.nf[1] * g.nf[2];
  const float4 *xc4 = reinterpret_cast<const float4 *>(xc);
  float4 *xf4 = reinterpret_cast<float4 *>(xf);
  for (size_t v = (size_t)blockIdx.x * MG_TB + threadIdx.x; v < nF; v += (size_t)gridDim.x * MG_TB) {
    const int kk = (int)(v % g.nf[2]), j = (int)((v / g.nf[2]) % g.nf[1]), i = (int)(v / ((size_t)g.nf[2] * g.nf[1]));
    const int a[3] = {g.fa[0][i], g.fa[1][j], g.fa[2][kk]};
    const float w[3] = {g.fw[0][i], g.fw[1][j], g.fw[2][kk]};
    float sx = 0.f, sy = 0.f, sz = 0.f;
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
          const float ww = wi * wj * wk;
          const float4 c = __ldg(xc4 + V);
          sx = fmaf(ww, c.x, sx); sy = fmaf(ww, c.y, sy); sz = fmaf(ww, c.z, sz);
        }
      }
    }
    float4 x = xf4[v];
    if (!maskF[3 * v]) x.x = fmaf(scale, sx, x.x);
    if (!maskF[3 * v + 1]) x.y = fmaf(scale, sy, x.y);
    if (!maskF[3 * v + 2]) x.z = fmaf(scale, sz, x.z);
    xf4[v] = x;
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
// coarsest level: the constrained matrix as a dense array in SHARED memory (n <= 96: 72 KB of doubles), Gauss-Jordan in place —
// column p eliminated per trip — and the inverse written out in FP32.  One CTA of 1024 threads; in global memory with 256
// threads the same elimination took 320 us per step (profiles/r02_mg_prepare_phases.txt).
constexpr int MG_DENSE_T = 1024;
__global__ void __launch_bounds__(MG_DENSE_T) k_mg_dense_inverse(int nV, const int *__restrict__ bp, const int *__restrict__ bc,
                                                                 const double *__restrict__ A, const unsigned char *__restrict__ mask,
                                                                 float *__restrict__ inv) {
  extern __shared__ double D[];
  __shared__ double prow[MG_MAX_DENSE], pcol[MG_MAX_DENSE];
  __shared__ double pivInv;
  const int n = 3 * nV;
  for (int i = threadIdx.x; i < n * n; i += MG_DENSE_T) D[i] = 0.0;
  __syncthreads();
  for (int item = threadIdx.x; item < n * 16; item += MG_DENSE_T) {   // (row, block of the row)
    const int row = item >> 4, j = item & 15;
    const int v = row / 3, k = row - 3 * v;
    if (mask[row]) { if (j == 0) D[row * n + row] = 1.0; continue; }
    const int rs = bp[v], nb = bp[v + 1] - rs;
    for (int jj = j; jj < nb; jj += 16)
      for (int l = 0; l < 3; l++) {
        const int col = 3 * bc[rs + jj] + l;
        if (!mask[col]) D[row * n + col] = A[9 * (size_t)rs + (size_t)k * 3 * nb + 3 * jj + l];
      }
  }
  __syncthreads();
  for (int p = 0; p < n; p++) {
    if (threadIdx.x == 0) pivInv = 1.0 / D[p * n + p];
    __syncthreads();
    for (int j = threadIdx.x; j < n; j += MG_DENSE_T) {
      prow[j] = D[p * n + j] * pivInv;
      pcol[j] = D[j * n + p];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < n * n; t += MG_DENSE_T) {
      const int i = t / n, j = t - i * n;
      double val;
      if (i == p) val = (j == p) ? pivInv : prow[j];
      else if (j == p) val = -pcol[i] * pivInv;
      else val = D[t] - pcol[i] * prow[j];
      D[t] = val;
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < n * n; t += MG_DENSE_T) {
    const int i = t / n, j = t - i * n;
    inv[t] = (mask[i] || mask[j]) ? 0.f : (float)D[t];
  }
}
__global__ void __launch_bounds__(128) k_mg_dense_apply(int n, const float *__restrict__ inv, const float *__restrict__ b,
                                                        float *__restrict__ x, const FbScalars *sc) {
  if (sc && sc->done) return;
  __shared__ float sb[MG_MAX_DENSE];
  for (int j = threadIdx.x; j < n; j += blockDim.x) sb[j] = b[j + j / 3];   // DOF 3v + k lives at VS v + k
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < n; j++) s = fmaf(inv[(size_t)i * n + j], sb[j], s);
    x[i + i / 3] = s;
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
__global__ void k_mg_fill_pattern(size_t n, const unsigned char *__restrict__ mask, float *__restrict__ o) {   // n = 3 nV DOFs
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    o[i + i / 3] = mask[i] ? 0.f : 1.0f + 0.37f * (float)((i * 2654435761ull) % 1000ull) * 1e-3f;
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
    r32[i + i / 3] = (float)ri;
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
  for (size_t i = (size_t)blockIdx.x * MG_TB + threadIdx.x; i < (size_t)n; i += (size_t)gridDim.x * MG_TB) d[i] = (double)z[i + i / 3];
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

// alpha = r.z / d.q; x += alpha d; r -= alpha q; m partial; r32 = r and — when Binv is given — the cycle's first smoothing
// sweep from the zero guess, x1 = omega Binv r32, in the same pass (one thread per vertex: it holds all three components)
__global__ void __launch_bounds__(MG_TB) k_mgcg_update(int nV, const double *__restrict__ d, const double *__restrict__ q,
                                                       const double *__restrict__ invD, double *__restrict__ x, double *__restrict__ r,
                                                       float *__restrict__ r32, const FbScalars *sc, int it, const double *dqSlots,
                                                       int nDq, double *slotsM, const float *__restrict__ Binv, float omega,
                                                       float *__restrict__ x1) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const double dq = cta_sum_slots<MG_TB>(dqSlots, nDq);
  const double alpha = sc->rho[(it - 1) & 1] / dq;
  double part = 0.0;
  for (size_t v = (size_t)blockIdx.x * MG_TB + threadIdx.x; v < (size_t)nV; v += (size_t)gridDim.x * MG_TB) {
    float rv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const size_t i = 3 * v + k;
      x[i] = fma(alpha, d[i], x[i]);
      const double ri = fma(-alpha, q[i], r[i]);
      r[i] = ri;
      rv[k] = (float)ri;
      part = fma(ri * ri, invD[i], part);
    }
    reinterpret_cast<float4 *>(r32)[v] = make_float4(rv[0], rv[1], rv[2], 0.f);
    if (Binv) {
      const float *bi = Binv + 9 * v;
      float4 o;
      o.x = omega * fmaf(bi[0], rv[0], fmaf(bi[1], rv[1], bi[2] * rv[2]));
      o.y = omega * fmaf(bi[3], rv[0], fmaf(bi[4], rv[1], bi[5] * rv[2]));
      o.z = omega * fmaf(bi[6], rv[0], fmaf(bi[7], rv[1], bi[8] * rv[2]));
      o.w = 0.f;
      reinterpret_cast<float4 *>(x1)[v] = o;
    }
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
    d[i] = fma(beta, d[i], (double)z[i + i / 3]);
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
  void *ptrs[] = {L.AE, L.AB, L.scale, L.Binv, L.b, L.x, L.xn, L.res, L.pv, L.u, L.twin, L.cidx[0], L.cidx[1], L.cidx[2], L.fa[0], L.fa[1], L.fa[2],
                  L.fw[0], L.fw[1], L.fw[2]};
  for (void *p : ptrs)
    if (p) fb_dev_free(p);
  if (li > 0 && L.ctx) fb_destroy(L.ctx);
  (void)owner;
  memset(&L, 0, sizeof(L));
}

int alloc_level_vectors(fb_context *c, MgLevel &L, bool half) {
  fb_context *lc = L.ctx;
  L.nV = lc->nV; L.r = lc->r;
  {
    FB_TRY(fb_dev_alloc(c, &L.scale, 4));
    const float init[4] = {1.f, 1.f, 0.f, 0.f};
    FB_CUDA(cudaMemcpyAsync(L.scale, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
    const unsigned int inf = 0x7f800000u;
    FB_CUDA(cudaMemcpyAsync(L.scale + 2, &inf, sizeof(inf), cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
    (void)half;
  }
  FB_TRY(fb_dev_alloc(c, &L.Binv, 9 * (size_t)lc->nV));
  float **vecs[] = {&L.b, &L.x, &L.xn, &L.res, &L.pv};
  for (float **v : vecs) {
    FB_TRY(fb_dev_alloc(c, v, (size_t)VS * lc->nV + 4));   // (x, y, z, 0) per vertex; the fourth entries stay zero for ever
    FB_CUDA(cudaMemsetAsync(*v, 0, sizeof(float) * ((size_t)VS * lc->nV + 4), c->stream));
  }
  int perSM = 1;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_mg_spmv<__half, 1, true>, MG_TB, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
  size_t want = ((size_t)lc->nV + 2 * (MG_TB / 16) - 1) / (2 * (MG_TB / 16));   // two block rows per 16-lane group and trip
  size_t cap = (size_t)c->sm_count * perSM;
  if (cap > MG_SLOTS) cap = MG_SLOTS;
  L.grid_spmv = (int)std::max<size_t>(1, std::min(want, cap));
  L.grid_vec = grid_for_n(c, (size_t)lc->r);
  L.lmax = 0.f;
  return FB_OK;
}

// the cycle's product on one level, storage type chosen at run time
template <int MODE, bool DOT>
void launch_mg_spmv(fb_context *c, const FbMg *mg, const MgLevel &L, bool pdl, const float *x, const float *b, float omega, float *out,
                    double *slots, const FbScalars *sc, float gamma = 0.f, int usePrev = 0);

void drop_subcycle_graph(FbMg *mg);

// Storage of one level's reduced-precision matrix: slot-major (structured) when the level is a tensor grid whose rows use at
// most 16 distinct grid offsets and FP16 storage is on; otherwise the padded block rows of k_mg_spmv.
int setup_level_matrix(fb_context *c, FbMg *mg, MgLevel &L, int li) {
  fb_context *lc = L.ctx;
  L.AE = nullptr; L.nSlots = 0;
  const bool grid = L.n[0] > 0 && (long long)L.n[0] * L.n[1] * L.n[2] == (long long)lc->nV;
  if (mg->half && mg->useEll && grid && lc->nV > 0 && lc->nV >= mg->ellMinV) {
    std::vector<int> bp, bc;
    FB_TRY(fb_fetch_structure(lc, bp, bc));
    const int ny = L.n[1], nz = L.n[2];
    bool used[27] = {false}, ok = true;
    for (int v = 0; v < lc->nV && ok; v++) {
      const int vk = v % nz, vj = (v / nz) % ny, vi = v / (nz * ny);
      for (int p = bp[v]; p < bp[v + 1]; p++) {
        const int cc = bc[p];
        const int dk = cc % nz - vk, dj = (cc / nz) % ny - vj, di = cc / (nz * ny) - vi;
        if (di < -1 || di > 1 || dj < -1 || dj > 1 || dk < -1 || dk > 1) { ok = false; break; }
        used[(di + 1) * 9 + (dj + 1) * 3 + (dk + 1)] = true;
      }
    }
    int n = 0;
    for (int o = 0; o < 27; o++) {
      L.slotOf[o] = -1;
      if (used[o]) { if (n < 16) { L.slotOf[o] = (signed char)n; L.slotOff[n] = (o / 9 - 1) * ny * nz + ((o / 3) % 3 - 1) * nz + (o % 3 - 1); } n++; }
    }
    if (ok && n <= 16 && used[13]) {
      L.nSlots = n;
      FB_TRY(fb_dev_alloc(c, &L.AE, (size_t)n * lc->nV * 12 + 8));
      FB_CUDA(cudaMemsetAsync(L.AE, 0, sizeof(__half) * ((size_t)n * lc->nV * 12 + 8), c->stream));   // empty slots stay zero for ever
      FB_CUDA(cudaMemcpyAsync(mg->slotOfDev + 27 * li, L.slotOf, 27, cudaMemcpyHostToDevice, c->stream));
      FB_CUDA(cudaStreamSynchronize(c->stream));
      int perSM = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_mg_spmv_ell<1, true>, ELL_T, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
      size_t cap = (size_t)c->sm_count * perSM;
      if (cap > MG_SLOTS) cap = MG_SLOTS;
      const size_t tiles = ((size_t)lc->nV + ELL_T - 1) / ELL_T;
      L.grid_ell = (int)std::max<size_t>(1, std::min(tiles, cap));
      FB_CUDA(cudaFuncSetAttribute(k_mg_pack_ell, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PK_SMEM));
      L.grid_pack = (int)std::max<size_t>(1, std::min(((size_t)lc->nV + PK_V - 1) / PK_V, (size_t)c->sm_count * 5));
      return FB_OK;
    }
  }
  unsigned char *ab = nullptr;   // 12 entries per block, 2 (FP16) or 4 (FP32) bytes each
  FB_TRY(fb_dev_alloc(c, &ab, (size_t)lc->nB * 12 * (mg->half ? 2 : 4) + 64));
  L.AB = ab;
  return FB_OK;
}

// lambda_max(Binv A) of one level by power iteration (host reads the norms: only at setup and every 128 solves)
int estimate_lmax(fb_context *c, FbMg *mg, MgLevel &L, int its) {
  cudaStream_t st = c->stream;
  fb_context *lc = L.ctx;
  if (L.r == 0) { L.lmax = 1.f; return FB_OK; }
  const bool cold = L.lmax == 0.f;
  if (cold) k_mg_fill_pattern<<<L.grid_vec, MG_TB, 0, st>>>((size_t)L.r, lc->rowmask, L.pv);
  double lam = L.lmax;
  std::vector<double> h((size_t)L.grid_vec);
  for (int k = 0; k < its; k++) {
    launch_mg_spmv<2, false>(c, mg, L, false, L.pv, nullptr, 0.f, L.res, nullptr, nullptr);
    k_mg_norm2<<<L.grid_vec, MG_TB, 0, st>>>((size_t)VS * L.nV, L.res, mg->slotsM);
    k_mg_norm2<<<L.grid_vec, MG_TB, 0, st>>>((size_t)VS * L.nV, L.pv, mg->slotsZ);
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
    if (getenv("FEMBRAIN_B200_MG_TRACE")) fprintf(stderr, "[mg lmax] nV %d it %d lambda %.6f\n", L.nV, k, lam);
    k_mg_scale_copy<<<L.grid_vec, MG_TB, 0, st>>>((size_t)VS * L.nV, (float)(1.0 / std::sqrt(ny)), L.res, L.pv);
    c->launches++;
  }
  L.lmax = (float)((lam > 0.0 && std::isfinite(lam)) ? lam : 3.0);
  drop_subcycle_graph(mg);   // omega is a constant of the captured kernels
  return FB_OK;
}

template <int MODE, bool DOT>
void launch_mg_spmv(fb_context *c, const FbMg *mg, const MgLevel &L, bool pdl, const float *x, const float *b, float omega, float *out,
                    double *slots, const FbScalars *sc, float gamma, int usePrev) {
  const fb_context *lc = L.ctx;
  if (L.AE && mg->useEll) {
    EllArgs g;
    g.nV = L.nV; g.nSlots = L.nSlots; g.nz = L.n[2]; g.nynz = L.n[1] * L.n[2];
    for (int q = 0; q < 16; q++) { g.slotOff[q] = L.slotOff[q]; g.slotSeg[q] = 4; g.slotDk[q] = 0; }
    for (int o27 = 0; o27 < 27; o27++)
      if (L.slotOf[o27] >= 0) { g.slotSeg[L.slotOf[o27]] = (signed char)(o27 / 3); g.slotDk[L.slotOf[o27]] = (signed char)(o27 % 3 - 1); }
    fb_launch(pdl, c->stream, k_mg_spmv_ell<MODE, DOT>, L.grid_ell, ELL_T, g, (const __half *)L.AE, (const float *)L.scale, x, b,
              (const float *)L.Binv, (const unsigned char *)lc->rowmask, omega, gamma, usePrev, out, slots, sc);
    return;
  }
  if (mg->half)
    fb_launch(pdl, c->stream, k_mg_spmv<__half, MODE, DOT>, L.grid_spmv, MG_TB, L.nV, lc->bp, lc->bc, (const __half *)L.AB, (const float *)L.scale, x, b,
              (const float *)L.Binv, (const unsigned char *)lc->rowmask, omega, gamma, usePrev, out, slots, sc);
  else
    fb_launch(pdl, c->stream, k_mg_spmv<float, MODE, DOT>, L.grid_spmv, MG_TB, L.nV, lc->bp, lc->bc, (const float *)L.AB, (const float *)L.scale, x, b,
              (const float *)L.Binv, (const unsigned char *)lc->rowmask, omega, gamma, usePrev, out, slots, sc);
}

// Damped block Jacobi: w = 1.4 / lambda_max every sweep.  Chebyshev (nu >= 2): the three-term recurrence for the interval
// [a, b] = [hi lambda_max / alpha, hi lambda_max] of Binv A (Saad, Iterative Methods, alg. 12.1): theta = (a + b) / 2,
// delta = (b - a) / 2, rho_0 = delta / theta, rho_k = 1 / (2 theta / delta - rho_(k-1)); w_0 = 1 / theta, w_k = 2 rho_k / delta,
// gm_k = rho_k rho_(k-1).  Both are polynomials in Binv A times Binv: symmetric, the same before and after the correction.
void smoother_coefficients(const FbMg *mg, float lmax, float *w, float *gm) {
  for (int k = 0; k < 8; k++) { w[k] = 1.4f / lmax; gm[k] = 0.f; }
  if (!mg->cheb || mg->nu < 2) return;
  const double b = (double)mg->chebHi * lmax, a = b / mg->chebAlpha, theta = 0.5 * (a + b), delta = 0.5 * (b - a);
  double rho = delta / theta;
  w[0] = (float)(1.0 / theta);
  for (int k = 1; k < 8; k++) {
    const double rn = 1.0 / (2.0 * theta / delta - rho);
    w[k] = (float)(2.0 * rn / delta);
    gm[k] = (float)(rn * rho);
    rho = rn;
  }
}

float *vcycle(fb_context *c, FbMg *mg, int li);

// The sub-cycle of levels >= 1 is ~30 kernels of a few microseconds each: launch latency, not work (185 us of a 1.08 ms
// iteration at 10M tets, profiles/r02_mg_iteration_phases.txt).  Captured once per hierarchy / lambda_max estimate into a CUDA
// graph (pointers and omegas are constants of the capture) and replayed with one launch per cycle.
float *subcycle_graph(fb_context *c, FbMg *mg) {
  cudaStream_t st = c->stream;
  if (mg->subFailed) return nullptr;
  if (!mg->subGraph) {
    const long long before = c->launches;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); mg->subFailed = 1; return nullptr; }
    mg->capturing = 1;
    mg->subResult = vcycle(c, mg, 1);
    mg->capturing = 0;
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    mg->subKernels = (int)(c->launches - before);
    c->launches = before;
    if (e != cudaSuccess || !graph) { cudaGetLastError(); mg->subFailed = 1; return nullptr; }
    cudaGraphExec_t exec = nullptr;
    e = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { cudaGetLastError(); mg->subFailed = 1; return nullptr; }
    mg->subGraph = exec;
  }
  if (cudaGraphLaunch((cudaGraphExec_t)mg->subGraph, st) != cudaSuccess) { cudaGetLastError(); mg->subFailed = 1; return nullptr; }
  c->launches += mg->subKernels;
  return mg->subResult;
}

void drop_subcycle_graph(FbMg *mg) {
  if (mg->subGraph) { cudaGraphExecDestroy((cudaGraphExec_t)mg->subGraph); mg->subGraph = nullptr; }
  mg->subFailed = 0;
}

// z = V(b) on level li; returns the vector holding the result (li = 0: with the dot b.z into slotsZ)
float *vcycle(fb_context *c, FbMg *mg, int li) {
  cudaStream_t st = c->stream;
  MgLevel &L = mg->L[li];
  fb_context *lc = L.ctx;
  const FbScalars *sc = c->sc;
  if (li == mg->nLevels - 1) {   // coarsest: dense inverse
    k_mg_dense_apply<<<1, 128, 0, st>>>(L.r, mg->denseInv, L.b, L.xn, sc);
    c->launches++;
    return L.xn;
  }
  MgLevel &C = mg->L[li + 1];
  const GridMaps g = maps_of(L);
  const int nu = (li > 0 && mg->nuCoarse > 0) ? mg->nuCoarse : mg->nu;
  // step k of the smoother: x_(k+1) = x_k + w[k] Binv (b - A x_k) + gm[k] (x_k - x_(k-1))
  float w[8], gm[8];
  smoother_coefficients(mg, L.lmax, w, gm);
  const float omega = w[0];
  const bool pdl = !mg->capturing;   // inside a captured sub-cycle the graph's own edges order the kernels
  // pre-smoothing from the zero guess: the first sweep needs no product (on the finest level the CG update kernel has
  // already written it together with r32, except before the first iteration)
  if (!(li == 0 && mg->preDone)) {
    fb_launch(pdl, st, k_mg_presmooth<false>, L.grid_vec, MG_TB, L.nV, L.Binv, L.b, omega, L.x, (double *)nullptr, sc);
    c->launches++;
  }
  float *cur = L.x, *alt = L.xn;
  for (int s = 1; s < nu; s++) {
    launch_mg_spmv<1, false>(c, mg, L, pdl, cur, L.b, w[s], alt, nullptr, sc, gm[s], s > 1);   // x_0 = 0
    c->launches++;
    std::swap(cur, alt);
  }
  launch_mg_spmv<0, false>(c, mg, L, pdl, cur, L.b, omega, L.res, nullptr, sc);
  if (li == 0) tm_mark(st, "presmooth + residual (level 0)");
  k_mg_restrict<<<C.grid_vec, MG_TB, 0, st>>>(g, L.res, C.ctx->rowmask, C.b, (float *)nullptr, sc);
  c->launches += 2;
  if (li == 0) tm_mark(st, "restrict (level 0 -> 1)");
  float *coarse = nullptr;
  if (li == 0 && mg->useGraph && !mg->capturing) coarse = subcycle_graph(c, mg);   // levels >= 1 as ONE graph launch (~30 small kernels)
  if (!coarse) coarse = vcycle(c, mg, li + 1);
  if (li == 0) tm_mark(st, "levels >= 1");
  k_mg_prolong_add<<<L.grid_vec, MG_TB, 0, st>>>(g, coarse, lc->rowmask, cur, mg->coarseScale, sc);
  c->launches++;
  if (li == 0) tm_mark(st, "prolong (level 1 -> 0)");
  for (int s = 0; s < nu; s++) {   // the same sweeps after the correction: the cycle stays symmetric
    const bool last = s == nu - 1;
    if (li == 0 && last) launch_mg_spmv<1, true>(c, mg, L, pdl, cur, L.b, w[s], alt, mg->slotsZ, sc, gm[s], 1);
    else launch_mg_spmv<1, false>(c, mg, L, pdl, cur, L.b, w[s], alt, nullptr, sc, gm[s], 1);
    c->launches++;
    std::swap(cur, alt);
  }
  return cur;
}

// z = preconditioner(r32) on the finest level: result in *z, r.z partials in slotsZ (nSlotsZ returned)
int apply_preconditioner(fb_context *c, FbMg *mg, float **z) {
  MgLevel &L = mg->L[0];
  *z = L.xn;
  if (mg->variant == FB_SOLVER_MG_PCG && mg->nLevels > 1) {
    *z = vcycle(c, mg, 0);
    mg->vcycles++;
    return (L.AE && mg->useEll) ? L.grid_ell : L.grid_spmv;
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
  drop_subcycle_graph(mg);
  for (int li = mg->nLevels - 1; li >= 0; li--) free_level(c, mg->L[li], li);
  if (mg->denseInv) fb_dev_free(mg->denseInv);
  if (mg->slotsM) fb_dev_free(mg->slotsM);
  if (mg->slotOfDev) fb_dev_free(mg->slotOfDev);
  if (mg->slotsZ) fb_dev_free(mg->slotsZ);
  delete mg;
  c->mg = nullptr;
}

static int mg_ensure(fb_context *c) {
  if (c->mg) return FB_OK;
  FbMg *mg = new FbMg();
  memset(mg, 0, sizeof(*mg));
  c->mg = mg;
  mg->half = !(getenv("FEMBRAIN_B200_MG_PREC") && !strcmp(getenv("FEMBRAIN_B200_MG_PREC"), "fp32"));
  // FEMBRAIN_B200_MG_ELL: unset = levels of >= 200,000 vertices (measured: 5 % faster steps at 1.77M vertices, even at 185k,
  // slower on small levels where one thread per row cannot fill 148 SMs); 1 = every tensor-grid level; 0 = none;
  // FEMBRAIN_B200_MG_ELL_MIN = the vertex threshold
  mg->useEll = !(getenv("FEMBRAIN_B200_MG_ELL") && atoi(getenv("FEMBRAIN_B200_MG_ELL")) == 0);
  mg->ellMinV = getenv("FEMBRAIN_B200_MG_ELL") ? 0 : 200000;   // (227k-vertex level 1 of the 10M-tet cube: -0.5 ms per step; 185k: +0.06)
  if (getenv("FEMBRAIN_B200_MG_ELL_MIN")) mg->ellMinV = atoi(getenv("FEMBRAIN_B200_MG_ELL_MIN"));
  mg->useGraph = !(getenv("FEMBRAIN_B200_MG_GRAPH") && atoi(getenv("FEMBRAIN_B200_MG_GRAPH")) == 0);
  mg->nu = 3;   // measured at 10M / 1M tets (profiles/r02_mg_smoother_sweep.txt): 13-17 iterations, the fastest step of nu = 1..4
  mg->cheb = !(getenv("FEMBRAIN_B200_MG_SMOOTHER") && !strcmp(getenv("FEMBRAIN_B200_MG_SMOOTHER"), "jacobi"));
  mg->chebAlpha = getenv("FEMBRAIN_B200_MG_CHEB_ALPHA") ? (float)atof(getenv("FEMBRAIN_B200_MG_CHEB_ALPHA")) : 20.f;
  mg->chebHi = getenv("FEMBRAIN_B200_MG_CHEB_HI") ? (float)atof(getenv("FEMBRAIN_B200_MG_CHEB_HI")) : 1.1f;
  if (!(mg->chebAlpha > 1.f)) mg->chebAlpha = 20.f;
  mg->nuCoarse = getenv("FEMBRAIN_B200_MG_NU_COARSE") ? std::max(0, std::min(8, atoi(getenv("FEMBRAIN_B200_MG_NU_COARSE")))) : 4;   // measured: 1 / 2 / 3 / 4 / 6 / 8 sweeps on the coarse levels -> 23-36 / 15-22 / 13-18 / 12-16 / 11-15 / 11-14 iterations at 10M tets; 4 is the fastest step
  mg->coarseScale = getenv("FEMBRAIN_B200_MG_CSCALE") ? (float)atof(getenv("FEMBRAIN_B200_MG_CSCALE")) : 1.f;
  if (!(mg->coarseScale > 0.f)) mg->coarseScale = 1.f;
  if (getenv("FEMBRAIN_B200_MG_NU") && atoi(getenv("FEMBRAIN_B200_MG_NU")) > 0) mg->nu = std::min(8, atoi(getenv("FEMBRAIN_B200_MG_NU")));
  FB_TRY(fb_dev_alloc(c, &mg->slotsM, 2 * (size_t)MG_SLOTS));
  FB_TRY(fb_dev_alloc(c, &mg->slotOfDev, 27 * (size_t)MG_MAX_LEVELS));
  FB_TRY(fb_dev_alloc(c, &mg->slotsZ, (size_t)MG_SLOTS));
  return FB_OK;
}

// builds (or rebuilds) the level hierarchy for the current variant; level 0 always exists for variants != 0
static int mg_build(fb_context *c) {
  FbMg *mg = c->mg;
  drop_subcycle_graph(mg);
  for (int li = mg->nLevels - 1; li >= 0; li--) free_level(c, mg->L[li], li);
  mg->nLevels = 0;
  mg->prepared = 0;
  if (mg->denseInv) { fb_dev_free(mg->denseInv); mg->denseInv = nullptr; }
  if (mg->variant == FB_SOLVER_JACOBI_PCG) return FB_OK;
  MgLevel &L0 = mg->L[0];
  memset(&L0, 0, sizeof(L0));
  L0.ctx = c;
  FB_TRY(alloc_level_vectors(c, L0, mg->half != 0));
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
    FB_TRY(alloc_level_vectors(c, C, mg->half != 0));
    FB_TRY(fb_dev_alloc(c, &C.u, (size_t)lc->r));
    FB_TRY(mg_upload(c, &C.twin, twinF));   // straight from the FINEST level: injection needs no intermediate vectors
    c->bytes += lc->bytes;
    twinFine.swap(twinF);
    for (int d = 0; d < 3; d++) ax[d].swap(cax[d]);
    li++;
  }
  for (int k = 0; k < mg->nLevels; k++) FB_TRY(setup_level_matrix(c, mg, mg->L[k], k));
  MgLevel &Lc = mg->L[mg->nLevels - 1];
  if (mg->nLevels > 1) {
    if (Lc.r > MG_MAX_DENSE) { fb_set_error("multigrid: coarsest level has %d unknowns (> %d)", Lc.r, MG_MAX_DENSE); return FB_ERR_NOT_SUPPORTED; }
    mg->nDense = Lc.r;
    FB_CUDA(cudaFuncSetAttribute(k_mg_dense_inverse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * MG_MAX_DENSE * MG_MAX_DENSE)));
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
  // developer aid (FEMBRAIN_B200_MG_TIMING=2): host clock around synchronised phases of the per-step preparation
  const bool laps = getenv("FEMBRAIN_B200_MG_TIMING") && atoi(getenv("FEMBRAIN_B200_MG_TIMING")) == 2;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](int li, const char *what) {
    if (!laps) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[mg prepare] level %d %-24s %8.1f us\n", li, what, std::chrono::duration<double, std::micro>(t1 - t0).count());
    t0 = std::chrono::steady_clock::now();
  };
  lap(-1, "(idle)");
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
      lap(li, "inject + assembly");
    }
    if (mg->variant == FB_SOLVER_MG_PCG && mg->nLevels > 1) {   // (the one-level variant applies Binv only)
      k_mg_scale_min<<<L.grid_vec, MG_TB, 0, st>>>((size_t)lc->r, lc->invD, reinterpret_cast<unsigned int *>(L.scale + 2));
      k_mg_scale_final<<<1, 1, 0, st>>>(L.scale, mg->half);
      if (L.AE) k_mg_pack_ell<<<L.grid_pack, PK_T, PK_SMEM, st>>>(lc->nV, L.n[1], L.n[2], lc->bp, lc->bc, lc->Keff, L.scale, mg->slotOfDev + 27 * li, L.AE);
      else if (mg->half) k_mg_pack<__half><<<grid_for_n(c, (size_t)lc->nB), MG_TB, 0, st>>>(lc->nB, lc->bp, lc->brow, lc->Keff, L.scale, (__half *)L.AB);
      else k_mg_pack<float><<<grid_for_n(c, (size_t)lc->nB), MG_TB, 0, st>>>(lc->nB, lc->bp, lc->brow, lc->Keff, L.scale, (float *)L.AB);
      lap(li, "scale + pack");
    }
    k_mg_block_inverse<<<(L.nV + 127) / 128, 128, 0, st>>>(L.nV, lc->bp, lc->diag, lc->Keff, lc->rowmask, L.Binv);
    c->launches += 4;
    lap(li, "block inverse");
    if (getenv("FEMBRAIN_B200_MG_SYNC")) {
      cudaError_t e = cudaStreamSynchronize(st);
      fprintf(stderr, "[mg sync] level %d (nV %d, nSlots %d, AE %p AB %p) after pack + block inverse: %s\n", li, L.nV, L.nSlots, (void *)L.AE, L.AB, cudaGetErrorString(e));
    }
  }
  if (mg->nLevels > 1) {
    MgLevel &Lc = mg->L[mg->nLevels - 1];
    k_mg_dense_inverse<<<1, MG_DENSE_T, sizeof(double) * (size_t)Lc.r * Lc.r, st>>>(Lc.nV, Lc.ctx->bp, Lc.ctx->bc, Lc.ctx->Keff, Lc.ctx->rowmask, mg->denseInv);
    c->launches += 1;
    lap(mg->nLevels - 1, "dense build + invert");
    const bool first = mg->L[0].lmax == 0.f;
    if (first || mg->solves >= 128) {
      for (int li = 0; li + 1 < mg->nLevels; li++) FB_TRY(estimate_lmax(c, mg, mg->L[li], first ? (getenv("FEMBRAIN_B200_MG_LMAX_ITS") ? atoi(getenv("FEMBRAIN_B200_MG_LMAX_ITS")) : 30) : 3));
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
  drop_subcycle_graph(mg);
  for (int li = mg->nLevels - 1; li >= 0; li--) free_level(c, mg->L[li], li);
  mg->nLevels = 0;
  mg->prepared = 0;
}

int fb_mg_active(const fb_context *c) { return (c->mg && c->mg->variant != FB_SOLVER_JACOBI_PCG) ? c->mg->variant : 0; }
int fb_mg_warm(const fb_context *c) { return c->mg ? c->mg->warm : 0; }

// PCG with the variant's preconditioner on Keff x = rhs (masked).  Same contract as fb_pcg_solve.
static int mg_pcg_solve_once(fb_context *c, double eps, int maxIt) {
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
  float *z = nullptr;
  mg->preDone = 0;
  int nZ = apply_preconditioner(c, mg, &z);
  const bool cyc = mg->variant == FB_SOLVER_MG_PCG && mg->nLevels > 1;
  float w0[8], g0[8];
  smoother_coefficients(mg, L.lmax, w0, g0);
  const float omega0 = cyc ? w0[0] : 0.f;
  k_mgcg_begin<<<gv, MG_TB, 0, st>>>(n, z, c->dir, c->sc, slotsM0, slotsM, mg->slotsZ, gv, nZ, eps, maxIt);
  c->launches++;
  const int CH = 6;
  g_tm.on = getenv("FEMBRAIN_B200_MG_TIMING") && atoi(getenv("FEMBRAIN_B200_MG_TIMING")) != 0;
  static const bool trace = getenv("FEMBRAIN_B200_MG_TRACE") && atoi(getenv("FEMBRAIN_B200_MG_TRACE")) != 0;
  int it = 1, slot = 0, pending = 0;
  bool finished = false;
  while (!finished && it <= maxIt) {
    const int end = (it + CH - 1 < maxIt) ? it + CH - 1 : maxIt;
    for (; it <= end; it++) {
      const bool sample = c->profiling && (it % 4 == 1) && c->nprof < 64;
      if (sample) cudaEventRecord(c->evProf[2 * c->nprof], st);
      int nDq = 0;
      g_tm.armed = g_tm.on && it == 3;
      g_tm.n = 0;
      tm_mark(st, "start");
      FB_TRY(fb_pcg_launch_product_dq(c, c->dir, c->Ad, &nDq));    // q = A d, d.q partials in c->partials
      tm_mark(st, "FP64 product A d");
      if (sample) { cudaEventRecord(c->evProf[2 * c->nprof + 1], st); c->nprof++; }
      fb_launch(true, st, k_mgcg_update, gv, MG_TB, L.nV, c->dir, c->Ad, c->invD, c->x, c->res, L.b, (const FbScalars *)c->sc, it,
                (const double *)c->partials, nDq, slotsM, (const float *)(cyc ? L.Binv : nullptr), omega0, L.x);
      mg->preDone = cyc ? 1 : 0;
      fb_launch(true, st, k_mgcg_check, 1, MG_TB, c->sc, it, (const double *)slotsM, gv);
      c->launches += 2;
      tm_mark(st, "update + check");
      nZ = apply_preconditioner(c, mg, &z);
      tm_mark(st, "postsmooth (level 0)");
      fb_launch(true, st, k_mgcg_direction, gv, MG_TB, n, (const float *)z, c->dir, c->sc, it, (const double *)mg->slotsZ, nZ);
      c->launches++;
      tm_mark(st, "direction");
      if (g_tm.armed) {
        cudaStreamSynchronize(st);
        for (int k = 1; k < g_tm.n; k++) {
          float ms = 0;
          cudaEventElapsedTime(&ms, g_tm.ev[k - 1], g_tm.ev[k]);
          fprintf(stderr, "[mg timing] %-36s %8.1f us\n", g_tm.name[k], 1e3 * ms);
        }
        g_tm.armed = false;
      }
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

// The multigrid cycle with Chebyshev smoothing is only a valid (positive definite) preconditioner while the smoothing interval
// covers the spectrum of Binv A: if lambda_max has grown past 1.1 x the running power-iteration estimate (the matrix changes
// every step), CG stalls or diverges.  The healthy solver needs < 100 iterations on every mesh tried, so the first attempt is
// capped at 300; on failure lambda_max is re-estimated on every level (30 more iterations), the interval's safety factor is
// raised to >= 1.3 for the rest of the context's life, and the solve is repeated from x0 = 0 with the caller's limit.
int fb_mg_pcg_solve(fb_context *c, double eps, int maxIt) {
  FbMg *mg = c->mg;
  const bool guarded = mg->variant == FB_SOLVER_MG_PCG && mg->cheb && mg->nu >= 2 && maxIt > 300;
  FB_TRY(mg_pcg_solve_once(c, eps, guarded ? 300 : maxIt));
  if (!guarded || c->last_iters >= 0 || mg->nLevels < 2) return FB_OK;
  const int spent = -c->last_iters;
  for (int li = 0; li + 1 < mg->nLevels; li++) FB_TRY(estimate_lmax(c, mg, mg->L[li], 30));
  if (mg->chebHi < 1.3f) mg->chebHi = 1.3f;
  drop_subcycle_graph(mg);
  mg->fallbacks++;
  c->have_solution = 0;
  if (getenv("FEMBRAIN_B200_MG_TRACE")) fprintf(stderr, "[mg] no convergence in %d iterations: lambda_max re-estimated (%.4f on level 0), interval factor %.2f, retrying\n", spent, mg->L[0].lmax, mg->chebHi);
  FB_TRY(mg_pcg_solve_once(c, eps, maxIt));
  if (c->last_iters > 0) c->last_iters += spent;   // the iterations the caller paid for
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
  if (c->mg->variant == FB_SOLVER_MG_PCG) {
    const int st = mg_build(c);
    if (st != FB_OK) { fb_mg_invalidate(c); c->mg->variant = FB_SOLVER_JACOBI_PCG; c->prm.solver_variant = FB_SOLVER_JACOBI_PCG; }
    return st;
  }
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
  const int st = mg_build(c);
  if (st != FB_OK) {   // e.g. out of memory while creating a level: fall back to the reference's solver, context stays usable
    fb_mg_invalidate(c);
    c->mg->variant = FB_SOLVER_JACOBI_PCG;
    c->prm.solver_variant = FB_SOLVER_JACOBI_PCG;
  }
  return st;
}

int fb_get_solver(const fb_context *c, int *variant, int *warm_start, int *levels) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  if (variant) *variant = c->mg ? c->mg->variant : FB_SOLVER_JACOBI_PCG;
  if (warm_start) *warm_start = c->mg ? c->mg->warm : 0;
  if (levels) *levels = c->mg ? c->mg->nLevels : 0;
  return FB_OK;
}

int fb_get_solver_levels(const fb_context *c, int capacity, int *num_vertices, long long *num_blocks) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  const int n = c->mg ? c->mg->nLevels : 0;
  for (int i = 0; i < n && i < capacity; i++) {
    if (num_vertices) num_vertices[i] = c->mg->L[i].ctx->nV;
    if (num_blocks) num_blocks[i] = c->mg->L[i].ctx->nB;
  }
  return FB_OK;
}

int fb_get_solver_smoother(const fb_context *c, int *sweeps, int *chebyshev, double *alpha, int *structured_levels) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  const FbMg *mg = c->mg;
  if (sweeps) *sweeps = mg ? mg->nu : 0;
  if (chebyshev) *chebyshev = (mg && mg->cheb && mg->nu >= 2) ? 1 : 0;
  if (alpha) *alpha = mg ? (double)mg->chebAlpha : 0.0;
  int n = 0;
  if (mg) for (int i = 0; i < mg->nLevels; i++) n += mg->L[i].AE != nullptr;
  if (structured_levels) *structured_levels = n;
  return FB_OK;
}

const char *fb_solver_name(int variant) {
  switch (variant) {
    case FB_SOLVER_JACOBI_PCG: return "jacobi_pcg (the reference's algorithm, CGSolver.cpp:129-190)";
    case FB_SOLVER_BLOCK_JACOBI_PCG: return "block_jacobi_pcg (variant: 3x3 block-diagonal preconditioner, FP32 apply)";
    case FB_SOLVER_MG_PCG: return "mg_pcg (variant: geometric multigrid V-cycle preconditioner, Chebyshev block-Jacobi smoothing, re-assembled coarse levels, FP16/FP32 cycle, FP64 CG)";
    default: return "unknown";
  }
}

}  // extern "C"
