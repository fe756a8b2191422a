// fb_pcg_persistent.cu — the whole Jacobi-PCG loop of CGSolver::SolveLinearSystemWithJacobiPreconditioner
// (reference src/3rdparty/vegafem/sparseSolver/CGSolver.cpp:129-190) in ONE cooperative kernel.
//
// Why: with three kernels per iteration the 1M-tet step spends ~15 of every 56 us between kernels (launch
// latency, drain and ramp-up of three dependent launches; ncu launch list profiles/r01_launches_pcg.csv) — and
// far more than that on the 200k-tet meshes of the batch configuration.  Here the grid is exactly one resident
// wave (launched with cudaLaunchCooperativeKernel, so co-residency is guaranteed and grid.sync() is safe),
// every CTA owns a contiguous range of block rows holding an equal share of the 3x3 blocks, and the three
// dependency points of an iteration (d.q -> alpha, sum r^2/diag -> beta, d visible before the next SpMV) are
// grid barriers instead of kernel boundaries.
//
// Arithmetic is the same as in the three-kernel path: same recurrences, refresh period and stopping rule; each
// dot product is a per-CTA sum (fixed shuffle tree) written to a slot, and after the barrier EVERY CTA adds all
// slots in the same fixed order, so all CTAs compute bit-identical alpha/beta/rho and take the same branch.
// Vectors that other CTAs write inside the kernel (d, x) are read with ordinary (coherent, L1-cacheable) loads,
// never the non-coherent path: grid.sync() is a gpu-scope acquire, which invalidates L1, so the gathers after it
// see the new values and still coalesce/hit in L1 within an iteration (an L2-only __ldcg gather moved 4.7x the
// matrix bytes in 32-byte sectors and ran at half speed).  Matrix values stream with ld.global.nc.L1::no_allocate.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "../fb_internal.h"
#include "../fb_pcg_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int PTB = 256;

struct PersistArgs {
  int nV;
  const int *ctaRows, *bp, *bc;
  const double *A;
  const unsigned char *mask;
  const double *b, *invD;
  double *x, *r, *d, *q;
  FbScalars *sc;
  double *slotsA, *slotsB;
  unsigned long long *prof;
  int profiling;
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// per-CTA sum with a fixed tree; valid in thread 0
__device__ __forceinline__ double block_sum(double v, double *s_w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (warp == 0) {
    s = (lane < PTB / 32) ? s_w[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  }
  __syncthreads();
  return s;
}

// every CTA adds all slots in the same order: lane i takes slots i, i+32, ..., then an xor butterfly
__device__ __forceinline__ double all_slots_sum(const double *slots, int n, double *s_bcast) {
  if (threadIdx.x < 32) {
    double s = 0.0;
    for (int k = threadIdx.x; k < n; k += 32) s += __ldcg(slots + k);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) *s_bcast = s;
  }
  __syncthreads();
  const double total = *s_bcast;
  __syncthreads();
  return total;
}

// rows [rowBeg, rowEnd) of y = mask(A xin) (MODE 1, returns sum xin.y) or y = mask(b - A xin) (MODE 2, returns
// sum y^2 invD); 16 lanes per block row, all loads of a row in flight (as k_spmv_rows3)
template <int MODE>
__device__ __forceinline__ double spmv_range(const PersistArgs &a, int rowBeg, int rowEnd, const double *xin, double *y) {
  const int lane = threadIdx.x & (TILE_G - 1);
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
  const int groups = PTB / TILE_G;
  double part = 0.0;
  int v = rowBeg + threadIdx.x / TILE_G;
  int rs = 0, re = 0;
  if (v < rowEnd) { rs = __ldg(a.bp + v); re = __ldg(a.bp + v + 1); }
  while (v < rowEnd) {
    const int vn = v + groups;
    int rsn = 0, ren = 0;
    if (vn < rowEnd) { rsn = __ldg(a.bp + vn); ren = __ldg(a.bp + vn + 1); }
    const int n3 = 3 * (re - rs);
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int base = 0; base < n3; base += TILE_CHUNK) {
      RowVals val;
      int col[3];
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int t = base + lane + TILE_G * p;
        col[p] = (t < n3) ? __ldg(a.bc + rs + t / 3) : -1;
      }
      load_row_chunk(a.A, rs, n3, base, lane, val);
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int t = base + lane + TILE_G * p;
        const double xv = (col[p] >= 0) ? xin[3 * (size_t)col[p] + (t % 3)] : 0.0;
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
      if (MODE == 1) {
        if (__ldg(a.mask + row)) s = 0.0;
        y[row] = s;
        part = fma(xin[row], s, part);
      } else {
        const double rres = __ldg(a.mask + row) ? 0.0 : (__ldg(a.b + row) - s);
        y[row] = rres;
        part += (rres * rres) * __ldg(a.invD + row);
      }
    }
    v = vn; rs = rsn; re = ren;
  }
  return part;
}

__global__ void __launch_bounds__(PTB) k_pcg_persistent(PersistArgs a) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double s_w[PTB / 32];
  __shared__ double s_bcast;
  const int tid = threadIdx.x;
  const int nCta = gridDim.x;
  const int rowBeg = a.ctaRows[blockIdx.x], rowEnd = a.ctaRows[blockIdx.x + 1];
  const size_t dBeg = 3 * (size_t)rowBeg, dEnd = 3 * (size_t)rowEnd;
  double rho = a.sc->rho[0];
  const double rho0 = a.sc->rho0, eps2 = a.sc->eps2;
  const int maxIt = a.sc->max_it;
  int it = 1;
  // while ((residualNorm2 > eps*eps*initialResidualNorm2) && (iteration <= maxIterations))     CGSolver.cpp:150
  while ((rho > eps2 * rho0) && (it <= maxIt)) {
    const bool sample = a.profiling && blockIdx.x == 0 && tid == 0 && (it % 16 == 1);
    unsigned long long t0 = 0;
    if (sample) t0 = global_ns();
    // q = A d, d.q
    double part = spmv_range<1>(a, rowBeg, rowEnd, a.d, a.q);
    part = block_sum(part, s_w);
    if (tid == 0) a.slotsA[blockIdx.x] = part;
    grid.sync();
    if (sample) { atomicAdd(a.prof, global_ns() - t0); atomicAdd(a.prof + 1, 1ull); }
    const double dq = all_slots_sum(a.slotsA, nCta, &s_bcast);
    const double alpha = rho / dq;
    part = 0.0;
    if (it % 30 == 0) {
      // x += alpha d, then the exact residual r = b - A x                                        CGSolver.cpp:161-167
      for (size_t i = dBeg + tid; i < dEnd; i += PTB) a.x[i] = fma(alpha, a.d[i], a.x[i]);
      grid.sync();
      part = spmv_range<2>(a, rowBeg, rowEnd, a.x, a.r);
    } else {
      for (size_t i = dBeg + tid; i < dEnd; i += PTB) {
        a.x[i] = fma(alpha, a.d[i], a.x[i]);
        const double ri = fma(-alpha, a.q[i], a.r[i]);
        a.r[i] = ri;
        part += (ri * ri) * __ldg(a.invD + i);
      }
    }
    part = block_sum(part, s_w);
    if (tid == 0) a.slotsB[blockIdx.x] = part;
    grid.sync();
    const double rhoNew = all_slots_sum(a.slotsB, nCta, &s_bcast);
    const double beta = rhoNew / rho;
    for (size_t i = dBeg + tid; i < dEnd; i += PTB) a.d[i] = fma(__ldg(a.invD + i), a.r[i], beta * a.d[i]);
    rho = rhoNew;
    it++;
    grid.sync();  // every CTA's slice of d is visible before the next SpMV gathers it
  }
  if (blockIdx.x == 0 && tid == 0) {
    a.sc->iters = it - 1;
    a.sc->rho[(it - 1) & 1] = rho;
    a.sc->done = 1;
  }
}

// first row r with bp[r] >= target block, for every CTA boundary: equal block counts per CTA, row aligned
__global__ void k_plan_cta_rows(int nV, int nB, int grid, const int *__restrict__ bp, int *__restrict__ ctaRows) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > grid) return;
  if (c == grid) { ctaRows[c] = nV; return; }
  const long long target = (long long)nB * c / grid;
  int lo = 0, hi = nV;  // bp is non-decreasing, bp[nV] = nB
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (bp[mid] >= target) hi = mid; else lo = mid + 1;
  }
  ctaRows[c] = lo;
}

}  // namespace

int fb_pcg_plan_persistent(fb_context *c) {
  c->pers_grid = 0;
  // Opt-in (FEMBRAIN_B200_PCG=persistent).  Measured on B200 (profiles/r01_pcg_persistent_vs_kernels.txt): correct, but
  // SLOWER than the kernels path — 79.5 vs 55.3 us/iteration at 1M tets, 797 vs 526 us at 10M, 29.1 vs 26.5 us at 200k:
  // cooperative-groups grid.sync() costs about as much as a kernel boundary here and the SpMV phase runs at 4
  // CTAs/SM (57 registers) behind a barrier that waits for the slowest CTA.
  const char *env = getenv("FEMBRAIN_B200_PCG");
  if (!(env && !strcmp(env, "persistent"))) return FB_OK;
  if (c->nV == 0 || c->nB == 0) return FB_OK;
  int coop = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->device) != cudaSuccess || !coop) { cudaGetLastError(); return FB_OK; }
  int perSM = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_pcg_persistent, PTB, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); return FB_OK; }
  long long grid = (long long)c->sm_count * perSM;
  const long long byRows = ((long long)c->nV + PTB / TILE_G - 1) / (PTB / TILE_G);  // at least one row per 16-lane group
  if (grid > byRows) grid = byRows;
  if (grid > FB_MAX_PARTIALS) grid = FB_MAX_PARTIALS;
  if (grid < 1) grid = 1;
  if (c->ctaRows) { fb_dev_free(c->ctaRows); c->ctaRows = nullptr; }
  FB_TRY(fb_dev_alloc(c, &c->ctaRows, (size_t)grid + 1));
  if (!c->pers_prof) {
    FB_TRY(fb_dev_alloc(c, &c->pers_prof, 2));
    FB_CUDA(cudaMemsetAsync(c->pers_prof, 0, 2 * sizeof(unsigned long long), c->stream));
  }
  k_plan_cta_rows<<<(int)((grid + 1 + 255) / 256), 256, 0, c->stream>>>(c->nV, c->nB, (int)grid, c->bp, c->ctaRows);
  c->launches++;
  FB_CUDA(cudaStreamSynchronize(c->stream));
  FB_CUDA(cudaGetLastError());
  c->pers_grid = (int)grid;
  return FB_OK;
}

int fb_pcg_launch_persistent(fb_context *c) {
  PersistArgs a;
  a.nV = c->nV;
  a.ctaRows = c->ctaRows; a.bp = c->bp; a.bc = c->bc;
  a.A = c->Keff;
  a.mask = c->rowmask;
  a.b = c->rhs; a.invD = c->invD;
  a.x = c->x; a.r = c->res; a.d = c->dir; a.q = c->Ad;
  a.sc = c->sc;
  a.slotsA = c->partials; a.slotsB = c->partials + 3 * (size_t)FB_MAX_PARTIALS;
  a.prof = c->pers_prof;
  a.profiling = c->profiling;
  void *args[] = {&a};
  FB_CUDA(cudaLaunchCooperativeKernel((const void *)k_pcg_persistent, dim3(c->pers_grid), dim3(PTB), args, 0, c->stream));
  c->launches++;
  return FB_OK;
}
