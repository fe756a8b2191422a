// fb_tma.cu — the PCG products with the matrix stream staged through shared memory by the bulk-copy engine (TMA, 1-D
// cp.async.bulk + mbarrier), opt-in with FEMBRAIN_B200_SPMV=tma.  EXPERIMENTAL: written at the end of round 1 without GPU time
// left to run it; tests/test_tma_gpu.py is skipped unless FEMBRAIN_B200_TEST_EXPERIMENTAL=1.
//
// Reference: the products of CGSolver::SolveLinearSystemWithJacobiPreconditioner (src/3rdparty/vegafem/sparseSolver/
// CGSolver.cpp:152, :165; SparseMatrix::MultiplyVector, sparseMatrix.cpp:405-413).  Same arithmetic per row as k_spmv_rows3
// (fb_pcg.cu): 16 lanes per block row, three planes, 16-lane shuffle tree, per-CTA sum in a fixed slot.
//
// Why: k_spmv_rows3 keeps ~70 kB of matrix bytes in flight per SM, all of them held by registers of stalled lanes, and ncu
// shows its L1TEX pipe 69 % busy next to 80-88 % of the HBM copy peak.  Here the matrix never passes L1TEX as a
// global load: block rows are cut into TILES of <= 32 rows / <= 480 blocks whose values, block columns and row pointers are
// each ONE contiguous byte range of Keff / bc / bp, so one producer lane per CTA brings a tile in with three bulk copies
// that complete on an mbarrier, three tiles deep, two CTAs per SM (~150 kB in flight per SM, no registers involved).  The 16
// consumer warps read values and columns with conflict-free ld.shared and only gather x (L1/L2 hits) from global memory.
#include <algorithm>
#include <vector>

#include "../fb_internal.h"
#include "../fb_pcg_common.cuh"

struct FbTma {
  int nTiles;
  int4 *tiles;   // {first block row, rows, first block, blocks}
  int grid;
  int *err;      // device flag: a bounded mbarrier wait ran out
};

namespace {

constexpr int TM_CONSUMER_WARPS = 16;
constexpr int TM_TB = 32 * (TM_CONSUMER_WARPS + 1);  // + the producer warp
constexpr int TM_ROWS = 2 * TM_CONSUMER_WARPS;       // 16 lanes per row: one row per half warp per tile
constexpr int TM_BLOCKS = 480;                       // blocks per tile: 34,560 B of values
constexpr int TM_STAGES = 3;
constexpr int TM_A_BYTES = 72 * TM_BLOCKS + 16;      // + alignment slack (the copy starts at the 16-byte line below the tile)
constexpr int TM_C_BYTES = 4 * TM_BLOCKS + 16;
constexpr int TM_P_BYTES = ((4 * (TM_ROWS + 1) + 15) / 16) * 16 + 16;
constexpr int TM_STAGE_BYTES = TM_A_BYTES + TM_C_BYTES + TM_P_BYTES;
constexpr int TM_SMEM = TM_STAGES * TM_STAGE_BYTES + 128;  // + barriers
constexpr unsigned TM_MAX_POLLS = 1u << 20;                // bounded waits: a protocol bug must not hang the GPU

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, unsigned parity, int *err) {
  const unsigned a = smem_u32(bar);
  for (unsigned polls = 0; polls < TM_MAX_POLLS; polls++) {
    unsigned ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  atomicExch(err, 1);
  return false;
}
// global -> shared bulk copy, completion counted in bytes on `bar`; addresses and size are multiples of 16
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// MODE 1: y = mask(A x), per-CTA sum x.y.  MODE 2: y = mask(b - A x), per-CTA sum y^2 invD.
template <int MODE>
__global__ void __launch_bounds__(TM_TB, 2) k_spmv_tma(int nTiles, const int4 *__restrict__ tiles, const int *__restrict__ bp,
                                                       const int *__restrict__ bc, const double *__restrict__ A,
                                                       const double *__restrict__ x, double *__restrict__ y,
                                                       const unsigned char *__restrict__ mask, const double *__restrict__ b,
                                                       const double *__restrict__ invD, const FbScalars *sc, double *slots, int *err) {
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + TM_STAGES * TM_STAGE_BYTES);
  unsigned long long *empty = full + TM_STAGES;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TM_STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], TM_CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int myTiles = (nTiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // tiles blockIdx.x, + gridDim.x, ...
  double part = 0.0;
  if (warp == TM_CONSUMER_WARPS) {
    // ---- producer: one lane keeps the ring of stages full ----------------------------------------------------------
    if ((threadIdx.x & 31) == 0) {
      for (int i = 0; i < myTiles; i++) {
        const int s = i % TM_STAGES;
        if (i >= TM_STAGES && !mbar_wait(&empty[s], ((i / TM_STAGES) - 1) & 1, err)) break;
        const int4 t = __ldg(tiles + blockIdx.x + (size_t)i * gridDim.x);
        unsigned char *st = smem + s * TM_STAGE_BYTES;
        const size_t offA = 72 * (size_t)t.z, offC = 4 * (size_t)t.z, offP = 4 * (size_t)t.x;
        const unsigned dA = (unsigned)(offA & 15), dC = (unsigned)(offC & 15), dP = (unsigned)(offP & 15);
        const unsigned nA = (72u * (unsigned)t.w + dA + 15u) & ~15u, nC = (4u * (unsigned)t.w + dC + 15u) & ~15u;
        const unsigned nP = (4u * (unsigned)(t.y + 1) + dP + 15u) & ~15u;
        mbar_expect_tx(&full[s], nA + nC + nP);
        bulk_g2s(st, reinterpret_cast<const unsigned char *>(A) + (offA - dA), nA, &full[s]);
        bulk_g2s(st + TM_A_BYTES, reinterpret_cast<const unsigned char *>(bc) + (offC - dC), nC, &full[s]);
        bulk_g2s(st + TM_A_BYTES + TM_C_BYTES, reinterpret_cast<const unsigned char *>(bp) + (offP - dP), nP, &full[s]);
      }
    }
  } else {
    // ---- consumers: half warp <-> block row of the tile --------------------------------------------------------------
    const int lane = threadIdx.x & (TILE_G - 1);
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);
    const int g = threadIdx.x / TILE_G;  // 0 .. TM_ROWS-1
    int4 tn = make_int4(0, 0, 0, 0);
    if (myTiles > 0) tn = __ldg(tiles + blockIdx.x);
    for (int i = 0; i < myTiles; i++) {
      const int s = i % TM_STAGES;
      const int4 t = tn;  // the tile record was requested one tile ago: no global-load latency at the head of a tile
      if (i + 1 < myTiles) tn = __ldg(tiles + blockIdx.x + (size_t)(i + 1) * gridDim.x);
      const bool mine = g < t.y;
      const int v = t.x + g;
      // the row owners' vector entries do not depend on the tile's bytes: requested before the wait
      const size_t row = 3 * (size_t)v + (lane < 3 ? lane : 0);
      double xr = 0.0, br = 0.0, wr = 0.0;
      unsigned char mk = 0;
      if (mine && lane < 3) {
        mk = __ldg(mask + row);
        if (MODE == 1) xr = __ldg(x + row);
        if (MODE == 2) { br = __ldg(b + row); wr = __ldg(invD + row); }
      }
      if (!mbar_wait(&full[s], (i / TM_STAGES) & 1, err)) break;
      if (mine) {
        const unsigned char *st = smem + s * TM_STAGE_BYTES;
        const unsigned dA = (unsigned)((72 * (size_t)t.z) & 15), dC = (unsigned)((4 * (size_t)t.z) & 15), dP = (unsigned)((4 * (size_t)t.x) & 15);
        const int *sP = reinterpret_cast<const int *>(st + TM_A_BYTES + TM_C_BYTES + dP);
        const int rs = sP[g] - t.z, n3 = 3 * (sP[g + 1] - sP[g]);  // first block of the row inside the tile
        const double *sA = reinterpret_cast<const double *>(st + dA) + 9 * (size_t)rs;
        const int *sC = reinterpret_cast<const int *>(st + TM_A_BYTES + dC) + rs;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
        for (int base = 0; base < n3; base += TILE_CHUNK) {
          double val[3][3];
          int col[3];
#pragma unroll
          for (int p = 0; p < 3; p++) {
            const int tt = base + lane + TILE_G * p;
            const bool ok = tt < n3;
            col[p] = ok ? sC[tt / 3] : -1;
            val[p][0] = ok ? sA[tt] : 0.0;
            val[p][1] = ok ? sA[tt + n3] : 0.0;
            val[p][2] = ok ? sA[tt + 2 * n3] : 0.0;
          }
#pragma unroll
          for (int p = 0; p < 3; p++) {
            const int tt = base + lane + TILE_G * p;
            const double xv = (col[p] >= 0) ? __ldg(x + 3 * (size_t)col[p] + (tt % 3)) : 0.0;
            acc0 = fma(val[p][0], xv, acc0); acc1 = fma(val[p][1], xv, acc1); acc2 = fma(val[p][2], xv, acc2);
          }
        }
#pragma unroll
        for (int o = TILE_G / 2; o > 0; o >>= 1) {
          acc0 += __shfl_xor_sync(gmask, acc0, o, TILE_G);
          acc1 += __shfl_xor_sync(gmask, acc1, o, TILE_G);
          acc2 += __shfl_xor_sync(gmask, acc2, o, TILE_G);
        }
        if (lane < 3) {
          double sres = (lane == 0) ? acc0 : ((lane == 1) ? acc1 : acc2);
          if (MODE == 1) {
            if (mk) sres = 0.0;
            y[row] = sres;
            part = fma(xr, sres, part);
          } else {
            const double rres = mk ? 0.0 : (br - sres);
            y[row] = rres;
            part += (rres * rres) * wr;
          }
        }
      }
      __syncwarp();  // both half warps are done reading the stage
      if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
    }
  }
  block_reduce_to_slot<TM_TB>(part, slots);
}

}  // namespace

void fb_tma_release(fb_context *c) {
  FbTma *t = c->tma;
  if (!t) return;
  if (t->tiles) fb_dev_free(t->tiles);
  if (t->err) fb_dev_free(t->err);
  delete t;
  c->tma = nullptr;
}

// Cuts the block rows into tiles (host, from a copy of bp).  FB_OK with c->tma == nullptr: not applicable (a block row with
// more than TM_BLOCKS blocks, or the kernel does not fit), the caller keeps the default kernels.
int fb_tma_plan(fb_context *c) {
  if (c->tma || c->nV == 0 || c->nB == 0) return FB_OK;
  cudaStream_t st = c->stream;
  std::vector<int> bp((size_t)c->nV + 1);
  FB_CUDA(cudaMemcpyAsync(bp.data(), c->bp, sizeof(int) * bp.size(), cudaMemcpyDeviceToHost, st));
  FB_CUDA(cudaStreamSynchronize(st));
  std::vector<int4> tiles;
  const int lo = c->row_lo, hi = c->row_hi;
  for (int v = lo; v < hi;) {
    int e = v;
    while (e < hi && e - v < TM_ROWS && bp[e + 1] - bp[v] <= TM_BLOCKS) e++;
    if (e == v) return FB_OK;  // a single row does not fit a tile
    tiles.push_back(make_int4(v, e - v, bp[v], bp[e] - bp[v]));
    v = e;
  }
  if (tiles.empty()) return FB_OK;
  if (cudaFuncSetAttribute(k_spmv_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM) != cudaSuccess ||
      cudaFuncSetAttribute(k_spmv_tma<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM) != cudaSuccess) {
    cudaGetLastError();
    return FB_OK;
  }
  int perSM = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_spmv_tma<1>, TM_TB, TM_SMEM) != cudaSuccess || perSM < 1) { cudaGetLastError(); return FB_OK; }
  FbTma *t = new FbTma();
  t->nTiles = (int)tiles.size();
  t->tiles = nullptr;
  t->err = nullptr;
  c->tma = t;
  int status = fb_dev_alloc(c, &t->tiles, tiles.size());
  if (status == FB_OK) status = fb_dev_alloc(c, &t->err, 1);
  if (status != FB_OK) { fb_tma_release(c); return status; }
  FB_CUDA(cudaMemcpyAsync(t->tiles, tiles.data(), sizeof(int4) * tiles.size(), cudaMemcpyHostToDevice, st));
  FB_CUDA(cudaMemsetAsync(t->err, 0, sizeof(int), st));
  FB_CUDA(cudaStreamSynchronize(st));
  t->grid = std::max(1, std::min(std::min(perSM * c->sm_count, t->nTiles), FB_MAX_PARTIALS));
  return FB_OK;
}

int fb_tma_grid(const fb_context *c) { return c->tma ? c->tma->grid : 0; }

// 1 when a bounded mbarrier wait gave up during the last products (results are then invalid)
int fb_tma_failed(fb_context *c) {
  if (!c->tma) return 0;
  int h = 0;
  if (cudaMemcpyAsync(&h, c->tma->err, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess) return 1;
  if (cudaStreamSynchronize(c->stream) != cudaSuccess) return 1;
  return h;
}

void fb_tma_launch(fb_context *c, int mode, const double *x, double *y, const double *b, double *slots) {
  FbTma *t = c->tma;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)t->grid);
  cfg.blockDim = dim3((unsigned)TM_TB);
  cfg.dynamicSmemBytes = TM_SMEM;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = c->pdl ? 1 : 0;
  const FbScalars *sc = c->sc;
  if (mode == 1)
    cudaLaunchKernelEx(&cfg, k_spmv_tma<1>, t->nTiles, (const int4 *)t->tiles, (const int *)c->bp, (const int *)c->bc, (const double *)c->Keff, x, y,
                       (const unsigned char *)c->rowmask, b, (const double *)c->invD, sc, slots, t->err);
  else
    cudaLaunchKernelEx(&cfg, k_spmv_tma<2>, t->nTiles, (const int4 *)t->tiles, (const int *)c->bp, (const int *)c->bc, (const double *)c->Keff, x, y,
                       (const unsigned char *)c->rowmask, b, (const double *)c->invD, sc, slots, t->err);
  c->launches++;
}
