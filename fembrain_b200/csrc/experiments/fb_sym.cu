// fb_sym.cu — the PCG products q = Keff d and r = b - Keff x from HALF the matrix bytes.
//
// Reference: CGSolver::SolveLinearSystemWithJacobiPreconditioner (src/3rdparty/vegafem/sparseSolver/CGSolver.cpp:129-190)
// multiplies with the full jagged-row matrix (SparseMatrix::MultiplyVector, sparseMatrix.cpp:405-413), 12 B per entry.
// Keff = M + h D + h^2 K is symmetric (each element contributes R K0 R^T; M and D = dampK K + dampM M likewise); as
// assembled in floating point the two halves differ by rounding only (measured on the reference's own matrix:
// max |A - A^T| = 2.6e-16 max |A|).  The SpMV is HBM bound, so this path stores the block-upper triangle only:
//   * U: for block row i the diagonal block and the blocks with column > i, in the layout of Keff (three planes of
//     3 nb doubles per block row), re-packed from Keff once per solve (k_sym_pack: reads half of Keff, ~35 us at 1M tets);
//   * row j streams its upper blocks from DRAM exactly like k_spmv_rows3 and takes its LOWER blocks (j, i < j) as the
//     transposes of U's blocks (i, j), which row i streamed a few thousand rows earlier: they come from L2 (126 MB; the
//     look-back distance on the 1M / 10M-tet cubes is 5 / 23 MB of U).
// DRAM bytes per product: 8.44 B per upper entry + 16 B per lower block (column, offset, plane stride) + 52 B per block
// row = 0.60 of the full-matrix kernel.  Values used for the lower half are U's, i.e. the product is with a matrix that
// differs from Keff by its own rounding asymmetry (<= 3e-16 relative to max |A|) — the same order as the changed summation
// order of any parallel SpMV; K, Keff, rhs as exported stay the reference's bits, PCG recurrences, refresh and stopping
// rule are unchanged (fb_pcg.cu).  Single-GPU, single-mesh contexts, three-kernel schedule only.
//
// MEASURED (B200, profiles/r01_spmv_sym_variants.txt, r01_spmv_sym_ncu_details.txt): correct (tests/test_sym_gpu.py) and the
// DRAM traffic does drop (150 MB instead of 200 MB per product at 1M tets, L2 hit rate 55 %), but the product is SLOWER: 48.6
// us in step against 38.5 us at 1M tets, 412 against 366 us at 10M.  The kernel is not HBM bound any more (40 % of peak) but
// bound by the L1TEX data pipe (l1tex__data_pipe_lsu_wavefronts 76 %, LSU write-back 70 % busy): every stored block still
// passes the load-store unit twice (once streamed, once gathered, the gather as nine 8-byte loads whose 32 lanes hit 32
// different lines), so the bytes through LSU per row are those of the full-matrix kernel (~1.6 kB) while its DRAM stream
// gets cheaper.  OPT-IN therefore (FEMBRAIN_B200_SPMV=sym).  What would have to change to profit from the halved DRAM
// traffic is written down in DESIGN.md §4 (each block through the LSU once: scatter contributions of the transposed
// block to a per-row list and sum them in fixed order, or stage the stream with TMA into shared memory).
#include <cub/cub.cuh>

#include "../fb_internal.h"
#include "../fb_pcg_common.cuh"

struct FbSym {
  int nBu, nBl;    // upper (incl. diagonal) and lower blocks
  int *ubp;        // [nV+1] upper block-row pointer
  int *ubc;        // [nBu]  upper block columns
  int *lbp;        // [nV+1] lower block-row pointer
  int4 *lmeta;     // [nBl]  {column i, offset of U(i,j)[0][0] in doubles, plane stride of row i in doubles, 0}
  double *U;       // [9 nBu]
  int grid[3];     // one resident wave per mode (1, 2 used)
  int gridPack;
};

namespace {

constexpr int SY_TB = 256;
constexpr int SY_G = 8;            // lanes per block row: the upper half of a row has ~8 blocks = 24 scalar columns
constexpr int SY_CHUNK = 3 * SY_G;

__global__ void k_sym_count(int nV, const int *__restrict__ bp, const int *__restrict__ diag, int *__restrict__ nu, int *__restrict__ nl) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > nV) return;
  int u = 0, l = 0;
  if (v < nV && diag[v] >= 0) { u = bp[v + 1] - diag[v]; l = diag[v] - bp[v]; }
  nu[v] = u;
  nl[v] = l;
}

// one thread per block of K: upper blocks copy their column, lower blocks look their mirror image up in row `col`
__global__ void k_sym_fill(int nB, const int *__restrict__ bp, const int *__restrict__ bc, const int *__restrict__ brow,
                           const int *__restrict__ diag, const int *__restrict__ ubp, const int *__restrict__ lbp,
                           int *__restrict__ ubc, int4 *__restrict__ lmeta, int *__restrict__ bad) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nB) return;
  const int v = brow[g], col = bc[g];
  if (g >= diag[v]) {
    ubc[ubp[v] + (g - diag[v])] = col;
    return;
  }
  // (v, col) with col < v: position of column v among the upper blocks of row col (ascending columns)
  int lo = diag[col], hi = bp[col + 1] - 1, pos = -1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int cm = bc[mid];
    if (cm == v) { pos = mid; break; }
    if (cm < v) lo = mid + 1; else hi = mid - 1;
  }
  if (pos < 0) { atomicExch(bad, 1); return; }  // structurally unsymmetric: cannot happen for an element-built pattern
  const int nu = bp[col + 1] - diag[col];
  lmeta[lbp[v] + (g - bp[v])] = make_int4(col, 9 * ubp[col] + 3 * (pos - diag[col]), 3 * nu, 0);
}

// U <- upper part of Keff: per block row and plane k, the tail of the plane from the diagonal block on
__global__ void __launch_bounds__(SY_TB) k_sym_pack(int nV, const int *__restrict__ bp, const int *__restrict__ diag,
                                                    const int *__restrict__ ubp, const double *__restrict__ A, double *__restrict__ U) {
  const int lane = threadIdx.x & (SY_G - 1);
  const int nGroups = gridDim.x * (SY_TB / SY_G);
  for (int v = blockIdx.x * (SY_TB / SY_G) + threadIdx.x / SY_G; v < nV; v += nGroups) {
    const int dg = __ldg(diag + v);
    if (dg < 0) continue;
    const int rs = __ldg(bp + v), re = __ldg(bp + v + 1), us = __ldg(ubp + v);
    const int n3 = 3 * (re - rs), n3u = 3 * (re - dg), skip = 3 * (dg - rs);
    const double *src = A + 9 * (size_t)rs + skip;
    double *dst = U + 9 * (size_t)us;
#pragma unroll
    for (int k = 0; k < 3; k++)
      for (int t = lane; t < n3u; t += SY_G) dst[(size_t)k * n3u + t] = ld_stream(src + (size_t)k * n3 + t);
  }
}

// MODE 1: y = mask(A x), per-CTA sum x.y.  MODE 2: y = mask(b - A x), per-CTA sum y^2 invD.  (modes of k_spmv_rows3)
template <int MODE>
__global__ void __launch_bounds__(SY_TB, 4) k_spmv_sym(int rowBeg, int nV, const int *__restrict__ ubp, const int *__restrict__ ubc,
                                                       const double *__restrict__ U, const int *__restrict__ lbp,
                                                       const int4 *__restrict__ lmeta, const double *__restrict__ x,
                                                       double *__restrict__ y, const unsigned char *__restrict__ mask,
                                                       const double *__restrict__ b, const double *__restrict__ invD,
                                                       const FbScalars *sc, double *slots) {
  pdl_wait();
  pdl_trigger();
  if (sc->done) return;
  const int lane = threadIdx.x & (SY_G - 1);
  const unsigned gmask = 0xffu << (threadIdx.x & 24);
  const int nGroups = gridDim.x * (SY_TB / SY_G);
  double part = 0.0;
  int v = rowBeg + blockIdx.x * (SY_TB / SY_G) + threadIdx.x / SY_G;
  int us = 0, ue = 0, ls = 0, le = 0;
  if (v < nV) { us = __ldg(ubp + v); ue = __ldg(ubp + v + 1); ls = __ldg(lbp + v); le = __ldg(lbp + v + 1); }
  while (v < nV) {
    const int vn = v + nGroups;
    int usn = 0, uen = 0, lsn = 0, len = 0;
    if (vn < nV) { usn = __ldg(ubp + vn); uen = __ldg(ubp + vn + 1); lsn = __ldg(lbp + vn); len = __ldg(lbp + vn + 1); }
    // the first lower blocks' records are requested now; they are back when the upper half has been issued
    int4 lm = make_int4(-1, 0, 0, 0);
    if (ls + lane < le) lm = __ldg(lmeta + ls + lane);
    const int n3 = 3 * (ue - us);
    const size_t row = 3 * (size_t)v + (lane < 3 ? lane : 0);
    double xr = 0.0, br = 0.0, wr = 0.0;
    unsigned char mk = 0;
    if (lane < 3) {
      mk = __ldg(mask + row);
      if (MODE == 1) xr = __ldg(x + row);
      if (MODE == 2) { br = __ldg(b + row); wr = __ldg(invD + row); }
    }
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    // upper half (diagonal block included): streamed, three planes, lane <-> scalar column
    const double *a0 = U + 9 * (size_t)us;
    for (int base = 0; base < n3; base += SY_CHUNK) {
      double val[3][3];
      int col[3];
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int t = base + lane + SY_G * p;
        col[p] = (t < n3) ? __ldg(ubc + us + t / 3) : -1;
      }
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int t = base + lane + SY_G * p;
        const bool ok = t < n3;
        const double *q = a0 + t;
        val[p][0] = ok ? ld_stream(q) : 0.0;
        val[p][1] = ok ? ld_stream(q + n3) : 0.0;
        val[p][2] = ok ? ld_stream(q + 2 * (size_t)n3) : 0.0;
      }
#pragma unroll
      for (int p = 0; p < 3; p++) {
        const int t = base + lane + SY_G * p;
        const double xv = (col[p] >= 0) ? __ldg(x + 3 * (size_t)col[p] + (t % 3)) : 0.0;
        acc0 = fma(val[p][0], xv, acc0); acc1 = fma(val[p][1], xv, acc1); acc2 = fma(val[p][2], xv, acc2);
      }
    }
    // lower half: lane <-> block (j, i), the transpose of U's block (i, j); its three 24-byte pieces are L2 (or L1) hits
    for (int pb = ls; pb < le; pb += SY_G) {
      if (pb != ls) {
        lm = make_int4(-1, 0, 0, 0);
        if (pb + lane < le) lm = __ldg(lmeta + pb + lane);
      }
      if (lm.x >= 0) {
        const double *s = U + lm.y;
        const double *xi = x + 3 * (size_t)lm.x;
        double w[3][3];
#pragma unroll
        for (int m = 0; m < 3; m++) {
#pragma unroll
          for (int k = 0; k < 3; k++) w[m][k] = __ldg(s + (size_t)m * lm.z + k);  // U(i,j)[m][k] = A(j,i)[k][m]
        }
        const double x0 = __ldg(xi), x1 = __ldg(xi + 1), x2 = __ldg(xi + 2);
        acc0 = fma(w[0][0], x0, acc0); acc1 = fma(w[0][1], x0, acc1); acc2 = fma(w[0][2], x0, acc2);
        acc0 = fma(w[1][0], x1, acc0); acc1 = fma(w[1][1], x1, acc1); acc2 = fma(w[1][2], x1, acc2);
        acc0 = fma(w[2][0], x2, acc0); acc1 = fma(w[2][1], x2, acc1); acc2 = fma(w[2][2], x2, acc2);
      }
    }
#pragma unroll
    for (int o = SY_G / 2; o > 0; o >>= 1) {
      acc0 += __shfl_xor_sync(gmask, acc0, o, SY_G);
      acc1 += __shfl_xor_sync(gmask, acc1, o, SY_G);
      acc2 += __shfl_xor_sync(gmask, acc2, o, SY_G);
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
    v = vn; us = usn; ue = uen; ls = lsn; le = len;
  }
  block_reduce_to_slot<SY_TB>(part, slots);
}

template <typename K>
int wave_grid(const fb_context *c, K kernel, size_t want) {
  int perSM = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kernel, SY_TB, 0) != cudaSuccess || perSM < 1) { cudaGetLastError(); perSM = 1; }
  const size_t wave = (size_t)perSM * (size_t)c->sm_count;
  size_t g = want < wave ? want : wave;
  if (g > FB_MAX_PARTIALS) g = FB_MAX_PARTIALS;
  return g < 1 ? 1 : (int)g;
}

}  // namespace

void fb_sym_release(fb_context *c) {
  FbSym *s = c->sym;
  if (!s) return;
  void *ptrs[] = {s->ubp, s->ubc, s->lbp, s->lmeta, s->U};
  for (void *p : ptrs)
    if (p) fb_dev_free(p);
  delete s;
  c->sym = nullptr;
}

// Builds the upper/lower structure (once per context, on its stream).  Returns FB_OK with c->sym == nullptr when the
// path does not apply (empty mesh, offsets beyond 32 bits): the caller then keeps the full-matrix kernels.
int fb_sym_plan(fb_context *c) {
  if (c->sym || c->nV == 0 || c->nB == 0) return FB_OK;
  if (9LL * (long long)c->nB >= 0x7fffffffLL) return FB_OK;  // offsets into U are ints
  cudaStream_t st = c->stream;
  FbSym *s = new FbSym();
  memset(s, 0, sizeof(*s));
  c->sym = s;
  const size_t n1 = (size_t)c->nV + 1;
  int *nu = nullptr, *nl = nullptr, *bad = nullptr;
  void *tmp = nullptr;
  int status = FB_OK;
#define SYM_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess && status == FB_OK) { fb_set_error("%s -> %s", #call, cudaGetErrorString(e__)); status = FB_ERR_CUDA; } } while (0)
  SYM_CUDA(fb_tmp_alloc(st, &nu, sizeof(int) * n1));
  SYM_CUDA(fb_tmp_alloc(st, &nl, sizeof(int) * n1));
  SYM_CUDA(fb_tmp_alloc(st, &bad, sizeof(int)));
  if (status == FB_OK) status = fb_dev_alloc(c, &s->ubp, n1);
  if (status == FB_OK) status = fb_dev_alloc(c, &s->lbp, n1);
  if (status == FB_OK) {
    SYM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    k_sym_count<<<(unsigned)((n1 + 255) / 256), 256, 0, st>>>(c->nV, c->bp, c->diag, nu, nl);
    size_t tb = 0;
    SYM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, nu, s->ubp, (int64_t)n1, st));
    SYM_CUDA(fb_tmp_alloc(st, (char **)&tmp, tb));
    SYM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, nu, s->ubp, (int64_t)n1, st));
    SYM_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, nl, s->lbp, (int64_t)n1, st));
    int tot[2] = {0, 0};
    SYM_CUDA(cudaMemcpyAsync(&tot[0], s->ubp + c->nV, sizeof(int), cudaMemcpyDeviceToHost, st));
    SYM_CUDA(cudaMemcpyAsync(&tot[1], s->lbp + c->nV, sizeof(int), cudaMemcpyDeviceToHost, st));
    SYM_CUDA(cudaStreamSynchronize(st));
    s->nBu = tot[0];
    s->nBl = tot[1];
  }
  if (status == FB_OK && s->nBu + s->nBl != c->nB) { fb_set_error("symmetric SpMV plan: %d + %d blocks, expected %d", s->nBu, s->nBl, c->nB); status = FB_ERR_CUDA; }
  if (status == FB_OK) status = fb_dev_alloc(c, &s->ubc, (size_t)s->nBu);
  if (status == FB_OK) status = fb_dev_alloc(c, &s->lmeta, (size_t)s->nBl);
  if (status == FB_OK) status = fb_dev_alloc(c, &s->U, 9 * (size_t)s->nBu);
  int badHost = 0;
  if (status == FB_OK) {
    k_sym_fill<<<(unsigned)(((size_t)c->nB + 255) / 256), 256, 0, st>>>(c->nB, c->bp, c->bc, c->brow, c->diag, s->ubp, s->lbp, s->ubc, s->lmeta, bad);
    SYM_CUDA(cudaMemcpyAsync(&badHost, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    SYM_CUDA(cudaStreamSynchronize(st));
    SYM_CUDA(cudaGetLastError());
  }
#undef SYM_CUDA
  fb_tmp_free(st, tmp);
  fb_tmp_free(st, nu);
  fb_tmp_free(st, nl);
  fb_tmp_free(st, bad);
  if (status != FB_OK) { fb_sym_release(c); return status; }
  if (badHost) { fb_sym_release(c); return FB_OK; }  // pattern not symmetric: keep the full-matrix kernels
  const size_t gpb = SY_TB / SY_G;
  const size_t want = ((size_t)c->nV + gpb - 1) / gpb;
  s->grid[0] = 0;
  s->grid[1] = wave_grid(c, k_spmv_sym<1>, want);
  s->grid[2] = wave_grid(c, k_spmv_sym<2>, want);
  s->gridPack = wave_grid(c, k_sym_pack, want);
  return FB_OK;
}

int fb_sym_grid(const fb_context *c, int mode) { return c->sym ? c->sym->grid[mode] : 0; }

size_t fb_sym_bytes_per_product(const fb_context *c) {
  if (!c->sym) return 0;
  const FbSym *s = c->sym;
  return (size_t)s->nBu * 76 + (size_t)s->nBl * 16 + (size_t)c->nV * 60;  // values + column; record; two pointers, x read, y write, mask etc.
}

// U <- upper(Keff), on the context's stream (start of every solve: Keff changes with every assembly)
int fb_sym_pack(fb_context *c) {
  FbSym *s = c->sym;
  k_sym_pack<<<s->gridPack, SY_TB, 0, c->stream>>>(c->nV, c->bp, c->diag, s->ubp, c->Keff, s->U);
  c->launches++;
  return FB_OK;
}

// the solver's product over block rows [row_lo, row_hi); per-CTA sums to slots[blockIdx.x] (the consumer adds them)
void fb_sym_launch(fb_context *c, int mode, const double *x, double *y, const double *b, double *slots) {
  FbSym *s = c->sym;
  if (mode == 1)
    fb_launch(c->pdl != 0, c->stream, k_spmv_sym<1>, s->grid[1], SY_TB, c->row_lo, c->row_hi, s->ubp, s->ubc, s->U, s->lbp, s->lmeta, x, y,
              c->rowmask, b, c->invD, c->sc, slots);
  else
    fb_launch(c->pdl != 0, c->stream, k_spmv_sym<2>, s->grid[2], SY_TB, c->row_lo, c->row_hi, s->ubp, s->ubc, s->U, s->lbp, s->lmeta, x, y,
              c->rowmask, b, c->invD, c->sc, slots);
  c->launches++;
}
