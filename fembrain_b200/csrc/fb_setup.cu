// fb_setup.cu — integer/index construction on the device, bit-exact with the reference.
//
// Reference (src/3rdparty/vegafem): CorotationalLinearFEM::GetStiffnessMatrixTopology
// (corotationalLinearFEM/corotationalLinearFEM.cpp:163-186) inserts, for every tet and every vertex
// pair (v_i, v_j), a 3x3 block into std::map rows (SparseMatrixOutline::AddEntry,
// sparseMatrix/sparseMatrix.cpp:128-138) => rows hold ascending, unique columns; BuildRowColumnIndices
// (corotationalLinearFEM.cpp:482-502) caches the position of v_j inside the row of v_i by a linear
// search (GetInverseIndex, sparseMatrix.cpp:613-620).  The reference spends 18.8 s per million tets
// here.
//
// Here: the 16 nT (v_i, v_j) pairs are radix-sorted as 64-bit keys (stable, so contributions to one
// block stay in ascending element order = the reference's accumulation order), run heads give the
// unique blocks, a scan gives block ids; row pointers come from an integer histogram + scan.
#include <cub/cub.cuh>

#include "fb_internal.h"

namespace {

__global__ void k_make_pairs(int nT, int nV, const int *__restrict__ tets, unsigned long long *__restrict__ keys,
                             unsigned int *__restrict__ vals, int *__restrict__ err) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)nT * 16) return;
  int el = (int)(t >> 4), ij = (int)(t & 15), i = ij >> 2, j = ij & 3;
  int vi = tets[4 * (size_t)el + i], vj = tets[4 * (size_t)el + j];
  if ((unsigned)vi >= (unsigned)nV || (unsigned)vj >= (unsigned)nV) {
    atomicExch(err, el + 1);
    vi = vj = 0;
  }
  if (i != j && vi == vj) atomicExch(err + 1, el + 1);  // repeated vertex inside one tet
  keys[t] = ((unsigned long long)(unsigned)vi << 32) | (unsigned)vj;
  vals[t] = (unsigned int)t;
}

__global__ void k_heads(size_t n, const unsigned long long *__restrict__ keys, int *__restrict__ head) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  head[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1 : 0;
}

// blk[t] = inclusive scan of head = 1-based block id of sorted contribution t
__global__ void k_fill_blocks(size_t n, const unsigned long long *__restrict__ keys, const int *__restrict__ blk,
                              int *__restrict__ bc, int *__restrict__ brow, int *__restrict__ seg,
                              int *__restrict__ diag, int *__restrict__ rowCount) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  bool isHead = (t == 0) || (keys[t] != keys[t - 1]);
  if (!isHead) return;
  int b = blk[t] - 1;
  int row = (int)(keys[t] >> 32), col = (int)(keys[t] & 0xffffffffu);
  bc[b] = col;
  brow[b] = row;
  seg[b] = (int)t;
  if (row == col) diag[row] = b;
  atomicAdd(&rowCount[row], 1);  // integer histogram: result independent of order
}

__global__ void k_fill_contrib(size_t n, const unsigned long long *__restrict__ keys, const unsigned int *__restrict__ vals,
                               const int *__restrict__ blk, const int *__restrict__ bp, unsigned int *__restrict__ src,
                               int *__restrict__ colIdx) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  unsigned int v = vals[t];
  src[t] = v;
  int row = (int)(keys[t] >> 32);
  colIdx[v] = (blk[t] - 1) - bp[row];
}

__global__ void k_set_int(int *p, int v) { *p = v; }


}  // namespace

int fb_build_topology(fb_context *c) {
  const size_t n = (size_t)c->nT * 16;
  cudaStream_t st = c->stream;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  unsigned int *vals = nullptr, *vals2 = nullptr;
  int *blk = nullptr, *err = nullptr, *rowCount = nullptr;
  void *tmp = nullptr;
  int status = FB_OK;
  auto cleanup = [&]() {
    for (void *q : {(void *)keys, (void *)keys2, (void *)vals, (void *)vals2, (void *)blk, (void *)err, (void *)rowCount, tmp}) fb_tmp_free(st, q);
  };
#define SETUP_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      fb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      cleanup();                                                                           \
      return (e__ == cudaErrorMemoryAllocation) ? FB_ERR_OUT_OF_MEMORY : FB_ERR_CUDA;      \
    }                                                                                      \
  } while (0)

  SETUP_CUDA(fb_tmp_alloc(st, &keys, sizeof(unsigned long long) * (n ? n : 1)));
  SETUP_CUDA(fb_tmp_alloc(st, &keys2, sizeof(unsigned long long) * (n ? n : 1)));
  SETUP_CUDA(fb_tmp_alloc(st, &vals, sizeof(unsigned int) * (n ? n : 1)));
  SETUP_CUDA(fb_tmp_alloc(st, &vals2, sizeof(unsigned int) * (n ? n : 1)));
  SETUP_CUDA(fb_tmp_alloc(st, &blk, sizeof(int) * (n ? n : 1)));
  SETUP_CUDA(fb_tmp_alloc(st, &err, sizeof(int) * 2));
  SETUP_CUDA(fb_tmp_alloc(st, &rowCount, sizeof(int) * ((size_t)c->nV + 1)));
  SETUP_CUDA(cudaMemsetAsync(err, 0, sizeof(int) * 2, st));
  SETUP_CUDA(cudaMemsetAsync(rowCount, 0, sizeof(int) * ((size_t)c->nV + 1), st));

  const int TB = 256;
  const unsigned gridN = (unsigned)((n + TB - 1) / TB);
  if (n) {
    k_make_pairs<<<gridN, TB, 0, st>>>(c->nT, c->nV, c->tets, keys, vals, err);
    c->launches++;
  }
  int herr[2] = {0, 0};
  SETUP_CUDA(cudaMemcpyAsync(herr, err, sizeof(herr), cudaMemcpyDeviceToHost, st));
  SETUP_CUDA(cudaStreamSynchronize(st));
  if (herr[0]) {
    fb_set_error("tetrahedron %d references a vertex outside [0, %d)", herr[0] - 1, c->nV);
    cleanup();
    return FB_ERR_BAD_MESH;
  }
  // (a tet with a repeated vertex is degenerate but the reference accepts it and produces NaNs —
  //  blobtree/tumor.veg element 9767 — so it is not rejected here either: herr[1] is informational)

  int vbits = 1;
  while ((1ll << vbits) < (long long)c->nV) vbits++;
  size_t tmpBytes = 0, tb2 = 0;
  cub::DoubleBuffer<unsigned long long> dk(keys, keys2);
  cub::DoubleBuffer<unsigned int> dv(vals, vals2);
  SETUP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, dk, dv, (int64_t)n, 0, 32 + vbits, st));
  SETUP_CUDA(cub::DeviceScan::InclusiveSum(nullptr, tb2, blk, blk, (int64_t)(n > (size_t)c->nV + 1 ? n : (size_t)c->nV + 1), st));
  if (tb2 > tmpBytes) tmpBytes = tb2;
  SETUP_CUDA(fb_tmp_alloc(st, &tmp, tmpBytes ? tmpBytes : 1));
  if (n) {
    SETUP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, dk, dv, (int64_t)n, 0, 32 + vbits, st));
    c->launches += 4;
  }
  unsigned long long *sk = dk.Current();
  unsigned int *sv = dv.Current();
  int *head = (int *)(sk == keys ? keys2 : keys);  // reuse the idle key buffer for the head flags
  int nB = 0;
  if (n) {
    k_heads<<<gridN, TB, 0, st>>>(n, sk, head);
    SETUP_CUDA(cub::DeviceScan::InclusiveSum(tmp, tmpBytes, head, blk, (int64_t)n, st));
    c->launches += 2;
    SETUP_CUDA(cudaMemcpyAsync(&nB, blk + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    SETUP_CUDA(cudaStreamSynchronize(st));
  }
  c->nB = nB;
  c->nnzK = 9ll * nB;

  status = fb_dev_alloc(c, &c->bp, (size_t)c->nV + 1 + 8);  // + slack for 16-byte-line bulk copies (fb_tma.cu)
  if (!status) status = fb_dev_alloc(c, &c->bc, (size_t)nB + 8);
  if (!status) status = fb_dev_alloc(c, &c->brow, (size_t)nB);
  if (!status) status = fb_dev_alloc(c, &c->diag, (size_t)c->nV);
  if (!status) status = fb_dev_alloc(c, &c->seg, (size_t)nB + 1);
  if (!status) status = fb_dev_alloc(c, &c->src, n);
  if (!status) status = fb_dev_alloc(c, &c->colIdx, n);
  if (status) { cleanup(); return status; }
  SETUP_CUDA(cudaMemsetAsync(c->diag, 0xff, sizeof(int) * (size_t)c->nV, st));  // -1 = vertex in no tet
  if (n) {
    k_fill_blocks<<<gridN, TB, 0, st>>>(n, sk, blk, c->bc, c->brow, c->seg, c->diag, rowCount);
    c->launches++;
  }
  k_set_int<<<1, 1, 0, st>>>(c->seg + nB, (int)n);
  SETUP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, rowCount, c->bp, (int64_t)c->nV + 1, st));
  c->launches += 2;
  if (n) {
    k_fill_contrib<<<gridN, TB, 0, st>>>(n, sk, sv, blk, c->bp, c->src, c->colIdx);
    c->launches++;
  }
  SETUP_CUDA(cudaStreamSynchronize(st));
  SETUP_CUDA(cudaGetLastError());
  cleanup();
#undef SETUP_CUDA
  return FB_OK;
}
