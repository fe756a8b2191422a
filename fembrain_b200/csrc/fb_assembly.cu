// fb_assembly.cu — row-gather assembly: K, T = hK + D, Keff and f in ONE pass over the elements' data, element
// matrices never written to HBM.  COMPILED WITH -fmad=false (bit-identical to the reference, like fb_fem.cu).
//
// Reference (src/3rdparty/vegafem): CorotationalLinearFEM::ComputeForceAndStiffnessMatrixOfSubmesh
// (corotationalLinearFEM/corotationalLinearFEM.cpp:219-293 warp=1 branch, scatter :456-468) and the matrix
// sequence of DoTimestep (src/deformable/PS_VolumeConservingIntegrator.cpp:84-116).
//
// The two-phase path of fb_fem.cu writes every 12x12 element matrix to a scratch array (1152 B/tet) and reads it
// back: ~2.9 kB of HBM traffic per tet against 218 B of compulsory bytes, and 1248 B/tet of device memory.  Here:
//   k_rotation         one thread per tet: F, polar decomposition, R written into the element's 192-byte record
//                      (R[9], the 4x3 of MInverse, volume, lambda, mu) — 72 B/tet, the only intermediate
//   k_pack_xu          rest position and displacement of every vertex side by side (48 B, one 16-byte-aligned record)
//   k_assemble_gather  one CTA per run of consecutive block rows (vertices).  An "incidence" is a (vertex, element,
//                      local index i) triple; the four 3x3 blocks (i, j=0..3) of that element land in the vertex's
//                      block row.  All global reads of the CTA are issued up front as 16-byte asynchronous copies
//                      into shared memory (element records by incidence, vertex records by block column).  Four
//                      adjacent lanes take the four j of one incidence: each recomputes its K0 block from the 4x3
//                      of MInverse, rotates it (R K0, R K0 R^T) and forms its three force terms; the running f_el
//                      of the reference (j = 0..3 in order) is passed lane to lane inside the quad.  The blocks
//                      overwrite the incidence's record in shared memory; then one thread per (block, scalar) of
//                      the CTA's rows adds its contribution list IN ASCENDING ELEMENT ORDER — the order in which
//                      the reference's element loop calls AddEntry — and applies the DoTimestep epilogue.
// No atomics, no float reassociation: K, f, T, Keff are bit-identical to the reference's (and to the two-phase
// path).  Blocks of a CTA are visited in order of decreasing list length (precomputed), so the lanes of a warp run
// the same number of trips: a diagonal block of the cube collects 24 contributions, an edge block 4-6.
#include <cub/cub.cuh>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "fb_element_math.h"
#include "fb_internal.h"

namespace {

constexpr int EREC = 24;  // doubles per element record: R[9], G[12], volume, lambda, mu

// CAP incidences (4 CAP slots) and at most BCAP blocks per CTA, TB threads: 4 CAP / TB compute passes
template <int CAP_, int TB_, int BCAP_, int MINB_>
struct GaCfg {
  static constexpr int CAP = CAP_, TB = TB_, BCAP = BCAP_, MINB = MINB_;
  static_assert((4 * CAP_) % TB_ == 0, "whole passes");
  static_assert(BCAP_ <= 4096 && 4 * CAP_ <= 65536, "16-bit list entries");
};

// Shared memory of one CTA:
//   vals [CAP][36]  per incidence: IN the element record (24 doubles)  ->  OUT four 3x3 blocks, slot = 4 li + j at slot*9
//   xb   [BCAP][6]  per block of the CTA's rows: x0 and u of the block's column vertex
//   fel  [CAP][3]   element force rows of the incidence
//   lists: contribution -> slot, slot -> block, list starts, block order, block -> local vertex, row starts, diagonals
template <class C>
struct GatherSmem {
  double vals[C::CAP * 36];
  double xb[C::BCAP * 6];
  double fel[C::CAP * 3];
  int segl[C::BCAP + 1];
  int bpl[C::CAP + 1];
  int dgl[C::CAP];
  unsigned short csl[C::CAP * 4];
  unsigned short sbl[C::CAP * 4];
  unsigned short bol[C::BCAP];
  unsigned short bvl[C::BCAP];
};

inline unsigned grid_for(size_t n, int tb) { return (unsigned)((n + tb - 1) / tb); }

// ---- setup -------------------------------------------------------------------------------------------------
// incidences of vertex v = the (i == j) entries of its diagonal block's contribution list (a degenerate tet that
// repeats a vertex also puts (i, j != i) entries there), already in ascending (element, i) order
__global__ void k_count_inc(int nV, const int *__restrict__ diag, const int *__restrict__ seg, const unsigned int *__restrict__ src,
                            int *__restrict__ cnt) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v > nV) return;
  int n = 0;
  if (v < nV) {
    const int b = diag[v];
    if (b >= 0)
      for (int s = seg[b]; s < seg[b + 1]; s++) {
        const unsigned ij = src[s] & 15u;
        n += ((ij >> 2) == (ij & 3u));
      }
  }
  cnt[v] = n;
}

__global__ void k_fill_inc(int nV, const int *__restrict__ diag, const int *__restrict__ seg, const unsigned int *__restrict__ src,
                           const int *__restrict__ incp, unsigned int *__restrict__ inc) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const int b = diag[v];
  if (b < 0) return;
  int o = incp[v];
  for (int s = seg[b]; s < seg[b + 1]; s++) {
    const unsigned c = src[s], ij = c & 15u;
    if ((ij >> 2) == (ij & 3u)) inc[o++] = ((c >> 4) << 2) | (ij >> 2);  // element*4 + i
  }
}

__global__ void k_vertex_cta(int nCta, const int *__restrict__ ctaV, int *__restrict__ vcta) {
  int c = blockIdx.x;
  for (int v = ctaV[c] + threadIdx.x; v < ctaV[c + 1]; v += blockDim.x) vcta[v] = c;
}

// one thread per block: CTA-local slot of every contribution (csrc, in seg/src order), CTA-local block of every slot
// (sblk, in slot order), and the sort key (cta, -length)
__global__ void k_fill_csrc(int nB, const int *__restrict__ brow, const int *__restrict__ bp, const int *__restrict__ seg,
                            const unsigned int *__restrict__ src, const int *__restrict__ incp, const unsigned int *__restrict__ inc,
                            const int *__restrict__ ctaV, const int *__restrict__ vcta, unsigned short *__restrict__ csrc,
                            unsigned short *__restrict__ sblk, unsigned long long *__restrict__ keys, int *__restrict__ ids) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nB) return;
  const int v = brow[b], cta = vcta[v], vFirst = ctaV[cta];
  const int incFirst = incp[vFirst];
  const int base = incp[v] - incFirst;
  const int bl = b - bp[vFirst];
  const int lo0 = incp[v], hi0 = incp[v + 1];
  const int s0 = seg[b], s1 = seg[b + 1];
  for (int s = s0; s < s1; s++) {
    const unsigned c = src[s];
    const unsigned want = ((c >> 4) << 2) | ((c >> 2) & 3u);  // element*4 + i
    int lo = lo0, hi = hi0;
    while (lo < hi) {  // inc[lo0..hi0) ascending
      const int mid = (lo + hi) >> 1;
      if (inc[mid] < want) lo = mid + 1; else hi = mid;
    }
    const int slot = ((base + (lo - lo0)) << 2) | (int)(c & 3u);
    csrc[s] = (unsigned short)slot;
    sblk[4 * (size_t)incFirst + slot] = (unsigned short)(bl | (int)(((c >> 2) & 3u) << 12));  // local block | i << 12
  }
  const int len = s1 - s0;
  keys[b] = ((unsigned long long)(unsigned)cta << 16) | (unsigned long long)(65535 - (len > 65535 ? 65535 : len));
  ids[b] = b;
}

// static part of the element records from the structure-of-arrays planes of fb_context::edata
__global__ void k_fill_erec(int nT, const double *__restrict__ ed, double *__restrict__ erec) {
  const int el = blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nT) return;
  double *o = erec + (size_t)el * EREC;
  for (int k = 0; k < 9; k++) o[k] = 0.0;
  for (int k = 0; k < 15; k++) o[9 + k] = ed[(size_t)k * nT + el];  // G[12], volume, lambda, mu
}

// ---- per step ------------------------------------------------------------------------------------------------
// one thread per tetrahedron: P = x0 + u, F = P MInverse, R from the polar decomposition, det < 0 -> -R
// (corotationalLinearFEM.cpp:246-268)
__global__ void __launch_bounds__(128) k_rotation(int nT, const int *__restrict__ tets, const double *__restrict__ x0,
                                                  const double *__restrict__ u, double tol, double *__restrict__ erec) {
  const int el = blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nT) return;
  const int4 vt = reinterpret_cast<const int4 *>(tets)[el];
  const int vi[4] = {vt.x, vt.y, vt.z, vt.w};
  double P[4][3];
#pragma unroll
  for (int v = 0; v < 4; v++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) P[v][cc] = x0[3 * (size_t)vi[v] + cc] + u[3 * (size_t)vi[v] + cc];
  double *rec = erec + (size_t)el * EREC;
  double G[12];
#pragma unroll
  for (int k = 0; k < 12; k++) G[k] = rec[9 + k];
  double F[9], R[9];
  fbm::deformation_gradient(P, G, F);
  const double det = fbm::polar_rotation(F, R, tol, nullptr);
  if (det < 0) {
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] *= -1.0;
  }
#pragma unroll
  for (int k = 0; k < 9; k++) rec[k] = R[k];
}

__global__ void k_pack_xu(int nV, const double *__restrict__ x0, const double *__restrict__ u, double *__restrict__ xu) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * nV) return;
  const int v = t / 3, k = t - 3 * v;
  xu[6 * (size_t)v + k] = x0[t];
  xu[6 * (size_t)v + 3 + k] = u[t];
}

struct GatherParams {
  double scale, h, dampK, dampM;
  int effective;
  const int *ctaV, *incp, *bp, *bc, *diag, *seg, *border;
  const unsigned int *inc;
  const unsigned short *csrc, *sblk;
  const double *erec, *xu, *mblk;
  const unsigned char *fixed;
  double *Kraw, *T, *Keff, *invD, *f;
};

__device__ __forceinline__ void cp_async16(double *smemDst, const double *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}

template <class C>
__global__ void __launch_bounds__(C::TB, C::MINB) k_assemble_gather(const GatherParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  GatherSmem<C> &S = *reinterpret_cast<GatherSmem<C> *>(smraw);
  constexpr int TB = C::TB;
  constexpr int PASSES = 4 * C::CAP / TB;
  const int tid = threadIdx.x;
  const int v0 = __ldg(p.ctaV + blockIdx.x), v1 = __ldg(p.ctaV + blockIdx.x + 1);
  const int inc0 = __ldg(p.incp + v0);
  const int nSlots = 4 * (__ldg(p.incp + v1) - inc0);
  const int blk0 = __ldg(p.bp + v0), nBlk = __ldg(p.bp + v1) - blk0;

  // ---- stage: every global read of the CTA is issued here; records travel as 16-byte asynchronous copies ----------
#pragma unroll
  for (int ps = 0; ps < PASSES; ps++) {
    const int slot = ps * TB + tid;
    if (slot < nSlots) {
      const unsigned ei = __ldg(p.inc + inc0 + (slot >> 2));
      const double *rec = p.erec + (size_t)(ei >> 2) * EREC + 6 * (slot & 3);  // lane j of the quad copies doubles 6j .. 6j+5
      double *dst = S.vals + (slot >> 2) * 36 + 6 * (slot & 3);
      cp_async16(dst, rec);
      cp_async16(dst + 2, rec + 2);
      cp_async16(dst + 4, rec + 4);
    }
  }
  for (int t = tid; t < nBlk; t += TB) {
    const double *rec = p.xu + 6 * (size_t)__ldg(p.bc + blk0 + t);
    double *dst = S.xb + 6 * t;
    cp_async16(dst, rec);
    cp_async16(dst + 2, rec + 2);
    cp_async16(dst + 4, rec + 4);
  }
  {
    const int segBase = __ldg(p.seg + blk0);  // the CTA's contributions are seg[blk0] .. seg[blk0 + nBlk): nSlots entries
    const unsigned short *sblk = p.sblk + 4 * (size_t)inc0;
    for (int t = tid; t < nSlots; t += TB) {
      S.csl[t] = __ldg(p.csrc + segBase + t);
      S.sbl[t] = __ldg(sblk + t);
    }
    for (int t = tid; t <= nBlk; t += TB) S.segl[t] = __ldg(p.seg + blk0 + t) - segBase;
    for (int t = tid; t < nBlk; t += TB) S.bol[t] = (unsigned short)(__ldg(p.border + blk0 + t) - blk0);
    for (int t = tid; t <= v1 - v0; t += TB) {
      const int rs = __ldg(p.bp + v0 + t) - blk0;
      S.bpl[t] = rs;
      if (t < v1 - v0) {
        S.dgl[t] = __ldg(p.diag + v0 + t) - blk0;
        const int re = __ldg(p.bp + v0 + t + 1) - blk0;
        for (int bl = rs; bl < re; bl++) S.bvl[bl] = (unsigned short)t;
      }
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // ---- compute: one slot per thread and pass, operands from shared memory ---------------------------------------
#pragma unroll 1
  for (int ps = 0; ps < PASSES; ps++) {
    if (ps * TB >= nSlots) break;
    const int slot = ps * TB + tid;
    const bool active = slot < nSlots;
    const int sl = active ? slot : (nSlots - 1);  // idle lanes of the last pass repeat the last slot (shuffles stay full-warp)
    const int li = sl >> 2, j = sl & 3;
    const int sb = S.sbl[sl];
    const int i = sb >> 12;
    double *ir = S.vals + li * 36;
    const double *xj = S.xb + 6 * (sb & 4095);
    double R[9], gi[3], gj[3], X0j[3], Pj[3];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = ir[k];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      gi[k] = ir[9 + 3 * i + k];
      gj[k] = ir[9 + 3 * j + k];
      X0j[k] = xj[k];
      Pj[k] = X0j[k] + xj[3 + k];
    }
    const double vol = ir[21], lambda = ir[22], mu = ir[23];
    double eb[9], K[9], RK[9], Kel[9];
    fbm::eb_products(gj, lambda, mu, eb);
    fbm::k0_block(gi, eb, vol, K);
    fbm::warp_block(R, K, RK, Kel);
    // fElement[3i+k] += Kel[k][l] P_j[l] - RK[k][l] x0_j[l], j = 0..3 outer, l inner, starting from 0
    // (corotationalLinearFEM.cpp:275-286): the running sums travel through the quad, lane j adds its three terms
    double d[9];
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int l = 0; l < 3; l++) d[3 * k + l] = Kel[3 * k + l] * Pj[l] - RK[3 * k + l] * X0j[l];
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
#pragma unroll
    for (int jj = 0; jj < 4; jj++) {
      if (j == jj) {
        f0 += d[0]; f0 += d[1]; f0 += d[2];
        f1 += d[3]; f1 += d[4]; f1 += d[5];
        f2 += d[6]; f2 += d[7]; f2 += d[8];
      }
      f0 = __shfl_sync(0xffffffffu, f0, jj, 4);
      f1 = __shfl_sync(0xffffffffu, f1, jj, 4);
      f2 = __shfl_sync(0xffffffffu, f2, jj, 4);
    }
    __syncwarp();  // the quad has read its record: the four blocks may now overwrite it
    if (active) {
#pragma unroll
      for (int q = 0; q < 9; q++) ir[9 * j + q] = Kel[q];
      if (j == 3) { S.fel[li * 3 + 0] = f0; S.fel[li * 3 + 1] = f1; S.fel[li * 3 + 2] = f2; }
    }
  }
  __syncthreads();

  // ---- K: one thread per (block, scalar) of the CTA's rows; SparseMatrix::ResetToZero, then AddEntry in element order
  const int nItems = 9 * nBlk;
  for (int item = tid; item < nItems; item += TB) {
    const int bi = item / 9, q = item - 9 * bi;
    const int bl = S.bol[bi];
    const int k = q / 3, l = q - 3 * k;
    const int b = blk0 + bl;
    const int vl = S.bvl[bl];
    const double m = (k == l && p.effective) ? __ldg(p.mblk + b) : 0.0;  // in flight during the sum
    const int s0 = S.segl[bl], s1 = S.segl[bl + 1];
    double acc = 0.0;
    int s = s0;
    for (; s + 4 <= s1; s += 4) {  // four shared-memory loads in flight, added in list order
      const double a0 = S.vals[(int)S.csl[s] * 9 + q], a1 = S.vals[(int)S.csl[s + 1] * 9 + q];
      const double a2 = S.vals[(int)S.csl[s + 2] * 9 + q], a3 = S.vals[(int)S.csl[s + 3] * 9 + q];
      acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; s < s1; s++) acc += S.vals[(int)S.csl[s] * 9 + q];
    const int rs = S.bpl[vl], nb = S.bpl[vl + 1] - rs;
    const size_t idx = 9 * (size_t)(blk0 + rs) + (size_t)(3 * nb) * k + 3 * (size_t)(bl - rs) + l;
    const double Kv = acc * p.scale;  // *tangentStiffnessMatrix *= internalForceScalingFactor  (:87)
    if (p.Kraw) p.Kraw[idx] = Kv;
    if (p.effective) {
      double D = Kv * p.dampK;                 // ScalarMultiply(dampingStiffnessCoef, rayleigh)   (:100)
      if (k == l) D += p.dampM * m;            // rayleigh->AddSubMatrix(dampingMassCoef, M)        (:102)
      double Tv = Kv * p.h;                    // K *= h                                            (:110)
      Tv += D;                                 // K += D                                            (:112)
      p.T[idx] = Tv;                           // (K += 1.0 * empty dampingMatrix: no entries)      (:113)
      double Ke = Tv * p.h;                    // K *= h                                            (:115)
      if (k == l) Ke += 1.0 * m;               // K->AddSubMatrix(1.0, M)                           (:116)
      p.Keff[idx] = Ke;
      if (k == l && bl == S.dgl[vl]) {
        const int dof = 3 * (v0 + vl) + k;
        p.invD[dof] = p.fixed[dof] ? 0.0 : 1.0 / Ke;  // CGSolver.cpp:134-136 on the constrained system
      }
    }
  }
  // ---- f: one thread per DOF of the CTA's rows, incident elements in ascending order (:288-293)
  const int nF = 3 * (v1 - v0);
  for (int item = tid; item < nF; item += TB) {
    const int vl = item / 3, k = item - 3 * vl;
    const int a = __ldg(p.incp + v0 + vl) - inc0, z = __ldg(p.incp + v0 + vl + 1) - inc0;
    double acc = 0.0;
    for (int t = a; t < z; t++) acc += S.fel[t * 3 + k];
    p.f[3 * (size_t)(v0 + vl) + k] = acc * p.scale;
  }
}

typedef GaCfg<256, 256, 256, 2> CfgBig;
typedef GaCfg<192, 256, 192, 2> CfgMid;
typedef GaCfg<128, 256, 128, 3> CfgSmall;

template <class C>
int launch_gather(fb_context *c, const GatherParams &p) {
  k_assemble_gather<C><<<c->ga_ctas, C::TB, sizeof(GatherSmem<C>), c->stream>>>(p);
  return FB_OK;
}

}  // namespace

// Builds the gather plan after fb_build_topology and fb_launch_element_data.  Leaves c->ga_ctas == 0 (two-phase path)
// when a vertex has more incident elements or a block row more blocks than one CTA holds, or when
// FEMBRAIN_B200_ASSEMBLY=twophase is set.
int fb_build_gather_plan(fb_context *c) {
  c->ga_ctas = 0;
  const char *env = getenv("FEMBRAIN_B200_ASSEMBLY");
  if (env && !strcmp(env, "twophase")) return FB_OK;
  if (c->nT == 0 || c->nV == 0 || c->nB == 0) return FB_OK;
  const char *capEnv = getenv("FEMBRAIN_B200_GA_CAP");
  const int capWant = capEnv ? atoi(capEnv) : 128;
  int CAP, BCAP;
  size_t smem;
  cudaError_t attr;
  if (capWant >= 256) {
    c->ga_cfg = 2; CAP = CfgBig::CAP; BCAP = CfgBig::BCAP; smem = sizeof(GatherSmem<CfgBig>);
    attr = cudaFuncSetAttribute(k_assemble_gather<CfgBig>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  } else if (capWant >= 192) {
    c->ga_cfg = 1; CAP = CfgMid::CAP; BCAP = CfgMid::BCAP; smem = sizeof(GatherSmem<CfgMid>);
    attr = cudaFuncSetAttribute(k_assemble_gather<CfgMid>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  } else {
    c->ga_cfg = 0; CAP = CfgSmall::CAP; BCAP = CfgSmall::BCAP; smem = sizeof(GatherSmem<CfgSmall>);
    attr = cudaFuncSetAttribute(k_assemble_gather<CfgSmall>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  }
  if (attr != cudaSuccess) { fb_set_error("cudaFuncSetAttribute(smem %zu) -> %s", smem, cudaGetErrorString(attr)); return FB_ERR_CUDA; }

  cudaStream_t st = c->stream;
  const int nV = c->nV, nB = c->nB;
  int *cnt = nullptr, *vcta = nullptr, *ids = nullptr, *ids2 = nullptr;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  void *tmp = nullptr;
  auto cleanup = [&]() { cudaFree(cnt); cudaFree(vcta); cudaFree(ids); cudaFree(ids2); cudaFree(keys); cudaFree(keys2); cudaFree(tmp); };
#define GP_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      fb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      cleanup();                                                                           \
      return (e__ == cudaErrorMemoryAllocation) ? FB_ERR_OUT_OF_MEMORY : FB_ERR_CUDA;      \
    }                                                                                      \
  } while (0)
#define GP_TRY(call) do { int s__ = (call); if (s__ != FB_OK) { cleanup(); return s__; } } while (0)

  GP_CUDA(cudaMalloc(&cnt, sizeof(int) * ((size_t)nV + 1)));
  GP_TRY(fb_dev_alloc(c, &c->ga_incp, (size_t)nV + 1));
  k_count_inc<<<grid_for((size_t)nV + 1, 256), 256, 0, st>>>(nV, c->diag, c->seg, c->src, cnt);
  size_t tmpBytes = 0, tb2 = 0;
  GP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cnt, c->ga_incp, (int64_t)nV + 1, st));
  cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
  cub::DoubleBuffer<int> dv(nullptr, nullptr);
  GP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, (int64_t)nB, 0, 64, st));
  if (tb2 > tmpBytes) tmpBytes = tb2;
  GP_CUDA(cudaMalloc(&tmp, tmpBytes ? tmpBytes : 1));
  GP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, cnt, c->ga_incp, (int64_t)nV + 1, st));
  c->launches += 2;
  std::vector<int> incp((size_t)nV + 1), bp((size_t)nV + 1);
  GP_CUDA(cudaMemcpyAsync(incp.data(), c->ga_incp, sizeof(int) * incp.size(), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaMemcpyAsync(bp.data(), c->bp, sizeof(int) * bp.size(), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaStreamSynchronize(st));
  const int nInc = incp[nV];
  // greedy packing of consecutive vertices: at most CAP incidences and BCAP blocks per CTA (host: one pass over nV)
  std::vector<int> ctaV;
  ctaV.reserve((size_t)nInc / (size_t)(CAP - CAP / 8) + 2);
  ctaV.push_back(0);
  int start = 0;
  for (int v = 0; v < nV; v++) {
    if (incp[v + 1] - incp[v] > CAP || bp[v + 1] - bp[v] > BCAP) { cleanup(); return FB_OK; }  // very high valence: two-phase path
    if (incp[v + 1] - incp[start] > CAP || bp[v + 1] - bp[start] > BCAP) { ctaV.push_back(v); start = v; }
  }
  ctaV.push_back(nV);
  const int nCta = (int)ctaV.size() - 1;

  GP_TRY(fb_dev_alloc(c, &c->ga_inc, (size_t)nInc));
  GP_TRY(fb_dev_alloc(c, &c->ga_ctaV, (size_t)nCta + 1));
  GP_TRY(fb_dev_alloc(c, &c->ga_csrc, 16 * (size_t)c->nT));
  GP_TRY(fb_dev_alloc(c, &c->ga_sblk, 4 * (size_t)nInc));
  GP_TRY(fb_dev_alloc(c, &c->ga_border, (size_t)nB));
  GP_TRY(fb_dev_alloc(c, &c->ga_erec, (size_t)EREC * (size_t)c->nT));
  GP_TRY(fb_dev_alloc(c, &c->ga_xu, 6 * (size_t)nV));
  GP_CUDA(cudaMalloc(&vcta, sizeof(int) * (size_t)nV));
  GP_CUDA(cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)nB));
  GP_CUDA(cudaMalloc(&keys2, sizeof(unsigned long long) * (size_t)nB));
  GP_CUDA(cudaMalloc(&ids, sizeof(int) * (size_t)nB));
  GP_CUDA(cudaMalloc(&ids2, sizeof(int) * (size_t)nB));
  GP_CUDA(cudaMemcpyAsync(c->ga_ctaV, ctaV.data(), sizeof(int) * ctaV.size(), cudaMemcpyHostToDevice, st));
  k_fill_inc<<<grid_for((size_t)nV, 256), 256, 0, st>>>(nV, c->diag, c->seg, c->src, c->ga_incp, c->ga_inc);
  k_vertex_cta<<<nCta, 64, 0, st>>>(nCta, c->ga_ctaV, vcta);
  k_fill_csrc<<<grid_for((size_t)nB, 128), 128, 0, st>>>(nB, c->brow, c->bp, c->seg, c->src, c->ga_incp, c->ga_inc, c->ga_ctaV, vcta,
                                                         c->ga_csrc, c->ga_sblk, keys, ids);
  k_fill_erec<<<grid_for((size_t)c->nT, 128), 128, 0, st>>>(c->nT, c->edata, c->ga_erec);
  int ctaBits = 1;
  while ((1ll << ctaBits) < (long long)nCta) ctaBits++;
  dk = cub::DoubleBuffer<unsigned long long>(keys, keys2);
  dv = cub::DoubleBuffer<int>(ids, ids2);
  GP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, dk, dv, (int64_t)nB, 0, 16 + ctaBits, st));
  GP_CUDA(cudaMemcpyAsync(c->ga_border, dv.Current(), sizeof(int) * (size_t)nB, cudaMemcpyDeviceToDevice, st));
  c->launches += 8;
  GP_CUDA(cudaStreamSynchronize(st));
  GP_CUDA(cudaGetLastError());
  cleanup();
#undef GP_CUDA
#undef GP_TRY
  c->ga_ctas = nCta;
  return FB_OK;
}

int fb_launch_assembly_gather(fb_context *c, const double *u, double *Kraw, bool effective) {
  k_rotation<<<grid_for((size_t)c->nT, 128), 128, 0, c->stream>>>(c->nT, c->tets, c->x0, u, c->prm.polar_tolerance, c->ga_erec);
  k_pack_xu<<<grid_for((size_t)c->r, 256), 256, 0, c->stream>>>(c->nV, c->x0, u, c->ga_xu);
  GatherParams p;
  p.scale = c->prm.internal_force_scaling; p.h = c->prm.timestep;
  p.dampK = c->prm.damping_stiffness; p.dampM = c->prm.damping_mass;
  p.effective = effective ? 1 : 0;
  p.ctaV = c->ga_ctaV; p.incp = c->ga_incp; p.bp = c->bp; p.bc = c->bc; p.diag = c->diag; p.seg = c->seg;
  p.border = c->ga_border; p.inc = c->ga_inc; p.csrc = c->ga_csrc; p.sblk = c->ga_sblk;
  p.erec = c->ga_erec; p.xu = c->ga_xu; p.mblk = c->mblk; p.fixed = c->rowmask;
  p.Kraw = Kraw; p.T = c->T; p.Keff = c->Keff; p.invD = c->invD; p.f = c->fint;
  if (c->ga_cfg == 2) launch_gather<CfgBig>(c, p);
  else if (c->ga_cfg == 1) launch_gather<CfgMid>(c, p);
  else launch_gather<CfgSmall>(c, p);
  c->launches += 3;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}
