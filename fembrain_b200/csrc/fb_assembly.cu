// fb_assembly.cu — row-gather assembly: K, Keff, f AND the right-hand side -h ((hK + D) qvel + f - fext) in ONE pass over the
// elements' data; element matrices and T = hK + D are never written to HBM.  COMPILED WITH -fmad=false (bit-identical to the
// reference, like fb_fem.cu).
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
//                      block row.  The CTA's index lists are one precomputed blob (GaLists) copied to shared memory,
//                      then all element / vertex records it names follow as 16-byte asynchronous copies: two
//                      dependent memory round trips per CTA.  ONE THREAD PER INCIDENCE takes its record into registers
//                      and walks j = 0..3: recomputes the K0 block from the 4x3 of MInverse, rotates it (R K0, R K0 R^T)
//                      and adds its nine force terms to the running f_el in the reference's order (j outer, 12
//                      sequential adds per component).  (Round 1 used four lanes per incidence exchanging the force
//                      terms through the record behind __syncwarp: ~3x the shared-memory traffic, 12 % slower at 10M tets;
//                      still selectable with FEMBRAIN_B200_GA_QUAD=1 for timing.)  The blocks
//                      overwrite the incidence's record in shared memory; then one thread per (block, row of 3
//                      scalars) of the CTA's rows adds its contribution list IN ASCENDING ELEMENT ORDER — the order
//                      in which the reference's element loop calls AddEntry — and applies the DoTimestep epilogue.
//                      T = hK + D exists only in registers: each (block, row) thread multiplies its three entries with the
//                      column vertex's velocity (staged beside x0/u), the nine products of every block pass through the
//                      now idle `vals`, and the thread that sums a DOF's element forces also adds the products of its row
//                      in CSR order from 0 — SparseMatrix::MultiplyVector's order (sparseMatrix.cpp:736-749) — and writes
//                      qresidual / rhs (PS_VolumeConservingIntegrator.cpp:119-123, :160).
// No atomics, no float reassociation: K, f, T, Keff are bit-identical to the reference's (and to the two-phase
// path).  Blocks of a CTA are visited in order of decreasing list length (precomputed), so the lanes of a warp run
// the same number of trips: a diagonal block of the cube collects 24 contributions, an edge block 4-6.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fb_element_math.h"
#include "fb_internal.h"

namespace {

constexpr int EREC = 24;  // doubles per element record: R[9], G[12], volume, lambda, mu
// Doubles between two incidences' 36-double areas in shared memory.  One thread per incidence reads its own area, so the lanes
// of a warp are VSTRIDE doubles apart: with 36 (72 words) only four bank groups are hit — an 8-way conflict on every access (ncu:
// 49.6M shared bank conflicts at 1M tets); 37 (74 words, = 2 mod 4) spreads a half-warp's 8-byte accesses over all 32 banks.  The
// areas are then only 8-byte aligned: the records are staged with 8-byte cp.async.
constexpr int VSTRIDE = 37;

// CAP incidences (4 CAP slots) and at most BCAP blocks per CTA, TB threads: 4 CAP / TB compute passes
template <int CAP_, int TB_, int BCAP_, int MINB_, bool PERINC_ = false>
struct GaCfg {
  static constexpr int CAP = CAP_, TB = TB_, BCAP = BCAP_, MINB = MINB_;
  static constexpr bool PERINC = PERINC_;   // compute phase: one thread per incidence (its four blocks in turn) instead of four lanes
  static_assert((4 * CAP_) % TB_ == 0, "whole passes");
  static_assert(BCAP_ <= 4096 && 4 * CAP_ <= 65535 && CAP_ * 37 <= 65535 && CAP_ % 8 == 0 && BCAP_ % 8 == 0, "16-bit list entries, 16-byte sections");
};

// Index lists of one CTA, built once at setup, CTA-relative, copied verbatim into shared memory (16-byte pieces).
template <int CAP, int BCAP>
struct GaLists {
  int v0, nVl, nSlots, blk0, nBlk, pad[3];
  unsigned int inc[CAP];          // element*4 + i of every incidence
  unsigned int bcol[BCAP];        // column vertex of every block
  double mb[BCAP];                // mass scalar of every block (M = mb (x) I3)
  unsigned short csl[CAP * 4];    // contribution (list order) -> offset of its slot's block in vals (doubles)
  unsigned short sbl[CAP * 4];    // slot -> block | i << 12
  unsigned short segl[BCAP + 8];  // list start of every block (+ end)
  unsigned short bol[BCAP];       // blocks in order of decreasing list length
  unsigned short bvl[BCAP];       // block -> local vertex
  unsigned short bpl[CAP + 8];    // local vertex -> first block (+ end)
  unsigned short dgl[CAP];        // local vertex -> its diagonal block
  unsigned short ipl[CAP + 8];    // local vertex -> first incidence (+ end)
};

// Shared memory of one CTA:
//   vals [CAP][VSTRIDE] per incidence (36 used): IN the element record (24 doubles)  ->  OUT four 3x3 blocks, slot = 4 li + j at slot*9
//   xb   [BCAP][6]  per block of the CTA's rows: x0 and u of the block's column vertex
//   qv   [BCAP][3]  ... and its velocity (right-hand side only)
//                   (between the two, the quad's 4 x 9 force terms pass through the same 36 doubles)
//   fel  [CAP][3]   element force rows of the incidence
template <class C>
struct GatherSmem {
  double vals[C::CAP * VSTRIDE];
  double xb[C::BCAP * 6];
  double qv[C::BCAP * 3];
  double fel[C::CAP * 3];
  GaLists<C::CAP, C::BCAP> L;
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

__global__ void k_vertex_cta(int nCta, const int *__restrict__ ctaV, int *__restrict__ vcta) {
  int c = blockIdx.x;
  for (int v = ctaV[c] + threadIdx.x; v < ctaV[c + 1]; v += blockDim.x) vcta[v] = c;
}

// one thread per vertex: header, incidences, row starts and diagonal of the vertex in its CTA's lists
template <int CAP, int BCAP>
__global__ void k_lists_vertex(int nV, const int *__restrict__ diag, const int *__restrict__ seg, const unsigned int *__restrict__ src,
                               const int *__restrict__ incp, const int *__restrict__ bp, const int *__restrict__ ctaV,
                               const int *__restrict__ vcta, unsigned int *__restrict__ inc, GaLists<CAP, BCAP> *__restrict__ lists) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nV) return;
  const int cta = vcta[v], vFirst = ctaV[cta], vEnd = ctaV[cta + 1];
  GaLists<CAP, BCAP> &L = lists[cta];
  const int vl = v - vFirst, blk0 = bp[vFirst], inc0 = incp[vFirst];
  if (v == vFirst) {
    L.v0 = vFirst; L.nVl = vEnd - vFirst; L.nSlots = 4 * (incp[vEnd] - inc0); L.blk0 = blk0; L.nBlk = bp[vEnd] - blk0;
    L.pad[0] = L.pad[1] = L.pad[2] = 0;
  }
  L.bpl[vl] = (unsigned short)(bp[v] - blk0);
  L.ipl[vl] = (unsigned short)(incp[v] - inc0);
  if (v == vEnd - 1) {
    L.bpl[vl + 1] = (unsigned short)(bp[vEnd] - blk0);
    L.ipl[vl + 1] = (unsigned short)(incp[vEnd] - inc0);
  }
  const int b = diag[v];
  L.dgl[vl] = (unsigned short)(b - blk0);
  int o = incp[v];
  for (int s = seg[b]; s < seg[b + 1]; s++) {
    const unsigned c = src[s], ij = c & 15u;
    if ((ij >> 2) == (ij & 3u)) {
      const unsigned val = ((c >> 4) << 2) | (ij >> 2);  // element*4 + i
      inc[o] = val;
      L.inc[o - inc0] = val;
      o++;
    }
  }
}

// one thread per block: slot of every contribution, block (and i) of every slot, list start, column, mass, local
// vertex; and the sort key (cta, -length)
template <int CAP, int BCAP>
__global__ void k_lists_block(int nB, const int *__restrict__ brow, const int *__restrict__ bc, const int *__restrict__ bp,
                              const int *__restrict__ seg, const unsigned int *__restrict__ src, const int *__restrict__ incp,
                              const unsigned int *__restrict__ inc, const int *__restrict__ ctaV, const int *__restrict__ vcta,
                              const double *__restrict__ mblk, GaLists<CAP, BCAP> *__restrict__ lists,
                              unsigned long long *__restrict__ keys, int *__restrict__ ids) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nB) return;
  const int v = brow[b], cta = vcta[v], vFirst = ctaV[cta], vEnd = ctaV[cta + 1];
  GaLists<CAP, BCAP> &L = lists[cta];
  const int blk0 = bp[vFirst], bl = b - blk0, segBase = seg[blk0];
  const int base = incp[v] - incp[vFirst];
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
    L.csl[s - segBase] = (unsigned short)((slot >> 2) * VSTRIDE + (slot & 3) * 9);   // offset of the slot's block in GatherSmem::vals
    L.sbl[slot] = (unsigned short)(bl | (int)(((c >> 2) & 3u) << 12));
  }
  L.segl[bl] = (unsigned short)(s0 - segBase);
  if (b == bp[vEnd] - 1) L.segl[bl + 1] = (unsigned short)(s1 - segBase);
  L.bcol[bl] = (unsigned int)bc[b];
  L.mb[bl] = mblk[b];
  L.bvl[bl] = (unsigned short)(v - vFirst);
  const int len = s1 - s0;
  keys[b] = ((unsigned long long)(unsigned)cta << 16) | (unsigned long long)(65535 - (len > 65535 ? 65535 : len));
  ids[b] = b;
}

// after the sort by (cta, -length): position t of the sorted sequence is position t - blk0 inside its CTA
template <int CAP, int BCAP>
__global__ void k_lists_order(int nB, const int *__restrict__ sorted, const int *__restrict__ brow, const int *__restrict__ bp,
                              const int *__restrict__ ctaV, const int *__restrict__ vcta, GaLists<CAP, BCAP> *__restrict__ lists) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nB) return;
  const int b = sorted[t], cta = vcta[brow[b]], blk0 = bp[ctaV[cta]];
  lists[cta].bol[t - blk0] = (unsigned short)(b - blk0);
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
__global__ void __launch_bounds__(128) k_rotation(int nT, const int *__restrict__ tets, const double *__restrict__ xu,
                                                  const double *__restrict__ ed, double tol, double *__restrict__ erec) {
  const int el = blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nT) return;
  const int4 vt = reinterpret_cast<const int4 *>(tets)[el];
  const int vi[4] = {vt.x, vt.y, vt.z, vt.w};
  double P[4][3];
#pragma unroll
  for (int v = 0; v < 4; v++) {   // the packed (x0, u) record of the vertex: three 16-byte loads instead of six scattered 8-byte ones
    const double2 *r2 = reinterpret_cast<const double2 *>(xu + 6 * (size_t)vi[v]);
    const double2 a = __ldg(r2), b = __ldg(r2 + 1), c = __ldg(r2 + 2);   // x0.x x0.y | x0.z u.x | u.y u.z
    P[v][0] = a.x + b.y;
    P[v][1] = a.y + c.x;
    P[v][2] = b.x + c.y;
  }
  double *rec = erec + (size_t)el * EREC;
  double G[12];  // from the structure-of-arrays planes: coalesced (the records' 192-byte stride cost 85 vs 53 us at 1M tets)
#pragma unroll
  for (int k = 0; k < 12; k++) G[k] = ed[(size_t)k * nT + el];
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
  int effective, wantRhs, prefetchAhead;
  const void *lists;
  const double *erec, *xu, *qvel, *fext;
  const unsigned char *fixed;
  double *Kraw, *Keff, *invD, *f, *qres, *rhs;
};

__device__ __forceinline__ void cp_async16(void *smemDst, const void *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smemDst, const void *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smemDst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <class C>
__global__ void __launch_bounds__(C::TB, C::MINB) k_assemble_gather(const GatherParams p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  GatherSmem<C> &S = *reinterpret_cast<GatherSmem<C> *>(smraw);
  typedef GaLists<C::CAP, C::BCAP> Lists;
  static_assert(sizeof(Lists) % 16 == 0 && offsetof(GatherSmem<C>, L) % 16 == 0, "16-byte copies");
  constexpr int TB = C::TB;
  constexpr int PASSES = 4 * C::CAP / TB;
  const int tid = threadIdx.x;

  // ---- stage 1: the CTA's index lists, verbatim ------------------------------------------------------------------
  if (C::PERINC) {   // the blob of the CTA that will run here two waves later: into L2 now, so that ITS first round trip is short
    const unsigned ahead = blockIdx.x + (unsigned)p.prefetchAhead;
    if (p.prefetchAhead > 0 && ahead < gridDim.x) {
      const unsigned char *g = reinterpret_cast<const unsigned char *>(reinterpret_cast<const Lists *>(p.lists) + ahead);
      for (int t = tid; t < (int)(sizeof(Lists) / 128); t += TB) asm volatile("prefetch.global.L2 [%0];" ::"l"(g + 128 * t));
    }
  }
  {
    const unsigned char *g = reinterpret_cast<const unsigned char *>(reinterpret_cast<const Lists *>(p.lists) + blockIdx.x);
    unsigned char *s = reinterpret_cast<unsigned char *>(&S.L);
    for (int t = tid; t < (int)(sizeof(Lists) / 16); t += TB) cp_async16(s + 16 * t, g + 16 * t);
    cp_async_wait_all();
    __syncthreads();
  }
  const int v0 = S.L.v0, nVl = S.L.nVl, nSlots = S.L.nSlots, blk0 = S.L.blk0, nBlk = S.L.nBlk;

  // ---- stage 2: element records by incidence, vertex records by block column -------------------------------------
#pragma unroll
  for (int ps = 0; ps < PASSES; ps++) {
    const int slot = ps * TB + tid;
    if (slot < nSlots) {
      const unsigned ei = S.L.inc[slot >> 2];
      const double *rec = p.erec + (size_t)(ei >> 2) * EREC + 6 * (slot & 3);  // lane j of the quad copies doubles 6j .. 6j+5
      double *dst = S.vals + (slot >> 2) * VSTRIDE + 6 * (slot & 3);
#pragma unroll
      for (int q = 0; q < 6; q++) cp_async8(dst + q, rec + q);
    }
  }
  for (int t = tid; t < nBlk; t += TB) {
    const double *rec = p.xu + 6 * (size_t)S.L.bcol[t];
    double *dst = S.xb + 6 * t;
    cp_async16(dst, rec);
    cp_async16(dst + 2, rec + 2);
    cp_async16(dst + 4, rec + 4);
    if (p.wantRhs) {
      const double *qsrc = p.qvel + 3 * (size_t)S.L.bcol[t];
      cp_async8(S.qv + 3 * t, qsrc);
      cp_async8(S.qv + 3 * t + 1, qsrc + 1);
      cp_async8(S.qv + 3 * t + 2, qsrc + 2);
    }
  }
  cp_async_wait_all();
  __syncthreads();

  if constexpr (C::PERINC) {
    // ---- compute, one thread per INCIDENCE: R, the 4x3 of MInverse and the material in registers, the four blocks (i, j = 0..3)
    // in turn, the element force row accumulated in registers in the reference's (j, l) order.  Against the four-lanes version:
    // ~13 instead of ~36 shared-memory reads and 9 instead of 18 writes per block, no __syncwarp (nobody else touches this
    // incidence's 36 doubles).  Block j overwrites doubles 9j .. 9j+8 of the record — everything it covers is in registers.
    const int nInc = nSlots >> 2;
    for (int li = tid; li < nInc; li += TB) {
      double *ir = S.vals + li * VSTRIDE;
      double R[9], G[12];
#pragma unroll
      for (int k = 0; k < 9; k++) R[k] = ir[k];
#pragma unroll
      for (int k = 0; k < 12; k++) G[k] = ir[9 + k];
      const double vol = ir[21], lambda = ir[22], mu = ir[23];
      const int i = S.L.sbl[4 * li] >> 12;
      double gi[3];
      gi[0] = i == 0 ? G[0] : (i == 1 ? G[3] : (i == 2 ? G[6] : G[9]));
      gi[1] = i == 0 ? G[1] : (i == 1 ? G[4] : (i == 2 ? G[7] : G[10]));
      gi[2] = i == 0 ? G[2] : (i == 1 ? G[5] : (i == 2 ? G[8] : G[11]));
      double facc[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double *xj = S.xb + 6 * (S.L.sbl[4 * li + j] & 4095);
        double X0j[3], Pj[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          X0j[k] = xj[k];
          Pj[k] = X0j[k] + xj[3 + k];
        }
        double eb[9], K[9], RK[9], Kel[9];
        fbm::eb_products(&G[3 * j], lambda, mu, eb);
        fbm::k0_block(gi, eb, vol, K);
        fbm::warp_block(R, K, RK, Kel);
        fbm::force_accumulate(Kel, RK, Pj, X0j, facc);   // fElement[3i+k], j outer, l inner, from 0 (corotationalLinearFEM.cpp:275-286)
#pragma unroll
        for (int q = 0; q < 9; q++) ir[9 * j + q] = Kel[q];
      }
      S.fel[li * 3 + 0] = facc[0]; S.fel[li * 3 + 1] = facc[1]; S.fel[li * 3 + 2] = facc[2];
    }
  } else {
  // ---- compute: one slot per thread and pass, operands from shared memory ---------------------------------------
  #pragma unroll 1
    for (int ps = 0; ps < PASSES; ps++) {
      if (ps * TB >= nSlots) break;
      const int slot = ps * TB + tid;
      const bool active = slot < nSlots;
      const int sl = active ? slot : (nSlots - 1);  // idle lanes of the last pass repeat the last slot
      const int li = sl >> 2, j = sl & 3;
      const int sb = S.L.sbl[sl];
      const int i = sb >> 12;
      double *ir = S.vals + li * VSTRIDE;
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
      // (corotationalLinearFEM.cpp:275-286): every lane leaves its nine terms in the incidence's 36 doubles (the record
      // has been read by all four lanes), lane k < 3 of the quad then adds the twelve terms of component k in the
      // reference's order; finally the four blocks take the place of the terms
      double d[9];
  #pragma unroll
      for (int k = 0; k < 3; k++)
  #pragma unroll
        for (int l = 0; l < 3; l++) d[3 * k + l] = Kel[3 * k + l] * Pj[l] - RK[3 * k + l] * X0j[l];
      __syncwarp();
      if (active) {
  #pragma unroll
        for (int q = 0; q < 9; q++) ir[9 * j + q] = d[q];
      }
      __syncwarp();
      if (active && j < 3) {
        double f = 0.0;
  #pragma unroll
        for (int jj = 0; jj < 4; jj++) {
          f += ir[9 * jj + 3 * j + 0]; f += ir[9 * jj + 3 * j + 1]; f += ir[9 * jj + 3 * j + 2];
        }
        S.fel[li * 3 + j] = f;
      }
      __syncwarp();
      if (active) {
  #pragma unroll
        for (int q = 0; q < 9; q++) ir[9 * j + q] = Kel[q];
      }
    }
  }
  __syncthreads();

  // ---- K: one thread per (block, row k) of the CTA's rows; SparseMatrix::ResetToZero, then AddEntry in element order
  const int nItems = 3 * nBlk;
  constexpr int NIT = (3 * C::BCAP + TB - 1) / TB;
  double pr[NIT][3];  // T[k][l] * qvel[col][l] of this thread's items
#pragma unroll
  for (int it = 0; it < NIT; it++) {
    pr[it][0] = pr[it][1] = pr[it][2] = 0.0;
    const int item = it * TB + tid;
    if (item < nItems) {
      const int bi = item / 3, k = item - 3 * bi;
      const int bl = S.L.bol[bi];
      const int vl = S.L.bvl[bl];
      const int s0 = S.L.segl[bl], s1 = S.L.segl[bl + 1];
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (int s = s0; s < s1; s++) {
        const double *c = S.vals + (int)S.L.csl[s] + 3 * k;
        a0 += c[0]; a1 += c[1]; a2 += c[2];
      }
      const int rs = S.L.bpl[vl], nb = S.L.bpl[vl + 1] - rs;
      const size_t idx = 9 * (size_t)(blk0 + rs) + (size_t)(3 * nb) * k + 3 * (size_t)(bl - rs);
      const double m = S.L.mb[bl];
      const bool isDiag = (bl == S.L.dgl[vl]);
#pragma unroll
      for (int l = 0; l < 3; l++) {
        const double acc = (l == 0) ? a0 : ((l == 1) ? a1 : a2);
        const double Kv = acc * p.scale;  // *tangentStiffnessMatrix *= internalForceScalingFactor  (:87)
        if (p.Kraw) p.Kraw[idx + l] = Kv;
        if (p.effective) {
          double D = Kv * p.dampK;                 // ScalarMultiply(dampingStiffnessCoef, rayleigh)   (:100)
          if (k == l) D += p.dampM * m;            // rayleigh->AddSubMatrix(dampingMassCoef, M)        (:102)
          double Tv = Kv * p.h;                    // K *= h                                            (:110)
          Tv += D;                                 // K += D                                            (:112)
                                                   // (K += 1.0 * empty dampingMatrix: no entries)      (:113)
          if (p.wantRhs) pr[it][l] = S.qv[3 * bl + l] * Tv;   // K->MultiplyVector(qvel, qresidual), one term     (:119)
          double Ke = Tv * p.h;                    // K *= h                                            (:115)
          if (k == l) Ke += 1.0 * m;               // K->AddSubMatrix(1.0, M)                           (:116)
          p.Keff[idx + l] = Ke;
          if (k == l && isDiag) {
            const int dof = 3 * (v0 + vl) + k;
            p.invD[dof] = p.fixed[dof] ? 0.0 : 1.0 / Ke;  // CGSolver.cpp:134-136 on the constrained system
          }
        }
      }
    }
  }
  if (p.wantRhs) {
    __syncthreads();  // every contribution has been read: vals now carries the products, nine per block in CSR position
#pragma unroll
    for (int it = 0; it < NIT; it++) {
      const int item = it * TB + tid;
      if (item < nItems) {
        const int bi = item / 3, k = item - 3 * bi;
        double *o = S.vals + 9 * (int)S.L.bol[bi] + 3 * k;
        o[0] = pr[it][0]; o[1] = pr[it][1]; o[2] = pr[it][2];
      }
    }
    __syncthreads();
  }
  // ---- f: one thread per DOF of the CTA's rows, incident elements in ascending order (:288-293); then the row's products in
  //      column order from 0, + (f - fext), * -h                                   (PS_VolumeConservingIntegrator.cpp:119-123)
  const int nF = 3 * nVl;
  for (int item = tid; item < nF; item += TB) {
    const int vl = item / 3, k = item - 3 * vl;
    const int a = S.L.ipl[vl], z = S.L.ipl[vl + 1];
    double acc = 0.0;
    for (int t = a; t < z; t++) acc += S.fel[t * 3 + k];
    const double fint = acc * p.scale;
    const size_t dof = 3 * (size_t)(v0 + vl) + k;
    p.f[dof] = fint;
    if (p.wantRhs) {
      double tq = 0.0;
      for (int bl = S.L.bpl[vl]; bl < (int)S.L.bpl[vl + 1]; bl++) {
        const double *t3 = S.vals + 9 * bl + 3 * k;
        tq += t3[0]; tq += t3[1]; tq += t3[2];
      }
      double v = tq;
      v += fint - p.fext[dof];
      v *= -p.h;
      p.qres[dof] = v;
      p.rhs[dof] = p.fixed[dof] ? 0.0 : v;
    }
  }
}

typedef GaCfg<256, 256, 256, 1, true> CfgBig;
typedef GaCfg<192, 192, 192, 2, true> CfgMid;
typedef GaCfg<128, 256, 128, 4> CfgSmall;  // four lanes per incidence (round 1), kept for A/B timing: 64 registers, 54 KB, 4 CTAs/SM
typedef GaCfg<128, 128, 128, 4, true> CfgSmallInc;  // same lists, one thread per incidence, up to 128 registers

template <class C>
cudaError_t prepare_gather(size_t *smem, size_t *listBytes) {
  *smem = sizeof(GatherSmem<C>);
  *listBytes = sizeof(GaLists<C::CAP, C::BCAP>);
  return cudaFuncSetAttribute(k_assemble_gather<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
}

template <class C>
void fill_lists(fb_context *c, const int *vcta, unsigned int *inc, unsigned long long *keys, int *ids) {
  typedef GaLists<C::CAP, C::BCAP> Lists;
  Lists *lists = reinterpret_cast<Lists *>(c->ga_lists);
  cudaStream_t st = c->stream;
  k_lists_vertex<C::CAP, C::BCAP><<<grid_for((size_t)c->nV, 128), 128, 0, st>>>(c->nV, c->diag, c->seg, c->src, c->ga_incp, c->bp, c->ga_ctaV,
                                                                                 vcta, inc, lists);
  k_lists_block<C::CAP, C::BCAP><<<grid_for((size_t)c->nB, 128), 128, 0, st>>>(c->nB, c->brow, c->bc, c->bp, c->seg, c->src, c->ga_incp, inc,
                                                                                c->ga_ctaV, vcta, c->mblk, lists, keys, ids);
}

template <class C>
void order_lists(fb_context *c, const int *sorted, const int *vcta) {
  typedef GaLists<C::CAP, C::BCAP> Lists;
  k_lists_order<C::CAP, C::BCAP><<<grid_for((size_t)c->nB, 256), 256, 0, c->stream>>>(c->nB, sorted, c->brow, c->bp, c->ga_ctaV, vcta,
                                                                                      reinterpret_cast<Lists *>(c->ga_lists));
}

template <class C>
void launch_gather(fb_context *c, const GatherParams &p) {
  k_assemble_gather<C><<<c->ga_ctas, C::TB, sizeof(GatherSmem<C>), c->stream>>>(p);
}

}  // namespace

// Builds the gather plan after fb_build_topology, fb_launch_element_data and fb_launch_mass.  Leaves c->ga_ctas == 0
// (two-phase path) when a vertex has more incident elements or a block row more blocks than one CTA holds, or when
// FEMBRAIN_B200_ASSEMBLY=twophase is set.
int fb_build_gather_plan(fb_context *c) {
  c->ga_ctas = 0;
  const char *env = getenv("FEMBRAIN_B200_ASSEMBLY");
  if (env && !strcmp(env, "twophase")) return FB_OK;
  if (c->nT == 0 || c->nV == 0 || c->nB == 0) return FB_OK;
  cudaStream_t st = c->stream;
  const int nV = c->nV, nB = c->nB;
  int *cnt = nullptr, *vcta = nullptr, *ids = nullptr, *ids2 = nullptr;
  unsigned int *inc = nullptr;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  void *tmp = nullptr;
  auto cleanup = [&]() { for (void *q : {(void *)cnt, (void *)vcta, (void *)ids, (void *)ids2, (void *)inc, (void *)keys, (void *)keys2, tmp}) fb_tmp_free(st, q); };
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

  const bool trace = getenv("FEMBRAIN_B200_TRACE_SETUP") != nullptr;
  auto t0 = std::chrono::steady_clock::now();
  auto lap = [&](const char *what) {
    if (!trace) return;
    cudaStreamSynchronize(st);
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[gather plan] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  };
  GP_CUDA(fb_tmp_alloc(st, &cnt, sizeof(int) * ((size_t)nV + 1)));
  GP_TRY(fb_dev_alloc(c, &c->ga_incp, (size_t)nV + 1));
  k_count_inc<<<grid_for((size_t)nV + 1, 256), 256, 0, st>>>(nV, c->diag, c->seg, c->src, cnt);
  size_t tmpBytes = 0, tb2 = 0;
  GP_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, cnt, c->ga_incp, (int64_t)nV + 1, st));
  cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
  cub::DoubleBuffer<int> dv(nullptr, nullptr);
  GP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, (int64_t)nB, 0, 64, st));
  if (tb2 > tmpBytes) tmpBytes = tb2;
  GP_CUDA(fb_tmp_alloc(st, &tmp, tmpBytes ? tmpBytes : 1));
  GP_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, cnt, c->ga_incp, (int64_t)nV + 1, st));
  c->launches += 2;
  std::vector<int> incp((size_t)nV + 1), bp((size_t)nV + 1);
  GP_CUDA(cudaMemcpyAsync(incp.data(), c->ga_incp, sizeof(int) * incp.size(), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaMemcpyAsync(bp.data(), c->bp, sizeof(int) * bp.size(), cudaMemcpyDeviceToHost, st));
  GP_CUDA(cudaStreamSynchronize(st));
  lap("count + scan + download");
  // Configuration by the mesh's largest vertex: the smallest CTA that holds every block row runs the most CTAs per SM.
  // Measured on B200, 998,250 / 10,110,954 tets (profiles/r01_assembly_gather.txt): 128 incidences per CTA at 4 CTAs/SM
  // 0.68 / 6.70 ms, at 3 CTAs/SM 0.73 / 7.13 ms, 192 per CTA (2 CTAs/SM) 0.81 / 7.95 ms; two-phase path 0.86 / 8.81 ms.
  int maxInc = 0, maxRow = 0;
  for (int v = 0; v < nV; v++) {
    maxInc = std::max(maxInc, incp[v + 1] - incp[v]);
    maxRow = std::max(maxRow, bp[v + 1] - bp[v]);
  }
  const char *capEnv = getenv("FEMBRAIN_B200_GA_CAP");
  const int need = std::max(std::max(maxInc, maxRow), capEnv ? atoi(capEnv) : 0);
  if (need > CfgBig::CAP) { cleanup(); return FB_OK; }  // a vertex with more than 256 incident tets: two-phase path
  int CAP, BCAP;
  size_t smem = 0, listBytes = 0;
  cudaError_t attr;
  if (need > CfgMid::CAP) { c->ga_cfg = 2; CAP = CfgBig::CAP; BCAP = CfgBig::BCAP; attr = prepare_gather<CfgBig>(&smem, &listBytes); }
  else if (need > CfgSmall::CAP) { c->ga_cfg = 1; CAP = CfgMid::CAP; BCAP = CfgMid::BCAP; attr = prepare_gather<CfgMid>(&smem, &listBytes); }
  else {
    c->ga_cfg = 0; CAP = CfgSmall::CAP; BCAP = CfgSmall::BCAP;
    attr = prepare_gather<CfgSmall>(&smem, &listBytes);
    if (attr == cudaSuccess) attr = prepare_gather<CfgSmallInc>(&smem, &listBytes);
  }
  if (attr != cudaSuccess) {
    fb_set_error("cudaFuncSetAttribute(smem %zu) -> %s", smem, cudaGetErrorString(attr));
    cleanup();
    return FB_ERR_CUDA;
  }
  const int nInc = incp[nV];
  // greedy packing of consecutive vertices: at most CAP incidences and BCAP blocks per CTA (host: one pass over nV)
  std::vector<int> ctaV;
  ctaV.reserve((size_t)nInc / (size_t)(CAP - CAP / 8) + 2);
  ctaV.push_back(0);
  int start = 0;
  for (int v = 0; v < nV; v++) {
    if (incp[v + 1] - incp[start] > CAP || bp[v + 1] - bp[start] > BCAP) { ctaV.push_back(v); start = v; }
  }
  ctaV.push_back(nV);
  const int nCta = (int)ctaV.size() - 1;
  lap("host packing");

  GP_TRY(fb_dev_alloc(c, &c->ga_ctaV, (size_t)nCta + 1));
  GP_TRY(fb_dev_alloc(c, &c->ga_lists, listBytes * (size_t)nCta));
  GP_TRY(fb_dev_alloc(c, &c->ga_erec, (size_t)EREC * (size_t)c->nT));
  GP_TRY(fb_dev_alloc(c, &c->ga_xu, 6 * (size_t)nV));
  GP_CUDA(cudaMemsetAsync(c->ga_lists, 0, listBytes * (size_t)nCta, st));
  GP_CUDA(fb_tmp_alloc(st, &inc, sizeof(unsigned int) * (size_t)(nInc ? nInc : 1)));
  GP_CUDA(fb_tmp_alloc(st, &vcta, sizeof(int) * (size_t)nV));
  GP_CUDA(fb_tmp_alloc(st, &keys, sizeof(unsigned long long) * (size_t)nB));
  GP_CUDA(fb_tmp_alloc(st, &keys2, sizeof(unsigned long long) * (size_t)nB));
  GP_CUDA(fb_tmp_alloc(st, &ids, sizeof(int) * (size_t)nB));
  GP_CUDA(fb_tmp_alloc(st, &ids2, sizeof(int) * (size_t)nB));
  GP_CUDA(cudaMemcpyAsync(c->ga_ctaV, ctaV.data(), sizeof(int) * ctaV.size(), cudaMemcpyHostToDevice, st));
  lap("allocations + memset");
  k_vertex_cta<<<nCta, 64, 0, st>>>(nCta, c->ga_ctaV, vcta);
  if (c->ga_cfg == 2) fill_lists<CfgBig>(c, vcta, inc, keys, ids);
  else if (c->ga_cfg == 1) fill_lists<CfgMid>(c, vcta, inc, keys, ids);
  else fill_lists<CfgSmall>(c, vcta, inc, keys, ids);
  lap("lists (vertex, block)");
  k_fill_erec<<<grid_for((size_t)c->nT, 128), 128, 0, st>>>(c->nT, c->edata, c->ga_erec);
  lap("element records");
  int ctaBits = 1;
  while ((1ll << ctaBits) < (long long)nCta) ctaBits++;
  dk = cub::DoubleBuffer<unsigned long long>(keys, keys2);
  dv = cub::DoubleBuffer<int>(ids, ids2);
  GP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmpBytes, dk, dv, (int64_t)nB, 0, 16 + ctaBits, st));
  if (c->ga_cfg == 2) order_lists<CfgBig>(c, dv.Current(), vcta);
  else if (c->ga_cfg == 1) order_lists<CfgMid>(c, dv.Current(), vcta);
  else order_lists<CfgSmall>(c, dv.Current(), vcta);
  c->launches += 9;
  GP_CUDA(cudaStreamSynchronize(st));
  GP_CUDA(cudaGetLastError());
  lap("sort + order");
  cleanup();
  lap("free temporaries");
#undef GP_CUDA
#undef GP_TRY
  c->ga_ctas = nCta;
  return FB_OK;
}

static bool ga_quad() {   // FEMBRAIN_B200_GA_QUAD=1: the four-lanes-per-incidence compute phase (A/B timing)
  static const bool v = getenv("FEMBRAIN_B200_GA_QUAD") != nullptr;
  return v;
}

int fb_launch_assembly_gather(fb_context *c, const double *u, double *Kraw, bool effective, bool rhs) {
  k_pack_xu<<<grid_for((size_t)c->r, 256), 256, 0, c->stream>>>(c->nV, c->x0, u, c->ga_xu);
  k_rotation<<<grid_for((size_t)c->nT, 128), 128, 0, c->stream>>>(c->nT, c->tets, c->ga_xu, c->edata, c->prm.polar_tolerance, c->ga_erec);
  GatherParams p;
  p.scale = c->prm.internal_force_scaling; p.h = c->prm.timestep;
  p.dampK = c->prm.damping_stiffness; p.dampM = c->prm.damping_mass;
  p.effective = effective ? 1 : 0;
  p.wantRhs = (effective && rhs) ? 1 : 0;
  static const int ahead = getenv("FEMBRAIN_B200_GA_PREFETCH") ? atoi(getenv("FEMBRAIN_B200_GA_PREFETCH")) : 2 * 4 * 148;
  p.prefetchAhead = ahead;
  p.lists = c->ga_lists;
  p.erec = c->ga_erec; p.xu = c->ga_xu; p.fixed = c->rowmask;
  p.qvel = c->qvel; p.fext = c->fext; p.qres = c->qres; p.rhs = c->rhs;
  p.Kraw = Kraw; p.Keff = c->Keff; p.invD = c->invD; p.f = c->fint;
  if (c->ga_cfg == 2) launch_gather<CfgBig>(c, p);
  else if (c->ga_cfg == 1) launch_gather<CfgMid>(c, p);
  else if (ga_quad()) launch_gather<CfgSmall>(c, p);
  else launch_gather<CfgSmallInc>(c, p);
  c->launches += 3;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}
