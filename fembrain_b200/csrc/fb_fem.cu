// fb_fem.cu — corotational element kernel, deterministic assembly, effective-matrix / rhs formation
// and state update.  COMPILED WITH -fmad=false: every value these kernels produce is bit-identical
// to the reference's (same operations, same order, IEEE double, no contraction).
//
// Reference (src/3rdparty/vegafem unless noted):
//   per-tet F, polar, R K0 R^T, f_el .. corotationalLinearFEM/corotationalLinearFEM.cpp:219-293
//   scatter into K / f .............. corotationalLinearFEM.cpp:288-293, 456-468
//   mass matrix ..................... volumetricMesh/generateMassMatrix.cpp:33-76, tetMesh.cpp:150-182
//   Keff / rhs / state update ....... src/deformable/PS_VolumeConservingIntegrator.cpp:84-123, 230-237
//
// Assembly is two-phase and atomic-free.  Phase 1 (k_element, one thread per tetrahedron, blocks
// written through a per-warp shared-memory transpose so global stores are 256-byte coalesced)
// leaves the sixteen 3x3 blocks of K_el in scrK[ij][el][9] and f_el in scrF[c][el].  Phase 2
// (k_reduce_K, one thread per scalar of K) walks the block's precomputed contribution list
// (ascending element id — the order in which the reference's element loop adds into the matrix)
// and sums sequentially, so K is bit-identical to the reference's; the same thread then forms
// T = h*K + D and Keff = M + h*D + h^2*K with the reference's sequence of roundings and writes
// them in the reference's CSR value order.
#include "fb_element_math.h"
#include "fb_internal.h"

namespace {

constexpr int EL_TB = 128;

// ---- setup: per-element G (4x3 of MInverse), volume, Lame parameters ---------------------------
__global__ void k_element_data(int nT, const int *__restrict__ tets, const double *__restrict__ x0,
                               const double *__restrict__ E, const double *__restrict__ nu,
                               const double *__restrict__ rho, double Eu, double nuu, double rhou,
                               double *__restrict__ ed) {
  int el = blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nT) return;
  int4 vt = reinterpret_cast<const int4 *>(tets)[el];
  int vi[4] = {vt.x, vt.y, vt.z, vt.w};
  double x[4][3];
  for (int v = 0; v < 4; v++)
    for (int c = 0; c < 3; c++) x[v][c] = x0[3 * (size_t)vi[v] + c];
  double G[12];
  fbm::minverse_4x3(x, G, nullptr);
  double vol = fbm::tet_volume(x[0], x[1], x[2], x[3]);
  for (int k = 0; k < 12; k++) ed[(size_t)k * nT + el] = G[k];
  ed[(size_t)12 * nT + el] = vol;
  double E_ = E ? E[el] : Eu, nu_ = nu ? nu[el] : nuu;
  // ENuMaterial::getLambda / getMu, volumetricMesh/volumetricMeshENuMaterial.h:61-62
  ed[(size_t)13 * nT + el] = (nu_ * E_) / ((1 + nu_) * (1 - 2 * nu_));
  ed[(size_t)14 * nT + el] = E_ / (2 * (1 + nu_));
  ed[(size_t)15 * nT + el] = rho ? rho[el] : rhou;
}

// ---- setup: consistent mass, one scalar per 3x3 block, accumulated in element order ---------------
__global__ void k_mass(int nB, int nT, const int *__restrict__ seg, const unsigned int *__restrict__ src,
                       const double *__restrict__ ed, double *__restrict__ mblk) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nB) return;
  double acc = 0;
  for (int s = seg[b]; s < seg[b + 1]; s++) {
    unsigned int cidx = src[s];
    int el = (int)(cidx >> 4), ij = (int)(cidx & 15);
    double density = ed[(size_t)15 * nT + el], vol = ed[(size_t)12 * nT + el];
    double factor = density * vol / 20;  // tetMesh.cpp:171-172
    double m = ((ij >> 2) == (ij & 3)) ? 2.0 : 1.0;
    acc += factor * m;
  }
  mblk[b] = acc;
}

// ---- phase 1: one thread per tetrahedron ------------------------------------------------------------
__global__ void __launch_bounds__(EL_TB) k_element(int nT, const int *__restrict__ tets, const double *__restrict__ x0,
                                                   const double *__restrict__ u, const double *__restrict__ ed,
                                                   double tol, double *__restrict__ scrK, double *__restrict__ scrF) {
  __shared__ double sm[EL_TB / 32][32 * 9];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int el = blockIdx.x * EL_TB + threadIdx.x;
  const int elw0 = blockIdx.x * EL_TB + warp * 32;  // first element of this warp
  const bool active = el < nT;
  const int e = active ? el : (nT - 1);
  double *sw = sm[warp];

  int4 vt = reinterpret_cast<const int4 *>(tets)[e];
  int vi[4] = {vt.x, vt.y, vt.z, vt.w};
  double X0[4][3], P[4][3];
#pragma unroll
  for (int v = 0; v < 4; v++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
      X0[v][cc] = x0[3 * (size_t)vi[v] + cc];
      P[v][cc] = X0[v][cc] + u[3 * (size_t)vi[v] + cc];
    }
  double G[12];
#pragma unroll
  for (int k = 0; k < 12; k++) G[k] = ed[(size_t)k * nT + e];
  const double vol = ed[(size_t)12 * nT + e], lambda = ed[(size_t)13 * nT + e], mu = ed[(size_t)14 * nT + e];

  double F[9], R[9];
  fbm::deformation_gradient(P, G, F);
  double det = fbm::polar_rotation(F, R, tol, nullptr);
  if (det < 0) {
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] *= -1.0;
  }

  // number of doubles this warp may write per plane (tail warp of the grid)
  const int wlimit = (nT - elw0 >= 32) ? 288 : ((nT - elw0 > 0) ? 9 * (nT - elw0) : 0);
#pragma unroll 1
  for (int i = 0; i < 4; i++) {
    double facc[3] = {0.0, 0.0, 0.0};
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
      double eb[9], K[9], RK[9], Kel[9];
      fbm::eb_products(G + 3 * j, lambda, mu, eb);
      fbm::k0_block(G + 3 * i, eb, vol, K);
      fbm::warp_block(R, K, RK, Kel);
      fbm::force_accumulate(Kel, RK, P[j], X0[j], facc);
      // transpose through shared memory: lane-major [lane][9] -> linear 288 doubles of the plane
#pragma unroll
      for (int q = 0; q < 9; q++) sw[lane * 9 + q] = Kel[q];
      __syncwarp();
      double *dst = scrK + ((size_t)(4 * i + j) * nT + elw0) * 9;
#pragma unroll
      for (int s = 0; s < 9; s++) {
        int t = lane + 32 * s;
        if (t < wlimit) dst[t] = sw[t];
      }
      __syncwarp();
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < 3; k++) scrF[(size_t)(3 * i + k) * nT + el] = facc[k];
    }
  }
}

// ---- phase 1 for warp = 0 (linear FEM) and warp = 2 (exact tangent): whole 12x12 element matrix per thread ------------------
// corotationalLinearFEM.cpp:232-449.  Not on the reference's hot path (Deformable.cpp:186 builds its force model with the
// default warp = 1); offered through fb_set_warp for callers of CorotationalLinearFEMForceModel(fem, warp).  Same scratch
// layout as k_element ([16 blocks][nT][9] and [12][nT]), so phase 2 is shared.
template <int WARP>
__global__ void __launch_bounds__(64) k_element_full(int nT, const int *__restrict__ tets, const double *__restrict__ x0,
                                                     const double *__restrict__ u, const double *__restrict__ ed, double tol,
                                                     double *__restrict__ scrK, double *__restrict__ scrF) {
  const int el = blockIdx.x * blockDim.x + threadIdx.x;
  if (el >= nT) return;
  const int4 vt = reinterpret_cast<const int4 *>(tets)[el];
  const int vi[4] = {vt.x, vt.y, vt.z, vt.w};
  double X0[4][3], U[4][3];
  for (int v = 0; v < 4; v++)
    for (int cc = 0; cc < 3; cc++) {
      X0[v][cc] = x0[3 * (size_t)vi[v] + cc];
      U[v][cc] = u[3 * (size_t)vi[v] + cc];
    }
  double G[12];
  for (int k = 0; k < 12; k++) G[k] = ed[(size_t)k * nT + el];
  const double vol = ed[(size_t)12 * nT + el], lambda = ed[(size_t)13 * nT + el], mu = ed[(size_t)14 * nT + el];
  double KE[144], fEl[12];
  fbm::element_full(WARP, X0, U, G, vol, lambda, mu, tol, KE, fEl);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double *dst = scrK + ((size_t)(4 * i + j) * nT + el) * 9;
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) dst[3 * k + l] = KE[12 * (3 * i + k) + 3 * j + l];
    }
  for (int i = 0; i < 12; i++) scrF[(size_t)i * nT + el] = fEl[i];
}

// ---- phase 2: one thread per scalar entry of K -----------------------------------------------------
struct ReduceParams {
  int nB, nT;
  double scale, h, dampK, dampM;
  bool effective;  // also write T, Keff, invD
};

__global__ void __launch_bounds__(256) k_reduce_K(ReduceParams p, const int *__restrict__ seg, const unsigned int *__restrict__ src,
                                                  const double *__restrict__ scrK, const int *__restrict__ brow,
                                                  const int *__restrict__ bp, const int *__restrict__ diag,
                                                  const double *__restrict__ mblk, const unsigned char *__restrict__ fixed,
                                                  double *__restrict__ Kraw, double *__restrict__ T,
                                                  double *__restrict__ Keff, double *__restrict__ invD) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (size_t)p.nB * 9) return;
  // (visiting blocks in order of list length, to even out the lanes of a warp, was measured SLOWER: 584 vs 480 us at
  //  1M tets — the scratch reads and matrix writes lose their locality)
  const int b = (int)(t / 9), e = (int)(t - (size_t)b * 9);
  int s0 = seg[b], s1 = seg[b + 1];
  double acc = 0.0;  // SparseMatrix::ResetToZero, then AddEntry in element order
  // four contributions in flight per trip (independent index + value loads), added in list order
  int s = s0;
  for (; s + 4 <= s1; s += 4) {
    const unsigned int c0 = __ldg(src + s), c1 = __ldg(src + s + 1), c2 = __ldg(src + s + 2), c3 = __ldg(src + s + 3);
    const double v0 = __ldg(scrK + ((size_t)(c0 & 15) * p.nT + (c0 >> 4)) * 9 + e);
    const double v1 = __ldg(scrK + ((size_t)(c1 & 15) * p.nT + (c1 >> 4)) * 9 + e);
    const double v2 = __ldg(scrK + ((size_t)(c2 & 15) * p.nT + (c2 >> 4)) * 9 + e);
    const double v3 = __ldg(scrK + ((size_t)(c3 & 15) * p.nT + (c3 >> 4)) * 9 + e);
    acc += v0; acc += v1; acc += v2; acc += v3;
  }
  for (; s < s1; s++) {
    const unsigned int cidx = __ldg(src + s);
    acc += __ldg(scrK + ((size_t)(cidx & 15) * p.nT + (cidx >> 4)) * 9 + e);
  }
  int v = brow[b];
  int rs = bp[v], nb = bp[v + 1] - rs;
  int k = e / 3, l = e - 3 * k;
  size_t idx = 9 * (size_t)rs + (size_t)(3 * nb) * k + 3 * (size_t)(b - rs) + l;
  double Kv = acc * p.scale;  // *tangentStiffnessMatrix *= internalForceScalingFactor  (:87)
  if (Kraw) Kraw[idx] = Kv;
  if (!p.effective) return;
  double D = Kv * p.dampK;                 // ScalarMultiply(dampingStiffnessCoef, rayleigh)   (:100)
  double m = (k == l) ? mblk[b] : 0.0;
  if (k == l) D += p.dampM * m;            // rayleigh->AddSubMatrix(dampingMassCoef, M)        (:102)
  double Tv = Kv * p.h;                    // K *= h                                            (:110)
  Tv += D;                                 // K += D                                            (:112)
  T[idx] = Tv;                             // (K += 1.0 * empty dampingMatrix: no entries)      (:113)
  double Ke = Tv * p.h;                    // K *= h                                            (:115)
  if (k == l) Ke += 1.0 * m;               // K->AddSubMatrix(1.0, M)                           (:116)
  Keff[idx] = Ke;
  if (k == l && b == diag[v]) {
    int dof = 3 * v + k;
    invD[dof] = fixed[dof] ? 0.0 : 1.0 / Ke;  // CGSolver.cpp:134-136 on the constrained system
  }
}

__global__ void k_reduce_f(int nV, int nT, double scale, const int *__restrict__ diag, const int *__restrict__ seg,
                           const unsigned int *__restrict__ src, const double *__restrict__ scrF, double *__restrict__ f) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * nV) return;
  int v = t / 3, l = t - 3 * v;
  int b = diag[v];
  double acc = 0.0;
  if (b >= 0) {
    for (int s = seg[b]; s < seg[b + 1]; s++) {
      unsigned int cidx = src[s];
      int ij = (int)(cidx & 15), i = ij >> 2;
      if (i != (ij & 3)) continue;  // only (i,i) pairs are vertex incidences (degenerate tets repeat a vertex)
      size_t el = cidx >> 4;
      acc += scrF[(size_t)(3 * i + l) * nT + el];
    }
  }
  f[t] = acc * scale;
}

// ---- exact-order SpMV for T * qvel (one thread per scalar row, sequential, no FMA) -------------------
__global__ void k_spmv_exact(int nV, const int *__restrict__ bp, const int *__restrict__ bc, const double *__restrict__ A,
                             const double *__restrict__ x, double *__restrict__ y) {
  int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= 3 * nV) return;
  int v = row / 3, k = row - 3 * v;
  int rs = bp[v], nb = bp[v + 1] - rs;
  const double *a = A + 9 * (size_t)rs + (size_t)(3 * nb) * k;
  double acc = 0;
  for (int j = 0; j < nb; j++) {
    int cb = 3 * bc[rs + j];
    acc += x[cb + 0] * a[3 * j + 0];
    acc += x[cb + 1] * a[3 * j + 1];
    acc += x[cb + 2] * a[3 * j + 2];
  }
  y[row] = acc;
}

// rhs: qres = (T qvel + (fint - fext)) * (-h); constrained rows -> 0 (RemoveRows drops them)  (:119-123, :160)
__global__ void k_rhs(int r, double h, const double *__restrict__ Tq, const double *__restrict__ fint,
                      const double *__restrict__ fext, const unsigned char *__restrict__ fixed,
                      double *__restrict__ qres, double *__restrict__ rhs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r) return;
  double v = Tq[i];
  v += fint[i] - fext[i];
  v *= -h;
  qres[i] = v;
  rhs[i] = fixed[i] ? 0.0 : v;
}

// qvel += qdelta; q += h*qvel; constrained DOFs zeroed; qaccel = 0   (:230-237, :55)
__global__ void k_state_update(int r, double h, const double *__restrict__ dv, const unsigned char *__restrict__ fixed,
                               double *__restrict__ q, double *__restrict__ qvel, double *__restrict__ qaccel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r) return;
  if (fixed[i]) {
    q[i] = 0.0; qvel[i] = 0.0; qaccel[i] = 0.0;
    return;
  }
  double v = qvel[i] + dv[i];
  qvel[i] = v;
  q[i] = q[i] + h * v;
  qaccel[i] = 0.0;
}

// inspection: expand G to MInverse's first three columns and recompute K0 (144) for a range of elements
__global__ void k_expand_element(int nT, int el0, int n, const double *__restrict__ ed, const double *__restrict__ x0,
                                 const int *__restrict__ tets, double *__restrict__ minv16, double *__restrict__ k0) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  int el = el0 + t;
  double G[12];
  for (int k = 0; k < 12; k++) G[k] = ed[(size_t)k * nT + el];
  double vol = ed[(size_t)12 * nT + el], lambda = ed[(size_t)13 * nT + el], mu = ed[(size_t)14 * nT + el];
  if (minv16) {
    // fourth column of MInverse (never used on the path) from the reference's cofactor formula is
    // not stored; report the 4x3 part and leave column 3 as NaN-free zeros
    for (int j = 0; j < 4; j++) {
      for (int cc = 0; cc < 3; cc++) minv16[16 * (size_t)t + 4 * j + cc] = G[3 * j + cc];
      minv16[16 * (size_t)t + 4 * j + 3] = 0.0;
    }
  }
  if (k0) {
    for (int j = 0; j < 4; j++) {
      double eb[9];
      fbm::eb_products(G + 3 * j, lambda, mu, eb);
      for (int i = 0; i < 4; i++) {
        double K[9];
        fbm::k0_block(G + 3 * i, eb, vol, K);
        for (int m = 0; m < 3; m++)
          for (int l = 0; l < 3; l++) k0[144 * (size_t)t + 12 * (3 * i + m) + 3 * j + l] = K[3 * m + l];
      }
    }
  }
}

inline unsigned grid_for(size_t n, int tb) { return (unsigned)((n + tb - 1) / tb); }

}  // namespace

int fb_launch_element_data(fb_context *c, const double *E, const double *nu, const double *rho) {
  if (c->nT == 0) return FB_OK;
  k_element_data<<<grid_for(c->nT, 128), 128, 0, c->stream>>>(c->nT, c->tets, c->x0, E, nu, rho, c->prm.youngs_modulus,
                                                               c->prm.poisson_ratio, c->prm.density, c->edata);
  c->launches++;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_launch_mass(fb_context *c) {
  if (c->nB == 0) return FB_OK;
  k_mass<<<grid_for(c->nB, 256), 256, 0, c->stream>>>(c->nB, c->nT, c->seg, c->src, c->edata, c->mblk);
  c->launches++;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

// rhs: also qresidual = (hK + D) qvel and the right-hand side of the velocity solve.  The gather path does it inside its one
// kernel (T = hK + D is never stored); the two-phase path stores T and runs the exact-order product + k_rhs.
int fb_launch_assembly(fb_context *c, const double *u, double *Kraw, bool effective, bool rhs) {
  if (c->nT == 0) {
    if (effective && rhs && c->r > 0) {   // no elements: T = 0
      FB_CUDA(cudaMemsetAsync(c->tmp, 0, sizeof(double) * (size_t)c->r, c->stream));
      return fb_launch_rhs(c);
    }
    return FB_OK;
  }
  if (c->ga_ctas > 0 && c->warp == 1) return fb_launch_assembly_gather(c, u, Kraw, effective, rhs);
  if (effective && !c->T) FB_TRY(fb_dev_alloc(c, &c->T, (size_t)c->nnzK));
  if (!c->scrK) FB_TRY(fb_dev_alloc(c, &c->scrK, 144 * (size_t)c->nT));
  if (!c->scrF) FB_TRY(fb_dev_alloc(c, &c->scrF, 12 * (size_t)c->nT));
  if (c->warp == 0)
    k_element_full<0><<<grid_for(c->nT, 64), 64, 0, c->stream>>>(c->nT, c->tets, c->x0, u, c->edata, c->prm.polar_tolerance, c->scrK, c->scrF);
  else if (c->warp == 2)
    k_element_full<2><<<grid_for(c->nT, 64), 64, 0, c->stream>>>(c->nT, c->tets, c->x0, u, c->edata, c->prm.polar_tolerance, c->scrK, c->scrF);
  else
    k_element<<<grid_for(c->nT, EL_TB), EL_TB, 0, c->stream>>>(c->nT, c->tets, c->x0, u, c->edata, c->prm.polar_tolerance,
                                                                c->scrK, c->scrF);
  ReduceParams p;
  p.nB = c->nB; p.nT = c->nT;
  p.scale = c->prm.internal_force_scaling; p.h = c->prm.timestep;
  p.dampK = c->prm.damping_stiffness; p.dampM = c->prm.damping_mass;
  p.effective = effective;
  k_reduce_K<<<grid_for((size_t)c->nB * 9, 256), 256, 0, c->stream>>>(p, c->seg, c->src, c->scrK, c->brow, c->bp, c->diag, c->mblk,
                                                                     c->rowmask, Kraw, c->T, c->Keff, c->invD);
  k_reduce_f<<<grid_for((size_t)c->r, 256), 256, 0, c->stream>>>(c->nV, c->nT, c->prm.internal_force_scaling, c->diag, c->seg,
                                                                 c->src, c->scrF, c->fint);
  c->launches += 3;
  FB_CUDA(cudaGetLastError());
  if (effective && rhs) {
    FB_TRY(fb_launch_spmv_exact(c, c->T, c->qvel, c->tmp));
    FB_TRY(fb_launch_rhs(c));
  }
  return FB_OK;
}

int fb_launch_spmv_exact(fb_context *c, const double *A, const double *x, double *y) {
  if (c->r == 0) return FB_OK;
  k_spmv_exact<<<grid_for((size_t)c->r, 128), 128, 0, c->stream>>>(c->nV, c->bp, c->bc, A, x, y);
  c->launches++;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_launch_rhs(fb_context *c) {
  if (c->r == 0) return FB_OK;
  k_rhs<<<grid_for((size_t)c->r, 256), 256, 0, c->stream>>>(c->r, c->prm.timestep, c->tmp, c->fint, c->fext, c->rowmask, c->qres, c->rhs);
  c->launches++;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_launch_state_update(fb_context *c) {
  if (c->r == 0) return FB_OK;
  k_state_update<<<grid_for((size_t)c->r, 256), 256, 0, c->stream>>>(c->r, c->prm.timestep, c->x, c->fixed, c->q, c->qvel, c->qaccel);
  c->launches++;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_launch_expand_element(fb_context *c, double *minv16_dev, double *k0_dev, int el0, int n) {
  if (n <= 0) return FB_OK;
  k_expand_element<<<grid_for((size_t)n, 128), 128, 0, c->stream>>>(c->nT, el0, n, c->edata, c->x0, c->tets, minv16_dev, k0_dev);
  c->launches++;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}
