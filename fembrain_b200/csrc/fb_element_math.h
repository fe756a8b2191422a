// fb_element_math.h — per-tetrahedron arithmetic of the corotational element, written so that
// every floating-point result is bit-identical to the reference's.
//
// Device code (fb_assembly.cu) is compiled with -fmad=false: no multiply-add contraction, IEEE
// double divide/sqrt, operations in the reference's order.  The header also compiles for the host
// (FB_HD empty) — ONLY so that tests/host_math_check.cpp can debug operation order against the
// oracle in a container without a GPU; the product library exports no host path.
//
// Reference (relative to /root/reference/src/3rdparty/vegafem):
//   inverse4x4 ............ corotationalLinearFEM/corotationalLinearFEM.cpp:529-572
//   tet volume ............ volumetricMesh/tetMesh.cpp:184-188 (+ minivector/vec3d.h:206-218)
//   polar decomposition ... polarDecomposition/polarDecomposition.cpp:8-108
//   K0 = V * B^T (E B) .... corotationalLinearFEM/corotationalLinearFEM.cpp:98-145
//   WarpMatrix ............ corotationalLinearFEM/corotationalLinearFEM.cpp:191-211
//   F, f_el ............... corotationalLinearFEM/corotationalLinearFEM.cpp:238-286
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD static inline
#endif

namespace fbm {

// Upper 4x3 part of inverse([x0 x1 x2 x3; 1 1 1 1]) (row-major 4x4 "M" of the reference with
// M[4*dim+vtx]); G[3*j+c] = MInverse[4*j+c], c < 3.  The last column of MInverse is never used on
// the path.  Same expressions, same order, as inverse4x4 with A[12..15] = 1 kept as variables.
FB_HD void minverse_4x3(const double x[4][3], double G[12], double *detOut) {
  double A[16];
  for (int vtx = 0; vtx < 4; vtx++)
    for (int dim = 0; dim < 3; dim++) A[4 * dim + vtx] = x[vtx][dim];
  A[12] = A[13] = A[14] = A[15] = 1.0;
  double I0 = -A[11] * A[14] * A[5] + A[10] * A[15] * A[5] + A[11] * A[13] * A[6] - A[10] * A[13] * A[7] - A[15] * A[6] * A[9] + A[14] * A[7] * A[9];
  double I1 = A[1] * A[11] * A[14] - A[1] * A[10] * A[15] - A[11] * A[13] * A[2] + A[10] * A[13] * A[3] + A[15] * A[2] * A[9] - A[14] * A[3] * A[9];
  double I2 = -A[15] * A[2] * A[5] + A[14] * A[3] * A[5] + A[1] * A[15] * A[6] - A[13] * A[3] * A[6] - A[1] * A[14] * A[7] + A[13] * A[2] * A[7];
  double I4 = A[11] * A[14] * A[4] - A[10] * A[15] * A[4] - A[11] * A[12] * A[6] + A[10] * A[12] * A[7] + A[15] * A[6] * A[8] - A[14] * A[7] * A[8];
  double I5 = -A[0] * A[11] * A[14] + A[0] * A[10] * A[15] + A[11] * A[12] * A[2] - A[10] * A[12] * A[3] - A[15] * A[2] * A[8] + A[14] * A[3] * A[8];
  double I6 = A[15] * A[2] * A[4] - A[14] * A[3] * A[4] - A[0] * A[15] * A[6] + A[12] * A[3] * A[6] + A[0] * A[14] * A[7] - A[12] * A[2] * A[7];
  double I8 = -A[11] * A[13] * A[4] + A[11] * A[12] * A[5] - A[15] * A[5] * A[8] + A[13] * A[7] * A[8] + A[15] * A[4] * A[9] - A[12] * A[7] * A[9];
  double I9 = -A[1] * A[11] * A[12] + A[0] * A[11] * A[13] + A[1] * A[15] * A[8] - A[13] * A[3] * A[8] - A[0] * A[15] * A[9] + A[12] * A[3] * A[9];
  double I10 = -A[1] * A[15] * A[4] + A[13] * A[3] * A[4] + A[0] * A[15] * A[5] - A[12] * A[3] * A[5] + A[1] * A[12] * A[7] - A[0] * A[13] * A[7];
  double I12 = A[10] * A[13] * A[4] - A[10] * A[12] * A[5] + A[14] * A[5] * A[8] - A[13] * A[6] * A[8] - A[14] * A[4] * A[9] + A[12] * A[6] * A[9];
  double I13 = A[1] * A[10] * A[12] - A[0] * A[10] * A[13] - A[1] * A[14] * A[8] + A[13] * A[2] * A[8] + A[0] * A[14] * A[9] - A[12] * A[2] * A[9];
  double I14 = A[1] * A[14] * A[4] - A[13] * A[2] * A[4] - A[0] * A[14] * A[5] + A[12] * A[2] * A[5] - A[1] * A[12] * A[6] + A[0] * A[13] * A[6];
  double det = A[0] * I0 + A[1] * I4 + A[2] * I8 + A[3] * I12;
  double invDet = 1.0 / det;
  G[0] = I0 * invDet;  G[1] = I1 * invDet;  G[2] = I2 * invDet;
  G[3] = I4 * invDet;  G[4] = I5 * invDet;  G[5] = I6 * invDet;
  G[6] = I8 * invDet;  G[7] = I9 * invDet;  G[8] = I10 * invDet;
  G[9] = I12 * invDet; G[10] = I13 * invDet; G[11] = I14 * invDet;
  if (detOut) *detOut = det;
}

// TetMesh::getTetVolume: 1/6 * | (a-d) . ((b-d) x (c-d)) |
FB_HD double tet_volume(const double a[3], const double b[3], const double c[3], const double d[3]) {
  double p0 = a[0] - d[0], p1 = a[1] - d[1], p2 = a[2] - d[2];
  double u0 = b[0] - d[0], u1 = b[1] - d[1], u2 = b[2] - d[2];
  double v0 = c[0] - d[0], v1 = c[1] - d[1], v2 = c[2] - d[2];
  double w0 = u1 * v2 - v1 * u2;
  double w1 = -u0 * v2 + v0 * u2;
  double w2 = u0 * v1 - v0 * u1;
  double dt = p0 * w0 + p1 * w1 + p2 * w2;
  return 1.0 / 6 * fabs(dt);
}

FB_HD double one_norm3(const double *A) {
  double norm = 0.0;
  for (int i = 0; i < 3; i++) {
    double s = fabs(A[i + 0]) + fabs(A[i + 3]) + fabs(A[i + 6]);
    if (s > norm) norm = s;
  }
  return norm;
}
FB_HD double inf_norm3(const double *A) {
  double norm = 0.0;
  for (int i = 0; i < 3; i++) {
    double s = fabs(A[3 * i + 0]) + fabs(A[3 * i + 1]) + fabs(A[3 * i + 2]);
    if (s > norm) norm = s;
  }
  return norm;
}
FB_HD void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

// PolarDecomposition::Compute.  Q row-major; S (optional: only the exact-tangent branch, warp = 2, reads it) = Q^T M, then
// symmetrised (polarDecomposition.cpp:94-105).  Identical iteration, scaling and stopping rule => identical trip count and bits.
FB_HD double polar_rotation(const double *M, double *Q, double tolerance, int *itersOut, double *S = nullptr) {
  double Mk[9], Ek[9];
  double det, M1, Minf, E1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Mk[3 * i + j] = M[3 * j + i];
  M1 = one_norm3(Mk);
  Minf = inf_norm3(Mk);
  int it = 0;
  do {
    double Adj[9];
    cross3(&Mk[3], &Mk[6], &Adj[0]);
    cross3(&Mk[6], &Mk[0], &Adj[3]);
    cross3(&Mk[0], &Mk[3], &Adj[6]);
    det = Mk[0] * Adj[0] + Mk[1] * Adj[1] + Mk[2] * Adj[2];
    if (det == 0.0) break;  // the reference prints a warning and breaks (polarDecomposition.cpp:63-67)
    double A1 = one_norm3(Adj), Ainf = inf_norm3(Adj);
    double gamma = sqrt(sqrt((A1 * Ainf) / (M1 * Minf)) / fabs(det));
    double g1 = gamma * 0.5;
    double g2 = 0.5 / (gamma * det);
    for (int i = 0; i < 9; i++) {
      Ek[i] = Mk[i];
      Mk[i] = g1 * Mk[i] + g2 * Adj[i];
      Ek[i] -= Mk[i];
    }
    E1 = one_norm3(Ek);
    M1 = one_norm3(Mk);
    Minf = inf_norm3(Mk);
    it++;
  } while (E1 > M1 * tolerance);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Q[3 * i + j] = Mk[3 * j + i];
  if (itersOut) *itersOut = it;
  if (S) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double acc = 0;
        for (int k = 0; k < 3; k++) acc += Mk[3 * i + k] * M[3 * k + j];
        S[3 * i + j] = acc;
      }
    for (int i = 0; i < 3; i++)
      for (int j = i; j < 3; j++) S[3 * i + j] = S[3 * j + i] = 0.5 * (S[3 * i + j] + S[3 * j + i]);
  }
  return det;
}

// F = P * MInverse, upper-left 3x3 (corotationalLinearFEM.cpp:252-259).
// P[l][k] = world coordinate l of vertex k; G as in minverse_4x3.
FB_HD void deformation_gradient(const double Pw[4][3], const double G[12], double F[9]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double acc = 0;
      for (int k = 0; k < 4; k++) acc += Pw[k][i] * G[3 * k + j];
      F[3 * i + j] = acc;
    }
}

// The nine products E*B that column-block j of EB contains (EB = E*B, corotationalLinearFEM.cpp:115-120):
// eb[0..2] = lambda*g, eb[3..5] = mu*g, eb[6..8] = (lambda+2mu)*g, g = G[3j..3j+2].
// Every entry of EB is a single product (all other terms of the reference's 6-term sums are
// exact zeros), so no rounding differs.
FB_HD void eb_products(const double *g, double lambda, double mu, double eb[9]) {
  const double A = lambda + 2 * mu;
  for (int c = 0; c < 3; c++) {
    eb[c] = lambda * g[c];
    eb[3 + c] = mu * g[c];
    eb[6 + c] = A * g[c];
  }
}

// 3x3 block (i,j) of K0 = volume * B^T (E B): K[3m+l] = K0[3i+m][3j+l].
// h = G[3i..], eb = eb_products of vertex j.  Each entry is the reference's 6-term sum over k with
// the structurally-zero products dropped (adding +-0 never changes a partial sum's value), the
// surviving products added in increasing k, then multiplied by the volume.
FB_HD void k0_block(const double *h, const double *eb, double volume, double K[9]) {
  const double *lg = eb, *mg = eb + 3, *ag = eb + 6;
  // row m=0 uses B rows 0 (h0), 3 (h1), 5 (h2)
  K[0] = ((h[0] * ag[0]) + h[1] * mg[1]) + h[2] * mg[2];
  K[1] = (h[0] * lg[1]) + h[1] * mg[0];
  K[2] = (h[0] * lg[2]) + h[2] * mg[0];
  // row m=1 uses B rows 1 (h1), 3 (h0), 4 (h2)
  K[3] = (h[1] * lg[0]) + h[0] * mg[1];
  K[4] = ((h[1] * ag[1]) + h[0] * mg[0]) + h[2] * mg[2];
  K[5] = (h[1] * lg[2]) + h[2] * mg[1];
  // row m=2 uses B rows 2 (h2), 4 (h1), 5 (h0)
  K[6] = (h[2] * lg[0]) + h[0] * mg[2];
  K[7] = (h[2] * lg[1]) + h[1] * mg[2];
  K[8] = ((h[2] * ag[2]) + h[1] * mg[1]) + h[0] * mg[0];
  for (int e = 0; e < 9; e++) K[e] *= volume;
}

// WarpMatrix for one 3x3 block: RK = R*K, RKRT = RK*R^T (sums over m = 0,1,2 in order, from 0).
FB_HD void warp_block(const double *R, const double *K, double RK[9], double RKRT[9]) {
  for (int k = 0; k < 3; k++)
    for (int l = 0; l < 3; l++) {
      double acc = 0;
      for (int m = 0; m < 3; m++) acc += R[3 * k + m] * K[3 * m + l];
      RK[3 * k + l] = acc;
    }
  for (int k = 0; k < 3; k++)
    for (int l = 0; l < 3; l++) {
      double acc = 0;
      for (int m = 0; m < 3; m++) acc += RK[3 * k + m] * R[3 * l + m];
      RKRT[3 * k + l] = acc;
    }
}

// f_el rows 3i..3i+2 accumulate, for vertex j: Kel*P - RK*x0 (corotationalLinearFEM.cpp:275-286).
// facc[k] is the running fElement[3i+k]; Pj / X0j are vertex j's world and rest coordinates.
FB_HD void force_accumulate(const double Kel[9], const double RK[9], const double *Pj, const double *X0j, double facc[3]) {
  for (int k = 0; k < 3; k++)
    for (int l = 0; l < 3; l++) facc[k] += Kel[3 * k + l] * Pj[l] - RK[3 * k + l] * X0j[l];
}

// ---- warp = 0 and warp = 2 (corotationalLinearFEM.cpp:296-449): whole 12x12 element matrices, one thread per element ------
// w = A v, w = A^T v, c = a b, c = a b^T in the association order of Vega's matrixMultiplyMacros.h (:224-270)
FB_HD void mv3(const double *A, const double *v, double *w) {
  w[0] = A[0] * v[0] + A[1] * v[1] + A[2] * v[2];
  w[1] = A[3] * v[0] + A[4] * v[1] + A[5] * v[2];
  w[2] = A[6] * v[0] + A[7] * v[1] + A[8] * v[2];
}
FB_HD void mtv3(const double *A, const double *v, double *w) {
  w[0] = A[0] * v[0] + A[3] * v[1] + A[6] * v[2];
  w[1] = A[1] * v[0] + A[4] * v[1] + A[7] * v[2];
  w[2] = A[2] * v[0] + A[5] * v[1] + A[8] * v[2];
}
FB_HD void mm3(const double *a, const double *b, double *c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
FB_HD void mmt3(const double *a, const double *b, double *c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
// CorotationalLinearFEM::inverse3x3 (:506-524)
FB_HD void inverse3(const double *A, double *AInv) {
  AInv[0] = -A[5] * A[7] + A[4] * A[8];
  AInv[1] = A[2] * A[7] - A[1] * A[8];
  AInv[2] = -A[2] * A[4] + A[1] * A[5];
  AInv[3] = A[5] * A[6] - A[3] * A[8];
  AInv[4] = -A[2] * A[6] + A[0] * A[8];
  AInv[5] = A[2] * A[3] - A[0] * A[5];
  AInv[6] = -A[4] * A[6] + A[3] * A[7];
  AInv[7] = A[1] * A[6] - A[0] * A[7];
  AInv[8] = -A[1] * A[3] + A[0] * A[4];
  const double invDet = 1.0 / (-A[2] * A[4] * A[6] + A[1] * A[5] * A[6] + A[2] * A[3] * A[7] - A[0] * A[5] * A[7] - A[1] * A[3] * A[8] +
                               A[0] * A[4] * A[8]);
  for (int i = 0; i < 9; i++) AInv[i] *= invDet;
}

// The exact-tangent terms of warp = 2 added to KElement (row-major 12x12), :296-428.  R, S from the polar decomposition (R already
// flipped when det < 0, S not — as in the reference), G = the 4x3 of MInverse, Pw / X0 world and rest positions of the four
// vertices, K0 / RK the element's undeformed and half-warped matrices.
FB_HD void exact_tangent_add(const double *R, const double *S, const double *G, const double Pw[4][3], const double X0[4][3],
                             const double *K0, const double *RK, double *KElement) {
  double Gm[9], temp[9], invG[9];
  const double tr = S[0] + S[4] + S[8];
  for (int i = 0; i < 9; i++) temp[i] = -S[i];
  temp[0] += tr; temp[4] += tr; temp[8] += tr;
  mmt3(temp, R, Gm);          // G = (tr(S) I - S) R^T
  inverse3(Gm, invG);
  double rhs[27], omega[27];  // 3 x 9, column-major
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double t[9];
      for (int k = 0; k < 9; k++) t[k] = 0.0;
      for (int k = 0; k < 3; k++) t[3 * k + j] = R[3 * i + k];   // i-th row of R into column j
      double *a = &rhs[3 * (3 * i + j)];                          // SKEW_PART
      a[0] = 0.5 * (t[7] - t[5]);
      a[1] = 0.5 * (t[2] - t[6]);
      a[2] = 0.5 * (t[3] - t[1]);
    }
  for (int i = 0; i < 27; i++) rhs[i] *= 2.0;
  for (int i = 0; i < 9; i++) mv3(invG, &rhs[3 * i], &omega[3 * i]);
  double dRdF[81];            // each column is skew(omega) R; column-major
  for (int i = 0; i < 9; i++) {
    const double *a = &omega[3 * i];
    double skew[9];
    skew[0] = 0;     skew[1] = -a[2]; skew[2] = a[1];
    skew[3] = a[2];  skew[4] = 0;     skew[5] = -a[0];
    skew[6] = -a[1]; skew[7] = a[0];  skew[8] = 0;
    mm3(skew, R, &dRdF[9 * i]);
  }
  double dRdx[108];           // d R / d x of the tet's 12 coordinates; column-major, each column a row-major 3x3
  for (int k = 0; k < 4; k++)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        double B[9];          // B[i][j][3 kk + l] = dRdF[9 (3 j + l) + 3 i + kk]
        for (int kk = 0; kk < 3; kk++)
          for (int l = 0; l < 3; l++) B[3 * kk + l] = dRdF[9 * (3 * j + l) + 3 * i + kk];
        mv3(B, &G[3 * k], &dRdx[9 * (3 * k + j) + 3 * i]);
      }
  // term 1: \hat{dR/dx_l} K (R^T x - m)
  double tv[12], a[12];
  for (int vtx = 0; vtx < 4; vtx++) {
    mtv3(R, Pw[vtx], &tv[3 * vtx]);
    for (int i = 0; i < 3; i++) tv[3 * vtx + i] -= X0[vtx][i];
  }
  for (int i = 0; i < 12; i++) {
    double acc = 0.0;
    for (int j = 0; j < 12; j++) acc += K0[12 * i + j] * tv[j];
    a[i] = acc;
  }
  for (int column = 0; column < 12; column++) {
    double b[12];
    for (int j = 0; j < 4; j++) mv3(&dRdx[9 * column], &a[3 * j], &b[3 * j]);
    for (int row = 0; row < 12; row++) KElement[12 * row + column] += b[row];
  }
  // term 2: (R K \hat{dR/dx_l}^T) x
  for (int vtx = 0; vtx < 4; vtx++)
    for (int i = 0; i < 3; i++) a[3 * vtx + i] = Pw[vtx][i];
  for (int column = 0; column < 12; column++) {
    double b[12];
    for (int j = 0; j < 4; j++) mtv3(&dRdx[9 * column], &a[3 * j], &b[3 * j]);
    for (int row = 0; row < 12; row++) {
      double contrib = 0.0;
      for (int j = 0; j < 12; j++) contrib += RK[12 * row + j] * b[j];
      KElement[12 * row + column] += contrib;
    }
  }
}

// Whole element, any warp: KE (row-major 12x12) and fEl (12) as ComputeForceAndStiffnessMatrixOfSubmesh forms them (:232-449).
FB_HD void element_full(int warp, const double X0[4][3], const double U[4][3], const double G[12], double vol, double lambda, double mu,
                        double tol, double *KE, double *fEl) {
  double K0[144];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double eb[9], K[9];
      eb_products(G + 3 * j, lambda, mu, eb);
      k0_block(G + 3 * i, eb, vol, K);
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) K0[12 * (3 * i + k) + 3 * j + l] = K[3 * k + l];
    }
  if (warp == 0) {
    // no warp: KElement = K0, f = K u with each vertex's three products summed first (:431-441)
    for (int q = 0; q < 144; q++) KE[q] = K0[q];
    for (int i = 0; i < 12; i++) {
      double acc = 0;
      for (int j = 0; j < 4; j++) acc += KE[12 * i + 3 * j + 0] * U[j][0] + KE[12 * i + 3 * j + 1] * U[j][1] + KE[12 * i + 3 * j + 2] * U[j][2];
      fEl[i] = acc;
    }
    return;
  }
  double P[4][3], F[9], R[9], S[9], RK[144];
  for (int v = 0; v < 4; v++)
    for (int cc = 0; cc < 3; cc++) P[v][cc] = X0[v][cc] + U[v][cc];
  deformation_gradient(P, G, F);
  const double det = polar_rotation(F, R, tol, nullptr, S);
  if (det < 0)
    for (int i = 0; i < 9; i++) R[i] *= -1.0;
  for (int i = 0; i < 4; i++) {
    double facc[3] = {0.0, 0.0, 0.0};
    for (int j = 0; j < 4; j++) {
      double Kb[9], RKb[9], Kel[9];
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) Kb[3 * k + l] = K0[12 * (3 * i + k) + 3 * j + l];
      warp_block(R, Kb, RKb, Kel);
      force_accumulate(Kel, RKb, P[j], X0[j], facc);
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) {
          RK[12 * (3 * i + k) + 3 * j + l] = RKb[3 * k + l];
          KE[12 * (3 * i + k) + 3 * j + l] = Kel[3 * k + l];
        }
    }
    for (int k = 0; k < 3; k++) fEl[3 * i + k] = facc[k];
  }
  if (warp == 2) exact_tangent_add(R, S, G, P, X0, K0, RK, KE);
}

}  // namespace fbm
