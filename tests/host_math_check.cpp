// host_math_check.cpp — TEST ONLY.  Instantiates fembrain_b200/csrc/fb_element_math.h for the host
// so that the element arithmetic's operation order can be checked bit-for-bit against the oracle
// in a container without a GPU (tests/test_element_math_host.py builds and loads it).  It is not
// part of the product: libfembrain_b200.so has no host compute path.
#include <string.h>
#include "../fembrain_b200/csrc/fb_element_math.h"

extern "C" {

// per-element G (4x3 of MInverse) and K0 (144, row-major 12x12)
void hm_element_data(int nT, const int *tets, const double *x0, double lambda, double mu, double *G12, double *K0_144) {
  for (int el = 0; el < nT; el++) {
    double x[4][3];
    for (int v = 0; v < 4; v++)
      for (int c = 0; c < 3; c++) x[v][c] = x0[3 * tets[4 * el + v] + c];
    double G[12];
    fbm::minverse_4x3(x, G, 0);
    double vol = fbm::tet_volume(x[0], x[1], x[2], x[3]);
    memcpy(G12 + 12 * (size_t)el, G, sizeof(G));
    for (int j = 0; j < 4; j++) {
      double eb[9];
      fbm::eb_products(G + 3 * j, lambda, mu, eb);
      for (int i = 0; i < 4; i++) {
        double K[9];
        fbm::k0_block(G + 3 * i, eb, vol, K);
        for (int m = 0; m < 3; m++)
          for (int l = 0; l < 3; l++) K0_144[144 * (size_t)el + 12 * (3 * i + m) + 3 * j + l] = K[3 * m + l];
      }
    }
  }
}

// in-order assembly exactly like the reference's scatter, driven by the oracle's CSR row starts
// (Kia) and element->position map (colIdx16)
void hm_assemble(int nV, int nT, const int *tets, const double *x0, const double *u, double lambda, double mu,
                 double tol, const int *Kia, const int *colIdx16, double *f, double *Ka, int *polarIters) {
  memset(f, 0, sizeof(double) * 3 * (size_t)nV);
  memset(Ka, 0, sizeof(double) * (size_t)Kia[3 * nV]);
  for (int el = 0; el < nT; el++) {
    const int *vt = tets + 4 * el;
    double X0[4][3], P[4][3];
    for (int v = 0; v < 4; v++)
      for (int c = 0; c < 3; c++) {
        X0[v][c] = x0[3 * vt[v] + c];
        P[v][c] = X0[v][c] + u[3 * vt[v] + c];
      }
    double G[12], F[9], R[9];
    fbm::minverse_4x3(X0, G, 0);
    double vol = fbm::tet_volume(X0[0], X0[1], X0[2], X0[3]);
    fbm::deformation_gradient(P, G, F);
    int it;
    double det = fbm::polar_rotation(F, R, tol, &it);
    if (polarIters) polarIters[el] = it;
    if (det < 0)
      for (int i = 0; i < 9; i++) R[i] *= -1.0;
    double eb[4][9];
    for (int j = 0; j < 4; j++) fbm::eb_products(G + 3 * j, lambda, mu, eb[j]);
    for (int i = 0; i < 4; i++) {
      double facc[3] = {0, 0, 0};
      for (int j = 0; j < 4; j++) {
        double K[9], RK[9], Kel[9];
        fbm::k0_block(G + 3 * i, eb[j], vol, K);
        fbm::warp_block(R, K, RK, Kel);
        fbm::force_accumulate(Kel, RK, P[j], X0[j], facc);
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++) Ka[Kia[3 * vt[i] + k] + 3 * colIdx16[16 * el + 4 * i + j] + l] += Kel[3 * k + l];
      }
      for (int k = 0; k < 3; k++) f[3 * vt[i] + k] += facc[k];
    }
  }
}

// the same scatter for any warp (0 linear, 1 corotational, 2 exact tangent) through fbm::element_full — the function the
// device kernel k_element_full calls
void hm_assemble_warp(int nV, int nT, const int *tets, const double *x0, const double *u, double lambda, double mu, double tol,
                      const int *Kia, const int *colIdx16, int warp, double *f, double *Ka) {
  memset(f, 0, sizeof(double) * 3 * (size_t)nV);
  memset(Ka, 0, sizeof(double) * (size_t)Kia[3 * nV]);
  for (int el = 0; el < nT; el++) {
    const int *vt = tets + 4 * el;
    double X0[4][3], U[4][3];
    for (int v = 0; v < 4; v++)
      for (int c = 0; c < 3; c++) {
        X0[v][c] = x0[3 * vt[v] + c];
        U[v][c] = u[3 * vt[v] + c];
      }
    double G[12], KE[144], fEl[12];
    fbm::minverse_4x3(X0, G, 0);
    double vol = fbm::tet_volume(X0[0], X0[1], X0[2], X0[3]);
    fbm::element_full(warp, X0, U, G, vol, lambda, mu, tol, KE, fEl);
    for (int j = 0; j < 4; j++)
      for (int l = 0; l < 3; l++) f[3 * vt[j] + l] += fEl[3 * j + l];
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++)
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++) Ka[Kia[3 * vt[i] + k] + 3 * colIdx16[16 * el + 4 * i + j] + l] += KE[12 * (3 * i + k) + 3 * j + l];
  }
}

double hm_polar(const double *F, double *R, double tol, int *iters) { return fbm::polar_rotation(F, R, tol, iters); }
}
