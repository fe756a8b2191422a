/*
 * vega_port.c — plain-C restatement of the reference's per-frame FEM step (fbport_* API).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h): only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The product
 * (fembrain_b200/csrc, CUDA only) never includes, links or calls it.
 *
 * PARITY PINNED: every function below is checked bit-for-bit (integers) / to the last ulp
 * (reals, same operation order, no FMA contraction: built with -ffp-contract=off) against the
 * UNMODIFIED reference compiled in place as oracle/_ref/libfembrain_ref.so (tests/test_oracle.py),
 * and against fixtures generated from it and committed under tests/golden/.
 *
 * All reference citations are relative to /root/reference/src; VEGA = 3rdparty/vegafem.
 * Data structures differ from the reference on purpose (flat CSR instead of jagged rows and
 * std::map); the integer results (row order, sorted columns, index maps) and the floating-point
 * operation order are the same.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define FBO_PREFIX fbport_
#include "oracle_api.h"

typedef struct {
  int nV, nT, r;
  double *x0; /* undeformedPositions, 3nV (corotationalLinearFEM.cpp:45-50) */
  int *tets;  /* 4nT */
  double rho, h, dampM, dampK;
  const double *matE, *matNu, *matRho; /* optional per-element materials (fbport_create_materials) */
  double *lambda, *mu; /* per element (corotationalLinearFEM.cpp:55-68) */
  double *MInv;        /* 16 per element */
  double *K0;          /* 144 per element */
  /* vertex adjacency (block structure of K): sorted unique neighbours incl. self */
  int *bp, *bc;
  /* tangent stiffness matrix, scalar CSR */
  int nnzK, *Kia, *Kja;
  double *Ka, *Da; /* Da = rayleighDampingMatrix values */
  int *rowIdx, *colIdx;
  /* mass matrix */
  int nnzM, *Mia, *Mja, *subIdx;
  double *Ma;
  /* constrained system */
  int nC, *cdofs;
  int nS, nnzS, *Sia, *Sja, *superRows, *superIdx, *diagIdx;
  double *Sa;
  /* integrator state (VEGA/integrator/integratorBase.cpp:38-65) */
  double *q, *qvel, *qaccel, *q1, *qvel1, *qaccel1, *fext, *fint, *qres, *qdelta, *buffer, *bufC;
  /* CG buffers (VEGA/sparseSolver/CGSolver.cpp:61-67) */
  double *cr, *cd, *cq, *invD;
  double tAsm, tSolve;
} Port;

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int cmp_u64(const void *a, const void *b) {
  uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
  return (x > y) - (x < y);
}
static int cmp_int(const void *a, const void *b) {
  int x = *(const int *)a, y = *(const int *)b;
  return (x > y) - (x < y);
}

/* position of `col` in the sorted array c[0..n) or -1 (SparseMatrix::GetInverseIndex,
 * VEGA/sparseMatrix/sparseMatrix.cpp:613-620 — linear there, same answer on sorted rows) */
static int find_pos(const int *c, int n, int col) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    if (c[mid] == col) return mid;
    if (c[mid] < col) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

/* CorotationalLinearFEM::inverse4x4, VEGA/corotationalLinearFEM/corotationalLinearFEM.cpp:529-572 */
static void inverse4x4(const double *A, double *AInv) {
  AInv[0] = -A[11] * A[14] * A[5] + A[10] * A[15] * A[5] + A[11] * A[13] * A[6] - A[10] * A[13] * A[7] - A[15] * A[6] * A[9] + A[14] * A[7] * A[9];
  AInv[1] = A[1] * A[11] * A[14] - A[1] * A[10] * A[15] - A[11] * A[13] * A[2] + A[10] * A[13] * A[3] + A[15] * A[2] * A[9] - A[14] * A[3] * A[9];
  AInv[2] = -A[15] * A[2] * A[5] + A[14] * A[3] * A[5] + A[1] * A[15] * A[6] - A[13] * A[3] * A[6] - A[1] * A[14] * A[7] + A[13] * A[2] * A[7];
  AInv[3] = A[11] * A[2] * A[5] - A[10] * A[3] * A[5] - A[1] * A[11] * A[6] + A[1] * A[10] * A[7] + A[3] * A[6] * A[9] - A[2] * A[7] * A[9];
  AInv[4] = A[11] * A[14] * A[4] - A[10] * A[15] * A[4] - A[11] * A[12] * A[6] + A[10] * A[12] * A[7] + A[15] * A[6] * A[8] - A[14] * A[7] * A[8];
  AInv[5] = -A[0] * A[11] * A[14] + A[0] * A[10] * A[15] + A[11] * A[12] * A[2] - A[10] * A[12] * A[3] - A[15] * A[2] * A[8] + A[14] * A[3] * A[8];
  AInv[6] = A[15] * A[2] * A[4] - A[14] * A[3] * A[4] - A[0] * A[15] * A[6] + A[12] * A[3] * A[6] + A[0] * A[14] * A[7] - A[12] * A[2] * A[7];
  AInv[7] = -A[11] * A[2] * A[4] + A[10] * A[3] * A[4] + A[0] * A[11] * A[6] - A[0] * A[10] * A[7] - A[3] * A[6] * A[8] + A[2] * A[7] * A[8];
  AInv[8] = -A[11] * A[13] * A[4] + A[11] * A[12] * A[5] - A[15] * A[5] * A[8] + A[13] * A[7] * A[8] + A[15] * A[4] * A[9] - A[12] * A[7] * A[9];
  AInv[9] = -A[1] * A[11] * A[12] + A[0] * A[11] * A[13] + A[1] * A[15] * A[8] - A[13] * A[3] * A[8] - A[0] * A[15] * A[9] + A[12] * A[3] * A[9];
  AInv[10] = -A[1] * A[15] * A[4] + A[13] * A[3] * A[4] + A[0] * A[15] * A[5] - A[12] * A[3] * A[5] + A[1] * A[12] * A[7] - A[0] * A[13] * A[7];
  AInv[11] = A[1] * A[11] * A[4] - A[0] * A[11] * A[5] + A[3] * A[5] * A[8] - A[1] * A[7] * A[8] - A[3] * A[4] * A[9] + A[0] * A[7] * A[9];
  AInv[12] = A[10] * A[13] * A[4] - A[10] * A[12] * A[5] + A[14] * A[5] * A[8] - A[13] * A[6] * A[8] - A[14] * A[4] * A[9] + A[12] * A[6] * A[9];
  AInv[13] = A[1] * A[10] * A[12] - A[0] * A[10] * A[13] - A[1] * A[14] * A[8] + A[13] * A[2] * A[8] + A[0] * A[14] * A[9] - A[12] * A[2] * A[9];
  AInv[14] = A[1] * A[14] * A[4] - A[13] * A[2] * A[4] - A[0] * A[14] * A[5] + A[12] * A[2] * A[5] - A[1] * A[12] * A[6] + A[0] * A[13] * A[6];
  AInv[15] = -A[1] * A[10] * A[4] + A[0] * A[10] * A[5] - A[2] * A[5] * A[8] + A[1] * A[6] * A[8] + A[2] * A[4] * A[9] - A[0] * A[6] * A[9];
  double invDet = 1.0 / (A[0] * AInv[0] + A[1] * AInv[4] + A[2] * AInv[8] + A[3] * AInv[12]);
  for (int i = 0; i < 16; i++) AInv[i] *= invDet;
}

/* TetMesh::getTetVolume, VEGA/volumetricMesh/tetMesh.cpp:184-188 with Vec3d dot/cross
 * (VEGA/minivector/vec3d.h:206-218): 1/6 * | (a-d) . ((b-d) x (c-d)) | */
static double tet_volume(const double *a, const double *b, const double *c, const double *d) {
  double p[3], u[3], v[3], w[3];
  for (int i = 0; i < 3; i++) { p[i] = a[i] - d[i]; u[i] = b[i] - d[i]; v[i] = c[i] - d[i]; }
  w[0] = u[1] * v[2] - v[1] * u[2];
  w[1] = -u[0] * v[2] + v[0] * u[2];
  w[2] = u[0] * v[1] - v[0] * u[1];
  double dt = p[0] * w[0] + p[1] * w[1] + p[2] * w[2];
  return 1.0 / 6 * fabs(dt);
}

/* ---- PolarDecomposition, VEGA/polarDecomposition/polarDecomposition.cpp:8-108 ---- */
static double one_norm(const double *A) {
  double norm = 0.0;
  for (int i = 0; i < 3; i++) {
    double s = fabs(A[i + 0]) + fabs(A[i + 3]) + fabs(A[i + 6]);
    if (s > norm) norm = s;
  }
  return norm;
}
static double inf_norm(const double *A) {
  double norm = 0.0;
  for (int i = 0; i < 3; i++) {
    double s = fabs(A[3 * i + 0]) + fabs(A[3 * i + 1]) + fabs(A[3 * i + 2]);
    if (s > norm) norm = s;
  }
  return norm;
}
static void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static double polar_compute(const double *M, double *Q, double *S, double tol) {
  double Mk[9], Ek[9], det, M1, Minf, E1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Mk[3 * i + j] = M[3 * j + i];
  M1 = one_norm(Mk);
  Minf = inf_norm(Mk);
  do {
    double Adj[9];
    cross3(&Mk[3], &Mk[6], &Adj[0]);
    cross3(&Mk[6], &Mk[0], &Adj[3]);
    cross3(&Mk[0], &Mk[3], &Adj[6]);
    det = Mk[0] * Adj[0] + Mk[1] * Adj[1] + Mk[2] * Adj[2];
    if (det == 0.0) break; /* reference prints a warning here (polarDecomposition.cpp:63-67) */
    double A1 = one_norm(Adj), Ainf = inf_norm(Adj);
    double gamma = sqrt(sqrt((A1 * Ainf) / (M1 * Minf)) / fabs(det));
    double g1 = gamma * 0.5;
    double g2 = 0.5 / (gamma * det);
    for (int i = 0; i < 9; i++) {
      Ek[i] = Mk[i];
      Mk[i] = g1 * Mk[i] + g2 * Adj[i];
      Ek[i] -= Mk[i];
    }
    E1 = one_norm(Ek);
    M1 = one_norm(Mk);
    Minf = inf_norm(Mk);
  } while (E1 > M1 * tol);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) Q[3 * i + j] = Mk[3 * j + i];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      S[3 * i + j] = 0;
      for (int k = 0; k < 3; k++) S[3 * i + j] += Mk[3 * i + k] * M[3 * k + j];
    }
  for (int i = 0; i < 3; i++)
    for (int j = i; j < 3; j++) S[3 * i + j] = S[3 * j + i] = 0.5 * (S[3 * i + j] + S[3 * j + i]);
  return det;
}

/* ---- setup ---- */

/* Block structure of GetStiffnessMatrixTopology (corotationalLinearFEM.cpp:163-186): vertex pair
 * (v_i, v_j) for every i,j of every tet; std::map rows => sorted unique columns. */
static void build_adjacency(Port *s) {
  size_t np = 16 * (size_t)s->nT;
  uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (np ? np : 1));
  size_t k = 0;
  for (int el = 0; el < s->nT; el++)
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++)
        keys[k++] = ((uint64_t)(uint32_t)s->tets[4 * el + i] << 32) | (uint32_t)s->tets[4 * el + j];
  qsort(keys, np, sizeof(uint64_t), cmp_u64);
  size_t nu = 0;
  for (size_t i = 0; i < np; i++)
    if (i == 0 || keys[i] != keys[i - 1]) keys[nu++] = keys[i];
  s->bp = (int *)calloc((size_t)s->nV + 1, sizeof(int));
  s->bc = (int *)malloc(sizeof(int) * (nu ? nu : 1));
  for (size_t i = 0; i < nu; i++) {
    s->bp[(keys[i] >> 32) + 1]++;
    s->bc[i] = (int)(keys[i] & 0xffffffffu);
  }
  for (int v = 0; v < s->nV; v++) s->bp[v + 1] += s->bp[v];
  free(keys);
}

static void build_K_structure(Port *s) {
  int nB = s->bp[s->nV];
  s->nnzK = 9 * nB;
  s->Kia = (int *)malloc(sizeof(int) * ((size_t)s->r + 1));
  s->Kja = (int *)malloc(sizeof(int) * (size_t)(s->nnzK ? s->nnzK : 1));
  s->Ka = (double *)calloc((size_t)(s->nnzK ? s->nnzK : 1), sizeof(double));
  s->Da = (double *)calloc((size_t)(s->nnzK ? s->nnzK : 1), sizeof(double));
  int cnt = 0;
  for (int v = 0; v < s->nV; v++)
    for (int k = 0; k < 3; k++) {
      s->Kia[3 * v + k] = cnt;
      for (int p = s->bp[v]; p < s->bp[v + 1]; p++)
        for (int l = 0; l < 3; l++) s->Kja[cnt++] = 3 * s->bc[p] + l;
    }
  s->Kia[s->r] = cnt;
  /* BuildRowColumnIndices, corotationalLinearFEM.cpp:482-502 */
  s->rowIdx = (int *)malloc(sizeof(int) * 4 * (size_t)s->nT);
  s->colIdx = (int *)malloc(sizeof(int) * 16 * (size_t)s->nT);
  for (int el = 0; el < s->nT; el++) {
    for (int i = 0; i < 4; i++) s->rowIdx[4 * el + i] = s->tets[4 * el + i];
    for (int i = 0; i < 4; i++) {
      int vi = s->tets[4 * el + i];
      for (int j = 0; j < 4; j++)
        s->colIdx[16 * el + 4 * i + j] = find_pos(s->bc + s->bp[vi], s->bp[vi + 1] - s->bp[vi], s->tets[4 * el + j]);
    }
  }
}

/* CorotationalLinearFEM ctor, corotationalLinearFEM.cpp:40-146 (MInverse and K0 = V * B^T E B) */
static void build_element_data(Port *s, double E_, double nu_) {
  s->lambda = (double *)malloc(sizeof(double) * (size_t)s->nT);
  s->mu = (double *)malloc(sizeof(double) * (size_t)s->nT);
  s->MInv = (double *)malloc(sizeof(double) * 16 * (size_t)s->nT);
  s->K0 = (double *)malloc(sizeof(double) * 144 * (size_t)s->nT);
  for (int el = 0; el < s->nT; el++) {
    /* ENuMaterial::getLambda/getMu, VEGA/volumetricMesh/volumetricMeshENuMaterial.h:61-62 */
    if (s->matE) E_ = s->matE[el];
    if (s->matNu) nu_ = s->matNu[el];
    s->lambda[el] = (nu_ * E_) / ((1 + nu_) * (1 - 2 * nu_));
    s->mu[el] = E_ / (2 * (1 + nu_));
    const int *vt = s->tets + 4 * el;
    double M[16];
    for (int vtx = 0; vtx < 4; vtx++)
      for (int dim = 0; dim < 3; dim++) M[4 * dim + vtx] = s->x0[3 * vt[vtx] + dim];
    M[12] = M[13] = M[14] = M[15] = 1.0;
    double *MInv = s->MInv + 16 * (size_t)el;
    inverse4x4(M, MInv);

    double B[72] = {MInv[0], 0, 0, MInv[4], 0, 0, MInv[8], 0, 0, MInv[12], 0, 0,
                    0, MInv[1], 0, 0, MInv[5], 0, 0, MInv[9], 0, 0, MInv[13], 0,
                    0, 0, MInv[2], 0, 0, MInv[6], 0, 0, MInv[10], 0, 0, MInv[14],
                    MInv[1], MInv[0], 0, MInv[5], MInv[4], 0, MInv[9], MInv[8], 0, MInv[13], MInv[12], 0,
                    0, MInv[2], MInv[1], 0, MInv[6], MInv[5], 0, MInv[10], MInv[9], 0, MInv[14], MInv[13],
                    MInv[2], 0, MInv[0], MInv[6], 0, MInv[4], MInv[10], 0, MInv[8], MInv[14], 0, MInv[12]};
    double lambda = s->lambda[el], mu = s->mu[el];
    double Em[36] = {lambda + 2 * mu, lambda, lambda, 0, 0, 0,
                     lambda, lambda + 2 * mu, lambda, 0, 0, 0,
                     lambda, lambda, lambda + 2 * mu, 0, 0, 0,
                     0, 0, 0, mu, 0, 0,
                     0, 0, 0, 0, mu, 0,
                     0, 0, 0, 0, 0, mu};
    double EB[72];
    memset(EB, 0, sizeof(EB));
    for (int i = 0; i < 6; i++)
      for (int j = 0; j < 12; j++)
        for (int k = 0; k < 6; k++) EB[12 * i + j] += Em[6 * i + k] * B[12 * k + j];
    double *K = s->K0 + 144 * (size_t)el;
    memset(K, 0, sizeof(double) * 144);
    for (int i = 0; i < 12; i++)
      for (int j = 0; j < 12; j++)
        for (int k = 0; k < 6; k++) K[12 * i + j] += B[12 * k + i] * EB[12 * k + j];
    double volume = tet_volume(s->x0 + 3 * vt[0], s->x0 + 3 * vt[1], s->x0 + 3 * vt[2], s->x0 + 3 * vt[3]);
    for (int i = 0; i < 144; i++) K[i] *= volume;
  }
}

/* GenerateMassMatrix::computeMassMatrix(inflate3Dim=true), VEGA/volumetricMesh/generateMassMatrix.cpp:33-76;
 * TetMesh::computeElementMassMatrix, VEGA/volumetricMesh/tetMesh.cpp:150-182 */
static void build_mass(Port *s) {
  int nB = s->bp[s->nV];
  double *mb = (double *)calloc((size_t)(nB ? nB : 1), sizeof(double)); /* one scalar per block */
  static const double mtx[16] = {2, 1, 1, 1, 1, 2, 1, 1, 1, 1, 2, 1, 1, 1, 1, 2};
  for (int el = 0; el < s->nT; el++) {
    const int *vt = s->tets + 4 * el;
    double vol = tet_volume(s->x0 + 3 * vt[0], s->x0 + 3 * vt[1], s->x0 + 3 * vt[2], s->x0 + 3 * vt[3]);
    double factor = (s->matRho ? s->matRho[el] : s->rho) * vol / 20;
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        double entry = factor * mtx[4 * j + i];
        int p = s->bp[vt[i]] + s->colIdx[16 * el + 4 * i + j];
        mb[p] += entry; /* outline AddEntry: first insert, then += (sparseMatrix.cpp:128-138) */
      }
  }
  s->nnzM = 3 * nB;
  s->Mia = (int *)malloc(sizeof(int) * ((size_t)s->r + 1));
  s->Mja = (int *)malloc(sizeof(int) * (size_t)(s->nnzM ? s->nnzM : 1));
  s->Ma = (double *)malloc(sizeof(double) * (size_t)(s->nnzM ? s->nnzM : 1));
  s->subIdx = (int *)malloc(sizeof(int) * (size_t)(s->nnzM ? s->nnzM : 1));
  int cnt = 0;
  for (int v = 0; v < s->nV; v++)
    for (int c = 0; c < 3; c++) {
      s->Mia[3 * v + c] = cnt;
      for (int p = s->bp[v]; p < s->bp[v + 1]; p++) {
        s->Mja[cnt] = 3 * s->bc[p] + c;
        s->Ma[cnt] = mb[p];
        /* BuildSubMatrixIndices (sparseMatrix.cpp:1004-1047): position of that column in K's row */
        s->subIdx[cnt] = 3 * (p - s->bp[v]) + c;
        cnt++;
      }
    }
  s->Mia[s->r] = cnt;
  free(mb);
}

/* systemMatrix = copy(K).RemoveRowsColumns(fixed) + BuildSuperMatrixIndices
 * (VEGA/integrator/implicitNewmarkSparse.cpp:69-71; sparseMatrix.cpp:1296-1357, 896-991) */
static void build_system(Port *s) {
  int *oldToNew = (int *)malloc(sizeof(int) * (size_t)(s->r ? s->r : 1));
  int dof = 0, cntd = 0;
  for (int i = 0; i < s->nC; i++) {
    while (dof < s->cdofs[i]) oldToNew[dof++] = cntd++;
    oldToNew[dof++] = -1;
  }
  while (dof < s->r) oldToNew[dof++] = cntd++;
  s->nS = s->r - s->nC;
  s->Sia = (int *)malloc(sizeof(int) * ((size_t)s->nS + 1));
  s->superRows = (int *)malloc(sizeof(int) * (size_t)(s->nS ? s->nS : 1));
  size_t cap = (size_t)(s->nnzK ? s->nnzK : 1);
  s->Sja = (int *)malloc(sizeof(int) * cap);
  s->superIdx = (int *)malloc(sizeof(int) * cap);
  int nnz = 0, row = 0;
  for (int i = 0; i < s->r; i++) {
    if (oldToNew[i] < 0) continue;
    s->Sia[row] = nnz;
    s->superRows[row] = i;
    for (int p = s->Kia[i]; p < s->Kia[i + 1]; p++) {
      int nc = oldToNew[s->Kja[p]];
      if (nc < 0) continue;
      s->Sja[nnz] = nc;
      s->superIdx[nnz] = p - s->Kia[i];
      nnz++;
    }
    row++;
  }
  s->Sia[s->nS] = nnz;
  s->nnzS = nnz;
  s->Sa = (double *)calloc((size_t)(nnz ? nnz : 1), sizeof(double));
  /* BuildDiagonalIndices, sparseMatrix.cpp:622-630 */
  s->diagIdx = (int *)malloc(sizeof(int) * (size_t)(s->nS ? s->nS : 1));
  for (int i = 0; i < s->nS; i++) s->diagIdx[i] = find_pos(s->Sja + s->Sia[i], s->Sia[i + 1] - s->Sia[i], i);
  free(oldToNew);
}

static double *dalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

static void *port_create(int nV, const double *verts, int nT, const int *tets, double E, double nu, double rho,
                         const double *Ee, const double *nue, const double *rhoe, int nFixedVerts, const int *fixedVerts,
                         double h, double dampM, double dampK);

void *fbport_create(int nV, const double *verts, int nT, const int *tets, double E, double nu,
                    double rho, int nFixedVerts, const int *fixedVerts, double h, double dampM,
                    double dampK) {
  return port_create(nV, verts, nT, tets, E, nu, rho, NULL, NULL, NULL, nFixedVerts, fixedVerts, h, dampM, dampK);
}

/* per-element ENU materials (what a .veg with several *MATERIAL / *REGION sections yields, volumetricMesh.cpp:45-535;
 * arrays are only read during construction) */
void *fbport_create_materials(int nV, const double *verts, int nT, const int *tets, const double *Ee, const double *nue,
                              const double *rhoe, int nFixedVerts, const int *fixedVerts, double h, double dampM, double dampK) {
  return port_create(nV, verts, nT, tets, 0, 0, 0, Ee, nue, rhoe, nFixedVerts, fixedVerts, h, dampM, dampK);
}

static void *port_create(int nV, const double *verts, int nT, const int *tets, double E, double nu, double rho,
                         const double *Ee, const double *nue, const double *rhoe, int nFixedVerts, const int *fixedVerts,
                         double h, double dampM, double dampK) {
  Port *s = (Port *)calloc(1, sizeof(Port));
  s->matE = Ee; s->matNu = nue; s->matRho = rhoe;
  s->nV = nV; s->nT = nT; s->r = 3 * nV;
  s->rho = rho; s->h = h; s->dampM = dampM; s->dampK = dampK;
  s->x0 = (double *)malloc(sizeof(double) * 3 * (size_t)(nV ? nV : 1));
  memcpy(s->x0, verts, sizeof(double) * 3 * (size_t)nV);
  s->tets = (int *)malloc(sizeof(int) * 4 * (size_t)(nT ? nT : 1));
  memcpy(s->tets, tets, sizeof(int) * 4 * (size_t)nT);
  build_adjacency(s);
  build_K_structure(s);
  build_element_data(s, E, nu);
  build_mass(s);
  /* Deformable::FixedVerticesToFixedDOF, deformable/Deformable.cpp:294-314 */
  int *fv = (int *)malloc(sizeof(int) * (size_t)(nFixedVerts ? nFixedVerts : 1));
  memcpy(fv, fixedVerts, sizeof(int) * (size_t)nFixedVerts);
  qsort(fv, (size_t)nFixedVerts, sizeof(int), cmp_int);
  s->nC = 3 * nFixedVerts;
  s->cdofs = (int *)malloc(sizeof(int) * (size_t)(s->nC ? s->nC : 1));
  for (int i = 0; i < nFixedVerts; i++)
    for (int c = 0; c < 3; c++) s->cdofs[3 * i + c] = 3 * fv[i] + c;
  free(fv);
  build_system(s);
  size_t r = (size_t)s->r;
  s->q = dalloc(r); s->qvel = dalloc(r); s->qaccel = dalloc(r);
  s->q1 = dalloc(r); s->qvel1 = dalloc(r); s->qaccel1 = dalloc(r);
  s->fext = dalloc(r); s->fint = dalloc(r); s->qres = dalloc(r); s->qdelta = dalloc(r);
  s->buffer = dalloc(r); s->bufC = dalloc((size_t)s->nS);
  s->cr = dalloc((size_t)s->nS); s->cd = dalloc((size_t)s->nS); s->cq = dalloc((size_t)s->nS);
  s->invD = dalloc((size_t)s->nS);
  s->matE = s->matNu = s->matRho = NULL;
  return s;
}

void fbport_destroy(void *p) {
  Port *s = (Port *)p;
  if (!s) return;
  free(s->x0); free(s->tets); free(s->lambda); free(s->mu); free(s->MInv); free(s->K0);
  free(s->bp); free(s->bc); free(s->Kia); free(s->Kja); free(s->Ka); free(s->Da);
  free(s->rowIdx); free(s->colIdx); free(s->Mia); free(s->Mja); free(s->subIdx); free(s->Ma);
  free(s->cdofs); free(s->Sia); free(s->Sja); free(s->superRows); free(s->superIdx);
  free(s->diagIdx); free(s->Sa);
  free(s->q); free(s->qvel); free(s->qaccel); free(s->q1); free(s->qvel1); free(s->qaccel1);
  free(s->fext); free(s->fint); free(s->qres); free(s->qdelta); free(s->buffer); free(s->bufC);
  free(s->cr); free(s->cd); free(s->cq); free(s->invD);
  free(s);
}

int fbport_r(void *p) { return ((Port *)p)->r; }
int fbport_nnz_K(void *p) { return ((Port *)p)->nnzK; }
int fbport_nnz_M(void *p) { return ((Port *)p)->nnzM; }
int fbport_rows_sys(void *p) { return ((Port *)p)->nS; }
int fbport_nnz_sys(void *p) { return ((Port *)p)->nnzS; }

static void copy_csr(int n, int nnz, const int *ia, const int *ja, const double *a, int *oia, int *oja, double *oa) {
  if (oia) memcpy(oia, ia, sizeof(int) * ((size_t)n + 1));
  if (oja) memcpy(oja, ja, sizeof(int) * (size_t)nnz);
  if (oa) memcpy(oa, a, sizeof(double) * (size_t)nnz);
}
void fbport_K_csr(void *p, int *ia, int *ja, double *a) { Port *s = (Port *)p; copy_csr(s->r, s->nnzK, s->Kia, s->Kja, s->Ka, ia, ja, a); }
void fbport_M_csr(void *p, int *ia, int *ja, double *a) { Port *s = (Port *)p; copy_csr(s->r, s->nnzM, s->Mia, s->Mja, s->Ma, ia, ja, a); }
void fbport_sys_csr(void *p, int *ia, int *ja, double *a) { Port *s = (Port *)p; copy_csr(s->nS, s->nnzS, s->Sia, s->Sja, s->Sa, ia, ja, a); }

void fbport_element_maps(void *p, int *rowIdx4, int *colIdx16) {
  Port *s = (Port *)p;
  memcpy(rowIdx4, s->rowIdx, sizeof(int) * 4 * (size_t)s->nT);
  memcpy(colIdx16, s->colIdx, sizeof(int) * 16 * (size_t)s->nT);
}
void fbport_element_data(void *p, double *MInv16, double *K0_144) {
  Port *s = (Port *)p;
  if (MInv16) memcpy(MInv16, s->MInv, sizeof(double) * 16 * (size_t)s->nT);
  if (K0_144) memcpy(K0_144, s->K0, sizeof(double) * 144 * (size_t)s->nT);
}
void fbport_super_maps(void *p, int *superRows, int *superIdx) {
  Port *s = (Port *)p;
  memcpy(superRows, s->superRows, sizeof(int) * (size_t)s->nS);
  memcpy(superIdx, s->superIdx, sizeof(int) * (size_t)s->nnzS);
}
void fbport_submatrix_map(void *p, int *idx) {
  Port *s = (Port *)p;
  memcpy(idx, s->subIdx, sizeof(int) * (size_t)s->nnzM);
}

/* ---- sparse primitives on flat CSR (VEGA/sparseMatrix/sparseMatrix.cpp:405-413, 509-568, 993-1002, 1088-1098) ---- */
static void spmv(int n, const int *ia, const int *ja, const double *a, const double *x, double *y) {
  for (int i = 0; i < n; i++) {
    y[i] = 0;
    for (int p = ia[i]; p < ia[i + 1]; p++) y[i] += x[ja[p]] * a[p];
  }
}
static void add_submatrix_M(Port *s, double *dst, double factor) {
  for (int i = 0; i < s->r; i++)
    for (int p = s->Mia[i]; p < s->Mia[i + 1]; p++) dst[s->Kia[i] + s->subIdx[p]] += factor * s->Ma[p];
}

/* CorotationalLinearFEM::ComputeForceAndStiffnessMatrixOfSubmesh, warp = 1
 * (corotationalLinearFEM.cpp:219-293, 456-468; WarpMatrix :191-211) */
static void compute_force_and_matrix(Port *s, const double *u, double *f, double *Ka) {
  if (f) memset(f, 0, sizeof(double) * (size_t)s->r);
  if (Ka) memset(Ka, 0, sizeof(double) * (size_t)s->nnzK);
  for (int el = 0; el < s->nT; el++) {
    const int *vt = s->tets + 4 * el;
    const double *MInv = s->MInv + 16 * (size_t)el;
    const double *K = s->K0 + 144 * (size_t)el;
    double KElement[144], P[16], F[9], R[9], S[9], RK[144], fElement[12];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 4; j++) P[4 * i + j] = s->x0[3 * vt[j] + i] + u[3 * vt[j] + i];
    for (int j = 0; j < 4; j++) P[12 + j] = 1;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        F[3 * i + j] = 0;
        for (int k = 0; k < 4; k++) F[3 * i + j] += P[4 * i + k] * MInv[4 * k + j];
      }
    double det = polar_compute(F, R, S, 1E-6);
    if (det < 0)
      for (int i = 0; i < 9; i++) R[i] *= -1.0;
    memset(RK, 0, sizeof(RK));
    memset(KElement, 0, sizeof(KElement));
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) {
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++)
            for (int m = 0; m < 3; m++)
              RK[12 * (3 * i + k) + (3 * j + l)] += R[3 * k + m] * K[12 * (3 * i + m) + (3 * j + l)];
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++)
            for (int m = 0; m < 3; m++)
              KElement[12 * (3 * i + k) + (3 * j + l)] += RK[12 * (3 * i + k) + (3 * j + m)] * R[3 * l + m];
      }
    for (int i = 0; i < 12; i++) {
      fElement[i] = 0;
      for (int j = 0; j < 4; j++)
        for (int l = 0; l < 3; l++)
          fElement[i] += KElement[12 * i + 3 * j + l] * P[4 * l + j] - RK[12 * i + 3 * j + l] * s->x0[3 * vt[j] + l];
    }
    if (f)
      for (int j = 0; j < 4; j++)
        for (int l = 0; l < 3; l++) f[3 * vt[j] + l] += fElement[3 * j + l];
    if (Ka) {
      const int *rowIndex = s->rowIdx + 4 * el;
      const int *columnIndex = s->colIdx + 16 * el;
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
          for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++)
              Ka[s->Kia[3 * rowIndex[i] + k] + 3 * columnIndex[4 * i + j] + l] += KElement[12 * (3 * i + k) + 3 * j + l];
    }
  }
}

void fbport_force_and_matrix(void *p, const double *u, double *f, double *Ka) {
  Port *s = (Port *)p;
  compute_force_and_matrix(s, u, f, s->Ka);
  if (Ka) memcpy(Ka, s->Ka, sizeof(double) * (size_t)s->nnzK);
}

void fbport_set_state(void *p, const double *q, const double *qvel) {
  Port *s = (Port *)p;
  memcpy(s->q, q, sizeof(double) * (size_t)s->r);
  if (qvel) memcpy(s->qvel, qvel, sizeof(double) * (size_t)s->r);
}
void fbport_get_state(void *p, double *q, double *qvel, double *qaccel) {
  Port *s = (Port *)p;
  if (q) memcpy(q, s->q, sizeof(double) * (size_t)s->r);
  if (qvel) memcpy(qvel, s->qvel, sizeof(double) * (size_t)s->r);
  if (qaccel) memcpy(qaccel, s->qaccel, sizeof(double) * (size_t)s->r);
}
void fbport_set_external_forces(void *p, const double *f) {
  Port *s = (Port *)p;
  memcpy(s->fext, f, sizeof(double) * (size_t)s->r);
}

/* CGSolver::SolveLinearSystemWithJacobiPreconditioner, VEGA/sparseSolver/CGSolver.cpp:129-190 */
static int pcg_jacobi(Port *s, double *x, const double *b, double eps, int maxIterations) {
  int n = s->nS;
  double *r = s->cr, *d = s->cd, *q = s->cq, *invD = s->invD;
  for (int i = 0; i < n; i++) invD[i] = 1.0 / s->Sa[s->Sia[i] + s->diagIdx[i]];
  int iteration = 1;
  spmv(n, s->Sia, s->Sja, s->Sa, x, r);
  for (int i = 0; i < n; i++) { r[i] = b[i] - r[i]; d[i] = invD[i] * r[i]; }
  double res2 = 0;
  for (int i = 0; i < n; i++) res2 += r[i] * r[i] * invD[i];
  double res20 = res2;
  while ((res2 > eps * eps * res20) && (iteration <= maxIterations)) {
    spmv(n, s->Sia, s->Sja, s->Sa, d, q);
    double dDotq = 0;
    for (int i = 0; i < n; i++) dDotq += d[i] * q[i];
    double alpha = res2 / dDotq;
    for (int i = 0; i < n; i++) x[i] += alpha * d[i];
    if (iteration % 30 == 0) {
      spmv(n, s->Sia, s->Sja, s->Sa, x, r);
      for (int i = 0; i < n; i++) r[i] = b[i] - r[i];
    } else {
      for (int i = 0; i < n; i++) r[i] = r[i] - alpha * q[i];
    }
    double old = res2;
    res2 = 0;
    for (int i = 0; i < n; i++) res2 += r[i] * r[i] * invD[i];
    double beta = res2 / old;
    for (int i = 0; i < n; i++) d[i] = invD[i] * r[i] + beta * d[i];
    iteration++;
  }
  return (iteration - 1) * ((res2 > eps * eps * res20) ? -1 : 1);
}

/* RemoveRows / InsertRows, VEGA/insertRows/insertRows.cpp:65-109, 29-63 */
static void remove_rows(int m, double *xc, const double *x, int nf, const int *fixed) {
  int n = 0, row = 0;
  for (int i = 0; i < nf; i++) {
    while (row < fixed[i]) xc[n++] = x[row++];
    row++;
  }
  while (row < m) xc[n++] = x[row++];
}
static void insert_rows(int m, const double *xc, double *x, int nf, const int *fixed) {
  int dst = 0, src = 0;
  for (int i = 0; i < nf; i++) {
    while (dst < fixed[i]) x[dst++] = xc[src++];
    x[dst++] = 0.0;
  }
  while (dst < m) x[dst++] = xc[src++];
}

/* VolumeConservingIntegrator::DoTimestep, deformable/PS_VolumeConservingIntegrator.cpp:46-260
 * (maxIterations = 1, dynamic solver, PCG branch, internalForceScalingFactor = 1). */
int fbport_do_timestep(void *p) {
  Port *s = (Port *)p;
  int r = s->r;
  const double internalForceScalingFactor = 1.0, epsilon = 1E-6;
  const int maxIterations = 1;
  int numIter = 0;
  double error0 = 0, errorQuotient;
  for (int i = 0; i < r; i++) {
    s->qaccel1[i] = s->qaccel[i] = 0;
    s->q1[i] = s->q[i];
    s->qvel1[i] = s->qvel[i];
  }
  do {
    double t0 = now_s();
    compute_force_and_matrix(s, s->q, s->fint, s->Ka);
    s->tAsm = now_s() - t0;
    for (int i = 0; i < r; i++) s->fint[i] *= internalForceScalingFactor;
    for (int k = 0; k < s->nnzK; k++) s->Ka[k] *= internalForceScalingFactor;
    memset(s->qres, 0, sizeof(double) * (size_t)r);
    /* D = dampK * K ; D += dampM * M */
    for (int k = 0; k < s->nnzK; k++) s->Da[k] = s->Ka[k] * s->dampK;
    add_submatrix_M(s, s->Da, s->dampM);
    /* K *= h ; K += D ; (K += 1.0 * empty damping matrix) ; qres = K qvel ; K *= h ; K += M */
    for (int k = 0; k < s->nnzK; k++) s->Ka[k] *= s->h;
    for (int k = 0; k < s->nnzK; k++) s->Ka[k] += s->Da[k];
    spmv(r, s->Kia, s->Kja, s->Ka, s->qvel, s->qres);
    for (int k = 0; k < s->nnzK; k++) s->Ka[k] *= s->h;
    add_submatrix_M(s, s->Ka, 1.0);
    for (int i = 0; i < r; i++) {
      s->qres[i] += s->fint[i] - s->fext[i];
      s->qres[i] *= -s->h;
      s->qdelta[i] = s->qres[i];
    }
    double error = 0;
    for (int i = 0; i < r; i++) error += s->qres[i] * s->qres[i];
    if (numIter == 0) { error0 = error; errorQuotient = 1.0; }
    else errorQuotient = error / error0;
    if (errorQuotient < epsilon * epsilon) break;
    remove_rows(r, s->bufC, s->qdelta, s->nC, s->cdofs);
    /* AssignSuperMatrix, sparseMatrix.cpp:993-1002 */
    for (int i = 0; i < s->nS; i++) {
      const double *row = s->Ka + s->Kia[s->superRows[i]];
      for (int pp = s->Sia[i]; pp < s->Sia[i + 1]; pp++) s->Sa[pp] = row[s->superIdx[pp]];
    }
    t0 = now_s();
    memset(s->buffer, 0, sizeof(double) * (size_t)r);
    int info = pcg_jacobi(s, s->buffer, s->bufC, 1e-6, 10000);
    if (info > 0) info = 0;
    if (info != 0) return 1; /* the reference prints and exit(-1)s here (:203-209) */
    s->tSolve = now_s() - t0;
    insert_rows(r, s->buffer, s->qdelta, s->nC, s->cdofs);
    for (int i = 0; i < r; i++) {
      s->qvel[i] += s->qdelta[i];
      s->q[i] += s->h * s->qvel[i];
    }
    for (int i = 0; i < s->nC; i++) s->q[s->cdofs[i]] = s->qvel[s->cdofs[i]] = s->qaccel[s->cdofs[i]] = 0.0;
    numIter++;
  } while (numIter < maxIterations);
  return 0;
}

void fbport_K_values(void *p, double *a) { Port *s = (Port *)p; memcpy(a, s->Ka, sizeof(double) * (size_t)s->nnzK); }
void fbport_rhs(void *p, double *b) { Port *s = (Port *)p; memcpy(b, s->bufC, sizeof(double) * (size_t)s->nS); }
void fbport_internal_forces(void *p, double *f) { Port *s = (Port *)p; memcpy(f, s->fint, sizeof(double) * (size_t)s->r); }
void fbport_qdelta(void *p, double *d) { Port *s = (Port *)p; memcpy(d, s->qdelta, sizeof(double) * (size_t)s->r); }

int fbport_solve(void *p, const double *b, double *x, double eps, int maxIter) {
  Port *s = (Port *)p;
  memset(x, 0, sizeof(double) * (size_t)s->nS);
  return pcg_jacobi(s, x, b ? b : s->bufC, eps, maxIter);
}
int fbport_solve_iters(void *p, const double *b, double *x, int iters) { return fbport_solve(p, b, x, 0.0, iters); }
void fbport_assign_system(void *p) {
  Port *s = (Port *)p;
  for (int i = 0; i < s->nS; i++) {
    const double *row = s->Ka + s->Kia[s->superRows[i]];
    for (int pp = s->Sia[i]; pp < s->Sia[i + 1]; pp++) s->Sa[pp] = row[s->superIdx[pp]];
  }
}
void fbport_sys_spmv(void *p, const double *x, double *y) {
  Port *s = (Port *)p;
  spmv(s->nS, s->Sia, s->Sja, s->Sa, x, y);
}
double fbport_assembly_time(void *p) { return ((Port *)p)->tAsm; }
double fbport_solve_time(void *p) { return ((Port *)p)->tSolve; }
double fbport_polar(const double *F9, double *R9, double *S9, double tol) { return polar_compute(F9, R9, S9, tol); }

#include "deformable_port.inc"
