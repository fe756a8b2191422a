/*
 * ref_harness.cpp — C wrapper (fbref_*) around the UNMODIFIED reference classes.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).  This file contains no reference code: it
 * instantiates the reference's own classes exactly as Deformable::syncForceModel does
 * (src/deformable/Deformable.cpp:178-216) and forwards to their methods.  It is compiled by
 * oracle/Makefile together with the reference's sources, read in place from /root/reference,
 * into oracle/_ref/libfembrain_ref.so (git-ignored).
 *
 * C++98 on purpose: the reference headers do not compile as C++11 (mat3d.h:125,404).
 * `protected` is opened for this translation unit only, to read systemMatrix, bufferConstrained,
 * the CG solver and the element index maps; class layout does not depend on access specifiers.
 */
#define protected public
#define private public
/* order matters: the integrator/ copy of integratorBaseSparse.h (declares setConstrainedDOF) must
 * be seen before implicitNewmarkSparse.h, and the include/ copy of implicitNewmarkSparse.h must be
 * the one that is parsed so that "integratorSolverSelection.h" resolves to the PCG variant
 * (vegafem/include/integratorSolverSelection.h:40) and not the PARDISO one beside the sources. */
#include "../integrator/integratorBaseSparse.h"
#include "implicitNewmarkSparse.h"
#include "PS_VolumeConservingIntegrator.h"
#include "corotationalLinearFEM.h"
#include "corotationalLinearFEMMT.h"
#include "corotationalLinearFEMForceModel.h"
#include "generateMassMatrix.h"
#include "polarDecomposition.h"
#include "tetMesh.h"
#include "volumetricMeshENuMaterial.h"
#include "CGSolver.h"
#undef protected
#undef private

#include <algorithm>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <stdio.h>
#include <vector>

#define FBO_PREFIX fbref_
#include "oracle_api.h"

namespace {
struct RefSim {
  TetMesh *mesh;
  CorotationalLinearFEM *fem;
  CorotationalLinearFEMForceModel *forceModel;
  SparseMatrix *mass;
  VolumeConservingIntegrator *integrator;
  int nV, nT;
  std::vector<int> fixedDofs;
};

void csr_of(const SparseMatrix *m, int *ia, int *ja, double *a) {
  m->GenerateCompressedRowMajorFormat(a, ia, ja, 0, 0);
}
}  // namespace

extern "C" {

void *fbref_create(int nV, const double *verts, int nT, const int *tets, double E, double nu,
                   double rho, int nFixedVerts, const int *fixedVerts, double h, double dampM,
                   double dampK) {
  RefSim *s = new RefSim();
  s->nV = nV;
  s->nT = nT;
  std::vector<double> v(verts, verts + 3 * (size_t)nV);
  std::vector<int> e(tets, tets + 4 * (size_t)nT);
  s->mesh = new TetMesh(nV, &v[0], nT, &e[0], E, nu, rho);
  s->fem = new CorotationalLinearFEM(s->mesh);
  s->forceModel = new CorotationalLinearFEMForceModel(s->fem);
  GenerateMassMatrix::computeMassMatrix(s->mesh, &s->mass, true);
  /* Deformable::FixedVerticesToFixedDOF (Deformable.cpp:294-314) */
  std::vector<int> fv(fixedVerts, fixedVerts + nFixedVerts);
  std::sort(fv.begin(), fv.end());
  s->fixedDofs.resize(3 * fv.size());
  for (size_t i = 0; i < fv.size(); i++) {
    s->fixedDofs[3 * i + 0] = 3 * fv[i] + 0;
    s->fixedDofs[3 * i + 1] = 3 * fv[i] + 1;
    s->fixedDofs[3 * i + 2] = 3 * fv[i] + 2;
  }
  int dummy = 0;
  s->integrator = new VolumeConservingIntegrator(
      3 * nV, h, s->mass, s->forceModel, 0, (int)s->fixedDofs.size(),
      s->fixedDofs.empty() ? &dummy : &s->fixedDofs[0], dampM, dampK, 1, 1E-6, 1);
  return s;
}

/* The same chain, with the mesh and its materials read by the reference's own .veg loader
 * (TetMesh(char*), volumetricMesh.cpp:45-535). */
void *fbref_create_from_veg(const char *path, int nFixedVerts, const int *fixedVerts, double h, double dampM, double dampK) {
  RefSim *s = new RefSim();
  try {
    s->mesh = new TetMesh((char *)path);
  } catch (int) {
    delete s;
    return NULL;
  }
  s->nV = s->mesh->getNumVertices();
  s->nT = s->mesh->getNumElements();
  s->fem = new CorotationalLinearFEM(s->mesh);
  s->forceModel = new CorotationalLinearFEMForceModel(s->fem);
  GenerateMassMatrix::computeMassMatrix(s->mesh, &s->mass, true);
  std::vector<int> fv(fixedVerts, fixedVerts + nFixedVerts);
  std::sort(fv.begin(), fv.end());
  s->fixedDofs.resize(3 * fv.size());
  for (size_t i = 0; i < fv.size(); i++)
    for (int k = 0; k < 3; k++) s->fixedDofs[3 * i + k] = 3 * fv[i] + k;
  int dummy = 0;
  s->integrator = new VolumeConservingIntegrator(3 * s->nV, h, s->mass, s->forceModel, 0, (int)s->fixedDofs.size(),
                                                 s->fixedDofs.empty() ? &dummy : &s->fixedDofs[0], dampM, dampK, 1, 1E-6, 1);
  return s;
}

/* TetMesh::save -> VolumetricMesh::save (volumetricMesh.cpp:646-757): the reference's own .veg writer on the mesh it holds */
int fbref_save_veg(void *p, const char *path) { return ((RefSim *)p)->mesh->save((char *)path); }

int fbref_num_vertices(void *p) { return ((RefSim *)p)->nV; }
int fbref_num_tets(void *p) { return ((RefSim *)p)->nT; }
/* mesh as the reference holds it after loading: vertices, 0-based tets, per-element E / nu / density */
void fbref_mesh(void *p, double *verts, int *tets, double *E, double *nu, double *rho) {
  RefSim *s = (RefSim *)p;
  for (int i = 0; i < s->nV; i++)
    for (int k = 0; k < 3; k++) verts[3 * i + k] = (*s->mesh->getVertex(i))[k];
  for (int el = 0; el < s->nT; el++) {
    for (int k = 0; k < 4; k++) tets[4 * el + k] = s->mesh->getVertexIndex(el, k);
    VolumetricMesh::ENuMaterial *m = downcastENuMaterial(s->mesh->getElementMaterial(el));
    E[el] = m ? m->getE() : 0.0;
    nu[el] = m ? m->getNu() : 0.0;
    rho[el] = s->mesh->getElementDensity(el);
  }
}

void fbref_destroy(void *p) {
  RefSim *s = (RefSim *)p;
  if (!s) return;
  delete s->integrator;
  delete s->mass;
  delete s->forceModel;
  delete s->fem;
  delete s->mesh;
  delete s;
}

int fbref_r(void *p) { return 3 * ((RefSim *)p)->nV; }
int fbref_nnz_K(void *p) { return ((RefSim *)p)->integrator->tangentStiffnessMatrix->GetNumEntries(); }
int fbref_nnz_M(void *p) { return ((RefSim *)p)->mass->GetNumEntries(); }
int fbref_rows_sys(void *p) { return ((RefSim *)p)->integrator->systemMatrix->GetNumRows(); }
int fbref_nnz_sys(void *p) { return ((RefSim *)p)->integrator->systemMatrix->GetNumEntries(); }

void fbref_K_csr(void *p, int *ia, int *ja, double *a) {
  csr_of(((RefSim *)p)->integrator->tangentStiffnessMatrix, ia, ja, a);
}
void fbref_M_csr(void *p, int *ia, int *ja, double *a) { csr_of(((RefSim *)p)->mass, ia, ja, a); }
void fbref_sys_csr(void *p, int *ia, int *ja, double *a) {
  csr_of(((RefSim *)p)->integrator->systemMatrix, ia, ja, a);
}

void fbref_element_maps(void *p, int *rowIdx4, int *colIdx16) {
  RefSim *s = (RefSim *)p;
  for (int el = 0; el < s->nT; el++) {
    memcpy(rowIdx4 + 4 * (size_t)el, s->fem->rowIndices[el], 4 * sizeof(int));
    memcpy(colIdx16 + 16 * (size_t)el, s->fem->columnIndices[el], 16 * sizeof(int));
  }
}

void fbref_element_data(void *p, double *MInv16, double *K0_144) {
  RefSim *s = (RefSim *)p;
  for (int el = 0; el < s->nT; el++) {
    if (MInv16) memcpy(MInv16 + 16 * (size_t)el, s->fem->MInverse[el], 16 * sizeof(double));
    if (K0_144) memcpy(K0_144 + 144 * (size_t)el, s->fem->KElementUndeformed[el], 144 * sizeof(double));
  }
}

void fbref_super_maps(void *p, int *superRows, int *superIdx) {
  SparseMatrix *m = ((RefSim *)p)->integrator->systemMatrix;
  size_t k = 0;
  for (int i = 0; i < m->numRows; i++) {
    superRows[i] = m->superRows[i];
    for (int j = 0; j < m->rowLength[i]; j++) superIdx[k++] = m->superMatrixIndices[i][j];
  }
}

void fbref_submatrix_map(void *p, int *idx) {
  RefSim *s = (RefSim *)p;
  SparseMatrix *K = s->integrator->tangentStiffnessMatrix;
  size_t k = 0;
  for (int i = 0; i < K->numRows; i++)
    for (int j = 0; j < s->mass->rowLength[i]; j++) idx[k++] = K->subMatrixIndices[0][i][j];
}

void fbref_force_and_matrix(void *p, const double *u, double *f, double *Ka) {
  RefSim *s = (RefSim *)p;
  std::vector<double> uu(u, u + 3 * (size_t)s->nV);
  SparseMatrix *K = s->integrator->tangentStiffnessMatrix;
  s->forceModel->GetForceAndMatrix(&uu[0], f, K);
  if (Ka) K->GenerateCompressedRowMajorFormat(Ka, NULL, NULL, 0, 0);
}

// CorotationalLinearFEM::ComputeForceAndStiffnessMatrix with an explicit warp (0 linear, 1 corotational, 2 exact tangent),
// corotationalLinearFEM.cpp:214-449 — what CorotationalLinearFEMForceModel(fem, warp)::GetForceAndMatrix calls
void fbref_force_and_matrix_warp(void *p, const double *u, int warp, double *f, double *Ka) {
  RefSim *s = (RefSim *)p;
  std::vector<double> uu(u, u + 3 * (size_t)s->nV);
  SparseMatrix *K = s->integrator->tangentStiffnessMatrix;
  s->fem->ComputeForceAndStiffnessMatrix(&uu[0], f, K, warp);
  if (Ka) K->GenerateCompressedRowMajorFormat(Ka, NULL, NULL, 0, 0);
}

// The reference's stronger CPU assembly baseline (optional): CorotationalLinearFEMMT
// (corotationalLinearFEMMT.cpp:126-177) — numThreads pthreads over element ranges, one private force vector and stiffness
// matrix per thread, summed afterwards.  Wall-clock seconds per ComputeForceAndStiffnessMatrix (best of `reps`), and the
// largest |f_mt - f_single| as a sanity figure (the summation order differs from the single-threaded loop, so not bit-exact).
double fbref_mt_assembly_seconds(void *p, const double *u, int threads, int reps, double *maxForceDiff) {
  RefSim *s = (RefSim *)p;
  std::vector<double> uu(u, u + 3 * (size_t)s->nV), f(3 * (size_t)s->nV), f1(3 * (size_t)s->nV);
  CorotationalLinearFEMMT *mt = new CorotationalLinearFEMMT(s->mesh, threads);
  SparseMatrix *K;
  mt->GetStiffnessMatrixTopology(&K);
  double best = 1e300;
  for (int i = 0; i < reps; i++) {
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    mt->ComputeForceAndStiffnessMatrix(&uu[0], &f[0], K, 1);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    if (dt < best) best = dt;
  }
  if (maxForceDiff) {
    s->fem->ComputeForceAndStiffnessMatrix(&uu[0], &f1[0], NULL, 1);
    double m = 0;
    for (size_t i = 0; i < f.size(); i++) { double d = f[i] - f1[i]; if (d < 0) d = -d; if (d > m) m = d; }
    *maxForceDiff = m;
  }
  delete K;
  delete mt;
  return best;
}

void fbref_set_state(void *p, const double *q, const double *qvel) {
  ((RefSim *)p)->integrator->SetqState(q, qvel, NULL);
}
void fbref_get_state(void *p, double *q, double *qvel, double *qaccel) {
  ((RefSim *)p)->integrator->GetqState(q, qvel, qaccel);
}
void fbref_set_external_forces(void *p, const double *f) {
  RefSim *s = (RefSim *)p;
  std::vector<double> ff(f, f + 3 * (size_t)s->nV);
  s->integrator->SetExternalForces(&ff[0]);
}
int fbref_do_timestep(void *p) { return ((RefSim *)p)->integrator->DoTimestep(); }

void fbref_K_values(void *p, double *a) {
  ((RefSim *)p)->integrator->tangentStiffnessMatrix->GenerateCompressedRowMajorFormat(a, NULL, NULL, 0, 0);
}
void fbref_rhs(void *p, double *b) {
  RefSim *s = (RefSim *)p;
  memcpy(b, s->integrator->bufferConstrained, sizeof(double) * s->integrator->systemMatrix->GetNumRows());
}
void fbref_internal_forces(void *p, double *f) {
  RefSim *s = (RefSim *)p;
  memcpy(f, s->integrator->internalForces, sizeof(double) * 3 * s->nV);
}
void fbref_qdelta(void *p, double *d) {
  RefSim *s = (RefSim *)p;
  memcpy(d, s->integrator->qdelta, sizeof(double) * 3 * s->nV);
}

int fbref_solve(void *p, const double *b, double *x, double eps, int maxIter) {
  RefSim *s = (RefSim *)p;
  int n = s->integrator->systemMatrix->GetNumRows();
  memset(x, 0, sizeof(double) * n);
  return s->integrator->jacobiPreconditionedCGSolver->SolveLinearSystemWithJacobiPreconditioner(
      x, b ? b : s->integrator->bufferConstrained, eps, maxIter, 0);
}
int fbref_solve_iters(void *p, const double *b, double *x, int iters) {
  return fbref_solve(p, b, x, 0.0, iters);
}
void fbref_assign_system(void *p) {
  RefSim *s = (RefSim *)p;
  s->integrator->systemMatrix->AssignSuperMatrix(s->integrator->tangentStiffnessMatrix);
}
void fbref_sys_spmv(void *p, const double *x, double *y) {
  ((RefSim *)p)->integrator->systemMatrix->MultiplyVector(x, y);
}

double fbref_assembly_time(void *p) { return ((RefSim *)p)->integrator->GetForceAssemblyTime(); }
double fbref_solve_time(void *p) { return ((RefSim *)p)->integrator->GetSystemSolveTime(); }

double fbref_polar(const double *F9, double *R9, double *S9, double tol) {
  return PolarDecomposition::Compute(F9, R9, S9, tol);
}

}  // extern "C"
