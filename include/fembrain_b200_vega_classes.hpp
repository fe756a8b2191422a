// fembrain_b200_vega_classes.hpp — the C ABI behind the reference's OWN virtual interfaces.
//
// Two classes DERIVED from Vega's types, for hosts that hold `ForceModel *` / `IntegratorBaseSparse *` pointers and
// therefore need real subclasses (include/fembrain_b200_vega.hpp is the lighter, duck-typed alternative that needs no
// Vega header):
//
//   fembrain_b200::CudaCorotationalForceModel       : public ForceModel
//       (VEGA/forceModel/forceModel.h:42-67; replaces CorotationalLinearFEMForceModel,
//        VEGA/elasticForceModel/corotationalLinearFEMForceModel.cpp:30-56, warp = 1)
//   fembrain_b200::CudaVolumeConservingIntegrator2  : public IntegratorBaseSparse
//       (VEGA/integrator/integratorBaseSparse.h:45-90; replaces VolumeConservingIntegrator,
//        DEF/PS_VolumeConservingIntegrator.h:13-37 — SAME constructor signature)
//
// With them the edit in DEF/Deformable.{h,cpp} is: the member type `ImplicitNewmarkSparse *` -> `IntegratorBaseSparse *`,
// `new CorotationalLinearFEMForceModel(m_lpDeformable)` -> `new CudaCorotationalForceModel(nV, verts, nT, tets, E, nu, rho)`
// and `new VolumeConservingIntegrator(` -> `new CudaVolumeConservingIntegrator2(`, all other lines unchanged
// (tests/test_dropin.py compiles the reference's Deformable.cpp with exactly these substitutions).
//
// State lives where Vega keeps it: IntegratorBase's host arrays q, qvel, qaccel, externalForces (its SetqState / GetqState /
// SetExternalForces are NOT virtual and act on those arrays, VEGA/integrator/integratorBase.cpp:84-122).  DoTimestep
// therefore uploads state and forces, runs fb_step on the device, and downloads the new state — the copy semantics the
// reference has, with the device doing the work in between.
//
// Requires Vega's headers on the include path (sparseMatrix.h, forceModel.h, integratorBaseSparse.h).  C++98.
#ifndef FEMBRAIN_B200_VEGA_CLASSES_HPP
#define FEMBRAIN_B200_VEGA_CLASSES_HPP

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fembrain_b200.h"
#include "corotationalLinearFEM.h"
#include "forceModel.h"
#include "integratorBaseSparse.h"
#include "sparseMatrix.h"
#include "tetMesh.h"
#include "volumetricMeshENuMaterial.h"

namespace fembrain_b200 {

inline void vega_fail(const char *what, int st) {
  printf("Error: fembrain_b200 %s failed: %s (%s)\n", what, fb_status_string(st), fb_last_error_string());
#ifndef FEMBRAIN_B200_NO_EXIT
  exit(1);  // the reference's idiom for unusable inputs (implicitNewmarkSparse.cpp:52-56, sparseMatrix.cpp:906-917)
#else
  throw 1;  // ... or corotationalLinearFEM.cpp:61-64's
#endif
}

// The corotational linear FEM force model of one tet mesh, evaluated on the device.  Owns the fb_context (mesh, block
// structure of K, element data); the integrator below borrows it, exactly as VolumeConservingIntegrator borrows the
// reference's ForceModel (implicitNewmarkSparse.cpp:85-90 does not delete it).
class CudaCorotationalForceModel : public ForceModel {
 public:
  // mesh + material: the arguments of `new TetMesh(...)` in Deformable::syncForceModel (DEF/Deformable.cpp:178)
  CudaCorotationalForceModel(int numVertices, const double *restPositions, int numElements, const int *elements, double E,
                             double nu, double density, int device = 0)
      : ctx_(NULL) {
    fb_params p;
    fb_default_params(&p);
    p.youngs_modulus = E; p.poisson_ratio = nu; p.density = density; p.device = device;
    int st = fb_create(&ctx_, numVertices, restPositions, numElements, elements, 0, NULL, &p);
    if (st != FB_OK) vega_fail("fb_create", st);
    r = 3 * numVertices;  // ForceModel::r
  }
  // Same argument as the class it replaces: `new CorotationalLinearFEMForceModel(corotationalLinearFEM)` becomes
  // `new CudaCorotationalForceModel(corotationalLinearFEM)` (DEF/Deformable.cpp:186).  Mesh and per-element E / nu / density
  // are read from the model's TetMesh; non-ENu materials are rejected the way CorotationalLinearFEM's constructor rejects
  // them (`throw 1`, corotationalLinearFEM.cpp:61-64).  (The CPU-side CorotationalLinearFEM object itself is then unused: a
  // host that wants the fast setup drops it and calls the mesh-array constructor above.)
  // `warp` as in CorotationalLinearFEMForceModel(fem, warp) (corotationalLinearFEMForceModel.h:42): 0 linear, 1 default, 2 exact tangent.
  explicit CudaCorotationalForceModel(CorotationalLinearFEM *fem, int warp = 1, int device = 0) : ctx_(NULL) {
    TetMesh *mesh = fem->GetTetMesh();
    const int nV = mesh->getNumVertices(), nT = mesh->getNumElements();
    std::vector<double> x(3 * (size_t)nV + 1), E((size_t)nT + 1), nu((size_t)nT + 1), rho((size_t)nT + 1);
    std::vector<int> t(4 * (size_t)nT + 1);
    for (int i = 0; i < nV; i++)
      for (int k = 0; k < 3; k++) x[3 * (size_t)i + k] = (*mesh->getVertex(i))[k];
    for (int el = 0; el < nT; el++) {
      for (int k = 0; k < 4; k++) t[4 * (size_t)el + k] = mesh->getVertexIndex(el, k);
      VolumetricMesh::ENuMaterial *m = downcastENuMaterial(mesh->getElementMaterial(el));
      if (m == NULL) {
        printf("Error: CudaCorotationalForceModel: mesh does not consist of E, nu materials.\n");
        throw 1;
      }
      E[el] = m->getE(); nu[el] = m->getNu(); rho[el] = mesh->getElementDensity(el);
    }
    fb_params p;
    fb_default_params(&p);
    p.device = device;
    int st = fb_create_with_materials(&ctx_, nV, &x[0], nT, &t[0], 0, NULL, &E[0], &nu[0], &rho[0], &p);
    if (st != FB_OK) vega_fail("fb_create_with_materials", st);
    if (warp != 1) {
      st = fb_set_warp(ctx_, warp);
      if (st != FB_OK) vega_fail("fb_set_warp", st);
    }
    r = 3 * nV;
  }
  virtual ~CudaCorotationalForceModel() { fb_destroy(ctx_); }

  virtual void GetInternalForce(double *u, double *internalForces) {
    check(fb_compute_force_and_matrix(ctx_, u, internalForces, NULL), "GetInternalForce");
  }
  // same rows, same ascending column order as CorotationalLinearFEM::GetStiffnessMatrixTopology
  // (corotationalLinearFEM.cpp:163-186 via SparseMatrixOutline's std::map)
  virtual void GetTangentStiffnessMatrixTopology(SparseMatrix **tangentStiffnessMatrix) {
    const long long nnz = fb_nnz_stiffness(ctx_);
    std::vector<int> ia((size_t)r + 1), ja((size_t)(nnz ? nnz : 1));
    check(fb_get_stiffness_csr(ctx_, &ia[0], &ja[0]), "GetTangentStiffnessMatrixTopology");
    SparseMatrixOutline outline(r);
    for (int row = 0; row < r; row++)
      for (int k = ia[row]; k < ia[row + 1]; k++) outline.AddEntry(row, ja[k], 0.0);
    *tangentStiffnessMatrix = new SparseMatrix(&outline);
  }
  virtual void GetTangentStiffnessMatrix(double *u, SparseMatrix *K) { GetForceAndMatrix(u, NULL, K); }
  virtual void GetForceAndMatrix(double *u, double *internalForces, SparseMatrix *K) {
    std::vector<double> a;
    if (K) a.resize((size_t)fb_nnz_stiffness(ctx_) + 1);
    check(fb_compute_force_and_matrix(ctx_, u, internalForces, K ? &a[0] : NULL), "GetForceAndMatrix");
    if (K) {  // values arrive in the reference's own CSR order (GenerateCompressedRowMajorFormat): row by row, ascending columns
      size_t k = 0;
      double **rows = K->GetEntries();
      for (int row = 0; row < K->GetNumRows(); row++) {
        const int len = K->GetRowLength(row);
        memcpy(rows[row], &a[k], sizeof(double) * (size_t)len);
        k += (size_t)len;
      }
    }
  }
  fb_context *context() { return ctx_; }

 private:
  fb_context *ctx_;
  void check(int st, const char *what) { if (st != FB_OK) vega_fail(what, st); }
  CudaCorotationalForceModel(const CudaCorotationalForceModel &);
  CudaCorotationalForceModel &operator=(const CudaCorotationalForceModel &);
};

// VolumeConservingIntegrator with the step on the device.  Constructor arguments are the reference's, in the reference's
// order (DEF/PS_VolumeConservingIntegrator.h:21-26); forceModel must be a CudaCorotationalForceModel (it holds the mesh on
// the device).  massMatrix is kept for GetKineticEnergy / GetTotalMass (IntegratorBaseSparse) — the step itself uses the
// device's own bit-identical mass matrix.
class CudaVolumeConservingIntegrator2 : public IntegratorBaseSparse {
 public:
  CudaVolumeConservingIntegrator2(int r, double timestep, SparseMatrix *massMatrix, ForceModel *forceModel,
                                  int positiveDefiniteSolver = 0, int numConstrainedDOFs = 0, int *constrainedDOFs = NULL,
                                  double dampingMassCoef = 0.0, double dampingStiffnessCoef = 0.0, int maxIterations = 1,
                                  double epsilon = 1E-6, int numSolverThreads = 0)
      : IntegratorBaseSparse(r, timestep, massMatrix, forceModel, numConstrainedDOFs, constrainedDOFs, dampingMassCoef,
                             dampingStiffnessCoef), ctx_(NULL) {
    (void)positiveDefiniteSolver; (void)maxIterations; (void)epsilon; (void)numSolverThreads;
    CudaCorotationalForceModel *fm = dynamic_cast<CudaCorotationalForceModel *>(forceModel);
    if (!fm) vega_fail("CudaVolumeConservingIntegrator2: forceModel is not a CudaCorotationalForceModel", FB_ERR_INVALID_ARGUMENT);
    ctx_ = fm->context();
    if (r != fb_num_dofs(ctx_)) {  // implicitNewmarkSparse.cpp:52-56: size mismatch -> exit(1)
      printf("Error: the provided mass matrix / force model does not have correct size. r=%d device r=%d\n", r, fb_num_dofs(ctx_));
      exit(1);
    }
    apply_constraints(numConstrainedDOFs, constrainedDOFs);
    check(fb_set_timestep(ctx_, timestep), "SetTimestep");
    check(fb_set_damping(ctx_, dampingMassCoef, dampingStiffnessCoef), "SetDamping");
  }
  virtual ~CudaVolumeConservingIntegrator2() {}

  virtual int SetState(double *q_, double *qvel_ = NULL) {  // ImplicitNewmarkSparse::SetState semantics that Deformable relies on: copy in
    memcpy(q, q_, sizeof(double) * (size_t)r);
    if (qvel_) memcpy(qvel, qvel_, sizeof(double) * (size_t)r);
    return 0;
  }

  // VolumeConservingIntegrator::DoTimestep (DEF/PS_VolumeConservingIntegrator.cpp:46-260): 0 on success; a PCG failure
  // prints the reference's message and exit(-1)s like the reference (:203-209) unless FEMBRAIN_B200_NO_EXIT is defined
  virtual int DoTimestep() {
    check(fb_set_timestep(ctx_, timestep), "SetTimestep");  // the setters of IntegratorBase are inline writes to these members
    check(fb_set_damping(ctx_, dampingMassCoef, dampingStiffnessCoef), "SetDamping");
    check(fb_set_internal_force_scaling(ctx_, internalForceScalingFactor), "SetInternalForceScalingFactor");
    check(fb_set_state(ctx_, q, qvel, qaccel), "SetqState");
    check(fb_set_external_forces(ctx_, externalForces), "SetExternalForces");
    int st = fb_step(ctx_);
    forceAssemblyTime = fb_force_assembly_seconds(ctx_);
    systemSolveTime = fb_system_solve_seconds(ctx_);
    if (st == FB_ERR_SOLVER_NOT_CONVERGED) {
      printf("Error: %s sparse solver returned non-zero exit status %d.\n", "PCG", fb_last_cg_iterations(ctx_));
#ifndef FEMBRAIN_B200_NO_EXIT
      exit(-1);
#endif
      return 1;
    }
    if (st != FB_OK) vega_fail("DoTimestep", st);
    check(fb_get_state(ctx_, q, qvel, qaccel), "GetqState");
    return 0;
  }

  // replaces the list AND rebuilds the constrained system (the reference only replaces the list,
  // integratorBaseSparse.cpp:73-87, leaving systemMatrix stale)
  virtual bool setConstrainedDOF(int num, int *dofs) {
    if (!IntegratorBaseSparse::setConstrainedDOF(num, dofs)) return false;
    return apply_constraints(num, dofs);
  }

  int GetLastCGIterations() { return fb_last_cg_iterations(ctx_); }
  fb_context *context() { return ctx_; }

 private:
  fb_context *ctx_;
  void check(int st, const char *what) { if (st != FB_OK) vega_fail(what, st); }
  // DOFs come in whole-vertex triples, as Deformable::FixedVerticesToFixedDOF produces them (DEF/Deformable.cpp:294-314)
  bool apply_constraints(int num, const int *dofs) {
    if (num % 3 != 0) return false;
    std::vector<int> verts((size_t)(num / 3) + 1);
    for (int i = 0; i < num / 3; i++) {
      if (dofs[3 * i] % 3 != 0 || dofs[3 * i + 1] != dofs[3 * i] + 1 || dofs[3 * i + 2] != dofs[3 * i] + 2) return false;
      verts[i] = dofs[3 * i] / 3;
    }
    int st = fb_set_fixed_vertices(ctx_, num / 3, &verts[0]);
    if (st != FB_OK) vega_fail("setConstrainedDOF", st);
    return true;
  }
};

}  // namespace fembrain_b200
#endif
