// fembrain_b200_vega.hpp — header-only C++98 adapter: the reference's integrator class shape on top of
// the C ABI (fembrain_b200.h).  A FemBrain maintainer swaps
//
//     m_lpIntegrator = new VolumeConservingIntegrator(m_dof, m_timeStep, m_lpMassMatrix,
//                          m_lpDeformableForceModel, ...);                 // DEF/Deformable.cpp:208-216
// for
//     m_lpIntegrator = new fembrain_b200::CudaVolumeConservingIntegrator(ctNodes, &vertices[0], ctCells,
//                          &elements[0], 1E7, 0.46, 1000, m_dof, m_timeStep, ...same trailing arguments...);
//
// and every later call site in DEF/Deformable.cpp compiles unchanged: SetExternalForcesToZero /
// SetExternalForces (:325,:343), DoTimestep (:346), GetqState / SetqState (:347,:402,:600), ResetToRest (:630),
// SetDampingStiffnessCoef / SetDampingMassCoef (:284,:290), setConstrainedDOF (:500), GetSystemSolveTime
// (Deformable.h:152).  Method names, argument meaning, copy semantics (all buffers are copied in/out,
// VEGA/integrator/integratorBase.cpp:84-122) and the 0/1 return convention of DoTimestep
// (VEGA/integrator/implicitNewmarkSparse.h:95-97) are the reference's.  Where the reference calls exit()
// (solver failure DEF/PS_VolumeConservingIntegrator.cpp:203-209, bad DOF list sparseMatrix.cpp:906-917) this
// class does the same thing by default so behaviour is identical; define FEMBRAIN_B200_NO_EXIT to get return
// codes / `throw 1` (the reference's own error idiom, corotationalLinearFEM.cpp:61-64) instead.
//
// The CPU-side Vega objects the reference constructor takes (SparseMatrix * massMatrix, ForceModel *) are not
// needed: mass matrix, corotational force model, stiffness topology and index maps are built on the GPU from
// the same mesh arrays Deformable::syncForceModel already holds (DEF/Deformable.cpp:140-178).
#ifndef FEMBRAIN_B200_VEGA_HPP
#define FEMBRAIN_B200_VEGA_HPP

#include <stdio.h>
#include <stdlib.h>

#include "fembrain_b200.h"

namespace fembrain_b200 {

class CudaVolumeConservingIntegrator {
 public:
  // constrainedDOFs: 0-indexed, pre-sorted ascending, copied (DEF/PS_VolumeConservingIntegrator.h:15-26).
  // positiveDefiniteSolver, maxIterations, epsilon (Newton loop) and numSolverThreads are accepted for
  // signature compatibility; like the reference's PCG build, one Newton iteration is performed and the first
  // and last are ignored (DEF/Deformable.cpp:100-103).
  CudaVolumeConservingIntegrator(int numVertices, const double *restPositions, int numElements, const int *elements,
                                 double E, double nu, double density, int r, double timestep,
                                 int positiveDefiniteSolver = 0, int numConstrainedDOFs = 0, int *constrainedDOFs = NULL,
                                 double dampingMassCoef = 0.0, double dampingStiffnessCoef = 0.0, int maxIterations = 1,
                                 double epsilon = 1E-6, int numSolverThreads = 0, int device = 0)
      : ctx_(NULL), r_(r), timestep_(timestep), dampingMassCoef_(dampingMassCoef), dampingStiffnessCoef_(dampingStiffnessCoef) {
    (void)positiveDefiniteSolver; (void)maxIterations; (void)epsilon; (void)numSolverThreads;
    fb_params p;
    fb_default_params(&p);
    p.youngs_modulus = E; p.poisson_ratio = nu; p.density = density;
    p.timestep = timestep; p.damping_mass = dampingMassCoef; p.damping_stiffness = dampingStiffnessCoef;
    p.device = device;
    if (r != 3 * numVertices) fail("r must equal 3 * numVertices", FB_ERR_INVALID_ARGUMENT);
    int st = fb_create_with_constrained_dofs(&ctx_, numVertices, restPositions, numElements, elements, numConstrainedDOFs,
                                             constrainedDOFs, &p);
    if (st != FB_OK) fail("fb_create_with_constrained_dofs", st);
  }
  virtual ~CudaVolumeConservingIntegrator() { fb_destroy(ctx_); }

  // --- IntegratorBase surface (VEGA/integrator/integratorBase.h:107-205) ---
  inline int Getr() { return r_; }
  virtual void ResetToRest() { check(fb_reset_to_rest(ctx_), "ResetToRest"); }
  virtual int SetState(double *q, double *qvel = NULL) { return check(fb_set_state(ctx_, q, qvel, NULL), "SetState"); }
  virtual void SetqState(const double *q, const double *qvel = NULL, const double *qaccel = NULL) {
    check(fb_set_state(ctx_, q, qvel, qaccel), "SetqState");
  }
  virtual void GetqState(double *q, double *qvel = NULL, double *qaccel = NULL) { check(fb_get_state(ctx_, q, qvel, qaccel), "GetqState"); }
  virtual void SetExternalForces(double *externalForces) { check(fb_set_external_forces(ctx_, externalForces), "SetExternalForces"); }
  virtual void AddExternalForces(double *externalForces) { check(fb_add_external_forces(ctx_, externalForces), "AddExternalForces"); }
  virtual void GetExternalForces(double *externalForces_copy) { check(fb_get_external_forces(ctx_, externalForces_copy), "GetExternalForces"); }
  virtual void SetExternalForcesToZero() { check(fb_set_external_forces_to_zero(ctx_), "SetExternalForcesToZero"); }
  virtual void SetTimestep(double timestep) { timestep_ = timestep; check(fb_set_timestep(ctx_, timestep), "SetTimestep"); }
  inline double GetTimestep() { return timestep_; }
  inline void SetDampingMassCoef(double c) { dampingMassCoef_ = c; check(fb_set_damping(ctx_, dampingMassCoef_, dampingStiffnessCoef_), "SetDampingMassCoef"); }
  inline void SetDampingStiffnessCoef(double c) { dampingStiffnessCoef_ = c; check(fb_set_damping(ctx_, dampingMassCoef_, dampingStiffnessCoef_), "SetDampingStiffnessCoef"); }
  inline double GetDampingMassCoef() { return dampingMassCoef_; }
  inline double GetDampingStiffnessCoef() { return dampingStiffnessCoef_; }
  inline void SetInternalForceScalingFactor(double s) { check(fb_set_internal_force_scaling(ctx_, s), "SetInternalForceScalingFactor"); }

  // --- the step: VolumeConservingIntegrator::DoTimestep (DEF/PS_VolumeConservingIntegrator.cpp:46-260) ---
  // returns 0 on success, 1 on failure (implicitNewmarkSparse.h:95-97); on solver failure the reference prints
  // and exit(-1)s, and so does this unless FEMBRAIN_B200_NO_EXIT is defined.
  virtual int DoTimestep() {
    int st = fb_step(ctx_);
    if (st == FB_OK) return 0;
    if (st == FB_ERR_SOLVER_NOT_CONVERGED) {
      printf("Error: %s sparse solver returned non-zero exit status %d.\n", "PCG", fb_last_cg_iterations(ctx_));
#ifndef FEMBRAIN_B200_NO_EXIT
      exit(-1);
#endif
      return 1;
    }
    fail("DoTimestep", st);
    return 1;
  }

  // --- IntegratorBaseSparse surface (VEGA/integrator/integratorBaseSparse.h:45-90) ---
  inline double GetForceAssemblyTime() { return fb_force_assembly_seconds(ctx_); }
  inline double GetSystemSolveTime() { return fb_system_solve_seconds(ctx_); }
  // Replaces the constrained-DOF list AND rebuilds the constrained system (the reference only replaces the
  // list, integratorBaseSparse.cpp:73-87).  DOFs must come in whole-vertex triples, as Deformable produces them.
  virtual bool setConstrainedDOF(int num, int *arrConstrainedDOFs_) {
    if (num == 0 || arrConstrainedDOFs_ == 0) return false;
    if (num % 3 != 0) return false;
    int *verts = (int *)malloc(sizeof(int) * (size_t)(num / 3));
    for (int i = 0; i < num / 3; i++) {
      if (arrConstrainedDOFs_[3 * i] % 3 != 0 || arrConstrainedDOFs_[3 * i + 1] != arrConstrainedDOFs_[3 * i] + 1 ||
          arrConstrainedDOFs_[3 * i + 2] != arrConstrainedDOFs_[3 * i] + 2) {
        free(verts);
        return false;
      }
      verts[i] = arrConstrainedDOFs_[3 * i] / 3;
    }
    int st = fb_set_fixed_vertices(ctx_, num / 3, verts);
    free(verts);
    return st == FB_OK;
  }

  // --- extras that have no reference counterpart ---
  inline int GetLastCGIterations() { return fb_last_cg_iterations(ctx_); }
  inline fb_context *context() { return ctx_; }
  // device pointer to q for a zero-copy hand-off to rendering (see fb_displacements_dev)
  inline const double *DeviceDisplacements() { return fb_displacements_dev(ctx_); }

 protected:
  fb_context *ctx_;
  int r_;
  double timestep_, dampingMassCoef_, dampingStiffnessCoef_;

  int check(int st, const char *what) {
    if (st != FB_OK) fail(what, st);
    return 0;
  }
  void fail(const char *what, int st) {
    printf("Error: fembrain_b200 %s failed: %s (%s)\n", what, fb_status_string(st), fb_last_error_string());
#ifndef FEMBRAIN_B200_NO_EXIT
    exit(1);
#else
    throw 1;
#endif
  }

 private:
  CudaVolumeConservingIntegrator(const CudaVolumeConservingIntegrator &);
  CudaVolumeConservingIntegrator &operator=(const CudaVolumeConservingIntegrator &);
};

}  // namespace fembrain_b200
#endif
