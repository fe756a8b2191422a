/*
 * oracle_api.h — C interface shared by the two CPU checkers of the FEM step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * The same function set is exported twice, with two prefixes:
 *   fbref_*   oracle/_ref/libfembrain_ref.so  — the UNMODIFIED reference sources
 *             (/root/reference/src/3rdparty/vegafem + src/deformable/PS_VolumeConservingIntegrator.cpp)
 *             compiled in place by oracle/Makefile, wrapped by oracle/ref_harness.cpp.
 *   fbport_*  oracle/libfembrain_port.so      — oracle/vega_port.c, a plain-C restatement of the
 *             same algorithm (each function cites the reference file:line it follows).
 *
 * Conventions: all indices 0-based ints, all reals double, CSR as produced by
 * SparseMatrix::GenerateCompressedRowMajorFormat (vegafem/sparseMatrix/sparseMatrix.cpp:1151-1175).
 */
#ifndef FEMBRAIN_ORACLE_API_H
#define FEMBRAIN_ORACLE_API_H

#ifndef FBO_PREFIX
#error "define FBO_PREFIX (fbref_ or fbport_) before including oracle_api.h"
#endif
#define FBO_CAT2(a, b) a##b
#define FBO_CAT(a, b) FBO_CAT2(a, b)
#define FBO(name) FBO_CAT(FBO_PREFIX, name)

#ifdef __cplusplus
extern "C" {
#endif

/* Setup chain of Deformable::syncForceModel (src/deformable/Deformable.cpp:127-220):
 * TetMesh -> CorotationalLinearFEM -> force model -> mass matrix (inflate3Dim) ->
 * FixedVerticesToFixedDOF (:294-314, sorts the vertex list) -> VolumeConservingIntegrator
 * (maxIterations=1, epsilon=1e-6).  Returns NULL on failure. */
void *FBO(create)(int nV, const double *verts, int nT, const int *tets, double E, double nu,
                  double rho, int nFixedVerts, const int *fixedVerts, double h, double dampM,
                  double dampK);
void FBO(destroy)(void *sim);

int FBO(r)(void *sim);        /* 3*nV */
int FBO(nnz_K)(void *sim);    /* scalar nnz of the tangent stiffness matrix */
int FBO(nnz_M)(void *sim);    /* scalar nnz of the (3x inflated) mass matrix */
int FBO(rows_sys)(void *sim); /* r - numConstrainedDOFs */
int FBO(nnz_sys)(void *sim);  /* scalar nnz of systemMatrix */

/* structure (+ current values where a != NULL) */
void FBO(K_csr)(void *sim, int *ia, int *ja, double *a);   /* tangentStiffnessMatrix */
void FBO(M_csr)(void *sim, int *ia, int *ja, double *a);   /* massMatrix */
void FBO(sys_csr)(void *sim, int *ia, int *ja, double *a); /* systemMatrix */

/* CorotationalLinearFEM::BuildRowColumnIndices (corotationalLinearFEM.cpp:482-502) */
void FBO(element_maps)(void *sim, int *rowIdx4, int *colIdx16);
/* per element MInverse[16] and KElementUndeformed[144] (corotationalLinearFEM.cpp:70-145) */
void FBO(element_data)(void *sim, double *MInv16, double *K0_144);
/* SparseMatrix::BuildSuperMatrixIndices (sparseMatrix.cpp:945-991) on systemMatrix */
void FBO(super_maps)(void *sim, int *superRows, int *superIdx);
/* SparseMatrix::BuildSubMatrixIndices(massMatrix) (sparseMatrix.cpp:1004-1047), flattened in M's CSR order */
void FBO(submatrix_map)(void *sim, int *idx);

/* ForceModel::GetForceAndMatrix(u, f, K) — corotationalLinearFEM.cpp:219-470, warp=1.
 * f has r entries, Ka nnz_K entries (CSR order).  Leaves K's values in tangentStiffnessMatrix. */
void FBO(force_and_matrix)(void *sim, const double *u, double *f, double *Ka);

/* IntegratorBase state API (integratorBase.cpp:84-122) */
void FBO(set_state)(void *sim, const double *q, const double *qvel);
void FBO(get_state)(void *sim, double *q, double *qvel, double *qaccel);
void FBO(set_external_forces)(void *sim, const double *f);
/* VolumeConservingIntegrator::DoTimestep (PS_VolumeConservingIntegrator.cpp:46-260); returns its return value */
int FBO(do_timestep)(void *sim);

/* after do_timestep: Keff values in K's CSR order, rhs (bufferConstrained), internal forces, qdelta */
void FBO(K_values)(void *sim, double *a);
void FBO(rhs)(void *sim, double *b);
void FBO(internal_forces)(void *sim, double *f);
void FBO(qdelta)(void *sim, double *d);

/* CGSolver::SolveLinearSystemWithJacobiPreconditioner (CGSolver.cpp:129-190) on the current
 * systemMatrix with rhs b (rows_sys entries; NULL = current bufferConstrained), x0 = 0.
 * Returns the solver's return value (+iterations converged, -iterations not converged). */
int FBO(solve)(void *sim, const double *b, double *x, double eps, int maxIter);
/* same, but stops after exactly `iters` iterations irrespective of convergence (timing sample);
 * implemented by calling the solver with eps = 0 */
int FBO(solve_iters)(void *sim, const double *b, double *x, int iters);
/* systemMatrix->AssignSuperMatrix(tangentStiffnessMatrix) (sparseMatrix.cpp:993-1002): load the constrained
 * matrix from whatever tangentStiffnessMatrix currently holds (bench.py's bounded CPU sample) */
void FBO(assign_system)(void *sim);
/* y = systemMatrix * x (sparseMatrix.cpp:405-413) */
void FBO(sys_spmv)(void *sim, const double *x, double *y);

double FBO(assembly_time)(void *sim); /* IntegratorBaseSparse::GetForceAssemblyTime */
double FBO(solve_time)(void *sim);    /* IntegratorBaseSparse::GetSystemSolveTime */

/* PolarDecomposition::Compute (polarDecomposition.cpp:37-108); returns det */
double FBO(polar)(const double *F9, double *R9, double *S9, double tol);

#ifdef __cplusplus
}
#endif
#endif
