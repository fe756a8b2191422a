/*
 * fembrain_b200.h — C ABI of the B200-native soft-tissue FEM step.
 *
 * Drop-in boundary for ONE path of pouryashirazian/FemBrain: Deformable::timestep() ->
 * VolumeConservingIntegrator::DoTimestep() -> corotational assembly -> Jacobi-PCG implicit-Euler
 * solve.  The reference has no C ABI or FFI for this path; its seam is the C++ interface of
 * `class Deformable` and Vega's ForceModel / IntegratorBaseSparse.  Each entry point below names
 * the reference interface it replaces (paths relative to /root/reference/src; VEGA =
 * 3rdparty/vegafem, DEF = deformable).  include/fembrain_b200_vega.hpp is the header-only C++
 * adapter with the reference's class shape on top of these calls; INTEGRATION.md shows the edit a
 * maintainer makes in DEF/Deformable.cpp.
 *
 * Rules of the ABI: plain pointers and sizes, no exceptions, no C++ or torch types; every call
 * returns an fb_status (0 = ok) unless documented otherwise; one context = one CUDA device + one
 * CUDA stream = one caller thread at a time (the reference is single-threaded and non-reentrant,
 * graphics/SceneGraph.cpp:187-193); contexts are independent of one another.  All reals are
 * double, all indices 0-based int, DOF d of vertex v is 3*v+d.  Buffer arguments named *_dev are
 * device pointers on the context's device; all others are host pointers.  There is no CPU
 * fallback: creation fails with FB_ERR_NO_DEVICE when no sm_100 device is usable.
 */
#ifndef FEMBRAIN_B200_H
#define FEMBRAIN_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FB_ABI_VERSION 1

typedef enum fb_status {
  FB_OK = 0,
  FB_ERR_INVALID_ARGUMENT = 1,     /* NULL pointer, negative size, index out of range, unsorted input where sorted is required */
  FB_ERR_NO_DEVICE = 2,            /* no CUDA device / device is not sm_100: there is no fallback */
  FB_ERR_CUDA = 3,                 /* a CUDA runtime call failed; see fb_last_error_string() */
  FB_ERR_OUT_OF_MEMORY = 4,
  FB_ERR_SOLVER_NOT_CONVERGED = 5, /* PCG hit cg_max_iter: the reference printf+exit(-1)s here
                                      (DEF/PS_VolumeConservingIntegrator.cpp:203-209); state is left unchanged */
  FB_ERR_BAD_MESH = 6,             /* vertex index out of range, or a vertex that belongs to no tetrahedron
                                      (the reference dereferences diagonal index -1 in that case, sparseMatrix.cpp:622-630) */
  FB_ERR_COMM = 7,                 /* NCCL failure in a partitioned context */
  FB_ERR_NOT_SUPPORTED = 8
} fb_status;

/* Parameters the reference hard-codes (defaults set by fb_default_params). */
typedef struct fb_params {
  double youngs_modulus;     /* 1e7      DEF/Deformable.cpp:178 */
  double poisson_ratio;      /* 0.46     DEF/Deformable.cpp:178 */
  double density;            /* 1000     DEF/Deformable.cpp:178 */
  double timestep;           /* 0.0333   DEF/Deformable.cpp:113 */
  double damping_mass;       /* 0.0      DEF/Deformable.cpp:107 */
  double damping_stiffness;  /* 0.01     DEF/Deformable.cpp:110 */
  double cg_epsilon;         /* 1e-6     DEF/PS_VolumeConservingIntegrator.cpp:196 */
  int cg_max_iterations;     /* 10000    DEF/PS_VolumeConservingIntegrator.cpp:197 */
  double polar_tolerance;    /* 1e-6     VEGA/corotationalLinearFEM/corotationalLinearFEM.cpp:263 */
  double internal_force_scaling; /* 1.0  VEGA/integrator/integratorBase.cpp:46 */
  int device;                /* CUDA device ordinal, default 0 */
  int keep_raw_stiffness;    /* 0: K is consumed in registers; 1: also store raw K every step for fb_get_stiffness_values */
  int solver_variant;        /* FB_SOLVER_JACOBI_PCG (0, default): the reference's solver.  See fb_set_solver */
  int warm_start;            /* solver variants only: start PCG from the previous step's solution instead of 0 */
  int reserved[4];
} fb_params;

/* Linear solver of the implicit step.  0 is the reference's algorithm and the parity path.  The others are LABELLED VARIANTS
 * (not in the reference): the same system Keff dv = rhs, the same stopping rule (Jacobi-weighted residual
 * sum r^2/diag <= eps^2 sum b^2/diag, CGSolver.cpp:147-150), a faster-converging preconditioner (fb_mg.cu). */
#define FB_SOLVER_JACOBI_PCG 0        /* CGSolver::SolveLinearSystemWithJacobiPreconditioner, CGSolver.cpp:129-190 */
#define FB_SOLVER_BLOCK_JACOBI_PCG 1  /* 3x3 block-diagonal preconditioner; any mesh */
#define FB_SOLVER_MG_PCG 2            /* geometric multigrid V-cycle preconditioner (Chebyshev smoothing); meshes on a tensor grid (fb_set_grid) */

typedef struct fb_context fb_context;

void fb_default_params(fb_params *p);
int fb_abi_version(void);
const char *fb_status_string(int status);
/* text of the most recent failure on the calling thread (CUDA error string, offending index, ...) */
const char *fb_last_error_string(void);

/* ---- lifetime --------------------------------------------------------------------------------
 * Replaces the setup chain of Deformable::syncForceModel (DEF/Deformable.cpp:127-220): TetMesh
 * (VEGA/volumetricMesh/tetMesh.cpp:129-131), CorotationalLinearFEM ctor (corotationalLinearFEM.cpp:40-146),
 * GenerateMassMatrix::computeMassMatrix(inflate3Dim) (generateMassMatrix.cpp:33-76),
 * FixedVerticesToFixedDOF (Deformable.cpp:294-314; the vertex list need not be sorted, duplicates are
 * an error) and the VolumeConservingIntegrator / ImplicitNewmarkSparse ctor
 * (implicitNewmarkSparse.cpp:39-83).  rest_positions: 3*num_vertices; tets: 4*num_tets.  Inputs are copied. */
int fb_create(fb_context **out, int num_vertices, const double *rest_positions, int num_tets,
              const int *tets, int num_fixed_vertices, const int *fixed_vertices, const fb_params *params);
/* Same, with per-element materials (E, nu, density arrays of num_tets entries; NULL = use params):
 * what a .veg file with several *MATERIAL / *REGION sections produces (volumetricMesh.cpp:45-...). */
int fb_create_with_materials(fb_context **out, int num_vertices, const double *rest_positions,
                             int num_tets, const int *tets, int num_fixed_vertices,
                             const int *fixed_vertices, const double *E, const double *nu,
                             const double *density, const fb_params *params);
/* Same as fb_create with an arbitrary sorted, 0-indexed constrained-DOF list, the argument the
 * integrator constructor itself takes (DEF/PS_VolumeConservingIntegrator.h:21-26). */
int fb_create_with_constrained_dofs(fb_context **out, int num_vertices, const double *rest_positions,
                                    int num_tets, const int *tets, int num_constrained_dofs,
                                    const int *constrained_dofs, const fb_params *params);
void fb_destroy(fb_context *ctx);
/* Deformable::syncForceModel on an existing model after its mesh changed (cutting: cutCompleted -> syncForceModel, main.cpp:614-617;
 * DEF/Deformable.cpp:127-220 deletes and rebuilds TetMesh, CorotationalLinearFEM, force model, mass matrix and integrator).
 * Full re-setup IN PLACE: the handle stays valid, parameters / timestep / damping / CG settings / gravity / floor / haptic radius
 * are kept, state restarts at rest, haptic forces and the edge list are dropped (they index the old mesh).  ~10 ms at 1M tets
 * from the memory pool (the reference's constructors: 18.8 s).  Uniform material (params); single-mesh contexts only. */
int fb_sync_force_model(fb_context *ctx, int num_vertices, const double *rest_positions, int num_tets, const int *tets,
                        int num_fixed_vertices, const int *fixed_vertices);

/* Deformable::setFixedVertices + the integrator rebuild it needs.  (IntegratorBaseSparse::setConstrainedDOF,
 * integratorBaseSparse.cpp:73-87, updates the list but not systemMatrix — a reference bug; here the
 * constrained system always matches the list.) */
int fb_set_fixed_vertices(fb_context *ctx, int num_fixed_vertices, const int *fixed_vertices);

/* ---- sizes ------------------------------------------------------------------------------------ */
int fb_num_vertices(const fb_context *ctx);
int fb_num_tets(const fb_context *ctx);
int fb_num_dofs(const fb_context *ctx);              /* r = 3*num_vertices (ForceModel::Getr) */
int fb_num_constrained_dofs(const fb_context *ctx);
int fb_num_local_dofs(const fb_context *ctx);        /* DOFs of the LOCAL system the inspection hooks describe: = fb_num_dofs except on a
                                                        partitioned context (owned + ghost vertices of this rank) */
long long fb_nnz_stiffness(const fb_context *ctx);   /* scalar nnz of K (SparseMatrix::GetNumEntries) */
long long fb_nnz_mass(const fb_context *ctx);
long long fb_nnz_system(const fb_context *ctx);      /* scalar nnz of the constrained systemMatrix */

/* ---- forces and state: IntegratorBase API (VEGA/integrator/integratorBase.cpp:84-132) ---------- */
int fb_set_external_forces(fb_context *ctx, const double *f);          /* SetExternalForces */
int fb_add_external_forces(fb_context *ctx, const double *f);          /* AddExternalForces */
int fb_set_external_forces_to_zero(fb_context *ctx);                   /* SetExternalForcesToZero */
int fb_get_external_forces(fb_context *ctx, double *f);                /* GetExternalForces */
int fb_set_state(fb_context *ctx, const double *q, const double *qvel, const double *qaccel); /* SetqState; qvel/qaccel may be NULL */
int fb_get_state(fb_context *ctx, double *q, double *qvel, double *qaccel);                   /* GetqState; any may be NULL */
int fb_reset_to_rest(fb_context *ctx);                                 /* ResetToRest */
/* device-resident variants (no host copy; pointers must live on the context's device) */
int fb_set_external_forces_dev(fb_context *ctx, const double *f_dev);
int fb_get_state_dev(fb_context *ctx, double *q_dev, double *qvel_dev, double *qaccel_dev);
/* read-only device pointer to the displacement vector q (r doubles): the hand-off to rendering
 * (GPUPoly::applyFemDisplacements, implicit/OclPolygonizer.cpp:1543-1584) without a host round trip */
const double *fb_displacements_dev(const fb_context *ctx);

int fb_set_timestep(fb_context *ctx, double h);                        /* IntegratorBase::SetTimestep */
/* the `warp` argument of CorotationalLinearFEMForceModel(fem, warp) / ComputeForceAndStiffnessMatrix
 * (VEGA/corotationalLinearFEM/corotationalLinearFEM.cpp:219-449): 1 = corotational with the approximate tangent R K0 R^T (the
 * default, what Deformable.cpp:186 builds and the only mode served by the one-pass gather assembly), 0 = linear FEM (no
 * rotations), 2 = exact tangent (adds the dR/dx terms, :296-428).  0 and 2 run the two-phase assembly with one thread per
 * element; K and f stay bit-identical to the reference's for every mode. */
int fb_set_warp(fb_context *ctx, int warp);
int fb_get_warp(const fb_context *ctx);
int fb_set_damping(fb_context *ctx, double damping_mass, double damping_stiffness); /* SetDampingMassCoef / SetDampingStiffnessCoef */
int fb_set_internal_force_scaling(fb_context *ctx, double s);          /* SetInternalForceScalingFactor */
int fb_set_cg(fb_context *ctx, double epsilon, int max_iterations);
/* Solver variants (single-mesh, single-GPU contexts).  fb_set_grid declares that the mesh's vertices are the nodes of an
 * nx x ny x nz tensor grid numbered (i*ny + j)*nz + k, as VolMeshSamples::CreateTruthCube numbers them (any axis spacing; checked
 * against the rest positions) — what FB_SOLVER_MG_PCG coarsens.  fb_set_solver selects the variant (FB_OK, or
 * FB_ERR_NOT_SUPPORTED / FB_ERR_INVALID_ARGUMENT with the state unchanged); fb_last_cg_iterations / _residual_ratio then
 * report the variant's iterations and its final weighted residual ratio.  fb_solve uses the selected solver too. */
int fb_set_grid(fb_context *ctx, int nx, int ny, int nz);
int fb_set_solver(fb_context *ctx, int variant, int warm_start);
int fb_get_solver(const fb_context *ctx, int *variant, int *warm_start, int *levels);
const char *fb_solver_name(int variant);
/* vertices and 3x3 blocks of every level of the variant's hierarchy, finest first (fb_get_solver gives the level count) */
int fb_get_solver_levels(const fb_context *ctx, int capacity, int *num_vertices, long long *num_blocks);
/* the cycle's smoother: sweeps on the finest level before and after the coarse correction (coarse levels: 4), 1 when they form a Chebyshev iteration on
 * [1.1 lambda_max / alpha, 1.1 lambda_max] (0: damped block Jacobi), and how many levels use the structured slot-major product */
int fb_get_solver_smoother(const fb_context *ctx, int *sweeps, int *chebyshev, double *alpha, int *structured_levels);

/* ---- the step ---------------------------------------------------------------------------------
 * VolumeConservingIntegrator::DoTimestep (DEF/PS_VolumeConservingIntegrator.cpp:46-260), PCG branch,
 * one Newton iteration: assemble f_int and K at q; Keff = M + h*D + h^2*K with D = dampK*K + dampM*M;
 * rhs = -h*((h*K + D)*qvel + f_int - f_ext); solve the constrained system by Jacobi-PCG from x0 = 0;
 * qvel += dv; q += h*qvel; constrained DOFs zeroed.  Returns FB_OK or FB_ERR_SOLVER_NOT_CONVERGED. */
int fb_step(fb_context *ctx);
/* One fb_step of each of n DISTINCT, independent contexts (the batch of BASELINE configs[3] when every mesh needs its own
 * context, e.g. with a solver variant), issued from `host_threads` host threads so that the contexts' kernels overlap on the
 * device.  status (may be NULL) receives every context's fb_step result; the return value is the first one that is not FB_OK. */
int fb_step_many(fb_context *const *ctxs, int n, int host_threads, int *status);

/* Deformable::timestep (DEF/Deformable.cpp:318-420) around fb_step: zero forces, optional gravity
 * (-10000 on y when enabled and no contact, :331-338), haptic forces with ring spreading
 * (applyHapticForces, :634-706), the step, then the floor-plane post-step (:350-402).  See
 * fb_deformable_* setters below. */
int fb_deformable_timestep(fb_context *ctx);
int fb_deformable_set_gravity(fb_context *ctx, int enabled);                       /* Deformable::setGravity */
int fb_deformable_set_floor(fb_context *ctx, int enabled, double floor_y);         /* m_collisionObj origin y (:351) */
/* Deformable::hapticSetCurrentForces (:712-717) + m_bHapticInProgress; forces: 3*count */
int fb_deformable_set_haptic_forces(fb_context *ctx, int count, const int *vertex_indices, const double *forces, int in_progress);
int fb_deformable_set_haptic_neighborhood(fb_context *ctx, int rings);             /* m_hapticForceNeighorhoodSize (5) */
int fb_deformable_contact_count(const fb_context *ctx);                            /* m_ctCollided */
/* Ring spreading walks VolMesh::get_node_neighbors (DEF/VolMesh.cpp:1346-1363), which indexes the GLOBAL edge array
 * with a counter that runs over the node's incident-edge COUNT.  Default here: true mesh adjacency.  Passing the
 * host's edge array (num_edges pairs from,to in VolMesh::m_vEdges order) with reference_quirk = 1 reproduces the
 * reference's rings exactly; reference_quirk = 0 or num_edges = 0 returns to true adjacency. */
int fb_deformable_set_edge_list(fb_context *ctx, int num_edges, const int *from_to, int reference_quirk);
/* Force producers' vertex queries on the CURRENT positions (rest + displacement), evaluated on the device:
 * Deformable::pickVertices (DEF/Deformable.cpp:430-448; closed box, graphics/AABB.h:84-92) — ascending indices,
 * up to `capacity` of them (and their coordinates if coords != NULL, 3 per vertex); *count = all vertices inside.
 * Deformable::pickVertex -> CuttableMesh::findClosestVertex (DEF/Deformable.cpp:422-428, DEF/CuttableMesh.cpp:511-526) —
 * lowest index among the nearest vertices, -1 for an empty mesh; dist / vertex may be NULL. */
int fb_deformable_pick_vertices(fb_context *ctx, const double *box_lo, const double *box_hi, int capacity, int *indices,
                                double *coords, int *count);
int fb_deformable_pick_vertex(fb_context *ctx, const double *world_pos, int *index, double *dist, double *vertex);

/* ---- mesh ingest and rendering hand-off (the callers either side of the path) ---------------------
 * .veg reader with the rules of the reference's loader (VolumetricMesh(char*), VEGA/volumetricMesh/volumetricMesh.cpp:45-535):
 * vertices, 0-based TET elements and one (E, nu, density) triple per element from the *MATERIAL / *SET / *REGION sections.
 * Output arrays are malloc'ed by the library (any may be NULL to skip) and released with fb_veg_free.  Host-only: no GPU needed. */
int fb_veg_load(const char *path, int *num_vertices, int *num_tets, double **vertices, int **tets, double **E, double **nu,
                double **density);
void fb_veg_free(void *array);
/* TetGen <basename>.node / <basename>.ele with the rules of TetMesh(char*, int) (VEGA/volumetricMesh/tetMesh.cpp:45-127):
 * 1-indexed consecutive lines, 3 coordinates, 4 vertices per element; one material (E 1e8, nu 0.45, density 1000).
 * Same output conventions as fb_veg_load (release with fb_veg_free).  Host-only. */
int fb_tetgen_load(const char *basename, int *num_vertices, int *num_tets, double **vertices, int **tets, double **E, double **nu,
                   double **density);
/* .veg writers.  Host-only.
 * FB_VEG_STYLE_FEMBRAIN = VolMeshIO::writeVega (DEF/VolMeshIO.cpp:171-224), the writer behind every "Generated by FemBrain" model
 * in data/models/blobtree: coordinates in the default ostream format (6 significant digits, printf %g), 1-based ids, and ALWAYS
 * the fixed material "ENU, 1000, 10000000, 0.45" on allElements — E / nu / density are ignored, as that writer has none.
 * FB_VEG_STYLE_VEGA = VolumetricMesh::save (VEGA/volumetricMesh/volumetricMesh.cpp:646-757): %.15G coordinates; one *MATERIAL
 * per distinct (density, E, nu) triple in order of first appearance (named material_<k>); with one material a single region
 * "allElements, material_0", otherwise one *SET set_<k> (1-based element ids, 8 per line) and one *REGION per material.
 * E / nu / density are per-element arrays as fb_veg_load returns them; all three NULL = no *MATERIAL / *REGION section (the
 * loaders then apply the reference's default material).  Reading the result back with fb_veg_load or the reference's loader
 * gives the same elements and per-element materials, and the reference's own TetMesh::save of that file reproduces it byte
 * for byte (tests/test_veg.py). */
#define FB_VEG_STYLE_FEMBRAIN 0
#define FB_VEG_STYLE_VEGA 1
int fb_veg_save(const char *path, int style, int num_vertices, const double *vertices, int num_tets, const int *tets,
                const double *E, const double *nu, const double *density);
/* TetMesh(char* filename) + the setup chain of fb_create_with_materials */
int fb_create_from_veg(fb_context **out, const char *path, int num_fixed_vertices, const int *fixed_vertices, const fb_params *params);
/* GPUPoly::applyFemDisplacements + ApplyVertexDeformations (implicit/OclPolygonizer.cpp:1543-1584, data/opencl/Polygonizer.cl:1417-1427):
 * out[i] = rest[i] + (float4)(q[3i], q[3i+1], q[3i+2], 0) for the first `count` vertices, in float, on the device.
 * rest_xyzw NULL = the context's rest positions cast to float with w = 1.  The _dev variant takes device pointers (e.g. a mapped
 * GL vertex buffer) and involves no host copy at all. */
int fb_export_positions_float4(fb_context *ctx, int count, const float *rest_xyzw, float *out_xyzw);
int fb_export_positions_float4_dev(fb_context *ctx, int count, const float *rest_xyzw_dev, float *out_xyzw_dev);

/* ---- timing / solver statistics (IntegratorBaseSparse::GetForceAssemblyTime / GetSystemSolveTime,
 * integratorBaseSparse.h:66-67; CGSolver return value, CGSolver.cpp:189) -------------------------- */
double fb_force_assembly_seconds(const fb_context *ctx); /* CUDA-event time of the last step's assembly kernels */
double fb_system_solve_seconds(const fb_context *ctx);   /* CUDA-event time of the last step's PCG */
double fb_step_seconds(const fb_context *ctx);           /* whole fb_step, device time */
int fb_last_cg_iterations(const fb_context *ctx);
double fb_last_cg_residual_ratio(const fb_context *ctx); /* rho_final / rho_0 (M^-1-weighted, squared) */
long long fb_kernel_launches(const fb_context *ctx);     /* kernels of this library launched so far on this context */
size_t fb_device_bytes(const fb_context *ctx);           /* device memory held by the context */
/* Contexts take device memory from the device's stream-ordered pool and fb_destroy returns it there, so that a
 * destroy -> create cycle (the reference's full re-setup after a cut: cutCompleted -> syncForceModel,
 * main.cpp:614-617, DEF/Deformable.cpp:127-220) does not pay for allocation again.  This gives unused pool memory
 * back to the driver. */
int fb_trim_memory(void);
/* 1 when the library was built with the shelved SpMV / PCG experiments of csrc/experiments/ (python -m fembrain_b200.build
 * --experiments; selected at run time with FEMBRAIN_B200_SPMV=sym|tma, FEMBRAIN_B200_PCG=persistent), 0 for the default build */
int fb_experiments_built(void);
/* Debug aid (environment FEMBRAIN_B200_GUARD=1 at process start): every device allocation of every context is wrapped in
 * 256-byte guard bands; this call reports how many live allocations were checked and how many had a band overwritten (an
 * out-of-bounds write by one of the library's kernels).  Always 0 / 0 when the guard mode is off. */
int fb_check_guards(long long *checked, long long *corrupted);

/* ---- inspection hooks for parity (outputs are host buffers sized by the fb_nnz_ / fb_num_ calls) -
 * CSR in the layout of SparseMatrix::GenerateCompressedRowMajorFormat (sparseMatrix.cpp:1151-1175). */
int fb_get_stiffness_csr(fb_context *ctx, int *ia, int *ja);                 /* GetStiffnessMatrixTopology, corotationalLinearFEM.cpp:163-186 */
int fb_get_mass_csr(fb_context *ctx, int *ia, int *ja, double *a);           /* computeMassMatrix */
int fb_get_system_csr(fb_context *ctx, int *ia, int *ja, double *a);         /* systemMatrix: RemoveRowsColumns (sparseMatrix.cpp:1296-1357) + AssignSuperMatrix of the last Keff */
int fb_get_element_maps(fb_context *ctx, int *row_indices4, int *column_indices16); /* BuildRowColumnIndices, corotationalLinearFEM.cpp:482-502 */
int fb_get_element_data(fb_context *ctx, double *minverse16, double *k0_144);/* MInverse / KElementUndeformed (corotationalLinearFEM.cpp:70-145) */
int fb_get_super_maps(fb_context *ctx, int *super_rows, int *super_indices); /* BuildSuperMatrixIndices, sparseMatrix.cpp:945-991 */
int fb_get_submatrix_map(fb_context *ctx, int *indices);                     /* BuildSubMatrixIndices(M), sparseMatrix.cpp:1004-1047 */
int fb_get_constrained_dofs(fb_context *ctx, int *dofs);
/* ForceModel::GetForceAndMatrix(u, f, K) (corotationalLinearFEM.cpp:219-470, warp=1): f has r entries,
 * K_values nnz_stiffness entries; either may be NULL.  Does not touch the integrator state. */
int fb_compute_force_and_matrix(fb_context *ctx, const double *u, double *f, double *K_values);
/* after fb_step: Keff (tangentStiffnessMatrix as DoTimestep leaves it), rhs (bufferConstrained,
 * r - num_constrained entries), internal forces, qdelta */
int fb_get_effective_stiffness_values(fb_context *ctx, double *a);
int fb_get_rhs(fb_context *ctx, double *b_constrained);
int fb_get_internal_forces(fb_context *ctx, double *f);
int fb_get_qdelta(fb_context *ctx, double *d);
/* CGSolver::SolveLinearSystemWithJacobiPreconditioner (CGSolver.cpp:129-190) on the current Keff:
 * b_constrained / x_constrained have r - num_constrained entries, x0 = 0; *iterations receives the
 * reference's return value (+n converged, -n not converged).  b NULL = the last step's rhs. */
int fb_solve(fb_context *ctx, const double *b_constrained, double *x_constrained, double epsilon,
             int max_iterations, int *iterations);
/* y = systemMatrix * x on constrained vectors (SparseMatrix::MultiplyVector, sparseMatrix.cpp:405-413) */
int fb_system_multiply(fb_context *ctx, const double *x_constrained, double *y_constrained);

/* ---- device-side timing of a region of calls (bench.py): events are recorded on the context's stream -- */
int fb_timer_start(fb_context *ctx);
int fb_timer_stop(fb_context *ctx, double *seconds);
/* When enabled, fb_step brackets every 16th PCG iteration's SpMV launch with a CUDA-event pair (at most 64
 * per step).  fb_get_spmv_profile returns the mean duration of those launches over all steps since the
 * profile was enabled, the number of samples, and the algorithmic bytes one launch streams
 * (values + block columns + row pointers + x read + y write). */
int fb_set_profiling(fb_context *ctx, int enabled);
int fb_get_spmv_profile(fb_context *ctx, double *mean_seconds, int *samples, double *bytes_per_launch);

/* ---- micro-benchmark hooks (bench.py roofline section): run `repeats` launches of one kernel on the
 * context's current matrices and return the mean device time per launch in seconds -------------- */
int fb_bench_spmv(fb_context *ctx, int repeats, double *seconds_per_launch);
int fb_bench_assembly(fb_context *ctx, int repeats, double *seconds_per_launch);
int fb_bench_cg_iteration(fb_context *ctx, int repeats, double *seconds_per_iteration);

/* ---- a batch of independent meshes in one context (BASELINE config 4) ----------------------------
 * In the reference every mesh is its own Deformable with its own integrator and CG solver.  A batch context holds `count`
 * meshes as ONE block-diagonal system: setup, assembly, right-hand side and state update are the single-mesh code on the
 * concatenated mesh (values bit-identical to a context per mesh); PCG keeps the scalars, the iteration count and the
 * stopping rule of CGSolver::SolveLinearSystemWithJacobiPreconditioner (VEGA/sparseSolver/CGSolver.cpp:129-190) PER MESH,
 * and every kernel launch covers all meshes that are still iterating.  Inputs are concatenated in mesh order: vertices,
 * tets and fixed vertices with MESH-LOCAL vertex ids.  Every vector of the force/state API is the concatenation in the
 * same order (fb_batch_offsets gives the first vertex / tet of every mesh, count + 1 entries each).  fb_step returns
 * FB_ERR_SOLVER_NOT_CONVERGED if any mesh did not converge; fb_last_cg_iterations is the largest count (negative in that
 * case), fb_batch_last_cg_iterations the reference's return value per mesh. */
int fb_create_batch(fb_context **out, int count, const int *num_vertices, const double *rest_positions, const int *num_tets,
                    const int *tets, const int *num_fixed_vertices, const int *fixed_vertices, const fb_params *params);
int fb_batch_count(const fb_context *ctx);
int fb_batch_offsets(const fb_context *ctx, int *vertex_offsets, int *tet_offsets);
int fb_batch_last_cg_iterations(const fb_context *ctx, int *iterations, double *residual_ratios);

/* ---- partitioned (multi-GPU) contexts: one process per GPU, row-block partition -----------------
 * No reference counterpart (the reference is single-threaded CPU code).  Every rank
 * passes the same global mesh; the context keeps rows [vertex_begin, vertex_end) of this rank plus
 * ghost columns.  comm_id: the 128-byte ncclUniqueId produced by fb_comm_unique_id on rank 0 and
 * broadcast by the host program (torch.distributed / MPI / files). */
int fb_comm_unique_id(void *id128);
int fb_create_partitioned(fb_context **out, int num_vertices, const double *rest_positions,
                          int num_tets, const int *tets, int num_fixed_vertices,
                          const int *fixed_vertices, const fb_params *params, int rank, int world,
                          const void *comm_id128);
int fb_partition_range(const fb_context *ctx, int *vertex_begin, int *vertex_end);
/* The row blocks are contiguous ranges of the caller's vertex numbering when that cuts few tets (structured meshes: the
 * benchmark cube gets slabs), otherwise of a Cuthill-McKee ordering of the vertex graph computed on the host (METIS is not
 * required): fb_partition_ordering returns that choice, order[new] = caller's vertex id (identity when *reordered = 0);
 * fb_partition_range and fb_plan_partition's vertex_begin/end are positions in it.  All vertex ids and vectors crossing
 * the ABI stay in the caller's numbering. */
int fb_partition_ordering(int num_vertices, int num_tets, const int *tets, int world, int *order, int *reordered);
int fb_partition_reordered(const fb_context *ctx);
/* 1 when the ranks exchange the PCG scalars and the halo of d through peer-memory stores over NVLink (CUDA IPC mappings,
 * flags in each rank's comm block), 0 when they go through NCCL calls between kernels (FEMBRAIN_B200_P2P=0, or the
 * mapping could not be set up). */
int fb_partition_peer_memory(const fb_context *ctx);
/* On a partitioned context every vector argument of the force/state calls is GLOBAL length (3*num_vertices of the whole
 * mesh): setters take the global vector and keep this rank's part; getters write this rank's OWNED entries and zeros
 * elsewhere, so the sum over ranks is the full vector.  fb_num_vertices/_tets/_dofs report the global mesh; the
 * inspection hooks (CSR, maps, rhs, ...) describe the rank's LOCAL system. */
/* The LOCAL system of a partitioned context (what the inspection hooks describe): local vertices [local_begin, local_end) are
 * the owned rows (contiguous), the others are ghosts; l2g[fb_num_local_dofs/3] maps a local vertex to the caller's vertex id. */
int fb_partition_local_range(const fb_context *ctx, int *local_begin, int *local_end);
int fb_partition_local_to_global(const fb_context *ctx, int *l2g);
/* Owned-range variants for callers that keep their data distributed (no global-length vector, no host gather): f_owned,
 * q, qvel, qaccel hold 3*(vertex_end - vertex_begin) doubles, the rows fb_partition_range reports, in the partition ordering
 * (= the caller's numbering unless fb_partition_reordered).  On an ordinary context they are the whole vectors.  These are
 * the calls a host program uses every frame (IntegratorBase::SetExternalForces / GetqState, integratorBase.cpp:84-122). */
int fb_set_external_forces_owned(fb_context *ctx, const double *f_owned);
int fb_get_state_owned(fb_context *ctx, double *q_owned, double *qvel_owned, double *qaccel_owned);
/* Host-only view of the partition (runs without a GPU): what rank `rank` of `world` owns and exchanges.
 * counts[7] = {vertex_begin, vertex_end, local vertices, local tets, neighbours, total send vertices, total recv vertices};
 * every other output may be NULL; sizes come from a first call: l2g[counts[2]] (the local mesh's vertices in local order: ascending position in the partition ordering, as caller's vertex ids), local_tets[counts[3]], nbr_ranks/send_counts/recv_counts[counts[4]], send_global[counts[5]],
 * recv_global[counts[6]] (concatenated per neighbour, ascending global ids). */
int fb_plan_partition(int num_vertices, int num_tets, const int *tets, int world, int rank, int *counts, int *l2g,
                      int *local_tets, int *nbr_ranks, int *send_counts, int *recv_counts, int *send_global,
                      int *recv_global);

#ifdef __cplusplus
}
#endif
#endif /* FEMBRAIN_B200_H */
