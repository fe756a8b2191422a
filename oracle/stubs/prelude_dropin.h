/* oracle/stubs/prelude_dropin.h — TEST INFRASTRUCTURE ONLY; forced first (-include) when oracle/Makefile compiles the
 * reference's OWN src/deformable/Deformable.cpp a SECOND time, as the drop-in demonstration (oracle/_ref/libfembrain_dropin.so).
 *
 * It applies, with the preprocessor and without touching the reference's files, exactly the edit INTEGRATION.md §4 asks a
 * maintainer to make in Deformable.{h,cpp}:
 *     ImplicitNewmarkSparse * m_lpIntegrator                    ->  IntegratorBaseSparse * m_lpIntegrator
 *     new CorotationalLinearFEMForceModel(m_lpDeformable)       ->  new fembrain_b200::CudaCorotationalForceModel(m_lpDeformable)
 *     new VolumeConservingIntegrator(...same arguments...)      ->  new fembrain_b200::CudaVolumeConservingIntegrator2(...)
 * Every other line of Deformable.cpp is compiled as the reference wrote it, so the result is the reference's Deformable
 * (timestep, haptic rings, floor post-step, picks) running its solves on the B200 through libfembrain_b200.so.
 * The Vega headers are parsed BEFORE the renaming macros (their include guards then keep them out of reach). */
#ifndef FB_STUB_PRELUDE_DROPIN_H
#define FB_STUB_PRELUDE_DROPIN_H
#include "prelude_deformable.h"
#include "corotationalLinearFEM.h"
#include "corotationalLinearFEMForceModel.h"
#include "generateMassMatrix.h"
#include "PS_VolumeConservingIntegrator.h"
#define FEMBRAIN_B200_NO_EXIT /* failures (no GPU in the build container) surface as `throw 1`, not exit(1) */
#include "fembrain_b200_vega_classes.hpp"
#define ImplicitNewmarkSparse IntegratorBaseSparse
#define CorotationalLinearFEMForceModel fembrain_b200::CudaCorotationalForceModel
#define VolumeConservingIntegrator fembrain_b200::CudaVolumeConservingIntegrator2
#endif
