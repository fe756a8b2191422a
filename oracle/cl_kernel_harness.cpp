/*
 * cl_kernel_harness.cpp — runs the reference's OpenCL kernel ApplyVertexDeformations (data/opencl/Polygonizer.cl:1417-1427)
 * on the CPU.  TEST INFRASTRUCTURE ONLY; contains no reference code: oracle/Makefile extracts the kernel's text from the
 * reference tree at build time into oracle/_ref/obj/apply_vertex_deformations.inc (a build output, never committed), and this
 * file supplies the handful of OpenCL C names the kernel uses (__kernel, __global, float4 with componentwise +,
 * get_global_id) so that the text compiles as C++.  fbcl_apply_fem_displacements adds the host half of the hand-off the way
 * GPUPoly::applyFemDisplacements does it (src/implicit/OclPolygonizer.cpp:1557-1565: doubles converted to float, w = 0).
 */
#include <stddef.h>

#include <vector>

typedef unsigned int U32;
struct float4 {
  float x, y, z, w;
};
static inline float4 operator+(const float4 &a, const float4 &b) {  /* OpenCL C: componentwise single-precision add */
  float4 r;
  r.x = a.x + b.x; r.y = a.y + b.y; r.z = a.z + b.z; r.w = a.w + b.w;
  return r;
}
#define __kernel static
#define __global
static int g_work_item = 0;
static inline int get_global_id(int) { return g_work_item; }

#include "apply_vertex_deformations.inc"

extern "C" {
/* the kernel over a 1-D range of `range` work items (>= ctVertices, as ComputeGlobalIndexSpace rounds it up) */
void fbcl_apply_vertex_deformations(U32 ctVertices, U32 range, const float *rest4, const float *disp4, float *out4) {
  for (g_work_item = 0; g_work_item < (int)range; g_work_item++)
    ApplyVertexDeformations(ctVertices, (float4 *)rest4, (float4 *)disp4, (float4 *)out4);
}
/* GPUPoly::applyFemDisplacements: repack, then the kernel */
void fbcl_apply_fem_displacements(U32 ctVertices, const float *rest4, const double *displacements, float *out4) {
  std::vector<float> h(4 * (size_t)ctVertices + 4);
  for (U32 i = 0; i < ctVertices; i++) {
    h[i * 4] = displacements[i * 3];
    h[i * 4 + 1] = displacements[i * 3 + 1];
    h[i * 4 + 2] = displacements[i * 3 + 2];
    h[i * 4 + 3] = 0;
  }
  fbcl_apply_vertex_deformations(ctVertices, ((ctVertices + 63) / 64) * 64, rest4, &h[0], out4);
}
}
