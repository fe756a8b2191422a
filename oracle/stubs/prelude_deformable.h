/* oracle/stubs/prelude_deformable.h — TEST INFRASTRUCTURE ONLY; forced first (-include) when oracle/Makefile compiles the
 * reference's OWN src/deformable/Deformable.cpp in place (C++11, -fpermissive).  Defines the include guards of headers that
 * need GL / OpenCL / Bullet and supplies the few declarations Deformable.{h,cpp} use from them.  No reference code. */
#ifndef FB_STUB_PRELUDE_DEFORMABLE_H
#define FB_STUB_PRELUDE_DEFORMABLE_H
#include <iostream>
/* Vega's mat3d.h:404 writes `s << ... << std::endl << s << ...`: valid before C++11 through ostream's conversion to void*,
 * gone since.  An overload that accepts it keeps the header parseable in the C++11 translation unit. */
inline std::ostream &operator<<(std::ostream &a, const std::ostream &) { return a; }
#include "prelude_volmesh.h"
#include "../ref_prelude.h"
#define SG_MESH_H                    /* src/graphics/SGMesh.h -> GLMeshBuffer (GL buffers) */
#define PS_SURFACEMESH_H             /* src/deformable/SurfaceMesh.h */
#define OCLVOLCONSERVEDINTEGRATOR_H_ /* src/deformable/OclVolConservedIntegrator.h -> ViennaCL / OpenCL */
#define _SCENEOBJECTDEFORMABLE_H_    /* vegafem sceneObjectDeformable.h -> GL */
#define SGBULLETCDSHAPE_H_           /* src/deformable/SGBulletCDShape.h -> Bullet */
#define VOLMESHRENDER_H_             /* src/deformable/VolMeshRender.h -> SGMesh */
#include "graphics/SGNode.h"
namespace PS { namespace GL {} namespace SG {
/* what Deformable needs of its base class: an SGNode with an empty draw() */
class SGMesh : public SGNode { public: SGMesh() {} virtual ~SGMesh() {} virtual void draw() {} virtual void drawNoEffect() {} };
} }
namespace PS { namespace FEM {} }
using namespace PS::GL;
/* src/graphics/GLFuncs.h: the two box-drawing helpers Deformable::draw / CuttableMesh::draw call; defined as no-ops in
 * oracle/deformable_harness.cpp */
void DrawAABB(const PS::MATH::AABB &box, const PS::MATH::vec3f &color = PS::MATH::vec3f(0, 0, 1));
void DrawAABB(const PS::MATH::vec3f &lo, const PS::MATH::vec3f &hi, const PS::MATH::vec3f &color, float lineWidth = 1.0f);
#endif
