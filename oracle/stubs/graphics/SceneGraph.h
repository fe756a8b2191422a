/* oracle/stubs — stand-in for src/graphics/SceneGraph.h (scene graph singleton, GL, Loki).  The mesh sources only ask it for
 * the camera position inside drawing code, which the checkers never call. */
#ifndef FB_STUB_SCENEGRAPH_H
#define FB_STUB_SCENEGRAPH_H
#include "graphics/SGNode.h"
namespace PS { namespace SG {
struct FbStubCamera { PS::MATH::vec3f getPos() const { return PS::MATH::vec3f(0, 0, 0); } };
struct FbStubSceneGraph { FbStubCamera camera() const { return FbStubCamera(); } };
struct TheSceneGraph { static FbStubSceneGraph &Instance() { static FbStubSceneGraph g; return g; } };
} }
#endif
