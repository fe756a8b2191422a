/* oracle/stubs/prelude_volmesh.h — TEST INFRASTRUCTURE ONLY; forced first (-include) when oracle/Makefile compiles the
 * reference's OWN mesh sources (src/deformable/VolMesh.cpp, ...) in place.  Headers that the reference includes with a
 * quoted, same-directory path cannot be shadowed through -I, so their include guards are defined here and the few
 * declarations the compiled files need from them are supplied by the stubs.  Contains no reference code. */
#ifndef FB_STUB_PRELUDE_VOLMESH_H
#define FB_STUB_PRELUDE_VOLMESH_H
/* standard headers the real Logger / SceneGraph headers used to bring in transitively */
#include <map>
#include <set>
#include <vector>
#include <string>
#include <list>
#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>
#include <stdio.h>
#define hifem_SGEffect_h       /* src/graphics/SGEffect.h (ShaderManager -> Loki, GL) */
#define SELECTGL_H_           /* src/graphics/selectgl.h (GL/glew.h, GL/freeglut.h) */
#include "graphics/selectgl.h"
#include "graphics/SGEffect.h" /* resolves to oracle/stubs/graphics/SGEffect.h (stubs come first on the include path) */
#endif
