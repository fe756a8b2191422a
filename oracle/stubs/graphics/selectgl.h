/* oracle/stubs — stand-in for src/graphics/selectgl.h (OpenGL / GLEW / GLUT are not installed): the GL entry points the
 * compiled reference files mention, as no-ops.  Drawing code is compiled but never called by the checkers. */
#ifndef FB_STUB_SELECTGL_H
#define FB_STUB_SELECTGL_H
#include "base/MathBase.h"
typedef unsigned int GLenum; typedef unsigned int GLuint; typedef int GLint; typedef int GLsizei; typedef float GLfloat;
typedef double GLdouble; typedef unsigned char GLubyte; typedef unsigned int GLbitfield; typedef void GLvoid;
#define FB_GL_NOP(name) template <typename... A> static inline void name(A...) {}
FB_GL_NOP(glPushAttrib) FB_GL_NOP(glPopAttrib) FB_GL_NOP(glEnable) FB_GL_NOP(glDisable) FB_GL_NOP(glBegin) FB_GL_NOP(glEnd)
FB_GL_NOP(glVertex3dv) FB_GL_NOP(glVertex3fv) FB_GL_NOP(glVertex3d) FB_GL_NOP(glVertex3f) FB_GL_NOP(glColor3f) FB_GL_NOP(glColor3d)
FB_GL_NOP(glColor4f) FB_GL_NOP(glColor4fv) FB_GL_NOP(glColor3fv) FB_GL_NOP(glColor3dv) FB_GL_NOP(glColor4d) FB_GL_NOP(glLineWidth) FB_GL_NOP(glPointSize) FB_GL_NOP(glPolygonOffset)
FB_GL_NOP(glPolygonMode) FB_GL_NOP(glPushMatrix) FB_GL_NOP(glPopMatrix) FB_GL_NOP(glTranslated) FB_GL_NOP(glTranslatef)
FB_GL_NOP(glMultMatrixf) FB_GL_NOP(glMultMatrixd) FB_GL_NOP(glNormal3dv) FB_GL_NOP(glNormal3fv) FB_GL_NOP(glBlendFunc) FB_GL_NOP(glDepthMask)
FB_GL_NOP(glMatrixMode) FB_GL_NOP(glLoadIdentity) FB_GL_NOP(glLoadMatrixf) FB_GL_NOP(glGetFloatv) FB_GL_NOP(glGetDoublev) FB_GL_NOP(glScalef) FB_GL_NOP(glScaled) FB_GL_NOP(glRasterPos3f) FB_GL_NOP(glRasterPos3d) FB_GL_NOP(glutBitmapCharacter) FB_GL_NOP(glutSolidSphere) FB_GL_NOP(glutWireSphere)
enum { GL_ALL_ATTRIB_BITS = 0, GL_POLYGON_OFFSET_POINT, GL_POLYGON_OFFSET_FILL, GL_POLYGON_OFFSET_LINE, GL_POINTS, GL_LINES, GL_TRIANGLES, GL_LINE_LOOP, GL_LINE_STRIP,
       GL_QUADS, GL_LIGHTING, GL_FRONT_AND_BACK, GL_LINE, GL_FILL, GL_BLEND, GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA, GL_DEPTH_TEST, GL_TRUE, GL_FALSE,
       GL_CULL_FACE, GL_FRONT, GL_BACK, GL_POINT, GL_LINE_SMOOTH, GL_POINT_SMOOTH, GL_ENABLE_BIT, GL_CURRENT_BIT, GL_LINE_BIT, GL_POLYGON_BIT, GL_TRIANGLE_STRIP, GL_QUAD_STRIP, GL_MODELVIEW_MATRIX, GL_MODELVIEW, GL_PROJECTION, GL_PROJECTION_MATRIX, GL_TRIANGLE_FAN, GL_POLYGON };
static void *const GLUT_BITMAP_8_BY_13 = 0;
static void *const GLUT_BITMAP_HELVETICA_12 = 0;
#endif
