/*
 * deformable_harness.cpp — C wrapper (fbdef_*) around the reference's OWN `class Deformable`
 * (src/deformable/Deformable.{h,cpp}) and the mesh classes it sits on (VolMesh, CuttableMesh, VolMeshSamples,
 * SGNode/SGTransform, AABB), all compiled UNMODIFIED and in place by oracle/Makefile into oracle/_ref/libfembrain_ref.so.
 *
 * TEST INFRASTRUCTURE ONLY.  This file contains no reference code: it constructs the reference's objects the way
 * src/main.cpp does (a VolMesh, `new Deformable(mesh, fixedVertices)`, a scene node as collision object) and forwards to
 * their methods.  What made Deformable.cpp compilable without TBB / Loki / OpenGL / Bullet / OpenCL is oracle/stubs/:
 * stand-in headers for those libraries (no-op GL entry points, a thread-count query, empty Loki templates) and
 * prelude_deformable.h, which defines the include guards of the GL/OpenCL-dependent headers.  None of the arithmetic on
 * the checked path comes from a stub: Deformable::timestep, applyHapticForces, pickVertices, pickVertex,
 * CuttableMesh::findClosestVertex, VolMesh::get_node_neighbors / displace, base/Vec.h and graphics/AABB.h are the
 * reference's own code.
 *
 * C++11 on purpose (VolMesh needs it); Vega's headers are made parseable in C++11 by the one operator<< overload in the
 * prelude.  `private` is opened for this translation unit only, to read m_q / m_arrExtForces / m_ctCollided.
 */
#define private public
#define protected public
#include "Deformable.h"
#include "VolMeshSamples.h"
#undef private
#undef protected

#include <string.h>

#include <vector>

/* the same wrapper is compiled twice: fbdef_* over the reference as it is (libfembrain_ref.so), fbdrop_* over the
 * reference's Deformable.cpp compiled with the drop-in substitutions of stubs/prelude_dropin.h (libfembrain_dropin.so) */
#ifndef FBDEF
#define FBDEF(name) fbdef_##name
#endif

/* the two link-time leftovers of the GL tree (src/graphics/GLFuncs.cpp) and of the SQLite logger (DBLogger.cpp) */
void DrawAABB(const PS::MATH::AABB &, const PS::MATH::vec3f &) {}
void DrawAABB(const PS::MATH::vec3f &, const PS::MATH::vec3f &, const PS::MATH::vec3f &, float) {}
std::string DBLogger::timestamp() { return std::string(); }

namespace {
struct FloorNode : public PS::SG::SGNode {
  void draw() {}
};
struct DefSim {
  PS::MESH::VolMesh *mesh;
  Deformable *def;
  FloorNode *floor;
};
}  // namespace

extern "C" {

/* VolMesh::setup (src/deformable/VolMesh.cpp:139-163) + Deformable(const VolMesh&, const vector<int>&)
 * (src/deformable/Deformable.cpp:44-56 -> syncForceModel :127-220).  The collision object starts far below the mesh. */
void *FBDEF(create)(int nV, const double *verts, int nT, const int *tets, int nFixed, const int *fixedVerts) {
  DefSim *s = new DefSim();
  s->mesh = new PS::MESH::VolMesh();
  std::vector<U32> el(tets, tets + 4 * (size_t)nT);
  if (!s->mesh->setup((U32)nV, verts, (U32)nT, el.empty() ? NULL : &el[0])) {
    delete s->mesh;
    delete s;
    return NULL;
  }
  std::vector<int> fv(fixedVerts, fixedVerts + nFixed);
  try {
    s->def = new Deformable(*s->mesh, fv);
  } catch (int) {  /* the drop-in build reports an unusable device this way (FEMBRAIN_B200_NO_EXIT) */
    delete s->mesh;
    delete s;
    return NULL;
  }
  s->def->setGravity(false);  /* Deformable::init (Deformable.cpp:84-123) leaves m_bApplyGravity uninitialised; main.cpp sets it */
  s->floor = new FloorNode();
  s->floor->transform()->translate(vec3f(0.0f, -1.0e30f, 0.0f));
  s->floor->transform()->syncMatrices();  /* the forward matrix follows translate() only on request (SGTransform.cpp:65-70) */
  s->def->setCollisionObject(s->floor);
  return s;
}

void FBDEF(destroy)(void *p) {
  DefSim *s = (DefSim *)p;
  if (!s) return;
  delete s->def;
  delete s->floor;
  delete s->mesh;
  delete s;
}

int FBDEF(num_vertices)(void *p) { return (int)((DefSim *)p)->def->m_lpVolMesh->countNodes(); }
int FBDEF(num_cells)(void *p) { return (int)((DefSim *)p)->def->m_lpVolMesh->countCells(); }
int FBDEF(num_edges)(void *p) { return (int)((DefSim *)p)->def->m_lpVolMesh->countEdges(); }

/* the mesh as Deformable holds it (its CuttableMesh copy): rest positions, cells, and the edge array in m_vEdges order */
void FBDEF(mesh)(void *p, double *restpos, int *cells, int *edgesFromTo) {
  PS::CuttableMesh *m = ((DefSim *)p)->def->m_lpVolMesh;
  for (U32 i = 0; i < m->countNodes(); i++) {
    const vec3d r = m->const_nodeAt(i).restpos;
    if (restpos) { restpos[3 * i] = r.x; restpos[3 * i + 1] = r.y; restpos[3 * i + 2] = r.z; }
  }
  if (cells)
    for (U32 c = 0; c < m->countCells(); c++)
      for (int k = 0; k < 4; k++) cells[4 * c + k] = (int)m->const_cellAt(c).nodes[k];
  if (edgesFromTo)
    for (U32 e = 0; e < m->countEdges(); e++) {
      edgesFromTo[2 * e] = (int)m->const_edgeAt(e).from;
      edgesFromTo[2 * e + 1] = (int)m->const_edgeAt(e).to;
    }
}

/* VolMesh::get_node_neighbors (src/deformable/VolMesh.cpp:1346-1363), with its indexing quirk, as compiled */
int FBDEF(node_neighbors)(void *p, int v, int capacity, int *out) {
  std::vector<U32> nb;
  U32 n = ((DefSim *)p)->def->m_lpVolMesh->get_node_neighbors((U32)v, nb);
  for (U32 i = 0; i < n && (int)i < capacity; i++) out[i] = (int)nb[i];
  return (int)n;
}

void FBDEF(set_gravity)(void *p, int on) { ((DefSim *)p)->def->setGravity(on != 0); }
void FBDEF(set_haptic_radius)(void *p, int rings) { ((DefSim *)p)->def->setHapticForceRadius(rings); }
/* the collision object's transform: Deformable::timestep maps the origin through it and uses the y coordinate (a FLOAT) */
void FBDEF(set_floor)(void *p, float y) {
  DefSim *s = (DefSim *)p;
  s->floor->resetTransform();
  s->floor->transform()->translate(vec3f(0.0f, y, 0.0f));
  s->floor->transform()->syncMatrices();
}
float FBDEF(floor_y)(void *p) { return ((DefSim *)p)->floor->transform()->forward().map(vec3f(0, 0, 0)).y; }

/* hapticStart(index) / hapticSetCurrentForces / hapticEnd (src/deformable/Deformable.cpp:510-539, 712-717) */
void FBDEF(set_haptic)(void *p, int n, const int *idx, const double *f3, int inProgress) {
  Deformable *d = ((DefSim *)p)->def;
  if (inProgress) d->hapticStart(n > 0 ? idx[0] : -1); else d->hapticEnd();
  std::vector<int> vi(idx, idx + n);
  std::vector<vec3d> vf((size_t)n);
  for (int i = 0; i < n; i++) vf[i] = vec3d(f3[3 * i], f3[3 * i + 1], f3[3 * i + 2]);
  d->hapticSetCurrentForces(vi, vf);
}

void FBDEF(timestep)(void *p) { ((DefSim *)p)->def->timestep(); }

void FBDEF(get_state)(void *p, double *q, double *qvel, double *qacc) {
  Deformable *d = ((DefSim *)p)->def;
  const size_t bytes = sizeof(double) * d->m_dof;
  if (q) memcpy(q, d->m_q, bytes);
  if (qvel) memcpy(qvel, d->m_qVel, bytes);
  if (qacc) memcpy(qacc, d->m_qAcc, bytes);
}
void FBDEF(set_state)(void *p, const double *q, const double *qvel) {
  Deformable *d = ((DefSim *)p)->def;
  std::vector<double> zero(d->m_dof, 0.0);
  d->m_lpIntegrator->SetqState(q, qvel, &zero[0]);
}
void FBDEF(get_external_forces)(void *p, double *f) {
  Deformable *d = ((DefSim *)p)->def;
  memcpy(f, d->m_arrExtForces, sizeof(double) * d->m_dof);
}
int FBDEF(contacts)(void *p) { return (int)((DefSim *)p)->def->m_ctCollided; }
void FBDEF(set_contacts)(void *p, int n) { ((DefSim *)p)->def->m_ctCollided = (U32)n; }
/* node positions after VolMesh::displace (src/deformable/VolMesh.cpp:1370-1385) and the mesh AABB Deformable publishes */
void FBDEF(positions)(void *p, double *pos) {
  PS::CuttableMesh *m = ((DefSim *)p)->def->m_lpVolMesh;
  for (U32 i = 0; i < m->countNodes(); i++) {
    const vec3d x = m->const_nodeAt(i).pos;
    pos[3 * i] = x.x; pos[3 * i + 1] = x.y; pos[3 * i + 2] = x.z;
  }
}
void FBDEF(aabb)(void *p, float *lo3, float *hi3) {
  const PS::MATH::AABB b = ((DefSim *)p)->def->aabb();
  lo3[0] = b.lower().x; lo3[1] = b.lower().y; lo3[2] = b.lower().z;
  hi3[0] = b.upper().x; hi3[1] = b.upper().y; hi3[2] = b.upper().z;
}

/* Deformable::pickVertices (src/deformable/Deformable.cpp:430-448) and ::pickVertex (:422-428) */
int FBDEF(pick_vertices)(void *p, const double *lo, const double *hi, int capacity, int *indices, double *coords) {
  std::vector<vec3d> c;
  std::vector<int> ix;
  int n = ((DefSim *)p)->def->pickVertices(vec3d(lo[0], lo[1], lo[2]), vec3d(hi[0], hi[1], hi[2]), c, ix);
  for (int i = 0; i < n && i < capacity; i++) {
    indices[i] = ix[i];
    if (coords) { coords[3 * i] = c[i].x; coords[3 * i + 1] = c[i].y; coords[3 * i + 2] = c[i].z; }
  }
  return n;
}
int FBDEF(pick_vertex)(void *p, const double *w, double *dist, double *vertex) {
  DefSim *s = (DefSim *)p;
  vec3d v;
  double d = 0.0;
  int i = s->def->m_lpVolMesh->findClosestVertex(vec3d(w[0], w[1], w[2]), d, v);
  if (dist) *dist = d;
  if (vertex) { vertex[0] = v.x; vertex[1] = v.y; vertex[2] = v.z; }
  return i;
}

/* VolMeshSamples (src/deformable/VolMeshSamples.cpp:15-253): which = 0 one tetra, 1 two tetra, 2 truth cube (a,b,c = nx,ny,nz;
 * x = cellsize), 3 egg shell (a,b = hseg,vseg; x = radius, y = thickness).  First call with NULL outputs for the sizes. */
int FBDEF(sample_mesh)(int which, int a, int b, int c, double x, double y, int *nV, int *nT, double *verts, int *cells) {
  PS::MESH::VolMesh *m = NULL;
  switch (which) {
    case 0: m = PS::MESH::VolMeshSamples::CreateOneTetra(); break;
    case 1: m = PS::MESH::VolMeshSamples::CreateTwoTetra(); break;
    case 2: m = PS::MESH::VolMeshSamples::CreateTruthCube(a, b, c, x); break;
    case 3: m = PS::MESH::VolMeshSamples::CreateEggShell(a, b, x, y); break;
    default: return -1;
  }
  if (!m) return -1;
  *nV = (int)m->countNodes();
  *nT = (int)m->countCells();
  if (verts)
    for (U32 i = 0; i < m->countNodes(); i++) {
      const vec3d r = m->const_nodeAt(i).restpos;
      verts[3 * i] = r.x; verts[3 * i + 1] = r.y; verts[3 * i + 2] = r.z;
    }
  if (cells)
    for (U32 k = 0; k < m->countCells(); k++)
      for (int j = 0; j < 4; j++) cells[4 * k + j] = (int)m->const_cellAt(k).nodes[j];
  delete m;
  return 0;
}

}  // extern "C"
