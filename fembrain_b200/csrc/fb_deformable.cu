// fb_deformable.cu — the caller either side of the integrator step: Deformable::timestep
// (reference src/deformable/Deformable.cpp:318-420) — external-force build (gravity :331-338, haptic forces with
// ring spreading applyHapticForces :634-706) before DoTimestep, floor-plane post-step (:350-402) after it.
//
// The force build runs on the device (k_force_base, k_haptic_spread): nothing of size r is built on or copied from the
// host per frame (41 MB at 10M tets).  The rings of applyHapticForces are breadth-first levels from each haptic vertex;
// one CTA walks the haptic vertices IN ORDER, so that every entry of the force vector receives its additions in the
// reference's order (direct forces by index, then per haptic vertex ring by ring) and the result is bit-identical to the
// reference's std::set arithmetic; inside one ring the new vertices are distinct, so their additions run in parallel.
// The post-step is one kernel over the vertices.  Neighbour rings: the reference walks VolMesh::get_node_neighbors
// (DEF/VolMesh.cpp:1346-1363), which indexes the GLOBAL edge array with a loop counter that runs over the node's
// incident-edge COUNT (const_edgeAt(i) instead of const_edgeAt(edges[i])) — so its "neighbours" of v are those
// among the first deg(v) edges of the mesh that touch v.  Two modes are offered and the choice is explicit:
//   default ............. true mesh adjacency (vertices sharing a tetrahedron edge), from the K block structure
//   reference quirk ..... bit-for-bit the reference's behaviour, given the host's edge array in VolMesh order
//                         (fb_deformable_set_edge_list; the FemBrain host has it as m_lpVolMesh->const_edgeAt(i))
#include <cub/cub.cuh>

#include <algorithm>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fb_internal.h"

namespace {

#define CHECK_CTX(c)                                                            \
  do {                                                                          \
    if (!(c)) { fb_set_error("NULL context"); return FB_ERR_INVALID_ARGUMENT; } \
    cudaError_t e_ = cudaSetDevice((c)->device);                                \
    if (e_ != cudaSuccess) { fb_set_error("cudaSetDevice(%d): %s", (c)->device, cudaGetErrorString(e_)); return FB_ERR_CUDA; } \
  } while (0)

// Deformable::timestep post-step (DEF/Deformable.cpp:350-402): count contacts (pc.y <= c.y), then for EVERY node
// v <- (v - v_n) - 0.4 v_n with n = (0,1,0), accelerations <- 0, and penetrating nodes snapped onto the plane.
// Vec3 arithmetic spelled out as base/Vec.h does it (dot = x*x + y*y + z*z, :472-475).
__global__ void k_floor_poststep(int nV, double floorY, const double *__restrict__ x0, double *__restrict__ q,
                                 double *__restrict__ qvel, double *__restrict__ qaccel, int *__restrict__ contacts) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nV) return;
  const double pry = x0[3 * (size_t)i + 1];
  const double qy = q[3 * (size_t)i + 1];
  const double pcy = pry + qy;
  const double vx = qvel[3 * (size_t)i], vy = qvel[3 * (size_t)i + 1], vz = qvel[3 * (size_t)i + 2];
  const double nx = 0.0, ny = 1.0, nz = 0.0;
  const double dotvn = vx * nx + vy * ny + vz * nz;
  const double vnx = nx * dotvn, vny = ny * dotvn, vnz = nz * dotvn;
  const double vpx = vx - vnx, vpy = vy - vny, vpz = vz - vnz;
  qvel[3 * (size_t)i] = vpx - vnx * 0.4;
  qvel[3 * (size_t)i + 1] = vpy - vny * 0.4;
  qvel[3 * (size_t)i + 2] = vpz - vnz * 0.4;
  qaccel[3 * (size_t)i] = 0.0;
  qaccel[3 * (size_t)i + 1] = 0.0;
  qaccel[3 * (size_t)i + 2] = 0.0;
  if (pcy <= floorY) {
    atomicAdd(contacts, 1);  // integer count
    q[3 * (size_t)i + 1] = floorY - pry;
  }
}

// Deformable::pickVertices (DEF/Deformable.cpp:430-448) with Contains<double> (graphics/AABB.h:84-92, closed box) on the
// current node positions pos = restpos + u (VolMesh::displace, DEF/VolMesh.cpp:1370-1385): one flag per vertex, then an
// order-preserving compaction so the indices come out ascending like the reference's push_back loop.
__global__ void k_pick_flags(int nV, const double *__restrict__ x0, const double *__restrict__ q, double lx, double ly, double lz,
                             double hx, double hy, double hz, int *__restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nV) return;
  const double px = x0[3 * (size_t)i] + q[3 * (size_t)i], py = x0[3 * (size_t)i + 1] + q[3 * (size_t)i + 1],
               pz = x0[3 * (size_t)i + 2] + q[3 * (size_t)i + 2];
  flag[i] = ((px >= lx) && (px <= hx) && (py >= ly) && (py <= hy) && (pz >= lz) && (pz <= hz)) ? 1 : 0;
}
__global__ void k_pick_scatter(int nV, const double *__restrict__ x0, const double *__restrict__ q, const int *__restrict__ flag,
                               const int *__restrict__ pos, int capacity, int *__restrict__ idx, double *__restrict__ coords) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nV || !flag[i]) return;
  const int o = pos[i];
  if (o >= capacity) return;
  idx[o] = i;
  if (coords)
    for (int k = 0; k < 3; k++) coords[3 * (size_t)o + k] = x0[3 * (size_t)i + k] + q[3 * (size_t)i + k];
}

// CuttableMesh::findClosestVertex (DEF/CuttableMesh.cpp:511-526): squared distance (dx*dx + dy*dy + dz*dz, Vec3::length2)
// to the current position, strict '<' in index order => the LOWEST index among equal minima.
struct PickBest {
  double d2;
  int idx;
};
__device__ __forceinline__ PickBest pick_min(PickBest a, PickBest b) {
  return (b.d2 < a.d2 || (b.d2 == a.d2 && b.idx < a.idx)) ? b : a;
}
__global__ void __launch_bounds__(256) k_pick_closest(int nV, const double *__restrict__ x0, const double *__restrict__ q, double wx,
                                                      double wy, double wz, PickBest *__restrict__ out) {
  __shared__ PickBest sm[256];
  PickBest best;
  best.d2 = DBL_MAX;  // GetMaxLimit<double>()
  best.idx = 0x7fffffff;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < nV; i += gridDim.x * 256) {
    const double dx = wx - (x0[3 * (size_t)i] + q[3 * (size_t)i]), dy = wy - (x0[3 * (size_t)i + 1] + q[3 * (size_t)i + 1]),
                 dz = wz - (x0[3 * (size_t)i + 2] + q[3 * (size_t)i + 2]);
    PickBest c;
    c.d2 = dx * dx + dy * dy + dz * dz;
    c.idx = i;
    if (c.d2 < DBL_MAX) best = pick_min(best, c);  // the reference never accepts a distance that is not below the limit (NaN, inf)
  }
  sm[threadIdx.x] = best;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] = pick_min(sm[threadIdx.x], sm[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

// SetExternalForcesToZero + memset + gravity (DEF/Deformable.cpp:325-338): 0, or 0 + (-10000) on the y components
__global__ void k_force_base(size_t r, int gravity, double *__restrict__ f) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < r; i += (size_t)gridDim.x * blockDim.x) {
    double v = 0.0;
    if (gravity && (i % 3 == 1)) v += -10000.0;
    f[i] = v;
  }
}

struct HapticArgs {
  int nH, rings, nV;
  const int *idx;          // [nH] haptic vertices
  const double *force;     // [3 nH]
  const int *bp, *bc;      // true adjacency: the block structure of K (vertices sharing a tet), self excluded
  int quirk, nEdges;       // reference rule: VolMesh::get_node_neighbors on the host's edge array
  const int *edges, *deg;
  unsigned int *stamp;     // [nV] visit stamps, never cleared: stampBase + haptic index + 1 marks "affected by this haptic vertex"
  unsigned int stampBase;
  int *listA, *listB;      // [nV] frontier lists
};

// applyHapticForces (DEF/Deformable.cpp:634-706).  ONE CTA.
__global__ void __launch_bounds__(1024) k_haptic_spread(HapticArgs a, double *__restrict__ f) {
  __shared__ int nLast, nFresh;
  // direct forces, in index order (a vertex may be listed twice)                                   (:641-647)
  if (threadIdx.x < 3)
    for (int i = 0; i < a.nH; i++) f[3 * (size_t)a.idx[i] + threadIdx.x] += a.force[3 * (size_t)i + threadIdx.x];
  __syncthreads();
  if (a.rings <= 1) return;
  for (int iv = 0; iv < a.nH; iv++) {
    const unsigned int cur = a.stampBase + (unsigned int)iv + 1u;
    const double fx = a.force[3 * (size_t)iv], fy = a.force[3 * (size_t)iv + 1], fz = a.force[3 * (size_t)iv + 2];
    int *last = a.listA, *fresh = a.listB;
    if (threadIdx.x == 0) {
      a.stamp[a.idx[iv]] = cur;   // affectedVertices = lastLayerVertices = { the haptic vertex }
      last[0] = a.idx[iv];
      nLast = 1;
    }
    __syncthreads();
    for (int j = 1; j < a.rings; j++) {
      // linear kernel (:657-658): 1.0 * (size - j) / static_cast<double>(size)
      const double mag = 1.0 * (a.rings - j) / static_cast<double>(a.rings);
      if (threadIdx.x == 0) nFresh = 0;
      __syncthreads();
      const int nl = nLast;
      // every neighbour of the last ring that is not yet affected joins the new ring once (atomicMax = test-and-set)
      for (int e = threadIdx.x; e < nl; e += blockDim.x) {
        const int v = last[e];
        if (a.quirk) {
          // VolMesh::get_node_neighbors (DEF/VolMesh.cpp:1346-1363): for (i = 0; i < incident_edges(v).size(); i++) e = const_edgeAt(i)
          const int d = min(a.deg[v], a.nEdges);
          for (int i = 0; i < d; i++) {
            const int from = a.edges[2 * i], to = a.edges[2 * i + 1];
            const int nb = (from == v) ? to : ((to == v) ? from : -1);
            if (nb >= 0 && atomicMax(&a.stamp[nb], cur) < cur) fresh[atomicAdd(&nFresh, 1)] = nb;
          }
        } else {
          for (int p = a.bp[v]; p < a.bp[v + 1]; p++) {
            const int nb = a.bc[p];
            if (nb != v && atomicMax(&a.stamp[nb], cur) < cur) fresh[atomicAdd(&nFresh, 1)] = nb;
          }
        }
      }
      __syncthreads();
      const int nf = nFresh;
      for (int e = threadIdx.x; e < nf; e += blockDim.x) {   // distinct vertices: one addition each (:687-690)
        const size_t v = (size_t)fresh[e];
        f[3 * v] += mag * fx;
        f[3 * v + 1] += mag * fy;
        f[3 * v + 2] += mag * fz;
      }
      __syncthreads();
      if (threadIdx.x == 0) nLast = nf;
      int *t = last; last = fresh; fresh = t;
      __syncthreads();
    }
  }
}

}  // namespace

extern "C" {

int fb_deformable_set_gravity(fb_context *c, int enabled) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  c->gravity = enabled != 0;
  return FB_OK;
}
int fb_deformable_set_floor(fb_context *c, int enabled, double y) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  c->floor_enabled = enabled != 0;
  c->floor_y = y;
  return FB_OK;
}
int fb_deformable_set_haptic_forces(fb_context *c, int count, const int *idx, const double *forces, int inProgress) {
  if (!c || count < 0 || (count > 0 && (!idx || !forces))) return FB_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < count; i++)
    if (idx[i] < 0 || idx[i] >= c->nV) { fb_set_error("haptic vertex %d out of range", idx[i]); return FB_ERR_INVALID_ARGUMENT; }
  free(c->haptic_idx_host);
  free(c->haptic_f_host);
  c->haptic_idx_host = (int *)malloc(sizeof(int) * (size_t)(count ? count : 1));
  c->haptic_f_host = (double *)malloc(sizeof(double) * 3 * (size_t)(count ? count : 1));
  memcpy(c->haptic_idx_host, idx, sizeof(int) * (size_t)count);
  memcpy(c->haptic_f_host, forces, sizeof(double) * 3 * (size_t)count);
  c->nHaptic = count;
  c->haptic_in_progress = inProgress != 0;
  // device copies for the force-build kernel
  if (cudaSetDevice(c->device) != cudaSuccess) { cudaGetLastError(); return FB_ERR_CUDA; }
  if (c->haptic_idx_dev) { fb_dev_free(c->haptic_idx_dev); c->haptic_idx_dev = nullptr; }
  if (c->haptic_f_dev) { fb_dev_free(c->haptic_f_dev); c->haptic_f_dev = nullptr; }
  FB_TRY(fb_dev_alloc(c, &c->haptic_idx_dev, (size_t)count));
  FB_TRY(fb_dev_alloc(c, &c->haptic_f_dev, 3 * (size_t)count));
  if (count) {
    FB_CUDA(cudaMemcpyAsync(c->haptic_idx_dev, idx, sizeof(int) * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaMemcpyAsync(c->haptic_f_dev, forces, sizeof(double) * 3 * (size_t)count, cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return FB_OK;
}
int fb_deformable_set_haptic_neighborhood(fb_context *c, int rings) {
  if (!c || rings < 0) return FB_ERR_INVALID_ARGUMENT;
  c->haptic_rings = rings;
  return FB_OK;
}
int fb_deformable_set_edge_list(fb_context *c, int numEdges, const int *fromTo, int referenceQuirk) {
  if (!c || numEdges < 0 || (numEdges > 0 && !fromTo)) return FB_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < 2 * numEdges; i++)
    if (fromTo[i] < 0 || fromTo[i] >= c->nV) { fb_set_error("edge endpoint %d out of range", fromTo[i]); return FB_ERR_INVALID_ARGUMENT; }
  free(c->edges_host);
  free(c->edge_degree_host);
  c->edges_host = nullptr;
  c->edge_degree_host = nullptr;
  c->nEdges = numEdges;
  c->haptic_quirk = (referenceQuirk != 0) && numEdges > 0;
  if (numEdges > 0) {
    c->edges_host = (int *)malloc(sizeof(int) * 2 * (size_t)numEdges);
    memcpy(c->edges_host, fromTo, sizeof(int) * 2 * (size_t)numEdges);
    c->edge_degree_host = (int *)calloc((size_t)(c->nV ? c->nV : 1), sizeof(int));
    for (int i = 0; i < numEdges; i++) {  // m_incident_edges_per_node: both endpoints (VolMesh.cpp:580-581)
      c->edge_degree_host[fromTo[2 * i]]++;
      c->edge_degree_host[fromTo[2 * i + 1]]++;
    }
  }
  if (cudaSetDevice(c->device) != cudaSuccess) { cudaGetLastError(); return FB_ERR_CUDA; }
  if (c->edges_dev) { fb_dev_free(c->edges_dev); c->edges_dev = nullptr; }
  if (c->edge_degree_dev) { fb_dev_free(c->edge_degree_dev); c->edge_degree_dev = nullptr; }
  if (numEdges > 0) {
    FB_TRY(fb_dev_alloc(c, &c->edges_dev, 2 * (size_t)numEdges));
    FB_TRY(fb_dev_alloc(c, &c->edge_degree_dev, (size_t)c->nV));
    FB_CUDA(cudaMemcpyAsync(c->edges_dev, c->edges_host, sizeof(int) * 2 * (size_t)numEdges, cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaMemcpyAsync(c->edge_degree_dev, c->edge_degree_host, sizeof(int) * (size_t)c->nV, cudaMemcpyHostToDevice, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return FB_OK;
}
int fb_deformable_contact_count(const fb_context *c) { return c ? c->contact_count : 0; }

int fb_deformable_pick_vertices(fb_context *c, const double *lo, const double *hi, int capacity, int *indices, double *coords,
                                int *count) {
  CHECK_CTX(c);
  if (!lo || !hi || !count || capacity < 0 || (capacity > 0 && !indices)) { fb_set_error("bad arguments to fb_deformable_pick_vertices"); return FB_ERR_INVALID_ARGUMENT; }
  if (c->dist) { fb_set_error("picking on a partitioned context is not supported"); return FB_ERR_NOT_SUPPORTED; }
  *count = 0;
  const int nV = c->nV;
  if (nV == 0) return FB_OK;
  cudaStream_t st = c->stream;
  int *flag = nullptr, *pos = nullptr, *dIdx = nullptr;
  double *dCo = nullptr;
  void *tmp = nullptr;
  size_t tmpBytes = 0;
  int status = FB_OK;
  auto cleanup = [&]() { for (void *q : {(void *)flag, (void *)pos, (void *)dIdx, (void *)dCo, tmp}) fb_tmp_free(st, q); };
#define PK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fb_set_error("%s -> %s", #call, cudaGetErrorString(e__)); cleanup(); return FB_ERR_CUDA; } } while (0)
  PK(fb_tmp_alloc(st, &flag, sizeof(int) * ((size_t)nV + 1)));
  PK(fb_tmp_alloc(st, &pos, sizeof(int) * ((size_t)nV + 1)));
  PK(cub::DeviceScan::ExclusiveSum(nullptr, tmpBytes, flag, pos, (int64_t)nV + 1, st));
  PK(fb_tmp_alloc(st, &tmp, tmpBytes));
  PK(cudaMemsetAsync(flag + nV, 0, sizeof(int), st));
  k_pick_flags<<<(nV + 255) / 256, 256, 0, st>>>(nV, c->x0, c->q, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], flag);
  PK(cub::DeviceScan::ExclusiveSum(tmp, tmpBytes, flag, pos, (int64_t)nV + 1, st));
  int found = 0;
  PK(cudaMemcpyAsync(&found, pos + nV, sizeof(int), cudaMemcpyDeviceToHost, st));
  PK(cudaStreamSynchronize(st));
  c->launches += 3;
  *count = found;  // the full count even when it exceeds the caller's capacity (call again with a larger buffer)
  const int n = found < capacity ? found : capacity;
  if (n > 0) {
    PK(fb_tmp_alloc(st, &dIdx, sizeof(int) * (size_t)n));
    if (coords) PK(fb_tmp_alloc(st, &dCo, sizeof(double) * 3 * (size_t)n));
    k_pick_scatter<<<(nV + 255) / 256, 256, 0, st>>>(nV, c->x0, c->q, flag, pos, n, dIdx, dCo);
    c->launches++;
    PK(cudaMemcpyAsync(indices, dIdx, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
    if (coords) PK(cudaMemcpyAsync(coords, dCo, sizeof(double) * 3 * (size_t)n, cudaMemcpyDeviceToHost, st));
    PK(cudaStreamSynchronize(st));
  }
  cleanup();
  return status;
}

int fb_deformable_pick_vertex(fb_context *c, const double *wpos, int *index, double *dist, double *vertex) {
  CHECK_CTX(c);
  if (!wpos || !index) { fb_set_error("bad arguments to fb_deformable_pick_vertex"); return FB_ERR_INVALID_ARGUMENT; }
  if (c->dist) { fb_set_error("picking on a partitioned context is not supported"); return FB_ERR_NOT_SUPPORTED; }
  *index = -1;
  if (dist) *dist = sqrt(DBL_MAX);
  const int nV = c->nV;
  if (nV == 0) return FB_OK;
  cudaStream_t st = c->stream;
  int grid = (nV + 255) / 256;
  if (grid > 4 * c->sm_count) grid = 4 * c->sm_count;
  PickBest *part = nullptr;
  void *tmpv = nullptr;
  auto cleanup = [&]() { fb_tmp_free(st, part); fb_tmp_free(st, tmpv); };
  PK(fb_tmp_alloc(st, &part, sizeof(PickBest) * ((size_t)grid + 1)));
  k_pick_closest<<<grid, 256, 0, st>>>(nV, c->x0, c->q, wpos[0], wpos[1], wpos[2], part);
  c->launches++;
  std::vector<PickBest> h((size_t)grid);
  PK(cudaMemcpyAsync(h.data(), part, sizeof(PickBest) * (size_t)grid, cudaMemcpyDeviceToHost, st));
  PK(cudaStreamSynchronize(st));
  PickBest best = h[0];
  for (int i = 1; i < grid; i++)
    if (h[i].d2 < best.d2 || (h[i].d2 == best.d2 && h[i].idx < best.idx)) best = h[i];
  cleanup();
#undef PK
  if (best.idx == 0x7fffffff) return FB_OK;  // no vertex at a finite distance: the reference returns -1 too
  *index = best.idx;
  if (dist) *dist = sqrt(best.d2);
  if (vertex) {
    double x[3], u[3];
    FB_CUDA(cudaMemcpyAsync(x, c->x0 + 3 * (size_t)best.idx, sizeof(x), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaMemcpyAsync(u, c->q + 3 * (size_t)best.idx, sizeof(u), cudaMemcpyDeviceToHost, st));
    FB_CUDA(cudaStreamSynchronize(st));
    for (int k = 0; k < 3; k++) vertex[k] = x[k] + u[k];
  }
  return FB_OK;
}

int fb_deformable_timestep(fb_context *c) {
  CHECK_CTX(c);
  if (c->dist) { fb_set_error("fb_deformable_timestep on a partitioned context: build the force vector on the host and call fb_step"); return FB_ERR_NOT_SUPPORTED; }
  const size_t r = (size_t)c->r;
  cudaStream_t st = c->stream;
  // SetExternalForcesToZero + memset(m_arrExtForces) + gravity (applyGravity = m_bApplyGravity && m_ctCollided == 0)   (:325-338)
  if (r) {
    int grid = (int)((r + 255) / 256);
    if (grid > 8 * c->sm_count) grid = 8 * c->sm_count;
    k_force_base<<<grid, 256, 0, st>>>(r, (c->gravity && c->contact_count == 0) ? 1 : 0, c->fext);
    c->launches++;
  }
  // applyHapticForces                                                                                                    (:634-706)
  if (c->nHaptic > 0 && c->haptic_in_progress) {
    if (!c->haptic_stamp) {
      FB_TRY(fb_dev_alloc(c, &c->haptic_stamp, (size_t)c->nV));
      FB_TRY(fb_dev_alloc(c, &c->haptic_listA, (size_t)c->nV));
      FB_TRY(fb_dev_alloc(c, &c->haptic_listB, (size_t)c->nV));
      FB_CUDA(cudaMemsetAsync(c->haptic_stamp, 0, sizeof(unsigned int) * (size_t)(c->nV ? c->nV : 1), st));
      c->haptic_stamp_base = 0;
    }
    if (c->haptic_stamp_base > 0xffffffffu - (unsigned int)c->nHaptic - 2u) {   // stamps would wrap: start over
      FB_CUDA(cudaMemsetAsync(c->haptic_stamp, 0, sizeof(unsigned int) * (size_t)(c->nV ? c->nV : 1), st));
      c->haptic_stamp_base = 0;
    }
    HapticArgs a;
    a.nH = c->nHaptic; a.rings = c->haptic_rings; a.nV = c->nV;
    a.idx = c->haptic_idx_dev; a.force = c->haptic_f_dev;
    a.bp = c->bp; a.bc = c->bc;
    a.quirk = (c->haptic_quirk && c->edges_dev) ? 1 : 0;
    a.nEdges = c->nEdges; a.edges = c->edges_dev; a.deg = c->edge_degree_dev;
    a.stamp = c->haptic_stamp; a.stampBase = c->haptic_stamp_base;
    a.listA = c->haptic_listA; a.listB = c->haptic_listB;
    k_haptic_spread<<<1, 1024, 0, st>>>(a, c->fext);
    c->launches++;
    c->haptic_stamp_base += (unsigned int)c->nHaptic + 1u;
  }
  FB_CUDA(cudaGetLastError());
  FB_TRY(fb_do_step(c));
  if (c->floor_enabled) {
    FB_CUDA(cudaMemsetAsync(c->contact_dev, 0, sizeof(int), c->stream));
    if (c->nV) {
      k_floor_poststep<<<(c->nV + 255) / 256, 256, 0, c->stream>>>(c->nV, c->floor_y, c->x0, c->q, c->qvel, c->qaccel, c->contact_dev);
      c->launches++;
    }
    FB_CUDA(cudaMemcpyAsync(&c->contact_count, c->contact_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    FB_CUDA(cudaStreamSynchronize(c->stream));
  }
  return FB_OK;
}

}  // extern "C"
