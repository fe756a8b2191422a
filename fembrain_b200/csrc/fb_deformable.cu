// fb_deformable.cu — the caller either side of the integrator step: Deformable::timestep
// (reference src/deformable/Deformable.cpp:318-420) — external-force build (gravity :331-338, haptic forces with
// ring spreading applyHapticForces :634-706) before DoTimestep, floor-plane post-step (:350-402) after it.
//
// The force build is O(picked vertices x rings) set arithmetic on the host, exactly as in the reference; the
// post-step is one kernel over the vertices.  Neighbour rings: the reference walks VolMesh::get_node_neighbors
// (DEF/VolMesh.cpp:1346-1363), which indexes the GLOBAL edge array with a loop counter that runs over the node's
// incident-edge COUNT (const_edgeAt(i) instead of const_edgeAt(edges[i])) — so its "neighbours" of v are those
// among the first deg(v) edges of the mesh that touch v.  Two modes are offered and the choice is explicit:
//   default ............. true mesh adjacency (vertices sharing a tetrahedron edge), from the K block structure
//   reference quirk ..... bit-for-bit the reference's behaviour, given the host's edge array in VolMesh order
//                         (fb_deformable_set_edge_list; the FemBrain host has it as m_lpVolMesh->const_edgeAt(i))
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <set>
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

// neighbours of vtx under the chosen rule
void node_neighbors(const fb_context *c, int vtx, std::vector<int> &out) {
  out.clear();
  if (c->haptic_quirk && c->edges_host) {
    // VolMesh::get_node_neighbors: for (i = 0; i < incident_edges(vtx).size(); i++) { e = const_edgeAt(i); ... }
    const int deg = c->edge_degree_host[vtx];
    for (int i = 0; i < deg && i < c->nEdges; i++) {
      const int from = c->edges_host[2 * i], to = c->edges_host[2 * i + 1];
      if (from == vtx) out.push_back(to);
      else if (to == vtx) out.push_back(from);
    }
    return;
  }
  for (int p = c->adj_host_bp[vtx]; p < c->adj_host_bp[vtx + 1]; p++)
    if (c->adj_host_bc[p] != vtx) out.push_back(c->adj_host_bc[p]);
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
  return FB_OK;
}
int fb_deformable_contact_count(const fb_context *c) { return c ? c->contact_count : 0; }

int fb_deformable_timestep(fb_context *c) {
  CHECK_CTX(c);
  if (c->dist) { fb_set_error("fb_deformable_timestep on a partitioned context: build the force vector on the host and call fb_step"); return FB_ERR_NOT_SUPPORTED; }
  const size_t r = (size_t)c->r;
  if (!c->fext_host) FB_CUDA(cudaMallocHost(&c->fext_host, sizeof(double) * (r ? r : 1)));
  double *f = c->fext_host;
  // SetExternalForcesToZero + memset(m_arrExtForces)                                    (:325-328)
  memset(f, 0, sizeof(double) * r);
  // gravity: applyGravity = m_bApplyGravity && m_ctCollided == 0; ext[3i+1] += -10000   (:331-338)
  if (c->gravity && c->contact_count == 0)
    for (size_t i = 1; i < r; i += 3) f[i] += -10000.0;
  // applyHapticForces                                                                    (:634-706)
  if (c->nHaptic > 0 && c->haptic_in_progress) {
    for (int i = 0; i < c->nHaptic; i++)
      for (int d = 0; d < 3; d++) f[3 * (size_t)c->haptic_idx_host[i] + d] += c->haptic_f_host[3 * (size_t)i + d];
    const int R = c->haptic_rings;
    if (R > 1) {
      if (!(c->haptic_quirk && c->edges_host) && !c->adj_host_bp) {
        std::vector<int> bp, bc;
        FB_TRY(fb_fetch_structure(c, bp, bc));
        c->adj_host_bp = (int *)malloc(sizeof(int) * bp.size());
        c->adj_host_bc = (int *)malloc(sizeof(int) * (bc.size() ? bc.size() : 1));
        memcpy(c->adj_host_bp, bp.data(), sizeof(int) * bp.size());
        memcpy(c->adj_host_bc, bc.data(), sizeof(int) * bc.size());
      }
      std::vector<int> nbors;
      for (int iv = 0; iv < c->nHaptic; iv++) {
        std::set<int> affected, last;
        affected.insert(c->haptic_idx_host[iv]);
        last.insert(c->haptic_idx_host[iv]);
        const double *ef = c->haptic_f_host + 3 * (size_t)iv;
        for (int j = 1; j < R; j++) {
          const double mag = 1.0 * (R - j) / static_cast<double>(R);  // linear kernel (:657-658)
          std::set<int> fresh;
          for (std::set<int>::iterator itv = last.begin(); itv != last.end(); ++itv) {
            node_neighbors(c, *itv, nbors);
            for (size_t k = 0; k < nbors.size(); k++)
              if (affected.find(nbors[k]) == affected.end()) fresh.insert(nbors[k]);
          }
          last.clear();
          for (std::set<int>::iterator itn = fresh.begin(); itn != fresh.end(); ++itn) {
            f[3 * (size_t)*itn] += mag * ef[0];
            f[3 * (size_t)*itn + 1] += mag * ef[1];
            f[3 * (size_t)*itn + 2] += mag * ef[2];
            last.insert(*itn);
            affected.insert(*itn);
          }
        }
      }
    }
  }
  if (r) FB_CUDA(cudaMemcpyAsync(c->fext, f, sizeof(double) * r, cudaMemcpyHostToDevice, c->stream));
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
