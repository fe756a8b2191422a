// fb_dist.cu — one big mesh over several GPUs (BASELINE.json config 5): row-block partition, NCCL halo exchange
// of the PCG search direction and 1-element FP64 all-reduces for the two dot products.
//
// No reference counterpart (the reference is single-threaded CPU code).  Design:
//  * Rows (vertices) are split into `world` contiguous ranges holding equal numbers of tet incidences (a proxy for
//    matrix nonzeros) — row blocks, METIS-style objective without METIS (not in the image); for the benchmark
//    cube the ranges are slabs along the slowest index with two neighbours each.
//  * Each rank keeps the tets that touch one of its rows and all their vertices (owned + ghost), renumbered in
//    ascending global order, and runs the ordinary single-GPU setup on that local mesh.  Cut tets are therefore
//    assembled on both sides: assembly needs NO communication, and since contributions are still summed in
//    ascending (global) element order the owned rows of K are bit-identical to the single-GPU matrix.
//  * Ghost rows are masked in the solver (rowmask); ghost COLUMNS carry the neighbour's values of d, refreshed by
//    a halo exchange after every direction update (pack kernel -> grouped ncclSend/ncclRecv -> unpack kernel).
//    x, q, qvel at ghosts follow by the same arithmetic on the same bits, so the state needs no exchange.
//  * d.q and sum r^2/diag are reduced per rank in fixed order, then summed across ranks by ncclAllReduce on the
//    device scalar; every rank sees the same bits, so the loop flag is consistent without a host round trip.
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fb_internal.h"
#include "fb_pcg_common.cuh"


struct FbDist {
  ncclComm_t ncomm;
  ncclComm_t comm_nccl() const { return ncomm; }
  int rank, world;
  int nV_global, nT_global;
  int vbeg, vend;            // owned vertex range, in the ordering the partition was cut from
  int reordered;             // 1: that ordering is Cuthill-McKee, not the caller's numbering
  std::vector<int> l2g;      // local vertex -> global vertex (ascending)
  std::vector<unsigned char> owned;  // per local vertex
  int nNbr;
  std::vector<int> nbrRank, sendOff, recvOff;  // offsets in vertices, size nNbr + 1
  int *sendIdx, *recvIdx;    // device: local vertex ids, concatenated per neighbour
  double *sendBuf, *recvBuf; // device: 3 doubles per vertex
  double *hostStage;         // pinned, local r doubles
  // peer-memory exchange (CUDA IPC): see fb_pcg_common.cuh
  int p2p;
  double *comm;                      // this rank's comm block (device)
  double *peerComm[FB_MAX_RANKS];    // every rank's comm block mapped here ([rank] = comm)
  double *peerDir[FB_MAX_NBR];       // neighbours' search-direction vectors mapped here
  int *remoteIdx;                    // device: neighbour-local vertex index of every send entry
  unsigned int *pushTicket;          // device
  unsigned char *pushFlag;           // device [nV]: vertex is in some send list
  int *pushPtr;                      // device [nV+1]
  int2 *pushEnt;                     // device: (neighbour slot, neighbour-local vertex) per send entry, by vertex
  unsigned long long solveCount;
  void *opened[FB_MAX_RANKS + FB_MAX_NBR];
  int nOpened;
};

#define FB_NCCL(call)                                                                        \
  do {                                                                                       \
    ncclResult_t r__ = (call);                                                               \
    if (r__ != ncclSuccess) {                                                                \
      fb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__));   \
      return FB_ERR_COMM;                                                                    \
    }                                                                                        \
  } while (0)

namespace {

__global__ void k_pack(int n, const int *__restrict__ idx, const double *__restrict__ vec, double *__restrict__ buf) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  int i = t / 3, k = t - 3 * i;
  buf[t] = vec[3 * (size_t)idx[i] + k];
}
__global__ void k_unpack(int n, const int *__restrict__ idx, const double *__restrict__ buf, double *__restrict__ vec) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * n) return;
  int i = t / 3, k = t - 3 * i;
  vec[3 * (size_t)idx[i] + k] = buf[t];
}
__global__ void k_mask_ghost(int nV, const unsigned char *__restrict__ ownedV, const unsigned char *__restrict__ fixed,
                             unsigned char *__restrict__ rowmask) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 3 * nV) return;
  rowmask[t] = (fixed[t] || !ownedV[t / 3]) ? 1 : 0;
}

struct HaloPushArgs {
  int nNbr;
  int off[FB_MAX_NBR + 1];
  int nbrRank[FB_MAX_NBR];
  double *peerVec[FB_MAX_NBR];
};

// Owned boundary values of `vec` stored straight into the neighbours' ghost entries over NVLink, then (last CTA) the
// halo flag of this rank raised in each neighbour's comm block.
__global__ void k_halo_push(HaloPushArgs h, FbPeerArgs pa, const int *__restrict__ sendIdx, const int *__restrict__ remoteIdx,
                            const double *__restrict__ vec, FbScalars *sc, unsigned int *ticket) {
  if (sc->done) return;
  const int total = 3 * h.off[h.nNbr];
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const int i = t / 3, k = t - 3 * i;
    int j = 0;
    while (j + 1 < h.nNbr && i >= h.off[j + 1]) j++;
    h.peerVec[j][3 * (size_t)remoteIdx[i] + k] = vec[3 * (size_t)sendIdx[i] + k];
  }
  __threadfence_system();
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  __syncthreads();
  if (last && threadIdx.x == 0) {
    *ticket = 0u;
    __threadfence_system();
    for (int j = 0; j < h.nNbr; j++)
      ((volatile unsigned long long *)pa.comm[h.nbrRank[j]])[FB_COMM_FLAG(pa.parity, FB_COMM_HALO, pa.rank)] = pa.epoch;
  }
}

// ---- the partition plan: pure host integer work, identical on every rank ------------------------------------
struct Plan {
  std::vector<int> bounds;       // world + 1 vertex boundaries
  std::vector<int> localTets;    // global tet ids, ascending
  std::vector<int> l2g;          // ascending global vertex ids of the local mesh
  std::vector<int> nbr;          // neighbour ranks, ascending
  std::vector<std::vector<int> > send, recv;  // global vertex ids per neighbour, ascending
};

int owner_of(const std::vector<int> &bounds, int v) {
  return (int)(std::upper_bound(bounds.begin(), bounds.end(), v) - bounds.begin()) - 1;
}

void make_bounds(int nV, int nT, const int *tets, int world, std::vector<int> &bounds) {
  std::vector<long long> w((size_t)nV + 1, 0);
  for (size_t i = 0; i < 4 * (size_t)nT; i++) w[(size_t)tets[i] + 1]++;
  for (int v = 0; v < nV; v++) w[(size_t)v + 1] += w[v];
  const long long total = w[nV];
  bounds.assign((size_t)world + 1, nV);
  bounds[0] = 0;
  for (int p = 1; p < world; p++) {
    const long long target = total * p / world;
    int v = (int)(std::lower_bound(w.begin(), w.end(), target) - w.begin());
    if (v > nV) v = nV;
    if (v < bounds[p - 1]) v = bounds[p - 1];
    bounds[p] = v;
  }
  bounds[world] = nV;
}

// number of tets whose vertices do not all belong to one rank
long long count_cut_tets(const std::vector<int> &bounds, int nT, const int *tets) {
  long long cut = 0;
  for (int el = 0; el < nT; el++) {
    const int *t = tets + 4 * (size_t)el;
    const int o = owner_of(bounds, t[0]);
    cut += (owner_of(bounds, t[1]) != o || owner_of(bounds, t[2]) != o || owner_of(bounds, t[3]) != o) ? 1 : 0;
  }
  return cut;
}

// Cuthill-McKee ordering of the vertex graph (vertices sharing a tet), per connected component from a pseudo-peripheral
// root (two breadth-first sweeps), neighbours visited by increasing degree.  order[new] = old.  Contiguous ranges of this
// ordering are breadth-first "slabs": the row-block partition's cut becomes a level-set surface instead of whatever the
// mesh generator's numbering implies.  Pure host integer work, identical on every rank.
void cuthill_mckee(int nV, int nT, const int *tets, std::vector<int> &order) {
  std::vector<long long> ptr((size_t)nV + 1, 0);
  for (size_t i = 0; i < 4 * (size_t)nT; i++) ptr[(size_t)tets[i] + 1] += 3;
  for (int v = 0; v < nV; v++) ptr[(size_t)v + 1] += ptr[v];
  std::vector<int> adj((size_t)ptr[nV]);
  {
    std::vector<long long> fill(ptr.begin(), ptr.end() - 1);
    for (int el = 0; el < nT; el++) {
      const int *t = tets + 4 * (size_t)el;
      for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++)
          if (i != j) adj[(size_t)fill[t[i]]++] = t[j];
    }
  }
  // unique neighbours per vertex, then sorted by (degree, id)
  std::vector<int> deg((size_t)nV, 0);
  for (int v = 0; v < nV; v++) {
    int *b = adj.data() + ptr[v], *e = adj.data() + ptr[v + 1];
    std::sort(b, e);
    deg[v] = (int)(std::unique(b, e) - b);
  }
  for (int v = 0; v < nV; v++) {
    int *b = adj.data() + ptr[v];
    std::sort(b, b + deg[v], [&](int x, int y) { return deg[x] != deg[y] ? deg[x] < deg[y] : x < y; });
  }
  order.clear();
  order.reserve((size_t)nV);
  std::vector<int> mark((size_t)nV, 0);  // 0 = unvisited; sweeps use increasing stamps
  std::vector<int> queue;
  queue.reserve((size_t)nV);
  int stamp = 0;
  auto sweep = [&](int root) {  // breadth-first from root over unfinished vertices; returns the last vertex reached
    stamp++;
    queue.clear();
    queue.push_back(root);
    mark[root] = stamp;
    for (size_t h = 0; h < queue.size(); h++) {
      const int v = queue[h];
      const int *b = adj.data() + ptr[v];
      for (int k = 0; k < deg[v]; k++)
        if (mark[b[k]] >= 0 && mark[b[k]] != stamp) { mark[b[k]] = stamp; queue.push_back(b[k]); }
    }
    return queue.back();
  };
  for (int seed = 0; seed < nV; seed++) {
    if (mark[seed] < 0) continue;           // already ordered
    const int far1 = sweep(seed);           // pseudo-peripheral root: the far end of the far end
    const int far2 = sweep(far1);
    sweep(far2);
    for (size_t h = 0; h < queue.size(); h++) { order.push_back(queue[h]); mark[queue[h]] = -1; }
  }
}

// The numbering the row blocks are cut from: the caller's, unless its cut is large AND Cuthill-McKee's is clearly smaller
// (structured inputs like the cube keep their numbering and their slabs).  order empty = identity.
void choose_ordering(int nV, int nT, const int *tets, int world, std::vector<int> &order) {
  order.clear();
  if (world < 2 || nT == 0) return;
  std::vector<int> bounds;
  make_bounds(nV, nT, tets, world, bounds);
  const long long cutI = count_cut_tets(bounds, nT, tets);
  if (cutI * 100 <= 15ll * nT) return;  // <= 15 % of the tets are cut (a banded numbering: slabs): keep it, skip the graph work
  std::vector<int> cm;
  cuthill_mckee(nV, nT, tets, cm);
  if ((int)cm.size() != nV) return;
  std::vector<int> inv((size_t)nV);
  for (int i = 0; i < nV; i++) inv[cm[i]] = i;
  std::vector<int> pt(4 * (size_t)nT);
  for (size_t i = 0; i < pt.size(); i++) pt[i] = inv[tets[i]];
  make_bounds(nV, nT, pt.data(), world, bounds);
  const long long cutR = count_cut_tets(bounds, nT, pt.data());
  if (cutR * 10 < cutI * 9) order.swap(cm);
}

void make_plan(int nV, int nT, const int *tets, int world, int rank, Plan &pl) {
  make_bounds(nV, nT, tets, world, pl.bounds);
  const int vb = pl.bounds[rank], ve = pl.bounds[rank + 1];
  std::vector<unsigned char> mark((size_t)nV, 0);
  pl.localTets.clear();
  for (int el = 0; el < nT; el++) {
    const int *t = tets + 4 * (size_t)el;
    bool mine = false;
    for (int i = 0; i < 4; i++) mine |= (t[i] >= vb && t[i] < ve);
    if (!mine) continue;
    pl.localTets.push_back(el);
    for (int i = 0; i < 4; i++) mark[t[i]] = 1;
  }
  pl.l2g.clear();
  for (int v = 0; v < nV; v++)
    if (mark[v]) pl.l2g.push_back(v);
  // halo lists from the local tets: a (mine) next to b (rank p) => a goes to p, b comes from p
  std::vector<std::vector<int> > send((size_t)world), recv((size_t)world);
  for (size_t k = 0; k < pl.localTets.size(); k++) {
    const int *t = tets + 4 * (size_t)pl.localTets[k];
    int own[4];
    for (int i = 0; i < 4; i++) own[i] = owner_of(pl.bounds, t[i]);
    for (int i = 0; i < 4; i++) {
      if (own[i] != rank) continue;
      for (int j = 0; j < 4; j++)
        if (own[j] != rank) { send[own[j]].push_back(t[i]); recv[own[j]].push_back(t[j]); }
    }
  }
  pl.nbr.clear(); pl.send.clear(); pl.recv.clear();
  for (int p = 0; p < world; p++) {
    if (send[p].empty() && recv[p].empty()) continue;
    std::sort(send[p].begin(), send[p].end());
    send[p].erase(std::unique(send[p].begin(), send[p].end()), send[p].end());
    std::sort(recv[p].begin(), recv[p].end());
    recv[p].erase(std::unique(recv[p].begin(), recv[p].end()), recv[p].end());
    pl.nbr.push_back(p);
    pl.send.push_back(send[p]);
    pl.recv.push_back(recv[p]);
  }
}

}  // namespace

// ---- hooks used by the solver ---------------------------------------------------------------------------------
int fb_dist_halo_exchange(fb_context *c, double *vec) {
  FbDist *d = c->dist;
  if (!d || d->nNbr == 0) return FB_OK;
  const int nS = d->sendOff[d->nNbr], nR = d->recvOff[d->nNbr];
  if (nS) { k_pack<<<(3 * nS + 255) / 256, 256, 0, c->stream>>>(nS, d->sendIdx, vec, d->sendBuf); c->launches++; }
  FB_NCCL(ncclGroupStart());
  for (int i = 0; i < d->nNbr; i++) {
    const int ns = d->sendOff[i + 1] - d->sendOff[i], nr = d->recvOff[i + 1] - d->recvOff[i];
    if (ns) FB_NCCL(ncclSend(d->sendBuf + 3 * (size_t)d->sendOff[i], 3 * (size_t)ns, ncclDouble, d->nbrRank[i], d->ncomm, c->stream));
    if (nr) FB_NCCL(ncclRecv(d->recvBuf + 3 * (size_t)d->recvOff[i], 3 * (size_t)nr, ncclDouble, d->nbrRank[i], d->ncomm, c->stream));
  }
  FB_NCCL(ncclGroupEnd());
  if (nR) { k_unpack<<<(3 * nR + 255) / 256, 256, 0, c->stream>>>(nR, d->recvIdx, d->recvBuf, vec); c->launches++; }
  return FB_OK;
}

int fb_dist_allreduce_scalar(fb_context *c, const double *dev_part, double *dev_total) {
  FbDist *d = c->dist;
  if (!d) return FB_OK;
  if (d->world == 1) {
    FB_CUDA(cudaMemcpyAsync(dev_total, dev_part, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    return FB_OK;
  }
  FB_NCCL(ncclAllReduce(dev_part, dev_total, 1, ncclDouble, ncclSum, d->ncomm, c->stream));
  return FB_OK;
}

int fb_dist_p2p(const fb_context *c) { return (c && c->dist) ? c->dist->p2p : 0; }

void fb_dist_peer_args(fb_context *c, FbPeerArgs *pa) {
  memset(pa, 0, sizeof(*pa));
  FbDist *d = c->dist;
  if (!d || !d->p2p) return;
  pa->enabled = 1;
  pa->rank = d->rank;
  pa->world = d->world;
  pa->parity = (int)(d->solveCount & 1ull);
  for (int p = 0; p < d->world; p++) pa->comm[p] = d->peerComm[p];
}

unsigned long long fb_dist_epoch(fb_context *c, int it, int family) {
  return (c->dist->solveCount << 32) | (unsigned long long)(3 * (long long)it + family + 1);
}

void fb_dist_next_solve(fb_context *c) {
  if (c->dist) c->dist->solveCount++;
}

int fb_dist_reset_tickets(fb_context *c) {
  if (c->dist && c->dist->pushTicket) FB_CUDA(cudaMemsetAsync(c->dist->pushTicket, 0, sizeof(unsigned int), c->stream));
  return FB_OK;
}

unsigned int fb_dist_halo_mask(const fb_context *c) {
  unsigned int m = 0;
  for (int i = 0; i < c->dist->nNbr; i++) m |= 1u << c->dist->nbrRank[i];
  return m;
}

int fb_dist_halo_push(fb_context *c, const double *vec, unsigned long long epoch) {
  FbDist *d = c->dist;
  if (!d || !d->p2p || d->nNbr == 0) return FB_OK;
  HaloPushArgs h;
  memset(&h, 0, sizeof(h));
  h.nNbr = d->nNbr;
  for (int i = 0; i <= d->nNbr; i++) h.off[i] = d->sendOff[i];
  for (int i = 0; i < d->nNbr; i++) { h.nbrRank[i] = d->nbrRank[i]; h.peerVec[i] = d->peerDir[i]; }
  FbPeerArgs pa;
  fb_dist_peer_args(c, &pa);
  pa.epoch = epoch;
  const int total = 3 * d->sendOff[d->nNbr];
  int grid = (total + 255) / 256;
  if (grid > 2 * c->sm_count) grid = 2 * c->sm_count;
  if (grid < 1) grid = 1;
  k_halo_push<<<grid, 256, 0, c->stream>>>(h, pa, d->sendIdx, d->remoteIdx, vec, c->sc, d->pushTicket);
  c->launches++;
  return FB_OK;
}

void fb_dist_push_args(fb_context *c, FbPushArgs *out, unsigned long long epoch) {
  memset(out, 0, sizeof(*out));
  FbDist *d = c->dist;
  if (!d || !d->p2p || d->nNbr == 0) return;
  out->nNbr = d->nNbr;
  for (int i = 0; i < d->nNbr; i++) { out->nbrRank[i] = d->nbrRank[i]; out->peerVec[i] = d->peerDir[i]; }
  out->pushFlag = d->pushFlag; out->pushPtr = d->pushPtr; out->pushEnt = d->pushEnt;
  out->epoch = epoch;
}

// Peer mappings for the exchange above: comm blocks of all ranks, `dir` vectors of the neighbours, and for every send
// entry the neighbour's local index of that vertex (the neighbour's recv list, which mirrors this rank's send list).
// Any failure leaves p2p = 0 and the context on the NCCL path.
static int setup_p2p(fb_context *c) {
  FbDist *d = c->dist;
  d->p2p = 0;
  const char *env = getenv("FEMBRAIN_B200_P2P");
  if (env && atoi(env) == 0) return FB_OK;
  if (d->world < 2) return FB_OK;
  cudaStream_t st = c->stream;
  // Every buffer the collectives below need is allocated BEFORE the first of them; after this point a rank-local failure
  // (allocation, IPC export, mapping) only clears `ok` and every rank still runs every collective, so no rank can be left
  // blocked inside NCCL by a peer that returned early.  The decision is taken once, from the final agreed vote.
  struct Handles { cudaIpcMemHandle_t comm, dir; };
  const int nS = d->sendOff[d->nNbr];
  int *vote = nullptr, *remoteScratch = nullptr;
  char *devH = nullptr;
  if (cudaMalloc(&vote, 2 * sizeof(int)) != cudaSuccess || cudaMalloc(&devH, sizeof(Handles) * (size_t)(d->world + 1)) != cudaSuccess ||
      cudaMalloc(&remoteScratch, sizeof(int) * (size_t)(nS + 1)) != cudaSuccess) {
    cudaGetLastError();
    fb_dev_free(vote); fb_dev_free(devH); fb_dev_free(remoteScratch);
    fb_set_error("peer-memory setup: scratch allocation failed");
    return FB_ERR_OUT_OF_MEMORY;
  }
  auto release = [&]() { fb_dev_free(vote); fb_dev_free(devH); fb_dev_free(remoteScratch); };
  {  // eligibility vote: all ranks take part, all ranks see the same answer
    int eligible = (d->world <= FB_MAX_RANKS && d->nNbr <= FB_MAX_NBR && c->use_rows3) ? 1 : 0, all = 0;
    cudaMemcpyAsync(vote, &eligible, sizeof(int), cudaMemcpyHostToDevice, st);
    ncclResult_t r0 = ncclAllReduce(vote, vote, 1, ncclInt, ncclMin, d->comm_nccl(), st);
    cudaMemcpyAsync(&all, vote, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    if (r0 != ncclSuccess || !all) { cudaGetLastError(); release(); return FB_OK; }
  }
  int ok = 1;
  if (fb_dev_alloc_plain(c, &d->comm, (size_t)FB_COMM_WORDS) != FB_OK) ok = 0;  // exported with CUDA IPC
  if (fb_dev_alloc(c, &d->pushTicket, 1) != FB_OK) ok = 0;
  if (fb_dev_alloc(c, &d->remoteIdx, (size_t)nS) != FB_OK) ok = 0;
  if (ok && (cudaMemsetAsync(d->comm, 0, sizeof(double) * FB_COMM_WORDS, st) != cudaSuccess ||
             cudaMemsetAsync(d->pushTicket, 0, sizeof(unsigned int), st) != cudaSuccess)) ok = 0;
  // 1. IPC handles of (comm, dir) of every rank
  Handles mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) ok = (cudaIpcGetMemHandle(&mine.comm, d->comm) == cudaSuccess) && (cudaIpcGetMemHandle(&mine.dir, c->dir) == cudaSuccess);
  if (cudaMemcpyAsync(devH + sizeof(Handles) * (size_t)d->world, &mine, sizeof(Handles), cudaMemcpyHostToDevice, st) != cudaSuccess) ok = 0;
  ncclResult_t nr = ncclAllGather(devH + sizeof(Handles) * (size_t)d->world, devH, sizeof(Handles), ncclChar, d->comm_nccl(), st);
  std::vector<Handles> all((size_t)d->world);
  if (nr == ncclSuccess) cudaMemcpyAsync(all.data(), devH, sizeof(Handles) * (size_t)d->world, cudaMemcpyDeviceToHost, st);
  // 2. the neighbours' local indices of my send vertices = their recv lists for me
  int *remoteDst = d->remoteIdx ? d->remoteIdx : remoteScratch;
  if (nr == ncclSuccess) nr = ncclGroupStart();
  for (int i = 0; i < d->nNbr && nr == ncclSuccess; i++) {
    const int ns = d->sendOff[i + 1] - d->sendOff[i], nrv = d->recvOff[i + 1] - d->recvOff[i];
    if (nrv) nr = ncclSend(d->recvIdx + d->recvOff[i], (size_t)nrv, ncclInt, d->nbrRank[i], d->comm_nccl(), st);
    if (ns && nr == ncclSuccess) nr = ncclRecv(remoteDst + d->sendOff[i], (size_t)ns, ncclInt, d->nbrRank[i], d->comm_nccl(), st);
  }
  if (nr == ncclSuccess) nr = ncclGroupEnd();
  cudaError_t ce = cudaStreamSynchronize(st);
  if (nr != ncclSuccess || ce != cudaSuccess) { cudaGetLastError(); ok = 0; }
  // 3. map the peers
  d->nOpened = 0;
  for (int p = 0; p < d->world && ok; p++) {
    if (p == d->rank) { d->peerComm[p] = d->comm; continue; }
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[p].comm, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    d->peerComm[p] = (double *)ptr;
    d->opened[d->nOpened++] = ptr;
  }
  for (int i = 0; i < d->nNbr && ok; i++) {
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[d->nbrRank[i]].dir, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
    d->peerDir[i] = (double *)ptr;
    d->opened[d->nOpened++] = ptr;
  }
  // 4. all ranks must agree (a rank that failed would otherwise wait for peers that never publish)
  int agreed = 0;
  if (cudaMemcpyAsync(vote, &ok, sizeof(int), cudaMemcpyHostToDevice, st) != cudaSuccess) ok = 0;
  if (ncclAllReduce(vote, vote, 1, ncclInt, ncclMin, d->comm_nccl(), st) != ncclSuccess) ok = 0;
  cudaMemcpyAsync(&agreed, vote, sizeof(int), cudaMemcpyDeviceToHost, st);
  cudaStreamSynchronize(st);
  cudaGetLastError();
  release();
  d->p2p = ok && agreed;
  if (d->p2p) {  // send entries regrouped by local vertex for the push fused into k_direction (no collectives from here on)
    std::vector<int> sIdx((size_t)nS), rIdx((size_t)nS);
    if (nS) {
      FB_CUDA(cudaMemcpyAsync(sIdx.data(), d->sendIdx, sizeof(int) * (size_t)nS, cudaMemcpyDeviceToHost, st));
      FB_CUDA(cudaMemcpyAsync(rIdx.data(), d->remoteIdx, sizeof(int) * (size_t)nS, cudaMemcpyDeviceToHost, st));
      FB_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<int> ptr((size_t)c->nV + 1, 0);
    std::vector<unsigned char> flag((size_t)c->nV, 0);
    for (int k = 0; k < nS; k++) { ptr[(size_t)sIdx[k] + 1]++; flag[sIdx[k]] = 1; }
    for (int v = 0; v < c->nV; v++) ptr[(size_t)v + 1] += ptr[v];
    std::vector<int2> ent((size_t)(nS ? nS : 1));
    std::vector<int> fill(ptr.begin(), ptr.end() - 1);
    for (int j = 0; j < d->nNbr; j++)
      for (int k = d->sendOff[j]; k < d->sendOff[j + 1]; k++) ent[(size_t)fill[sIdx[k]]++] = make_int2(j, rIdx[k]);
    FB_TRY(fb_dev_alloc(c, &d->pushFlag, (size_t)c->nV));
    FB_TRY(fb_dev_alloc(c, &d->pushPtr, (size_t)c->nV + 1));
    FB_TRY(fb_dev_alloc(c, &d->pushEnt, (size_t)nS));
    FB_CUDA(cudaMemcpyAsync(d->pushFlag, flag.data(), flag.size(), cudaMemcpyHostToDevice, st));
    FB_CUDA(cudaMemcpyAsync(d->pushPtr, ptr.data(), sizeof(int) * ptr.size(), cudaMemcpyHostToDevice, st));
    if (nS) FB_CUDA(cudaMemcpyAsync(d->pushEnt, ent.data(), sizeof(int2) * (size_t)nS, cudaMemcpyHostToDevice, st));
    FB_CUDA(cudaStreamSynchronize(st));
  }
  return FB_OK;
}

int fb_dist_refresh_rowmask(fb_context *c) {
  FbDist *d = c->dist;
  if (!d) return FB_OK;
  unsigned char *ownedDev = nullptr;
  FB_CUDA(cudaMalloc(&ownedDev, (size_t)(c->nV ? c->nV : 1)));
  FB_CUDA(cudaMemcpyAsync(ownedDev, d->owned.data(), (size_t)c->nV, cudaMemcpyHostToDevice, c->stream));
  if (c->nV) { k_mask_ghost<<<(3 * c->nV + 255) / 256, 256, 0, c->stream>>>(c->nV, ownedDev, c->fixed, c->rowmask); c->launches++; }
  FB_CUDA(cudaStreamSynchronize(c->stream));
  fb_dev_free(ownedDev);
  return FB_OK;
}

int fb_dist_global_sizes(const fb_context *c, int *nV, int *nT) {
  if (!c || !c->dist) return FB_ERR_NOT_SUPPORTED;
  *nV = c->dist->nV_global;
  *nT = c->dist->nT_global;
  return FB_OK;
}

int fb_dist_upload_global(fb_context *c, const double *g, double *localDev) {
  FbDist *d = c->dist;
  double *h = d->hostStage;
  for (int v = 0; v < c->nV; v++) {
    const size_t gv = 3 * (size_t)d->l2g[v];
    h[3 * (size_t)v] = g[gv]; h[3 * (size_t)v + 1] = g[gv + 1]; h[3 * (size_t)v + 2] = g[gv + 2];
  }
  if (c->r) FB_CUDA(cudaMemcpyAsync(localDev, h, sizeof(double) * (size_t)c->r, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

int fb_dist_download_owned(fb_context *c, const double *localDev, double *g) {
  FbDist *d = c->dist;
  double *h = d->hostStage;
  if (c->r) FB_CUDA(cudaMemcpyAsync(h, localDev, sizeof(double) * (size_t)c->r, cudaMemcpyDeviceToHost, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  memset(g, 0, sizeof(double) * 3 * (size_t)d->nV_global);
  for (int v = 0; v < c->nV; v++) {
    if (!d->owned[v]) continue;
    const size_t gv = 3 * (size_t)d->l2g[v];
    g[gv] = h[3 * (size_t)v]; g[gv + 1] = h[3 * (size_t)v + 1]; g[gv + 2] = h[3 * (size_t)v + 2];
  }
  return FB_OK;
}

void fb_dist_destroy(fb_context *c) {
  FbDist *d = c->dist;
  if (!d) return;
  if (d->sendIdx) fb_dev_free(d->sendIdx);
  if (d->recvIdx) fb_dev_free(d->recvIdx);
  if (d->sendBuf) fb_dev_free(d->sendBuf);
  if (d->recvBuf) fb_dev_free(d->recvBuf);
  if (d->hostStage) cudaFreeHost(d->hostStage);
  if (c->rowmask && c->rowmask != c->fixed) { fb_dev_free(c->rowmask); c->rowmask = nullptr; }
  for (int i = 0; i < d->nOpened; i++) cudaIpcCloseMemHandle(d->opened[i]);
  if (d->comm) fb_dev_free(d->comm);
  if (d->pushTicket) fb_dev_free(d->pushTicket);
  if (d->remoteIdx) fb_dev_free(d->remoteIdx);
  if (d->pushFlag) fb_dev_free(d->pushFlag);
  if (d->pushPtr) fb_dev_free(d->pushPtr);
  if (d->pushEnt) fb_dev_free(d->pushEnt);
  if (d->ncomm) ncclCommDestroy(d->ncomm);
  delete d;
  c->dist = nullptr;
}

// ===================================================================================================================
extern "C" {

int fb_comm_unique_id(void *id128) {
  if (!id128) return FB_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  FB_NCCL(ncclGetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return FB_OK;
}

// Host-only inspection of the partition (no GPU needed): what rank `rank` of `world` would own and exchange.
// counts: [0] vertex_begin, [1] vertex_end, [2] local vertices, [3] local tets, [4] neighbours, [5] total send
// vertices, [6] total recv vertices.  Optional outputs (NULL to skip) are sized from a first call:
// l2g[counts[2]], local_tets[counts[3]], nbr_ranks[counts[4]], send_counts/recv_counts[counts[4]],
// send_global[counts[5]], recv_global[counts[6]] (concatenated per neighbour, ascending global ids).
int fb_plan_partition(int nV, int nT, const int *tets, int world, int rank, int *counts, int *l2g, int *local_tets,
                      int *nbr_ranks, int *send_counts, int *recv_counts, int *send_global, int *recv_global) {
  if (nV < 0 || nT < 0 || (nT > 0 && !tets) || world < 1 || rank < 0 || rank >= world || !counts) {
    fb_set_error("bad arguments to fb_plan_partition");
    return FB_ERR_INVALID_ARGUMENT;
  }
  for (size_t i = 0; i < 4 * (size_t)nT; i++)
    if (tets[i] < 0 || tets[i] >= nV) { fb_set_error("tetrahedron %zu references a vertex outside [0, %d)", i / 4, nV); return FB_ERR_BAD_MESH; }
  Plan pl;
  std::vector<int> order, ptets;
  choose_ordering(nV, nT, tets, world, order);
  if (!order.empty()) {  // plan on the renumbered mesh, report in the caller's vertex ids
    std::vector<int> inv((size_t)nV);
    for (int i = 0; i < nV; i++) inv[order[i]] = i;
    ptets.resize(4 * (size_t)nT);
    for (size_t i = 0; i < ptets.size(); i++) ptets[i] = inv[tets[i]];
    make_plan(nV, nT, ptets.data(), world, rank, pl);
    for (size_t i = 0; i < pl.l2g.size(); i++) pl.l2g[i] = order[pl.l2g[i]];
    for (size_t i = 0; i < pl.nbr.size(); i++) {
      for (size_t k = 0; k < pl.send[i].size(); k++) pl.send[i][k] = order[pl.send[i][k]];
      for (size_t k = 0; k < pl.recv[i].size(); k++) pl.recv[i][k] = order[pl.recv[i][k]];
    }
  } else {
    make_plan(nV, nT, tets, world, rank, pl);
  }
  size_t ns = 0, nr = 0;
  for (size_t i = 0; i < pl.nbr.size(); i++) { ns += pl.send[i].size(); nr += pl.recv[i].size(); }
  counts[0] = pl.bounds[rank]; counts[1] = pl.bounds[rank + 1];
  counts[2] = (int)pl.l2g.size(); counts[3] = (int)pl.localTets.size(); counts[4] = (int)pl.nbr.size();
  counts[5] = (int)ns; counts[6] = (int)nr;
  if (l2g) memcpy(l2g, pl.l2g.data(), sizeof(int) * pl.l2g.size());
  if (local_tets) memcpy(local_tets, pl.localTets.data(), sizeof(int) * pl.localTets.size());
  size_t so = 0, ro = 0;
  for (size_t i = 0; i < pl.nbr.size(); i++) {
    if (nbr_ranks) nbr_ranks[i] = pl.nbr[i];
    if (send_counts) send_counts[i] = (int)pl.send[i].size();
    if (recv_counts) recv_counts[i] = (int)pl.recv[i].size();
    if (send_global) memcpy(send_global + so, pl.send[i].data(), sizeof(int) * pl.send[i].size());
    if (recv_global) memcpy(recv_global + ro, pl.recv[i].data(), sizeof(int) * pl.recv[i].size());
    so += pl.send[i].size();
    ro += pl.recv[i].size();
  }
  return FB_OK;
}

int fb_partition_ordering(int nV, int nT, const int *tets, int world, int *order, int *reordered) {
  if (nV < 0 || nT < 0 || (nT > 0 && !tets) || world < 1 || !reordered) { fb_set_error("bad arguments to fb_partition_ordering"); return FB_ERR_INVALID_ARGUMENT; }
  for (size_t i = 0; i < 4 * (size_t)nT; i++)
    if (tets[i] < 0 || tets[i] >= nV) { fb_set_error("tetrahedron %zu references a vertex outside [0, %d)", i / 4, nV); return FB_ERR_BAD_MESH; }
  std::vector<int> ord;
  choose_ordering(nV, nT, tets, world, ord);
  *reordered = ord.empty() ? 0 : 1;
  if (order)
    for (int i = 0; i < nV; i++) order[i] = ord.empty() ? i : ord[i];
  return FB_OK;
}

int fb_create_partitioned(fb_context **out, int nV, const double *x0, int nT, const int *tetsIn, int nFixed, const int *fixedVerts,
                          const fb_params *prm, int rank, int world, const void *comm_id128) {
  const int *tets = tetsIn;
  if (!out) return FB_ERR_INVALID_ARGUMENT;
  *out = nullptr;
  if (nV < 0 || nT < 0 || (nV > 0 && !x0) || (nT > 0 && !tets) || world < 1 || rank < 0 || rank >= world || nFixed < 0 ||
      (nFixed > 0 && !fixedVerts) || (world > 1 && !comm_id128)) {
    fb_set_error("bad arguments to fb_create_partitioned");
    return FB_ERR_INVALID_ARGUMENT;
  }
  for (size_t i = 0; i < 4 * (size_t)nT; i++)
    if (tets[i] < 0 || tets[i] >= nV) { fb_set_error("tetrahedron %zu references a vertex outside [0, %d)", i / 4, nV); return FB_ERR_BAD_MESH; }
  {  // a vertex in no tetrahedron is an error exactly as in fb_create (checked globally here)
    std::vector<unsigned char> used((size_t)nV, 0);
    for (size_t i = 0; i < 4 * (size_t)nT; i++) used[tets[i]] = 1;
    for (int v = 0; v < nV; v++)
      if (!used[v]) { fb_set_error("vertex %d belongs to no tetrahedron", v); return FB_ERR_BAD_MESH; }
  }
  // The partition is cut from the caller's numbering or, for unstructured numberings, from a Cuthill-McKee ordering
  // (choose_ordering).  From here on "global" ids are ids in that ordering; order[] maps them back to the caller's, which
  // is what the vector API (fb_dist_upload_global / _download_owned, through l2g) and the fixed-vertex list speak.
  std::vector<int> order, ptets, inv;
  choose_ordering(nV, nT, tets, world, order);
  if (!order.empty()) {
    inv.resize((size_t)nV);
    for (int i = 0; i < nV; i++) inv[order[i]] = i;
    ptets.resize(4 * (size_t)nT);
    for (size_t i = 0; i < ptets.size(); i++) ptets[i] = inv[tetsIn[i]];
    tets = ptets.data();
  }
  Plan pl;
  make_plan(nV, nT, tets, world, rank, pl);
  const int nLV = (int)pl.l2g.size(), nLT = (int)pl.localTets.size();
  std::vector<int> g2l((size_t)nV, -1);
  for (int i = 0; i < nLV; i++) g2l[pl.l2g[i]] = i;
  std::vector<double> lx(3 * (size_t)nLV);
  for (int i = 0; i < nLV; i++) {
    const int callers = order.empty() ? pl.l2g[i] : order[pl.l2g[i]];
    for (int k = 0; k < 3; k++) lx[3 * (size_t)i + k] = x0[3 * (size_t)callers + k];
  }
  std::vector<int> lt(4 * (size_t)nLT);
  for (int e = 0; e < nLT; e++)
    for (int k = 0; k < 4; k++) lt[4 * (size_t)e + k] = g2l[tets[4 * (size_t)pl.localTets[e] + k]];
  std::vector<int> fv(fixedVerts, fixedVerts + nFixed);
  std::sort(fv.begin(), fv.end());
  std::vector<int> cd, fl;
  for (int i = 0; i < nFixed; i++) {
    if (fv[i] < 0 || fv[i] >= nV) { fb_set_error("fixed vertex %d out of range [0, %d)", fv[i], nV); return FB_ERR_INVALID_ARGUMENT; }
    if (i && fv[i] == fv[i - 1]) { fb_set_error("fixed vertex %d listed twice", fv[i]); return FB_ERR_INVALID_ARGUMENT; }
    const int l = g2l[order.empty() ? fv[i] : inv[fv[i]]];
    if (l >= 0) fl.push_back(l);
  }
  std::sort(fl.begin(), fl.end());  // ascending LOCAL ids (the partition ordering need not follow the caller's numbering)
  for (size_t i = 0; i < fl.size(); i++) { cd.push_back(3 * fl[i]); cd.push_back(3 * fl[i] + 1); cd.push_back(3 * fl[i] + 2); }
  fb_context *c = nullptr;
  FB_TRY(fb_create_local(&c, nLV, lx.data(), nLT, lt.data(), (int)cd.size(), cd.data(), nullptr, nullptr, nullptr, prm));
  FbDist *d = new FbDist();
  d->ncomm = nullptr; d->sendIdx = d->recvIdx = nullptr; d->sendBuf = d->recvBuf = nullptr; d->hostStage = nullptr;
  d->p2p = 0; d->comm = nullptr; d->remoteIdx = nullptr; d->pushTicket = nullptr; d->pushFlag = nullptr; d->pushPtr = nullptr; d->pushEnt = nullptr; d->solveCount = 0; d->nOpened = 0;
  d->rank = rank; d->world = world; d->nV_global = nV; d->nT_global = nT;
  d->vbeg = pl.bounds[rank]; d->vend = pl.bounds[rank + 1];
  d->reordered = order.empty() ? 0 : 1;
  d->owned.resize((size_t)nLV);
  for (int i = 0; i < nLV; i++) d->owned[i] = (pl.l2g[i] >= d->vbeg && pl.l2g[i] < d->vend) ? 1 : 0;
  d->l2g = pl.l2g;  // local vertex -> the CALLER's vertex id
  if (!order.empty())
    for (int i = 0; i < nLV; i++) d->l2g[i] = order[pl.l2g[i]];
  d->nNbr = (int)pl.nbr.size();
  d->nbrRank = pl.nbr;
  d->sendOff.assign((size_t)d->nNbr + 1, 0);
  d->recvOff.assign((size_t)d->nNbr + 1, 0);
  std::vector<int> sIdx, rIdx;
  for (int i = 0; i < d->nNbr; i++) {
    for (size_t k = 0; k < pl.send[i].size(); k++) sIdx.push_back(g2l[pl.send[i][k]]);
    for (size_t k = 0; k < pl.recv[i].size(); k++) rIdx.push_back(g2l[pl.recv[i][k]]);
    d->sendOff[i + 1] = (int)sIdx.size();
    d->recvOff[i + 1] = (int)rIdx.size();
  }
  c->dist = d;
  {  // rows the solver's products visit: the owned rows, contiguous in the local numbering; ghost rows are never multiplied
    int lo = 0, hi = nLV;
    while (lo < nLV && !d->owned[lo]) lo++;
    while (hi > lo && !d->owned[hi - 1]) hi--;
    bool contiguous = true;
    for (int i = lo; i < hi; i++) contiguous &= (d->owned[i] != 0);
    if (!contiguous) { fb_set_error("owned rows are not contiguous in the local numbering"); fb_destroy(c); return FB_ERR_INVALID_ARGUMENT; }
    c->row_lo = lo;
    c->row_hi = hi;
  }
  int st = FB_OK;
#define DCHK(call) do { st = (call); if (st != FB_OK) { fb_destroy(c); return st; } } while (0)
#define DCUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { fb_set_error("%s -> %s", #call, cudaGetErrorString(e__)); fb_destroy(c); return FB_ERR_CUDA; } } while (0)
  DCHK(fb_dev_alloc(c, &d->sendIdx, sIdx.size()));
  DCHK(fb_dev_alloc(c, &d->recvIdx, rIdx.size()));
  DCHK(fb_dev_alloc(c, &d->sendBuf, 3 * sIdx.size()));
  DCHK(fb_dev_alloc(c, &d->recvBuf, 3 * rIdx.size()));
  if (!sIdx.empty()) DCUDA(cudaMemcpyAsync(d->sendIdx, sIdx.data(), sizeof(int) * sIdx.size(), cudaMemcpyHostToDevice, c->stream));
  if (!rIdx.empty()) DCUDA(cudaMemcpyAsync(d->recvIdx, rIdx.data(), sizeof(int) * rIdx.size(), cudaMemcpyHostToDevice, c->stream));
  DCUDA(cudaStreamSynchronize(c->stream));
  DCUDA(cudaMallocHost(&d->hostStage, sizeof(double) * (size_t)(c->r ? c->r : 1)));
  unsigned char *mask = nullptr;
  DCHK(fb_dev_alloc(c, &mask, (size_t)c->r));
  c->rowmask = mask;
  DCHK(fb_dist_refresh_rowmask(c));
  if (world > 1) {
    ncclUniqueId id;
    memcpy(&id, comm_id128, sizeof(id));
    ncclResult_t r = ncclCommInitRank(&d->ncomm, world, id, rank);
    if (r != ncclSuccess) {
      fb_set_error("ncclCommInitRank -> %s", ncclGetErrorString(r));
      d->ncomm = nullptr;
      fb_destroy(c);
      return FB_ERR_COMM;
    }
    DCHK(setup_p2p(c));
  }
#undef DCHK
#undef DCUDA
  *out = c;
  return FB_OK;
}

// ---- owned-range vector API: this rank's rows only, no global-length staging -----------------------------------
// The owned rows are contiguous in the local numbering (row_lo .. row_hi), so these are single copies of
// 3 * (vertex_end - vertex_begin) doubles straight between the caller's buffer and the device vectors.
static int owned_range(fb_context *c, size_t *off, size_t *bytes) {
  if (!c) { fb_set_error("NULL context"); return FB_ERR_INVALID_ARGUMENT; }
  if (cudaSetDevice(c->device) != cudaSuccess) { cudaGetLastError(); return FB_ERR_CUDA; }
  *off = 3 * (size_t)c->row_lo;   // ordinary contexts: row_lo = 0, row_hi = nV (the whole vector)
  *bytes = sizeof(double) * 3 * (size_t)(c->row_hi - c->row_lo);
  return FB_OK;
}

int fb_set_external_forces_owned(fb_context *c, const double *f_owned) {
  size_t off, bytes;
  FB_TRY(owned_range(c, &off, &bytes));
  if (!f_owned) { fb_set_error("f_owned is NULL"); return FB_ERR_INVALID_ARGUMENT; }
  // forces on ghost vertices are not needed: ghost rows are masked in the solve and their state follows from the
  // neighbour's search direction (fb_dist.cu header)
  if (bytes) FB_CUDA(cudaMemcpyAsync(c->fext + off, f_owned, bytes, cudaMemcpyHostToDevice, c->stream));
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

int fb_get_state_owned(fb_context *c, double *q, double *qvel, double *qaccel) {
  size_t off, bytes;
  FB_TRY(owned_range(c, &off, &bytes));
  if (bytes) {
    if (q) FB_CUDA(cudaMemcpyAsync(q, c->q + off, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (qvel) FB_CUDA(cudaMemcpyAsync(qvel, c->qvel + off, bytes, cudaMemcpyDeviceToHost, c->stream));
    if (qaccel) FB_CUDA(cudaMemcpyAsync(qaccel, c->qaccel + off, bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  FB_CUDA(cudaStreamSynchronize(c->stream));
  return FB_OK;
}

int fb_partition_local_range(const fb_context *c, int *local_begin, int *local_end) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  if (local_begin) *local_begin = c->row_lo;
  if (local_end) *local_end = c->row_hi;
  return FB_OK;
}

int fb_partition_local_to_global(const fb_context *c, int *l2g) {
  if (!c || !l2g) return FB_ERR_INVALID_ARGUMENT;
  if (c->dist) memcpy(l2g, c->dist->l2g.data(), sizeof(int) * c->dist->l2g.size());
  else for (int v = 0; v < c->nV; v++) l2g[v] = v;
  return FB_OK;
}

int fb_partition_peer_memory(const fb_context *c) { return fb_dist_p2p(c); }
int fb_partition_reordered(const fb_context *c) { return (c && c->dist) ? c->dist->reordered : 0; }

int fb_partition_range(const fb_context *c, int *b, int *e) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  if (b) *b = c->dist ? c->dist->vbeg : 0;
  if (e) *e = c->dist ? c->dist->vend : c->nV;
  return FB_OK;
}

}  // extern "C"
