// fb_pcg_common.cuh — device helpers shared by the PCG translation units.
#pragma once
#include <cstring>

#include "fb_internal.h"

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may
// have its CTAs scheduled while the previous kernel of the stream is still draining; pdl_wait() blocks until that
// kernel has completed and its writes are visible (a no-op for ordinary launches), pdl_trigger() lets the NEXT
// kernel's CTAs be scheduled as soon as this kernel's CTAs leave their SM slots.  Every PCG kernel calls both at its
// top, before its first read of anything another kernel produced, so only launch latency and ramp-up overlap.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch with the attribute when enabled; same argument conversion rules as <<<>>>
template <typename... KArgs, typename... Args>
static inline void fb_launch(bool pdl, cudaStream_t st, void (*kernel)(KArgs...), int grid, int block, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}


// lanes per block row and scalars covered by the three unrolled passes of k_spmv_rows3
constexpr int TILE_G = 16;
constexpr int TILE_CHUNK = 3 * TILE_G;

struct RowVals {
  double v[3][3];  // [pass][k]
};

// The same streaming load with an L2 eviction hint (createpolicy + .L2::cache_hint): matrix lines marked evict-first leave L2
// before the PCG vectors (d, q, r, x, 1/diag: 21 MB at 1M tets) do, so the vector kernels and the x gathers of the next
// product can find them there.  Opt-in (FEMBRAIN_B200_L2EVICT=1), not measured yet.
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double ld_stream_hint(const double *p, unsigned long long policy) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
  return v;
}
__device__ __forceinline__ void load_row_chunk_hint(const double *__restrict__ A, int rs, int n3, int base, int lane, RowVals &o,
                                                    unsigned long long policy) {
  const double *a0 = A + 9 * (size_t)rs + base;
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int t = base + lane + TILE_G * p;
    const bool ok = t < n3;
    const double *q = a0 + lane + TILE_G * p;
    o.v[p][0] = ok ? ld_stream_hint(q, policy) : 0.0;
    o.v[p][1] = ok ? ld_stream_hint(q + n3, policy) : 0.0;
    o.v[p][2] = ok ? ld_stream_hint(q + 2 * (size_t)n3, policy) : 0.0;
  }
}

__device__ __forceinline__ void load_row_chunk(const double *__restrict__ A, int rs, int n3, int base, int lane, RowVals &o) {
  const double *a0 = A + 9 * (size_t)rs + base;
#pragma unroll
  for (int p = 0; p < 3; p++) {
    const int t = base + lane + TILE_G * p;
    const bool ok = t < n3;
    const double *q = a0 + lane + TILE_G * p;
    o.v[p][0] = ok ? ld_stream(q) : 0.0;
    o.v[p][1] = ok ? ld_stream(q + n3) : 0.0;
    o.v[p][2] = ok ? ld_stream(q + 2 * (size_t)n3) : 0.0;
  }
}


// Deferred variant: the per-CTA sum goes to its slot and the CONSUMER kernel adds the slots (cta_sum_slots), every CTA
// redundantly and in the same fixed order.  This takes the ticket atomics and the serial last-CTA pass off the tail of
// the producer (the SpMV) — the slots are L2-resident and read in parallel by all CTAs of the next kernel.
template <int TB>
__device__ __forceinline__ void block_reduce_to_slot(double v, double *slots) {
  __shared__ double wsum1[TB / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) wsum1[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double s = (lane < TB / 32) ? wsum1[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) slots[blockIdx.x] = s;
  }
}

// same association order as the last-CTA pass of block_reduce_to_total: thread t adds slots t, t+TB, ..., then the tree
template <int TB>
__device__ __forceinline__ double cta_sum_slots(const double *__restrict__ slots, int n) {
  __shared__ double wsum2[TB / 32];
  __shared__ double bcast;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += TB) s += __ldcg(slots + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if (lane == 0) wsum2[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double z = (lane < TB / 32) ? wsum2[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_down_sync(0xffffffffu, z, o);
    if (lane == 0) bcast = z;
  }
  __syncthreads();
  return bcast;
}

// ---- peer-memory exchange between the ranks of a partitioned context (one process per GPU, NVLink/NVSwitch) --------
// Every rank owns a small "comm block" in device memory, mapped into all other ranks with CUDA IPC.  A producer kernel
// stores its value into slot [rank] of EVERY rank's block, fences at system scope, then stores the epoch into flag
// [rank]; a consumer kernel spins (bounded) on its LOCAL flags until all carry the epoch, then adds the slots in rank
// order — all ranks add the same bits in the same order, so alpha/beta/rho and the loop flag agree bit for bit.
// INSIDE a solve a rank cannot overwrite a slot before every other rank has consumed it, because its next write of that
// kind sits behind a wait on something each of them publishes only after consuming (fb_dist.cu).  ACROSS solves that
// argument has one hole: the first publish of solve S+1 (k_cg_init, rho0) waits for nothing, so a rank that has left solve
// S could overwrite its RHO slot while a lagging peer has not yet collected rho(final) of solve S.  Slots and flags are
// therefore double-buffered by the parity of the solve counter: to touch parity p again a rank must have FINISHED solve
// S+1, which needs every rank's publishes of S+1, which are stream-ordered after that rank's last collect of solve S.
// With that, a flag can never legitimately run ahead of the epoch a consumer waits for: `flag > epochWait` is reported
// as an overrun (comm_error = 2) instead of being taken as satisfied.
#define FB_MAX_RANKS 16
enum { FB_COMM_DQ = 0, FB_COMM_RHO = 1, FB_COMM_HALO = 2 };  // slot/flag families
// block layout in 8-byte words, per solve parity: slots[family][FB_MAX_RANKS], then flags[family][FB_MAX_RANKS]
#define FB_COMM_SLOT(par, fam, r) ((par) * 6 * FB_MAX_RANKS + (fam) * FB_MAX_RANKS + (r))
#define FB_COMM_FLAG(par, fam, r) ((par) * 6 * FB_MAX_RANKS + 3 * FB_MAX_RANKS + (fam) * FB_MAX_RANKS + (r))
#define FB_COMM_WORDS (12 * FB_MAX_RANKS)
#define FB_SPIN_LIMIT (1ll << 24)  // x ~100 ns: a lost peer turns into FB_ERR_COMM after ~2 s instead of a hung GPU

struct FbPeerArgs {
  int enabled, rank, world;
  int parity;                    // solve counter & 1: which half of the comm blocks this solve uses
  unsigned long long epoch;      // epoch of the value this kernel PUBLISHES (0 = none)
  unsigned long long epochWait;  // epoch of the value this kernel COLLECTS or waits for (0 = none)
  unsigned int haloMask;         // ranks whose halo flag the kernel waits for (SpMV)
  double *comm[FB_MAX_RANKS];
};

// Halo of the search direction pushed by the kernel that computes it (k_direction): per local vertex the list of
// (neighbour slot, neighbour-local vertex) it must be stored to; pushFlag skips the other vertices with one byte.
#define FB_MAX_NBR 8
struct FbPushArgs {
  int nNbr;                          // 0 = nothing to push (one GPU, NCCL path)
  int nbrRank[FB_MAX_NBR];
  double *peerVec[FB_MAX_NBR];       // the neighbours' direction vectors, mapped here
  const unsigned char *pushFlag;     // [nV]
  const int *pushPtr;                // [nV + 1]
  const int2 *pushEnt;               // (neighbour slot, neighbour-local vertex)
  unsigned long long epoch;          // halo epoch raised in the neighbours' comm blocks when all stores are out
};

// executed by ONE thread (the thread that holds the rank's total)
__device__ __forceinline__ void peer_publish(const FbPeerArgs &pa, int family, double value) {
  for (int p = 0; p < pa.world; p++) ((volatile double *)pa.comm[p])[FB_COMM_SLOT(pa.parity, family, pa.rank)] = value;
  __threadfence_system();
  for (int p = 0; p < pa.world; p++) ((volatile unsigned long long *)pa.comm[p])[FB_COMM_FLAG(pa.parity, family, pa.rank)] = pa.epoch;
}

// executed by ALL threads of a CTA (TB >= 32); returns the rank-ordered sum in every thread; on timeout marks the solve
template <int TB>
__device__ __forceinline__ double peer_collect(const FbPeerArgs &pa, int family, FbScalars *sc) {
  __shared__ double s_total;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const volatile unsigned long long *flags = (const volatile unsigned long long *)pa.comm[pa.rank];
    bool ok = true, overrun = false;
    if (lane < pa.world) {
      long long spins = 0;
      unsigned long long f;
      while ((f = flags[FB_COMM_FLAG(pa.parity, family, lane)]) < pa.epochWait) {
        __nanosleep(64);
        if (++spins > FB_SPIN_LIMIT) { ok = false; break; }
      }
      overrun = f > pa.epochWait;  // the producer has already published a LATER value into this slot
    }
    ok = __all_sync(0xffffffffu, ok);
    overrun = __any_sync(0xffffffffu, overrun);
    __threadfence_system();
    const double v = (lane < pa.world) ? ((const volatile double *)pa.comm[pa.rank])[FB_COMM_SLOT(pa.parity, family, lane)] : 0.0;
    double tot = 0.0;
    for (int r = 0; r < pa.world; r++) tot += __shfl_sync(0xffffffffu, v, r);
    if (lane == 0) {
      s_total = tot;
      if (!ok) { sc->comm_error = 1; sc->done = 1; }
      else if (overrun) { sc->comm_error = 2; sc->done = 1; }
    }
  }
  __syncthreads();
  const double t = s_total;
  __syncthreads();
  return t;
}

// executed by ALL threads of a CTA: wait until every rank in haloMask has pushed its halo for epochWait
__device__ __forceinline__ void peer_wait_halo(const FbPeerArgs &pa, FbScalars *sc) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const volatile unsigned long long *flags = (const volatile unsigned long long *)pa.comm[pa.rank];
    if (lane < pa.world && ((pa.haloMask >> lane) & 1u)) {
      long long spins = 0;
      while (flags[FB_COMM_FLAG(pa.parity, FB_COMM_HALO, lane)] < pa.epochWait) {
        __nanosleep(64);
        if (++spins > FB_SPIN_LIMIT) { sc->comm_error = 1; sc->done = 1; break; }
      }
    }
    __threadfence_system();
  }
  __syncthreads();
}
