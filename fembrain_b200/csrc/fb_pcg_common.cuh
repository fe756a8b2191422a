// fb_pcg_common.cuh — device helpers shared by the PCG translation units.
#pragma once
#include "fb_internal.h"

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

