// fb_experiments_off.cu — the default build: the measured-and-shelved SpMV / PCG experiments are NOT in the product library.
//
// Three round-1 experiments live under csrc/experiments/ with their measurements under profiles/ and are compiled only by
// `python -m fembrain_b200.build --experiments` (FEMBRAIN_B200_BUILD_EXPERIMENTS=1):
//   fb_sym.cu             products from the block-upper triangle of Keff   48.6 vs 38.5 us at 1M tets, 412 vs 366 at 10M  (lost)
//   fb_pcg_persistent.cu  whole PCG loop in one cooperative kernel          79.5 vs 55.3 us per iteration at 1M            (lost)
//   fb_tma.cu             matrix stream staged by cp.async.bulk + mbarrier  product alone 331 vs 352 us at 10M (0.94 of the copy
//                         peak), but 471 vs 467 us per PCG iteration; 43.6 vs 38.8 us at 1M (profiles/r02_spmv_variants_*.txt)
// These stubs answer "not available", so FEMBRAIN_B200_SPMV=sym|tma and FEMBRAIN_B200_PCG=persistent fall back to the default
// kernels of fb_pcg.cu.
#include "fb_internal.h"

extern "C" int fb_experiments_built(void) { return 0; }
int fb_pcg_plan_persistent(fb_context *c) { c->pers_grid = 0; return FB_OK; }
int fb_pcg_launch_persistent(fb_context *) { fb_set_error("persistent PCG kernel not built (experiments)"); return FB_ERR_NOT_SUPPORTED; }
int fb_sym_plan(fb_context *c) { c->sym = nullptr; return FB_OK; }
int fb_sym_pack(fb_context *) { return FB_OK; }
void fb_sym_launch(fb_context *, int, const double *, double *, const double *, double *) {}
int fb_sym_grid(const fb_context *, int) { return 0; }
size_t fb_sym_bytes_per_product(const fb_context *) { return 0; }
void fb_sym_release(fb_context *) {}
int fb_tma_plan(fb_context *c) { c->tma = nullptr; return FB_OK; }
void fb_tma_launch(fb_context *, int, const double *, double *, const double *, double *) {}
int fb_tma_grid(const fb_context *) { return 0; }
int fb_tma_failed(fb_context *) { return 0; }
void fb_tma_release(fb_context *) {}
