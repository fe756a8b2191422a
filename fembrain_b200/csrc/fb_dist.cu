// fb_dist.cu — partitioned (multi-GPU) contexts: NCCL halo exchange and scalar all-reduce.
// (placeholder: filled in once the single-GPU path is parity-green)
#include "fb_internal.h"

int fb_dist_halo_exchange(fb_context *, double *) { return FB_ERR_NOT_SUPPORTED; }
int fb_dist_allreduce_scalar(fb_context *, double *) { return FB_ERR_NOT_SUPPORTED; }
void fb_dist_destroy(fb_context *) {}

extern "C" {
int fb_comm_unique_id(void *) { fb_set_error("partitioned contexts not built yet"); return FB_ERR_NOT_SUPPORTED; }
int fb_create_partitioned(fb_context **, int, const double *, int, const int *, int, const int *, const fb_params *, int, int,
                          const void *) {
  fb_set_error("partitioned contexts not built yet");
  return FB_ERR_NOT_SUPPORTED;
}
int fb_partition_range(const fb_context *c, int *b, int *e) {
  if (!c) return FB_ERR_INVALID_ARGUMENT;
  if (b) *b = 0;
  if (e) *e = c->nV;
  return FB_OK;
}
}
