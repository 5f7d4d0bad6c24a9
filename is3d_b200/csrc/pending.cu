// Entry points whose kernels are not written yet: they fail loudly (never fall back to a CPU path).
#include "ctx.h"

namespace is3d {
static is3d_status pending(is3d_ctx *ctx, const char *what)
{
  ctx->set_error(std::string(what) + ": CUDA kernel not implemented in this build");
  return IS3D_ERR_UNSUPPORTED;
}
is3d_status run_spectra_famod(is3d_ctx *ctx, double *, is3d_stats *) { return pending(ctx, "spectra df_mode 5"); }
is3d_status run_total_yield(is3d_ctx *ctx, double *, is3d_stats *) { return pending(ctx, "total yield"); }
is3d_status run_cell_yields(is3d_ctx *ctx, double *, double *, is3d_stats *) { return pending(ctx, "cell yields"); }
is3d_status run_sampler(is3d_ctx *ctx, int64_t, is3d_particle **, int64_t *, int64_t *, is3d_stats *) { return pending(ctx, "sampler"); }
}  // namespace is3d

extern "C" {
is3d_status is3d_total_yield(is3d_ctx *ctx, double *, is3d_stats *) { return is3d::pending(ctx, "total yield"); }
is3d_status is3d_cell_yields(is3d_ctx *ctx, double *, double *, is3d_stats *) { return is3d::pending(ctx, "cell yields"); }
is3d_status is3d_sample(is3d_ctx *ctx, int64_t, is3d_particle **, int64_t *, int64_t *, is3d_stats *) { return is3d::pending(ctx, "sampler"); }
void is3d_free_particles(is3d_particle *p) { (void)p; }
is3d_status is3d_sample_histograms(is3d_ctx *ctx, double *, double *, double *, double *, double *, double *, double *, double *, double *, double *) { return is3d::pending(ctx, "sampler histograms"); }
}
