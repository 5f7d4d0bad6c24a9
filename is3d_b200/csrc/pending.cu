// Entry points whose kernels are not written yet: they fail loudly (never fall back to a CPU path).
#include "ctx.h"

namespace is3d {
static is3d_status pending(is3d_ctx *ctx, const char *what)
{
  ctx->set_error(std::string(what) + ": CUDA kernel not implemented in this build");
  return IS3D_ERR_UNSUPPORTED;
}
is3d_status run_spectra_famod(is3d_ctx *ctx, double *, is3d_stats *) { return pending(ctx, "spectra df_mode 5"); }
}  // namespace is3d
