// de_dense64.cu -- wide-block (M = 64) tall-skinny kernels that are bound by the FP64 tensor pipe rather than by HBM:
// the two-operand Gram matrix X^T Y (kernels_gram2.cuh). Kept in a translation unit of its own (build time).
#include "de_internal.hpp"
#include "kernels_sparse.cuh"
#include "kernels_gram2.cuh"
#include "kernels_lincomb2.cuh"

using namespace dei;

namespace dei
{
  template <int M>
  static int launch_gram2_t(de_context *ctx, de::TsArgs a, double *out)
  {
    using C = de::Tg3Cfg<M>;
    DE_TRY(ensure_func_smem(ctx, (const void *)de::ts2_gram2_kernel<M>, C::SMEM));
    const long long nt = (a.n + C::TR - 1) / C::TR;
    const int grid = (int)std::max<long long>(1, std::min<long long>(nt, (long long)ctx->sm_count));
    a.partials = reduction_partials(ctx);
    a.done = ctx->done_ptr;
    {
      ProfScope prof(ctx, DE_PROF_GRAM);
      DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::ts2_gram2_kernel<M>), dim3(grid), dim3(de::kTg2Threads), C::SMEM, ctx->stream, a));
    }
    DE_LAUNCH_CHECK(ctx);
    return reduce_partials(ctx, a.partials, grid, M * M, out);
  }

  template <int M>
  static int launch_lincomb2_t(de_context *ctx, const de::Lc2Args &a)
  {
    using C = de::Lc2Cfg<M>;
    DE_TRY(ensure_func_smem(ctx, (const void *)de::ts2_lincomb_kernel<M>, C::SMEM));
    const long long nt = (a.n + C::TR - 1) / C::TR;
    const int grid = (int)std::max<long long>(1, std::min<long long>(nt, (long long)ctx->sm_count));
    {
      ProfScope prof(ctx, DE_PROF_UPDATE);
      DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::ts2_lincomb_kernel<M>), dim3(grid), dim3(C::THREADS), C::SMEM, ctx->stream, a));
    }
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  /** out = sum_s S_s C_s, out2 = sum_{s >= 1} S_s C_s (identity0: out = S_0 + alpha * sum_{s >= 1} S_s C_s) on the tensor-core
   *  kernel of kernels_lincomb2.cuh; widths 8 / 16 / 32 / 64, blocks with leading dimension w */
  int lincomb2_device(de_context *ctx, int w, long long n, int ns, const double *const *S, const double *const *Cm, double *out,
                      double *out2, bool identity0, double alpha)
  {
    if (ns < 1 || ns > de::kLc2MaxSrc)
      return set_error(ctx, DE_ERR_INVALID, "lincomb: 1 to 3 sources");
    de::Lc2Args a{};
    a.n = n;
    a.ns = ns;
    for (int s = 0; s < ns; ++s)
    {
      a.S[s] = S[s];
      a.C[s] = Cm[s];
    }
    a.out = out;
    a.out2 = ns > 1 ? out2 : nullptr;
    a.identity0 = identity0 ? 1 : 0;
    a.alpha = alpha;
    a.done = ctx->done_ptr;
    switch (w)
    {
    case 8:
      return launch_lincomb2_t<8>(ctx, a);
    case 16:
      return launch_lincomb2_t<16>(ctx, a);
    case 32:
      return launch_lincomb2_t<32>(ctx, a);
    case 64:
      return launch_lincomb2_t<64>(ctx, a);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "tensor-core lincomb kernel: unsupported width");
  }

  bool gram2_supported(int w) { return w == 64; }

  int gram2_device(de_context *ctx, int w, long long n, const double *X, int ldx, const double *Y, int ldy, double *out)
  {
    de::TsArgs a{};
    a.n = n;
    a.X = X;
    a.ldx = ldx;
    a.Y = Y;
    a.ldy = ldy;
    switch (w)
    {
    case 64:
      return launch_gram2_t<64>(ctx, a, out);
    case 32:
      return launch_gram2_t<32>(ctx, a, out);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "two-operand tensor-core Gram kernel: unsupported width");
  }
} // namespace dei
