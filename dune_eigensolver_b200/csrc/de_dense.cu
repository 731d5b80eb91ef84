// de_dense.cu -- tall-skinny dense kernels of the path and their launch logic: partial-sum reductions with fused tails
// (all-reduce, Cholesky, convergence test), diag-dot, Gram, block update / projection, CholQR2 (B-)orthonormalisation.
// Replaces dot_products_* / orthonormalize_* / B_orthonormalize_* of the reference (kernels_cpp.hh:7-96, :121-591).
#include <cstdlib>

#include "de_internal.hpp"
#include "kernels_dense.cuh"
#include "kernels_sparse.cuh"
#include "kernels_peer.cuh"
#include "kernels_tail.cuh"
#include "kernels_tallskinny.cuh"
#include "kernels_tallskinny2.cuh"

using namespace dei;

namespace dei
{
  // ---- reductions -----------------------------------------------------------------------------------
  int flush_pending_dot(de_context *ctx)
  {
    if (!ctx->pending_dot.valid)
      return DE_OK;
    const de_context::PendingDot p = ctx->pending_dot;
    ctx->pending_dot.valid = false;
    ctx->tail = p.conv;
    ctx->tail_armed = true;
    ctx->tail_did_allreduce = ctx->tail_did_op = false;
    DE_TRY(reduce_partials(ctx, ctx->partials, p.nparts, p.m, ctx->dDG()));
    DE_TRY(allreduce_sum(ctx, ctx->dDG(), (size_t)p.m));
    if (ctx->tail_did_op)
      ctx->tail_did_op = false;
    else
    {
      DE_REG(de::convergence_kernel), de::convergence_kernel<<<1, 64, 0, ctx->stream>>>(p.conv.k, p.conv.m, p.conv.shift, p.conv.tol, ctx->dDP(), p.conv.s_prev,
                                                        p.conv.hist, p.conv.flags);
      DE_LAUNCH_CHECK(ctx);
    }
    ctx->tail_armed = false;
    return DE_OK;
  }

  int reduce_partials(de_context *ctx, const double *partials, int nparts, int len, double *out)
  {
    dim3 block(32, 32);
    const bool multi = ctx->nranks > 1;
    if (ctx->pending_dot.valid)
    {
      const de_context::PendingDot p = ctx->pending_dot;
      const bool fusable = ctx->tail_armed && ctx->tail.kind == de::kTailChol && ctx->dtail_ticket &&
                           partials == ctx->partials + kDotPartialsReserve &&
                           (!multi || (ctx->peer_ready && len + p.m <= de::kPeerSlotDoubles));
      if (!fusable)
      {
        // cannot share a launch: the partials of this reduction sit behind the reserve, the deferred ones are flushed first
        DE_TRY(flush_pending_dot(ctx));
      }
      else
      {
        // [dp | G] <- reduce both partial sets -> ONE all-reduce -> convergence test -> Cholesky (kernels_tail.cuh).
        // The result lives in dDG() (dp at its head, G behind it), not in `out`: nothing reads G after the Cholesky.
        ProfScope prof(ctx, DE_PROF_SMALL);
        ctx->pending_dot.valid = false;
        de::TailArgs t = ctx->tail;
        t.kind = de::kTailChol | de::kTailConv;
        t.k = p.conv.k;
        t.shift = p.conv.shift;
        t.tol = p.conv.tol;
        t.s_prev = p.conv.s_prev;
        t.hist = p.conv.hist;
        t.flags = p.conv.flags;
        t.partials2 = ctx->partials;
        t.nparts2 = p.nparts;
        t.len2 = p.m;
        t.do_allreduce = multi ? 1 : 0;
        if (multi)
          t.pa = peer_args(ctx, ++ctx->ar_epoch);
        t.ticket = ctx->dtail_ticket;
        ctx->tail_armed = false;
        ctx->tail_did_allreduce = true;
        ctx->tail_did_op = true;
        DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::reduce_tail_kernel), dim3((len + p.m + 31) / 32), block, 0, ctx->stream, partials, nparts,
                                len, ctx->dDG(), ctx->done_ptr, t));
        DE_LAUNCH_CHECK(ctx);
        return DE_OK;
      }
    }
    ProfScope prof(ctx, DE_PROF_SMALL);
    if (ctx->tail_armed && ctx->dtail_ticket && (!multi || (ctx->peer_ready && len <= de::kPeerSlotDoubles)))
    {
      // reduce -> (all-reduce) -> Cholesky / convergence test in ONE launch
      de::TailArgs t = ctx->tail;
      t.do_allreduce = multi ? 1 : 0;
      if (multi && t.channel == 1)
      {
        t.pa = peer_args(ctx, ++ctx->ar_epoch_b); // the skippable all-reduce of the second CholQR sweep: its own channel
        t.pa.channel = 1;
      }
      else if (multi)
        t.pa = peer_args(ctx, ++ctx->ar_epoch);
      t.ticket = ctx->dtail_ticket;
      ctx->tail_armed = false;
      ctx->tail_did_allreduce = true;
      ctx->tail_did_op = true;
      DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::reduce_tail_kernel), dim3((len + 31) / 32), block, 0, ctx->stream, partials, nparts, len, out,
                              ctx->done_ptr, t));
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    ctx->tail_armed = false;
    DE_REG(de::reduce_partials_kernel), de::reduce_partials_kernel<<<(len + 31) / 32, block, 0, ctx->stream>>>(partials, nparts, len, out, ctx->done_ptr);
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  de::PeerArgs peer_args(de_context *ctx, unsigned long long epoch)
  {
    de::PeerArgs pa;
    pa.rank = ctx->rank;
    pa.nranks = ctx->nranks;
    for (int q = 0; q < de::kPeerMaxRanks; ++q)
      pa.base[q] = ctx->peer_base[q];
    pa.epoch = epoch;
    pa.channel = 0;
    pa.done = ctx->done_ptr;
    pa.err = ctx->dticket + 1;
    pa.timeout = ctx->peer_timeout_cycles;
    static const bool trace = std::getenv("DE_TRACE_PEER") != nullptr; // debugging aid only: no behaviour depends on it
    if (trace)
      std::fprintf(stderr, "[de peer] rank %d epoch %llu (ar %llu halo %llu)\n", ctx->rank, epoch, ctx->ar_epoch, ctx->halo_epoch);
    return pa;
  }

  int allreduce_sum(de_context *ctx, double *buf, size_t count)
  {
    if (ctx->tail_did_allreduce)
    {
      ctx->tail_did_allreduce = false; // the fused tail of the reduction already did it
      return DE_OK;
    }
    if (ctx->nranks <= 1)
      return DE_OK;
    if (ctx->peer_ready && count <= (size_t)de::kPeerSlotDoubles)
    {
      ProfScope prof(ctx, DE_PROF_SMALL);
      DE_REG(de::peer_allreduce_kernel), de::peer_allreduce_kernel<<<1, 1024, 0, ctx->stream>>>(peer_args(ctx, ++ctx->ar_epoch), buf, (int)count);
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    if (!ctx->comm)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "all-reduce: vector exceeds the peer window slot and the context has no NCCL communicator");
    DE_NCCL(ctx, nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    return DE_OK;
  }

  // ---- diag-dot ---------------------------------------------------------------------------------------
  int diag_dot_device(de_context *ctx, long long n, int m, const double *X, const double *Y, double *out)
  {
    const int hp = m / 2;
    dim3 block(hp, 256 / hp);
    const long long need = (n + block.y - 1) / block.y;
    const int grid = (int)std::max<long long>(1, std::min<long long>(need, kMaxPartials));
    double *part = reduction_partials(ctx);
    {
      ProfScope prof(ctx, DE_PROF_DOT);
      DE_REG(de::diag_dot_kernel), de::diag_dot_kernel<<<grid, block, 0, ctx->stream>>>(n, X, m, Y, m, m, part);
    }
    DE_LAUNCH_CHECK(ctx);
    DE_TRY(reduce_partials(ctx, part, grid, m, out));
    return allreduce_sum(ctx, out, m);
  }

  // ---- pipelined tall-skinny kernel (m = 8/16/32/64) -------------------------------------------------------

  template <int M, bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  int launch_ts_t(de_context *ctx, de::TsArgs a, double *gram_out)
  {
    if constexpr (DO_UPDATE && M == 64 && DO_GRAM && UPPER && SAME)
    {
      // no fused kernel at this width (72 Gram accumulators + the result block do not fit the register file: 168
      // registers with spills when tried): block update, then the Gram matrix of the result -- two passes, both on the
      // warp-specialised kernels
      de::TsArgs u = a;
      DE_TRY((launch_ts_t<M, true, false, false, true>(ctx, u, nullptr)));
      de::TsArgs g = a;
      g.X = a.Out;
      g.ldx = a.ldo;
      g.skip_flag = a.skip_flag; // (null except for the second sweep of a one-sweep-capable orthonormalisation)
      return launch_ts_t<M, false, true, true, true>(ctx, g, gram_out);
    }
    if constexpr (DO_UPDATE && (M <= 32 || !DO_GRAM) && (!DO_GRAM || (UPPER && SAME)))
    {
      // block update (+ Gram of the result): the register-to-register tensor-core kernel (kernels_tallskinny2.cuh)
      using C2 = de::Ts2Cfg<M>;
      DE_TRY(ensure_func_smem(ctx, (const void *)de::ts2_update_kernel<M, DO_GRAM>, C2::SMEM));
      const long long nt2 = (a.n + C2::TR - 1) / C2::TR;
      const int grid2 = (int)std::max<long long>(1, std::min<long long>(nt2, (long long)ctx->sm_count));
      a.partials = reduction_partials(ctx);
      a.done = ctx->done_ptr;
      a.push = de::PushRanges{};
      if (ctx->push_pending.n > 0 && a.ldo == M) // orthonormalize_device(..., next_spmm): halo rows go out with the update
        a.push = ctx->push_pending;
      if (a.push.n > 0)
      {
        DE_TRY(ensure_func_smem(ctx, (const void *)de::ts2_update_kernel<M, DO_GRAM, true>, C2::SMEM));
        ProfScope prof(ctx, DE_PROF_UPDATE);
        DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::ts2_update_kernel<M, DO_GRAM, true>), dim3(grid2), dim3(C2::THREADS), C2::SMEM, ctx->stream, a));
      }
      else
      {
        ProfScope prof(ctx, DE_PROF_UPDATE);
        DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::ts2_update_kernel<M, DO_GRAM>), dim3(grid2), dim3(C2::THREADS), C2::SMEM, ctx->stream, a));
      }
      DE_LAUNCH_CHECK(ctx);
      if (DO_GRAM)
        DE_TRY(reduce_partials(ctx, a.partials, grid2, M * M, gram_out));
      return DE_OK;
    }
    if constexpr (!DO_UPDATE && DO_GRAM && UPPER && SAME)
    {
      // G = X^T X of one block: warp-specialised tensor-core kernel (kernels_tallskinny2.cuh)
      using C3 = de::Tg2Cfg<M>;
      DE_TRY(ensure_func_smem(ctx, (const void *)de::ts2_gram_kernel<M>, C3::SMEM));
      const long long nt3 = (a.n + C3::TR - 1) / C3::TR;
      const int grid3 = (int)std::max<long long>(1, std::min<long long>(nt3, (long long)ctx->sm_count));
      a.partials = reduction_partials(ctx);
      a.done = ctx->done_ptr;
      {
        ProfScope prof(ctx, DE_PROF_GRAM);
        DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::ts2_gram_kernel<M>), dim3(grid3), dim3(de::kTg2Threads), C3::SMEM, ctx->stream, a));
      }
      DE_LAUNCH_CHECK(ctx);
      return reduce_partials(ctx, a.partials, grid3, M * M, gram_out);
    }
    constexpr int NOPS = (DO_GRAM && !SAME) ? 2 : 1;
    using C = de::TsCfg<M, UPPER, NOPS>;
    constexpr size_t smem = de::tall_skinny_smem_bytes<M, DO_UPDATE, DO_GRAM, UPPER, SAME>();
    DE_TRY(ensure_func_smem(ctx, (const void *)de::tall_skinny_kernel<M, DO_UPDATE, DO_GRAM, UPPER, SAME>, smem));
    int ctas_per_sm = 1; // resident CTAs per SM of this instantiation on this device (registers / shared memory decide)
    DE_TRY(func_occupancy(ctx, (const void *)de::tall_skinny_kernel<M, DO_UPDATE, DO_GRAM, UPPER, SAME>, C::THREADS, smem,
                          &ctas_per_sm));
    ctas_per_sm = std::min(ctas_per_sm, 2);
    const long long ntiles = (a.n + C::TR - 1) / C::TR;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * ctas_per_sm));
    a.partials = reduction_partials(ctx);
    a.done = ctx->done_ptr;
    {
      ProfScope prof(ctx, DO_UPDATE ? DE_PROF_UPDATE : DE_PROF_GRAM);
      DE_REG(de::tall_skinny_kernel<M, DO_UPDATE, DO_GRAM, UPPER, SAME>), de::tall_skinny_kernel<M, DO_UPDATE, DO_GRAM, UPPER, SAME><<<grid, C::THREADS, smem, ctx->stream>>>(a);
    }
    DE_LAUNCH_CHECK(ctx);
    if (DO_GRAM)
      DE_TRY(reduce_partials(ctx, a.partials, grid, M * M, gram_out));
    return DE_OK;
  }

  template <bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  int launch_ts(de_context *ctx, int w, const de::TsArgs &a, double *gram_out)
  {
    switch (w)
    {
    case 8:
      return launch_ts_t<8, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    case 16:
      return launch_ts_t<16, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    case 32:
      return launch_ts_t<32, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    case 64:
      return launch_ts_t<64, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "pipelined tall-skinny kernel: unsupported width");
  }

  // ---- Gram -------------------------------------------------------------------------------------------
  template <int M, bool UPPER, bool SAME>
  int launch_gram_t(de_context *ctx, long long n, const double *X, int ldx, const double *Y, int ldy, double *out)
  {
    using C = de::GramCfg<M, UPPER, SAME>;
    const long long ntiles = (n + C::TR - 1) / C::TR;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, kMaxPartials));
    double *part = reduction_partials(ctx);
    {
      ProfScope prof(ctx, DE_PROF_GRAM);
      DE_REG(de::gram_kernel<M, UPPER, SAME>), de::gram_kernel<M, UPPER, SAME><<<grid, C::THREADS, 0, ctx->stream>>>(n, X, ldx, Y, ldy, part);
    }
    DE_LAUNCH_CHECK(ctx);
    return reduce_partials(ctx, part, grid, M * M, out);
  }

  template <bool UPPER, bool SAME>
  int launch_gram_m(de_context *ctx, int w, long long n, const double *X, int ldx, const double *Y, int ldy, double *out)
  {
    switch (w)
    {
    case 8:
      return launch_gram_t<8, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 16:
      return launch_gram_t<16, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 24:
      return launch_gram_t<24, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 32:
      return launch_gram_t<32, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 40:
      return launch_gram_t<40, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 48:
      return launch_gram_t<48, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 56:
      return launch_gram_t<56, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 64:
      return launch_gram_t<64, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "gram: column count must be a multiple of 8 in [8,64]");
  }

  /** out (device, w*w) = X^T Y over the local rows, all-reduced over the ranks.
   *  symmetric: the result is known to be symmetric (only upper blocks are computed and mirrored). */
  int gram_device(de_context *ctx, int w, long long n, const double *X, int ldx, const double *Y, int ldy,
                  bool symmetric, double *out)
  {
    const bool same = (X == Y && ldx == ldy);
    if (!same && gram2_supported(w))
    {
      // wide two-operand Gram: tensor-pipe-bound, warp-specialised kernel (kernels_gram2.cuh); all tiles are computed
      DE_TRY(gram2_device(ctx, w, n, X, ldx, Y, ldy, out));
      return allreduce_sum(ctx, out, (size_t)w * w);
    }
    if (ts_supported(w))
    {
      de::TsArgs a{};
      a.n = n;
      a.X = X;
      a.ldx = ldx;
      a.Y = Y;
      a.ldy = ldy;
      if (symmetric && same)
        DE_TRY((launch_ts<false, true, true, true>(ctx, w, a, out)));
      else if (symmetric)
        DE_TRY((launch_ts<false, true, true, false>(ctx, w, a, out)));
      else if (same)
        DE_TRY((launch_ts<false, true, false, true>(ctx, w, a, out)));
      else
        DE_TRY((launch_ts<false, true, false, false>(ctx, w, a, out)));
      return allreduce_sum(ctx, out, (size_t)w * w);
    }
    if (symmetric && same)
      DE_TRY((launch_gram_m<true, true>(ctx, w, n, X, ldx, Y, ldy, out)));
    else if (symmetric)
      DE_TRY((launch_gram_m<true, false>(ctx, w, n, X, ldx, Y, ldy, out)));
    else if (same)
      DE_TRY((launch_gram_m<false, true>(ctx, w, n, X, ldx, Y, ldy, out)));
    else
      DE_TRY((launch_gram_m<false, false>(ctx, w, n, X, ldx, Y, ldy, out)));
    return allreduce_sum(ctx, out, (size_t)w * w);
  }

  // ---- block update -----------------------------------------------------------------------------------
  template <int M, int MODE>
  int launch_update_t(de_context *ctx, long long n, const double *X, int ldx, const double *R, double *Y, int ldy,
                      int upper)
  {
    using C = de::UpdCfg<M>;
    DE_TRY(ensure_func_smem(ctx, (const void *)de::update_kernel<M, MODE>, C::SMEM_BYTES));
    const long long ntiles = (n + C::TR - 1) / C::TR;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / C::SMEM_BYTES));
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    ProfScope prof(ctx, DE_PROF_UPDATE);
    DE_REG(de::update_kernel<M, MODE>), de::update_kernel<M, MODE><<<grid, C::THREADS, C::SMEM_BYTES, ctx->stream>>>(n, X, ldx, R, Y, ldy, upper);
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  template <int MODE>
  int update_device_t(de_context *ctx, int w, long long n, const double *X, int ldx, const double *R, double *Y, int ldy,
                      int upper, const int *skip_flag)
  {
    if (MODE == 0 && ts_supported(w))
    {
      de::TsArgs a{};
      a.n = n;
      a.X = X;
      a.ldx = ldx;
      a.R = R;
      a.Out = Y;
      a.ldo = ldy;
      a.upper = upper;
      a.skip_flag = skip_flag;
      return launch_ts<true, false, false, true>(ctx, w, a, nullptr);
    }
    switch (w)
    {
    case 8:
      return launch_update_t<8, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 16:
      return launch_update_t<16, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 24:
      return launch_update_t<24, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 32:
      return launch_update_t<32, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 40:
      return launch_update_t<40, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 48:
      return launch_update_t<48, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 56:
      return launch_update_t<56, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 64:
      return launch_update_t<64, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "block update: column count must be a multiple of 8 in [8,64]");
  }

  int update_device(de_context *ctx, int mode, int w, long long n, const double *X, int ldx, const double *R, double *Y,
                    int ldy, int upper, const int *skip_flag)
  {
    if (mode == 1 && ts_supported(w) && ldx == w && ldy == w && skip_flag == nullptr && ctx->use_lincomb2)
    {
      // projection Y -= X R (kernels_cpp.hh:335-348) on the tensor-core kernel: Y = Y + (-1) * X R
      const double *S[2] = {Y, X};
      const double *Cm[2] = {nullptr, R};
      return lincomb2_device(ctx, w, n, 2, S, Cm, Y, nullptr, true, -1.0);
    }
    return mode == 0 ? update_device_t<0>(ctx, w, n, X, ldx, R, Y, ldy, upper, skip_flag)
                     : update_device_t<1>(ctx, w, n, X, ldx, R, Y, ldy, upper, skip_flag);
  }

  // ---- (B-)orthonormalisation: CholQR2 ------------------------------------------------------------------
  int chol_inverse(de_context *ctx, int m, const double *G, double *Rinv, double *info, int *identity_flag)
  {
    if (ctx->tail_did_op)
    {
      ctx->tail_did_op = false; // done by the fused tail of the reduction that produced G
      return DE_OK;
    }
    ProfScope prof(ctx, DE_PROF_SMALL);
    if (m <= 32)
      DE_REG(de::chol_inverse2_kernel<32>), de::chol_inverse2_kernel<32><<<1, 1024, 0, ctx->stream>>>(m, G, Rinv, ctx->dstatus, info, identity_flag,
                                                                const_cast<int *>(ctx->done_ptr));
    else
      DE_REG(de::chol_inverse2_kernel<64>), de::chol_inverse2_kernel<64><<<1, 1024, 0, ctx->stream>>>(m, G, Rinv, ctx->dstatus, info, identity_flag,
                                                                const_cast<int *>(ctx->done_ptr));
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  /** the next partial-sum reduction (of a Gram matrix into G) is followed, in the same launch, by the all-reduce and by
   *  Rinv = chol(G)^-1 -- exactly what chol_inverse(ctx, m, G, Rinv, info, identity_flag) would do afterwards */
  void arm_chol_tail(de_context *ctx, int m, double *Rinv, double *info, int *identity_flag)
  {
    ctx->tail = de::TailArgs{};
    ctx->tail.kind = de::kTailChol;
    ctx->tail.m = m;
    ctx->tail.Rinv = Rinv;
    ctx->tail.status = ctx->dstatus;
    ctx->tail.info = info;
    ctx->tail.identity_flag = identity_flag;
    ctx->tail.done = const_cast<int *>(ctx->done_ptr);
    ctx->tail_armed = true;
    ctx->tail_did_allreduce = ctx->tail_did_op = false;
  }

  /** X <- X R^-1 (thin QR with positive-diagonal triangular R; reference orthonormalize_blocked,
   *  kernels_cpp.hh:180-351). Two CholQR sweeps over the WHOLE block: G = X^T X, R = chol(G), X <- X R^-1.
   *  The triangular factor of a full-rank block is unique, so the result equals the reference's block
   *  Gram-Schmidt up to round-off; the second sweep restores orthogonality to O(eps) for cond(X) < ~1e7. */
  int orthonormalize_device(de_context *ctx, long long n, int m, double *X, const double *G_ready, const de_matrix *next_spmm)
  {
    if (ts_supported(m))
    {
      // next_spmm: the caller's next distributed SpMM is A X with this X -- its halo rows leave with the block updates
      struct PushGuard
      {
        de_context *c;
        ~PushGuard() { c->push_pending.n = 0; }
      } push_guard{ctx};
      if (next_spmm != nullptr && next_spmm->n == n && plan_fused_push(ctx, next_spmm, m))
      {
        ctx->prepushed_X = X;
        ctx->prepushed_A = next_spmm;
        ctx->prepushed_epoch = ctx->halo_epoch + 1;
        ctx->prepushed_m = m;
        ctx->prepushed_released = false;
      }
      // sweep 1: G = X^T X (already known if the SpMM that produced X ran its Gram epilogue) ; R1 = chol(G) ;
      // X <- X R1^-1 fused with G2 = X^T X of the result
      // One sweep where one is enough (round 2): the tail of the first Gram reduction decides ON THE DEVICE, from the scaled
      // Gram matrix, whether cond is so small that X chol(G)^-1 is orthonormal to the accuracy the second sweep's own test
      // would accept (kernels_dense.cuh, kWellCond). If so a plain update runs, and the fused update + Gram, its reduction /
      // all-reduce / Cholesky tail and the last update all skip themselves -- in the steady state of the subspace iteration
      // that is every iteration. The launch sequence is the same either way (CUDA graph, several ranks in lock step).
      const bool multi = ctx->nranks > 1;
      const bool one_sweep_ok = ctx->use_one_sweep && ctx->dwell != nullptr && ctx->dtail_ticket != nullptr && G_ready == nullptr &&
                                (!multi || (ctx->peer_ready && (size_t)m * m + m <= (size_t)de::kPeerSlotDoubles));
      bool one_sweep = false;
      if (G_ready == nullptr)
      {
        arm_chol_tail(ctx, m, ctx->dR(), nullptr, nullptr);
        if (one_sweep_ok)
        {
          ctx->tail.wellcond = ctx->dwell;
          ctx->tail.flags_identity = ctx->dflags;
        }
        DE_TRY(gram_device(ctx, m, n, X, m, X, m, true, ctx->dG()));
        one_sweep = one_sweep_ok && ctx->tail_did_op; // the decision exists only if the fused tail ran
      }
      DE_TRY(chol_inverse(ctx, m, G_ready ? G_ready : ctx->dG(), ctx->dR(), nullptr, nullptr));
      de::TsArgs a{};
      a.n = n;
      a.X = X;
      a.ldx = m;
      a.R = ctx->dR();
      a.Out = X;
      a.ldo = m;
      a.upper = 1;
      if (one_sweep)
      {
        de::TsArgs u = a;
        u.skip_flag = ctx->dwell + 1; // runs iff one sweep is enough
        DE_TRY((launch_ts<true, false, false, true>(ctx, m, u, nullptr)));
        a.skip_flag = ctx->dwell;     // the fused update + Gram runs iff it is not
      }
      // sweep 2: R2 = chol(G2) ; X <- X R2^-1, skipped on the device when G2 = I to working precision. The factor of
      // sweep 1 is read at the start of the update kernel and overwritten by the tail of its reduction: the factor
      // fragments are in registers long before the last CTA of the reduction runs (it is a later launch).
      arm_chol_tail(ctx, m, ctx->dR(), nullptr, ctx->dflags);
      if (one_sweep)
      {
        ctx->tail.skip = ctx->dwell;
        ctx->tail.channel = 1;
      }
      DE_TRY((launch_ts<true, true, true, true>(ctx, m, a, ctx->dG())));
      DE_TRY(allreduce_sum(ctx, ctx->dG(), (size_t)m * m));
      DE_TRY(chol_inverse(ctx, m, ctx->dG(), ctx->dR(), nullptr, ctx->dflags));
      if (ctx->push_pending.n > 0)
      {
        ctx->push_pending.release = 1; // the last launch that can change X: it also raises the halo flags
        ctx->prepushed_released = true;
      }
      return update_device(ctx, 0, m, n, X, m, ctx->dR(), X, m, 1, ctx->dflags);
    }
    for (int sweep = 0; sweep < 2; ++sweep)
    {
      DE_TRY(gram_device(ctx, m, n, X, m, X, m, true, ctx->dG()));
      DE_TRY(chol_inverse(ctx, m, ctx->dG(), ctx->dR(), nullptr));
      DE_TRY(update_device(ctx, 0, m, n, X, m, ctx->dR(), X, m, 1));
    }
    return DE_OK;
  }

  /** X^T B X = I (reference B_orthonormalize_blocked, kernels_cpp.hh:356-591). BX = B X is formed once with the
   *  SpMM kernel and then carried through both sweeps with the same triangular factor (the reference keeps
   *  P = B V_k updated the same way, :527-539), so on return BX = B X for the new X. */
  int b_orthonormalize_device(de_context *ctx, const de_matrix *B, long long n, int m, double *X, double *BX,
                              bool want_info)
  {
    DE_TRY(spmm_device(ctx, B, X, BX, m, false));
    for (int sweep = 0; sweep < 2; ++sweep)
    {
      double *info = (want_info && sweep == 0) ? ctx->dInfo() : nullptr;
      arm_chol_tail(ctx, m, ctx->dR(), info, nullptr); // reduce -> all-reduce -> Cholesky in one launch (kernels_tail.cuh)
      DE_TRY(gram_device(ctx, m, n, X, m, BX, m, true, ctx->dG()));
      DE_TRY(chol_inverse(ctx, m, ctx->dG(), ctx->dR(), info));
      DE_TRY(update_device(ctx, 0, m, n, X, m, ctx->dR(), X, m, 1));
      DE_TRY(update_device(ctx, 0, m, n, BX, m, ctx->dR(), BX, m, 1));
    }
    return DE_OK;
  }

  int reset_status(de_context *ctx)
  {
    DE_CUDA(ctx, cudaMemsetAsync(ctx->dstatus, 0, sizeof(int), ctx->stream));
    return DE_OK;
  }

  /** copy `count` doubles of device scratch and the sticky status to the host and wait for them */
  int fetch_small(de_context *ctx, const double *dsrc, double *hdst, size_t count)
  {
    if (count > 0)
      DE_CUDA(ctx, cudaMemcpyAsync(ctx->hsmall, dsrc, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->hstatus, ctx->dstatus, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->hflags[8] = 0;
    if (ctx->peer_ready)
      DE_CUDA(ctx, cudaMemcpyAsync(ctx->hflags + 8, ctx->dticket + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->hflags[8] != 0)
    {
      const int w = ctx->hflags[8];
      return set_error(ctx, DE_ERR_NCCL,
                       std::string("NVLink peer window: rank ") + std::to_string(ctx->rank) + " gave up waiting for the " +
                           ((w & 15) == 2 ? "halo rows" : "all-reduce contribution") + " of rank " + std::to_string((w >> 4) & 15) +
                           " (epoch " + std::to_string((w >> 8) & 0xfff) + ", flag still at " + std::to_string((w >> 20) & 0x7ff) + "; this rank has issued " + std::to_string(ctx->ar_epoch) +
                           " all-reduces and " + std::to_string(ctx->halo_epoch) + " halo exchanges)");
    }
    if (count > 0 && hdst != ctx->hsmall)
      std::memcpy(hdst, ctx->hsmall, count * sizeof(double));
    if (*ctx->hstatus != 0)
      return set_error(ctx, DE_ERR_SINGULAR,
                       "orthonormalize: Gram matrix is not positive definite (pivot " + std::to_string(*ctx->hstatus - 1) +
                           "); the block is numerically rank deficient");
    return DE_OK;
  }

} // namespace dei

extern "C"
{

  int de_diag_dot(double *dp_host, const de_mv *X, const de_mv *Y)
  {
    if (!dp_host || !X || !Y)
      return set_error(nullptr, DE_ERR_INVALID, "de_diag_dot: null argument");
    de_context *ctx = X->ctx;
    if (X->n != Y->n)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of rows does not match"); // kernels_cpp.hh:30
    if (X->m != Y->m)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of columns does not match"); // :32
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    DE_TRY(diag_dot_device(ctx, X->n, X->m, X->d, Y->d, ctx->dDP()));
    return fetch_small(ctx, ctx->dDP(), dp_host, X->m);
  }

  int de_gram(double *G_host, const de_mv *X, const de_mv *Y)
  {
    if (!G_host || !X || !Y)
      return set_error(nullptr, DE_ERR_INVALID, "de_gram: null argument");
    de_context *ctx = X->ctx;
    if (X->n != Y->n)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of rows does not match"); // kernels_cpp.hh:62
    if (X->m != Y->m)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of columns does not match"); // :64
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    // X^T X is symmetric: only the upper block triangle is computed (and mirrored) when both operands are the same block
    DE_TRY(gram_device(ctx, X->m, X->n, X->d, X->m, Y->d, Y->m, X->d == Y->d, ctx->dG()));
    return fetch_small(ctx, ctx->dG(), G_host, (size_t)X->m * X->m);
  }

  int de_block_update(de_mv *X, const double *Q_host)
  {
    if (!X || !Q_host)
      return set_error(nullptr, DE_ERR_INVALID, "de_block_update: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    const size_t cnt = (size_t)X->m * X->m;
    std::memcpy(ctx->hsmall, Q_host, cnt * sizeof(double));
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->dR(), ctx->hsmall, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DE_TRY(update_device(ctx, 0, X->m, X->n, X->d, X->m, ctx->dR(), X->d, X->m, 0));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // hsmall may be reused by the next call
    return DE_OK;
  }

  int de_block_project(de_mv *X, int j0, int k0, int w, const double *S_host)
  {
    if (!X || !S_host)
      return set_error(nullptr, DE_ERR_INVALID, "de_block_project: null argument");
    de_context *ctx = X->ctx;
    if (w <= 0 || w % 8 != 0 || j0 % 8 != 0 || k0 % 8 != 0 || j0 < 0 || k0 < 0 || j0 + w > X->m || k0 + w > X->m ||
        (j0 < k0 + w && k0 < j0 + w))
      return set_error(ctx, DE_ERR_INVALID, "de_block_project: panels must be disjoint, 8-aligned and inside the block");
    DE_TRY(bind_device(ctx));
    const size_t cnt = (size_t)w * w;
    std::memcpy(ctx->hsmall, S_host, cnt * sizeof(double));
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->dR(), ctx->hsmall, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DE_TRY(update_device(ctx, 1, w, X->n, X->d + k0, X->m, ctx->dR(), X->d + j0, X->m, 0));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_orthonormalize(de_mv *X)
  {
    if (!X)
      return set_error(nullptr, DE_ERR_INVALID, "de_orthonormalize: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    DE_TRY(orthonormalize_device(ctx, X->n, X->m, X->d));
    return fetch_small(ctx, nullptr, nullptr, 0);
  }

  int de_b_orthonormalize(const de_matrix *B, de_mv *X, de_mv *BX, double *norm)
  {
    if (!B || !X)
      return set_error(nullptr, DE_ERR_INVALID, "de_b_orthonormalize: null argument");
    de_context *ctx = X->ctx;
    if (B->n != X->n || (BX && (BX->n != X->n || BX->m != X->m)))
      return set_error(ctx, DE_ERR_INVALID, "B_orthonormalize: shape mismatch");
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    ScopedBlocks tmp;
    double *bx = BX ? BX->d : nullptr;
    if (!bx)
      DE_TRY(tmp.alloc(ctx, &bx, (size_t)X->n * X->m));
    DE_TRY(b_orthonormalize_device(ctx, B, X->n, X->m, X->d, bx, norm != nullptr));
    double info = 0.0;
    DE_TRY(fetch_small(ctx, ctx->dInfo(), &info, norm ? 1 : 0));
    if (norm)
      *norm = info;
    return DE_OK;
  }

} // extern "C"
