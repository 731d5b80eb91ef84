// de_drivers.cu -- the device-resident driver loops: StandardLargest / StandardInverse / GeneralizedInverse (reference
// eigensolver.hh:28-112, :116-198, :204-351) and the LOBPCG drivers (new) with their combination kernels.
#include "de_internal.hpp"
#include "kernels_sparse.cuh" // convergence_kernel
#include "kernels_lobpcg.cuh"
#include "lobpcg_core.hpp"
#include "host_eig.hpp"

using namespace dei;

namespace dei
{
  /** The drivers could take G = Y^T Y from the SpMM epilogue instead of a separate Gram pass (de_spmm_gram does).
   *  Measured on B200 (100^3 Q1, m = 32): the epilogue adds 0.12 ms to a 0.19 ms SpMM, the separate pass costs
   *  0.056 ms + one launch. Off until the epilogue is cheaper. */
  constexpr bool kFuseGramIntoSpmm = false;

  // ---- LOBPCG: device implementation of the Ops interface of lobpcg_core.hpp ----------------------------------
  template <int M>
  int launch_lincomb_t(de_context *ctx, long long n, int ns, const double *const *S, const double *C, double *out,
                       double *out2)
  {
    using K = de::LinCfg<M>;
    DE_TRY(ensure_func_smem(ctx, (const void *)de::lincomb_kernel<M>, K::SMEM_BYTES));
    const long long ntiles = (n + K::TR - 1) / K::TR;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (size_t)(200 * 1024) / K::SMEM_BYTES));
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    ProfScope prof(ctx, DE_PROF_UPDATE);
    DE_REG(de::lincomb_kernel<M>), de::lincomb_kernel<M><<<grid, K::THREADS, K::SMEM_BYTES, ctx->stream>>>(n, ns, S[0], ns > 1 ? S[1] : nullptr,
                                                                           ns > 2 ? S[2] : nullptr, C, out, out2);
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  int lincomb_device(de_context *ctx, int m, long long n, int ns, const double *const *S, const double *C, double *out,
                     double *out2)
  {
    if (ts_supported(m) && ctx->use_lincomb2)
    {
      const double *Cm[3] = {C, C + (size_t)m * m, C + (size_t)2 * m * m};
      return lincomb2_device(ctx, m, n, ns, S, Cm, out, out2, false, 1.0);
    }
    switch (m)
    {
    case 8:
      return launch_lincomb_t<8>(ctx, n, ns, S, C, out, out2);
    case 16:
      return launch_lincomb_t<16>(ctx, n, ns, S, C, out, out2);
    case 24:
      return launch_lincomb_t<24>(ctx, n, ns, S, C, out, out2);
    case 32:
      return launch_lincomb_t<32>(ctx, n, ns, S, C, out, out2);
    case 40:
      return launch_lincomb_t<40>(ctx, n, ns, S, C, out, out2);
    case 48:
      return launch_lincomb_t<48>(ctx, n, ns, S, C, out, out2);
    case 56:
      return launch_lincomb_t<56>(ctx, n, ns, S, C, out, out2);
    case 64:
      return launch_lincomb_t<64>(ctx, n, ns, S, C, out, out2);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "lincomb: column count must be a multiple of 8 in [8,64]");
  }

  struct LobpcgDeviceOps
  {
    using Blk = double *;
    de_context *ctx = nullptr;
    const de_matrix *A = nullptr, *B = nullptr;
    const de_factor *T = nullptr; // optional preconditioner: W <- T^-1 W (factored apply, kernels_trsv.cuh)
    long long n = 0;
    int m = 0;
    ScopedBlocks blocks;
    double *dcoef = nullptr;  // 3 m^2 coefficients | m Ritz values
    double *dgrams = nullptr; // 12 m^2

    int init()
    {
      DE_TRY(blocks.alloc(ctx, &dcoef, (size_t)3 * m * m + m));
      DE_TRY(blocks.alloc(ctx, &dgrams, (size_t)12 * m * m));
      return DE_OK;
    }
    int alloc(Blk *b) { return blocks.alloc(ctx, b, (size_t)n * m); }
    /** wait for the stream; surfaces the sticky Cholesky status and the peer-window error flag */
    int sync_check() { return fetch_small(ctx, nullptr, nullptr, 0); }
    int orthonormalize(Blk X, Blk BX)
    {
      DE_TRY(reset_status(ctx));
      if (B)
        DE_TRY(b_orthonormalize_device(ctx, B, n, m, X, BX, false));
      else
        DE_TRY(orthonormalize_device(ctx, n, m, X));
      return sync_check();
    }
    int apply_A(Blk Y, Blk X) { return spmm_device(ctx, A, X, Y, m, false); }
    int apply_B(Blk Y, Blk X) { return spmm_device(ctx, B, X, Y, m, false); }
    int residual(Blk W, Blk AX, Blk BX, const double *theta, double *norm2)
    {
      double *dtheta = dcoef + (size_t)3 * m * m;
      DE_CUDA(ctx, cudaMemcpyAsync(dtheta, theta, sizeof(double) * m, cudaMemcpyHostToDevice, ctx->stream));
      const long long pairs = n * m / 2;
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        DE_REG(de::residual_kernel), de::residual_kernel<<<elementwise_grid(pairs), 256, 0, ctx->stream>>>(pairs, m, AX, BX, dtheta, W);
      }
      DE_LAUNCH_CHECK(ctx);
      DE_TRY(diag_dot_device(ctx, n, m, W, W, ctx->dDP()));
      return fetch_small(ctx, ctx->dDP(), norm2, m);
    }
    int precondition(Blk W)
    {
      if (!T)
        return DE_OK;
      // in place: the apply permutes W into the factor's own work block before anything is written back
      return factor_apply_device(ctx, T, W, W, m);
    }
    int elementwise_grid(long long pairs) const
    {
      return (int)std::max<long long>(1, std::min<long long>((pairs + 255) / 256, (long long)ctx->sm_count * 8));
    }
    /** Jacobi scale dinv = 1 / diag(A) and the Gershgorin bound of D^-1 A over all ranks' rows. The all-reduce of this
     *  library sums, so every rank deposits its local maximum in its own slot of a zeroed vector and the maximum is
     *  taken on the host. */
    double *ddinv = nullptr;
    int spectral_bound(double *b)
    {
      const int nr = std::max(1, ctx->nranks);
      double *slots = ctx->dDP(); // nr <= 64 doubles of scratch
      DE_TRY(blocks.alloc(ctx, &ddinv, (size_t)std::max<long long>(n, 1)));
      DE_CUDA(ctx, cudaMemsetAsync(slots, 0, sizeof(double) * nr, ctx->stream));
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        DE_REG(de::gershgorin_kernel), de::gershgorin_kernel<<<elementwise_grid(A->n), 256, 0, ctx->stream>>>(
            A->n, A->rowptr, A->col, A->val, ddinv, reinterpret_cast<unsigned long long *>(slots + ctx->rank));
      }
      DE_LAUNCH_CHECK(ctx);
      DE_TRY(allreduce_sum(ctx, slots, nr));
      std::vector<double> h(nr, 0.0);
      DE_TRY(fetch_small(ctx, slots, h.data(), nr));
      *b = *std::max_element(h.begin(), h.end()); // +inf if some row has no positive diagonal entry: the caller then
      return DE_OK;                               // runs without the preconditioner (lobpcg_core.hpp)
    }
    /** (m/2, 256/(m/2)) thread blocks of the row-wise streaming kernels */
    dim3 row_block() const { return dim3((unsigned)(m / 2), (unsigned)(256 / (m / 2))); }
    int row_grid() const
    {
      const long long rpb = 256 / (m / 2);
      return (int)std::max<long long>(1, std::min<long long>((n + rpb - 1) / rpb, (long long)ctx->sm_count * 8));
    }
    int cheb_start(Blk Z, Blk Zold, Blk R, double s)
    {
      ProfScope prof(ctx, DE_PROF_MISC);
      DE_REG(de::cheb_start_kernel), de::cheb_start_kernel<<<row_grid(), row_block(), 0, ctx->stream>>>(n, m / 2, s, ddinv, R, Z, Zold);
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    int cheb_step(Blk Zold, Blk Z, Blk R, Blk AZ, double alpha, double beta)
    {
      ProfScope prof(ctx, DE_PROF_MISC);
      DE_REG(de::cheb_step_kernel), de::cheb_step_kernel<<<row_grid(), row_block(), 0, ctx->stream>>>(n, m / 2, alpha, beta, ddinv, Z, R, AZ, Zold);
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    /** one Chebyshev step: AZ = A Z ; Zold <- Z + alpha (Z - Zold) + beta D^-1 (R - AZ) -- as the epilogue of the SpMM kernel
     *  where the matrix has its tensor-core form (AZ is then not touched), else as two passes */
    int apply_A_cheb(Blk Zold, Blk Z, Blk R, Blk AZ, double alpha, double beta)
    {
      bool fused = false;
      DE_TRY(spmm_cheb_device(ctx, A, Z, Zold, R, ddinv, alpha, beta, m, &fused));
      if (fused)
        return DE_OK;
      DE_TRY(apply_A(AZ, Z));
      return cheb_step(Zold, Z, R, AZ, alpha, beta);
    }
    int project(Blk W, Blk X, Blk BX)
    {
      DE_TRY(gram_device(ctx, m, n, BX, m, W, m, false, ctx->dG()));
      return update_device(ctx, 1, m, n, X, m, ctx->dG(), W, m, 0); // W -= X G
    }
    int grams(int count, const Blk *L, const Blk *R, const char *sym, double *out)
    {
      for (int g = 0; g < count; ++g)
        DE_TRY(gram_device(ctx, m, n, L[g], m, R[g], m, sym[g] != 0, dgrams + (size_t)g * m * m));
      DE_CUDA(ctx, cudaMemcpyAsync(out, dgrams, sizeof(double) * (size_t)count * m * m, cudaMemcpyDeviceToHost,
                                   ctx->stream));
      return sync_check();
    }
    int rotate(Blk X, const double *C)
    {
      DE_CUDA(ctx, cudaMemcpyAsync(dcoef, C, sizeof(double) * (size_t)m * m, cudaMemcpyHostToDevice, ctx->stream));
      const double *S[1] = {X};
      return lincomb_device(ctx, m, n, 1, S, dcoef, X, nullptr);
    }
    int lincomb(int ns, const Blk *S, const double *C, Blk out, Blk out2)
    {
      DE_CUDA(ctx, cudaMemcpyAsync(dcoef, C, sizeof(double) * (size_t)ns * m * m, cudaMemcpyHostToDevice, ctx->stream));
      const double *src[3] = {S[0], ns > 1 ? S[1] : nullptr, ns > 2 ? S[2] : nullptr};
      return lincomb_device(ctx, m, n, ns, src, dcoef, out, out2);
    }
  };

  /** LOBPCG on the device block X (n x m, start block on entry, Ritz vectors on return) */
  int lobpcg_device(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *T, bool largest,
                    int cheb_degree, double tol, int maxiter, int nev, int m, double *X, de::LobpcgResult &res,
                    int verbose)
  {
    LobpcgDeviceOps ops;
    ops.ctx = ctx;
    ops.A = A;
    ops.B = B;
    ops.T = T;
    ops.n = A->n;
    ops.m = m;
    DE_TRY(reset_status(ctx));
    DE_TRY(ops.init());
    de::LobpcgParams prm;
    prm.m = m;
    prm.nev = nev;
    prm.tol = tol;
    prm.maxiter = maxiter;
    prm.verbose = verbose;
    prm.has_B = B != nullptr;
    prm.largest = largest;
    prm.cheb_degree = (T || largest) ? 0 : std::max(0, cheb_degree); // a factored preconditioner takes precedence
    prm.name = B ? "GeneralizedLOBPCG" : "StandardLOBPCG";
    const int rc = de::lobpcg_run(ops, prm, X, res);
    if (rc == de::kLobpcgRitzFailed)
      return set_error(ctx, DE_ERR_SINGULAR,
                       "LOBPCG: the Rayleigh-Ritz problem on [X W] is numerically singular (or residuals are not finite)");
    if (rc != DE_OK)
      return rc;
    return ops.sync_check();
  }

  int lobpcg_check_args(de_context *ctx, const char *who, const de_matrix *A, const de_matrix *B, const de_factor *T,
                        int nev, int m)
  {
    if (!valid_cols(m))
      return set_error(ctx, DE_ERR_UNSUPPORTED, std::string(who) + ": nev exceeds DE_MAX_COLS (64)");
    if (nev <= 0 || nev > m)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": nev must be in [1, number of columns]");
    if ((B && B->n != A->n) || (T && T->n != A->n))
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": A, B and the preconditioner must have the same size");
    if (T && ctx->nranks > 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, std::string(who) + ": the factored preconditioner is single-GPU");
    return DE_OK;
  }

  int lobpcg_driver(de_context *ctx, const char *who, const de_matrix *A, const de_matrix *B, const de_factor *T,
                    double tol, int maxiter, int nev, const double *start_panel8, double *eval, double *evec, int verbose,
                    int *iterations)
  {
    if (!ctx || !A || !start_panel8 || !eval || !evec || nev <= 0)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": bad arguments");
    const int m = padded_cols(nev);
    DE_TRY(lobpcg_check_args(ctx, who, A, B, T, nev, m));
    DE_TRY(bind_device(ctx));
    const long long n = A->n;
    ScopedBlocks blk;
    double *X;
    DE_TRY(blk.alloc(ctx, &X, (size_t)n * m));
    DE_TRY(upload_panel8_device(ctx, n, m, start_panel8, X));
    de::LobpcgResult res;
    const auto t0 = std::chrono::steady_clock::now();
    DE_TRY(lobpcg_device(ctx, A, B, T, false, DE_LOBPCG_DEFAULT_CHEB_DEGREE, tol, maxiter, nev, m, X, res, verbose));
    if (iterations)
      *iterations = res.iterations;
    if (verbose > 0) // one summary line in the style of eigensolver.hh:345-350
    {
      double worst = 0.0;
      for (int j = 0; j < nev; ++j)
        worst = std::max(worst, res.resnorm[j] / std::max(std::abs(res.theta[j]), std::numeric_limits<double>::min()));
      std::printf("%s:  time_total=%g iterations=%d restarts=%d relres=%g\n", who,
                  std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), res.iterations,
                  res.restarts, worst);
    }
    return copy_out(ctx, n, m, nev, X, res.theta, eval, evec);
  }

} // namespace dei

extern "C"
{

  // ---- drivers --------------------------------------------------------------------------------------------------
  /** shared skeleton of StandardLargest (eigensolver.hh:28-112) and StandardInverse (:116-198) on device blocks.
   *  On entry Qa holds the start block; on exit Qa holds the orthonormal iterate (reference Q1 after the swap)
   *  and Qb the block it was mapped to (reference Q2). For the largest-eigenvalue variant the product A*Qa that
   *  the reference recomputes at the top of the loop (:78) is the one it already formed for the Rayleigh
   *  quotients (:84) -- it is reused, bit-identically. */
  /** Asynchronous form of the StandardLargest loop (no factor): the convergence test runs on the device
   *  (convergence_kernel) and raises a flag that turns every kernel of the iterations already enqueued into a no-op,
   *  so the host never waits for the GPU inside the loop -- it enqueues kPollEvery iterations, requests a copy of the
   *  flags, and only looks at the copy requested one batch earlier. The blocks are frozen in the state of the
   *  converged iteration; which buffer is which follows from the parity of the iteration count. */
  static int standard_core_async(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, int m,
                                 double *&Qa, double *&Qb, std::vector<double> &s2, int verbose, int *k_exit_out)
  {
    constexpr int kPollEvery = 4;
    const long long n = A->n;
    struct Guard
    {
      de_context *c;
      ~Guard()
      {
        c->done_ptr = nullptr;
        c->defer_dot = false;
        c->pending_dot.valid = false;
      }
    } guard{ctx};
    const size_t need = 64 + (size_t)std::min(std::max(maxiter, 2), de::kConvHistory) + 1; // s_prev | capped history
    if (ctx->dconv_cap < need)
    {
      dev_free(ctx->dconv);
      ctx->dconv = nullptr;
      ctx->dconv_cap = 0;
      DE_TRY(dev_alloc(ctx, &ctx->dconv, need));
      ctx->dconv_cap = need;
    }
    double *s_prev = ctx->dconv, *hist = ctx->dconv + 64;
    DE_TRY(reset_status(ctx));
    DE_CUDA(ctx, cudaMemsetAsync(ctx->dflags, 0, 4 * sizeof(int), ctx->stream));
    DE_CUDA(ctx, cudaMemsetAsync(ctx->dconv, 0, need * sizeof(double), ctx->stream));
    ctx->done_ptr = ctx->dflags + 1;
    // one all-reduce wait less per iteration: the Rayleigh quotients travel with the next Gram matrix
    ctx->defer_dot = ts_supported(m) && (ctx->nranks <= 1 || ctx->peer_ready) && ctx->dtail_ticket != nullptr;
    DE_TRY(orthonormalize_device(ctx, n, m, Qa)); // (:69)
    s2.assign(m, 0.0);
    int enqueued = 0;
    bool have_product = false, finished = false;
    bool have_gram = false; // dDG() + m holds Qb^T Qb of the block the next orthonormalisation works on
    int pending[2] = {0, 0}; // poll slot in flight?
    const auto t_enq0 = std::chrono::steady_clock::now();
    double t_wait = 0.0;
    int slot = 0;
    // one iteration of the loop as a sequence of launches. The iteration number is kept on the device (dflags[3], see
    // convergence_body): the same sequence can then be replayed from a CUDA graph.
    auto iteration = [&]() -> int {
      if (!have_product)
        DE_TRY(spmm_device(ctx, A, Qa, Qb, m, false)); // Qb = A Qa (:78)
      DE_TRY(orthonormalize_device(ctx, n, m, Qb, have_gram ? ctx->dDG() + m : nullptr, A)); // (:81)
      // Qa = A Qb, dp = diag(Qb^T Qa) (:84-85) and the convergence test as the tail of the reduction of the dot-product
      // partials -- or, deferred, of the next iteration's first Gram reduction
      ctx->tail = de::TailArgs{};
      ctx->tail.kind = de::kTailConv;
      ctx->tail.m = m;
      ctx->tail.k = -1;
      ctx->tail.shift = shift;
      ctx->tail.tol = tol;
      ctx->tail.s_prev = s_prev;
      ctx->tail.hist = hist;
      ctx->tail.flags = ctx->dflags;
      ctx->tail_armed = true;
      ctx->tail_did_allreduce = ctx->tail_did_op = false;
      DE_TRY(spmm_device(ctx, A, Qb, Qa, m, true, kFuseGramIntoSpmm ? &have_gram : nullptr));
      if (ctx->pending_dot.valid)
        ; // deferred: reduced, all-reduced and tested in the tail of the next iteration's first Gram reduction
      else if (ctx->tail_did_op)
        ctx->tail_did_op = false;
      else
      {
        DE_REG(de::convergence_kernel), de::convergence_kernel<<<1, 64, 0, ctx->stream>>>(-1, m, shift, tol, ctx->dDP(), s_prev, hist, ctx->dflags);
        DE_LAUNCH_CHECK(ctx);
      }
      ctx->tail_armed = false;
      std::swap(Qa, Qb); // now Qa orthonormal, Qb = A*Qa
      have_product = true;
      return DE_OK;
    };
    auto poll = [&]() -> int {
      // look at the copy requested one batch ago (it has almost always landed), then request a new one
      const int prev = slot ^ 1;
      if (pending[prev])
      {
        const auto tw0 = std::chrono::steady_clock::now();
        DE_CUDA(ctx, cudaEventSynchronize(ctx->ev_poll[prev]));
        t_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count();
        pending[prev] = 0;
        if (ctx->hflags[4 * prev + 1] != 0)
          finished = true;
      }
      DE_CUDA(ctx, cudaMemcpyAsync(ctx->hflags + 4 * slot, ctx->dflags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      DE_CUDA(ctx, cudaEventRecord(ctx->ev_poll[slot], ctx->stream));
      pending[slot] = 1;
      slot ^= 1;
      return DE_OK;
    };
    // Steady state = kPollEvery iterations (an even number: the blocks swap roles every iteration) captured ONCE in a CUDA
    // graph and replayed: one launch call instead of ~28, and graph-internal launch latency between the dependent
    // kernels. Only on one GPU (the peer epochs of the multi-GPU path are baked into the launches) and without the
    // per-kernel timers. The first two iterations run as plain launches (the first one has an extra SpMM).
    static_assert(kPollEvery % 2 == 0, "a graph of the loop must leave the two blocks in their original roles");
    cudaGraphExec_t loop_graph = nullptr;
    long long graph_nodes = 0;
    struct GraphGuard
    {
      cudaGraphExec_t &g;
      ~GraphGuard()
      {
        if (g)
          cudaGraphExecDestroy(g);
      }
    } graph_guard{loop_graph};
    const bool want_graph = ctx->nranks <= 1 && !ctx->profiling && ctx->use_loop_graph && ctx->defer_dot;
    int k = 1;
    while (k < maxiter && !finished)
    {
      if (want_graph && loop_graph == nullptr && k == 3 && k + kPollEvery <= maxiter)
      {
        cudaGraph_t graph = nullptr;
        const long long before = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
        {
          int rc = DE_OK;
          for (int q = 0; q < kPollEvery && rc == DE_OK; ++q)
            rc = iteration();
          const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
          graph_nodes = ctx->launches - before;
          ctx->launches = before;
          if (rc != DE_OK || ce != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&loop_graph, graph, 0) != cudaSuccess)
            loop_graph = nullptr;
          if (graph)
            cudaGraphDestroy(graph);
          cudaGetLastError();
          if (rc != DE_OK)
            return rc;
          if (loop_graph == nullptr)
            return set_error(ctx, DE_ERR_CUDA, "StandardLargest: capturing the iteration graph failed");
        }
      }
      if (loop_graph != nullptr && k + kPollEvery <= maxiter)
      {
        DE_CUDA(ctx, cudaGraphLaunch(loop_graph, ctx->stream));
        ctx->launches += graph_nodes;
        k += kPollEvery;
        enqueued += kPollEvery;
        DE_TRY(poll());
        continue;
      }
      DE_TRY(iteration());
      ++k;
      ++enqueued;
      if (enqueued % kPollEvery == 0)
        DE_TRY(poll());
    }
    DE_TRY(flush_pending_dot(ctx)); // the last iteration's Rayleigh quotients have no following Gram reduction
    ctx->defer_dot = false;
    ctx->done_ptr = nullptr;
    const double t_enq = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_enq0).count();
    // final state
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->hflags, ctx->dflags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DE_TRY(fetch_small(ctx, s_prev, s2.data(), (size_t)m)); // synchronises; reports a failed Cholesky
    const int done = ctx->hflags[1];
    int k_exit = ctx->hflags[2];
    if (!done)
      k_exit = enqueued; // ran to maxiter - 1
    if (maxiter <= 1)
      k_exit = std::min(1, maxiter - 1);
    if ((enqueued - k_exit) % 2 != 0)
      std::swap(Qa, Qb); // iterations enqueued after convergence did nothing: undo their pointer swaps
    if (verbose > 1)
      std::printf("async loop: %d iterations enqueued in %.3f ms of host time (%.3f ms of it waiting for poll copies), "
                  "drained %.3f ms later\n",
                  enqueued, t_enq * 1e3, t_wait * 1e3,
                  std::chrono::duration<double>(std::chrono::steady_clock::now() - t_enq0).count() * 1e3 - t_enq * 1e3);
    if (verbose > 0 && k_exit > 1)
    {
      const int k_listed = std::min(k_exit, de::kConvHistory - 1);
      std::vector<double> h((size_t)k_listed + 1);
      DE_CUDA(ctx, cudaMemcpy(h.data(), hist, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (int k = 2; k <= k_listed; ++k)
        std::printf("Iter=%d %g\n", k, h[k]);
    }
    if (k_exit_out)
      *k_exit_out = k_exit;
    return DE_OK;
  }

  static int standard_core(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                           int maxiter, int m, double *&Qa, double *&Qb, std::vector<double> &s2, int verbose,
                           int *k_exit_out)
  {
    const long long n = A->n;
    if (!F && ts_supported(m) && maxiter > 1)
      return standard_core_async(ctx, A, shift, tol, maxiter, m, Qa, Qb, s2, verbose, k_exit_out);
    DE_TRY(reset_status(ctx));
    DE_TRY(orthonormalize_device(ctx, n, m, Qa)); // (:69, :159)
    std::vector<double> s1(m, 0.0);
    s2.assign(m, 0.0);
    int k_exit = std::min(1, maxiter - 1);
    bool have_product = false; // Qb == A*Qa already?
    for (int k = 1; k < maxiter; ++k)
    {
      k_exit = k;
      if (F)
        DE_TRY(factor_apply_device(ctx, F, Qa, Qb, m)); // Qb = A^-1 Qa (:168)
      else if (!have_product)
        DE_TRY(spmm_device(ctx, A, Qa, Qb, m, false)); // Qb = A Qa (:78)
      DE_TRY(orthonormalize_device(ctx, n, m, Qb));    // (:81, :171)
      DE_TRY(spmm_device(ctx, A, Qb, Qa, m, true));    // Qa = A Qb and s1 = diag(Qb^T Qa) (:84-85, :174-175)
      DE_TRY(fetch_small(ctx, ctx->dDP(), s1.data(), m));
      double distance = 0.0;
      for (int i = 0; i < m; ++i)
      {
        s1[i] -= shift;
        distance = std::max(distance, std::abs(s1[i] - s2[i]));
      }
      if (verbose > 0 && k > 1)
        std::printf("%s=%d %g\n", F ? "iter" : "Iter", k, distance);
      std::swap(s1, s2);
      std::swap(Qa, Qb); // now Qa orthonormal, Qb = A*Qa
      have_product = true;
      if (k > 1 && distance < tol) // absolute change of the Rayleigh quotients (:101-102, :188-189)
        break;
    }
    if (k_exit_out)
      *k_exit_out = k_exit;
    return DE_OK;
  }

  static int standard_driver(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                             int maxiter, int nev, const double *start_panel8, double *eval, double *evec, int verbose,
                             int *iterations)
  {
    if (!ctx || !A || !start_panel8 || !eval || !evec || nev <= 0)
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: bad arguments");
    const int m = padded_cols(nev);
    if (!valid_cols(m))
      return set_error(ctx, DE_ERR_UNSUPPORTED, "standard eigensolver driver: nev exceeds DE_MAX_COLS (64)");
    if (F && F->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: factorisation does not match the matrix");
    DE_TRY(bind_device(ctx));
    const long long n = A->n;
    ScopedBlocks blk;
    double *Qa, *Qb;
    DE_TRY(blk.alloc(ctx, &Qa, (size_t)n * m));
    DE_TRY(blk.alloc(ctx, &Qb, (size_t)n * m));
    SetupTrace trace("standard_driver", ctx->rank);
    DE_TRY(upload_panel8_device(ctx, n, m, start_panel8, Qa));
    trace.lap("start block uploaded");
    std::vector<double> s2;
    DE_TRY(standard_core(ctx, A, F, shift, tol, maxiter, m, Qa, Qb, s2, verbose, iterations));
    trace.lap("solve");
    const int rc = copy_out(ctx, n, m, nev, Qa, s2, eval, evec);
    trace.lap("eigenvectors copied out");
    return rc;
  }

  /** device-resident variant: Q holds the start block on entry and the eigenvector block on return */
  static int standard_driver_mv(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                                int maxiter, de_mv *Q, double *eval_m, int verbose, int *iterations)
  {
    if (!ctx || !A || !Q || !eval_m)
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: bad arguments");
    if (Q->n != A->n || (F && F->n != A->n))
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: block / factorisation do not match the matrix");
    DE_TRY(bind_device(ctx));
    double *Qa = Q->d, *Qb = nullptr;
    DE_TRY(dev_alloc(ctx, &Qb, (size_t)Q->n * Q->m));
    std::vector<double> s2;
    int s = standard_core(ctx, A, F, shift, tol, maxiter, Q->m, Qa, Qb, s2, verbose, iterations);
    Q->d = Qa; // the buffers may have swapped roles; Q keeps the one with the result
    dev_free(Qb);
    if (s != DE_OK)
      return s;
    for (int j = 0; j < Q->m; ++j)
      eval_m[j] = s2[j];
    return DE_OK;
  }

  int de_standard_largest_mv(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, de_mv *Q,
                             double *eval_m, int verbose, int *iterations)
  {
    return standard_driver_mv(ctx, A, nullptr, shift, tol, maxiter, Q, eval_m, verbose, iterations);
  }

  int de_standard_inverse_mv(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                             int maxiter, de_mv *Q, double *eval_m, int verbose, int *iterations)
  {
    if (!F)
      return set_error(ctx, DE_ERR_INVALID, "de_standard_inverse_mv: factorisation is null");
    return standard_driver_mv(ctx, A, F, shift, tol, maxiter, Q, eval_m, verbose, iterations);
  }

  int de_standard_largest(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, int nev,
                          const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    return standard_driver(ctx, A, nullptr, shift, tol, maxiter, nev, start_panel8, eval, evec, verbose, iterations);
  }

  int de_standard_inverse(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                          int maxiter, int nev, const double *start_panel8, double *eval, double *evec, int verbose,
                          int *iterations)
  {
    if (!F)
      return set_error(ctx, DE_ERR_INVALID, "de_standard_inverse: factorisation is null");
    return standard_driver(ctx, A, F, shift, tol, maxiter, nev, start_panel8, eval, evec, verbose, iterations);
  }

  int de_generalized_inverse(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *F, double shift,
                             double tol, int maxiter, int nev, const double *start_panel8, double *eval, double *evec,
                             int verbose, int *iterations, double *relerror_out)
  {
    if (!ctx || !A || !B || !F || !start_panel8 || !eval || !evec || nev <= 0)
      return set_error(ctx, DE_ERR_INVALID, "de_generalized_inverse: bad arguments");
    const int m = padded_cols(nev);
    if (!valid_cols(m))
      return set_error(ctx, DE_ERR_UNSUPPORTED, "de_generalized_inverse: nev exceeds DE_MAX_COLS (64)");
    if (A->n != B->n || F->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "de_generalized_inverse: A, B and the factorisation must have the same size");
    DE_TRY(bind_device(ctx));
    const long long n = A->n;
    ScopedBlocks blk;
    double *Q1, *Q2, *BQ;
    DE_TRY(blk.alloc(ctx, &Q1, (size_t)n * m));
    DE_TRY(blk.alloc(ctx, &Q2, (size_t)n * m));
    DE_TRY(blk.alloc(ctx, &BQ, (size_t)n * m));
    DE_TRY(reset_status(ctx));
    DE_TRY(upload_panel8_device(ctx, n, m, start_panel8, Q1));
    std::vector<double> ra1(m, 0.0), ra2(m, 0.0), sA(m, 0.0);
    DE_TRY(b_orthonormalize_device(ctx, B, n, m, Q1, BQ, false)); // (:270) ; BQ = B*Q1
    DE_TRY(spmm_device(ctx, A, Q1, Q2, m, true));                 // (:271-272)
    DE_TRY(fetch_small(ctx, ctx->dDP(), sA.data(), m));
    for (int i = 0; i < m; ++i)
      ra2[i] = sA[i] - shift;
    int iter = 0;
    double relerror = 0.0;
    while (iter < maxiter)
    {
      // Q2 = B*Q1 (:295) is the product BQ that B-orthonormalisation carried along; Q1 = A^-1 * (B*Q1) (:296)
      DE_TRY(factor_apply_device(ctx, F, BQ, Q1, m));
      DE_TRY(b_orthonormalize_device(ctx, B, n, m, Q1, BQ, false)); // (:297)
      iter += 1;
      DE_TRY(spmm_device(ctx, A, Q1, Q2, m, true)); // (:308-309)
      DE_TRY(fetch_small(ctx, ctx->dDP(), sA.data(), m));
      relerror = 0.0;
      for (int i = 0; i < m; ++i)
      {
        ra1[i] = sA[i] - shift;
        relerror = std::max(relerror, std::abs(ra1[i] - ra2[i]));
      }
      relerror /= *std::max_element(ra1.begin(), ra1.end());
      if (verbose > 2)
        std::printf("iter=%d relerror=%g\n", iter, relerror);
      std::swap(ra1, ra2);
      if (iter > 10 && relerror < tol) // (:323)
        break;
    }
    if (iterations)
      *iterations = iter;
    if (relerror_out)
      *relerror_out = relerror;
    return copy_out(ctx, n, m, nev, Q1, ra2, eval, evec);
  }


  // ---- LOBPCG drivers (new; no reference counterpart) ----------------------------------------
  int de_standard_lobpcg(de_context *ctx, const de_matrix *A, double tol, int maxiter, int nev,
                         const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    return lobpcg_driver(ctx, "StandardLOBPCG", A, nullptr, nullptr, tol, maxiter, nev, start_panel8, eval, evec, verbose,
                         iterations);
  }

  int de_generalized_lobpcg(de_context *ctx, const de_matrix *A, const de_matrix *B, double tol, int maxiter, int nev,
                            const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    if (!B)
      return set_error(ctx, DE_ERR_INVALID, "GeneralizedLOBPCG: bad arguments");
    return lobpcg_driver(ctx, "GeneralizedLOBPCG", A, B, nullptr, tol, maxiter, nev, start_panel8, eval, evec, verbose,
                         iterations);
  }

  int de_lobpcg_mv(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *T, int largest,
                   int cheb_degree, double tol, int maxiter, int nev, de_mv *Q, double *eval_m, double *resnorm_m, int verbose, int *iterations,
                   int *restarts, int *converged)
  {
    if (!ctx || !A || !Q || !eval_m)
      return set_error(ctx, DE_ERR_INVALID, "de_lobpcg_mv: bad arguments");
    if (Q->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "de_lobpcg_mv: the block does not match the matrix");
    DE_TRY(lobpcg_check_args(ctx, "de_lobpcg_mv", A, B, T, nev, Q->m));
    DE_TRY(bind_device(ctx));
    de::LobpcgResult res;
    DE_TRY(lobpcg_device(ctx, A, B, T, largest != 0, cheb_degree, tol, maxiter, nev, Q->m, Q->d, res, verbose));
    for (int j = 0; j < Q->m; ++j)
    {
      eval_m[j] = res.theta[j];
      if (resnorm_m)
        resnorm_m[j] = res.resnorm[j];
    }
    if (iterations)
      *iterations = res.iterations;
    if (restarts)
      *restarts = res.restarts;
    if (converged)
      *converged = res.converged ? 1 : 0;
    return DE_OK;
  }

  int de_block_lincomb(de_mv *out, de_mv *out2, int ns, const de_mv *const *S, const double *C_host)
  {
    if (!out || !S || !C_host || ns < 1 || ns > 3)
      return set_error(nullptr, DE_ERR_INVALID, "de_block_lincomb: bad arguments");
    de_context *ctx = out->ctx;
    for (int s = 0; s < ns; ++s)
      if (!S[s] || S[s]->n != out->n || S[s]->m != out->m)
        return set_error(ctx, DE_ERR_INVALID, "de_block_lincomb: blocks must have the same shape");
    if (out2 && (out2->n != out->n || out2->m != out->m || out2->d == S[0]->d || out2->d == out->d))
      return set_error(ctx, DE_ERR_INVALID, "de_block_lincomb: out2 must have the same shape and alias neither out nor S[0]");
    for (int s = 1; s < ns; ++s)
      if (out->d == S[s]->d)
        return set_error(ctx, DE_ERR_INVALID, "de_block_lincomb: out may alias S[0] only");
    DE_TRY(bind_device(ctx));
    const int m = out->m;
    ScopedBlocks tmp;
    double *dC = nullptr;
    DE_TRY(tmp.alloc(ctx, &dC, (size_t)3 * m * m));
    DE_CUDA(ctx, cudaMemcpyAsync(dC, C_host, sizeof(double) * (size_t)ns * m * m, cudaMemcpyHostToDevice, ctx->stream));
    const double *src[3] = {S[0]->d, ns > 1 ? S[1]->d : nullptr, ns > 2 ? S[2]->d : nullptr};
    DE_TRY(lincomb_device(ctx, m, out->n, ns, src, dC, out->d, (out2 && ns > 1) ? out2->d : nullptr));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // dC returns to the allocator, C_host to the caller
    return DE_OK;
  }

  int de_host_sym_eig(int n, const double *A, double *w, double *V)
  {
    if (n < 0 || (n > 0 && (!A || !w || !V)))
      return set_error(nullptr, DE_ERR_INVALID, "de_host_sym_eig: bad arguments");
    return de::hosteig::sym_eig(n, A, w, V) == 0 ? DE_OK
                                                  : set_error(nullptr, DE_ERR_SINGULAR, "de_host_sym_eig: QL iteration failed");
  }

  int de_host_sym_gen_eig(int n, const double *GA, const double *GB, double *w, double *C, double *min_pivot)
  {
    if (n < 0 || (n > 0 && (!GA || !GB || !w || !C)))
      return set_error(nullptr, DE_ERR_INVALID, "de_host_sym_gen_eig: bad arguments");
    const int rc = de::hosteig::sym_gen_eig(n, GA, GB, w, C, 0.0, min_pivot);
    return rc == 0 ? DE_OK : set_error(nullptr, DE_ERR_SINGULAR, "de_host_sym_gen_eig: GB is not positive definite");
  }

} // extern "C"
