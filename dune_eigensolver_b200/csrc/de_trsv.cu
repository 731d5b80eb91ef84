// de_trsv.cu -- factored inverse apply: the UMFPACK field contract (reference umfpacktools.hh:26-44) re-ordered into
// level schedules at upload, and Y = Q U^-1 L^-1 P R X for all m columns at once
// (replaces matmul_inverse_tallskinny_{blocked,avx2_b8,neon_b8}, reference kernels_cpp.hh:660-755).
#include "de_internal.hpp"
#include "kernels_sparse.cuh"
#include "kernels_trsv.cuh"

using namespace dei;

namespace dei
{
  // ---- triangular solves --------------------------------------------------------------------------------
  constexpr int kChainMaxRows = 32;

  /** rows sorted by level for a strictly-triangular CSR whose dependencies point to already-solved rows */
  int build_schedule(de_context *ctx, long long n, const std::vector<int> &ptr, const std::vector<int> &col,
                     const std::vector<double> &val, const std::vector<double> *invdiag, bool lower, TrsvSchedule &S)
  {
    std::vector<int> level(n, 0);
    int nlev = 0;
    if (lower)
      for (long long i = 0; i < n; ++i)
      {
        int l = 0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k)
          l = std::max(l, level[col[k]] + 1);
        level[i] = l;
        nlev = std::max(nlev, l + 1);
      }
    else
      for (long long i = n - 1; i >= 0; --i)
      {
        int l = 0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k)
          l = std::max(l, level[col[k]] + 1);
        level[i] = l;
        nlev = std::max(nlev, l + 1);
      }
    if (n == 0)
      nlev = 0;
    S.nlevels = nlev;
    S.h_level_ptr.assign(nlev + 1, 0);
    for (long long i = 0; i < n; ++i)
      S.h_level_ptr[level[i] + 1]++;
    for (int l = 0; l < nlev; ++l)
      S.h_level_ptr[l + 1] += S.h_level_ptr[l];
    std::vector<int> rows(n), fill(S.h_level_ptr.begin(), S.h_level_ptr.end() - (nlev >= 0 ? 1 : 0));
    for (long long i = 0; i < n; ++i)
      rows[fill[level[i]]++] = (int)i;
    // segments: runs of narrow levels are chained in one CTA, wide levels get their own launch
    S.segments.clear();
    for (int l = 0; l < nlev;)
    {
      const int width = S.h_level_ptr[l + 1] - S.h_level_ptr[l];
      if (width <= kChainMaxRows)
      {
        int e = l + 1;
        while (e < nlev && S.h_level_ptr[e + 1] - S.h_level_ptr[e] <= kChainMaxRows)
          ++e;
        S.segments.push_back(TrsvSegment{1, l, e});
        l = e;
      }
      else
      {
        S.segments.push_back(TrsvSegment{0, l, l + 1});
        ++l;
      }
    }
    S.nnz = (long long)col.size();
    DE_TRY(upload_converted(ctx, &S.rows, rows.data(), rows.size()));
    DE_TRY(upload_converted(ctx, &S.rowptr, ptr.data(), ptr.size()));
    DE_TRY(upload_converted(ctx, &S.col, col.data(), col.size()));
    DE_TRY(upload_converted(ctx, &S.val, val.data(), val.size()));
    DE_TRY(upload_converted(ctx, &S.level_ptr, S.h_level_ptr.data(), S.h_level_ptr.size()));
    if (invdiag)
      DE_TRY(upload_converted(ctx, &S.invdiag, invdiag->data(), invdiag->size()));
    return DE_OK;
  }

  void free_schedule(TrsvSchedule &S)
  {
    dev_free(S.rows);
    dev_free(S.rowptr);
    dev_free(S.col);
    dev_free(S.val);
    dev_free(S.level_ptr);
    dev_free(S.invdiag);
  }

  template <int LC>
  int run_schedule_t(de_context *ctx, const TrsvSchedule &S, double *W, int m)
  {
    de::TrsvArgs a{S.rows, S.rowptr, S.col, S.val, S.invdiag, W, m};
    for (const TrsvSegment &seg : S.segments)
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      if (seg.chain)
        DE_REG(de::trsv_chain_kernel<LC>), de::trsv_chain_kernel<LC><<<1, 1024, 0, ctx->stream>>>(a, S.level_ptr, seg.a, seg.b);
      else
      {
        const int first = S.h_level_ptr[seg.a], count = S.h_level_ptr[seg.a + 1] - first;
        DE_REG(de::trsv_level_kernel<LC>), de::trsv_level_kernel<LC><<<(count + 7) / 8, 256, 0, ctx->stream>>>(a, first, count);
      }
      DE_LAUNCH_CHECK(ctx);
    }
    return DE_OK;
  }

  int run_schedule(de_context *ctx, const TrsvSchedule &S, double *W, int m)
  {
    const int hp = m / 2;
    if (hp <= 4)
      return run_schedule_t<4>(ctx, S, W, m);
    if (hp <= 8)
      return run_schedule_t<8>(ctx, S, W, m);
    if (hp <= 16)
      return run_schedule_t<16>(ctx, S, W, m);
    return run_schedule_t<32>(ctx, S, W, m);
  }

  int ensure_factor_work(de_context *ctx, de_factor *F, int m)
  {
    if (F->W_m >= m)
      return DE_OK;
    if (F->W)
      dev_free(F->W);
    F->W = nullptr;
    if (F->sweep_graph) // captured on the old work block
    {
      cudaGraphExecDestroy(F->sweep_graph);
      F->sweep_graph = nullptr;
      F->sweep_graph_m = 0;
    }
    DE_TRY(dev_alloc(ctx, &F->W, (size_t)F->n * m));
    F->W_m = m;
    return DE_OK;
  }

  /** Y = (factored A)^-1 X (reference matmul_inverse_tallskinny_blocked, kernels_cpp.hh:660-755) */
  int factor_apply_device(de_context *ctx, const de_factor *Fc, const double *X, double *Y, int m)
  {
    de_factor *F = const_cast<de_factor *>(Fc);
    if (ctx->nranks > 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "factored apply is single-GPU (triangular solves do not row-shard)");
    if (F->sn)
      return sn_apply_device(ctx, F, X, Y, m);
    DE_TRY(ensure_factor_work(ctx, F, m));
    const long long total = F->n * (m / 2);
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->sm_count * 8));
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      DE_REG(de::permute_rows_kernel), de::permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(F->n, m, F->P, F->rowscale, X, F->W, 0);
    }
    DE_LAUNCH_CHECK(ctx);
    // forward sweep over the levels of L, backward sweep over the levels of U: a fixed sequence of launches on the
    // fixed block W -> a CUDA graph, replayed with one call (per-launch CPU cost and front-end latency dominate these
    // sweeps for 2D problems: hundreds of levels of a few hundred rows). With per-kernel timers on, launch one by one.
    bool replayed = false;
    if (!ctx->profiling)
    {
      if (F->sweep_graph == nullptr || F->sweep_graph_m != m)
      {
        if (F->sweep_graph)
          cudaGraphExecDestroy(F->sweep_graph);
        F->sweep_graph = nullptr;
        F->sweep_graph_m = 0;
        cudaGraph_t graph = nullptr;
        const long long before = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
        {
          const int s1 = run_schedule(ctx, F->L, F->W, m);
          const int s2 = (s1 == DE_OK) ? run_schedule(ctx, F->U, F->W, m) : s1;
          const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
          F->sweep_graph_nodes = ctx->launches - before;
          ctx->launches = before; // counted again at every replay
          if (s2 == DE_OK && ce == cudaSuccess && graph != nullptr &&
              cudaGraphInstantiate(&F->sweep_graph, graph, 0) == cudaSuccess)
            F->sweep_graph_m = m;
          else
            F->sweep_graph = nullptr;
          if (graph)
            cudaGraphDestroy(graph);
          cudaGetLastError();
        }
      }
      if (F->sweep_graph != nullptr && F->sweep_graph_m == m)
      {
        DE_CUDA(ctx, cudaGraphLaunch(F->sweep_graph, ctx->stream));
        ctx->launches += F->sweep_graph_nodes;
        replayed = true;
      }
    }
    if (!replayed)
    {
      DE_TRY(run_schedule(ctx, F->L, F->W, m));
      DE_TRY(run_schedule(ctx, F->U, F->W, m));
    }
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      DE_REG(de::permute_rows_kernel), de::permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(F->n, m, F->Q, nullptr, F->W, Y, 1);
    }
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

} // namespace dei

extern "C"
{

  // ---- factored apply -----------------------------------------------------------------------------------------
  int de_factor_upload(de_context *ctx, int64_t n, const long *Lp, const long *Lj, const double *Lx, const long *Up,
                       const long *Ui, const double *Ux, const long *P, const long *Q, const double *Rs, long do_recip,
                       de_factor **out)
  {
    if (!ctx || !out || n < 0 || !Lp || !Up || !P || !Q || !Rs)
      return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: bad arguments");
    *out = nullptr;
    DE_TRY(bind_device(ctx));
    const long lnz = Lp[n], unz = Up[n];
    if (lnz >= (1L << 31) || unz >= (1L << 31) || n >= (1L << 31) - 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "factor too large for 32-bit indices");
    // L: CSR with the (unit) diagonal stored last in each row -> strip it (kernels_cpp.hh:717 skips it the same way)
    std::vector<int> lptr(n + 1, 0), lcol;
    std::vector<double> lval;
    lcol.reserve(lnz > n ? lnz - n : 0);
    lval.reserve(lnz > n ? lnz - n : 0);
    for (int64_t i = 0; i < n; ++i)
    {
      if (Lp[i + 1] - Lp[i] < 1 || Lj[Lp[i + 1] - 1] != i)
        return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: L rows must end with their diagonal entry");
      for (long k = Lp[i]; k < Lp[i + 1] - 1; ++k)
      {
        if (Lj[k] < 0 || Lj[k] >= i)
          return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: L is not strictly lower triangular");
        lcol.push_back((int)Lj[k]);
        lval.push_back(Lx[k]);
      }
      lptr[i + 1] = (int)lcol.size();
    }
    // U: CSC with the diagonal last in each column -> CSR of the strictly upper part + inverse diagonal
    std::vector<int> uptr(n + 1, 0);
    std::vector<double> invd(n, 0.0);
    for (int64_t j = 0; j < n; ++j)
    {
      if (Up[j + 1] - Up[j] < 1 || Ui[Up[j + 1] - 1] != j)
        return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: U columns must end with their diagonal entry");
      const double d = Ux[Up[j + 1] - 1];
      if (d == 0.0 || !std::isfinite(d))
        return set_error(ctx, DE_ERR_SINGULAR, "UMFPackFactorizedMatrix: input matrix is singular"); // umfpacktools.hh:163
      invd[j] = 1.0 / d;
      for (long k = Up[j]; k < Up[j + 1] - 1; ++k)
      {
        if (Ui[k] < 0 || Ui[k] >= j)
          return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: U is not strictly upper triangular");
        uptr[Ui[k] + 1]++;
      }
    }
    for (int64_t i = 0; i < n; ++i)
      uptr[i + 1] += uptr[i];
    std::vector<int> ucol(uptr[n]);
    std::vector<double> uval(uptr[n]);
    {
      std::vector<int> w(uptr.begin(), uptr.end() - 1);
      for (int64_t j = 0; j < n; ++j) // ascending j => ascending columns inside every row
        for (long k = Up[j]; k < Up[j + 1] - 1; ++k)
        {
          const int dst = w[Ui[k]]++;
          ucol[dst] = (int)j;
          uval[dst] = Ux[k];
        }
    }
    std::vector<double> rowscale(n);
    std::vector<char> seenp(n, 0), seenq(n, 0);
    for (int64_t k = 0; k < n; ++k)
    {
      if (P[k] < 0 || P[k] >= n || Q[k] < 0 || Q[k] >= n || seenp[P[k]] || seenq[Q[k]])
        return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: P / Q are not permutations");
      seenp[P[k]] = seenq[Q[k]] = 1;
      rowscale[k] = do_recip ? Rs[P[k]] : 1.0 / Rs[P[k]]; // kernels_cpp.hh:687, :699
    }
    de_factor *F = new de_factor();
    F->ctx = ctx;
    context_retain(ctx);
    F->n = n;
    F->lnz = lnz;
    F->unz = unz;
    auto fail = [&](int s) {
      de_factor_destroy(F);
      return s;
    };
    int s;
    if ((s = build_schedule(ctx, n, lptr, lcol, lval, nullptr, true, F->L)) != DE_OK)
      return fail(s);
    if ((s = build_schedule(ctx, n, uptr, ucol, uval, &invd, false, F->U)) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &F->P, P, (size_t)n)) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &F->Q, Q, (size_t)n)) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &F->rowscale, rowscale.data(), (size_t)n)) != DE_OK)
      return fail(s);
    *out = F;
    return DE_OK;
  }

  int de_factor_destroy(de_factor *F)
  {
    if (!F)
      return DE_OK;
    de_context *ctx = F->ctx;
    cudaSetDevice(ctx->device);
    free_schedule(F->L);
    free_schedule(F->U);
    if (F->sweep_graph)
      cudaGraphExecDestroy(F->sweep_graph);
    dev_free(F->P);
    dev_free(F->Q);
    dev_free(F->rowscale);
    dev_free(F->W);
    dev_free(F->W2);
    sn_release(F->sn);
    delete F;
    context_release(ctx);
    return DE_OK;
  }

  int de_factor_apply(de_mv *Y, const de_factor *F, de_mv *X)
  {
    if (!Y || !F || !X)
      return set_error(nullptr, DE_ERR_INVALID, "de_factor_apply: null argument");
    de_context *ctx = F->ctx;
    if (Y->n != X->n || Y->m != X->m)
      return set_error(ctx, DE_ERR_INVALID, "matmul_inverse_tallskinny_blocked: Qout/Qin size mismatch"); // kernels_cpp.hh:665
    if (F->n != X->n)
      return set_error(ctx, DE_ERR_INVALID,
                       "matmul_inverse_tallskinny_blocked: Factorization does not match size of Qout/Qin"); // :667
    DE_TRY(bind_device(ctx));
    return factor_apply_device(ctx, F, X->d, Y->d, X->m);
  }

  int de_factor_info(const de_factor *F, int64_t *n, int64_t *lnz, int64_t *unz, int *levels_L, int *levels_U)
  {
    if (!F)
      return set_error(nullptr, DE_ERR_INVALID, "null factor");
    if (n)
      *n = F->n;
    if (lnz)
      *lnz = F->lnz;
    if (unz)
      *unz = F->unz;
    if (levels_L)
      *levels_L = F->L.nlevels;
    if (levels_U)
      *levels_U = F->U.nlevels;
    return DE_OK;
  }

} // extern "C"
