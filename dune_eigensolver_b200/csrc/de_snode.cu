// de_snode.cu -- supernodal factors on the device: upload of a supernodal Cholesky factor (host provider:
// include/dune/eigensolver/supernodal_cholesky.hh), its panel / level schedule, and the factored apply on it
// (kernels_snode.cuh). The second form of de_factor next to the level-scheduled scalar factors of de_trsv.cu; the drivers
// and de_factor_apply do not care which one they hold.
#include "de_internal.hpp"
#include "../../include/dune/eigensolver/supernodal_cholesky.hh"
#include "kernels_snode.cuh"
#include "kernels_trsv.cuh" // permute_rows_kernel

using namespace dei;

struct de_sn_device
{
  de::SnPanel *panels = nullptr;
  int2 *items = nullptr;
  double *val = nullptr, *dinv = nullptr, *ones = nullptr;
  int *rowidx = nullptr, *ticket = nullptr, *perm = nullptr;
  std::vector<long long> level_ptr; // items of level l: [level_ptr[l], level_ptr[l + 1])
  long long npanels = 0, nitems = 0, stored = 0;
};

namespace dei
{
  void sn_release(de_sn_device *S)
  {
    if (!S)
      return;
    dev_free(S->panels);
    dev_free(S->items);
    dev_free(S->val);
    dev_free(S->dinv);
    dev_free(S->ones);
    dev_free(S->rowidx);
    dev_free(S->ticket);
    dev_free(S->perm);
    delete S;
  }

  static int sn_sweeps(de_context *ctx, const de_factor *F, double *W, double *Z, int m)
  {
    const de_sn_device *S = F->sn;
    de::SnArgs a{S->panels, S->items, S->val, S->rowidx, S->dinv, W, Z, S->ticket, m};
    const int nlev = (int)S->level_ptr.size() - 1;
    DE_TRY(ensure_func_smem(ctx, (const void *)de::sn_backward_kernel, de::kSnSmem));
    DE_TRY(ensure_func_smem(ctx, (const void *)de::sn_forward_kernel, de::kSnSmem));
    for (int l = 0; l < nlev; ++l)
    {
      const long long i0 = S->level_ptr[l], cnt = S->level_ptr[l + 1] - i0;
      if (cnt <= 0)
        continue;
      ProfScope prof(ctx, DE_PROF_TRSV);
      DE_REG(de::sn_forward_kernel), de::sn_forward_kernel<<<(unsigned)cnt, de::kSnThreads, de::kSnSmem, ctx->stream>>>(a, (int)i0);
      DE_LAUNCH_CHECK(ctx);
    }
    for (int l = nlev - 1; l >= 0; --l)
    {
      const long long i0 = S->level_ptr[l], cnt = S->level_ptr[l + 1] - i0;
      if (cnt <= 0)
        continue;
      ProfScope prof(ctx, DE_PROF_TRSV);
      DE_REG(de::sn_backward_kernel), de::sn_backward_kernel<<<(unsigned)cnt, de::kSnThreads, de::kSnSmem, ctx->stream>>>(a, (int)i0);
      DE_LAUNCH_CHECK(ctx);
    }
    return DE_OK;
  }

  /** Y = A^-1 X through the supernodal factor: W = X(perm) -> forward -> backward -> Y(perm) = Z */
  int sn_apply_device(de_context *ctx, de_factor *F, const double *X, double *Y, int m)
  {
    de_sn_device *S = F->sn;
    // work blocks: W (right-hand sides, updated in place) and Z (solutions)
    if (F->W_m < m)
    {
      dev_free(F->W);
      dev_free(F->W2);
      F->W = F->W2 = nullptr;
      if (F->sweep_graph)
      {
        cudaGraphExecDestroy(F->sweep_graph);
        F->sweep_graph = nullptr;
        F->sweep_graph_m = 0;
      }
      DE_TRY(dev_alloc(ctx, &F->W, (size_t)F->n * m));
      DE_TRY(dev_alloc(ctx, &F->W2, (size_t)F->n * m));
      F->W_m = m;
    }
    const long long total = F->n * (m / 2);
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->sm_count * 8));
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      DE_REG(de::permute_rows_kernel), de::permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(F->n, m, S->perm, S->ones, X, F->W, 0);
    }
    DE_LAUNCH_CHECK(ctx);
    bool replayed = false;
    if (!ctx->profiling)
    {
      // the level launches are a fixed sequence on fixed blocks: captured once per width, replayed with one call
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
          const int s1 = sn_sweeps(ctx, F, F->W, F->W2, m);
          const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
          F->sweep_graph_nodes = ctx->launches - before;
          ctx->launches = before;
          if (s1 == DE_OK && ce == cudaSuccess && graph != nullptr && cudaGraphInstantiate(&F->sweep_graph, graph, 0) == cudaSuccess)
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
      DE_TRY(sn_sweeps(ctx, F, F->W, F->W2, m));
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      DE_REG(de::permute_rows_kernel), de::permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(F->n, m, S->perm, nullptr, F->W2, Y, 1);
    }
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  /** device form of a supernodal factor: panels, inverted diagonal blocks, level schedule */
  int sn_upload(de_context *ctx, const de_b200::SupernodalFactor &H, de_factor *F)
  {
    using I = long;
    const I nsuper = H.nsuper, n = H.n;
    if ((long long)H.val.size() >= (1LL << 40))
      return set_error(ctx, DE_ERR_UNSUPPORTED, "supernodal factor too large");
    de_sn_device *S = new de_sn_device();
    F->sn = S;
    F->n = n;
    F->lnz = F->unz = H.lnz;
    // panels + inverted diagonal blocks + levels
    std::vector<de::SnPanel> panels;
    std::vector<int> plevel;
    std::vector<double> dinv;
    std::vector<int> last_level(nsuper, 0); // forward level of the last panel of supernode s
    std::vector<int> child_max(nsuper, 0);
    int nlev = 0;
    for (I s = 0; s < nsuper; ++s)
    {
      const I ns = H.cols(s), r = H.rows(s);
      const double *L = H.val.data() + H.valptr[s];
      int lev = child_max[s]; // first panel: one level above the last panels of all children
      for (I j0 = 0; j0 < ns; j0 += de::kSnPanel)
      {
        const int w = (int)std::min<I>(de::kSnPanel, ns - j0);
        de::SnPanel P;
        P.lofs = H.valptr[s] + j0 * r; // column j0; row index a is local to the supernode
        P.rofs = H.rowptr[s];
        P.dofs = (long long)dinv.size();
        P.r = (int)r;
        P.ns = (int)ns;
        P.j0 = (int)j0;
        P.w = w;
        P.c0 = (int)(H.sfirst[s] + j0);
        const I below = r - (j0 + w);
        P.ntiles = (int)std::max<I>(1, (below + de::kSnTile - 1) / de::kSnTile);
        // inverse of the w x w lower triangular diagonal block, row-major
        const size_t d0 = dinv.size();
        dinv.resize(d0 + (size_t)w * w, 0.0);
        double *D = dinv.data() + d0;
        for (int c = 0; c < w; ++c)
        {
          // column c of the inverse: solve L x = e_c
          for (int i = c; i < w; ++i)
          {
            double sum = (i == c) ? 1.0 : 0.0;
            for (int k = c; k < i; ++k)
              sum -= L[(j0 + i) + (j0 + k) * r] * D[k * w + c];
            D[i * w + c] = sum / L[(j0 + i) + (j0 + i) * r];
          }
        }
        panels.push_back(P);
        plevel.push_back(lev);
        nlev = std::max(nlev, lev + 1);
        ++lev;
      }
      last_level[s] = lev; // = level of the last panel + 1
      if (H.sparent[s] != -1)
        child_max[H.sparent[s]] = std::max(child_max[H.sparent[s]], lev);
    }
    // items sorted by level
    S->npanels = (long long)panels.size();
    S->level_ptr.assign(nlev + 1, 0);
    for (size_t p = 0; p < panels.size(); ++p)
      S->level_ptr[plevel[p] + 1] += panels[p].ntiles;
    for (int l = 0; l < nlev; ++l)
      S->level_ptr[l + 1] += S->level_ptr[l];
    S->nitems = S->level_ptr[nlev];
    if (S->nitems >= (1LL << 31) || S->npanels >= (1LL << 31))
      return set_error(ctx, DE_ERR_UNSUPPORTED, "supernodal factor: too many panels");
    std::vector<int2> items((size_t)S->nitems);
    {
      std::vector<long long> w(S->level_ptr.begin(), S->level_ptr.end() - 1);
      for (size_t p = 0; p < panels.size(); ++p)
        for (int t = 0; t < panels[p].ntiles; ++t)
          items[(size_t)w[plevel[p]]++] = make_int2((int)p, t);
    }
    S->stored = (long long)H.val.size();
    DE_TRY(dev_alloc(ctx, &S->panels, panels.size()));
    DE_TRY(dev_alloc(ctx, &S->items, items.size()));
    DE_TRY(dev_alloc(ctx, &S->val, H.val.size()));
    DE_TRY(dev_alloc(ctx, &S->dinv, dinv.size()));
    DE_TRY(dev_alloc(ctx, &S->rowidx, H.rowidx.size()));
    DE_TRY(dev_alloc(ctx, &S->ticket, panels.size()));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(S->panels, panels.data(), panels.size() * sizeof(de::SnPanel), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(S->items, items.data(), items.size() * sizeof(int2), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(S->rowidx, H.rowidx.data(), H.rowidx.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaMemsetAsync(S->ticket, 0, panels.size() * sizeof(int), ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    DE_TRY(upload_parallel(ctx, S->val, H.val.data(), H.val.size()));
    DE_TRY(upload_parallel(ctx, S->dinv, dinv.data(), dinv.size()));
    std::vector<double> ones((size_t)n, 1.0);
    DE_TRY(upload_converted(ctx, &S->ones, ones.data(), ones.size()));
    DE_TRY(upload_converted(ctx, &S->perm, H.perm.data(), H.perm.size()));
    return DE_OK;
  }
  int sn_expand_contract(de_host_factor *H)
  {
    if (!H->sn)
      return DE_OK;
    if (H->sn->lnz >= (1L << 31))
      return set_error(nullptr, DE_ERR_UNSUPPORTED,
                       "de_host_factor_arrays: this supernodal factor is too large for the explicit L / U arrays of the "
                       "UMFPACK contract; upload it with de_factor_upload_host");
    de_b200::supernodal_to_contract(*H->sn, H->F);
    return DE_OK;
  }
} // namespace dei

extern "C"
{

  int de_host_factorize_spd(int64_t n, const int64_t *rowptr, const int64_t *col, const double *val, int ordering, int nthreads,
                            de_host_factor **out)
  {
    if (!out || n < 0 || !rowptr)
      return set_error(nullptr, DE_ERR_INVALID, "de_host_factorize_spd: bad arguments");
    *out = nullptr;
    de_host_factor *F = new de_host_factor();
    F->sn.reset(new de_b200::SupernodalFactor());
    try
    {
      de_b200::supernodal_cholesky((long)n, rowptr, col, val, (de_b200::Ordering)ordering, nthreads, *F->sn);
    }
    catch (const std::exception &e)
    {
      delete F;
      const std::string msg = e.what();
      return set_error(nullptr, msg.find("positive definite") != std::string::npos ? DE_ERR_SINGULAR : DE_ERR_INVALID, msg);
    }
    *out = F;
    return DE_OK;
  }

  int de_host_factor_info(const de_host_factor *F, int *supernodal, int64_t *n, int64_t *lnz, int64_t *stored, double *flops,
                          double *seconds3)
  {
    if (!F)
      return set_error(nullptr, DE_ERR_INVALID, "null host factor");
    const bool sn = F->sn != nullptr;
    if (supernodal)
      *supernodal = sn ? 1 : 0;
    if (n)
      *n = sn ? F->sn->n : F->F.n;
    if (lnz)
      *lnz = sn ? F->sn->lnz : F->F.lnz;
    if (stored)
      *stored = sn ? (int64_t)F->sn->val.size() : F->F.lnz + F->F.unz;
    if (flops)
      *flops = sn ? F->sn->flops : 0.0;
    if (seconds3)
    {
      seconds3[0] = sn ? F->sn->seconds_ordering : 0.0;
      seconds3[1] = sn ? F->sn->seconds_symbolic : 0.0;
      seconds3[2] = sn ? F->sn->seconds_numeric : 0.0;
    }
    return DE_OK;
  }

  int de_factor_upload_host(de_context *ctx, const de_host_factor *H, de_factor **out)
  {
    if (!ctx || !H || !out)
      return set_error(ctx, DE_ERR_INVALID, "de_factor_upload_host: bad arguments");
    *out = nullptr;
    if (!H->sn)
    {
      const de_b200::FactorArrays &A = H->F;
      return de_factor_upload(ctx, A.n, A.Lp.data(), A.Lj.data(), A.Lx.data(), A.Up.data(), A.Ui.data(), A.Ux.data(), A.P.data(),
                              A.Q.data(), A.Rs.data(), A.do_recip, out);
    }
    DE_TRY(bind_device(ctx));
    de_factor *F = new de_factor();
    F->ctx = ctx;
    context_retain(ctx);
    const int s = sn_upload(ctx, *H->sn, F);
    if (s != DE_OK)
    {
      de_factor_destroy(F);
      return s;
    }
    *out = F;
    return DE_OK;
  }

} // extern "C"
