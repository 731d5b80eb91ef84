// de_spmm.cu -- sparse matrices on the device (CSR + BRB forms, row-partitioned parts with their halo plans) and the
// SpMM launch logic: replaces matmul_sparse_tallskinny_{naive,blocked,avx2_b8,neon_b8} (reference kernels_cpp.hh:596-657).
#include "de_internal.hpp"
#include "kernels_sparse.cuh"
#include "kernels_spmm_blocked.cuh"
#include "brb_format.hpp"
#include "kernels_brb_build.cuh"
#include "kernels_peer.cuh"

using namespace dei;

namespace dei
{
  // ---- SpMM -----------------------------------------------------------------------------------------
  template <bool DOT>
  int launch_spmm_rows(de_context *ctx, const de_matrix *A, const double *X, double *Y, int m, const int *rowlist,
                       long long nrows, double *partials, int *grid_out)
  {
    *grid_out = 0;
    if (nrows <= 0)
      return DE_OK;
    de::SpmmArgs a;
    a.nrows = nrows;
    a.rowlist = rowlist;
    a.rowptr = A->rowptr;
    a.col = A->col;
    a.val = A->val;
    a.X = X;
    a.H = A->halo_view;
    a.n_owned = A->n;
    a.ld = m;
    a.m = m;
    a.Y = Y;
    a.partials = partials;
    a.done = ctx->done_ptr;
    const int hp = m / 2;
    const int tpr = hp <= 4 ? 4 : (hp <= 8 ? 8 : (hp <= 16 ? 16 : 32));
    const int rpb = 256 / tpr;
    const long long need = (nrows + rpb - 1) / rpb;
    const int cap = DOT ? kMaxPartials : ctx->sm_count * 8;
    int grid = (int)std::min<long long>(need, cap);
    ProfScope prof(ctx, DE_PROF_SPMM);
    const bool exact = (m == 2 * tpr) && ((A->n + A->n_halo) * (long long)(m / 2) < (1LL << 31));
    const bool halo = A->n_halo > 0;
    if (exact)
    {
      // spmm_kernel_v2 is compiled for 3 resident CTAs per SM: launch exactly one wave of the grid-stride loop
      grid = (int)std::min<long long>(need, (long long)ctx->sm_count * 3);
#define DE_SPMM_V2(T)                                                                                        \
  if (halo)                                                                                                  \
    DE_REG(de::spmm_kernel_v2<T, DOT, true>), de::spmm_kernel_v2<T, DOT, true><<<grid, 256, 0, ctx->stream>>>(a);                                      \
  else                                                                                                       \
    DE_REG(de::spmm_kernel_v2<T, DOT, false>), de::spmm_kernel_v2<T, DOT, false><<<grid, 256, 0, ctx->stream>>>(a);
      switch (tpr)
      {
      case 4:
        DE_SPMM_V2(4)
        break;
      case 8:
        DE_SPMM_V2(8)
        break;
      case 16:
        DE_SPMM_V2(16)
        break;
      default:
        DE_SPMM_V2(32)
        break;
      }
#undef DE_SPMM_V2
    }
    else
      switch (tpr)
      {
      case 4:
        DE_REG(de::spmm_kernel<4, 1, DOT>), de::spmm_kernel<4, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      case 8:
        DE_REG(de::spmm_kernel<8, 1, DOT>), de::spmm_kernel<8, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      case 16:
        DE_REG(de::spmm_kernel<16, 1, DOT>), de::spmm_kernel<16, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      default:
        DE_REG(de::spmm_kernel<32, 1, DOT>), de::spmm_kernel<32, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      }
    DE_LAUNCH_CHECK(ctx);
    *grid_out = grid;
    return DE_OK;
  }

  /** cut rows 0..nrows of a (host) row pointer into blocks of <= 256 rows and <= 2048 nonzeros */
  template <class Ptr>
  std::vector<int4> cut_row_blocks(long long nrows, const Ptr *rowptr)
  {
    std::vector<int4> meta;
    long long r0 = 0;
    while (r0 < nrows)
    {
      long long r1 = r0;
      while (r1 < nrows && r1 - r0 < de::kStageMaxRows && (long long)(rowptr[r1 + 1] - rowptr[r0]) <= de::kStageCapNnz)
        ++r1;
      if (r1 == r0)
        r1 = r0 + 1; // a single row longer than the staging capacity: read directly from global memory
      meta.push_back(make_int4((int)r0, (int)r1, (int)rowptr[r0], (int)rowptr[r1]));
      r0 = r1;
    }
    return meta;
  }

  /** staged view of the whole matrix (shares the CSR arrays already on the device) */
  int build_staged_all(de_context *ctx, de_matrix *A, const int64_t *rowptr)
  {
    std::vector<int4> meta = cut_row_blocks(A->n, rowptr);
    StagedRows &S = A->st_all;
    S.rowptr = A->rowptr;
    S.col = A->col;
    S.val = A->val;
    S.owns_csr = false;
    S.nblocks = (int)meta.size();
    DE_TRY(dev_alloc(ctx, &S.blk_meta, meta.size()));
    DE_CUDA(ctx, cudaMemcpyAsync(S.blk_meta, meta.data(), meta.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S.valid = true;
    return DE_OK;
  }

  /** staged copy of the rows in `list` (interior or boundary rows of a distributed matrix), permuted to be contiguous */
  int build_staged_subset(de_context *ctx, const std::vector<int> &list, const int64_t *rowptr, const int64_t *col,
                          const double *val, StagedRows &S)
  {
    S.valid = false;
    if (list.empty())
      return DE_OK;
    std::vector<int> ptr(list.size() + 1, 0), c;
    std::vector<double> v;
    for (size_t i = 0; i < list.size(); ++i)
    {
      const int r = list[i];
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
      {
        c.push_back((int)col[k]);
        v.push_back(val[k]);
      }
      ptr[i + 1] = (int)c.size();
    }
    std::vector<int4> meta = cut_row_blocks((long long)list.size(), ptr.data());
    S.owns_csr = true;
    S.nblocks = (int)meta.size();
    DE_TRY(upload_converted(ctx, &S.rowptr, ptr.data(), ptr.size()));
    DE_TRY(upload_converted(ctx, &S.col, c.data(), c.size()));
    DE_TRY(upload_converted(ctx, &S.val, v.data(), v.size()));
    DE_TRY(upload_converted(ctx, &S.rowmap, list.data(), list.size()));
    DE_TRY(dev_alloc(ctx, &S.blk_meta, meta.size()));
    DE_CUDA(ctx, cudaMemcpyAsync(S.blk_meta, meta.data(), meta.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S.valid = true;
    return DE_OK;
  }

  inline bool staged_usable(const de_matrix *A, int m)
  {
    return (m == 8 || m == 16 || m == 32 || m == 64) && (A->n + A->n_halo) * (long long)(m / 2) < (1LL << 31);
  }

  template <bool DOT>
  int launch_spmm_staged(de_context *ctx, const de_matrix *A, const StagedRows &S, const double *X, double *Y, int m,
                         double *partials, int *grid_out)
  {
    *grid_out = 0;
    if (!S.valid || S.nblocks <= 0)
      return DE_OK;
    de::StagedArgs a;
    a.nblocks = S.nblocks;
    a.blk_meta = S.blk_meta;
    a.rowmap = S.rowmap;
    a.rowptr = S.rowptr;
    a.col = S.col;
    a.val = S.val;
    a.X = X;
    a.H = A->halo_view;
    a.n_owned = (int)A->n;
    a.m = m;
    a.Y = Y;
    a.partials = partials;
    a.done = ctx->done_ptr;
    constexpr size_t smem = de::spmm_staged_smem_bytes();
    const int grid = std::min(S.nblocks, ctx->sm_count * 3); // 3 resident CTAs per SM
    const bool halo = A->n_halo > 0;
    const int tpr = m / 2;
    ProfScope prof(ctx, DE_PROF_SPMM);
#define DE_SPMM_ST(T)                                                                                        \
  {                                                                                                          \
    DE_TRY(ensure_func_smem(ctx, (const void *)de::spmm_staged_kernel<T, DOT, true>, smem));                 \
    DE_TRY(ensure_func_smem(ctx, (const void *)de::spmm_staged_kernel<T, DOT, false>, smem));                \
    if (halo)                                                                                                \
      DE_REG(de::spmm_staged_kernel<T, DOT, true>), de::spmm_staged_kernel<T, DOT, true><<<grid, 256, smem, ctx->stream>>>(a);                             \
    else                                                                                                     \
      DE_REG(de::spmm_staged_kernel<T, DOT, false>), de::spmm_staged_kernel<T, DOT, false><<<grid, 256, smem, ctx->stream>>>(a);                            \
  }
    switch (tpr)
    {
    case 4:
      DE_SPMM_ST(4)
      break;
    case 8:
      DE_SPMM_ST(8)
      break;
    case 16:
      DE_SPMM_ST(16)
      break;
    default:
      DE_SPMM_ST(32)
      break;
    }
#undef DE_SPMM_ST
    DE_LAUNCH_CHECK(ctx);
    *grid_out = grid;
    return DE_OK;
  }

  /** BRB form of the matrix. The host only plans which rows form the tiles (brb::plan: pattern detection on a few
   *  sample rows + O(n) index arithmetic); the blobs are built on the device from the CSR arrays already uploaded
   *  (kernels_brb_build.cuh: COUNT pass -> offsets on the host -> FILL pass). A matrix the format cannot represent
   *  (or an empty one) simply keeps the CSR kernels: not an error. */
  int build_brb(de_context *ctx, de_matrix *A, long long n, long long ncols, const int64_t *rowptr, const int64_t *col,
                const double *val)
  {
    BrbDevice &B = A->brb;
    B.release();
    if (n <= 0 || rowptr[n] <= 0)
      return DE_OK;
    static_assert(sizeof(de::brb::TileDesc) == sizeof(int4), "tile descriptors are loaded as int4");
    de::brb::Plan P;
    int first = 0;
    int *d_rows = nullptr, *d_cut = nullptr;
    int4 *d_info = nullptr, *d_place = nullptr;
    auto cleanup = [&]()
    {
      dev_free(d_rows);
      dev_free(d_cut);
      dev_free(d_info);
      dev_free(d_place);
      d_rows = d_cut = nullptr;
      d_info = d_place = nullptr;
    };
    struct Scope
    {
      decltype(cleanup) &f;
      ~Scope() { f(); }
    } scope{cleanup};
    while (de::brb::plan(n, ncols, rowptr, col, val, n, first, P))
    {
      first = P.next;
      cleanup();
      const int ntiles = (int)P.order.tilecut.size() - 1;
      if (ntiles <= 0)
        continue;
      DE_TRY(dev_alloc(ctx, &d_rows, P.order.rows.size()));
      DE_TRY(dev_alloc(ctx, &d_cut, P.order.tilecut.size()));
      DE_TRY(dev_alloc(ctx, &d_info, (size_t)ntiles));
      DE_TRY(dev_alloc(ctx, &d_place, (size_t)ntiles));
      DE_CUDA(ctx, cudaMemcpyAsync(d_rows, P.order.rows.data(), P.order.rows.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      DE_CUDA(ctx, cudaMemcpyAsync(d_cut, P.order.tilecut.data(), P.order.tilecut.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      de::BrbBuildArgs a{};
      a.ntiles = ntiles;
      a.tilecut = d_cut;
      a.rows = d_rows;
      a.rowptr = A->rowptr;
      a.col = A->col;
      a.val = A->val;
      a.n_owned = (int)n;
      a.info = d_info;
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        DE_REG(de::brb_build_kernel<false>), de::brb_build_kernel<false><<<ntiles, de::kBldThreads, 0, ctx->stream>>>(a);
      }
      DE_LAUNCH_CHECK(ctx);
      std::vector<int4> info((size_t)ntiles);
      DE_CUDA(ctx, cudaMemcpyAsync(info.data(), d_info, info.size() * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
      DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      // sizes -> placement (tile order of the plan); the SpMM kernel sees the tiles interior-first
      std::vector<int4> place((size_t)ntiles);
      size_t blob16 = 0, ucols = 0;
      int max_len16 = 0, max_u = 0;
      long long nsteps = 0, nvals = 0;
      bool ok = true;
      for (int t = 0; t < ntiles && ok; ++t)
      {
        const int nb = P.order.tilecut[t + 1] - P.order.tilecut[t];
        const int4 f = info[t];
        if (f.w & 2)
        {
          ok = false;
          break;
        }
        const size_t words = (((size_t)4 + (nb + 1) + 12 * (size_t)nb + 3) & ~(size_t)3) + 4 * (size_t)f.x + (((size_t)2 * f.y + 3) & ~(size_t)3);
        const int len16 = (int)(words / 4);
        place[t] = make_int4((int)blob16, len16, (int)ucols, f.z);
        blob16 += (size_t)len16;
        ucols += (size_t)f.z;
        max_len16 = std::max(max_len16, len16);
        max_u = std::max(max_u, f.z);
        nsteps += f.x;
        nvals += f.y;
        if (blob16 >= ((size_t)1 << 31) || ucols >= ((size_t)1 << 31))
          ok = false;
      }
      if (!ok || !de::brb::fits_budget(max_len16, max_u))
        continue; // next candidate (smaller tiles / consecutive rows)
      DE_TRY(dev_alloc(ctx, &B.blob, blob16 + 4));
      DE_TRY(dev_alloc(ctx, &B.ucol, ucols + 64));
      DE_TRY(dev_alloc(ctx, &B.tile, (size_t)ntiles));
      DE_CUDA(ctx, cudaMemsetAsync(B.blob, 0, (blob16 + 4) * sizeof(int4), ctx->stream));
      DE_CUDA(ctx, cudaMemcpyAsync(d_place, place.data(), place.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
      a.place = d_place;
      a.blob = reinterpret_cast<int *>(B.blob);
      a.ucol = B.ucol;
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        DE_REG(de::brb_build_kernel<true>), de::brb_build_kernel<true><<<ntiles, de::kBldThreads, 0, ctx->stream>>>(a);
      }
      DE_LAUNCH_CHECK(ctx);
      std::vector<int4> tiles;
      tiles.reserve((size_t)ntiles);
      int n_interior = 0;
      for (int pass = 0; pass < 2; ++pass)
      {
        for (int t = 0; t < ntiles; ++t)
          if (((info[t].w & 1) != 0) == (pass == 1))
            tiles.push_back(place[t]);
        if (pass == 0)
          n_interior = (int)tiles.size();
      }
      DE_CUDA(ctx, cudaMemcpyAsync(B.tile, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
      DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host vectors die here
      B.ntiles = ntiles;
      B.n_interior = n_interior;
      B.max_len16 = max_len16;
      B.max_u = max_u;
      B.nblocks = (long long)(P.order.rows.size() / 8);
      B.nsteps = nsteps;
      B.nvals = nvals;
      B.grid = P.grid;
      B.tw = P.tw;
      B.th = P.th;
      B.td = P.td;
      B.blob16 = blob16;
      B.nucol = ucols;
      B.valid = true;
      return DE_OK;
    }
    return DE_OK;
  }

  /** AUTO policy (measured, profiles/r01_spmm_lab_*.log): the BRB kernel wins for long rows at every width (27-point:
   *  1.5-2.4x) and for short rows from m = 32 up; narrow blocks on a 7-point matrix leave the tensor-core steps mostly
   *  zero-filled and the CSR kernel is faster there. */
  inline bool brb_usable(const de_matrix *A, int m)
  {
    if (!A->brb.valid || A->spmm_format == DE_SPMM_CSR || m % 8 != 0 || m < 8 || m > DE_MAX_COLS)
      return false;
    if (A->spmm_format == DE_SPMM_BRB)
      return true;
    return m >= 32 || A->nnz >= 12 * A->n;
  }

  constexpr int kBrbMaxSmem = 227 * 1024;

  template <int NP, bool DOT, bool HALO, bool GRAM, int EPI = 0>
  int launch_brb_pass(de_context *ctx, const de::BrbArgs &a, int grid)
  {
    DE_TRY(ensure_func_smem(ctx, (const void *)de::spmm_brb_kernel<NP, DOT, HALO, GRAM, EPI>, kBrbMaxSmem));
    const size_t smem = de::spmm_brb_smem_bytes(NP, a.blob_cap16, a.xs_cap, a.stages);
    DE_CUDA(ctx, launch_pdl(ctx->pdl, DE_KERNEL(de::spmm_brb_kernel<NP, DOT, HALO, GRAM, EPI>), dim3(grid), dim3(de::brb_threads(GRAM)), smem, ctx->stream, a));
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  /** The drivers could take G = Y^T Y from the SpMM epilogue instead of a separate Gram pass (de_spmm_gram does).
   *  Measured on B200 (100^3 Q1, m = 32): the epilogue adds 0.12 ms to a 0.19 ms SpMM (20 more DMMA per row block
   *  on accumulators that serialise, 40 more registers, 10 instead of 12 consumer warps), the separate pass costs
   *  0.056 ms + one launch: the solve went from 22.4 to 25.9 ms. Off until the epilogue is cheaper. */
  constexpr bool kFuseGramIntoSpmm = false;

  /** can the Gram matrix Y^T Y be accumulated in the SpMM epilogue? (single pass over the columns) */
  inline bool brb_gram_epilogue(int m) { return m == 8 || m == 16 || m == 32; }

  /** Y = A X on tiles [t0, t0 + nt) of the BRB form. m columns are covered by passes of 32 / 16 / 8 columns (one
   *  kernel launch each; every pass re-streams the matrix blobs, so m = 64 costs two passes).
   *  DOT: per-CTA partials of diag(X^T Y) at partials[cta * pstride + j]; GRAM (needs brb_gram_epilogue(m)):
   *  additionally the CTA's partial of Y^T Y at partials[cta * pstride + m + i * m + j], pstride = m + m * m. */
  template <bool DOT, bool GRAM>
  int launch_spmm_brb(de_context *ctx, const de_matrix *A, int t0, int nt, const double *X, double *Y, int m, double *partials,
                      int *grid_out, int category = DE_PROF_SPMM)
  {
    *grid_out = 0;
    if (nt <= 0)
      return DE_OK;
    const BrbDevice &B = A->brb;
    const int grid = std::min(nt, ctx->sm_count);
    const bool halo = A->n_halo > 0;
    ProfScope prof(ctx, category);
    for (int c0 = 0; c0 < m;)
    {
      const int left = (m - c0) / 8;
      const int np = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
      de::BrbArgs a;
      a.ntiles = nt;
      a.n = A->n;
      a.tile = B.tile + t0;
      a.blob = B.blob;
      a.ucol = B.ucol;
      a.X = X + c0;
      a.H = A->halo_view ? A->halo_view + c0 : nullptr;
      a.n_owned = (int)A->n;
      a.ldx = m;
      a.Y = Y + c0;
      a.partials = partials ? partials + c0 : nullptr;
      a.pstride = GRAM ? m + m * m : m;
      a.gram_off = m;
      a.blob_cap16 = B.max_len16;
      a.xs_cap = B.max_u;
      a.done = ctx->done_ptr;
      a.E0 = nullptr;
      a.E1 = nullptr;
      a.edinv = nullptr;
      a.ealpha = a.ebeta = 0.0;
      bool epi = false;
      if constexpr (!DOT && !GRAM)
      {
        if (ctx->epi.valid) // spmm_cheb_device: the launch updates the Chebyshev iterate instead of storing Y
        {
          epi = true;
          a.E0 = ctx->epi.zold + c0;
          a.E1 = ctx->epi.r + c0;
          a.edinv = ctx->epi.dinv;
          a.ealpha = ctx->epi.alpha;
          a.ebeta = ctx->epi.beta;
        }
      }
      int stages = de::kBrbMaxStages;
      while (stages > 2 && de::spmm_brb_smem_bytes(np, a.blob_cap16, a.xs_cap, stages) > (size_t)kBrbMaxSmem)
        --stages;
      a.stages = stages;
      if (de::spmm_brb_smem_bytes(np, a.blob_cap16, a.xs_cap, stages) > (size_t)kBrbMaxSmem)
        return set_error(ctx, DE_ERR_UNSUPPORTED, "BRB tile does not fit in shared memory");
#define DE_BRB(NPV)                                                             \
  {                                                                             \
    if constexpr (!DOT && !GRAM)                                                \
    {                                                                           \
      if (epi && halo)                                                          \
        DE_TRY((launch_brb_pass<NPV, false, true, false, 1>(ctx, a, grid)));    \
      else if (epi)                                                             \
        DE_TRY((launch_brb_pass<NPV, false, false, false, 1>(ctx, a, grid)));   \
    }                                                                           \
    if (epi)                                                                    \
      ;                                                                         \
    else if (halo)                                                              \
      DE_TRY((launch_brb_pass<NPV, DOT, true, GRAM>(ctx, a, grid)));            \
    else                                                                        \
      DE_TRY((launch_brb_pass<NPV, DOT, false, GRAM>(ctx, a, grid)));           \
  }
      if (np == 4)
        DE_BRB(4)
      else if (np == 2)
        DE_BRB(2)
      else
        DE_BRB(1)
#undef DE_BRB
      c0 += 8 * np;
    }
    *grid_out = grid;
    return DE_OK;
  }

  int ensure_halo_buffers(de_context *ctx, de_matrix *A, int m)
  {
    if (A->buf_m >= m)
      return DE_OK;
    if (A->send_buf)
      dev_free(A->send_buf);
    if (A->halo_buf)
      dev_free(A->halo_buf);
    A->send_buf = A->halo_buf = nullptr;
    DE_TRY(dev_alloc(ctx, &A->send_buf, (size_t)A->n_send * m));
    DE_TRY(dev_alloc(ctx, &A->halo_buf, (size_t)A->n_halo * m));
    A->buf_m = m;
    return DE_OK;
  }

  /** Y = A X (+ dp = diag(X^T Y) into ctx->dDP when DOT). Distributed matrices first start the halo exchange
   *  (pack -> NCCL send/recv over NVLink on the communication stream), run the interior rows meanwhile, then the
   *  boundary rows once the halo rows have landed. */
  template <bool DOT>
  int spmm_device_t(de_context *ctx, const de_matrix *Ac, const double *X, double *Y, int m, bool *gram_out)
  {
    de_matrix *A = const_cast<de_matrix *>(Ac);
    int g1 = 0, g2 = 0;
    const bool dist = ctx->nranks > 1 && (A->n_halo > 0 || A->n_send > 0);
    // peer-store halo exchange? decided from quantities that are equal on all ranks; the epoch advances on every rank
    // of the job, also on one that has neither halo rows nor rows to send for this matrix
    const bool peer_path = ctx->nranks > 1 && ctx->peer_ready && A->peer_halo &&
                           (size_t)A->halo_rows_max * m * sizeof(double) <= ctx->halo_cap && A->npeers <= de::kPeerMaxRanks;
    const unsigned long long halo_epoch = peer_path ? ++ctx->halo_epoch : 0ull;
    struct ClearPrepushed
    {
      de_context *c;
      ~ClearPrepushed() { c->prepushed_X = nullptr; }
    } clear_prepushed{ctx};
    // GRAM epilogue: the caller can use G = Y^T Y (in ctx->dDG() + m); only the tensor-core kernel has it
    const bool gram = DOT && gram_out != nullptr && brb_usable(A, m) && brb_gram_epilogue(m);
    if (gram_out)
      *gram_out = gram;
    if (!dist)
    {
      if (brb_usable(A, m))
      {
        if (gram)
          DE_TRY((launch_spmm_brb<DOT, DOT>(ctx, A, 0, A->brb.ntiles, X, Y, m, ctx->partials, &g1)));
        else
          DE_TRY((launch_spmm_brb<DOT, false>(ctx, A, 0, A->brb.ntiles, X, Y, m, ctx->partials, &g1)));
      }
      else if (A->st_all.valid && staged_usable(A, m))
        DE_TRY(launch_spmm_staged<DOT>(ctx, A, A->st_all, X, Y, m, ctx->partials, &g1));
      else
        DE_TRY(launch_spmm_rows<DOT>(ctx, A, X, Y, m, nullptr, A->n, ctx->partials, &g1));
    }
    else
    {
      const bool peer = peer_path;
      de::PeerArgs pa{};
      if (peer)
      {
        // halo rows go straight into the neighbours' windows; the boundary tiles wait for this epoch's flags
        pa = peer_args(ctx, halo_epoch);
        A->halo_view = reinterpret_cast<double *>(ctx->window + de::kPeerHaloOff + (size_t)(pa.epoch & 1ull) * ctx->halo_cap);
        if (A->n_send > 0)
        {
          de::HaloPushArgs h{};
          h.npeers = A->npeers;
          for (int p = 0; p < A->npeers; ++p)
          {
            h.peer_rank[p] = A->peer[p];
            h.send_off[p] = A->send_off[p];
            h.deposit[p] = A->deposit[p];
          }
          h.send_off[A->npeers] = A->n_send;
          h.send_rows = A->send_rows;
          h.X = X;
          h.m = m;
          h.halo_cap_bytes = ctx->halo_cap;
          h.ticket = ctx->dticket;
          // rows already stored by the block-update kernels of the orthonormalisation that produced X (plan_fused_push)?
          const bool prepushed = ctx->prepushed_X == X && ctx->prepushed_A == A && ctx->prepushed_epoch == halo_epoch &&
                                 ctx->prepushed_m == m;
          if (prepushed)
            for (int p = 0; p <= A->npeers; ++p)
              h.send_off[p] = 0; // nothing to copy: the launch only releases the flags
          const long long total = prepushed ? 0 : A->n_send * (m / 2);
          const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->sm_count * 4));
          if (!(prepushed && ctx->prepushed_released)) // else: the last update launch has raised the flags already
          {
            ProfScope prof(ctx, DE_PROF_HALO_PUSH);
            DE_REG(de::halo_push_kernel), de::halo_push_kernel<<<grid, 256, 0, ctx->stream>>>(pa, h);
            DE_LAUNCH_CHECK(ctx);
          }
        }
      }
      else
      {
        if (!ctx->comm)
          return set_error(ctx, DE_ERR_UNSUPPORTED,
                           "SpMM: the halo block does not fit the peer window (or the matrix has no peer deposits) and the "
                           "context has no NCCL communicator; create the window with a larger halo_bytes");
        DE_TRY(ensure_halo_buffers(ctx, A, m));
        A->halo_view = A->halo_buf;
        if (A->n_send > 0)
        {
          const long long total = A->n_send * (m / 2);
          const int grid = (int)std::min<long long>((total + 255) / 256, ctx->sm_count * 8);
          ProfScope prof(ctx, DE_PROF_MISC);
          DE_REG(de::pack_rows_kernel), de::pack_rows_kernel<<<grid, 256, 0, ctx->stream>>>(A->n_send, A->send_rows, m, X, A->send_buf);
          DE_LAUNCH_CHECK(ctx);
        }
        DE_CUDA(ctx, cudaEventRecord(ctx->ev_pack, ctx->stream));
        DE_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_pack, 0));
        NcclApi &nc = nccl_api();
        DE_NCCL(ctx, nc.GroupStart());
        for (int p = 0; p < A->npeers; ++p)
        {
          if (A->send_count[p] > 0)
            DE_NCCL(ctx, nc.Send(A->send_buf + (size_t)A->send_off[p] * m, (size_t)A->send_count[p] * m, ncclDouble,
                                 A->peer[p], ctx->comm, ctx->comm_stream));
          if (A->recv_count[p] > 0)
            DE_NCCL(ctx, nc.Recv(A->halo_buf + (size_t)A->recv_off[p] * m, (size_t)A->recv_count[p] * m, ncclDouble,
                                 A->peer[p], ctx->comm, ctx->comm_stream));
        }
        DE_NCCL(ctx, nc.GroupEnd());
        DE_CUDA(ctx, cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
      }
      const bool staged = staged_usable(A, m) && (A->st_interior.valid || A->n_interior == 0) &&
                          (A->st_boundary.valid || A->n_boundary == 0);
      const bool brb = brb_usable(A, m);
      const size_t pstride = gram ? (size_t)m + (size_t)m * m : (size_t)m;
      if (brb)
      {
        if (gram)
          DE_TRY((launch_spmm_brb<DOT, DOT>(ctx, A, 0, A->brb.n_interior, X, Y, m, ctx->partials, &g1)));
        else
          DE_TRY((launch_spmm_brb<DOT, false>(ctx, A, 0, A->brb.n_interior, X, Y, m, ctx->partials, &g1)));
      }
      else if (staged)
        DE_TRY(launch_spmm_staged<DOT>(ctx, A, A->st_interior, X, Y, m, ctx->partials, &g1));
      else
        DE_TRY(launch_spmm_rows<DOT>(ctx, A, X, Y, m, A->interior, A->n_interior, ctx->partials, &g1));
      if (peer)
      {
        de::PeerList pl{};
        for (int p = 0; p < A->npeers; ++p)
          if (A->recv_count[p] > 0)
            pl.rank[pl.n++] = A->peer[p];
        ProfScope prof(ctx, DE_PROF_HALO_WAIT);
        DE_REG(de::halo_wait_kernel), de::halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(pa, pl);
        DE_LAUNCH_CHECK(ctx);
      }
      else
        DE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
      if (brb)
      {
        if (gram)
          DE_TRY((launch_spmm_brb<DOT, DOT>(ctx, A, A->brb.n_interior, A->brb.ntiles - A->brb.n_interior, X, Y, m,
                                            ctx->partials + (size_t)g1 * pstride, &g2, DE_PROF_SPMM_BOUNDARY)));
        else
          DE_TRY((launch_spmm_brb<DOT, false>(ctx, A, A->brb.n_interior, A->brb.ntiles - A->brb.n_interior, X, Y, m,
                                              ctx->partials + (size_t)g1 * pstride, &g2, DE_PROF_SPMM_BOUNDARY)));
      }
      else if (staged)
        DE_TRY(launch_spmm_staged<DOT>(ctx, A, A->st_boundary, X, Y, m, ctx->partials + (size_t)g1 * m, &g2));
      else
        DE_TRY(launch_spmm_rows<DOT>(ctx, A, X, Y, m, A->boundary, A->n_boundary, ctx->partials + (size_t)g1 * m, &g2));
    }
    if (DOT)
    {
      if (ctx->defer_dot && !gram && g1 + g2 > 0 && ctx->tail_armed && ctx->tail.kind == de::kTailConv)
      {
        // asynchronous driver loop: the partials stay where they are; the tail of the next Gram reduction reduces and
        // all-reduces them together with the Gram matrix and runs the convergence test (reduce_partials, de_dense.cu)
        ctx->pending_dot.valid = true;
        ctx->pending_dot.nparts = g1 + g2;
        ctx->pending_dot.m = m;
        ctx->pending_dot.conv = ctx->tail;
        ctx->tail_armed = false;
        return DE_OK;
      }
      // dp (and G = Y^T Y when the Gram epilogue ran) are reduced, and all-reduced, as ONE vector dDG = [dp | G]
      const int len = gram ? m + m * m : m;
      if (g1 + g2 > 0)
        DE_TRY(reduce_partials(ctx, ctx->partials, g1 + g2, len, ctx->dDG()));
      else
        DE_CUDA(ctx, cudaMemsetAsync(ctx->dDG(), 0, sizeof(double) * len, ctx->stream));
      DE_TRY(allreduce_sum(ctx, ctx->dDG(), (size_t)len));
    }
    return DE_OK;
  }

  /** Can the block-update kernels of an orthonormalisation store the halo rows of the NEXT SpMM with A (width m)
   *  straight into the neighbours' windows? Needs the peer path, at most two neighbours and consecutive send rows (what a
   *  slab partition gives). Fills ctx->push_pending for halo epoch ctx->halo_epoch + 1. The decision depends on this
   *  rank's lists only: a rank that pushes early and one that pushes in its SpMM call interoperate (the receiver only
   *  looks at the flag). */
  bool plan_fused_push(de_context *ctx, const de_matrix *A, int m)
  {
    ctx->push_pending.n = 0;
    if (!ctx->fused_push || !A || ctx->nranks <= 1 || !ctx->peer_ready || !A->peer_halo || A->n_send <= 0 ||
        (size_t)A->halo_rows_max * m * sizeof(double) > ctx->halo_cap || A->npeers > de::kPeerMaxRanks)
      return false;
    int cnt = 0;
    for (int p = 0; p < A->npeers; ++p)
      if (A->send_count[p] > 0)
      {
        if (cnt == 2 || (size_t)p >= A->send_first.size() || A->send_first[p] < 0)
          return false;
        ++cnt;
      }
    const unsigned long long epoch = ctx->halo_epoch + 1;
    de::PushRanges &r = ctx->push_pending;
    for (int p = 0; p < A->npeers; ++p)
      if (A->send_count[p] > 0)
      {
        r.lo[r.n] = A->send_first[p];
        r.hi[r.n] = A->send_first[p] + A->send_count[p];
        r.dst[r.n] = reinterpret_cast<double *>(ctx->peer_base[A->peer[p]] + de::kPeerHaloOff + (size_t)(epoch & 1ull) * ctx->halo_cap) +
                     (size_t)A->deposit[p] * m;
        // = peer_halo_flag(base of the peer, parity, this rank), kernels_peer.cuh
        r.flag[r.n] = de::peer_halo_flag(ctx->peer_base[A->peer[p]], (int)(epoch & 1ull), ctx->rank);
        ++r.n;
      }
    r.release = 0;
    r.first_row = 0;
    for (int q = 0; q < r.n; ++q)
      if (r.lo[q] > r.first_row && r.hi[q] >= A->n - 8) // a range at the end of the slab: the sweep starts there
        r.first_row = r.lo[q];
    r.epoch = epoch;
    r.ticket = ctx->dticket;
    // the flags of ALL send peers must be raised by whoever releases: with a peer that receives no rows from this rank
    // (send_count 0) but is listed, halo_push_kernel raises its flag too -- nobody waits for it, so it can be left out
    return r.n > 0;
  }

  /** Zold <- Z + alpha (Z - Zold) + beta D^-1 (R - A Z) in ONE pass when A has its tensor-core (BRB) form for this width:
   *  the Chebyshev update is the epilogue of the SpMM kernel (BrbArgs::E0 ...). *fused says whether that happened; if not the
   *  caller runs the SpMM into a scratch block and cheb_step_kernel. */
  int spmm_cheb_device(de_context *ctx, const de_matrix *A, const double *Z, double *Zold, const double *R, const double *dinv,
                       double alpha, double beta, int m, bool *fused)
  {
    *fused = false;
    if (!ctx->use_cheb_epilogue || !brb_usable(A, m))
      return DE_OK;
    ctx->epi.valid = true;
    ctx->epi.zold = Zold;
    ctx->epi.r = R;
    ctx->epi.dinv = dinv;
    ctx->epi.alpha = alpha;
    ctx->epi.beta = beta;
    const int rc = spmm_device_t<false>(ctx, A, Z, Zold /* never written as Y */, m, nullptr);
    ctx->epi.valid = false;
    *fused = rc == DE_OK;
    return rc;
  }

  void brb_set_plane_points(long long points) { de::brb::plane_points_setting() = points; }

  int spmm_device(de_context *ctx, const de_matrix *A, const double *X, double *Y, int m, bool dot, bool *gram_out)
  {
    return dot ? spmm_device_t<true>(ctx, A, X, Y, m, gram_out) : spmm_device_t<false>(ctx, A, X, Y, m, nullptr);
  }

} // namespace dei

extern "C"
{

  // ---- matrices ---------------------------------------------------------------------------------------
  static int matrix_upload(de_context *ctx, long long n, long long ncols, long long nnz, const int64_t *rowptr,
                           const int64_t *col, const double *val, de_matrix *A)
  {
    if (nnz >= (1LL << 31) || n >= (1LL << 31) - 1 || ncols >= (1LL << 31) - 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "matrix too large for 32-bit indices on one GPU");
    if (rowptr[0] != 0 || rowptr[n] != nnz)
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create: rowptr does not match nnz");
    A->n = n;
    A->nnz = nnz;
    // 16 bytes of tail padding: the staged kernels copy 16-byte chunks
    DE_TRY(dev_alloc(ctx, &A->rowptr, (size_t)n + 1 + 4));
    DE_TRY(dev_alloc(ctx, &A->col, (size_t)nnz + 4));
    DE_TRY(dev_alloc(ctx, &A->val, (size_t)nnz + 2));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // recycled blocks: their previous users are done
    long long range[2];
    DE_TRY(upload_parallel(ctx, A->rowptr, rowptr, (size_t)n + 1));
    DE_TRY(upload_parallel(ctx, A->col, col, (size_t)nnz, range));
    if (nnz > 0 && (range[0] < 0 || range[1] >= ncols))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create: column index out of range");
    DE_TRY(upload_parallel(ctx, A->val, val, (size_t)nnz));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_matrix_create_csr(de_context *ctx, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                           const double *val, de_matrix **out)
  {
    if (!ctx || !out || n < 0 || nnz < 0 || !rowptr || (nnz > 0 && (!col || !val)))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_csr: bad arguments");
    *out = nullptr;
    DE_TRY(bind_device(ctx));
    de_matrix *A = new de_matrix();
    A->ctx = ctx;
    context_retain(ctx);
    SetupTrace trace("create_csr", ctx->rank);
    int s = matrix_upload(ctx, n, n, nnz, rowptr, col, val, A);
    trace.lap("matrix upload");
    if (s == DE_OK)
      s = build_staged_all(ctx, A, rowptr);
    trace.lap("staged row blocks");
    if (s == DE_OK)
      s = build_brb(ctx, A, n, n, rowptr, col, val);
    trace.lap("BRB form built");
    if (s != DE_OK)
    {
      de_matrix_destroy(A);
      return s;
    }
    *out = A;
    return DE_OK;
  }

  int de_matrix_create_bcsr(de_context *ctx, int64_t nb, int64_t nnzb, int k, const int64_t *rowptr, const int64_t *col,
                            const double *val, de_matrix **out)
  {
    if (!ctx || !out || nb < 0 || nnzb < 0 || k < 1 || !rowptr || (nnzb > 0 && (!col || !val)))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_bcsr: bad arguments");
    if (k == 1)
      return de_matrix_create_csr(ctx, nb, nnzb, rowptr, col, val, out);
    if (rowptr[0] != 0 || rowptr[nb] != nnzb)
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_bcsr: rowptr does not match nnzb");
    // the scalar matrix the blocks denote; one-time host work like the BCRS -> CSR flattening of the 1 x 1 case
    const int64_t n = nb * k, nnz = nnzb * k * k;
    std::vector<int64_t> rp((size_t)n + 1), ci((size_t)nnz);
    std::vector<double> v((size_t)nnz);
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(8, nnz >> 20));
    auto work = [&](int t)
    {
      const int64_t b0 = nb * t / nth, b1 = nb * (t + 1) / nth;
      for (int64_t ib = b0; ib < b1; ++ib)
      {
        const int64_t e0 = rowptr[ib], e1 = rowptr[ib + 1], len = (e1 - e0) * k;
        for (int r = 0; r < k; ++r)
        {
          const int64_t row = ib * k + r, base = e0 * k * k + (int64_t)r * len;
          rp[row] = base;
          for (int64_t e = e0; e < e1; ++e)
            for (int c = 0; c < k; ++c)
            {
              ci[base + (e - e0) * k + c] = col[e] * k + c;
              v[base + (e - e0) * k + c] = val[(e * k + r) * k + c];
            }
        }
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nth; ++t)
      th.emplace_back(work, t);
    work(0);
    for (auto &x : th)
      x.join();
    rp[n] = nnz;
    return de_matrix_create_csr(ctx, n, nnz, rp.data(), ci.data(), v.data(), out);
  }

  int de_matrix_create_distributed(de_context *ctx, int64_t n_owned, int64_t n_halo, int64_t nnz,
                                   const int64_t *rowptr, const int64_t *col_local, const double *val, int npeers,
                                   const int *peer_ranks, const int64_t *recv_counts, const int64_t *send_offsets,
                                   const int64_t *send_rows, de_matrix **out)
  {
    if (!ctx || !out || n_owned < 0 || n_halo < 0 || nnz < 0 || !rowptr || npeers < 0 ||
        (npeers > 0 && (!peer_ranks || !recv_counts || !send_offsets)))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: bad arguments");
    *out = nullptr;
    DE_TRY(bind_device(ctx));
    de_matrix *A = new de_matrix();
    A->ctx = ctx;
    context_retain(ctx);
    auto fail = [&](int s) {
      de_matrix_destroy(A);
      return s;
    };
    SetupTrace trace("create_distributed", ctx->rank);
    int s = matrix_upload(ctx, n_owned, n_owned + n_halo, nnz, rowptr, col_local, val, A);
    if (s != DE_OK)
      return fail(s);
    trace.lap("matrix upload");
    A->n_halo = n_halo;
    A->npeers = npeers;
    long long roff = 0;
    for (int p = 0; p < npeers; ++p)
    {
      if (peer_ranks[p] < 0 || peer_ranks[p] >= ctx->nranks || peer_ranks[p] == ctx->rank ||
          (p > 0 && peer_ranks[p] <= peer_ranks[p - 1]))
        return fail(set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: peers must be ascending ranks != own"));
      A->peer.push_back(peer_ranks[p]);
      A->recv_count.push_back(recv_counts[p]);
      A->recv_off.push_back(roff);
      roff += recv_counts[p];
      A->send_count.push_back(send_offsets[p + 1] - send_offsets[p]);
      A->send_off.push_back(send_offsets[p]);
    }
    if (roff != n_halo)
      return fail(set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: recv_counts do not sum to n_halo"));
    A->n_send = npeers > 0 ? send_offsets[npeers] : 0;
    for (long long k = 0; k < A->n_send; ++k)
      if (send_rows[k] < 0 || send_rows[k] >= n_owned)
        return fail(set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: send row out of range"));
    if ((s = upload_converted(ctx, &A->send_rows, send_rows, (size_t)A->n_send)) != DE_OK)
      return fail(s);
    for (int p = 0; p < npeers; ++p) // slab partitions send whole planes: consecutive rows (plan_fused_push)
    {
      long long first = send_offsets[p + 1] > send_offsets[p] ? send_rows[send_offsets[p]] : -1;
      for (int64_t k = send_offsets[p]; k < send_offsets[p + 1] && first >= 0; ++k)
        if (send_rows[k] != first + (k - send_offsets[p]))
          first = -1;
      A->send_first.push_back(first);
    }
    // interior rows touch owned columns only and can run while the halo is in flight
    std::vector<int> in, bd;
    {
      const int nth = (int)std::max<long long>(1, std::min<long long>(8, nnz >> 20));
      std::vector<std::vector<int>> pin((size_t)nth), pbd((size_t)nth);
      auto scan = [&](int t)
      {
        const long long i0 = n_owned * t / nth, i1 = n_owned * (t + 1) / nth;
        for (long long i = i0; i < i1; ++i)
        {
          bool halo = false;
          for (int64_t k = rowptr[i]; k < rowptr[i + 1] && !halo; ++k)
            halo = col_local[k] >= n_owned;
          (halo ? pbd[t] : pin[t]).push_back((int)i);
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nth; ++t)
        th.emplace_back(scan, t);
      scan(0);
      for (auto &x : th)
        x.join();
      for (int t = 0; t < nth; ++t) // chunks are ascending row ranges: concatenation keeps the lists sorted
      {
        in.insert(in.end(), pin[t].begin(), pin[t].end());
        bd.insert(bd.end(), pbd[t].begin(), pbd[t].end());
      }
    }
    trace.lap("interior / boundary scan");
    A->n_interior = (long long)in.size();
    A->n_boundary = (long long)bd.size();
    if ((s = upload_converted(ctx, &A->interior, in.data(), in.size())) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &A->boundary, bd.data(), bd.size())) != DE_OK)
      return fail(s);
    trace.lap("row lists uploaded");
    if ((s = build_brb(ctx, A, n_owned, n_owned + n_halo, rowptr, col_local, val)) != DE_OK)
      return fail(s);
    trace.lap("BRB form built");
    // the row-permuted CSR copies for the staged kernel are only built when the CSR family can be chosen by the AUTO
    // policy (brb_usable); a forced DE_SPMM_CSR then still works through the row-list kernel
    if (!(A->brb.valid && nnz >= 12 * n_owned))
    {
      if ((s = build_staged_subset(ctx, in, rowptr, col_local, val, A->st_interior)) != DE_OK)
        return fail(s);
      if ((s = build_staged_subset(ctx, bd, rowptr, col_local, val, A->st_boundary)) != DE_OK)
        return fail(s);
    }
    *out = A;
    return DE_OK;
  }

  int de_matrix_destroy(de_matrix *A)
  {
    if (!A)
      return DE_OK;
    de_context *ctx = A->ctx;
    cudaSetDevice(ctx->device);
    dev_free(A->rowptr);
    dev_free(A->col);
    dev_free(A->val);
    dev_free(A->send_rows);
    dev_free(A->interior);
    dev_free(A->boundary);
    dev_free(A->send_buf);
    dev_free(A->halo_buf);
    A->st_all.release();
    A->st_interior.release();
    A->st_boundary.release();
    A->brb.release();
    delete A;
    context_release(ctx);
    return DE_OK;
  }

  int de_matrix_rows(const de_matrix *A, int64_t *n_owned, int64_t *nnz)
  {
    if (!A)
      return set_error(nullptr, DE_ERR_INVALID, "null matrix");
    if (n_owned)
      *n_owned = A->n;
    if (nnz)
      *nnz = A->nnz;
    return DE_OK;
  }

  int de_brb_format_check(int64_t n, int64_t ncols, int64_t n_owned, const int64_t *rowptr, const int64_t *col,
                          const double *val, int nthreads, int64_t *info8, double *max_abs_diff)
  {
    if (n < 0 || ncols < 0 || !rowptr || !info8 || !max_abs_diff)
      return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: bad arguments");
    for (int i = 0; i < 8; ++i)
      info8[i] = 0;
    *max_abs_diff = 0.0;
    de::brb::Format F;
    if (n == 0 || !de::brb::build(n, ncols, rowptr, col, val, n_owned, F, nthreads) || !F.valid)
      return DE_OK; // info8[0] == 0: no BRB form
    info8[0] = 1;
    info8[1] = F.grid ? 1 : 0;
    info8[2] = F.ntiles;
    info8[3] = F.n_interior;
    info8[4] = F.nblocks;
    info8[5] = F.nsteps;
    info8[6] = F.max_u;
    info8[7] = (int64_t)F.tw | ((int64_t)F.th << 16) | ((int64_t)F.td << 32);
    // decode every tile exactly as the kernel does and apply it to a probe vector
    auto probe = [](int64_t c) { return 1.0 + (double)((c * 2654435761ull) % 1021) / 1021.0; };
    std::vector<double> y((size_t)n, 0.0);
    std::vector<char> seen((size_t)n, 0);
    for (const de::brb::TileDesc &d : F.tile)
    {
      const int *h = &F.blob[(size_t)d.blob16 * 4];
      const int nb = h[0], ns = h[1];
      const int *blkstep = h + 4, *blkrows = blkstep + nb + 1;
      const int o_step = (4 + (nb + 1) + 8 * nb + 4 * nb + 3) & ~3;
      const unsigned short *self = reinterpret_cast<const unsigned short *>(blkrows + 8 * nb);
      const int *st = h + o_step;
      const double *v = reinterpret_cast<const double *>(h + o_step + 4 * ns);
      if (d.nu != h[3] || (size_t)d.len16 * 4 < (size_t)o_step + 4 * (size_t)ns + 2 * (size_t)h[2])
        return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: inconsistent tile header");
      for (int b = 0; b < nb; ++b)
      {
        for (int g = 0; g < 8; ++g)
          if (blkrows[8 * b + g] >= 0)
          {
            seen[blkrows[8 * b + g]]++;
            // the recorded position of the row's own column must hold exactly that column
            const unsigned sl = self[8 * b + g];
            if (sl != 0xffffu && ((int)sl >= d.nu || F.ucol[(size_t)d.ucol0 + sl] != blkrows[8 * b + g]))
              return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: wrong self column id");
          }
        for (int q = blkstep[b]; q < blkstep[b + 1]; ++q)
        {
          const unsigned lc[4] = {(unsigned)st[4 * q] & 0xffffu, (unsigned)st[4 * q] >> 16, (unsigned)st[4 * q + 1] & 0xffffu,
                                  (unsigned)st[4 * q + 1] >> 16};
          const unsigned mask = (unsigned)st[4 * q + 2];
          int k = st[4 * q + 3];
          for (int bit = 0; bit < 32; ++bit)
            if ((mask >> bit) & 1u)
            {
              const int row = blkrows[8 * b + bit / 4];
              if (row < 0 || row >= n || (int)lc[bit % 4] >= d.nu)
                return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: entry outside its tile");
              y[row] += v[k++] * probe(F.ucol[(size_t)d.ucol0 + lc[bit % 4]]);
            }
        }
      }
    }
    double mx = 0.0;
    for (int64_t r = 0; r < n; ++r)
    {
      if (seen[r] != 1)
        return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: a row is not covered exactly once");
      double ref = 0.0;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        ref += val[k] * probe(col[k]);
      mx = std::max(mx, std::fabs(ref - y[r]));
    }
    *max_abs_diff = mx;
    return DE_OK;
  }

  int de_matrix_brb_selfcheck(const de_matrix *A, int64_t n, int64_t ncols, const int64_t *rowptr, const int64_t *col,
                              const double *val, int64_t *mismatches)
  {
    if (!A || !rowptr || !mismatches)
      return set_error(A ? A->ctx : nullptr, DE_ERR_INVALID, "de_matrix_brb_selfcheck: bad arguments");
    de_context *ctx = A->ctx;
    *mismatches = -1;
    if (!A->brb.valid)
      return DE_OK;
    if (n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_brb_selfcheck: matrix size does not match");
    DE_TRY(bind_device(ctx));
    de::brb::Format F;
    if (!de::brb::build(n, ncols, rowptr, col, val, n, F))
    {
      *mismatches = -2; // the host builder found no BRB form although the device did
      return DE_OK;
    }
    const BrbDevice &B = A->brb;
    int64_t bad = 0;
    bad += (B.ntiles != F.ntiles) + (B.n_interior != F.n_interior) + (B.max_len16 != F.max_len16) + (B.max_u != F.max_u) +
           (B.nsteps != F.nsteps) + (B.nvals != F.nvals) + (B.blob16 * 4 != F.blob.size()) + (B.nucol != F.ucol.size());
    if (bad == 0)
    {
      std::vector<int> blob(F.blob.size()), ucol(F.ucol.size());
      std::vector<de::brb::TileDesc> tile(F.tile.size());
      DE_CUDA(ctx, cudaMemcpy(blob.data(), B.blob, blob.size() * sizeof(int), cudaMemcpyDeviceToHost));
      DE_CUDA(ctx, cudaMemcpy(ucol.data(), B.ucol, ucol.size() * sizeof(int), cudaMemcpyDeviceToHost));
      DE_CUDA(ctx, cudaMemcpy(tile.data(), B.tile, tile.size() * sizeof(int4), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < blob.size(); ++i)
        bad += blob[i] != F.blob[i];
      for (size_t i = 0; i < ucol.size(); ++i)
        bad += ucol[i] != F.ucol[i];
      bad += std::memcmp(tile.data(), F.tile.data(), tile.size() * sizeof(int4)) != 0;
    }
    else
      bad += 1000000;
    *mismatches = bad;
    return DE_OK;
  }

  int de_matrix_set_spmm_format(de_matrix *A, int format)
  {
    if (!A)
      return set_error(nullptr, DE_ERR_INVALID, "null matrix");
    if (format != DE_SPMM_AUTO && format != DE_SPMM_CSR && format != DE_SPMM_BRB)
      return set_error(A->ctx, DE_ERR_INVALID, "de_matrix_set_spmm_format: unknown format");
    if (format == DE_SPMM_BRB && !A->brb.valid)
      return set_error(A->ctx, DE_ERR_UNSUPPORTED, "de_matrix_set_spmm_format: this matrix has no BRB form");
    A->spmm_format = format;
    return DE_OK;
  }

  int de_matrix_spmm_info(const de_matrix *A, int *format, int64_t *tiles, int64_t *row_blocks, int64_t *steps,
                          int64_t *union_rows_max, int *tile_shape3)
  {
    if (!A)
      return set_error(nullptr, DE_ERR_INVALID, "null matrix");
    const bool brb = A->brb.valid && A->spmm_format != DE_SPMM_CSR;
    if (format)
      *format = brb ? DE_SPMM_BRB : DE_SPMM_CSR;
    if (tiles)
      *tiles = A->brb.valid ? A->brb.ntiles : 0;
    if (row_blocks)
      *row_blocks = A->brb.valid ? A->brb.nblocks : 0;
    if (steps)
      *steps = A->brb.valid ? A->brb.nsteps : 0;
    if (union_rows_max)
      *union_rows_max = A->brb.valid ? A->brb.max_u : 0;
    if (tile_shape3)
    {
      tile_shape3[0] = A->brb.grid ? A->brb.tw : 0;
      tile_shape3[1] = A->brb.grid ? A->brb.th : 0;
      tile_shape3[2] = A->brb.grid ? A->brb.td : 0;
    }
    return DE_OK;
  }

  // ---- kernels --------------------------------------------------------------------------------------------
  static int check_spmm_shapes(de_context *ctx, const char *who, const de_mv *Y, const de_matrix *A, const de_mv *X)
  {
    if (Y->n != A->n || X->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": number of rows does not match the matrix");
    if (Y->m != X->m)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": number of columns does not match");
    if (Y->d == X->d)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": output must not alias input");
    return DE_OK;
  }

  int de_spmm(de_mv *Y, const de_matrix *A, const de_mv *X)
  {
    if (!Y || !A || !X)
      return set_error(nullptr, DE_ERR_INVALID, "de_spmm: null argument");
    de_context *ctx = A->ctx;
    DE_TRY(check_spmm_shapes(ctx, "matmul_sparse_tallskinny", Y, A, X));
    DE_TRY(bind_device(ctx));
    return spmm_device(ctx, A, X->d, Y->d, X->m, false);
  }

  int de_spmm_diag_dot(de_mv *Y, const de_matrix *A, const de_mv *X, double *dp_host)
  {
    if (!Y || !A || !X || !dp_host)
      return set_error(nullptr, DE_ERR_INVALID, "de_spmm_diag_dot: null argument");
    de_context *ctx = A->ctx;
    DE_TRY(check_spmm_shapes(ctx, "matmul_sparse_tallskinny", Y, A, X));
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    DE_TRY(spmm_device(ctx, A, X->d, Y->d, X->m, true));
    return fetch_small(ctx, ctx->dDP(), dp_host, X->m);
  }

  int de_spmm_gram(de_mv *Y, const de_matrix *A, const de_mv *X, double *dp_host, double *G_host)
  {
    DE_TRY(check_spmm_shapes(Y ? Y->ctx : nullptr, "de_spmm_gram", Y, A, X));
    if (!dp_host || !G_host)
      return set_error(Y->ctx, DE_ERR_INVALID, "de_spmm_gram: null output");
    de_context *ctx = Y->ctx;
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    const int m = X->m;
    bool fused = false;
    DE_TRY(spmm_device(ctx, A, X->d, Y->d, m, true, &fused));
    const double *G = ctx->dDG() + m;
    if (!fused)
    {
      // no Gram epilogue for this matrix / width: a separate pass over Y
      DE_TRY(gram_device(ctx, m, Y->n, Y->d, m, Y->d, m, true, ctx->dG()));
      G = ctx->dG();
    }
    DE_TRY(fetch_small(ctx, ctx->dDP(), dp_host, (size_t)m));
    DE_CUDA(ctx, cudaMemcpy(G_host, G, sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    return DE_OK;
  }


} // extern "C"
