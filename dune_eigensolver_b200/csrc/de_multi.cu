// de_multi.cu -- row-partitioned multi-GPU front ends of the C ABI (new; the reference is single-threaded).
//
//  * de_matrix_create_rowblock: one rank's part of a row-partitioned matrix from its rows with GLOBAL column indices. The
//    halo plan (which rows of the vector block travel where) is derived here from the column indices; the only thing the
//    caller supplies is an all-gather over the ranks of the job (torch.distributed / MPI / the in-process one below).
//  * de_multi_*: ONE process, one host thread and one context per GPU ("de_context_create(const int*
//    device_ids, int ndev, ...)"). The C++ drop-in headers use it when more than one device is configured, so the
//    reference's free functions run row-partitioned without Python, MPI or NCCL: the windows of the NVLink data path
//    (kernels_peer.cuh) are plain peer-mapped allocations of the same process.
#include <condition_variable>
#include <cstdlib>
#include <memory>

#include "de_internal.hpp"

using namespace dei;

namespace
{
  /** barrier of the rank threads that can be broken: a rank that fails releases the others instead of hanging them */
  struct RankBarrier
  {
    std::mutex mu;
    std::condition_variable cv;
    int n = 1, waiting = 0;
    unsigned long long generation = 0;
    bool broken = false;
    bool wait()
    {
      std::unique_lock<std::mutex> lock(mu);
      if (broken)
        return false;
      const unsigned long long g = generation;
      if (++waiting == n)
      {
        waiting = 0;
        ++generation;
        cv.notify_all();
        return true;
      }
      cv.wait(lock, [&] { return generation != g || broken; });
      return !broken;
    }
    void abort()
    {
      std::lock_guard<std::mutex> lock(mu);
      broken = true;
      cv.notify_all();
    }
  };
}

struct de_multi
{
  int ndev = 0;
  std::vector<int> device;
  std::vector<de_context *> ctx;
  std::string err;
  // in-process all-gather
  RankBarrier barrier;
  std::vector<const void *> ag_src;
};

namespace
{
  struct RankUser
  {
    de_multi *M;
    int rank;
  };

  /** de_allgather_fn over the rank threads of one process */
  int inprocess_allgather(void *user, const void *send, void *recv, int64_t bytes)
  {
    RankUser *u = static_cast<RankUser *>(user);
    de_multi *M = u->M;
    M->ag_src[u->rank] = send;
    if (!M->barrier.wait())
      return 1;
    for (int q = 0; q < M->ndev; ++q)
      std::memcpy(static_cast<unsigned char *>(recv) + (size_t)q * bytes, M->ag_src[q], (size_t)bytes);
    if (!M->barrier.wait()) // nobody may overwrite its send buffer before everyone has copied
      return 1;
    return 0;
  }

  int multi_error(de_multi *M, int code, const std::string &msg)
  {
    if (M)
      M->err = msg;
    return set_error(nullptr, code, msg);
  }

  /** contiguous row partition; cuts only at multiples of `align` rows (a grid plane keeps the halo one plane thick) */
  std::vector<int64_t> partition_rows(int64_t n, int nranks, int64_t align)
  {
    if (align < 1 || n % align != 0 || n / align < nranks)
      align = 1;
    const int64_t units = n / align, base = units / nranks, extra = units % nranks;
    std::vector<int64_t> part(nranks + 1, 0);
    for (int r = 0; r < nranks; ++r)
      part[r + 1] = part[r] + (base + (r < extra ? 1 : 0)) * align;
    part[nranks] = n;
    return part;
  }

  enum DriverKind
  {
    kLargest,
    kLobpcg,
    kGenLobpcg
  };

  struct DriverCall
  {
    DriverKind kind;
    int64_t n, nnz;
    const int64_t *rowptr, *col;
    const double *val;
    const int64_t *b_rowptr, *b_col; // generalized problem: B (same row partition)
    const double *b_val;
    int64_t row_align;
    double shift, tol;
    int maxiter, nev;
    const double *start_panel8;
    double *eval, *evec;
    int verbose;
    int *iterations;
  };

  int run_rank(de_multi *M, int rank, const DriverCall &c, const std::vector<int64_t> &part, std::string &err, int *iters)
  {
    de_context *ctx = M->ctx[rank];
    auto fail = [&](int s) {
      err = de_last_error_string(nullptr);
      if (err.empty())
        err = de_last_error_string(ctx);
      M->barrier.abort();
      return s;
    };
    if (cudaSetDevice(M->device[rank]) != cudaSuccess)
      return fail(set_error(ctx, DE_ERR_CUDA, "de_multi: cudaSetDevice failed"));
    const int64_t r0 = part[rank], r1 = part[rank + 1], nl = r1 - r0;
    const int m = padded_cols(c.nev);
    RankUser user{M, rank};
    auto slice = [&](const int64_t *rowptr, const int64_t *col, const double *val, de_matrix **out) {
      std::vector<int64_t> rp((size_t)nl + 1);
      for (int64_t i = 0; i <= nl; ++i)
        rp[i] = rowptr[r0 + i] - rowptr[r0];
      return de_matrix_create_rowblock(ctx, nl, rp[nl], rp.data(), col + rowptr[r0], val + rowptr[r0], part.data(),
                                       inprocess_allgather, &user, out);
    };
    de_matrix *A = nullptr, *B = nullptr;
    int s = slice(c.rowptr, c.col, c.val, &A);
    if (s == DE_OK && c.kind == kGenLobpcg)
      s = slice(c.b_rowptr, c.b_col, c.b_val, &B);
    if (s != DE_OK)
    {
      de_matrix_destroy(A);
      de_matrix_destroy(B);
      return fail(s);
    }
    // this rank's rows of the global start block, in the panel layout of its own n (multivector.hh:130-133)
    std::vector<double> start((size_t)nl * m), evec((size_t)nl * c.nev), eval((size_t)c.nev);
    for (int p = 0; p < m / 8; ++p)
      std::memcpy(start.data() + (size_t)p * nl * 8, c.start_panel8 + ((size_t)p * c.n + r0) * 8, sizeof(double) * 8 * (size_t)nl);
    int it = 0;
    switch (c.kind)
    {
    case kLargest:
      s = de_standard_largest(ctx, A, c.shift, c.tol, c.maxiter, c.nev, start.data(), eval.data(), evec.data(), rank == 0 ? c.verbose : 0, &it);
      break;
    case kLobpcg:
      s = de_standard_lobpcg(ctx, A, c.tol, c.maxiter, c.nev, start.data(), eval.data(), evec.data(), rank == 0 ? c.verbose : 0, &it);
      break;
    case kGenLobpcg:
      s = de_generalized_lobpcg(ctx, A, B, c.tol, c.maxiter, c.nev, start.data(), eval.data(), evec.data(), rank == 0 ? c.verbose : 0, &it);
      break;
    }
    de_matrix_destroy(A);
    de_matrix_destroy(B);
    if (s != DE_OK)
      return fail(s);
    for (int j = 0; j < c.nev; ++j)
      std::memcpy(c.evec + (size_t)j * c.n + r0, evec.data() + (size_t)j * nl, sizeof(double) * (size_t)nl);
    if (rank == 0)
      for (int j = 0; j < c.nev; ++j)
        c.eval[j] = eval[j];
    *iters = it;
    return DE_OK;
  }

  int run_driver(de_multi *M, const DriverCall &c)
  {
    if (!M || !c.rowptr || !c.start_panel8 || !c.eval || !c.evec || c.n < 0 || c.nev <= 0 || (c.nnz > 0 && (!c.col || !c.val)) ||
        (c.kind == kGenLobpcg && (!c.b_rowptr || !c.b_col || !c.b_val)))
      return multi_error(M, DE_ERR_INVALID, "de_multi driver: bad arguments");
    if (c.rowptr[c.n] != c.nnz)
      return multi_error(M, DE_ERR_INVALID, "de_multi driver: rowptr does not match nnz");
    const int R = M->ndev;
    const std::vector<int64_t> part = partition_rows(c.n, R, c.row_align);
    std::vector<int> status(R, DE_OK), iters(R, 0);
    std::vector<std::string> errs(R);
    {
      std::lock_guard<std::mutex> lock(M->barrier.mu);
      M->barrier.broken = false;
      M->barrier.waiting = 0;
    }
    std::vector<std::thread> th;
    for (int r = 1; r < R; ++r)
      th.emplace_back([&, r] { status[r] = run_rank(M, r, c, part, errs[r], &iters[r]); });
    status[0] = run_rank(M, 0, c, part, errs[0], &iters[0]);
    for (auto &t : th)
      t.join();
    // a rank that fails stops contributing, and its peers then time out waiting for it: report every rank's message
    int first = DE_OK;
    std::string all;
    for (int r = 0; r < R; ++r)
      if (status[r] != DE_OK)
      {
        if (first == DE_OK || (first == DE_ERR_NCCL && status[r] != DE_ERR_NCCL))
          first = status[r];
        all += (all.empty() ? "rank " : " | rank ") + std::to_string(r) + ": " + errs[r];
      }
    if (first != DE_OK)
      return multi_error(M, first, all);
    if (c.iterations)
      *c.iterations = iters[0];
    return DE_OK;
  }
}

extern "C"
{

  int de_matrix_create_rowblock(de_context *ctx, int64_t n_owned, int64_t nnz, const int64_t *rowptr,
                                const int64_t *col_global, const double *val, const int64_t *part,
                                de_allgather_fn allgather, void *user, de_matrix **out)
  {
    if (!ctx || !out || n_owned < 0 || nnz < 0 || !rowptr || !part || (nnz > 0 && (!col_global || !val)))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_rowblock: bad arguments");
    *out = nullptr;
    const int R = ctx->nranks, me = ctx->rank;
    if (R > 1 && !allgather)
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_rowblock: an all-gather is needed on more than one rank");
    if (part[me + 1] - part[me] != n_owned || rowptr[n_owned] != nnz)
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_rowblock: partition / rowptr do not match the row block");
    if (R == 1)
      return de_matrix_create_csr(ctx, n_owned, nnz, rowptr, col_global, val, out);
    // (no value-initialisation: 2 x 8 nnz bytes of zero fill would cost more than the planning itself)
    static const bool trace = std::getenv("DE_TRACE_SETUP") != nullptr; // debugging aid: prints where the setup time goes
    const auto t_begin = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) {
      if (trace)
        std::fprintf(stderr, "[de setup] rank %d %-28s %8.2f ms\n", me, what,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() * 1e3);
    };
    std::unique_ptr<int64_t[]> col_local_buf(new int64_t[(size_t)std::max<int64_t>(nnz, 1)]);
    std::unique_ptr<int64_t[]> halo_buf(new int64_t[(size_t)std::max<int64_t>(nnz, 1)]);
    struct Span
    {
      int64_t *p;
      int64_t *data() const { return p; }
      int64_t *begin() const { return p; }
    } col_local{col_local_buf.get()}, halo{halo_buf.get()};
    std::vector<int64_t> recv((size_t)R, 0);
    int64_t n_halo = 0;
    DE_TRY(de_halo_plan_local(n_owned, rowptr, col_global, R, me, part, col_local.data(), halo.data(), &n_halo, recv.data()));
    lap("halo_plan_local");
    // every rank's per-owner halo counts, then every rank's halo list (padded to the longest)
    std::vector<int64_t> counts((size_t)R * R);
    if (allgather(user, recv.data(), counts.data(), (int64_t)sizeof(int64_t) * R) != 0)
      return set_error(ctx, DE_ERR_NCCL, "de_matrix_create_rowblock: all-gather of the halo counts failed");
    int64_t max_halo = 0;
    std::vector<int64_t> total((size_t)R, 0);
    for (int q = 0; q < R; ++q)
    {
      for (int p = 0; p < R; ++p)
        total[q] += counts[(size_t)q * R + p];
      max_halo = std::max(max_halo, total[q]);
    }
    std::vector<int64_t> mine((size_t)std::max<int64_t>(max_halo, 1), -1), lists((size_t)R * std::max<int64_t>(max_halo, 1));
    std::copy(halo.begin(), halo.begin() + n_halo, mine.begin());
    if (allgather(user, mine.data(), lists.data(), (int64_t)sizeof(int64_t) * std::max<int64_t>(max_halo, 1)) != 0)
      return set_error(ctx, DE_ERR_NCCL, "de_matrix_create_rowblock: all-gather of the halo lists failed");
    std::vector<int> peers((size_t)R);
    std::vector<int64_t> recv_counts((size_t)R), send_off((size_t)R + 1), deposit((size_t)R);
    int64_t n_send = 0;
    for (int q = 0; q < R; ++q)
      if (q != me)
        n_send += counts[(size_t)q * R + me];
    std::vector<int64_t> send_rows((size_t)std::max<int64_t>(n_send, 1));
    int npeers = 0, all_symmetric = 0;
    const int64_t stride = std::max<int64_t>(max_halo, 1);
    DE_TRY(de_halo_plan_peers(R, me, part, counts.data(), lists.data(), stride, &npeers, peers.data(), recv_counts.data(),
                              send_off.data(), send_rows.data(), deposit.data(), &max_halo, &all_symmetric));
    lap("all-gathers + peer plan");
    DE_TRY(de_matrix_create_distributed(ctx, n_owned, n_halo, nnz, rowptr, col_local.data(), val, npeers, peers.data(),
                                        recv_counts.data(), send_off.data(), send_rows.data(), out));
    lap("create_distributed");
    // The credit-free flow control of the peer-store halo exchange (kernels_peer.cuh) needs every send peer to be a
    // receive peer as well; an unsymmetric pattern keeps the NCCL path (the same decision on all ranks: it is derived
    // from the all-gathered counts alone).
    if (ctx->peer_ready && all_symmetric)
    {
      const int s = de_matrix_set_peer_deposit(*out, deposit.data(), max_halo);
      if (s != DE_OK)
      {
        de_matrix_destroy(*out);
        *out = nullptr;
        return s;
      }
    }
    return DE_OK;
  }

  int de_halo_plan_peers(int nranks, int rank, const int64_t *part, const int64_t *counts_all, const int64_t *lists_all,
                         int64_t list_stride, int *npeers, int *peer_ranks, int64_t *recv_counts, int64_t *send_offsets,
                         int64_t *send_rows, int64_t *deposit_rows, int64_t *max_halo_rows, int *symmetric)
  {
    if (nranks < 1 || rank < 0 || rank >= nranks || !part || !counts_all || !lists_all || list_stride < 0 || !npeers ||
        !peer_ranks || !recv_counts || !send_offsets || !send_rows || !deposit_rows)
      return set_error(nullptr, DE_ERR_INVALID, "de_halo_plan_peers: bad arguments");
    const int R = nranks, me = rank;
    int np = 0;
    int64_t ns = 0, mx = 0;
    send_offsets[0] = 0;
    for (int q = 0; q < R; ++q)
    {
      int64_t tot = 0;
      for (int p = 0; p < R; ++p)
        tot += counts_all[(size_t)q * R + p];
      if (tot > list_stride)
        return set_error(nullptr, DE_ERR_INVALID, "de_halo_plan_peers: a halo list is longer than list_stride");
      mx = std::max(mx, tot);
      if (q == me)
        continue;
      const int64_t to_q = counts_all[(size_t)q * R + me], from_q = counts_all[(size_t)me * R + q];
      if (to_q == 0 && from_q == 0)
        continue;
      peer_ranks[np] = q;
      recv_counts[np] = from_q;
      int64_t off = 0; // q's halo block is ordered by owner rank: my rows start behind those of the ranks below me
      for (int p = 0; p < me; ++p)
        off += counts_all[(size_t)q * R + p];
      deposit_rows[np] = off;
      for (int64_t k = 0; k < to_q; ++k)
      {
        const int64_t g = lists_all[(size_t)q * list_stride + off + k];
        if (g < part[me] || g >= part[me + 1])
          return set_error(nullptr, DE_ERR_INVALID, "de_halo_plan_peers: inconsistent halo lists");
        send_rows[ns++] = g - part[me]; // in q's halo order
      }
      send_offsets[++np] = ns;
    }
    *npeers = np;
    if (max_halo_rows)
      *max_halo_rows = mx;
    if (symmetric)
    {
      *symmetric = 1;
      for (int a = 0; a < R; ++a)
        for (int b = 0; b < R; ++b)
          if (a != b && (counts_all[(size_t)a * R + b] > 0) != (counts_all[(size_t)b * R + a] > 0))
            *symmetric = 0;
    }
    return DE_OK;
  }

  int de_multi_create(const int *device_ids, int ndev, int64_t halo_bytes, de_multi **out)
  {
    if (!out || !device_ids || ndev < 1 || ndev > de::kPeerMaxRanks || halo_bytes < 0)
      return multi_error(nullptr, DE_ERR_INVALID, "de_multi_create: needs 1..8 device ordinals");
    *out = nullptr;
    de_multi *M = new de_multi();
    M->ndev = ndev;
    M->device.assign(device_ids, device_ids + ndev);
    M->ctx.assign(ndev, nullptr);
    M->ag_src.assign(ndev, nullptr);
    M->barrier.n = ndev;
    auto bail = [&](int code, const std::string &msg) {
      de_multi_destroy(M);
      return multi_error(nullptr, code, msg);
    };
    for (int r = 0; r < ndev; ++r)
    {
      const int s = de_context_create(device_ids[r], nullptr, &M->ctx[r]);
      if (s != DE_OK)
        return bail(s, std::string("de_multi_create: ") + de_last_error_string(nullptr));
      M->ctx[r]->rank = r;
      M->ctx[r]->nranks = ndev;
    }
    if (ndev == 1)
    {
      *out = M;
      return DE_OK;
    }
    // several ranks on one device (test configuration): no programmatic dependent launch -- see launch_pdl
    bool shared_device = false;
    for (int r = 0; r < ndev; ++r)
      for (int q = 0; q < r; ++q)
        shared_device = shared_device || device_ids[q] == device_ids[r];
    if (shared_device)
      for (int r = 0; r < ndev; ++r)
        M->ctx[r]->pdl = false;
    // windows of the NVLink data path: ordinary allocations, mapped into the peers by enabling peer access
    const size_t cap = (((size_t)(halo_bytes > 0 ? halo_bytes : (int64_t)128 << 20)) + 255) & ~(size_t)255;
    for (int r = 0; r < ndev; ++r)
    {
      de_context *c = M->ctx[r];
      cudaSetDevice(c->device);
      for (int q = 0; q < ndev; ++q)
        if (device_ids[q] != c->device)
        {
          int can = 0;
          cudaDeviceCanAccessPeer(&can, c->device, device_ids[q]);
          if (!can)
            return bail(DE_ERR_UNSUPPORTED, "de_multi_create: devices " + std::to_string(c->device) + " and " +
                                                std::to_string(device_ids[q]) + " have no peer access");
          const cudaError_t e = cudaDeviceEnablePeerAccess(device_ids[q], 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return bail(DE_ERR_CUDA, std::string("de_multi_create: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
          cudaGetLastError();
        }
      c->halo_cap = cap;
      c->window_bytes = de::kPeerHaloOff + 2 * cap;
      if (cudaMalloc((void **)&c->window, c->window_bytes) != cudaSuccess || cudaMemset(c->window, 0, c->window_bytes) != cudaSuccess ||
          cudaMalloc((void **)&c->dticket, 2 * sizeof(int)) != cudaSuccess || cudaMemset(c->dticket, 0, 2 * sizeof(int)) != cudaSuccess ||
          cudaDeviceSynchronize() != cudaSuccess)
        return bail(DE_ERR_ALLOC, std::string("de_multi_create: window allocation failed: ") + cudaGetErrorString(cudaGetLastError()));
    }
    for (int r = 0; r < ndev; ++r)
    {
      de_context *c = M->ctx[r];
      // no kernel may be loaded lazily once ranks can spin on each other's flags (see preload_kernels)
      if (preload_kernels(c) != DE_OK)
        return bail(DE_ERR_CUDA, std::string("de_multi_create: loading the kernels failed: ") + de_last_error_string(c));
      for (int q = 0; q < ndev; ++q)
        c->peer_base[q] = M->ctx[q]->window;
      c->peer_ipc = false;
      c->ar_epoch = c->ar_epoch_b = c->halo_epoch = 0;
      c->peer_ready = true;
    }
    *out = M;
    return DE_OK;
  }

  int de_multi_destroy(de_multi *M)
  {
    if (!M)
      return DE_OK;
    for (de_context *c : M->ctx)
      if (c)
      {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize(); // no peer may still be writing into a window that is about to be freed
      }
    for (de_context *c : M->ctx)
      de_context_destroy(c);
    delete M;
    return DE_OK;
  }

  int de_multi_set_timeout(de_multi *M, double seconds)
  {
    if (!M || !(seconds > 0.0))
      return multi_error(M, DE_ERR_INVALID, "de_multi_set_timeout: bad arguments");
    for (de_context *c : M->ctx)
    {
      int khz = 2000000;
      cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device);
      c->peer_timeout_cycles = (long long)(seconds * 1e3 * (double)khz);
    }
    return DE_OK;
  }

  int de_multi_size(const de_multi *M, int *ndev)
  {
    if (!M || !ndev)
      return multi_error(nullptr, DE_ERR_INVALID, "de_multi_size: null argument");
    *ndev = M->ndev;
    return DE_OK;
  }

  int de_multi_context(de_multi *M, int rank, de_context **ctx)
  {
    if (!M || !ctx || rank < 0 || rank >= M->ndev)
      return multi_error(M, DE_ERR_INVALID, "de_multi_context: bad arguments");
    *ctx = M->ctx[rank];
    return DE_OK;
  }

  const char *de_multi_last_error(const de_multi *M) { return M ? M->err.c_str() : de_last_error_string(nullptr); }

  int de_multi_launch_count(const de_multi *M, int64_t *count)
  {
    if (!M || !count)
      return multi_error(nullptr, DE_ERR_INVALID, "de_multi_launch_count: null argument");
    *count = 0;
    for (const de_context *c : M->ctx)
      *count += c->launches;
    return DE_OK;
  }

  int de_multi_standard_largest(de_multi *M, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                                const double *val, int64_t row_align, double shift, double tol, int maxiter, int nev,
                                const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    DriverCall c{kLargest, n, nnz, rowptr, col, val, nullptr, nullptr, nullptr, row_align, shift, tol, maxiter, nev, start_panel8, eval, evec,
                 verbose, iterations};
    return run_driver(M, c);
  }

  int de_multi_standard_lobpcg(de_multi *M, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                               const double *val, int64_t row_align, double tol, int maxiter, int nev,
                               const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    DriverCall c{kLobpcg, n, nnz, rowptr, col, val, nullptr, nullptr, nullptr, row_align, 0.0, tol, maxiter, nev, start_panel8, eval, evec,
                 verbose, iterations};
    return run_driver(M, c);
  }

  int de_multi_generalized_lobpcg(de_multi *M, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                                  const double *val, int64_t b_nnz, const int64_t *b_rowptr, const int64_t *b_col,
                                  const double *b_val, int64_t row_align, double tol, int maxiter, int nev,
                                  const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    if (b_rowptr && b_rowptr[n] != b_nnz)
      return multi_error(M, DE_ERR_INVALID, "de_multi_generalized_lobpcg: B's rowptr does not match its nnz");
    DriverCall c{kGenLobpcg, n, nnz, rowptr, col, val, b_rowptr, b_col, b_val, row_align, 0.0, tol, maxiter, nev, start_panel8, eval, evec,
                 verbose, iterations};
    return run_driver(M, c);
  }

} // extern "C"
