#ifndef DUNE_EIGENSOLVER_B200_SUPERNODAL_CHOLESKY_HH
#define DUNE_EIGENSOLVER_B200_SUPERNODAL_CHOLESKY_HH

/** \file
 *  Host-side factorisation provider for LARGE symmetric positive definite matrices (3D pencils: BASELINE.json configs[2],
 *  A + shift B on 128^3 nodes): supernodal multifrontal Cholesky P A P^T = L L^T with a nested-dissection ordering.
 *
 *  Why a second provider next to sparse_lu.hh: the reference obtains its factors from UMFPACK (umfpacktools.hh:46-199),
 *  which is not available here; sparse_lu.hh is a scalar up-looking LU -- fine for the 2D configuration (n = 4 * 10^4) but
 *  O(flops) scalar work and explicit L and U with 8-byte indices (16 bytes per factor entry, twice) make a 3D factor of
 *  10^9 entries impossible. Here the factor is kept in SUPERNODAL form -- for every supernode one dense column-major block
 *  (rows x columns of the supernode) and one row-index list -- which is 8 bytes per entry, is what the dense kernels of the
 *  factorisation produce, and is exactly what the device apply wants (csrc/kernels_snode.cuh: dense panels on the FP64
 *  tensor pipe instead of scalar level schedules). The factorisation itself is one-time host setup (north_star).
 *
 *  Algorithm (all standard; own implementation):
 *    ordering      METIS nested dissection (compute_ordering, sparse_lu.hh) composed with a postorder of the elimination tree
 *    symbolic      elimination tree (Liu), column counts in O(nnz) (Gilbert, Ng, Peyton 1994: skeleton + least common
 *                  ancestors with path compression), fundamental supernodes + relaxed amalgamation of small children,
 *                  row structure of every supernode by merging the children's update rows
 *    numeric       multifrontal: assemble the frontal matrix (entries of A + children's update matrices, extend-add),
 *                  dense partial Cholesky (blocked right-looking: POTRF of the pivot block, TRSM of the rows below, SYRK
 *                  into the update matrix) with a register-blocked AVX2 kernel; independent subtrees run on different
 *                  threads, the fronts above them use all threads inside the dense kernels
 *  A non-positive pivot throws std::invalid_argument (the matrix is not positive definite: use sparse_lu.hh).
 *
 *  supernodal_to_contract() expands a (small) factor into the UMFPACK field contract of the reference
 *  (L unit lower triangular by rows, U = D L^T by columns, P = Q, Rs = 1), so the reference's own apply
 *  (kernels_cpp.hh:660-755) can run on the same factorisation in the parity tests.
 */

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
#include <stdexcept>
#include <thread>
#include <utility>
#include <vector>

#include "sparse_lu.hh"

namespace de_b200
{
  /** allocator whose default construction leaves the memory as it is: resize() of the factor's value array (29 GB for the
   *  128^3 pencil) must not zero-fill it on one thread -- the numeric phase writes every entry, in parallel */
  template <class T>
  struct default_init_allocator : std::allocator<T>
  {
    template <class U>
    struct rebind
    {
      using other = default_init_allocator<U>;
    };
    default_init_allocator() = default;
    template <class U>
    default_init_allocator(const default_init_allocator<U> &) noexcept {}
    template <class U, class... Args>
    void construct(U *p, Args &&...args)
    {
      if constexpr (sizeof...(Args) == 0)
        ::new ((void *)p) U;
      else
        ::new ((void *)p) U(std::forward<Args>(args)...);
    }
  };

  struct SupernodalFactor
  {
    long n = 0, nsuper = 0;
    long lnz = 0;      // entries of L (lower triangle incl. diagonal)
    double flops = 0.0; // of the numeric factorisation
    double seconds_ordering = 0.0, seconds_symbolic = 0.0, seconds_numeric = 0.0;
    std::vector<long> perm;   // perm[k] = original index of pivot k
    std::vector<long> sfirst; // [nsuper + 1] first column of supernode s
    std::vector<long> rowptr; // [nsuper + 1] offsets into rowidx
    std::vector<int> rowidx;  // rows of supernode s: its own columns first, then the update rows, ascending
    std::vector<long> valptr; // [nsuper + 1] offsets into val
    std::vector<double, default_init_allocator<double>> val; // block of supernode s: rows(s) x cols(s), column-major, leading dimension rows(s);
                              // the strict upper triangle of the diagonal block is zero
    std::vector<long> sparent; // supernodal elimination tree (-1: root)
    long rows(long s) const { return rowptr[s + 1] - rowptr[s]; }
    long cols(long s) const { return sfirst[s + 1] - sfirst[s]; }
  };

  namespace sn_detail
  {
    using I = long;

    /** a small persistent thread pool: parallel_for(n, f) runs f(i) for i < n on all threads (dynamic distribution) */
    class Pool
    {
      std::vector<std::thread> th_;
      std::mutex mu_;
      std::condition_variable cv_, done_;
      std::function<void(long)> fn_;
      std::atomic<long> next_{0};
      long count_ = 0, active_ = 0;
      unsigned long generation_ = 0;
      bool stop_ = false;

      void worker()
      {
        unsigned long seen = 0;
        for (;;)
        {
          {
            std::unique_lock<std::mutex> lock(mu_);
            cv_.wait(lock, [&] { return stop_ || generation_ != seen; });
            if (stop_)
              return;
            seen = generation_;
          }
          run();
          {
            std::lock_guard<std::mutex> lock(mu_);
            if (--active_ == 0)
              done_.notify_all();
          }
        }
      }
      void run()
      {
        for (long i = next_.fetch_add(1); i < count_; i = next_.fetch_add(1))
          fn_(i);
      }

    public:
      explicit Pool(int nthreads)
      {
        for (int t = 1; t < nthreads; ++t)
          th_.emplace_back([this] { worker(); });
      }
      ~Pool()
      {
        {
          std::lock_guard<std::mutex> lock(mu_);
          stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : th_)
          t.join();
      }
      int size() const { return (int)th_.size() + 1; }
      void parallel_for(long n, const std::function<void(long)> &f)
      {
        if (n <= 0)
          return;
        if (n == 1 || th_.empty())
        {
          for (long i = 0; i < n; ++i)
            f(i);
          return;
        }
        {
          std::lock_guard<std::mutex> lock(mu_);
          fn_ = f;
          count_ = n;
          next_ = 0;
          active_ = (long)th_.size();
          ++generation_;
        }
        cv_.notify_all();
        run();
        std::unique_lock<std::mutex> lock(mu_);
        done_.wait(lock, [&] { return active_ == 0; });
      }
    };

    typedef double v4d __attribute__((vector_size(32), aligned(8)));
    typedef double v8d __attribute__((vector_size(64), aligned(8)));

#if defined(__x86_64__) && defined(__GNUC__)
#define DE_B200_HAVE_AVX512_DISPATCH 1
    /** one strip of 4 columns of gemm_nt_sub on AVX-512 (chosen at run time): 24 x 4 register tiles -- 12 accumulators of 8
     *  doubles, 3 loads of A and 4 broadcasts of B per 12 FMAs. Returns the number of rows it handled (a multiple of 24).
     *  Measured on the pool's Xeons (one core, k = 48): 26-30 GFLOP/s against 15-16 for the 8 x 4 AVX2 tile. */
    __attribute__((target("avx512f"))) inline long gemm_nt_sub_strip_avx512(double *C, long ldc, const double *A, long lda,
                                                                            const double *B, long ldb, long m, long k)
    {
      long i = 0;
      for (; i + 24 <= m; i += 24)
      {
        v8d c[4][3];
        for (int q = 0; q < 4; ++q)
          for (int r = 0; r < 3; ++r)
            std::memcpy(&c[q][r], C + i + 8 * r + q * ldc, 64);
        for (long p = 0; p < k; ++p)
        {
          v8d a0, a1, a2;
          std::memcpy(&a0, A + i + p * lda, 64);
          std::memcpy(&a1, A + i + 8 + p * lda, 64);
          std::memcpy(&a2, A + i + 16 + p * lda, 64);
          for (int q = 0; q < 4; ++q)
          {
            const double b = B[q + p * ldb];
            c[q][0] -= a0 * b;
            c[q][1] -= a1 * b;
            c[q][2] -= a2 * b;
          }
        }
        for (int q = 0; q < 4; ++q)
          for (int r = 0; r < 3; ++r)
            std::memcpy(C + i + 8 * r + q * ldc, &c[q][r], 64);
      }
      return i;
    }
    inline bool cpu_has_avx512()
    {
      static const bool has = __builtin_cpu_supports("avx512f");
      return has;
    }
#endif

    /** C(m x n) -= A(m x k) B(n x k)^T, all column-major; lower_only: skip the part of C strictly above its diagonal
     *  (C square, m == n). The work horse of the factorisation: 8 x 4 register tile, broadcast of B, FMA on 4-wide
     *  vectors (AVX2 with -march=x86-64-v3; any other target compiles to what it has). */
    inline void gemm_nt_sub(double *C, long ldc, const double *A, long lda, const double *B, long ldb, long m, long n, long k)
    {
      long j = 0;
#ifdef DE_B200_HAVE_AVX512_DISPATCH
      const bool wide = m >= 24 && cpu_has_avx512();
#endif
      for (; j + 4 <= n; j += 4)
      {
        long i = 0;
#ifdef DE_B200_HAVE_AVX512_DISPATCH
        if (wide)
          i = gemm_nt_sub_strip_avx512(C + j * ldc, ldc, A, lda, B + j, ldb, m, k);
#endif
        for (; i + 8 <= m; i += 8)
        {
          v4d c00, c01, c10, c11, c20, c21, c30, c31;
          std::memcpy(&c00, C + i + (j + 0) * ldc, 32);
          std::memcpy(&c01, C + i + 4 + (j + 0) * ldc, 32);
          std::memcpy(&c10, C + i + (j + 1) * ldc, 32);
          std::memcpy(&c11, C + i + 4 + (j + 1) * ldc, 32);
          std::memcpy(&c20, C + i + (j + 2) * ldc, 32);
          std::memcpy(&c21, C + i + 4 + (j + 2) * ldc, 32);
          std::memcpy(&c30, C + i + (j + 3) * ldc, 32);
          std::memcpy(&c31, C + i + 4 + (j + 3) * ldc, 32);
          for (long p = 0; p < k; ++p)
          {
            v4d a0, a1;
            std::memcpy(&a0, A + i + p * lda, 32);
            std::memcpy(&a1, A + i + 4 + p * lda, 32);
            const double b0 = B[j + p * ldb], b1 = B[j + 1 + p * ldb], b2 = B[j + 2 + p * ldb], b3 = B[j + 3 + p * ldb];
            c00 -= a0 * b0;
            c01 -= a1 * b0;
            c10 -= a0 * b1;
            c11 -= a1 * b1;
            c20 -= a0 * b2;
            c21 -= a1 * b2;
            c30 -= a0 * b3;
            c31 -= a1 * b3;
          }
          std::memcpy(C + i + (j + 0) * ldc, &c00, 32);
          std::memcpy(C + i + 4 + (j + 0) * ldc, &c01, 32);
          std::memcpy(C + i + (j + 1) * ldc, &c10, 32);
          std::memcpy(C + i + 4 + (j + 1) * ldc, &c11, 32);
          std::memcpy(C + i + (j + 2) * ldc, &c20, 32);
          std::memcpy(C + i + 4 + (j + 2) * ldc, &c21, 32);
          std::memcpy(C + i + (j + 3) * ldc, &c30, 32);
          std::memcpy(C + i + 4 + (j + 3) * ldc, &c31, 32);
        }
        for (; i < m; ++i)
          for (long q = 0; q < 4; ++q)
          {
            double s = C[i + (j + q) * ldc];
            for (long p = 0; p < k; ++p)
              s -= A[i + p * lda] * B[j + q + p * ldb];
            C[i + (j + q) * ldc] = s;
          }
      }
      for (; j < n; ++j)
        for (long i = 0; i < m; ++i)
        {
          double s = C[i + j * ldc];
          for (long p = 0; p < k; ++p)
            s -= A[i + p * lda] * B[j + p * ldb];
          C[i + j * ldc] = s;
        }
    }

    /** unblocked Cholesky of the leading w x w block of a column-major panel with `rows` rows (ld): the rows below the
     *  block are solved against it as well (L21 = A21 L11^-T). Returns false on a non-positive pivot. */
    inline bool panel_factor(double *P, long ld, long rows, long w)
    {
      for (long j = 0; j < w; ++j)
      {
        double *cj = P + j * ld;
        for (long p = 0; p < j; ++p)
        {
          const double l = P[j + p * ld];
          if (l != 0.0)
          {
            const double *cp = P + p * ld;
            for (long i = j; i < rows; ++i)
              cj[i] -= cp[i] * l;
          }
        }
        const double d = cj[j];
        if (!(d > 0.0) || !std::isfinite(d))
          return false;
        const double r = std::sqrt(d), inv = 1.0 / r;
        cj[j] = r;
        for (long i = j + 1; i < rows; ++i)
          cj[i] *= inv;
      }
      return true;
    }

    constexpr long kPanel = 96;  // pivot columns per panel of the blocked factorisation
    constexpr long kTile = 192;  // rows / columns per task of a trailing update

    /** dense partial Cholesky of a front: Fr = r x r column-major (ld r), lower triangle assembled; the first ns columns are
     *  eliminated (blocked right-looking: panel factorisation, then the rank-w update of the trailing lower triangle in
     *  tiles), the trailing (r - ns)^2 block is left holding the update matrix. Pool may be null. */
    struct FactorTrace
    {
      bool on = std::getenv("DE_TRACE_FACTOR") != nullptr; // debugging aid: prints where the numeric phase spends its time
      double t_panel = 0, t_trailing = 0, t_assemble = 0, t_copyout = 0, t_subtrees = 0;
    };
    inline FactorTrace &factor_trace()
    {
      static FactorTrace t;
      return t;
    }
    inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

    inline bool front_factor(double *Fr, long r, long ns, Pool *pool)
    {
      FactorTrace &tr = factor_trace();
      const bool trace = tr.on && pool != nullptr;
      for (long k0 = 0; k0 < ns; k0 += kPanel)
      {
        const double tp0 = trace ? now_s() : 0.0;
        const long w = std::min(kPanel, ns - k0);
        double *P = Fr + k0 + k0 * r; // panel: rows k0.., columns k0..k0+w
        const long prow = r - k0;
        if (prow * w > 20000 && pool)
        {
          // factor the w x w block, then solve the rows below it in parallel chunks
          if (!panel_factor(P, r, w, w))
            return false;
          const long below = prow - w, chunk = 256, nch = (below + chunk - 1) / chunk;
          pool->parallel_for(nch, [&](long c) {
            const long i0 = w + c * chunk, i1 = std::min(prow, i0 + chunk);
            for (long j = 0; j < w; ++j)
            {
              double *cj = P + j * r;
              for (long p = 0; p < j; ++p)
              {
                const double l = P[j + p * r];
                const double *cp = P + p * r;
                for (long i = i0; i < i1; ++i)
                  cj[i] -= cp[i] * l;
              }
              const double inv = 1.0 / cj[j];
              for (long i = i0; i < i1; ++i)
                cj[i] *= inv;
            }
          });
        }
        else if (!panel_factor(P, r, prow, w))
          return false;
        const double tp1 = trace ? now_s() : 0.0;
        if (trace)
          tr.t_panel += tp1 - tp0;
        struct TrailTimer
        {
          FactorTrace &t;
          bool on;
          double t0;
          ~TrailTimer()
          {
            if (on)
              t.t_trailing += now_s() - t0;
          }
        } trail_timer{tr, trace, tp1};
        const long t0 = k0 + w, nt = r - t0; // trailing part of the front
        if (nt <= 0)
          continue;
        const double *Pan = Fr + t0 + k0 * r; // rows t0.. of the panel
        double *T = Fr + t0 + t0 * r;         // trailing block
        const long nb = (nt + kTile - 1) / kTile;
        auto tile = [&](long ti, long tj) {
          const long i0 = ti * kTile, i1 = std::min(nt, i0 + kTile), j0 = tj * kTile, j1 = std::min(nt, j0 + kTile);
          const long mi = i1 - i0, nj = j1 - j0;
          // the two panel slices of the tile, packed: in the front they are w columns a whole front column apart (one
          // page each for a large front), which the k loop of the kernel walks through for every 24 x 4 register tile
          thread_local std::vector<double> pa, pb;
          pa.resize((size_t)mi * w);
          for (long p = 0; p < w; ++p)
            std::memcpy(pa.data() + p * mi, Pan + i0 + p * r, sizeof(double) * (size_t)mi);
          if (ti != tj)
          {
            pb.resize((size_t)nj * w);
            for (long p = 0; p < w; ++p)
              std::memcpy(pb.data() + p * nj, Pan + j0 + p * r, sizeof(double) * (size_t)nj);
            gemm_nt_sub(T + i0 + j0 * r, r, pa.data(), mi, pb.data(), nj, mi, nj, w);
          }
          else
            for (long jj = j0; jj < j1; jj += 4) // lower part of a diagonal tile in strips of 4 columns (the few entries
            {                                      // above the diagonal inside a strip are never read)
              const long wj = std::min<long>(4, j1 - jj);
              gemm_nt_sub(T + jj + jj * r, r, pa.data() + (jj - i0), mi, pa.data() + (jj - i0), mi, i1 - jj, wj, w);
            }
        };
        const long ntasks = nb * (nb + 1) / 2;
        if (pool && (double)nt * nt * w > 2.0e6)
          pool->parallel_for(ntasks, [&](long t) {
            long ti = (long)((std::sqrt(8.0 * (double)t + 1.0) - 1.0) / 2.0);
            while (ti * (ti + 1) / 2 > t)
              --ti;
            while ((ti + 1) * (ti + 2) / 2 <= t)
              ++ti;
            tile(ti, t - ti * (ti + 1) / 2);
          });
        else
          for (long ti = 0; ti < nb; ++ti)
            for (long tj = 0; tj <= ti; ++tj)
              tile(ti, tj);
      }
      return true;
    }
  } // namespace sn_detail

  /** Supernodal multifrontal Cholesky of the symmetric positive definite CSR matrix (full pattern; only entries with
   *  col <= row of the permuted matrix are read). nthreads <= 0: all hardware threads. */
  template <class Int>
  inline void supernodal_cholesky(long n, const Int *rowptr, const Int *col, const double *val, Ordering ordering, int nthreads,
                                  SupernodalFactor &F)
  {
    using namespace sn_detail;
    if (n >= (1L << 31) - 1)
      throw std::invalid_argument("supernodal_cholesky: more than 2^31 rows");
    if (nthreads <= 0)
      nthreads = (int)std::max(1u, std::thread::hardware_concurrency());
    F = SupernodalFactor();
    F.n = n;
    const auto t_start = std::chrono::steady_clock::now();
    auto since = [](std::chrono::steady_clock::time_point t) {
      return std::chrono::duration<double>(std::chrono::steady_clock::now() - t).count();
    };
    std::vector<I> perm = compute_ordering(n, rowptr, col, ordering);
    const double t_order = since(t_start);
    const auto t_sym = std::chrono::steady_clock::now();
    auto t_lap = t_sym;
    auto lap = [&](const char *what) { // DE_TRACE_FACTOR: where the symbolic phase spends its time
      if (sn_detail::factor_trace().on)
      {
        std::fprintf(stderr, "[de factor] symbolic: %-28s %.2f s\n", what, since(t_lap));
        t_lap = std::chrono::steady_clock::now();
      }
    };

    // ---- lower part of the permuted matrix by rows: row k holds (column i < k, value), sorted; diagonal apart ------------
    std::vector<I> iperm(n), Bp, Bj;
    std::vector<double> Bx, diag;
    auto permute = [&]() {
      for (I k = 0; k < n; ++k)
        iperm[perm[k]] = k;
      Bp.assign(n + 1, 0);
      for (I k = 0; k < n; ++k)
      {
        const I i = perm[k];
        I c = 0;
        for (Int q = rowptr[i]; q < rowptr[i + 1]; ++q)
          c += iperm[col[q]] < k;
        Bp[k + 1] = Bp[k] + c;
      }
      Bj.resize(Bp[n]);
      Bx.resize(Bp[n]);
      diag.assign(n, 0.0);
      std::vector<std::pair<I, double>> rowbuf;
      for (I k = 0; k < n; ++k)
      {
        const I i = perm[k];
        rowbuf.clear();
        for (Int q = rowptr[i]; q < rowptr[i + 1]; ++q)
        {
          const I c = iperm[col[q]];
          if (c < k)
            rowbuf.emplace_back(c, val[q]);
          else if (c == k)
            diag[k] += val[q];
        }
        std::sort(rowbuf.begin(), rowbuf.end());
        I w = Bp[k];
        for (auto &e : rowbuf)
        {
          Bj[w] = e.first;
          Bx[w++] = e.second;
        }
      }
    };
    auto etree = [&](std::vector<I> &parent) {
      parent.assign(n, -1);
      std::vector<I> anc(n, -1);
      for (I k = 0; k < n; ++k)
        for (I q = Bp[k]; q < Bp[k + 1]; ++q)
        {
          I i = Bj[q];
          while (i != -1 && i < k)
          {
            const I nxt = anc[i];
            anc[i] = k;
            if (nxt == -1)
              parent[i] = k;
            i = nxt;
          }
        }
    };
    std::vector<I> parent;
    permute();
    etree(parent);
    {
      // postorder of the elimination tree (children before parents, subtrees contiguous), composed into the permutation
      std::vector<I> head(n, -1), next(n, -1), post;
      post.reserve(n);
      for (I j = n - 1; j >= 0; --j)
        if (parent[j] != -1)
        {
          next[j] = head[parent[j]];
          head[parent[j]] = j;
        }
      std::vector<I> stack;
      for (I root = 0; root < n; ++root)
      {
        if (parent[root] != -1)
          continue;
        stack.push_back(root);
        while (!stack.empty())
        {
          const I j = stack.back(), c = head[j];
          if (c == -1)
          {
            post.push_back(j);
            stack.pop_back();
          }
          else
          {
            head[j] = next[c];
            stack.push_back(c);
          }
        }
      }
      std::vector<I> np(n);
      for (I k = 0; k < n; ++k)
        np[k] = perm[post[k]];
      perm.swap(np);
      permute();
      etree(parent);
    }

    lap("permute + etree + postorder");
    // ---- column counts (Gilbert / Ng / Peyton): the tree is postordered, so post[k] = k ---------------------------------
    std::vector<I> cc(n, 0);
    {
      // upper part by rows = lower part by columns: for column j the rows i > j with an entry
      std::vector<I> Tp(n + 1, 0);
      for (I k = 0; k < n; ++k)
        for (I q = Bp[k]; q < Bp[k + 1]; ++q)
          Tp[Bj[q] + 1]++;
      for (I k = 0; k < n; ++k)
        Tp[k + 1] += Tp[k];
      std::vector<I> Ti(Tp[n]), w(Tp.begin(), Tp.end() - 1);
      for (I k = 0; k < n; ++k)
        for (I q = Bp[k]; q < Bp[k + 1]; ++q)
          Ti[w[Bj[q]]++] = k;
      std::vector<I> anc(n), maxfirst(n, -1), prevleaf(n, -1), first(n, -1);
      for (I k = 0; k < n; ++k)
      {
        I j = k;
        cc[j] = (first[j] == -1) ? 1 : 0; // leaf of the tree
        for (; j != -1 && first[j] == -1; j = parent[j])
          first[j] = k;
      }
      std::iota(anc.begin(), anc.end(), I(0));
      for (I j = 0; j < n; ++j)
      {
        if (parent[j] != -1)
          cc[parent[j]]--;
        for (I q = Tp[j]; q < Tp[j + 1]; ++q)
        {
          const I i = Ti[q]; // i > j, A(i, j) != 0
          if (first[j] <= maxfirst[i])
            continue; // j is not a leaf of the row subtree of i
          maxfirst[i] = first[j];
          const I jprev = prevleaf[i];
          prevleaf[i] = j;
          cc[j]++; // (i, j) is in the skeleton
          if (jprev != -1)
          {
            I qa = jprev;
            while (qa != anc[qa])
              qa = anc[qa];
            for (I s = jprev; s != qa;)
            {
              const I sp = anc[s];
              anc[s] = qa;
              s = sp;
            }
            cc[qa]--; // overlap at the least common ancestor
          }
        }
        if (parent[j] != -1)
          anc[j] = parent[j];
      }
      for (I j = 0; j < n; ++j)
        if (parent[j] != -1)
          cc[parent[j]] += cc[j];
    }

    lap("column counts");
    // ---- supernodes: fundamental, then relaxed amalgamation of a last child into its parent ------------------------------
    std::vector<I> nchild(n, 0);
    for (I j = 0; j < n; ++j)
      if (parent[j] != -1)
        nchild[parent[j]]++;
    std::vector<I> sfirst;
    sfirst.push_back(0);
    for (I j = 1; j < n; ++j)
    {
      const bool same = parent[j - 1] == j && cc[j] == cc[j - 1] - 1 && nchild[j] == 1;
      if (!same)
        sfirst.push_back(j);
    }
    sfirst.push_back(n);
    {
      // relaxed amalgamation: merge supernode s into the FOLLOWING supernode t when t is its parent (s is t's last child,
      // so the columns stay contiguous) and the explicit zeros this adds stay below a size-dependent fraction
      std::vector<I> merged;
      merged.push_back(0);
      I cur_first = 0;
      const I ns0 = (I)sfirst.size() - 1;
      for (I s = 0; s < ns0; ++s)
      {
        const I l = sfirst[s + 1] - 1; // last column of this fundamental supernode
        bool merge = false;
        if (s + 1 < ns0 && parent[l] == l + 1)
        {
          // columns cur_first..l would get the structure of column l+1 plus their own position: extra zeros per column
          const I wcur = l + 1 - cur_first, wnext = sfirst[s + 2] - sfirst[s + 1];
          const I rows_next = cc[l + 1]; // entries of column l + 1 (diagonal included)
          I zeros = 0;                   // explicit zeros of the merged block: new column length minus true length
          for (I c = cur_first; c <= l; ++c)
            zeros += (rows_next + (l + 1 - c)) - cc[c];
          const I wtot = wcur + wnext;
          const I total = wtot * rows_next + wcur * (wcur + 1) / 2 + wcur * (wnext - 1);
          const double frac = total > 0 ? (double)zeros / (double)total : 1.0;
          merge = (wtot <= 4) || (wtot <= 16 && frac < 0.8) || (wtot <= 48 && frac < 0.2) || (frac < 0.02);
        }
        if (!merge)
        {
          merged.push_back(sfirst[s + 1]);
          cur_first = sfirst[s + 1];
        }
      }
      if (merged.back() != n)
        merged.push_back(n);
      sfirst.swap(merged);
    }
    const I nsuper = (I)sfirst.size() - 1;
    std::vector<I> snode_of(n);
    for (I s = 0; s < nsuper; ++s)
      for (I c = sfirst[s]; c < sfirst[s + 1]; ++c)
        snode_of[c] = s;
    std::vector<I> sparent(nsuper, -1);
    for (I s = 0; s < nsuper; ++s)
    {
      const I p = parent[sfirst[s + 1] - 1];
      sparent[s] = p == -1 ? -1 : snode_of[p];
    }

    lap("supernodes");
    // ---- row structure of every supernode (own columns, then update rows ascending) -------------------------------------
    std::vector<I> rptr(nsuper + 1, 0);
    std::vector<std::vector<int>> upd_rows(nsuper);
    {
      // lower part by columns again (Tp / Ti), for the entries of A in the supernode's columns
      std::vector<I> Tp(n + 1, 0);
      for (I k = 0; k < n; ++k)
        for (I q = Bp[k]; q < Bp[k + 1]; ++q)
          Tp[Bj[q] + 1]++;
      for (I k = 0; k < n; ++k)
        Tp[k + 1] += Tp[k];
      std::vector<I> Ti(Tp[n]), w(Tp.begin(), Tp.end() - 1);
      for (I k = 0; k < n; ++k)
        for (I q = Bp[k]; q < Bp[k + 1]; ++q)
          Ti[w[Bj[q]]++] = k;
      std::vector<I> mark(n, -1);
      std::vector<std::vector<I>> children(nsuper);
      for (I s = 0; s < nsuper; ++s)
        if (sparent[s] != -1)
          children[sparent[s]].push_back(s);
      for (I s = 0; s < nsuper; ++s)
      {
        const I f = sfirst[s], l = sfirst[s + 1] - 1;
        std::vector<int> &R = upd_rows[s];
        for (I c = f; c <= l; ++c)
          for (I q = Tp[c]; q < Tp[c + 1]; ++q)
          {
            const I i = Ti[q];
            if (i > l && mark[i] != s)
            {
              mark[i] = s;
              R.push_back((int)i);
            }
          }
        for (I c : children[s])
          for (int i : upd_rows[c])
            if (i > l && mark[i] != s)
            {
              mark[i] = s;
              R.push_back(i);
            }
        std::sort(R.begin(), R.end());
      }
    }
    lap("row structure");
    F.nsuper = nsuper;
    F.perm = perm;
    F.sfirst = sfirst;
    F.sparent = sparent;
    F.rowptr.assign(nsuper + 1, 0);
    F.valptr.assign(nsuper + 1, 0);
    for (I s = 0; s < nsuper; ++s)
    {
      const I ns = sfirst[s + 1] - sfirst[s], r = ns + (I)upd_rows[s].size();
      F.rowptr[s + 1] = F.rowptr[s] + r;
      F.valptr[s + 1] = F.valptr[s] + r * ns;
      F.lnz += r * ns - ns * (ns - 1) / 2;
      for (I j = 0; j < ns; ++j)
      {
        const double len = (double)(r - j);
        F.flops += len * len; // column j: scale + rank-1 update of the len x len trailing part (lower half, 2 flops)
      }
    }
    F.rowidx.resize(F.rowptr[nsuper]);
    for (I s = 0; s < nsuper; ++s)
    {
      int *R = F.rowidx.data() + F.rowptr[s];
      const I ns = sfirst[s + 1] - sfirst[s];
      for (I j = 0; j < ns; ++j)
        R[j] = (int)(sfirst[s] + j);
      std::copy(upd_rows[s].begin(), upd_rows[s].end(), R + ns);
      std::vector<int>().swap(upd_rows[s]);
    }
    F.val.resize(F.valptr[nsuper]); // NOT zero-filled (default_init_allocator): every entry is written by the numeric phase

    lap("index lists + factor storage");
    const double t_symbolic = since(t_sym);
    const auto t_num = std::chrono::steady_clock::now();
    // ---- numeric: subtrees on different threads, the fronts above them with all threads inside the dense kernels ---------
    std::vector<double> work(nsuper, 0.0); // flops of the subtree rooted at s
    for (I s = 0; s < nsuper; ++s)
    {
      const I ns = F.cols(s), r = F.rows(s);
      double fl = 0.0;
      for (I j = 0; j < ns; ++j)
        fl += (double)(r - j) * (double)(r - j);
      work[s] += fl;
      if (sparent[s] != -1)
        work[sparent[s]] += work[s];
    }
    const double total_work = std::accumulate(work.begin(), work.end(), 0.0, [](double a, double b) { return std::max(a, b); });
    // a supernode belongs to the "top" if its subtree holds more than 1 / (8 nthreads) of the work; the maximal subtrees
    // below the top are the independent tasks
    const double cut = nthreads > 1 ? total_work / (8.0 * nthreads) : 2.0 * total_work;
    std::vector<char> top(nsuper, 0);
    for (I s = 0; s < nsuper; ++s)
      top[s] = work[s] > cut;
    std::vector<I> task_root;
    for (I s = 0; s < nsuper; ++s)
      if (!top[s] && (sparent[s] == -1 || top[sparent[s]]))
        task_root.push_back(s);
    // first supernode of the subtree of s (postorder: the subtree is the contiguous range [sub_first[s], s])
    std::vector<I> sub_first(nsuper);
    for (I s = 0; s < nsuper; ++s)
      sub_first[s] = s;
    for (I s = 0; s < nsuper; ++s)
      if (sparent[s] != -1)
        sub_first[sparent[s]] = std::min(sub_first[sparent[s]], sub_first[s]);

    // update matrices (nu x nu, column-major, only the lower triangle is written and read), alive until the parent is
    // assembled; uninitialised storage: value-initialising 2 GB on one thread cost more than the copy into it
    std::vector<std::unique_ptr<double[]>> upd(nsuper);
    std::unique_ptr<double[]> top_front; // the frontal matrix of the "top" supernodes: one grow-only buffer (no page faults per front)
    size_t top_front_cap = 0;
    std::vector<std::vector<I>> children(nsuper);
    for (I s = 0; s < nsuper; ++s)
      if (sparent[s] != -1)
        children[sparent[s]].push_back(s);
    std::atomic<int> failed{0};
    std::atomic<long> failed_col{-1};

    auto process = [&](I s, std::vector<int> &loc, Pool *pool) {
      const double t_proc0 = sn_detail::now_s();
      const I f = sfirst[s], ns = F.cols(s), r = F.rows(s), nu = r - ns;
      const int *R = F.rowidx.data() + F.rowptr[s];
      // the frontal matrix, column-major, lower triangle used. Large fronts (pool != nullptr: the top of the tree, processed
      // one after the other) live in one reused buffer and are cleared, assembled and copied out by all threads: on the
      // 64^3 pencil those serial O(r^2) steps were 39 % of the numeric phase.
      const bool par = pool != nullptr && (double)r * (double)r > 1.0e6;
      std::vector<double> small_front;
      double *Fr;
      if (pool != nullptr)
      {
        if (top_front_cap < (size_t)r * r)
        {
          top_front.reset();
          top_front.reset(new double[(size_t)r * r]);
          top_front_cap = (size_t)r * r;
        }
        Fr = top_front.get();
        const I cchunk = 64, nch = (r + cchunk - 1) / cchunk;
        auto clear = [&](long c) {
          for (I j = c * cchunk; j < std::min<I>(r, (c + 1) * cchunk); ++j)
          {
            // lower triangle, plus the three entries above the diagonal that the 4-column strips of the diagonal tiles
            // update (and never read): they must not hold NaN patterns from an earlier front
            const I i0 = std::max<I>(0, j - 3);
            std::memset(Fr + i0 + j * r, 0, sizeof(double) * (size_t)(r - i0));
          }
        };
        if (par)
          pool->parallel_for(nch, clear);
        else
          for (I c = 0; c < nch; ++c)
            clear(c);
      }
      else
      {
        small_front.assign((size_t)r * r, 0.0);
        Fr = small_front.data();
      }
      for (I a = 0; a < r; ++a)
        loc[R[a]] = (int)a;
      // entries of A: columns f .. f + ns - 1, rows in the front (diagonal kept apart)
      for (I j = 0; j < ns; ++j)
        Fr[j + j * r] = diag[f + j];
      for (I a = 0; a < r; ++a)
      {
        const I k = R[a];
        const I *b = Bj.data() + Bp[k], *e = Bj.data() + Bp[k + 1];
        for (const I *p = std::lower_bound(b, e, f); p < e && *p < f + ns; ++p)
          Fr[a + (*p - f) * r] += Bx[p - Bj.data()];
      }
      // extend-add of the children's update matrices
      for (I c : children[s])
      {
        const I nsc = F.cols(c), nuc = F.rows(c) - nsc;
        const int *Rc = F.rowidx.data() + F.rowptr[c] + nsc;
        const double *Uc = upd[c].get();
        auto add_cols = [&](I b0, I b1) { // distinct columns b go to distinct columns of the front
          for (I b = b0; b < b1; ++b)
          {
            double *dst = Fr + (I)loc[Rc[b]] * r;
            const double *src = Uc + b * nuc;
            for (I a = b; a < nuc; ++a)
              dst[loc[Rc[a]]] += src[a];
          }
        };
        if (par && (double)nuc * (double)nuc > 1.0e6)
        {
          // column b carries nuc - b entries: chunks of equal work, not of equal width
          const I nch = 4 * pool->size();
          pool->parallel_for(nch, [&](long q) {
            const double total = 0.5 * (double)nuc * (double)nuc;
            auto col_at = [&](double w) { return (I)((double)nuc - std::sqrt(std::max(0.0, (double)nuc * (double)nuc - 2.0 * w))); };
            const I b0 = q == 0 ? 0 : col_at(total * (double)q / (double)nch);
            const I b1 = q == nch - 1 ? nuc : col_at(total * (double)(q + 1) / (double)nch);
            add_cols(std::min(b0, nuc), std::min(std::max(b1, b0), nuc));
          });
        }
        else
          add_cols(0, nuc);
        upd[c].reset();
      }
      sn_detail::FactorTrace &tr = sn_detail::factor_trace();
      const bool trace = tr.on && pool != nullptr;
      if (trace)
        tr.t_assemble += sn_detail::now_s() - t_proc0;
      if (!front_factor(Fr, r, ns, pool))
      {
        failed = 1;
        failed_col = f;
        return;
      }
      const double t_co0 = trace ? sn_detail::now_s() : 0.0;
      double *Lb = F.val.data() + F.valptr[s]; // r x ns; the strict upper triangle of the pivot block stays zero
      upd[s].reset(nu > 0 ? new double[(size_t)nu * nu] : nullptr);
      double *U = upd[s].get();
      auto copy_out = [&](long c) {
        const I cchunk = 64;
        for (I j = c * cchunk; j < std::min<I>(r, (c + 1) * cchunk); ++j)
        {
          if (j < ns)
          {
            std::memset(Lb + j * r, 0, sizeof(double) * (size_t)j); // the strict upper triangle of the pivot block is zero
            std::memcpy(Lb + j + j * r, Fr + j + j * r, sizeof(double) * (size_t)(r - j));
          }
          else
            std::memcpy(U + (j - ns) + (j - ns) * nu, Fr + j + j * r, sizeof(double) * (size_t)(r - j));
        }
      };
      {
        const I nch = (r + 63) / 64;
        if (par)
          pool->parallel_for(nch, copy_out);
        else
          for (I c = 0; c < nch; ++c)
            copy_out(c);
      }
      if (trace)
        tr.t_copyout += sn_detail::now_s() - t_co0;
    };

    {
      Pool pool(nthreads);
      std::vector<std::vector<int>> locs((size_t)pool.size());
      std::atomic<int> next_id{0};
      // subtree tasks, heaviest first
      std::sort(task_root.begin(), task_root.end(), [&](I a, I b) { return work[a] > work[b]; });
      std::mutex idmu;
      std::vector<std::thread::id> ids;
      pool.parallel_for((long)task_root.size(), [&](long t) {
        int me;
        {
          std::lock_guard<std::mutex> lock(idmu);
          const auto id = std::this_thread::get_id();
          me = (int)(std::find(ids.begin(), ids.end(), id) - ids.begin());
          if (me == (int)ids.size())
            ids.push_back(id);
        }
        std::vector<int> &loc = locs[me];
        if ((I)loc.size() != n)
          loc.assign(n, 0);
        const I root = task_root[t];
        for (I s = sub_first[root]; s <= root && !failed; ++s)
          process(s, loc, nullptr);
      });
      (void)next_id;
      if (sn_detail::factor_trace().on)
        sn_detail::factor_trace().t_subtrees = since(t_num);
      std::vector<int> &loc = locs[0];
      if ((I)loc.size() != n)
        loc.assign(n, 0);
      for (I s = 0; s < nsuper && !failed; ++s)
        if (top[s])
          process(s, loc, &pool);
    }
    F.seconds_ordering = t_order;
    F.seconds_symbolic = t_symbolic;
    F.seconds_numeric = since(t_num);
    if (sn_detail::factor_trace().on)
    {
      const sn_detail::FactorTrace &tr = sn_detail::factor_trace();
      std::fprintf(stderr, "[de factor] numeric %.2f s: subtrees %.2f | top fronts: assemble %.2f panel %.2f trailing %.2f copy-out %.2f\n",
                   F.seconds_numeric, tr.t_subtrees, tr.t_assemble, tr.t_panel, tr.t_trailing, tr.t_copyout);
    }
    if (failed)
      throw std::invalid_argument("supernodal_cholesky: matrix is not positive definite (pivot column " +
                                  std::to_string((long)failed_col) + " of the permuted matrix)");
  }

  /** x <- A^-1 x on the host for an n x m row-major block (test helper; the product's apply runs on the GPU) */
  inline void supernodal_solve_host(const SupernodalFactor &F, double *x, int m)
  {
    const long n = F.n;
    std::vector<double> y((size_t)n * m);
    for (long k = 0; k < n; ++k)
      std::memcpy(y.data() + (size_t)k * m, x + (size_t)F.perm[k] * m, sizeof(double) * m);
    for (long s = 0; s < F.nsuper; ++s) // forward: L z = y
    {
      const long ns = F.cols(s), r = F.rows(s);
      const int *R = F.rowidx.data() + F.rowptr[s];
      const double *L = F.val.data() + F.valptr[s];
      for (long j = 0; j < ns; ++j)
      {
        double *yj = y.data() + (size_t)R[j] * m;
        const double inv = 1.0 / L[j + j * r];
        for (int c = 0; c < m; ++c)
          yj[c] *= inv;
        for (long a = j + 1; a < r; ++a)
        {
          const double l = L[a + j * r];
          double *ya = y.data() + (size_t)R[a] * m;
          for (int c = 0; c < m; ++c)
            ya[c] -= l * yj[c];
        }
      }
    }
    for (long s = F.nsuper - 1; s >= 0; --s) // backward: L^T w = z
    {
      const long ns = F.cols(s), r = F.rows(s);
      const int *R = F.rowidx.data() + F.rowptr[s];
      const double *L = F.val.data() + F.valptr[s];
      for (long j = ns - 1; j >= 0; --j)
      {
        double *yj = y.data() + (size_t)R[j] * m;
        for (long a = j + 1; a < r; ++a)
        {
          const double l = L[a + j * r];
          const double *ya = y.data() + (size_t)R[a] * m;
          for (int c = 0; c < m; ++c)
            yj[c] -= l * ya[c];
        }
        const double inv = 1.0 / L[j + j * r];
        for (int c = 0; c < m; ++c)
          yj[c] *= inv;
      }
    }
    for (long k = 0; k < n; ++k)
      std::memcpy(x + (size_t)F.perm[k] * m, y.data() + (size_t)k * m, sizeof(double) * m);
  }

  /** Expand into the reference's UMFPACK field contract (umfpacktools.hh:26-44): P A P^T = (L D^-1)(D L^T) with
   *  D = diag(L): L_c unit lower triangular by rows (diagonal last), U_c = D L^T by columns (diagonal last), P = Q = perm,
   *  Rs = 1, do_recip = 1. Explicit zeros of the supernodal blocks are dropped. Only for factors of moderate size. */
  inline void supernodal_to_contract(const SupernodalFactor &S, FactorArrays &F)
  {
    using I = long;
    const I n = S.n;
    F = FactorArrays();
    F.n = F.n_row = F.n_col = n;
    F.P = S.perm;
    F.Q = S.perm;
    F.Rs.assign(n, 1.0);
    F.do_recip = 1;
    std::vector<double> d(n);
    // column-wise entries (row i > j, value L_ij) -> U_c column i holds (row j, d_j L_ij)?  U_c = D L^T: U_c(j, i) = d_j L(i, j)
    // L_c = L D^-1: L_c(i, j) = L(i, j) / d_j, stored by rows.
    std::vector<I> lcount(n, 0), ucount(n, 0);
    for (I s = 0; s < S.nsuper; ++s)
    {
      const I ns = S.cols(s), r = S.rows(s);
      const int *R = S.rowidx.data() + S.rowptr[s];
      const double *L = S.val.data() + S.valptr[s];
      for (I j = 0; j < ns; ++j)
      {
        d[R[j]] = L[j + j * r];
        for (I a = j + 1; a < r; ++a)
          if (L[a + j * r] != 0.0)
          {
            lcount[R[a]]++; // row R[a] of L_c gets column R[j]
            ucount[R[a]]++; // column R[a] of U_c gets row R[j]
          }
      }
    }
    F.Lp.assign(n + 1, 0);
    F.Up.assign(n + 1, 0);
    for (I i = 0; i < n; ++i)
    {
      F.Lp[i + 1] = F.Lp[i] + lcount[i] + 1;
      F.Up[i + 1] = F.Up[i] + ucount[i] + 1;
    }
    F.lnz = F.Lp[n];
    F.unz = F.Up[n];
    F.nz_udiag = n;
    F.Lj.resize(F.lnz);
    F.Lx.resize(F.lnz);
    F.Ui.resize(F.unz);
    F.Ux.resize(F.unz);
    std::vector<I> lw(F.Lp.begin(), F.Lp.end() - 1), uw(F.Up.begin(), F.Up.end() - 1);
    for (I s = 0; s < S.nsuper; ++s) // supernodes ascending, columns ascending: entries arrive with ascending column / row index
    {
      const I ns = S.cols(s), r = S.rows(s);
      const int *R = S.rowidx.data() + S.rowptr[s];
      const double *L = S.val.data() + S.valptr[s];
      for (I j = 0; j < ns; ++j)
        for (I a = j + 1; a < r; ++a)
        {
          const double l = L[a + j * r];
          if (l == 0.0)
            continue;
          const I i = R[a], c = R[j];
          F.Lj[lw[i]] = c;
          F.Lx[lw[i]++] = l / d[c];
          F.Ui[uw[i]] = c;
          F.Ux[uw[i]++] = d[c] * l;
        }
    }
    for (I i = 0; i < n; ++i)
    {
      F.Lj[lw[i]] = i;
      F.Lx[lw[i]] = 1.0;
      F.Ui[uw[i]] = i;
      F.Ux[uw[i]] = d[i] * d[i];
    }
  }
} // namespace de_b200

#endif
