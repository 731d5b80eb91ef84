#ifndef DUNE_EIGENSOLVER_B200_SPARSE_LU_HH
#define DUNE_EIGENSOLVER_B200_SPARSE_LU_HH

/** \file
 *  Host-side factorisation provider: fills the factor-array contract that the reference obtains
 *  from SuiteSparse UMFPACK (reference umfpacktools.hh:26-44, :170-186), for sites where UMFPACK is
 *  not available (this image has none). The factorisation is one-time setup and stays on the host
 *  (BASELINE.json north_star); only the *apply* (reference kernels_cpp.hh:660-755) is a GPU kernel.
 *
 *  Contract produced (identical to umfpack_dl_get_numeric as used by the reference):
 *    P A Q = L U  after row scaling, with
 *    L : n x n unit lower triangular, compressed ROW storage, columns ascending, diagonal stored LAST
 *    U : n x n upper triangular, compressed COLUMN storage, rows ascending, diagonal stored LAST
 *    P[k] = i : original row i is pivot row k ;  Q[k] = j : original column j is pivot column k
 *    Rs, do_recip : row i of A is multiplied by Rs[i] if do_recip else divided by Rs[i]
 *
 *  Algorithm: static symmetric fill-reducing ordering (P = Q) on the pattern of A + A^T, elimination
 *  tree + row-pattern symbolic analysis, then an up-looking (row by row) sparse LU WITHOUT numerical
 *  pivoting. That is sufficient for the pencils of the eigensolver path (symmetric, shifted to be
 *  positive definite); a vanishing pivot raises the same "input matrix is singular" error the
 *  reference raises (umfpacktools.hh:160-164). Values may be unsymmetric (row scaling makes them so).
 */

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <string>
#include <stdexcept>
#include <vector>

#ifdef DE_B200_HAVE_METIS
// libmetis_static.a shipped with the CUDA toolkit is built with 64-bit idx_t (probed in this image).
extern "C" int METIS_NodeND(int64_t *nvtxs, int64_t *xadj, int64_t *adjncy, int64_t *vwgt,
                            int64_t *options, int64_t *perm, int64_t *iperm);
extern "C" int METIS_SetDefaultOptions(int64_t *options);
#endif

namespace de_b200
{
  //! The UMFPACK-style factor arrays (same names and meaning as reference umfpacktools.hh:26-44)
  struct FactorArrays
  {
    using IntType = long;
    IntType n = 0, lnz = 0, unz = 0, n_row = 0, n_col = 0, nz_udiag = 0;
    std::vector<IntType> Lp, Lj, Up, Ui, P, Q;
    std::vector<double> Lx, Ux, Rs;
    IntType do_recip = 1;
  };

  enum class Ordering : int
  {
    natural = 0,
    nested_dissection = 1, // geometric on detected structured grids with diagonal neighbours, else METIS (else RCM)
    rcm = 2,
    graph_nested_dissection = 3, // always the graph partitioner (METIS when compiled in, else RCM)
    geometric_nested_dissection = 4 // plane separators on a detected structured grid of any size and stencil (else as 1)
  };
  constexpr long kGeometricFromRows = 50000; // nested_dissection: below this size the graph partitioner is fast enough (and is
                                             // what the GPU parity tests of the small configurations were verified with)

  namespace detail
  {
    using I = long;

    //! adjacency (pattern of A + A^T without the diagonal), CSR-like, sorted and unique
    template <class Int>
    inline void symmetric_adjacency(I n, const Int *rowptr, const Int *col, std::vector<int64_t> &xadj,
                                    std::vector<int64_t> &adj)
    {
      std::vector<int64_t> deg(n, 0);
      for (I i = 0; i < n; ++i)
        for (Int k = rowptr[i]; k < rowptr[i + 1]; ++k)
          if ((I)col[k] != i)
          {
            deg[i]++;
            deg[col[k]]++;
          }
      std::vector<int64_t> start(n + 1, 0);
      for (I i = 0; i < n; ++i)
        start[i + 1] = start[i] + deg[i];
      std::vector<int64_t> tmp(start[n]);
      std::vector<int64_t> fill(start.begin(), start.end() - 1);
      for (I i = 0; i < n; ++i)
        for (Int k = rowptr[i]; k < rowptr[i + 1]; ++k)
          if ((I)col[k] != i)
          {
            tmp[fill[i]++] = col[k];
            tmp[fill[col[k]]++] = i;
          }
      xadj.assign(n + 1, 0);
      adj.clear();
      adj.reserve(tmp.size());
      for (I i = 0; i < n; ++i)
      {
        std::sort(tmp.begin() + start[i], tmp.begin() + start[i + 1]);
        auto e = std::unique(tmp.begin() + start[i], tmp.begin() + start[i + 1]);
        for (auto p = tmp.begin() + start[i]; p != e; ++p)
          adj.push_back(*p);
        xadj[i + 1] = (int64_t)adj.size();
      }
    }

    //! reverse Cuthill-McKee; perm[k] = original index placed at position k
    inline std::vector<I> rcm_order(I n, const std::vector<int64_t> &xadj, const std::vector<int64_t> &adj)
    {
      std::vector<I> order;
      order.reserve(n);
      std::vector<char> seen(n, 0);
      std::vector<I> byDegree(n);
      std::iota(byDegree.begin(), byDegree.end(), I(0));
      std::stable_sort(byDegree.begin(), byDegree.end(),
                       [&](I a, I b) { return xadj[a + 1] - xadj[a] < xadj[b + 1] - xadj[b]; });
      std::vector<I> nb;
      for (I s : byDegree)
      {
        if (seen[s])
          continue;
        std::size_t head = order.size();
        order.push_back(s);
        seen[s] = 1;
        while (head < order.size())
        {
          I v = order[head++];
          nb.clear();
          for (int64_t k = xadj[v]; k < xadj[v + 1]; ++k)
            if (!seen[adj[k]])
            {
              seen[adj[k]] = 1;
              nb.push_back((I)adj[k]);
            }
          std::sort(nb.begin(), nb.end(),
                    [&](I a, I b) { return xadj[a + 1] - xadj[a] < xadj[b + 1] - xadj[b]; });
          order.insert(order.end(), nb.begin(), nb.end());
        }
      }
      std::reverse(order.begin(), order.end());
      return order;
    }
    /** Is the pattern that of a stencil of radius 1 on an nx x ny x nz grid numbered lexicographically (x fastest)?
     *  Candidate strides are read off the neighbour offsets of the rows around the middle of the matrix (the first offset
     *  beyond 1 is S1 - 1 or S1, the first beyond the x-y plane is one of S2 - S1 - 1 ... S2: stencils may lack face or
     *  diagonal neighbours -- the Q1 Laplace stiffness matrix has exact zeros on the faces, which providers drop) and a
     *  candidate is accepted only if EVERY entry of the matrix is a grid neighbour (|dx|, |dy|, |dz| <= 1). nz = 1 for
     *  2D. diagonal_neighbours: some entry has two or more coordinates differing (9- / 27-point type). */
    template <class Int>
    inline bool detect_structured_grid(I n, const Int *rowptr, const Int *col, I &nx, I &ny, I &nz, bool &diagonal_neighbours)
    {
      if (n < 8)
        return false;
      std::vector<I> pos;
      for (I i = std::max<I>(0, n / 2 - 64); i < std::min<I>(n, n / 2 + 64); ++i)
        for (Int k = rowptr[i]; k < rowptr[i + 1]; ++k)
          if ((I)col[k] > i)
            pos.push_back((I)col[k] - i);
      std::sort(pos.begin(), pos.end());
      pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
      auto valid = [&](I ax, I ay, I az) {
        if (ax < 2 || ay < 2 || az < 1 || ax * ay * az != n)
          return false;
        bool diag = false;
        for (I i = 0; i < n; ++i)
        {
          const I x = i % ax, y = (i / ax) % ay, z = i / (ax * ay);
          for (Int k = rowptr[i]; k < rowptr[i + 1]; ++k)
          {
            const I c = (I)col[k];
            if (c < 0 || c >= n)
              return false;
            const I dx = std::abs(c % ax - x), dy = std::abs((c / ax) % ay - y), dz = std::abs(c / (ax * ay) - z);
            if (dx > 1 || dy > 1 || dz > 1)
              return false;
            diag = diag || dx + dy + dz > 1;
          }
        }
        nx = ax;
        ny = ay;
        nz = az;
        diagonal_neighbours = diag;
        return true;
      };
      I first = 0; // first offset beyond 1: S1 - 1 or S1
      for (I o : pos)
        if (o > 1)
        {
          first = o;
          break;
        }
      if (first == 0)
        return false;
      for (I S1 : {first, first + 1})
      {
        if (n % S1 != 0)
          continue;
        I beyond = 0; // first offset beyond the row above: S2 - S1 - 1 ... S2, or none (2D)
        for (I o : pos)
          if (o > S1 + 1)
          {
            beyond = o;
            break;
          }
        if (beyond == 0)
        {
          if (valid(S1, n / S1, 1))
            return true;
          continue;
        }
        for (I S2 : {beyond + S1 + 1, beyond + S1, beyond + S1 - 1, beyond + 1, beyond})
          if (S2 % S1 == 0 && n % S2 == 0 && valid(S1, S2 / S1, n / S2))
            return true;
      }
      return false;
    }

    /** geometric nested dissection of a box of grid points: the two halves first, the separating plane (one layer: the
     *  stencil has radius 1) last; boxes of at most kLeaf points are numbered lexicographically. perm is appended to. */
    inline void grid_nested_dissection(I nx, I ny, const I lo[3], const I hi[3], std::vector<I> &perm)
    {
      constexpr I kLeaf = 48;
      const I ext[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
      if (ext[0] <= 0 || ext[1] <= 0 || ext[2] <= 0)
        return;
      int d = 0;
      for (int e = 1; e < 3; ++e)
        if (ext[e] > ext[d])
          d = e;
      if (ext[0] * ext[1] * ext[2] <= kLeaf || ext[d] < 3)
      {
        for (I z = lo[2]; z < hi[2]; ++z)
          for (I y = lo[1]; y < hi[1]; ++y)
            for (I x = lo[0]; x < hi[0]; ++x)
              perm.push_back((z * ny + y) * nx + x);
        return;
      }
      const I cut = lo[d] + ext[d] / 2;
      I a_lo[3] = {lo[0], lo[1], lo[2]}, a_hi[3] = {hi[0], hi[1], hi[2]};
      a_hi[d] = cut;
      grid_nested_dissection(nx, ny, a_lo, a_hi, perm);
      I b_lo[3] = {lo[0], lo[1], lo[2]}, b_hi[3] = {hi[0], hi[1], hi[2]};
      b_lo[d] = cut + 1;
      grid_nested_dissection(nx, ny, b_lo, b_hi, perm);
      I s_lo[3] = {lo[0], lo[1], lo[2]}, s_hi[3] = {hi[0], hi[1], hi[2]};
      s_lo[d] = cut;
      s_hi[d] = cut + 1;
      // the separator is a dense block of the factor whatever its internal order: lexicographic
      for (I z = s_lo[2]; z < s_hi[2]; ++z)
        for (I y = s_lo[1]; y < s_hi[1]; ++y)
          for (I x = s_lo[0]; x < s_hi[0]; ++x)
            perm.push_back((z * ny + y) * nx + x);
    }
  } // namespace detail

  //! compute a fill-reducing symmetric ordering; perm[k] = original index that becomes pivot k
  template <class Int>
  inline std::vector<long> compute_ordering(long n, const Int *rowptr, const Int *col, Ordering ord)
  {
    std::vector<long> perm(n);
    std::iota(perm.begin(), perm.end(), 0L);
    if (ord == Ordering::natural || n < 3)
      return perm;
    if (ord == Ordering::geometric_nested_dissection || (ord == Ordering::nested_dissection && n >= kGeometricFromRows))
    {
      // structured grids (every matrix of BASELINE.json): geometric nested dissection in O(n) instead of a graph partitioner
      // (METIS: 31 s for the 128^3 pencil, a fifth of the whole factorisation)
      // Only for stencils with diagonal neighbours (9- / 27-point: every matrix of the 3D configurations), where plane
      // separators match METIS (27-point 64^3: 1.784e8 factor entries against 1.783e8); on 5- / 7-point stencils METIS
      // finds 30-35 % less fill than planes and keeps the job.
      long nx = 0, ny = 0, nz = 0;
      bool diagonal_neighbours = false;
      if (detail::detect_structured_grid(n, rowptr, col, nx, ny, nz, diagonal_neighbours) &&
          (diagonal_neighbours || ord == Ordering::geometric_nested_dissection))
      {
        std::vector<long> g;
        g.reserve(n);
        const long lo[3] = {0, 0, 0}, hi[3] = {nx, ny, nz};
        detail::grid_nested_dissection(nx, ny, lo, hi, g);
        if ((long)g.size() == n)
          return g;
      }
    }
    std::vector<int64_t> xadj, adj;
    detail::symmetric_adjacency(n, rowptr, col, xadj, adj);
#ifdef DE_B200_HAVE_METIS
    if (ord == Ordering::nested_dissection || ord == Ordering::graph_nested_dissection || ord == Ordering::geometric_nested_dissection)
    {
      int64_t nv = n;
      std::vector<int64_t> p(n), ip(n);
      int64_t options[40];
      METIS_SetDefaultOptions(options);
      int rc = METIS_NodeND(&nv, xadj.data(), adj.data(), nullptr, options, p.data(), ip.data());
      if (rc == 1) // METIS_OK
      {
        // METIS: A' = A(perm, perm) in its convention means new position k holds old vertex perm[k]
        for (long k = 0; k < n; ++k)
          perm[k] = (long)p[k];
        return perm;
      }
    }
#endif
    return detail::rcm_order(n, xadj, adj);
  }

  /** \brief sparse LU with static symmetric ordering and no numerical pivoting (see file comment)
   *
   *  \param scale_rows  if true mimic UMFPACK's default row scaling: Rs[i] = sum_j |a_ij|, do_recip = 0
   *                     (rows are divided by Rs); otherwise Rs = 1, do_recip = 1
   */
  template <class Int>
  inline void sparse_lu(long n, const Int *rowptr, const Int *col, const double *val,
                        const std::vector<long> &perm, FactorArrays &F, bool scale_rows = false)
  {
    using I = long;
    if ((I)perm.size() != n)
      throw std::invalid_argument("sparse_lu: permutation has wrong size");
    std::vector<I> iperm(n, -1);
    for (I k = 0; k < n; ++k)
    {
      if (perm[k] < 0 || perm[k] >= n || iperm[perm[k]] != -1)
        throw std::invalid_argument("sparse_lu: not a permutation");
      iperm[perm[k]] = k;
    }

    // row scaling
    F.Rs.assign(n, 1.0);
    F.do_recip = 1;
    if (scale_rows)
    {
      F.do_recip = 0;
      for (I i = 0; i < n; ++i)
      {
        double s = 0.0;
        for (Int k = rowptr[i]; k < rowptr[i + 1]; ++k)
          s += std::abs(val[k]);
        F.Rs[i] = (s > 0.0) ? s : 1.0;
      }
    }

    // --- permuted matrix in CSR (rows/cols in pivot order, explicit zeros dropped like the reference's
    //     BCRS -> CSC conversion does, umfpacktools.hh:68,89)
    std::vector<I> Bp(n + 1, 0);
    for (I k = 0; k < n; ++k)
    {
      I i = perm[k];
      I cnt = 0;
      for (Int q = rowptr[i]; q < rowptr[i + 1]; ++q)
        if (val[q] != 0.0)
          ++cnt;
      Bp[k + 1] = Bp[k] + cnt;
    }
    std::vector<I> Bj(Bp[n]);
    std::vector<double> Bx(Bp[n]);
    for (I k = 0; k < n; ++k)
    {
      I i = perm[k];
      I w = Bp[k];
      const double sc = scale_rows ? 1.0 / F.Rs[i] : 1.0;
      for (Int q = rowptr[i]; q < rowptr[i + 1]; ++q)
        if (val[q] != 0.0)
        {
          Bj[w] = iperm[col[q]];
          Bx[w] = scale_rows ? val[q] * sc : val[q];
          ++w;
        }
    }

    // --- pattern of the strict upper triangle of B + B^T by columns (needed for etree / row patterns)
    std::vector<I> Cp(n + 1, 0);
    for (I k = 0; k < n; ++k)
      for (I q = Bp[k]; q < Bp[k + 1]; ++q)
      {
        I c = Bj[q];
        if (c != k)
          Cp[std::max(c, k) + 1]++;
      }
    for (I k = 0; k < n; ++k)
      Cp[k + 1] += Cp[k];
    std::vector<I> Ci(Cp[n]);
    {
      std::vector<I> w(Cp.begin(), Cp.end() - 1);
      for (I k = 0; k < n; ++k)
        for (I q = Bp[k]; q < Bp[k + 1]; ++q)
        {
          I c = Bj[q];
          if (c != k)
            Ci[w[std::max(c, k)]++] = std::min(c, k);
        }
    }

    // --- elimination tree (Liu's algorithm with path compression)
    std::vector<I> parent(n, -1), ancestor(n, -1);
    for (I k = 0; k < n; ++k)
      for (I q = Cp[k]; q < Cp[k + 1]; ++q)
      {
        I i = Ci[q];
        while (i != -1 && i < k)
        {
          I nxt = ancestor[i];
          ancestor[i] = k;
          if (nxt == -1)
            parent[i] = k;
          i = nxt;
        }
      }

    // --- symbolic: row patterns of L (strictly lower part), sorted ascending
    std::vector<I> Lrp(n + 1, 0);
    std::vector<I> Lrj;
    Lrj.reserve(std::size_t(8) * n);
    std::vector<I> mark(n, -1);
    std::vector<I> ucount(n, 0);
    for (I k = 0; k < n; ++k)
    {
      mark[k] = k;
      std::size_t first = Lrj.size();
      for (I q = Cp[k]; q < Cp[k + 1]; ++q)
      {
        I i = Ci[q];
        while (mark[i] != k)
        {
          Lrj.push_back(i);
          mark[i] = k;
          i = parent[i];
        }
      }
      std::sort(Lrj.begin() + first, Lrj.end());
      for (std::size_t q = first; q < Lrj.size(); ++q)
        ucount[Lrj[q]]++;
      Lrp[k + 1] = (I)Lrj.size();
    }

    // U rows (CSR): diagonal first, then the columns c > j for which L(c,j) is structurally nonzero
    std::vector<I> Urp(n + 1, 0);
    for (I j = 0; j < n; ++j)
      Urp[j + 1] = Urp[j] + 1 + ucount[j];
    std::vector<I> Urj(Urp[n]);
    {
      std::vector<I> w(n);
      for (I j = 0; j < n; ++j)
      {
        Urj[Urp[j]] = j;
        w[j] = Urp[j] + 1;
      }
      for (I c = 0; c < n; ++c)
        for (I q = Lrp[c]; q < Lrp[c + 1]; ++q)
          Urj[w[Lrj[q]]++] = c; // ascending in c by construction
    }
    std::vector<double> Lrx(Lrj.size(), 0.0), Urx(Urj.size(), 0.0);

    // --- numeric, row by row
    std::vector<double> x(n, 0.0);
    I nz_udiag = 0;
    double amax = 0.0;
    for (double v : Bx)
      amax = std::max(amax, std::abs(v));
    for (I k = 0; k < n; ++k)
    {
      for (I q = Bp[k]; q < Bp[k + 1]; ++q)
        x[Bj[q]] += Bx[q];
      for (I q = Lrp[k]; q < Lrp[k + 1]; ++q)
      {
        const I j = Lrj[q];
        const double l = x[j] / Urx[Urp[j]];
        x[j] = 0.0;
        Lrx[q] = l;
        if (l != 0.0)
          for (I t = Urp[j] + 1; t < Urp[j + 1]; ++t)
            x[Urj[t]] -= l * Urx[t];
      }
      for (I t = Urp[k]; t < Urp[k + 1]; ++t)
      {
        Urx[t] = x[Urj[t]];
        x[Urj[t]] = 0.0;
      }
      // a pivot that is zero up to round-off of the elimination counts as structurally singular
      const double piv = Urx[Urp[k]];
      if (std::isfinite(piv) && std::abs(piv) > 64.0 * 2.220446049250313e-16 * amax)
        ++nz_udiag;
    }

    // --- export in the UMFPACK layout
    F.n = F.n_row = F.n_col = n;
    F.nz_udiag = nz_udiag;
    F.lnz = (I)Lrj.size() + n;
    F.unz = (I)Urj.size();
    if (nz_udiag < n)
      throw std::invalid_argument("UMFPackFactorizedMatrix: input matrix is singular");

    F.Lp.assign(n + 1, 0);
    F.Lj.resize(F.lnz);
    F.Lx.resize(F.lnz);
    for (I k = 0; k < n; ++k)
    {
      I w = Lrp[k] + k; // k diagonals have been emitted before row k
      F.Lp[k] = w;
      for (I q = Lrp[k]; q < Lrp[k + 1]; ++q, ++w)
      {
        F.Lj[w] = Lrj[q];
        F.Lx[w] = Lrx[q];
      }
      F.Lj[w] = k;
      F.Lx[w] = 1.0;
    }
    F.Lp[n] = F.lnz;

    // U: CSR -> CSC, diagonal (largest row index of the column) ends up last
    F.Up.assign(n + 1, 0);
    F.Ui.resize(F.unz);
    F.Ux.resize(F.unz);
    for (I t = 0; t < F.unz; ++t)
      F.Up[Urj[t] + 1]++;
    for (I c = 0; c < n; ++c)
      F.Up[c + 1] += F.Up[c];
    {
      std::vector<I> w(F.Up.begin(), F.Up.end() - 1);
      for (I j = 0; j < n; ++j)
        for (I t = Urp[j]; t < Urp[j + 1]; ++t)
        {
          I dst = w[Urj[t]]++;
          F.Ui[dst] = j;
          F.Ux[dst] = Urx[t];
        }
    }
    F.P.assign(perm.begin(), perm.end());
    F.Q.assign(perm.begin(), perm.end());
  }

  /** Backward-error check of a factorisation in the field contract: solves A x = b for one deterministic right-hand side
   *  through the factor arrays exactly as the reference's apply does (scale + permute, L, U, scatter: kernels_cpp.hh:682-750)
   *  and returns ||A x - b||_inf / (||A||_inf ||x||_inf + ||b||_inf). The LU above uses a STATIC ordering without
   *  numerical pivoting (UMFPACK pivots): for a matrix that is indefinite after the shift, or far from symmetric, small
   *  pivots can make the factor silently inaccurate -- this check turns that into an error. O(nnz + lnz + unz). */
  template <class Int>
  inline double factorization_backward_error(long n, const Int *rowptr, const Int *col, const double *val, const FactorArrays &F)
  {
    using I = long;
    std::vector<double> xt(n), b(n, 0.0), w(n), x(n);
    double anorm = 0.0;
    for (I i = 0; i < n; ++i)
      xt[i] = 1.0 + 0.5 * std::sin(0.7 * (double)i + 0.3);
    for (I i = 0; i < n; ++i)
    {
      double s = 0.0, rs = 0.0;
      for (Int q = rowptr[i]; q < rowptr[i + 1]; ++q)
      {
        s += val[q] * xt[col[q]];
        rs += std::abs(val[q]);
      }
      b[i] = s;
      anorm = std::max(anorm, rs);
    }
    for (I k = 0; k < n; ++k)
      w[k] = (F.do_recip ? F.Rs[F.P[k]] : 1.0 / F.Rs[F.P[k]]) * b[F.P[k]];
    for (I i = 0; i < n; ++i) // L: unit lower, rows, diagonal last
    {
      double s = w[i];
      for (I q = F.Lp[i]; q < F.Lp[i + 1] - 1; ++q)
        s -= F.Lx[q] * w[F.Lj[q]];
      w[i] = s;
    }
    for (I j = n - 1; j >= 0; --j) // U: columns, diagonal last
    {
      const double xj = w[j] / F.Ux[F.Up[j + 1] - 1];
      w[j] = xj;
      for (I q = F.Up[j]; q < F.Up[j + 1] - 1; ++q)
        w[F.Ui[q]] -= F.Ux[q] * xj;
    }
    for (I j = 0; j < n; ++j)
      x[F.Q[j]] = w[j];
    double res = 0.0, xn = 0.0, bn = 0.0;
    for (I i = 0; i < n; ++i)
    {
      double s = -b[i];
      for (Int q = rowptr[i]; q < rowptr[i + 1]; ++q)
        s += val[q] * x[col[q]];
      if (!(s == s))
        return std::numeric_limits<double>::infinity();
      res = std::max(res, std::abs(s));
      xn = std::max(xn, std::abs(x[i]));
      bn = std::max(bn, std::abs(b[i]));
    }
    return res / std::max(anorm * xn + bn, std::numeric_limits<double>::min());
  }

  constexpr double kMaxBackwardError = 1e-9; // a backward-stable factorisation gives ~1e-16; static pivoting gone wrong: >> 1e-9

  //! convenience: ordering + factorisation + backward-error check (throws like a singular matrix does if it fails)
  template <class Int>
  inline void factorize_csr(long n, const Int *rowptr, const Int *col, const double *val, FactorArrays &F,
                            Ordering ord = Ordering::nested_dissection, bool scale_rows = false)
  {
    std::vector<long> perm = compute_ordering(n, rowptr, col, ord);
    sparse_lu(n, rowptr, col, val, perm, F, scale_rows);
    const double be = n > 0 ? factorization_backward_error(n, rowptr, col, val, F) : 0.0;
    if (!(be <= kMaxBackwardError))
      throw std::invalid_argument("UMFPackFactorizedMatrix: the factorisation failed its backward-error check (relative residual " +
                                  std::to_string(be) + "): this provider uses a static ordering WITHOUT numerical pivoting, which is "
                                  "only safe for matrices that are (nearly) symmetric positive definite after the shift; the "
                                  "input matrix is singular or indefinite to working precision");
  }
} // namespace de_b200

#endif
