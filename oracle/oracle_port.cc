// TEST INFRASTRUCTURE ONLY (oracle). Never linked into, imported by or called from the product.
//
// "port": this repo's CPU restatement of the reference's hot path, in plain C++ over flat arrays
// (CSR with 64-bit indices; multivectors in the reference MultiVector<double,8> layout, element (i,j)
// at ((j/8)*n + i)*8 + j%8, reference multivector.hh:130-133). It exists so that a checker is
// available where the reference mount is not (and as a second opinion next to oracle/_ref).
//
// PINNING: tests/test_oracle.py checks every function here against (a) the reference compiled
// verbatim (oracle/_ref, built from /root/reference by oracle/Makefile), (b) the golden vectors in
// tests/golden/ generated from that build, (c) the analytic spectra of the reference's own test
// (src/dune-eigensolver.cc:437-446).
// UNPINNED: the LU factorisation feeding the shift-invert drivers (UMFPACK is not available; see
// include/dune/eigensolver/sparse_lu.hh). The factored *apply* is pinned.
//
// Each function names the reference lines it follows. Loop orders and accumulation orders are kept
// where they determine round-off (row-wise CSR accumulation, panel order in the Gram-Schmidt sweeps).

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include <dune/eigensolver/sparse_lu.hh>

namespace
{
  constexpr long PB = 8; // panel width b (reference eigensolver.hh:35,123,214)

  struct Csr
  {
    long n;
    const long *ptr;
    const long *col;
    const double *val;
  };

  inline double *panel(double *x, long n, long bj) { return x + bj * n; } // start of the panel that holds column bj

  // ---- reference kernels_cpp.hh:626-657 (matmul_sparse_tallskinny_blocked) -------------------------
  void spmm(const Csr &A, long m, const double *xin, double *yout)
  {
    for (long bj = 0; bj < m; bj += PB)
    {
      const double *x = xin + bj * A.n;
      double *y = yout + bj * A.n;
      for (long i = 0; i < A.n; ++i)
      {
        double acc[PB] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (long k = A.ptr[i]; k < A.ptr[i + 1]; ++k)
        {
          const double a = A.val[k];
          const double *xr = x + A.col[k] * PB;
          for (long j = 0; j < PB; ++j)
            acc[j] += a * xr[j];
        }
        for (long j = 0; j < PB; ++j)
          y[i * PB + j] = acc[j];
      }
    }
  }

  // ---- reference kernels_cpp.hh:24-55 (dot_products_diagonal_blocked) ------------------------------
  void diag_dot(long n, long m, const double *x1, const double *x2, double *dp)
  {
    for (long bj = 0; bj < m; bj += PB)
    {
      double s[PB] = {0, 0, 0, 0, 0, 0, 0, 0};
      const double *a = x1 + bj * n, *b = x2 + bj * n;
      for (long i = 0; i < n; ++i)
        for (long j = 0; j < PB; ++j)
          s[j] += a[i * PB + j] * b[i * PB + j];
      for (long j = 0; j < PB; ++j)
        dp[bj + j] = s[j];
    }
  }

  // ---- reference kernels_cpp.hh:58-96 (dot_products_all_blocked): G = X1^T X2 ----------------------
  void gram(long n, long m, const double *x1, const double *x2, double *g)
  {
    for (long b1 = 0; b1 < m; b1 += PB)
      for (long b2 = 0; b2 < m; b2 += PB)
      {
        double s[PB][PB] = {};
        const double *a = x1 + b1 * n, *b = x2 + b2 * n;
        for (long i = 0; i < n; ++i)
          for (long j1 = 0; j1 < PB; ++j1)
            for (long j2 = 0; j2 < PB; ++j2)
              s[j1][j2] += a[i * PB + j1] * b[i * PB + j2];
        for (long j1 = 0; j1 < PB; ++j1)
          for (long j2 = 0; j2 < PB; ++j2)
            g[(b1 + j1) * m + b2 + j2] = s[j1][j2];
      }
  }

  // projection of panel bj against the (already orthonormal) panel whose B-image / self is `w`
  // reference kernels_cpp.hh:320-348 and :553-583.  Returns max S entry (only used by the B variant).
  double project_panel(long n, const double *w, const double *qk, double *qj)
  {
    double s[PB][PB] = {};
    for (long i = 0; i < n; ++i)
      for (long k = 0; k < PB; ++k)
        for (long j = 0; j < PB; ++j)
          s[k][j] += w[i * PB + k] * qj[i * PB + j];
    double nrm = -INFINITY;
    for (long k = 0; k < PB; ++k)
      for (long j = 0; j < PB; ++j)
        nrm = std::max(nrm, s[k][j]);
    for (long i = 0; i < n; ++i)
      for (long k = 0; k < PB; ++k)
        for (long j = 0; j < PB; ++j)
          qj[i * PB + j] -= s[k][j] * qk[i * PB + k];
    return nrm;
  }

  // ---- reference kernels_cpp.hh:180-351 (orthonormalize_blocked, live `if (true)` branch) ----------
  void orthonormalize(long n, long m, double *q)
  {
    for (long bk = 0; bk < m; bk += PB)
    {
      double *v = panel(q, n, bk);
      // column-by-column modified Gram-Schmidt inside the panel (:204-228)
      for (long k = 0; k < PB; ++k)
      {
        double s[PB] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (long i = 0; i < n; ++i)
          for (long j = k; j < PB; ++j)
            s[j] += v[i * PB + k] * v[i * PB + j];
        for (long j = k + 1; j < PB; ++j)
          s[j] /= s[k];
        s[k] = 1.0 / std::sqrt(s[k]);
        for (long i = 0; i < n; ++i)
        {
          for (long j = k + 1; j < PB; ++j)
            v[i * PB + j] -= s[j] * v[i * PB + k];
          v[i * PB + k] *= s[k];
        }
      }
      // project every later panel against this one (:309-349)
      for (long bj = bk + PB; bj < m; bj += PB)
        project_panel(n, v, v, panel(q, n, bj));
    }
  }

  // ---- reference kernels_cpp.hh:121-155 (orthonormalize_naive, b = 1: plain column-major) ----------
  void orthonormalize_naive(long n, long m, double *q)
  {
    for (long k = 0; k < m; ++k)
    {
      double *qk = q + k * n;
      double s = 0.0;
      for (long i = 0; i < n; ++i)
        s += qk[i] * qk[i];
      s = 1.0 / std::sqrt(s);
      for (long i = 0; i < n; ++i)
        qk[i] *= s;
      for (long j = k + 1; j < m; ++j)
      {
        double *qj = q + j * n;
        s = 0.0;
        for (long i = 0; i < n; ++i)
          s += qk[i] * qj[i];
        for (long i = 0; i < n; ++i)
          qj[i] -= s * qk[i];
      }
    }
  }

  // U = L^-T D^-1/2 from the un-pivoted LU of the symmetric 8x8 Gram S (reference kernels_cpp.hh:468-512)
  void cholqr_factor(const double s[PB][PB], double U[PB][PB])
  {
    double LU[PB][PB];
    for (long k = 0; k < PB; ++k)
      for (long j = 0; j < PB; ++j)
        LU[k][j] = s[k][j];
    for (long k = 0; k < PB; ++k)
      for (long i = k + 1; i < PB; ++i)
      {
        LU[i][k] /= LU[k][k];
        for (long j = k + 1; j < PB; ++j)
          LU[i][j] -= LU[i][k] * LU[k][j];
      }
    double D[PB];
    for (long i = 0; i < PB; ++i)
      D[i] = 1.0 / std::sqrt(LU[i][i]);
    for (long i = 0; i < PB; ++i)
    {
      LU[i][i] = 1.0;
      for (long j = i + 1; j < PB; ++j)
        LU[i][j] = 0.0;
    }
    for (long i = 0; i < PB; ++i)
      for (long j = 0; j < PB; ++j)
        U[i][j] = (i == j) ? 1.0 : 0.0;
    for (long i = 1; i < PB; ++i) // rows of L^-1
      for (long j = 0; j < i; ++j)
        for (long k = 0; k < PB; ++k)
          U[i][k] -= LU[i][j] * U[j][k];
    for (long i = 0; i < PB; ++i)
      for (long j = 0; j < i; ++j)
        std::swap(U[i][j], U[j][i]);
    for (long i = 0; i < PB; ++i)
      for (long j = i; j < PB; ++j)
        U[i][j] *= D[j];
  }

  // V <- V U with U upper triangular, in place, right to left (reference kernels_cpp.hh:514-539)
  void right_multiply_upper(long n, double *v, const double U[PB][PB])
  {
    for (long i = 0; i < n; ++i)
      for (long j = PB - 1; j >= 0; --j)
      {
        double sum = 0.0;
        for (long k = 0; k <= j; ++k)
          sum += v[i * PB + k] * U[k][j];
        v[i * PB + j] = sum;
      }
  }

  // ---- reference kernels_cpp.hh:356-591 (B_orthonormalize_blocked, live CholQR branch) --------------
  double b_orthonormalize(const Csr &B, long m, double *q)
  {
    const long n = B.n;
    std::vector<double> p(n * PB);
    double norm = 0.0;
    for (long bk = 0; bk < m; bk += PB)
    {
      double *v = panel(q, n, bk);
      // p = B * v (:380-395)
      for (long i = 0; i < n; ++i)
      {
        double acc[PB] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (long k = B.ptr[i]; k < B.ptr[i + 1]; ++k)
          for (long j = 0; j < PB; ++j)
            acc[j] += B.val[k] * v[B.col[k] * PB + j];
        for (long j = 0; j < PB; ++j)
          p[i * PB + j] = acc[j];
      }
      // S = p^T v, upper triangle then mirrored (:451-463)
      double s[PB][PB] = {};
      for (long i = 0; i < n; ++i)
        for (long k = 0; k < PB; ++k)
          for (long j = k; j < PB; ++j)
            s[k][j] += p[i * PB + k] * v[i * PB + j];
      for (long k = 0; k < PB; ++k)
        for (long j = 0; j < k; ++j)
          s[k][j] = s[j][k];
      for (long k = 0; k < PB; ++k)
        for (long j = k + 1; j < PB; ++j)
          norm = std::max(norm, s[k][j]);
      double U[PB][PB];
      cholqr_factor(s, U);
      right_multiply_upper(n, v, U);        // (:514-526)
      right_multiply_upper(n, p.data(), U); // keeps p = B v (:527-539)
      for (long bj = bk + PB; bj < m; bj += PB)
        norm = std::max(norm, project_panel(n, p.data(), v, panel(q, n, bj)));
    }
    return norm;
  }

  // ---- reference kernels_cpp.hh:660-755 (matmul_inverse_tallskinny_blocked) -------------------------
  struct Factors
  {
    long n;
    const long *Lp, *Lj;
    const double *Lx;
    const long *Up, *Ui;
    const double *Ux;
    const long *P, *Q;
    const double *Rs;
    long do_recip;
  };

  void factor_apply(const Factors &F, long m, double *xin, double *xout)
  {
    const long n = F.n;
    for (long bj = 0; bj < m; bj += PB)
    {
      double *in = xin + bj * n, *out = xout + bj * n;
      // row scaling + row permutation into `out` (:682-705)
      for (long k = 0; k < n; ++k)
      {
        const double sc = F.do_recip ? F.Rs[F.P[k]] : 1.0 / F.Rs[F.P[k]];
        for (long s = 0; s < PB; ++s)
          out[k * PB + s] = sc * in[F.P[k] * PB + s];
      }
      // forward solve with unit-lower L in CSR, diagonal stored last and skipped (:710-728); result in `in`
      for (long i = 0; i < n; ++i)
      {
        double sum[PB];
        for (long s = 0; s < PB; ++s)
          sum[s] = out[i * PB + s];
        for (long k = F.Lp[i]; k < F.Lp[i + 1] - 1; ++k)
          for (long s = 0; s < PB; ++s)
            sum[s] -= F.Lx[k] * in[F.Lj[k] * PB + s];
        for (long s = 0; s < PB; ++s)
          in[i * PB + s] = sum[s];
      }
      // backward solve with U in CSC, diagonal last in each column, column-oriented updates; the
      // solution row j is written to row Q[j] of `out` (:732-753)
      for (long j = n - 1; j >= 0; --j)
      {
        const double d = F.Ux[F.Up[j + 1] - 1];
        double r[PB];
        for (long s = 0; s < PB; ++s)
          r[s] = in[j * PB + s] / d;
        for (long k = F.Up[j]; k < F.Up[j + 1] - 1; ++k)
          for (long s = 0; s < PB; ++s)
            in[F.Ui[k] * PB + s] -= F.Ux[k] * r[s];
        for (long s = 0; s < PB; ++s)
          out[F.Q[j] * PB + s] = r[s];
      }
    }
  }

  // ---- start block (reference eigensolver.hh:50-55): libstdc++ mt19937 + normal_distribution ---------
  void start_block(long n, long m, unsigned seed, double *out)
  {
    std::mt19937 urbg{seed};
    std::normal_distribution<double> gen{0.0, 1.0};
    for (long bj = 0; bj < m; bj += PB)
      for (long i = 0; i < n; ++i)
        for (long j = 0; j < PB; ++j)
          out[(bj / PB * n + i) * PB + j] = gen(urbg);
  }

  inline long padded_cols(int nev) { return (nev / PB + std::min<long>(nev % PB, 1)) * PB; } // eigensolver.hh:43

  void add_to_diagonal(long n, const long *ptr, const long *col, double *val, double s)
  {
    for (long i = 0; i < n; ++i)
      for (long k = ptr[i]; k < ptr[i + 1]; ++k)
        if (col[k] == i)
          val[k] += s;
  }

  void copy_out(long n, long m, int nev, const double *q, const double *s, double *eval, double *evec)
  {
    for (int j = 0; j < nev; ++j)
    {
      eval[j] = s[j];
      const double *pj = q + (j / PB) * n * PB + j % PB;
      for (long i = 0; i < n; ++i)
        evec[(std::size_t)j * n + i] = pj[i * PB];
    }
    (void)m;
  }

  int g_ordering = (int)de_b200::Ordering::nested_dissection;
  int g_scale_rows = 0;

  Factors view(const de_b200::FactorArrays &S)
  {
    return Factors{S.n,        S.Lp.data(), S.Lj.data(), S.Lx.data(), S.Up.data(), S.Ui.data(),
                   S.Ux.data(), S.P.data(),  S.Q.data(),  S.Rs.data(), S.do_recip};
  }

  // ---- reference eigensolver.hh:28-112 (largest) and :116-198 (inverse): one skeleton ---------------
  long standard_driver(bool inverse, long n, const long *ptr, const long *col, double *val, double shift, double tol,
                       int maxiter, int nev, unsigned seed, double *eval, double *evec)
  {
    const long m = padded_cols(nev);
    std::vector<double> Q1(n * m), Q2(n * m);
    start_block(n, m, seed, Q1.data());
    if (shift != 0.0)
      add_to_diagonal(n, ptr, col, val, shift); // mutates the caller's matrix (:57-66)
    Csr A{n, ptr, col, val};
    de_b200::FactorArrays store;
    if (inverse)
      de_b200::factorize_csr(n, ptr, col, val, store, (de_b200::Ordering)g_ordering, g_scale_rows != 0);
    orthonormalize(n, m, Q1.data());
    std::vector<double> s1(m, 0.0), s2(m, 0.0);
    long k_exit = std::min(1, maxiter - 1);
    for (long k = 1; k < maxiter; ++k)
    {
      k_exit = k;
      if (inverse)
        factor_apply(view(store), m, Q1.data(), Q2.data()); // Q1 is scratch afterwards (:168)
      else
        spmm(A, m, Q1.data(), Q2.data());
      orthonormalize(n, m, Q2.data());
      spmm(A, m, Q2.data(), Q1.data());
      diag_dot(n, m, Q2.data(), Q1.data(), s1.data());
      double distance = 0.0;
      for (long i = 0; i < m; ++i)
      {
        s1[i] -= shift;
        distance = std::max(distance, std::abs(s1[i] - s2[i]));
      }
      std::swap(s1, s2);
      std::swap(Q1, Q2);
      if (k > 1 && distance < tol) // absolute change of the Rayleigh quotients (:101-102)
        break;
    }
    copy_out(n, m, nev, Q1.data(), s2.data(), eval, evec);
    return k_exit;
  }

  // ---- reference eigensolver.hh:204-351 (GeneralizedInverse) ---------------------------------------
  long generalized_inverse(long n, const long *ptrA, const long *colA, const double *valA_in, const long *ptrB,
                           const long *colB, const double *valB, double shift, double reg, double tol, int maxiter,
                           int nev, unsigned seed, double *eval, double *evec)
  {
    const long m = padded_cols(nev);
    std::vector<double> valA(valA_in, valA_in + ptrA[n]); // the driver works on a copy (:208)
    std::vector<double> Q1(n * m), Q2(n * m);
    start_block(n, m, seed, Q1.data());
    if (shift != 0.0) // A += shift * B, pattern(B) must be contained in pattern(A) (:241-242)
      for (long i = 0; i < n; ++i)
      {
        long k = ptrA[i];
        for (long kb = ptrB[i]; kb < ptrB[i + 1]; ++kb)
        {
          while (k < ptrA[i + 1] && colA[k] < colB[kb])
            ++k;
          if (k == ptrA[i + 1] || colA[k] != colB[kb])
            throw std::invalid_argument("GeneralizedInverse: pattern of B not contained in pattern of A");
          valA[k] += shift * valB[kb];
        }
      }
    if (reg != 0.0)
      add_to_diagonal(n, ptrA, colA, valA.data(), reg);
    Csr A{n, ptrA, colA, valA.data()}, B{n, ptrB, colB, valB};
    de_b200::FactorArrays store;
    de_b200::factorize_csr(n, ptrA, colA, valA.data(), store, (de_b200::Ordering)g_ordering, g_scale_rows != 0);

    std::vector<double> ra1(m, 0.0), ra2(m, 0.0), sA(m, 0.0);
    b_orthonormalize(B, m, Q1.data());
    spmm(A, m, Q1.data(), Q2.data());
    diag_dot(n, m, Q2.data(), Q1.data(), sA.data());
    for (long i = 0; i < m; ++i)
      ra2[i] = sA[i] - shift;
    long iter = 0;
    while (iter < maxiter)
    {
      spmm(B, m, Q1.data(), Q2.data());
      factor_apply(view(store), m, Q2.data(), Q1.data());
      b_orthonormalize(B, m, Q1.data());
      ++iter;
      spmm(A, m, Q1.data(), Q2.data());
      diag_dot(n, m, Q2.data(), Q1.data(), sA.data());
      double relerror = 0.0;
      for (long i = 0; i < m; ++i)
      {
        ra1[i] = sA[i] - shift;
        relerror = std::max(relerror, std::abs(ra1[i] - ra2[i]));
      }
      relerror /= *std::max_element(ra1.begin(), ra1.end());
      std::swap(ra1, ra2);
      if (iter > 10 && relerror < tol) // relative change, at least 11 iterations (:315-324)
        break;
    }
    copy_out(n, m, nev, Q1.data(), ra2.data(), eval, evec);
    return iter;
  }

  thread_local std::string g_err;
  template <class F>
  int guarded(F &&f)
  {
    try
    {
      f();
      return 0;
    }
    catch (const std::exception &e)
    {
      g_err = e.what();
      return 1;
    }
  }
} // namespace

extern "C"
{
  const char *orc_kind() { return "port"; }
  const char *orc_last_error() { return g_err.c_str(); }
  void orc_set_factor_options(int ordering, int scale_rows)
  {
    g_ordering = ordering;
    g_scale_rows = scale_rows;
  }
  int orc_start_block(long n, long m, unsigned seed, double *out)
  {
    return guarded([&] { start_block(n, m, seed, out); });
  }
  int orc_spmm(long n, const long *rowptr, const long *col, const double *val, long m, const double *xin,
               double *yout)
  {
    return guarded([&] { spmm(Csr{n, rowptr, col, val}, m, xin, yout); });
  }
  int orc_diag_dot(long n, long m, const double *x1, const double *x2, double *dp)
  {
    return guarded([&] { diag_dot(n, m, x1, x2, dp); });
  }
  int orc_gram(long n, long m, const double *x1, const double *x2, double *g)
  {
    return guarded([&] { gram(n, m, x1, x2, g); });
  }
  int orc_orthonormalize(long n, long m, double *x)
  {
    return guarded([&] { orthonormalize(n, m, x); });
  }
  int orc_orthonormalize_naive(long n, long m, double *x)
  {
    return guarded([&] { orthonormalize_naive(n, m, x); });
  }
  int orc_b_orthonormalize(long n, const long *rowptr, const long *col, const double *val, long m, double *x,
                           double *norm_out)
  {
    return guarded([&] {
      double nrm = b_orthonormalize(Csr{n, rowptr, col, val}, m, x);
      if (norm_out)
        *norm_out = nrm;
    });
  }
  int orc_factor_apply(long n, long m, long *Lp, long *Lj, double *Lx, long *Up, long *Ui, double *Ux, long *P,
                       long *Q, double *Rs, long do_recip, double *xin, double *xout)
  {
    return guarded([&] { factor_apply(Factors{n, Lp, Lj, Lx, Up, Ui, Ux, P, Q, Rs, do_recip}, m, xin, xout); });
  }
  int orc_standard_largest(long n, const long *rowptr, const long *col, double *val, double shift, double tol,
                           int maxiter, int nev, unsigned seed, double *eval, double *evec, long *iterations)
  {
    return guarded([&] {
      long k = standard_driver(false, n, rowptr, col, val, shift, tol, maxiter, nev, seed, eval, evec);
      if (iterations)
        *iterations = k;
    });
  }
  int orc_standard_inverse(long n, const long *rowptr, const long *col, double *val, double shift, double tol,
                           int maxiter, int nev, unsigned seed, double *eval, double *evec, long *iterations)
  {
    return guarded([&] {
      long k = standard_driver(true, n, rowptr, col, val, shift, tol, maxiter, nev, seed, eval, evec);
      if (iterations)
        *iterations = k;
    });
  }
  int orc_generalized_inverse(long n, const long *rowptrA, const long *colA, const double *valA,
                              const long *rowptrB, const long *colB, const double *valB, double shift, double reg,
                              double tol, int maxiter, int nev, unsigned seed, double *eval, double *evec,
                              long *iterations)
  {
    return guarded([&] {
      long it = generalized_inverse(n, rowptrA, colA, valA, rowptrB, colB, valB, shift, reg, tol, maxiter, nev, seed,
                                    eval, evec);
      if (iterations)
        *iterations = it;
    });
  }

  // reference cost models, kernels_cpp.hh:98-106, :108-116, :157-175
  double orc_flops_orthonormalize(int n, int m)
  {
    double f = 0.0;
    for (int k = m; k > 0; --k)
      f += 3.0 * n + 4.0 * n * (k - 1);
    return f;
  }
  double orc_bytes_orthonormalize_naive(int n, int m)
  {
    double w = 0.0;
    for (int k = m; k > 0; --k)
      w += 3.0 * n + 5.0 * n * (k - 1);
    return 8.0 * w;
  }
  double orc_bytes_orthonormalize_blocked(int n, int m, int b)
  {
    double w = 0.0;
    for (int bk = 0; bk < m; bk += b)
    {
      for (int k = b; k > 0; --k)
        w += (double)n * k + (double)n * (k + 1);
      for (int bj = bk + b; bj < m; bj += b)
        w += 5.0 * b * n;
    }
    return 8.0 * w;
  }
}
