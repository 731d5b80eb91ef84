// TEST INFRASTRUCTURE ONLY (oracle). Never linked into, imported by or called from the product.
//
// "oracle/_ref": the reference's OWN implementation of the hot path, compiled verbatim.
// This translation unit #includes /root/reference/dune/eigensolver/eigensolver.hh from the
// read-only mount (nothing is copied into this repo) against the small DUNE header shim in
// oracle/shim/, and exposes the reference functions through a plain C interface so that the
// Python tests, smoke() and bench.py's cpu_baseline leg can call them with flat arrays.
//
// Third-party hole: SuiteSparse UMFPACK is neither vendored nor installed, so
// UMFPackFactorizedMatrix (reference umfpacktools.hh:16-220, compiled out because
// HAVE_SUITESPARSE_UMFPACK is undefined) is supplied here as an explicit specialisation with the
// same public fields, filled by the repo's host factorisation provider
// (include/dune/eigensolver/sparse_lu.hh). The factorisation itself is therefore "parity unpinned"
// (SURVEY.md §8c); the factored APPLY and everything else is the reference's code.
//
// Layout of every multivector argument: the reference MultiVector<double,8> layout
// (reference multivector.hh:130-133): element (i,j) at ((j/8)*n + i)*8 + j%8.

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <new>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <dune/common/fmatrix.hh>
#include <dune/common/timer.hh>
#include <dune/istl/bcrsmatrix.hh>
#include <dune/eigensolver/sparse_lu.hh>

using RefBlock = Dune::FieldMatrix<double, 1, 1>;
using RefMatrix = Dune::BCRSMatrix<RefBlock>;

static int g_ordering = (int)de_b200::Ordering::nested_dissection;
static int g_scale_rows = 0;

// ---- stand-in for the UMFPACK-backed class (same public field names as the reference) ----------
template <typename T>
class UMFPackFactorizedMatrix
{
};

template <>
class UMFPackFactorizedMatrix<RefMatrix>
{
  de_b200::FactorArrays store_;

public:
  using IntType = long;
  IntType n, lnz, unz, n_row, n_col, nz_udiag;
  IntType *Lp, *Lj;
  double *Lx;
  IntType *Up, *Ui;
  double *Ux;
  IntType *P, *Q;
  IntType do_recip;
  double *Rs;

  //! borrow externally supplied factor arrays (for the kernel-level apply test)
  UMFPackFactorizedMatrix(IntType n_, IntType *Lp_, IntType *Lj_, double *Lx_, IntType *Up_, IntType *Ui_,
                          double *Ux_, IntType *P_, IntType *Q_, double *Rs_, IntType do_recip_)
      : n(n_), lnz(Lp_[n_]), unz(Up_[n_]), n_row(n_), n_col(n_), nz_udiag(n_), Lp(Lp_), Lj(Lj_), Lx(Lx_),
        Up(Up_), Ui(Ui_), Ux(Ux_), P(P_), Q(Q_), do_recip(do_recip_), Rs(Rs_)
  {
  }

  //! factorise (what the reference does through UMFPACK in umfpacktools.hh:46-199)
  UMFPackFactorizedMatrix(const RefMatrix &A, int /*verbose*/ = 0)
  {
    if (A.N() != A.M())
      throw std::invalid_argument("UMFPackFactorizedMatrix: input matrix must be square");
    const long nn = (long)A.N();
    std::vector<double> v(A.nonzeroes());
    for (std::size_t k = 0; k < v.size(); ++k)
      v[k] = A.raw_val()[k];
    de_b200::factorize_csr(nn, A.raw_ptr().data(), A.raw_col().data(), v.data(), store_,
                           (de_b200::Ordering)g_ordering, g_scale_rows != 0);
    n = n_row = n_col = nn;
    lnz = store_.lnz;
    unz = store_.unz;
    nz_udiag = store_.nz_udiag;
    Lp = store_.Lp.data();
    Lj = store_.Lj.data();
    Lx = store_.Lx.data();
    Up = store_.Up.data();
    Ui = store_.Ui.data();
    Ux = store_.Ux.data();
    P = store_.P.data();
    Q = store_.Q.data();
    Rs = store_.Rs.data();
    do_recip = store_.do_recip;
  }
  UMFPackFactorizedMatrix(const UMFPackFactorizedMatrix &) = delete;
  UMFPackFactorizedMatrix &operator=(const UMFPackFactorizedMatrix &) = delete;
};

// ---- the reference, verbatim, from the read-only mount (-I/root/reference) ----------------------
#include <dune/eigensolver/eigensolver.hh>

namespace
{
  using MV8 = MultiVector<double, 8>;

  //! the reference prints from constructors and drivers; capture std::cout for the duration of a call
  struct CoutCapture
  {
    std::ostringstream buf;
    std::streambuf *old;
    CoutCapture() : old(std::cout.rdbuf(buf.rdbuf())) {}
    ~CoutCapture() { std::cout.rdbuf(old); }
  };

  MV8 load_mv(long n, long m, const double *src)
  {
    MV8 Q{(std::size_t)n, (std::size_t)m};
    if (n * m > 0)
      std::memcpy(&Q(0, 0), src, sizeof(double) * n * m);
    return Q;
  }
  void store_mv(const MV8 &Q, double *dst)
  {
    if (Q.rows() * Q.cols() > 0)
      std::memcpy(dst, &Q(0, 0), sizeof(double) * Q.rows() * Q.cols());
  }

  //! last integer following `key` in the captured log, or fallback
  long last_int_after(const std::string &log, const std::string &key, long fallback)
  {
    std::size_t pos = log.rfind(key);
    if (pos == std::string::npos)
      return fallback;
    return std::strtol(log.c_str() + pos + key.size(), nullptr, 10);
  }

  thread_local std::string g_err;
  int fail(const std::exception &e)
  {
    g_err = e.what();
    return 1;
  }
} // namespace

extern "C"
{
  const char *orc_kind() { return "reference"; }
  const char *orc_last_error() { return g_err.c_str(); }
  void orc_set_factor_options(int ordering, int scale_rows)
  {
    g_ordering = ordering;
    g_scale_rows = scale_rows;
  }

  //! start block exactly as the reference drivers fill it (eigensolver.hh:50-55)
  int orc_start_block(long n, long m, unsigned seed, double *out)
  {
    std::mt19937 urbg{seed};
    std::normal_distribution<double> generator{0.0, 1.0};
    for (long bj = 0; bj < m; bj += 8)
      for (long i = 0; i < n; ++i)
        for (long j = 0; j < 8; ++j)
          out[(bj / 8 * n + i) * 8 + j] = generator(urbg);
    return 0;
  }

  int orc_spmm(long n, const long *rowptr, const long *col, const double *val, long m, const double *xin,
               double *yout)
  {
    try
    {
      CoutCapture cap;
      RefMatrix A(n, n, rowptr, col, val);
      MV8 X = load_mv(n, m, xin), Y{(std::size_t)n, (std::size_t)m};
      matmul_sparse_tallskinny_blocked(Y, A, X);
      store_mv(Y, yout);
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  int orc_diag_dot(long n, long m, const double *x1, const double *x2, double *dp)
  {
    try
    {
      CoutCapture cap;
      MV8 A = load_mv(n, m, x1), B = load_mv(n, m, x2);
      std::vector<double> d;
      dot_products_diagonal_blocked(d, A, B);
      std::copy(d.begin(), d.end(), dp);
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  //! full Gram G = X1^T X2, row-major m x m
  int orc_gram(long n, long m, const double *x1, const double *x2, double *g)
  {
    try
    {
      CoutCapture cap;
      MV8 A = load_mv(n, m, x1), B = load_mv(n, m, x2);
      std::vector<std::vector<double>> d;
      dot_products_all_blocked(d, A, B);
      for (long i = 0; i < m; ++i)
        for (long j = 0; j < m; ++j)
          g[i * m + j] = d[i][j];
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  int orc_orthonormalize(long n, long m, double *x)
  {
    try
    {
      CoutCapture cap;
      MV8 Q = load_mv(n, m, x);
      orthonormalize_blocked(Q);
      store_mv(Q, x);
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  //! b = 1 variant; layout is then plain column-major n x m
  int orc_orthonormalize_naive(long n, long m, double *x)
  {
    try
    {
      CoutCapture cap;
      MultiVector<double, 1> Q{(std::size_t)n, (std::size_t)m};
      std::memcpy(&Q(0, 0), x, sizeof(double) * n * m);
      orthonormalize_naive(Q);
      std::memcpy(x, &Q(0, 0), sizeof(double) * n * m);
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  int orc_b_orthonormalize(long n, const long *rowptr, const long *col, const double *val, long m, double *x,
                           double *norm_out)
  {
    try
    {
      CoutCapture cap;
      RefMatrix B(n, n, rowptr, col, val);
      MV8 Q = load_mv(n, m, x);
      double nrm = B_orthonormalize_blocked(B, Q);
      if (norm_out)
        *norm_out = nrm;
      store_mv(Q, x);
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  //! xout = (factored A)^-1 xin ; xin is destroyed exactly as in the reference
  int orc_factor_apply(long n, long m, long *Lp, long *Lj, double *Lx, long *Up, long *Ui, double *Ux, long *P,
                       long *Q, double *Rs, long do_recip, double *xin, double *xout)
  {
    try
    {
      CoutCapture cap;
      UMFPackFactorizedMatrix<RefMatrix> F(n, Lp, Lj, Lx, Up, Ui, Ux, P, Q, Rs, do_recip);
      MV8 X = load_mv(n, m, xin), Y{(std::size_t)n, (std::size_t)m};
      matmul_inverse_tallskinny_blocked(Y, F, X);
      store_mv(Y, xout);
      store_mv(X, xin);
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  //! val is modified in place when shift != 0 (the reference mutates the caller's matrix)
  int orc_standard_largest(long n, const long *rowptr, const long *col, double *val, double shift, double tol,
                           int maxiter, int nev, unsigned seed, double *eval, double *evec, long *iterations)
  {
    try
    {
      CoutCapture cap;
      RefMatrix A(n, n, rowptr, col, val);
      std::vector<double> ev(nev, 0.0);
      std::vector<std::vector<double>> V(nev, std::vector<double>(n, 0.0));
      StandardLargest(A, shift, tol, maxiter, nev, ev, V, 1, seed);
      for (std::size_t k = 0; k < A.nonzeroes(); ++k)
        val[k] = A.raw_val()[k];
      std::copy(ev.begin(), ev.end(), eval);
      for (int j = 0; j < nev; ++j)
        std::copy(V[j].begin(), V[j].end(), evec + (std::size_t)j * n);
      // loop index at exit: the last "Iter=k" printed (k>1), else min(1, maxiter-1)
      if (iterations)
        *iterations = last_int_after(cap.buf.str(), "Iter=", std::min(1, maxiter - 1));
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  int orc_standard_inverse(long n, const long *rowptr, const long *col, double *val, double shift, double tol,
                           int maxiter, int nev, unsigned seed, double *eval, double *evec, long *iterations)
  {
    try
    {
      CoutCapture cap;
      RefMatrix A(n, n, rowptr, col, val);
      std::vector<double> ev(nev, 0.0);
      std::vector<std::vector<double>> V(nev, std::vector<double>(n, 0.0));
      StandardInverse(A, shift, tol, maxiter, nev, ev, V, 1, seed);
      for (std::size_t k = 0; k < A.nonzeroes(); ++k)
        val[k] = A.raw_val()[k];
      std::copy(ev.begin(), ev.end(), eval);
      for (int j = 0; j < nev; ++j)
        std::copy(V[j].begin(), V[j].end(), evec + (std::size_t)j * n);
      if (iterations)
        *iterations = last_int_after(cap.buf.str(), "iter=", std::min(1, maxiter - 1));
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  int orc_generalized_inverse(long n, const long *rowptrA, const long *colA, const double *valA,
                              const long *rowptrB, const long *colB, const double *valB, double shift, double reg,
                              double tol, int maxiter, int nev, unsigned seed, double *eval, double *evec,
                              long *iterations)
  {
    try
    {
      CoutCapture cap;
      RefMatrix A(n, n, rowptrA, colA, valA), B(n, n, rowptrB, colB, valB);
      std::vector<double> ev;
      std::vector<std::vector<double>> V;
      GeneralizedInverse(A, B, shift, reg, tol, maxiter, nev, ev, V, 1, seed);
      std::copy(ev.begin(), ev.end(), eval);
      for (int j = 0; j < nev; ++j)
        std::copy(V[j].begin(), V[j].end(), evec + (std::size_t)j * n);
      if (iterations)
        *iterations = last_int_after(cap.buf.str(), "iterations=", -1);
      return 0;
    }
    catch (const std::exception &e)
    {
      return fail(e);
    }
  }

  // the reference's analytic cost models (kernels_cpp.hh:98-116,157-175)
  double orc_flops_orthonormalize(int n, int m) { return flops_orthonormalize(n, m); }
  double orc_bytes_orthonormalize_naive(int n, int m) { return bytes_orthonormalize_naive(n, m); }
  double orc_bytes_orthonormalize_blocked(int n, int m, int b) { return bytes_orthonormalize_blocked(n, m, b); }
}
