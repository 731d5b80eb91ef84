// CPU check of the LOBPCG orchestration (dune_eigensolver_b200/csrc/lobpcg_core.hpp) and of the host Rayleigh-Ritz
// solver (host_eig.hpp): the same template the library instantiates with its device kernels is instantiated here with
// plain host loops. TEST INFRASTRUCTURE ONLY -- nothing in the library links or calls this.
//
// usage: lobpcg_host_test N nev tol generalized(0|1) largest(0|1) [verbose] [mgs(0|1)] [cheb_degree] [cheb_ratio] [maxiter]
// prints "iterations k", "restarts r", "eval ...", "maxres ...", "orth ..." ; exit code 0 if converged.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../dune_eigensolver_b200/csrc/lobpcg_core.hpp"

struct Csr
{
  int n = 0;
  std::vector<int> ptr, col;
  std::vector<double> val;
  void apply(const double *x, double *y, int m) const
  {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
      {
        double s = 0.0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k)
          s += val[k] * x[(size_t)col[k] * m + j];
        y[(size_t)i * m + j] = s;
      }
  }
};

// 5-point stencil on an N x N grid: center `c`, neighbours `o` (the reference's Laplacian is c = 4, o = -1,
// src/dune-eigensolver.cc:98-103)
static Csr stencil2d(int N, double c, double o)
{
  Csr A;
  A.n = N * N;
  A.ptr.push_back(0);
  for (int y = 0; y < N; ++y)
    for (int x = 0; x < N; ++x)
    {
      auto add = [&](int xx, int yy, double v) {
        if (xx >= 0 && xx < N && yy >= 0 && yy < N)
        {
          A.col.push_back(yy * N + xx);
          A.val.push_back(v);
        }
      };
      add(x, y - 1, o);
      add(x - 1, y, o);
      add(x, y, c);
      add(x + 1, y, o);
      add(x, y + 1, o);
      A.ptr.push_back((int)A.col.size());
    }
  return A;
}

struct HostOps
{
  using Blk = double *;
  const Csr *A = nullptr, *B = nullptr;
  int n = 0, m = 0;
  std::vector<std::vector<double>> pool;

  int alloc(Blk *b)
  {
    pool.emplace_back((size_t)n * m, 0.0);
    *b = pool.back().data();
    return 0;
  }
  void gram(const double *L, const double *R, double *G) const
  {
    for (int a = 0; a < m; ++a)
      for (int b = 0; b < m; ++b)
      {
        double s = 0.0;
        for (int i = 0; i < n; ++i)
          s += L[(size_t)i * m + a] * R[(size_t)i * m + b];
        G[(size_t)a * m + b] = s;
      }
  }
  /** CholQR2 with the pivot test of the device kernel (chol_inverse2_kernel, kernels_dense.cuh: a pivot that is not
   *  above 4 m eps G_kk reports rank deficiency), so that the orchestration sees the same failures on the CPU as on
   *  the GPU; set `mgs` for modified Gram-Schmidt instead. */
  bool mgs = false;
  int orthonormalize(Blk X, Blk BX)
  {
    if (mgs)
      return orthonormalize_mgs(X, BX);
    std::vector<double> G((size_t)m * m), R((size_t)m * m), T((size_t)n * m);
    for (int sweep = 0; sweep < 2; ++sweep)
    {
      if (B)
        B->apply(X, BX, m);
      gram(X, B ? BX : X, G.data());
      // G = R^T R, R upper triangular
      std::fill(R.begin(), R.end(), 0.0);
      for (int k = 0; k < m; ++k)
      {
        double d = G[(size_t)k * m + k];
        for (int q = 0; q < k; ++q)
          d -= R[(size_t)q * m + k] * R[(size_t)q * m + k];
        if (!(d > 4.0 * m * 2.220446049250313e-16 * G[(size_t)k * m + k]))
          return 5;
        R[(size_t)k * m + k] = std::sqrt(d);
        for (int j = k + 1; j < m; ++j)
        {
          double v = G[(size_t)k * m + j];
          for (int q = 0; q < k; ++q)
            v -= R[(size_t)q * m + k] * R[(size_t)q * m + j];
          R[(size_t)k * m + j] = v / R[(size_t)k * m + k];
        }
      }
      // X <- X R^-1 : forward substitution per row
      for (int i = 0; i < n; ++i)
        for (int j = 0; j < m; ++j)
        {
          double v = X[(size_t)i * m + j];
          for (int q = 0; q < j; ++q)
            v -= T[(size_t)i * m + q] * R[(size_t)q * m + j];
          T[(size_t)i * m + j] = v / R[(size_t)j * m + j];
        }
      std::copy(T.begin(), T.end(), X);
    }
    if (B)
      B->apply(X, BX, m);
    return 0;
  }
  int orthonormalize_mgs(Blk X, Blk BX)
  {
    // modified Gram-Schmidt in the B inner product, twice
    std::vector<double> bx((size_t)n);
    for (int sweep = 0; sweep < 2; ++sweep)
      for (int j = 0; j < m; ++j)
      {
        for (int k = 0; k <= j; ++k)
        {
          // bx = B x_k (k < j: projection; k == j: norm)
          for (int i = 0; i < n; ++i)
          {
            double s = 0.0;
            if (B)
              for (int q = B->ptr[i]; q < B->ptr[i + 1]; ++q)
                s += B->val[q] * X[(size_t)B->col[q] * m + k];
            else
              s = X[(size_t)i * m + k];
            bx[i] = s;
          }
          double d = 0.0;
          for (int i = 0; i < n; ++i)
            d += bx[i] * X[(size_t)i * m + j];
          if (k < j)
            for (int i = 0; i < n; ++i)
              X[(size_t)i * m + j] -= d * X[(size_t)i * m + k];
          else
          {
            if (!(d > 0.0))
              return 5;
            const double inv = 1.0 / std::sqrt(d);
            for (int i = 0; i < n; ++i)
              X[(size_t)i * m + j] *= inv;
          }
        }
      }
    if (B)
      B->apply(X, BX, m);
    return 0;
  }
  int apply_A(Blk Y, Blk X)
  {
    ++spmm_count;
    A->apply(X, Y, m);
    return 0;
  }
  int apply_B(Blk Y, Blk X)
  {
    B->apply(X, Y, m);
    return 0;
  }
  int residual(Blk W, Blk AX, Blk BX, const double *theta, double *norm2)
  {
    for (int j = 0; j < m; ++j)
      norm2[j] = 0.0;
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
      {
        const double r = AX[(size_t)i * m + j] - theta[j] * BX[(size_t)i * m + j];
        W[(size_t)i * m + j] = r;
        norm2[j] += r * r;
      }
    return 0;
  }
  int precondition(Blk) { return 0; }
  /** Jacobi-scaled Chebyshev: the bound is the Gershgorin bound of D^-1 A and every correction is scaled by D^-1 */
  std::vector<double> dinv;
  int spectral_bound(double *b)
  {
    double g = 0.0;
    dinv.assign(n, 1.0);
    for (int i = 0; i < n; ++i)
    {
      double r = 0.0, d = 0.0;
      for (int k = A->ptr[i]; k < A->ptr[i + 1]; ++k)
      {
        r += std::abs(A->val[k]);
        if (A->col[k] == i)
          d = A->val[k];
      }
      if (!(d > 0.0))
      {
        *b = HUGE_VAL; // like gershgorin_kernel: no Jacobi scale, the caller runs unpreconditioned
        return 0;
      }
      dinv[i] = 1.0 / d;
      g = std::max(g, r / d);
    }
    *b = g;
    return 0;
  }
  int cheb_start(Blk Z, Blk Zold, Blk R, double s)
  {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
      {
        Z[(size_t)i * m + j] = s * dinv[i] * R[(size_t)i * m + j];
        Zold[(size_t)i * m + j] = 0.0;
      }
    return 0;
  }
  int apply_A_cheb(Blk Zold, Blk Z, Blk R, Blk AZ, double alpha, double beta)
  {
    apply_A(AZ, Z);
    return cheb_step(Zold, Z, R, AZ, alpha, beta);
  }
  int cheb_step(Blk Zold, Blk Z, Blk R, Blk AZ, double alpha, double beta)
  {
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
      {
        const size_t e = (size_t)i * m + j;
        Zold[e] = Z[e] + alpha * (Z[e] - Zold[e]) + beta * dinv[i] * (R[e] - AZ[e]);
      }
    return 0;
  }
  long spmm_count = 0;
  int project(Blk W, Blk X, Blk BX)
  {
    std::vector<double> G((size_t)m * m);
    gram(BX, W, G.data());
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < m; ++j)
      {
        double s = 0.0;
        for (int k = 0; k < m; ++k)
          s += X[(size_t)i * m + k] * G[(size_t)k * m + j];
        W[(size_t)i * m + j] -= s;
      }
    return 0;
  }
  int grams(int count, const Blk *L, const Blk *R, const char *, double *out)
  {
    for (int g = 0; g < count; ++g)
      gram(L[g], R[g], out + (size_t)g * m * m);
    return 0;
  }
  int rotate(Blk X, const double *C)
  {
    const Blk S[1] = {X};
    return lincomb(1, S, C, X, nullptr);
  }
  int lincomb(int ns, const Blk *S, const double *C, Blk out, Blk out2)
  {
    std::vector<double> row(m), row2(m);
    for (int i = 0; i < n; ++i)
    {
      for (int j = 0; j < m; ++j)
        row[j] = row2[j] = 0.0;
      for (int s = ns - 1; s >= 0; --s)
      {
        for (int j = 0; j < m; ++j)
        {
          double acc = 0.0;
          for (int k = 0; k < m; ++k)
            acc += S[s][(size_t)i * m + k] * C[(size_t)s * m * m + (size_t)k * m + j];
          row[j] += acc;
          if (s >= 1)
            row2[j] += acc;
        }
      }
      for (int j = 0; j < m; ++j)
      {
        out[(size_t)i * m + j] = row[j];
        if (out2 && ns > 1)
          out2[(size_t)i * m + j] = row2[j];
      }
    }
    return 0;
  }
};

// CSR matrix from a binary file: int64 n, int64 nnz, int64 rowptr[n+1], int64 col[nnz], double val[nnz]
static bool load_csr(const char *path, Csr &A)
{
  FILE *f = std::fopen(path, "rb");
  if (!f)
    return false;
  long long n = 0, nnz = 0;
  bool ok = std::fread(&n, 8, 1, f) == 1 && std::fread(&nnz, 8, 1, f) == 1;
  std::vector<long long> rp(ok ? n + 1 : 0), ci(ok ? nnz : 0);
  A.val.resize(ok ? nnz : 0);
  ok = ok && std::fread(rp.data(), 8, n + 1, f) == (size_t)(n + 1) && std::fread(ci.data(), 8, nnz, f) == (size_t)nnz &&
       std::fread(A.val.data(), 8, nnz, f) == (size_t)nnz;
  std::fclose(f);
  A.n = (int)n;
  A.ptr.assign(rp.begin(), rp.end());
  A.col.assign(ci.begin(), ci.end());
  return ok;
}

int main(int argc, char **argv)
{
  const int N = argc > 1 ? std::atoi(argv[1]) : 20;
  const int nev = argc > 2 ? std::atoi(argv[2]) : 8;
  const double tol = argc > 3 ? std::atof(argv[3]) : 1e-8;
  const int generalized = argc > 4 ? std::atoi(argv[4]) : 0;
  const int largest = argc > 5 ? std::atoi(argv[5]) : 0;

  Csr A = stencil2d(N, 4.0, -1.0);
  Csr B = stencil2d(N, 4.0, 0.5); // SPD "mass-like" matrix on the same pattern
  if (std::getenv("LOBPCG_TEST_NEGATE_A")) // a negative definite matrix: no positive diagonal
    for (double &v : A.val)
      v = -v;
  if (const char *fa = std::getenv("LOBPCG_TEST_A")) // any CSR matrix instead (tests with variable coefficients)
    if (!load_csr(fa, A))
      return 3;
  if (const char *fb = std::getenv("LOBPCG_TEST_B"))
    if (!load_csr(fb, B))
      return 3;
  HostOps ops;
  ops.A = &A;
  ops.B = generalized ? &B : nullptr;
  ops.n = A.n;
  ops.mgs = argc > 7 && std::atoi(argv[7]) != 0;
  ops.m = (nev / 8 + (nev % 8 ? 1 : 0)) * 8;

  de::LobpcgParams prm;
  prm.m = ops.m;
  prm.nev = nev;
  prm.tol = tol;
  prm.maxiter = argc > 10 ? std::atoi(argv[10]) : 2000;
  prm.verbose = argc > 6 ? std::atoi(argv[6]) : 0;
  prm.has_B = generalized != 0;
  prm.largest = largest != 0;
  prm.cheb_degree = argc > 8 ? std::atoi(argv[8]) : 0;
  prm.cheb_ratio = argc > 9 ? std::atof(argv[9]) : 0.0;

  double *X;
  ops.alloc(&X);
  std::mt19937 gen(123);
  std::normal_distribution<double> dist(0.0, 1.0);
  for (size_t i = 0; i < (size_t)ops.n * ops.m; ++i)
    X[i] = dist(gen);

  de::LobpcgResult res;
  const int rc = de::lobpcg_run(ops, prm, X, res);
  std::printf("rc %d\n", rc);
  std::printf("iterations %d\nrestarts %d\nconverged %d\nspmm %ld\n", res.iterations, res.restarts, (int)res.converged,
              ops.spmm_count);
  std::printf("eval");
  for (int j = 0; j < nev; ++j)
    std::printf(" %.15e", res.theta[j]);
  std::printf("\n");

  // independent check of the returned pairs: residuals and B-orthonormality from scratch
  const int m = ops.m, n = ops.n;
  std::vector<double> AXv((size_t)n * m), BXv((size_t)n * m);
  A.apply(X, AXv.data(), m);
  if (generalized)
    B.apply(X, BXv.data(), m);
  else
    BXv.assign(X, X + (size_t)n * m);
  double maxres = 0.0, orth = 0.0;
  for (int j = 0; j < nev; ++j)
  {
    double s = 0.0;
    for (int i = 0; i < n; ++i)
    {
      const double r = AXv[(size_t)i * m + j] - res.theta[j] * BXv[(size_t)i * m + j];
      s += r * r;
    }
    maxres = std::max(maxres, std::sqrt(s) / std::abs(res.theta[j]));
  }
  for (int a = 0; a < m; ++a)
    for (int b = 0; b < m; ++b)
    {
      double s = 0.0;
      for (int i = 0; i < n; ++i)
        s += X[(size_t)i * m + a] * BXv[(size_t)i * m + b];
      orth = std::max(orth, std::abs(s - (a == b ? 1.0 : 0.0)));
    }
  std::printf("maxres %.3e\north %.3e\n", maxres, orth);
  return (rc == 0 && res.converged) ? 0 : 1;
}
