// Test program for the C++ drop-in headers (include/dune/eigensolver/*.hh): the same source a DUNE user would
// write against the reference's eigensolver.hh, compiled against this repo's headers + libdune_eigensolver_b200.so.
// The matrix type is the minimal BCRSMatrix stand-in from oracle/shim (test infrastructure; real dune-istl at a
// user's site). Prints results as "key value..." lines that tests/test_cpp_dropin.py compares with the oracle.
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include <dune/common/fmatrix.hh>
#include <dune/istl/bcrsmatrix.hh>

#include <dune/eigensolver/eigensolver.hh>

using Block = Dune::FieldMatrix<double, 1, 1>;
using Matrix = Dune::BCRSMatrix<Block>;

// 2D 5-point matrices of the reference's driver (reference src/dune-eigensolver.cc:98-143)
static Matrix laplacian(int N, const char *kind, int overlap)
{
  std::vector<long> ptr(1, 0), col;
  std::vector<double> val;
  auto pu = [&](int x, int y) { return (x < overlap || x > N - 1 - overlap || y < overlap || y > N - 1 - overlap) ? 0.0 : 1.0; };
  for (int y = 0; y < N; ++y)
    for (int x = 0; x < N; ++x)
    {
      const int nb = (y > 0) + (x > 0) + (x < N - 1) + (y < N - 1);
      auto put = [&](int xx, int yy, double v) {
        if (!std::strcmp(kind, "B"))
          v *= pu(x, y) * pu(xx, yy);
        col.push_back(yy * N + xx);
        val.push_back(v);
      };
      const double diag = !std::strcmp(kind, "neumann") ? (double)nb : 4.0;
      const double off = !std::strcmp(kind, "mass") ? 0.5 : -1.0; // "mass": an SPD matrix on the same pattern
      if (y > 0) put(x, y - 1, off);
      if (x > 0) put(x - 1, y, off);
      put(x, y, diag);
      if (x < N - 1) put(x + 1, y, off);
      if (y < N - 1) put(x, y + 1, off);
      ptr.push_back((long)col.size());
    }
  return Matrix((std::size_t)N * N, (std::size_t)N * N, ptr.data(), col.data(), val.data());
}

static void print(const char *key, const std::vector<double> &v);

// BCSR: the 2D Laplacian with k x k blocks L_ij * C, C = tridiag(0.5, 2 + c, 0.5) -- the matrix kron(L, C)
template <int K>
static Dune::BCRSMatrix<Dune::FieldMatrix<double, K, K>> block_laplacian(int N)
{
  using Blk = Dune::FieldMatrix<double, K, K>;
  Blk C;
  for (int r = 0; r < K; ++r)
    for (int c = 0; c < K; ++c)
      C[r][c] = r == c ? 2.0 + r : ((r - c == 1 || c - r == 1) ? 0.5 : 0.0);
  std::vector<long> ptr(1, 0), col;
  std::vector<Blk> val;
  for (int y = 0; y < N; ++y)
    for (int x = 0; x < N; ++x)
    {
      auto put = [&](int xx, int yy, double s) {
        Blk b;
        for (int r = 0; r < K; ++r)
          for (int c = 0; c < K; ++c)
            b[r][c] = s * C[r][c];
        col.push_back(yy * N + xx);
        val.push_back(b);
      };
      if (y > 0) put(x, y - 1, -1.0);
      if (x > 0) put(x - 1, y, -1.0);
      put(x, y, 4.0);
      if (x < N - 1) put(x + 1, y, -1.0);
      if (y < N - 1) put(x, y + 1, -1.0);
      ptr.push_back((long)col.size());
    }
  return Dune::BCRSMatrix<Blk>((std::size_t)N * N, (std::size_t)N * N, ptr.data(), col.data(), val.data());
}

template <int K>
static int run_bcsr(int N, int nev, double tol)
{
  auto A = block_laplacian<K>(N);
  const std::size_t nb = (std::size_t)N * N;
  std::vector<double> eval(nev);
  // a BlockVector<FieldVector<double,K>> stand-in: nb block entries of K components
  std::vector<std::vector<std::array<double, K>>> evec(nev, std::vector<std::array<double, K>>(nb));
  StandardLargest(A, 0.0, tol, 4000, nev, eval, evec, 0, 123);
  print("eval", eval);
  std::vector<double> head;
  for (int j = 0; j < nev; ++j)
    for (int c = 0; c < K; ++c)
      head.push_back(evec[j][1][c]); // scalar rows K .. 2K-1
  print("evec_block1", head);
  // the kernel-level entry point on the same block matrix: Y = A X for a block with nb * K rows
  MultiVector<double, 8> X = de_b200::random_start_block(nb * K, 8, 7), Y{nb * K, 8};
  matmul_sparse_tallskinny_blocked(Y, A, X);
  std::vector<double> y0;
  for (int j = 0; j < 8; ++j)
    y0.push_back(Y(K + 1, j));
  print("spmm_row", y0);
  return 0;
}

static void print(const char *key, const std::vector<double> &v)
{
  std::printf("%s", key);
  for (double x : v)
    std::printf(" %.17g", x);
  std::printf("\n");
}

int main(int argc, char **argv)
{
  const std::string mode = argc > 1 ? argv[1] : "largest";
  const int N = argc > 2 ? std::atoi(argv[2]) : 12;
  const int nev = argc > 3 ? std::atoi(argv[3]) : 8;
  const double tol = argc > 4 ? std::atof(argv[4]) : 1e-10;
  // 5th argument: comma-separated CUDA ordinals, e.g. "0,1" (or "0,0": two ranks on one GPU) -> the drivers run
  // row-partitioned through the single-process multi-GPU front end (ini key parallel.numgpus of the driver program)
  const std::string devs = argc > 5 ? argv[5] : "";
  const std::size_t n = (std::size_t)N * N;
  if (mode == "costmodel")
  {
    // the product header's analytic cost models (kernels_b200.hh; reference kernels_cpp.hh:98-116, :157-175): no GPU needed
    for (int i = 2; i + 2 < argc; i += 3)
    {
      const int n = std::atoi(argv[i]), m = std::atoi(argv[i + 1]), b = std::atoi(argv[i + 2]);
      std::printf("cost %d %d %d %.17g %.17g %.17g\n", n, m, b, flops_orthonormalize(n, m), bytes_orthonormalize_naive(n, m),
                  bytes_orthonormalize_blocked(n, m, b));
    }
    return 0;
  }
  try
  {
    if (!devs.empty())
    {
      std::vector<int> d;
      for (std::size_t i = 0; i < devs.size();)
      {
        const std::size_t j = devs.find(',', i);
        d.push_back(std::atoi(devs.substr(i, j == std::string::npos ? j : j - i).c_str()));
        if (j == std::string::npos)
          break;
        i = j + 1;
      }
      de_b200::Parallel::instance().set_devices(d);
      de_b200::Parallel::instance().set_row_align(N); // one grid line of the 2D grid
      std::printf("gpus %d\n", de_b200::Parallel::instance().num_gpus());
    }
    std::vector<double> eval(nev);
    std::vector<std::vector<double>> evec(nev, std::vector<double>(n));
    if (mode == "largest")
    {
      Matrix A = laplacian(N, "dirichlet", 0);
      StandardLargest(A, 0.0, tol, 4000, nev, eval, evec, 0, 123);
    }
    else if (mode == "inverse")
    {
      Matrix A = laplacian(N, "dirichlet", 0);
      StandardInverse(A, 1e-3, tol, 4000, nev, eval, evec, 0, 123);
      // the caller's matrix was shifted in place (reference eigensolver.hh:145-153)
      std::printf("diag0 %.17g\n", (double)(*(A.begin()->begin())));
    }
    else if (mode == "generalized")
    {
      Matrix A = laplacian(N, "neumann", 0), B = laplacian(N, "B", 3);
      std::vector<double> ev;
      std::vector<std::vector<double>> V;
      GeneralizedInverse(A, B, 1e-3, 0.0, tol, 4000, nev, ev, V, 1, 123);
      eval = ev;
      evec = V;
    }
    else if (mode == "lobpcg")
    {
      Matrix A = laplacian(N, "dirichlet", 0);
      StandardLOBPCG(A, tol, 4000, nev, eval, evec, 0, 123);
    }
    else if (mode == "globpcg")
    {
      Matrix A = laplacian(N, "dirichlet", 0), B = laplacian(N, "mass", 0);
      std::vector<double> ev;
      std::vector<std::vector<double>> V;
      GeneralizedLOBPCG(A, B, tol, 4000, nev, ev, V, 0, 123); // resizes its outputs like GeneralizedInverse
      eval = ev;
      evec = V;
    }
    else if (mode == "bcsr2")
      return run_bcsr<2>(N, nev, tol);
    else if (mode == "bcsr3")
      return run_bcsr<3>(N, nev, tol);
    else if (mode == "cholesky")
    {
      // the second factorisation provider through the header: supernodal Cholesky inside the library, factored apply on
      // the GPU (dense panels), public UMFPACK-contract members expanded from it
      Matrix A = laplacian(N, "dirichlet", 0);
      UMFPackFactorizedMatrix<Matrix> F(A, 0, de_b200::Ordering::nested_dissection, de_b200::Provider::cholesky);
      MultiVector<double, 8> X = de_b200::random_start_block(n, 16, 5), B{n, 16}, Y{n, 16};
      matmul_sparse_tallskinny_blocked(B, A, X);
      matmul_inverse_tallskinny_blocked(Y, F, B);
      double err = 0.0;
      for (std::size_t i = 0; i < n; ++i)
        for (std::size_t j = 0; j < 16; ++j)
          err = std::max(err, std::abs(Y(i, j) - X(i, j)));
      std::printf("supernodal %d\n", (int)F.supernodal());
      std::printf("contract %d %ld %ld\n", (int)(F.Lp != nullptr && F.Ux != nullptr), (long)F.n, (long)F.lnz);
      std::printf("solve_error %.3e\n", err);
      // Provider::automatic picks it for large symmetric matrices: shift-invert driver on the same grid
      StandardInverse(A, 1e-3, tol, 4000, nev, eval, evec, 0, 123);
    }
    else if (mode == "kernels")
    {
      Matrix A = laplacian(N, "dirichlet", 0);
      MultiVector<double, 8> X = de_b200::random_start_block(n, 16, 123), Y{n, 16};
      matmul_sparse_tallskinny_blocked(Y, A, X);
      std::vector<double> dp;
      dot_products_diagonal_blocked(dp, X, Y);
      print("diagdot", dp);
      orthonormalize_blocked(X);
      std::vector<std::vector<double>> G = dot_products_diagonal(X);
      double off = 0.0;
      for (std::size_t i = 0; i < 16; ++i)
        for (std::size_t j = 0; j < 16; ++j)
          off = std::max(off, std::abs(G[i][j] - (i == j ? 1.0 : 0.0)));
      std::printf("ortho_defect %.3e\n", off);
      try
      {
        MultiVector<double, 8> bad{n, 12};
      }
      catch (const std::invalid_argument &e)
      {
        std::printf("caught %s\n", e.what());
      }
      return 0;
    }
    print("eval", eval);
    std::vector<double> head;
    for (int j = 0; j < nev; ++j)
      head.push_back(evec[j][0]);
    print("evec_row0", head);
  }
  catch (const std::exception &e)
  {
    std::printf("exception %s\n", e.what());
    return 3;
  }
  return 0;
}
