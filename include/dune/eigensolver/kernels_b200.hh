#ifndef DUNE_EIGENSOLVER_B200_KERNELS_HH
#define DUNE_EIGENSOLVER_B200_KERNELS_HH

/** \file
 *  Drop-in replacements for the reference's block kernels (reference dune/eigensolver/kernels_cpp.hh and the
 *  kernels_avx2.hh / kernels_neon.hh variants of the same functions): same names, same argument meaning, same
 *  exceptions. Each call uploads its host MultiVector arguments, runs the sm_100a kernel through the C ABI and
 *  downloads the result, so user code that calls the kernels one by one keeps working; the eigensolver drivers in
 *  eigensolver.hh do NOT go through these wrappers but keep everything device-resident for the whole solve.
 *  There is no CPU implementation behind any of them.
 */

#include <stdexcept>
#include <vector>

#include "b200_runtime.hh"
#include "multivector.hh"
#include "umfpacktools.hh"

//! Qout = A * Qin (reference matmul_sparse_tallskinny_blocked, kernels_cpp.hh:626-657)
template <typename MV, typename ISTLM>
void matmul_sparse_tallskinny_blocked(MV &Qout, const ISTLM &A, const MV &Qin)
{
  de_b200::require_scalar_blocks<ISTLM>("matmul_sparse_tallskinny_blocked");
  de_b200::require_block8<MV>("matmul_sparse_tallskinny_blocked");
  auto &ctx = de_b200::Context::thread_default();
  de_b200::DeviceMatrix dA(ctx, A);
  de_b200::DeviceMV dIn(ctx, Qin), dOut(ctx, Qout.rows(), Qout.cols());
  de_b200::check(de_spmm(dOut.get(), dA.get(), dIn.get()), ctx.get());
  dOut.download(Qout);
}

//! the b = 1 variant of the reference (kernels_cpp.hh:596-621) has the same result; kept for API parity
template <typename MV, typename ISTLM>
void matmul_sparse_tallskinny_naive(MV &Qout, const ISTLM &A, const MV &Qin)
{
  de_b200::require_scalar_blocks<ISTLM>("matmul_sparse_tallskinny_naive");
  if (MV::blocksize != 1)
    throw std::invalid_argument("matmul_sparse_tallskinny_naive: blocksize must be one");
  const std::size_t n = Qin.rows(), m = Qin.cols(), mp = (m + 7) / 8 * 8;
  MultiVector<double, 8> in(n, mp), out(n, mp);
  for (std::size_t j = 0; j < m; ++j)
    for (std::size_t i = 0; i < n; ++i)
      in(i, j) = Qin(i, j);
  matmul_sparse_tallskinny_blocked(out, A, in);
  for (std::size_t j = 0; j < m; ++j)
    for (std::size_t i = 0; i < n; ++i)
      Qout(i, j) = out(i, j);
}

//! dp[j] = <Q1[:,j], Q2[:,j]> (reference dot_products_diagonal_blocked, kernels_cpp.hh:24-55)
template <typename MV>
void dot_products_diagonal_blocked(std::vector<double> &dp, const MV &Q1, const MV &Q2)
{
  de_b200::require_block8<MV>("dot_products_diagonal_blocked");
  if (dp.size() != Q1.cols())
    dp.resize(Q1.cols());
  if (Q1.rows() != Q2.rows())
    throw std::invalid_argument("dot_products_blocked: number of rows does not match");
  if (Q1.cols() != Q2.cols())
    throw std::invalid_argument("dot_products_blocked: number of columns does not match");
  auto &ctx = de_b200::Context::thread_default();
  de_b200::DeviceMV d1(ctx, Q1), d2(ctx, Q2);
  de_b200::check(de_diag_dot(dp.data(), d1.get(), d2.get()), ctx.get());
}

//! dp = Q1^T Q2 (reference dot_products_all_blocked, kernels_cpp.hh:58-96)
template <typename MV>
void dot_products_all_blocked(std::vector<std::vector<double>> &dp, const MV &Q1, const MV &Q2)
{
  de_b200::require_block8<MV>("dot_products_all_blocked");
  if (Q1.rows() != Q2.rows())
    throw std::invalid_argument("dot_products_blocked: number of rows does not match");
  if (Q1.cols() != Q2.cols())
    throw std::invalid_argument("dot_products_blocked: number of columns does not match");
  const std::size_t m = Q1.cols();
  auto &ctx = de_b200::Context::thread_default();
  de_b200::DeviceMV d1(ctx, Q1), d2(ctx, Q2);
  std::vector<double> flat(m * m);
  de_b200::check(de_gram(flat.data(), d1.get(), d2.get()), ctx.get());
  dp.assign(m, std::vector<double>(m));
  for (std::size_t i = 0; i < m; ++i)
    for (std::size_t j = 0; j < m; ++j)
      dp[i][j] = flat[i * m + j];
}

//! full Gram matrix Q^T Q (reference dot_products_diagonal(Q), kernels_cpp.hh:7-21)
template <typename MV>
std::vector<std::vector<double>> dot_products_diagonal(const MV &Q)
{
  std::vector<std::vector<double>> dp;
  dot_products_all_blocked(dp, Q, Q);
  return dp;
}

//! in-place thin QR, triangular factor with positive diagonal (reference orthonormalize_blocked, kernels_cpp.hh:180-351)
template <typename MV>
void orthonormalize_blocked(MV &Q)
{
  de_b200::require_block8<MV>("orthonormalize_blocked");
  auto &ctx = de_b200::Context::thread_default();
  de_b200::DeviceMV d(ctx, Q);
  de_b200::check(de_orthonormalize(d.get()), ctx.get());
  d.download(Q);
}

//! B-orthonormalisation (reference B_orthonormalize_blocked, kernels_cpp.hh:356-591)
template <typename ISTLM, typename MV>
double B_orthonormalize_blocked(const ISTLM &B, MV &Q)
{
  de_b200::require_scalar_blocks<ISTLM>("B_orthonormalize_blocked");
  de_b200::require_block8<MV>("B_orthonormalize_blocked");
  auto &ctx = de_b200::Context::thread_default();
  de_b200::DeviceMatrix dB(ctx, B);
  de_b200::DeviceMV d(ctx, Q);
  double norm = 0.0;
  de_b200::check(de_b_orthonormalize(dB.get(), d.get(), nullptr, &norm), ctx.get());
  d.download(Q);
  return norm;
}

//! Qout = A^-1 Qin through the factors; Qin may be overwritten (reference matmul_inverse_tallskinny_blocked,
//! kernels_cpp.hh:660-755)
template <typename MV, typename MAT>
void matmul_inverse_tallskinny_blocked(MV &Qout, UMFPackFactorizedMatrix<MAT> &F, MV &Qin)
{
  de_b200::require_block8<MV>("matmul_inverse_tallskinny_blocked");
  if (Qout.rows() != Qin.rows() || Qout.cols() != Qin.cols())
    throw std::invalid_argument("matmul_inverse_tallskinny_blocked: Qout/Qin size mismatch");
  if ((std::size_t)F.n != Qin.rows() || (std::size_t)F.n != Qout.rows())
    throw std::invalid_argument("matmul_inverse_tallskinny_blocked: Factorization does not match size of Qout/Qin");
  auto &ctx = de_b200::Context::thread_default();
  de_factor *dF = nullptr;
  de_b200::check(F.upload(ctx.get(), &dF),
                 ctx.get());
  try
  {
    de_b200::DeviceMV dIn(ctx, Qin), dOut(ctx, Qout.rows(), Qout.cols());
    de_b200::check(de_factor_apply(dOut.get(), dF, dIn.get()), ctx.get());
    dOut.download(Qout);
  }
  catch (...)
  {
    de_factor_destroy(dF);
    throw;
  }
  de_factor_destroy(dF);
}

// ---- the reference's analytic cost models, kept as the reporting convention (kernels_cpp.hh:98-116, :157-175) ----
inline double flops_orthonormalize(int n, int m)
{
  double total = 0.0;
  for (int k = 1; k <= m; ++k)
    total += 3.0 * n + 4.0 * n * (k - 1); // norm + scaling of column k, projection of it from the k-1 earlier ones
  return total;
}

inline double bytes_orthonormalize_naive(int n, int m, int numbersize = 8)
{
  double words = 0.0;
  for (int k = 1; k <= m; ++k)
    words += 3.0 * n + 5.0 * n * (k - 1);
  return words * numbersize;
}

inline double bytes_orthonormalize_blocked(int n, int m, int b, int numbersize = 8)
{
  double words = 0.0;
  for (int first = 0; first < m; first += b)
  {
    for (int k = 1; k <= b; ++k) // the diagonal panel, column by column
      words += (double)n * k + (double)n * (k + 1);
    const int later_panels = (m - first - 1) / b;
    words += 5.0 * b * n * later_panels; // each later panel: read both, read both again, write one
  }
  return words * numbersize;
}

#endif
