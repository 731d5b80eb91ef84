// host_eig.hpp -- the small dense symmetric eigenproblems of the Rayleigh-Ritz step, on the host.
//
// BASELINE.json north_star: "The p x p Rayleigh-Ritz eigenproblem stays a tiny host LAPACK call". The reference has no
// Rayleigh-Ritz step (its drivers use diagonal Rayleigh quotients, eigensolver.hh:84-87), and no LAPACK
// can be linked in this image, so the two classic routines are written out here: Householder tridiagonalisation
// followed by the implicit QL iteration (the EISPACK tred2 / tql2 pair), plus the reduction of the generalized problem
// GA c = theta GB c to standard form through a Cholesky factor of the diagonally scaled GB.
// Sizes are k*m x k*m with k <= 3 blocks of m <= 64 columns (<= 192): a few MFLOP per call.
// Pure host code, no CUDA; exported for tests as de_host_sym_eig / de_host_sym_gen_eig (no GPU needed).
#pragma once

#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace de
{
namespace hosteig
{

  /** Householder reduction of the symmetric n x n matrix V (full storage) to tridiagonal form: on return d = diagonal,
   *  e[1..n-1] = sub-diagonal, V = the accumulated orthogonal transformation STORED TRANSPOSED (element (i,j) at
   *  V[j*n + i]): the algorithm walks down columns, which are contiguous that way (2x faster at n = 192), and the QL
   *  iteration below wants the transposed matrix anyway. The input is symmetric, so its storage order does not matter. */
  inline void tridiagonalize(int n, double *V, double *d, double *e)
  {
    auto at = [&](int i, int j) -> double & { return V[(size_t)j * n + i]; };
    for (int j = 0; j < n; ++j)
      d[j] = at(n - 1, j);
    for (int i = n - 1; i > 0; --i)
    {
      double scale = 0.0, h = 0.0;
      for (int k = 0; k < i; ++k)
        scale += std::abs(d[k]);
      if (scale == 0.0)
      {
        e[i] = d[i - 1];
        for (int j = 0; j < i; ++j)
        {
          d[j] = at(i - 1, j);
          at(i, j) = 0.0;
          at(j, i) = 0.0;
        }
      }
      else
      {
        for (int k = 0; k < i; ++k)
        {
          d[k] /= scale;
          h += d[k] * d[k];
        }
        double f = d[i - 1];
        double g = std::sqrt(h);
        if (f > 0)
          g = -g;
        e[i] = scale * g;
        h -= f * g;
        d[i - 1] = f - g;
        for (int j = 0; j < i; ++j)
          e[j] = 0.0;
        for (int j = 0; j < i; ++j)
        {
          f = d[j];
          at(j, i) = f;
          g = e[j] + at(j, j) * f;
          for (int k = j + 1; k <= i - 1; ++k)
          {
            g += at(k, j) * d[k];
            e[k] += at(k, j) * f;
          }
          e[j] = g;
        }
        f = 0.0;
        for (int j = 0; j < i; ++j)
        {
          e[j] /= h;
          f += e[j] * d[j];
        }
        const double hh = f / (h + h);
        for (int j = 0; j < i; ++j)
          e[j] -= hh * d[j];
        for (int j = 0; j < i; ++j)
        {
          f = d[j];
          g = e[j];
          for (int k = j; k <= i - 1; ++k)
            at(k, j) -= (f * e[k] + g * d[k]);
          d[j] = at(i - 1, j);
          at(i, j) = 0.0;
        }
      }
      d[i] = h;
    }
    for (int i = 0; i < n - 1; ++i)
    {
      at(n - 1, i) = at(i, i);
      at(i, i) = 1.0;
      const double h = d[i + 1];
      if (h != 0.0)
      {
        for (int k = 0; k <= i; ++k)
          d[k] = at(k, i + 1) / h;
        for (int j = 0; j <= i; ++j)
        {
          double g = 0.0;
          for (int k = 0; k <= i; ++k)
            g += at(k, i + 1) * at(k, j);
          for (int k = 0; k <= i; ++k)
            at(k, j) -= g * d[k];
        }
      }
      for (int k = 0; k <= i; ++k)
        at(k, i + 1) = 0.0;
    }
    for (int j = 0; j < n; ++j)
    {
      d[j] = at(n - 1, j);
      at(n - 1, j) = 0.0;
    }
    at(n - 1, n - 1) = 1.0;
    e[0] = 0.0;
  }

  /** implicit QL iteration on the tridiagonal (d, e); the rotations are applied to the ROWS of Vt (the transposed
   *  eigenvector matrix: a rotation then touches two contiguous rows, which the compiler vectorises).
   *  returns 0, or 1 if an eigenvalue needed more than 60 iterations */
  inline int tridiagonal_ql(int n, double *Vt, double *d, double *e)
  {
    for (int i = 1; i < n; ++i)
      e[i - 1] = e[i];
    e[n - 1] = 0.0;
    double f = 0.0, tst1 = 0.0;
    const double eps = std::ldexp(1.0, -52);
    for (int l = 0; l < n; ++l)
    {
      tst1 = std::max(tst1, std::abs(d[l]) + std::abs(e[l]));
      int mm = l;
      while (mm < n)
      {
        if (std::abs(e[mm]) <= eps * tst1)
          break;
        ++mm;
      }
      if (mm > l)
      {
        int iter = 0;
        do
        {
          if (++iter > 60)
            return 1;
          double g = d[l];
          double p = (d[l + 1] - g) / (2.0 * e[l]);
          double r = std::hypot(p, 1.0);
          if (p < 0)
            r = -r;
          d[l] = e[l] / (p + r);
          d[l + 1] = e[l] * (p + r);
          const double dl1 = d[l + 1];
          double h = g - d[l];
          for (int i = l + 2; i < n; ++i)
            d[i] -= h;
          f += h;
          p = d[mm];
          double c = 1.0, c2 = c, c3 = c;
          const double el1 = e[l + 1];
          double s = 0.0, s2 = 0.0;
          for (int i = mm - 1; i >= l; --i)
          {
            c3 = c2;
            c2 = c;
            s2 = s;
            g = c * e[i];
            h = c * p;
            r = std::hypot(p, e[i]);
            e[i + 1] = s * r;
            s = e[i] / r;
            c = p / r;
            p = c * d[i] - s * g;
            d[i + 1] = h + s * (c * g + s * d[i]);
            double *__restrict__ v0 = Vt + (size_t)i * n, *__restrict__ v1 = Vt + (size_t)(i + 1) * n;
            for (int k = 0; k < n; ++k)
            {
              const double hk = v1[k];
              v1[k] = s * v0[k] + c * hk;
              v0[k] = c * v0[k] - s * hk;
            }
          }
          p = -s * s2 * c3 * el1 * e[l] / dl1;
          e[l] = s * p;
          d[l] = c * p;
        } while (std::abs(e[l]) > eps * tst1);
      }
      d[l] = d[l] + f;
      e[l] = 0.0;
    }
    return 0;
  }

  /** A = V diag(w) V^T for the symmetric n x n matrix A (row-major; only read). w ascending; column j of V
   *  (V[i*n + j], i = 0..n-1) is the unit eigenvector of w[j]. returns 0 on success. */
  inline int sym_eig(int n, const double *A, double *w, double *V)
  {
    if (n <= 0)
      return 0;
    std::vector<double> T((size_t)n * n), d(n), e(n);
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        T[(size_t)i * n + j] = 0.5 * (A[(size_t)i * n + j] + A[(size_t)j * n + i]);
    tridiagonalize(n, T.data(), d.data(), e.data());
    if (tridiagonal_ql(n, T.data(), d.data(), e.data()) != 0) // T now holds the eigenvectors as rows
      return 1;
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return d[a] < d[b]; });
    for (int j = 0; j < n; ++j)
    {
      w[j] = d[order[j]];
      for (int i = 0; i < n; ++i)
        V[(size_t)i * n + j] = T[(size_t)order[j] * n + i];
    }
    return 0;
  }

  /** GA c = theta GB c with GA symmetric, GB symmetric positive definite (both n x n row-major, only read).
   *  w ascending; column j of C is the eigenvector of w[j], normalised so that C^T GB C = I.
   *  GB is scaled to unit diagonal before its Cholesky factorisation; *min_pivot receives the smallest pivot of that
   *  scaled factorisation (its square is about 1 / cond of the scaled GB).
   *  returns 0 on success, 2 if GB is not (numerically) positive definite: a diagonal entry <= 0 or a pivot^2 below
   *  pivot_floor, 1 if the QL iteration failed. */
  inline int sym_gen_eig(int n, const double *GA, const double *GB, double *w, double *C, double pivot_floor,
                         double *min_pivot)
  {
    if (min_pivot)
      *min_pivot = 0.0;
    if (n <= 0)
      return 0;
    std::vector<double> dsc(n), L((size_t)n * n, 0.0), M((size_t)n * n), V((size_t)n * n);
    for (int i = 0; i < n; ++i)
    {
      const double b = GB[(size_t)i * n + i];
      if (!(b > 0.0) || !std::isfinite(b))
        return 2;
      dsc[i] = 1.0 / std::sqrt(b);
    }
    // Cholesky of the scaled GB (lower factor, row by row)
    double pmin = 1.0;
    for (int i = 0; i < n; ++i)
    {
      for (int j = 0; j <= i; ++j)
      {
        double s = 0.5 * (GB[(size_t)i * n + j] + GB[(size_t)j * n + i]) * dsc[i] * dsc[j];
        for (int k = 0; k < j; ++k)
          s -= L[(size_t)i * n + k] * L[(size_t)j * n + k];
        if (i == j)
        {
          if (!(s > pivot_floor) || !std::isfinite(s))
          {
            if (min_pivot)
              *min_pivot = s;
            return 2;
          }
          pmin = std::min(pmin, s);
          L[(size_t)i * n + i] = std::sqrt(s);
        }
        else
          L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
      }
    }
    if (min_pivot)
      *min_pivot = std::sqrt(pmin);
    // M = L^-1 (D GA D) L^-T : first T = L^-1 (D GA D) row by row (forward substitution on columns), then M = T L^-T
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        M[(size_t)i * n + j] = 0.5 * (GA[(size_t)i * n + j] + GA[(size_t)j * n + i]) * dsc[i] * dsc[j];
    // forward substitution on all columns at once (row operations: contiguous inner loops), twice with a transpose
    // in between: T = L^-1 (D GA D), then L^-1 T^T = (T L^-T)^T = M^T = M
    for (int pass = 0; pass < 2; ++pass)
    {
      for (int i = 0; i < n; ++i)
      {
        double *__restrict__ ri = M.data() + (size_t)i * n;
        for (int k = 0; k < i; ++k)
        {
          const double l = L[(size_t)i * n + k];
          const double *__restrict__ rk = M.data() + (size_t)k * n;
          for (int j = 0; j < n; ++j)
            ri[j] -= l * rk[j];
        }
        const double inv = 1.0 / L[(size_t)i * n + i];
        for (int j = 0; j < n; ++j)
          ri[j] *= inv;
      }
      for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
          std::swap(M[(size_t)i * n + j], M[(size_t)j * n + i]);
    }
    if (sym_eig(n, M.data(), w, V.data()) != 0)
      return 1;
    // C = D L^-T V : back substitution with L^T on all columns of V at once (row operations)
    for (int i = n - 1; i >= 0; --i)
    {
      double *__restrict__ ri = V.data() + (size_t)i * n;
      for (int k = i + 1; k < n; ++k)
      {
        const double l = L[(size_t)k * n + i];
        const double *__restrict__ rk = V.data() + (size_t)k * n;
        for (int j = 0; j < n; ++j)
          ri[j] -= l * rk[j];
      }
      const double inv = 1.0 / L[(size_t)i * n + i];
      for (int j = 0; j < n; ++j)
        ri[j] *= inv;
    }
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j)
        C[(size_t)i * n + j] = dsc[i] * V[(size_t)i * n + j];
    return 0;
  }

} // namespace hosteig
} // namespace de
