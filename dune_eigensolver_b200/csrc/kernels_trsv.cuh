// Level-scheduled sparse triangular solves with m right-hand sides (K6) for sm_100a:
// the factored inverse apply  X <- Q U^-1 L^-1 P R X  of reference kernels_cpp.hh:660-755.
//
// At upload (host, once) the UMFPACK-style factors are re-organised: L (CSR, unit diagonal stripped) and
// U (CSC -> CSR, diagonal stripped and inverted) get a level schedule, rows sorted by level. The apply works
// in place on one n x m row-major work block W:
//   1. W(k,:) = rowscale(k) * X(P[k],:)                        (kernels_cpp.hh:682-705)
//   2. forward sweep over the levels of L:  W(i,:) -= sum_j L(i,j) W(j,:)   (kernels_cpp.hh:710-728)
//   3. backward sweep over the levels of U: W(i,:) = (W(i,:) - sum_j U(i,j) W(j,:)) / U(i,i)   (:732-746)
//   4. Y(Q[j],:) = W(j,:)                                       (kernels_cpp.hh:747-750)
// Rows inside a level are independent. Wide levels get one launch each (a warp per row); runs of consecutive
// narrow levels (the sequential top of the elimination tree) are chained inside ONE CTA with __syncthreads
// between levels and the row's nonzeros split over several warps.
//
// Roofline: HBM-bound on streaming the factors once for all m columns: 12*(lnz+unz) + 16*n + 16*n*m bytes
// (the reference re-streams L and U once per 8-column panel, kernels_cpp.hh:676).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_sparse.cuh"

namespace de
{

  /** W(k,:) = scale[k] * X(perm[k],:)   or (scatter != 0)   Y(perm[k],:) = W(k,:) */
  static __global__ void __launch_bounds__(256) permute_rows_kernel(long long n, int m, const int *__restrict__ perm,
                                                             const double *__restrict__ scale,
                                                             const double *__restrict__ src,
                                                             double *__restrict__ dst, int scatter)
  {
    const int hp = m / 2;
    const long long total = n * hp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x)
    {
      const long long k = e / hp;
      const int c = 2 * (int)(e % hp);
      const long long p = perm[k];
      if (scatter)
        st2(dst + (size_t)p * m + c, ldg2(src + (size_t)k * m + c));
      else
      {
        double2 v = ldg2(src + (size_t)p * m + c);
        const double s = scale[k];
        v.x *= s;
        v.y *= s;
        st2(dst + (size_t)k * m + c, v);
      }
    }
  }

  struct TrsvArgs
  {
    const int *rows;      // row indices sorted by level
    const int *rowptr;    // CSR of the strictly triangular part (original row numbering)
    const int *col;
    const double *val;
    const double *invdiag; // null for the unit-diagonal L
    double *W;             // n x m work block, updated in place
    int m;
  };

  /** accumulate -sum_k val[k] * W(col[k], cpair) over the slice k = kbeg + part, kbeg + part + nparts, ... */
  template <int LC>
  __device__ __forceinline__ double2 trsv_row_partial(const TrsvArgs &a, int kbeg, int kend, int part, int nparts,
                                                      int c, bool active)
  {
    double2 acc = make_double2(0.0, 0.0);
    int k = kbeg + part;
    for (; k + 3 * nparts < kend; k += 4 * nparts)
    {
      int j[4];
      double v[4];
      double2 w[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        j[u] = a.col[k + u * nparts];
        v[u] = a.val[k + u * nparts];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        w[u] = active ? ld2(a.W + (size_t)j[u] * a.m + c) : make_double2(0.0, 0.0);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        fma2(acc, -v[u], w[u]);
    }
    for (; k < kend; k += nparts)
    {
      const int j = a.col[k];
      const double v = a.val[k];
      if (active)
        fma2(acc, -v, ld2(a.W + (size_t)j * a.m + c));
    }
    return acc;
  }

  /** One level, one warp per row. LC lanes cover the column pairs (LC = pow2 >= m/2, <= 32), the remaining
   *  32/LC lane groups split the row's nonzeros and are combined with shuffles. */
  template <int LC>
  static __global__ void __launch_bounds__(256) trsv_level_kernel(const TrsvArgs a, int first, int count)
  {
    constexpr int NS = 32 / LC;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= count)
      return;
    const int cl = lane % LC, part = lane / LC;
    const int c = 2 * cl;
    const bool active = c < a.m;
    const int i = a.rows[first + warp];
    double2 acc = trsv_row_partial<LC>(a, a.rowptr[i], a.rowptr[i + 1], part, NS, c, active);
#pragma unroll
    for (int off = LC; off < 32; off <<= 1)
    {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
      acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    }
    if (part == 0 && active)
    {
      double *w = a.W + (size_t)i * a.m + c;
      double2 r = ld2(w);
      r.x += acc.x;
      r.y += acc.y;
      if (a.invdiag)
      {
        const double d = a.invdiag[i];
        r.x *= d;
        r.y *= d;
      }
      st2(w, r);
    }
  }

  /** A run of consecutive narrow levels [lev_begin, lev_end) in ONE CTA of 32 warps. A level with R <= 32 rows
   *  gives each row 32/R warps (rounded down to a power of two); partial sums meet in shared memory.
   *  W is read with plain (coherent) loads: rows written before the barrier are consumed after it. */
  template <int LC>
  static __global__ void __launch_bounds__(1024) trsv_chain_kernel(const TrsvArgs a, const int *__restrict__ level_ptr,
                                                            int lev_begin, int lev_end)
  {
    constexpr int NS = 32 / LC;
    __shared__ double2 part_sum[32][LC]; // [warp][column pair]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cl = lane % LC, part = lane / LC;
    const int c = 2 * cl;
    const bool active = c < a.m;
    for (int lev = lev_begin; lev < lev_end; ++lev)
    {
      const int first = level_ptr[lev];
      const int R = level_ptr[lev + 1] - first; // 1..32
      int wpr = 1;
      while (wpr * 2 * R <= 32)
        wpr *= 2;
      const int myrow = warp / wpr, wpart = warp % wpr;
      const bool has_row = myrow < R;
      int i = 0;
      if (has_row)
      {
        i = a.rows[first + myrow];
        double2 acc = trsv_row_partial<LC>(a, a.rowptr[i], a.rowptr[i + 1], wpart * NS + part, wpr * NS, c, active);
#pragma unroll
        for (int off = LC; off < 32; off <<= 1)
        {
          acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
          acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
        }
        if (part == 0)
          part_sum[warp][cl] = acc;
      }
      __syncthreads();
      if (has_row && wpart == 0 && part == 0 && active)
      {
        double *w = a.W + (size_t)i * a.m + c;
        double2 r = ld2(w);
        for (int q = 0; q < wpr; ++q) // fixed order
        {
          r.x += part_sum[warp + q][cl].x;
          r.y += part_sum[warp + q][cl].y;
        }
        if (a.invdiag)
        {
          const double d = a.invdiag[i];
          r.x *= d;
          r.y *= d;
        }
        st2(w, r);
      }
      __syncthreads(); // results of this level are visible to the whole CTA before the next level reads them
    }
  }

} // namespace de
