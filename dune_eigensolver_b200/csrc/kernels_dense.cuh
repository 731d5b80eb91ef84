// Tall-skinny dense kernels for sm_100a: Gram matrices (K3), block update / projection (K4) and the tiny
// m x m Cholesky + triangular inverse that turns them into (B-)orthonormalisation (K5).
//
// All operands are row-major n x M fp64 blocks (or column views of them, given by pointer + leading dimension).
// Roofline: Gram 8*n*m (X = Y) or 16*n*m bytes, 2*n*m^2 flops; update 16*n*m bytes, 2*n*m^2
// flops (half for a triangular factor). HBM-bound for m <= 32; at m = 64 the arithmetic intensity m/8 flop/B
// sits at the FP64 ridge, so the kernels use 4x4 / 4x8 register tiles fed from shared memory to keep the FP64
// pipe (not the LSU) the second limiter. tcgen05 has no FP64 kind, so there is no tensor-core path here.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_sparse.cuh"

namespace de
{

  constexpr int ilog2_floor(int x) { return x <= 1 ? 0 : 1 + ilog2_floor(x / 2); }
  constexpr int pow2_floor(int x) { return 1 << ilog2_floor(x); }

  // ------------------------------------------------------------------------------------------------
  // Gram:  G = X^T Y  (M x M), reduction over the n rows
  // ------------------------------------------------------------------------------------------------
  template <int M, bool UPPER, bool SAME>
  struct GramCfg
  {
    static constexpr int NB = M / 4;                                 // 4x4 output blocks per dimension
    static constexpr int NT = UPPER ? NB * (NB + 1) / 2 : NB * NB;     // threads that tile G once
    static constexpr int RG = pow2_floor(256 / NT) < 1 ? 1 : pow2_floor(256 / NT); // row groups (split of the tile rows)
    static constexpr int THREADS = NT * RG;
    static constexpr int TR0 = 2048 / M;                             // ~16 KB of X per tile
    static constexpr int TR = ((TR0 + RG - 1) / RG) * RG;            // tile rows, multiple of RG
    static constexpr int TILE = TR * M;                              // doubles per operand tile
    static constexpr int SMEM_DOUBLES = (SAME ? TILE : 2 * TILE) > THREADS * 16 ? (SAME ? TILE : 2 * TILE) : THREADS * 16;
  };

  /** Per-CTA partial Gram matrices. Each thread owns a 4x4 block of G for one residue class of tile rows.
   *  UPPER (G known to be symmetric: X^T X of CholQR, X^T (B X) of the B variant; cf. the upper-triangle
   *  products of kernels_cpp.hh:236-242, :451-463) enumerates only the blocks on or above the diagonal and
   *  mirrors them on output, halving the FP64 work. SAME: Y aliases X, only one operand tile is staged.
   *  Rows beyond n are zero-filled. partials: [gridDim.x][M*M], reduced afterwards in fixed order. */
  template <int M, bool UPPER, bool SAME>
  __global__ void __launch_bounds__(GramCfg<M, UPPER, SAME>::THREADS)
      gram_kernel(long long n, const double *__restrict__ X, int ldx, const double *__restrict__ Y, int ldy,
                  double *__restrict__ partials)
  {
    using C = GramCfg<M, UPPER, SAME>;
    __shared__ __align__(16) double sm[C::SMEM_DOUBLES];
    double *Xs = sm;
    double *Ys = SAME ? sm : sm + C::TILE;

    const int tid = threadIdx.x;
    const int g = tid / C::NT; // row group
    const int t = tid % C::NT; // block id
    int bi, bj;
    if (UPPER)
    {
      int rem = t;
      bi = 0;
      while (rem >= C::NB - bi)
      {
        rem -= C::NB - bi;
        ++bi;
      }
      bj = bi + rem;
    }
    else
    {
      bi = t / C::NB;
      bj = t % C::NB;
    }

    double acc[4][4];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        acc[p][q] = 0.0;

    constexpr int HP = M / 2;
    const long long ntiles = (n + C::TR - 1) / C::TR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
      const long long r0 = tile * C::TR;
      for (int e = tid; e < C::TR * HP; e += C::THREADS)
      {
        const int r = e / HP, c = 2 * (e % HP);
        const bool in = r0 + r < n;
        st2(Xs + r * M + c, in ? ldg2(X + (size_t)(r0 + r) * ldx + c) : make_double2(0.0, 0.0));
        if (!SAME)
          st2(Ys + r * M + c, in ? ldg2(Y + (size_t)(r0 + r) * ldy + c) : make_double2(0.0, 0.0));
      }
      __syncthreads();
#pragma unroll 4
      for (int r = g; r < C::TR; r += C::RG)
      {
        const double2 x01 = ld2(Xs + r * M + 4 * bi), x23 = ld2(Xs + r * M + 4 * bi + 2);
        const double2 y01 = ld2(Ys + r * M + 4 * bj), y23 = ld2(Ys + r * M + 4 * bj + 2);
        const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
        const double yv[4] = {y01.x, y01.y, y23.x, y23.y};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            acc[p][q] = fma(xv[p], yv[q], acc[p][q]);
      }
      __syncthreads();
    }

    // combine the row groups in fixed order, then emit this CTA's partial
    double *red = sm; // THREADS*16 doubles, tile storage is dead now
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        red[(p * 4 + q) * C::THREADS + tid] = acc[p][q];
    __syncthreads();
    if (g == 0)
    {
      double *out = partials + (size_t)blockIdx.x * M * M;
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q)
        {
          double s = 0.0;
          for (int gg = 0; gg < C::RG; ++gg)
            s += red[(p * 4 + q) * C::THREADS + gg * C::NT + t];
          out[(4 * bi + p) * M + 4 * bj + q] = s;
          if (UPPER && bi != bj)
            out[(4 * bj + q) * M + 4 * bi + p] = s;
        }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // Block update:  Y = X R   or   Y -= X R      (X: n x M view, R: M x M row-major, Y: n x M view)
  // ------------------------------------------------------------------------------------------------
  template <int M>
  struct UpdCfg
  {
    static constexpr int CT = 8;                               // output columns per thread
    static constexpr int RT = M >= 48 ? 4 : (M >= 24 ? 2 : 1); // rows per thread
    static constexpr int TR = 128;                             // rows per tile
    static constexpr int NCG = M / CT;                         // column groups
    static constexpr int NRB = TR / (32 * RT);                 // 32*RT-row blocks per tile
    static constexpr int UNITS = NCG * NRB;                    // warp work units per tile
    static constexpr int WARPS = UNITS < 8 ? UNITS : 8;
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int LDS = M + 1;                          // odd stride: 32 rows at one k hit 32 banks pairs
    static constexpr size_t SMEM_BYTES = sizeof(double) * ((size_t)TR * LDS + (size_t)M * M);
  };

  /** One warp owns (32*RT rows) x (8 columns) of the output tile: lane = row, so the R(k, 8 cols) operand is a
   *  warp-wide shared-memory broadcast and the X(row,k) operand is conflict-free thanks to the odd row stride.
   *  For an upper-triangular factor (`upper`, the R^-1 of CholQR and U of kernels_cpp.hh:293-305) the k loop
   *  of column group c stops at 8c+7, which halves the FP64 work. The X tile is fully staged in shared memory
   *  before any output is written, so Y may alias X (in-place X <- X R).
   *  MODE 0: Y = X R.  MODE 1: Y -= X R (projection kernels_cpp.hh:335-348; X and Y are disjoint column views). */
  template <int M, int MODE>
  __global__ void __launch_bounds__(UpdCfg<M>::THREADS)
      update_kernel(long long n, const double *X, int ldx, const double *__restrict__ R, double *Y, int ldy,
                    int upper)
  {
    using C = UpdCfg<M>;
    extern __shared__ __align__(16) double dyn_smem[];
    double *Xs = dyn_smem;               // TR x LDS
    double *Rs = dyn_smem + C::TR * C::LDS; // M x M

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < M * M; e += C::THREADS)
      Rs[e] = __ldg(R + e);

    const long long ntiles = (n + C::TR - 1) / C::TR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
      const long long r0 = tile * C::TR;
      __syncthreads(); // previous tile's readers are done (also orders the Rs fill on the first pass)
      for (int e = tid; e < C::TR * (M / 2); e += C::THREADS)
      {
        const int r = e / (M / 2), c = 2 * (e % (M / 2));
        double2 v = make_double2(0.0, 0.0);
        if (r0 + r < n)
          v = ld2(X + (size_t)(r0 + r) * ldx + c);
        Xs[r * C::LDS + c] = v.x;
        Xs[r * C::LDS + c + 1] = v.y;
      }
      __syncthreads();

      for (int unit = warp; unit < C::UNITS; unit += C::WARPS)
      {
        const int cg = unit % C::NCG, rb = unit / C::NCG;
        const int c0 = cg * C::CT;
        const int kmax = upper ? (c0 + C::CT < M ? c0 + C::CT : M) : M;
        double acc[C::RT][C::CT];
#pragma unroll
        for (int q = 0; q < C::RT; ++q)
#pragma unroll
          for (int c = 0; c < C::CT; ++c)
            acc[q][c] = 0.0;
        const double *xrow = Xs + (rb * 32 * C::RT + lane) * C::LDS;
#pragma unroll 2
        for (int k = 0; k < kmax; ++k)
        {
          double rr[C::CT];
#pragma unroll
          for (int c = 0; c < C::CT; c += 2)
          {
            const double2 v = ld2(Rs + k * M + c0 + c);
            rr[c] = v.x;
            rr[c + 1] = v.y;
          }
#pragma unroll
          for (int q = 0; q < C::RT; ++q)
          {
            const double xk = xrow[q * 32 * C::LDS + k];
#pragma unroll
            for (int c = 0; c < C::CT; ++c)
              acc[q][c] = fma(xk, rr[c], acc[q][c]);
          }
        }
#pragma unroll
        for (int q = 0; q < C::RT; ++q)
        {
          const long long row = r0 + rb * 32 * C::RT + q * 32 + lane;
          if (row < n)
          {
            double *y = Y + (size_t)row * ldy + c0;
#pragma unroll
            for (int c = 0; c < C::CT; c += 2)
            {
              double2 v = make_double2(acc[q][c], acc[q][c + 1]);
              if (MODE == 1)
              {
                const double2 o = ld2(y + c);
                v.x = o.x - v.x;
                v.y = o.y - v.y;
              }
              st2(y + c, v);
            }
          }
        }
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // m x m Cholesky + inverse of the triangular factor (one CTA; replicated on every GPU of a multi-GPU run)
  // ------------------------------------------------------------------------------------------------
  /** G = R^T R (R upper, positive diagonal), Rinv = R^-1 (upper); only the upper triangle of G is read. This is the
   *  L D L^T / U = L^-T D^-1/2 construction of the reference (kernels_cpp.hh:247-291, :468-512) for the whole block.
   *  status[0] (sticky) = 1 + index of the first pivot that is non-finite or not above 4 m eps G(k,k); untouched on success.
   *  info[0] (optional) = largest strict-upper entry of G (the `norm` the reference's B-orthonormalisation returns,
   *  kernels_cpp.hh:464-466); identity_flag[0] (optional) = 1 iff max |G - I| <= 1e-14 over the upper triangle, in which
   *  case nothing is factored (Rinv = I).
   *  Cholesky + triangular inverse of the (all-reduced) Gram matrix: Rinv = R^-1 with G = R^T R, R upper triangular with positive diagonal; a pivot
   *  <= 4 m eps G_kk reports rank deficiency (status = k + 1, Rinv = I, `done` raised).
   *  ncu (profiles/r01_ncu_launches_brb.csv) had the first version (a shared-memory right-looking factorisation, kept in
   *  the history) at 30 us per call, 13 % of a StandardLargest iteration: integer divisions in the trailing update, three
   *  CTA barriers per pivot and a 32-thread back substitution. Here the matrix is padded with the identity to MP x MP (MP = 32 or 64), every thread keeps its
   *  (MP/32)^2 elements in registers, row k lives in ONE warp (its pivot is a warp shuffle away), and both phases are
   *  rank-1 updates with ONE barrier per pivot:
   *    factorisation   scaled row k -> shared; barrier; a_ij -= r_ki r_kj          (i > k)
   *    inverse         Gauss-Jordan from the bottom: x_k /= r_kk; x_i -= r_ik x_k   (i < k), X starts as I
   *  1024 threads: element (i, j) = (w + 32 a, l + 32 b) belongs to warp w, lane l, slot (a, b). */
  /** wellcond (optional, int[2]) and wc_identity (int[1]): the caller asks whether ONE CholQR sweep with this Gram matrix is
   *  enough. With C = D^-1/2 G D^-1/2 (unit diagonal; Cholesky is invariant under this scaling to first order) the eigenvalues
   *  of C lie within ||C - I||_F of 1, so ||C - I||_F <= kWellCond bounds cond(C) by (1 + d) / (1 - d) = 3: the orthogonality
   *  defect of X chol(G)^-1 is then a small multiple of the rounding error of G itself -- what the second sweep's own test
   *  (max |G2 - I| <= 1e-14 on a computed G2) would accept. wellcond[0] = yes, wellcond[1] = no, wc_identity[0] = yes (the
   *  flag the second sweep's update skips itself on). */
  constexpr double kWellCond = 0.5;

  template <int MP>
  __device__ __forceinline__ void chol_inverse2_body(int tid, int m, const double *__restrict__ G, double *__restrict__ Rinv,
                                                     int *__restrict__ status, double *__restrict__ info,
                                                     int *__restrict__ identity_flag, int *__restrict__ done,
                                                     int *__restrict__ wellcond = nullptr, int *__restrict__ wc_identity = nullptr)
  {
    constexpr int E = MP / 32;
    __shared__ double rowk[2][MP];
    __shared__ double colk[2][MP];
    __shared__ double dorig[MP];
    __shared__ double dinv[MP]; // 1 / r_kk
    __shared__ double redmx[32], reddev[32];
    __shared__ int bad;
    const int lane = tid & 31, warp = tid >> 5;
    if (tid == 0)
      bad = 0;

    double r[E][E], x[E][E];
    double mx = -1.0e300, dev = 0.0;
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
      for (int b = 0; b < E; ++b)
      {
        const int i = warp + 32 * a, j = lane + 32 * b;
        double v = (i == j) ? 1.0 : 0.0; // identity padding
        if (i < m && j < m)
        {
          v = (i <= j) ? __ldcg(G + i * m + j) : 0.0; // L2: G may have been written by other CTAs of this launch
          if (i < j)
            mx = fmax(mx, v);
          if (i <= j)
            dev = fmax(dev, fabs(v - (i == j ? 1.0 : 0.0)));
          if (!(v == v))
            dev = 1.0e300; // NaN: never "identity"
        }
        r[a][b] = v;
        x[a][b] = (i == j) ? 1.0 : 0.0;
        if (i == j)
          dorig[i] = v;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    }
    if (lane == 0)
    {
      redmx[warp] = mx;
      reddev[warp] = dev;
    }
    __syncthreads();
    if (tid == 0)
    {
      double t = -1.0e300, u = 0.0;
      for (int q = 0; q < 32; ++q)
      {
        t = fmax(t, redmx[q]);
        u = fmax(u, reddev[q]);
      }
      if (info != nullptr)
        info[0] = (m > 1) ? t : 0.0;
      if (identity_flag != nullptr)
        identity_flag[0] = (u <= 1.0e-14) ? 1 : 0;
      bad = (identity_flag != nullptr && u <= 1.0e-14) ? -1 : 0; // -1: G = I to working precision, nothing to factor
    }
    __syncthreads();
    if (wellcond != nullptr)
    {
      // ||C - I||_F^2 = 2 sum_{i < j} G_ij^2 / (G_ii G_jj); dorig[] was filled above, redmx[] is free again
      double off = 0.0;
#pragma unroll
      for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b)
        {
          const int i = warp + 32 * a, j = lane + 32 * b;
          if (i < j && j < m)
          {
            const double den = dorig[i] * dorig[j];
            off += (den > 0.0) ? 2.0 * r[a][b] * r[a][b] / den : 1.0e300;
          }
        }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
        off += __shfl_xor_sync(0xffffffffu, off, o);
      __syncthreads(); // everyone has read redmx / reddev of the first reduction
      if (lane == 0)
        redmx[warp] = off;
      __syncthreads();
      if (tid == 0)
      {
        double t = 0.0;
        for (int q = 0; q < 32; ++q)
          t += redmx[q];
        const int ok = (t == t && t <= kWellCond * kWellCond) ? 1 : 0;
        wellcond[0] = ok;
        wellcond[1] = 1 - ok;
        if (wc_identity != nullptr)
          wc_identity[0] = ok;
      }
    }
    if (bad == -1)
    {
      // the update that would use the factor skips itself on the same flag; leave Rinv = I for any other reader
#pragma unroll
      for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b)
        {
          const int i = warp + 32 * a, j = lane + 32 * b;
          if (i < m && j < m)
            Rinv[i * m + j] = (i == j) ? 1.0 : 0.0;
        }
      return;
    }

    // ---- factorisation ----
    for (int k = 0; k < MP; ++k)
    {
      const int ka = k >> 5, kw = k & 31, buf = k & 1;
      if (warp == kw)
      {
        // this warp owns row k (slot a = ka); lane kw of slot b = ka holds the pivot
        const double d = __shfl_sync(0xffffffffu, r[E == 1 ? 0 : ka][E == 1 ? 0 : ka], kw);
        const bool ok = (d > 4.0 * m * 2.220446049250313e-16 * dorig[k]) && isfinite(d);
        // one reciprocal square root per pivot; the division and the square root of the textbook form are the
        // longest links of the dependency chain (ncu: 800 cycles per pivot with sqrt + div, mostly in those two)
        const double rinv = ok ? rsqrt(d) : 1.0;
        const double rkk = ok ? d * rinv : 1.0;
        if (lane == 0)
        {
          dinv[k] = rinv;
          if (!ok)
            bad = k + 1;
        }
#pragma unroll
        for (int b = 0; b < E; ++b)
        {
          const int j = lane + 32 * b;
          double v = r[E == 1 ? 0 : ka][b];
          v = (j == k) ? rkk : (j > k ? v * rinv : 0.0);
          r[E == 1 ? 0 : ka][b] = v;
          rowk[buf][j] = v;
        }
      }
      __syncthreads();
      if (bad != 0)
        break;
#pragma unroll
      for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b)
        {
          const int i = warp + 32 * a, j = lane + 32 * b;
          if (i > k && j >= i)
            r[a][b] = fma(-rowk[buf][i], rowk[buf][j], r[a][b]);
        }
    }
    __syncthreads();
    if (bad != 0)
    {
      if (tid == 0)
      {
        status[0] = bad; // sticky: success never clears an earlier failure; the host resets it per driver call
        if (done != nullptr)
          *done = 1;
      }
#pragma unroll
      for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b)
        {
          const int i = warp + 32 * a, j = lane + 32 * b;
          if (i < m && j < m)
            Rinv[i * m + j] = (i == j) ? 1.0 : 0.0;
        }
      return;
    }

    // ---- inverse: Gauss-Jordan from the last row up ----
    for (int k = MP - 1; k >= 0; --k)
    {
      const int ka = k >> 5, kw = k & 31, buf = k & 1;
      if (warp == kw)
      {
        const double rinv = dinv[k];
#pragma unroll
        for (int b = 0; b < E; ++b)
        {
          const double v = x[E == 1 ? 0 : ka][b] * rinv;
          x[E == 1 ? 0 : ka][b] = v;
          rowk[buf][lane + 32 * b] = v;
        }
      }
      if (lane == kw)
      {
        // column k of R: element (i, k) lives in lane kw of warp i % 32, slot (i / 32, ka)
#pragma unroll
        for (int a = 0; a < E; ++a)
          colk[buf][warp + 32 * a] = r[a][E == 1 ? 0 : ka];
      }
      __syncthreads();
#pragma unroll
      for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b)
        {
          const int i = warp + 32 * a;
          if (i < k)
            x[a][b] = fma(-colk[buf][i], rowk[buf][lane + 32 * b], x[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
      for (int b = 0; b < E; ++b)
      {
        const int i = warp + 32 * a, j = lane + 32 * b;
        if (i < m && j < m)
          Rinv[i * m + j] = (i <= j) ? x[a][b] : 0.0;
      }
  }

  template <int MP>
  static __global__ void __launch_bounds__(1024) chol_inverse2_kernel(int m, const double *__restrict__ G, double *__restrict__ Rinv,
                                                               int *__restrict__ status, double *__restrict__ info,
                                                               int *__restrict__ identity_flag, int *__restrict__ done)
  {
    if (done != nullptr && *done != 0)
      return;
    chol_inverse2_body<MP>(threadIdx.x, m, G, Rinv, status, info, identity_flag, done);
  }

} // namespace de
