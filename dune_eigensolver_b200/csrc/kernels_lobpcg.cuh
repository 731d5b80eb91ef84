// Kernels of the LOBPCG drivers (lobpcg_core.hpp) that the subspace-iteration path did not need:
//   lincomb_kernel   X <- X Cx + W Cw + P Cp  and  P <- W Cw + P Cp  in ONE pass over the three blocks
//   residual_kernel  W = A X - B X diag(theta)
//   cheb_start_kernel / cheb_step_kernel / gershgorin_kernel   the Jacobi-scaled Chebyshev polynomial preconditioner
//                    (three-term recurrence between two SpMMs: 40*n*m bytes per step, pure streaming), its diagonal
//                    scale and its spectral bound
// Everything else of an LOBPCG iteration (SpMM, Gram, CholQR2, projection) reuses the kernels of the reference path.
// No reference counterpart: normallytangent/dune-eigensolver has no LOBPCG; the closest relatives are
// its block update V <- V U (kernels_cpp.hh:293-305) and projection Q_j -= Q_k S (:335-348).
//
// Roofline: lincomb moves 8*n*m*(ns + 2) bytes (ns sources read once, two outputs written) for 2*n*m^2*(ns) flops:
// HBM-bound up to m = 16, FP64-pipe-bound from m = 32 (3 products: 6*n*m^2 flops, AI = 6m/40 flop/B = 4.8 at m = 32
// against a ridge of 34.8 TFLOP/s / 6.5 TB/s = 5.3). residual: 32*n*m bytes, pure streaming.
#pragma once

#include <cuda_runtime.h>

#include "kernels_sparse.cuh"

namespace de
{

  template <int M>
  struct LinCfg
  {
    static constexpr int CT = 8;                               // output columns per thread
    static constexpr int RT = M >= 48 ? 4 : (M >= 24 ? 2 : 1); // rows per thread
    static constexpr int TR = 128;                             // rows per tile
    static constexpr int NCG = M / CT;                         // column groups
    static constexpr int NRB = TR / (32 * RT);                 // 32*RT-row blocks per tile
    static constexpr int WARPS = NCG * NRB;                    // one (row block, column group) unit per warp: 4..10
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int LDS = M + 1;                          // odd stride: the 32 rows of a warp hit distinct banks
    static constexpr int MAXSRC = 3;
    static constexpr size_t SMEM_BYTES = sizeof(double) * ((size_t)TR * LDS + (size_t)MAXSRC * M * M);
  };

  /** out = sum_{s < ns} S_s C_s ;  out2 = sum_{1 <= s < ns} S_s C_s  (out2 may be null; ignored for ns = 1).
   *  S_s: row-major n x M blocks with leading dimension M; C: ns row-major M x M matrices, contiguous, device memory.
   *  A warp owns (32*RT rows) x (8 columns) of both outputs for the whole tile; the sources are staged one after the
   *  other (last source first) in the same shared-memory tile and accumulated in registers, so every source row is
   *  read from HBM exactly once. out2 is written after the sources s >= 1, out after source 0.
   *  Aliasing: out may be S_0 and out2 may be any S_s with s >= 1 (not S_0): an output row is written only after this
   *  CTA has staged that row of every source that aliases it, and different CTAs own different rows. */
  template <int M>
  __global__ void __launch_bounds__(LinCfg<M>::THREADS)
      lincomb_kernel(long long n, int ns, const double *S0, const double *S1, const double *S2,
                     const double *__restrict__ C, double *out, double *out2)
  {
    using K = LinCfg<M>;
    extern __shared__ __align__(16) double dyn_smem[];
    double *Xs = dyn_smem;                  // TR x LDS
    double *Cs = dyn_smem + K::TR * K::LDS; // ns x M x M

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < ns * M * M; e += K::THREADS)
      Cs[e] = __ldg(C + e);

    const int cg = warp % K::NCG, rb = warp / K::NCG;
    const int c0 = cg * K::CT;
    const long long ntiles = (n + K::TR - 1) / K::TR;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
    {
      const long long r0 = tile * K::TR;
      double acc[K::RT][K::CT];
#pragma unroll
      for (int q = 0; q < K::RT; ++q)
#pragma unroll
        for (int c = 0; c < K::CT; ++c)
          acc[q][c] = 0.0;

      for (int s = ns - 1; s >= 0; --s)
      {
        const double *src = s == 0 ? S0 : (s == 1 ? S1 : S2);
        __syncthreads(); // readers of the previous staged tile are done (also orders the Cs fill on the first pass)
        for (int e = tid; e < K::TR * (M / 2); e += K::THREADS)
        {
          const int r = e / (M / 2), c = 2 * (e % (M / 2));
          double2 v = make_double2(0.0, 0.0);
          if (r0 + r < n)
            v = ld2(src + (size_t)(r0 + r) * M + c);
          Xs[r * K::LDS + c] = v.x;
          Xs[r * K::LDS + c + 1] = v.y;
        }
        __syncthreads();

        const double *xrow = Xs + (rb * 32 * K::RT + lane) * K::LDS;
        const double *cs = Cs + (size_t)s * M * M + c0;
#pragma unroll 2
        for (int k = 0; k < M; ++k)
        {
          double rr[K::CT];
#pragma unroll
          for (int c = 0; c < K::CT; c += 2)
          {
            const double2 v = ld2(cs + k * M + c);
            rr[c] = v.x;
            rr[c + 1] = v.y;
          }
#pragma unroll
          for (int q = 0; q < K::RT; ++q)
          {
            const double xk = xrow[q * 32 * K::LDS + k];
#pragma unroll
            for (int c = 0; c < K::CT; ++c)
              acc[q][c] = fma(xk, rr[c], acc[q][c]);
          }
        }

        double *dst = (s == 0) ? out : ((s == 1 && out2 != nullptr) ? out2 : nullptr);
        if (dst != nullptr)
        {
#pragma unroll
          for (int q = 0; q < K::RT; ++q)
          {
            const long long row = r0 + rb * 32 * K::RT + q * 32 + lane;
            if (row < n)
            {
              double *y = dst + (size_t)row * M + c0;
#pragma unroll
              for (int c = 0; c < K::CT; c += 2)
                st2(y + c, make_double2(acc[q][c], acc[q][c + 1]));
            }
          }
        }
      }
    }
  }

  /** W = AX - BX diag(theta): one thread per pair of adjacent columns, grid-stride over the n*m/2 pairs */
  __global__ void __launch_bounds__(256)
      residual_kernel(long long pairs, int m, const double *__restrict__ AX, const double *__restrict__ BX,
                      const double *__restrict__ theta, double *__restrict__ W)
  {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < pairs; e += stride)
    {
      const int c = (int)((2 * e) % m);
      const double2 a = ld2(AX + 2 * e), b = ld2(BX + 2 * e);
      const double t0 = __ldg(theta + c), t1 = __ldg(theta + c + 1);
      st2(W + 2 * e, make_double2(fma(-t0, b.x, a.x), fma(-t1, b.y, a.y)));
    }
  }

  // The three kernels below use a (m/2, 256/(m/2)) thread block: threadIdx.x walks the double2 pairs of a row,
  // threadIdx.y the rows, so consecutive threads touch consecutive 16-byte words of the row-major block and the
  // per-row Jacobi scale dinv[i] = 1 / a_ii is one load per row -- no index division.

  /** Z = s D^-1 R ; Zold = 0  (first iterate of the Jacobi-scaled Chebyshev iteration, z_1 = D^-1 r / theta, z_0 = 0) */
  __global__ void __launch_bounds__(256)
      cheb_start_kernel(long long n, int hp, double s, const double *__restrict__ dinv, const double *__restrict__ R,
                        double *__restrict__ Z, double *__restrict__ Zold)
  {
    const long long rstride = (long long)gridDim.x * blockDim.y;
    for (long long i = (long long)blockIdx.x * blockDim.y + threadIdx.y; i < n; i += rstride)
    {
      const double sc = s * __ldg(dinv + i);
      const size_t e = ((size_t)i * hp + threadIdx.x) * 2;
      const double2 r = ld2(R + e);
      st2(Z + e, make_double2(sc * r.x, sc * r.y));
      st2(Zold + e, make_double2(0.0, 0.0));
    }
  }

  /** Zold <- Z + alpha (Z - Zold) + beta D^-1 (R - AZ): the next Chebyshev iterate overwrites the one before the current */
  __global__ void __launch_bounds__(256)
      cheb_step_kernel(long long n, int hp, double alpha, double beta, const double *__restrict__ dinv,
                       const double *__restrict__ Z, const double *__restrict__ R, const double *__restrict__ AZ,
                       double *__restrict__ Zold)
  {
    const long long rstride = (long long)gridDim.x * blockDim.y;
    for (long long i = (long long)blockIdx.x * blockDim.y + threadIdx.y; i < n; i += rstride)
    {
      const double bd = beta * __ldg(dinv + i);
      const size_t e = ((size_t)i * hp + threadIdx.x) * 2;
      const double2 z = ld2(Z + e), zo = ld2(Zold + e), r = ld2(R + e), a = ld2(AZ + e);
      st2(Zold + e, make_double2(z.x + alpha * (z.x - zo.x) + bd * (r.x - a.x),
                                 z.y + alpha * (z.y - zo.y) + bd * (r.y - a.y)));
    }
  }

  /** dinv[i] = 1 / a_ii and out[0] = max(out[0], max_i sum_k |a_ik| / a_ii) over this rank's rows: the Jacobi scale and
   *  the Gershgorin bound of the spectrum of D^-1 A (equal to that of D^-1/2 A D^-1/2). Row i's diagonal entry is the
   *  one with column index i (distributed matrices number their owned columns first). A row without a positive
   *  diagonal entry makes the bound +inf, which the caller reports. Non-negative doubles compare like their bit
   *  patterns, so the maximum is an integer atomicMax. */
  __global__ void __launch_bounds__(256)
      gershgorin_kernel(long long n, const int *__restrict__ rowptr, const int *__restrict__ col,
                        const double *__restrict__ val, double *__restrict__ dinv, unsigned long long *out)
  {
    double best = 0.0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    {
      double sum = 0.0, d = 0.0;
      for (int k = rowptr[i]; k < rowptr[i + 1]; ++k)
      {
        const double v = val[k];
        sum += fabs(v);
        if (col[k] == i)
          d = v;
      }
      const bool ok = d > 0.0;
      dinv[i] = ok ? 1.0 / d : 1.0;
      best = fmax(best, ok ? sum / d : __longlong_as_double(0x7ff0000000000000LL));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      best = fmax(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best > 0.0)
      atomicMax(out, (unsigned long long)__double_as_longlong(best));
  }

} // namespace de
