// Supernodal triangular solves with m right-hand sides: the factored inverse apply (K6; reference
// matmul_inverse_tallskinny_blocked, kernels_cpp.hh:660-755) for LARGE factors, where the factor is a supernodal
// Cholesky P A P^T = L L^T (include/dune/eigensolver/supernodal_cholesky.hh) instead of scalar rows.
//
// Data: every supernode is one dense column-major block (rows x columns, leading dimension rows) plus its row-index
// list; supernodes are cut into PANELS of at most kSnPanel columns, whose diagonal blocks are inverted once at upload
// (a w x w triangular solve becomes a small GEMM). A panel is the unit of the level schedule: the panels of one supernode
// form a chain, panels of independent subtrees share a level. One launch per level, one CTA per (panel, tile of 64
// rows below the panel):
// All products run on the FP64 tensor pipe (mma.sync m8n8k4, SASS DMMA) out of shared memory.
//   forward  (L z = y):    x_p = Dinv_p W_p (recomputed by every CTA of the panel: w^2 m flops, nothing to wait for);
//                          tile 0 stores x_p into Z; every CTA subtracts L(tile rows, panel) x_p from W -- plain stores for
//                          rows inside the same supernode (only this CTA touches them in this level), fp64 atomics for the
//                          update rows, which panels of sibling subtrees may hit in the same level
//   backward (L^T x = z):  every CTA adds -L(tile rows, panel)^T Z(tile rows) into Z_p with atomics; the CTA that finishes
//                          last (ticket) applies Dinv_p^T. The rows it reads belong to ancestors: final since an earlier level.
// The apply reads the factor once per sweep for all m columns: 8 (lnz_stored) bytes + 16 n m; it is FP64-pipe-bound
// (2 lnz m flops per sweep, 8 flop / byte at m = 64).
// The atomics make the summation order -- hence the last bits -- run-dependent; the reference's own apply has one fixed order.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_sparse.cuh"

namespace de
{

  constexpr int kSnPanel = 32; // columns per panel
  constexpr int kSnTile = 64;  // rows below the panel per CTA (128: 1.6x slower -- the top levels are latency-bound and want many short CTAs)
  constexpr int kSnThreads = 256;

  struct SnPanel
  {
    long long lofs; // val[lofs + a + j * r]: row a (local to the supernode, a >= j0) of panel column j
    long long rofs; // rowidx[rofs + a]: global row of the supernode's local row a
    long long dofs; // dinv[dofs + i * w + j]: inverse of the panel's diagonal block, row-major, lower triangular
    int r, ns;      // rows / columns of the supernode
    int j0, w;      // first column of the panel inside the supernode, its width
    int c0;         // global index of the panel's first row (= first column)
    int ntiles;     // CTAs of this panel (>= 1)
  };

  struct SnArgs
  {
    const SnPanel *panels;
    const int2 *items; // (panel, tile) of every CTA, sorted by level
    const double *val;
    const int *rowidx;
    const double *dinv;
    double *W, *Z;
    int *ticket; // one per panel (backward sweep), self-resetting
    int m;
  };

  // ---- shared-memory layout (doubles): all strides are = 4 (mod 16) so that the FP64 tensor-core operand fragments
  // (lane (g, k) reads [4 ks + k][8 b + g]) are conflict-free 64-bit loads, as in kernels_tallskinny2.cuh -----------------
  constexpr int kSnLdX = DE_KERNEL_MAX_M + 4; // sX / sZ: [k or row][column]
  constexpr int kSnLdL = kSnTile + 4;         // sL: [panel column][row of the tile]
  constexpr int kSnLdD = kSnPanel + 4;        // sD: [row][column] of the inverted diagonal block
  constexpr size_t kSnSmem = sizeof(double) * ((size_t)kSnPanel * kSnLdX + (size_t)kSnPanel * kSnLdL + (size_t)kSnTile * kSnLdX);

  __device__ __forceinline__ void sn_dmma(double &c0, double &c1, double a, double b)
  {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
  }

  /** forward sweep, one level: items [item0, item0 + gridDim.x). 8 warps; all products on the FP64 tensor pipe:
   *  D(8 x 8) += A(8 x 4) B(4 x 8), lane (g = lane / 4, k = lane % 4) holds A(g, k), B(k, g), D(g, 2k .. 2k + 1). */
  static __global__ void __launch_bounds__(kSnThreads) sn_forward_kernel(const SnArgs a, int item0)
  {
    extern __shared__ __align__(16) unsigned char sn_dyn[];
    double *sX = reinterpret_cast<double *>(sn_dyn);        // [kSnPanel][kSnLdX]: W_p, then x_p
    double *sL = sX + kSnPanel * kSnLdX;                     // [kSnPanel][kSnLdL]: L(tile rows, panel)^T ; first: Dinv^T
    double *sW = sL + kSnPanel * kSnLdL;                     // [kSnPanel][kSnLdX]: staging of W_p (part of the sZ area)
    const int2 it = a.items[item0 + blockIdx.x];
    const SnPanel P = a.panels[it.x];
    const int m = a.m, w = P.w, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, k = lane & 3;
    const int wp = (w + 3) & ~3, nrb = (w + 7) >> 3, ncb = m >> 3;
    // W_p (rows c0 .. c0 + w, contiguous; zero rows up to a multiple of 8) and Dinv^T: sL[j][i] = Dinv(i, j)
    for (int e = tid; e < 8 * nrb * m; e += kSnThreads)
    {
      const int i = e / m, c = e % m;
      sW[i * kSnLdX + c] = i < w ? a.W[(size_t)(P.c0 + i) * m + c] : 0.0;
    }
    for (int e = tid; e < wp * 8 * nrb; e += kSnThreads)
    {
      const int j = e / (8 * nrb), i = e % (8 * nrb);
      sL[j * kSnLdL + i] = (i < w && j <= i) ? a.dinv[P.dofs + i * w + j] : 0.0;
    }
    __syncthreads();
    // x_p = Dinv_p W_p: tiles (rb, cb), rb < nrb, cb < ncb, dealt out to the warps
    for (int t = warp; t < nrb * ncb; t += kSnThreads / 32)
    {
      const int rb = t / ncb, cb = t % ncb;
      double c0 = 0.0, c1 = 0.0;
      for (int ks = 0; ks < wp / 4; ++ks)
        sn_dmma(c0, c1, sL[(4 * ks + k) * kSnLdL + 8 * rb + g], sW[(4 * ks + k) * kSnLdX + 8 * cb + g]);
      const int i = 8 * rb + g, c = 8 * cb + 2 * k;
      sX[i * kSnLdX + c] = c0;
      sX[i * kSnLdX + c + 1] = c1;
      if (it.y == 0 && i < w)
      {
        a.Z[(size_t)(P.c0 + i) * m + c] = c0;
        a.Z[(size_t)(P.c0 + i) * m + c + 1] = c1;
      }
    }
    // rows below the panel handled by this CTA
    const int below0 = P.j0 + w + it.y * kSnTile, nrow = min(kSnTile, P.r - below0);
    if (nrow <= 0)
      return;
    __syncthreads(); // x_p complete, Dinv^T no longer needed
    const int nrp = (nrow + 7) & ~7;
    for (int e = tid; e < nrp * wp; e += kSnThreads)
    {
      const int aa = e % nrp, j = e / nrp; // coalesced along the rows of a column
      sL[j * kSnLdL + aa] = (aa < nrow && j < w) ? a.val[P.lofs + (size_t)(below0 + aa) + (size_t)j * P.r] : 0.0;
    }
    __syncthreads();
    // U(64 x m) = L(tile rows, panel) x_p: warp = 8-row block, all column blocks
    for (int rb = warp; rb < nrp / 8; rb += kSnThreads / 32)
    {
      double acc[DE_KERNEL_MAX_M / 8][2];
#pragma unroll
      for (int cb = 0; cb < DE_KERNEL_MAX_M / 8; ++cb)
        acc[cb][0] = acc[cb][1] = 0.0;
      for (int ks = 0; ks < wp / 4; ++ks)
      {
        const double av = sL[(4 * ks + k) * kSnLdL + 8 * rb + g];
#pragma unroll
        for (int cb = 0; cb < DE_KERNEL_MAX_M / 8; ++cb)
          if (cb < ncb)
            sn_dmma(acc[cb][0], acc[cb][1], av, sX[(4 * ks + k) * kSnLdX + 8 * cb + g]);
      }
      const int aa = 8 * rb + g;
      if (aa < nrow)
      {
        const int la = below0 + aa;
        double *dst = a.W + (size_t)a.rowidx[P.rofs + la] * m + 2 * k;
        const bool own = la < P.ns; // a later panel of the same supernode: nobody else touches this row in this level
#pragma unroll
        for (int cb = 0; cb < DE_KERNEL_MAX_M / 8; ++cb)
          if (cb < ncb)
          {
            if (own)
            {
              dst[8 * cb] -= acc[cb][0];
              dst[8 * cb + 1] -= acc[cb][1];
            }
            else
            {
              atomicAdd(dst + 8 * cb, -acc[cb][0]);
              atomicAdd(dst + 8 * cb + 1, -acc[cb][1]);
            }
          }
      }
    }
  }

  /** backward sweep, one level */
  static __global__ void __launch_bounds__(kSnThreads) sn_backward_kernel(const SnArgs a, int item0)
  {
    extern __shared__ __align__(16) unsigned char sn_dyn[];
    double *sX = reinterpret_cast<double *>(sn_dyn);        // [kSnPanel][kSnLdX]: z_p at the end
    double *sL = sX + kSnPanel * kSnLdX;                     // [kSnPanel][kSnLdL]: L(tile rows, panel)^T ; at the end: Dinv
    double *sZ = sL + kSnPanel * kSnLdL;                     // [kSnTile][kSnLdX]: Z(tile rows)
    __shared__ int last;
    const int2 it = a.items[item0 + blockIdx.x];
    const SnPanel P = a.panels[it.x];
    const int m = a.m, w = P.w, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, k = lane & 3;
    const int njb = (w + 7) >> 3, ncb = m >> 3;
    const int below0 = P.j0 + w + it.y * kSnTile, nrow = min(kSnTile, P.r - below0);
    if (nrow > 0)
    {
      const int nrp = (nrow + 3) & ~3;
      for (int e = tid; e < nrp * 8 * njb; e += kSnThreads)
      {
        const int aa = e % nrp, j = e / nrp;
        sL[j * kSnLdL + aa] = (aa < nrow && j < w) ? a.val[P.lofs + (size_t)(below0 + aa) + (size_t)j * P.r] : 0.0;
      }
      for (int e = tid; e < nrp * m; e += kSnThreads)
      {
        const int aa = e / m, c = e % m;
        sZ[aa * kSnLdX + c] = aa < nrow ? __ldcg(a.Z + (size_t)a.rowidx[P.rofs + below0 + aa] * m + c) : 0.0; // ancestors' rows: final
      }
      __syncthreads();
      // S(w x m) = L(tile rows, panel)^T Z(tile rows): A(i = g, kk = k) = L(row 4 ks + k, column 8 jb + g)
      for (int t = warp; t < njb * ncb; t += kSnThreads / 32)
      {
        const int jb = t / ncb, cb = t % ncb;
        double c0 = 0.0, c1 = 0.0;
        for (int ks = 0; ks < nrp / 4; ++ks)
          sn_dmma(c0, c1, sL[(8 * jb + g) * kSnLdL + 4 * ks + k], sZ[(4 * ks + k) * kSnLdX + 8 * cb + g]);
        const int j = 8 * jb + g;
        if (j < w)
        {
          atomicAdd(a.Z + (size_t)(P.c0 + j) * m + 8 * cb + 2 * k, -c0);
          atomicAdd(a.Z + (size_t)(P.c0 + j) * m + 8 * cb + 2 * k + 1, -c1);
        }
      }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0)
      last = (atomicAdd(a.ticket + it.x, 1) == P.ntiles - 1) ? 1 : 0;
    __syncthreads();
    if (!last)
      return;
    if (tid == 0)
      a.ticket[it.x] = 0;
    __threadfence();
    // x_p = Dinv_p^T z_p: x(i, c) = sum_{j >= i} Dinv(j, i) z(j, c); A(i = g, kk = k) = Dinv(4 ks + k, 8 ib + g)
    const int wp = (w + 3) & ~3;
    for (int e = tid; e < wp * 8 * njb; e += kSnThreads)
    {
      const int j = e / (8 * njb), i = e % (8 * njb);
      sL[j * kSnLdL + i] = (j < w && i <= j) ? a.dinv[P.dofs + j * w + i] : 0.0;
    }
    for (int e = tid; e < wp * m; e += kSnThreads)
    {
      const int j = e / m, c = e % m;
      sX[j * kSnLdX + c] = j < w ? __ldcg(a.Z + (size_t)(P.c0 + j) * m + c) : 0.0;
    }
    __syncthreads();
    for (int t = warp; t < njb * ncb; t += kSnThreads / 32)
    {
      const int ib = t / ncb, cb = t % ncb;
      double c0 = 0.0, c1 = 0.0;
      for (int ks = 0; ks < wp / 4; ++ks)
        sn_dmma(c0, c1, sL[(4 * ks + k) * kSnLdL + 8 * ib + g], sX[(4 * ks + k) * kSnLdX + 8 * cb + g]);
      const int i = 8 * ib + g;
      if (i < w)
      {
        a.Z[(size_t)(P.c0 + i) * m + 8 * cb + 2 * k] = c0;
        a.Z[(size_t)(P.c0 + i) * m + 8 * cb + 2 * k + 1] = c1;
      }
    }
  }

} // namespace de
