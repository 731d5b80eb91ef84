// Linear combination of up to three tall-skinny blocks on the FP64 tensor pipe (M = 8/16/32/64):
//     out  = S_0 C_0 + S_1 C_1 + S_2 C_2        (LOBPCG: X <- X Cx + W Cw + P Cp)
//     out2 =           S_1 C_1 + S_2 C_2        (        P <-        W Cw + P Cp)      in ONE pass over the sources,
// or, with identity0, the projection  out = S_0 + alpha (S_1 C_1 + ...)  (W <- W - X (BX^T W); reference Q_j -= Q_k S,
// kernels_cpp.hh:335-348) and the rotation X <- X C (ns = 1; kernels_cpp.hh:293-305).
//
// The first version (lincomb_kernel, kernels_lobpcg.cuh; update_kernel<M, 1>, kernels_dense.cuh) staged one source tile
// at a time behind two CTA-wide barriers and multiplied with DFMA out of shared memory: 0.37 of the HBM roofline at
// M = 32 and 0.20 at M = 64 (0.35 of the FMA pipe), 14 % of an LOBPCG solve and the largest single item at M = 64.
// This one is ts2_update_kernel's scheme (kernels_tallskinny2.cuh) with the sources of a row tile as consecutive
// pipeline stages: 4 producer warps fill a ring of (tile, source) stages with cp.async copies arriving on mbarriers,
// 12 consumer warps own whole 8-row blocks and keep ONE set of accumulators across the sources of a tile -- processed
// last source first, so that the accumulators hold out2 after source 1 and out after source 0. The product is formed
// transposed, Out^T(8 cols x 8 rows) += C_s^T(8 x 4) S_s^T(4 x 8): the coefficient fragments come from a staged copy of
// the coefficient matrices (row stride M + 4: conflict-free), the source fragments straight from the staged tile, the
// results go to global memory from registers. No CTA-wide barrier in the loop.
//
// Bounds: 8 n M (ns + nout) bytes of HBM traffic against 2 n M^2 ns flops at the measured 37 TFLOP/s of the FP64 tensor
// pipe: HBM-bound up to M = 16, about even at M = 32, tensor-pipe-bound at M = 64.
//
// Aliasing: out may be S_0 and out2 may be any S_s with s >= 1 (not S_0). A consumer warp writes rows of a tile only after
// it has read the staged copy of every source of that tile that can alias the output, and the producers only run ahead
// into OTHER row tiles.
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

#include "kernels_spmm_blocked.cuh" // mbarrier / cp.async helpers, dmma884_sp
#include "kernels_tallskinny.cuh"   // cp_async16

namespace de
{

  constexpr int kLc2ProducerWarps = 4;
  constexpr int kLc2ConsumerWarps = 12;
  constexpr int kLc2MaxSrc = 3;

  template <int M>
  struct Lc2Cfg
  {
    static constexpr int NB = M / 8;                                            // 8-column blocks
    static constexpr int KS = M / 4;                                            // k steps of one product
    static constexpr int LDT = M + 4;                                           // staged row stride (doubles)
    static constexpr int BPW = (M == 64) ? 1 : (M == 32 ? 2 : (M == 16 ? 4 : 8)); // 8-row blocks per consumer warp and tile
    static constexpr int THREADS = 32 * (kLc2ProducerWarps + kLc2ConsumerWarps);
    static constexpr int TR = 8 * kLc2ConsumerWarps * BPW;                      // rows per tile
    static constexpr int STAGES = (M == 64) ? 2 : 3;
    static constexpr size_t STAGE_BYTES = (size_t)TR * LDT * sizeof(double);
    static constexpr size_t CBYTES = (size_t)kLc2MaxSrc * M * LDT * sizeof(double); // staged coefficient matrices
    static constexpr size_t SMEM = 128 + CBYTES + STAGES * STAGE_BYTES;
  };

  struct Lc2Args
  {
    long long n;
    int ns;                      // sources (1..3)
    const double *S[kLc2MaxSrc]; // row-major n x M, leading dimension M
    const double *C[kLc2MaxSrc]; // row-major M x M coefficient matrices (device); C[0] unused with identity0
    double *out, *out2;          // out2 may be null
    int identity0;               // source 0 enters as itself: out = S_0 + alpha * (sum over s >= 1)
    double alpha;
    const int *done;             // optional: a converged driver loop, the launch is a no-op
  };

  template <int M>
  __global__ void __launch_bounds__(Lc2Cfg<M>::THREADS, 1) ts2_lincomb_kernel(const Lc2Args a)
  {
    using C = Lc2Cfg<M>;
    constexpr int NPW = kLc2ProducerWarps, NCW = kLc2ConsumerWarps, ST = C::STAGES;
    extern __shared__ __align__(128) unsigned char dynl[];
    pdl_prologue();
    if (a.done != nullptr && *a.done != 0)
      return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bar0 = smem_u32(dynl); // full[s] at +8 s, empty[s] at +8 (ST + s)
    double *Cs = reinterpret_cast<double *>(dynl + 128); // [source][M][LDT]
    double *tiles = reinterpret_cast<double *>(dynl + 128 + C::CBYTES);
    if (tid == 0)
    {
      for (int s = 0; s < ST; ++s)
      {
        mbar_init(bar0 + 8 * s, 32 * NPW);
        mbar_init(bar0 + 8 * (ST + s), NCW);
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int s = a.identity0 ? 1 : 0; s < a.ns; ++s)
    {
      const double *cm = s == 0 ? a.C[0] : (s == 1 ? a.C[1] : a.C[2]);
      for (int e = tid; e < M * M; e += C::THREADS)
        Cs[((size_t)s * M + e / M) * C::LDT + e % M] = __ldg(cm + e);
    }
    __syncthreads();

    const long long ntiles = (a.n + C::TR - 1) / C::TR;
    if (warp < NPW)
    {
      // ---------------- producers: (tile, source) -> stage, sources last to first ----------------
      constexpr int CPR = M / 2;
      const int ptid = warp * 32 + lane;
      int st = 0, use = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      {
        const long long r0 = t * C::TR;
        for (int s = a.ns - 1; s >= 0; --s)
        {
          if (use > 0)
            mbar_wait(bar0 + 8 * (ST + st), (unsigned)((use - 1) & 1));
          double *dst = tiles + (size_t)st * C::TR * C::LDT;
          const double *src = s == 0 ? a.S[0] : (s == 1 ? a.S[1] : a.S[2]); // (no dynamic index into the parameters)
#pragma unroll 4
          for (int e = ptid; e < C::TR * CPR; e += 32 * NPW)
          {
            const int r = e / CPR, c = 2 * (e % CPR);
            const bool in = r0 + r < a.n;
            const long long row = in ? r0 + r : 0;
            cp_async16(dst + r * C::LDT + c, src + (size_t)row * M + c, in);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar0 + 8 * st) : "memory");
          if (++st == ST)
          {
            st = 0;
            ++use;
          }
        }
      }
    }
    else
    {
      // ---------------- consumers ----------------
      const int cw = warp - NPW;
      const int g = lane >> 2, k = lane & 3;
      int st = 0, use = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      {
        const long long r0 = t * C::TR;
        double c[C::BPW][C::NB][2];
#pragma unroll
        for (int b = 0; b < C::BPW; ++b)
#pragma unroll
          for (int jb = 0; jb < C::NB; ++jb)
            c[b][jb][0] = c[b][jb][1] = 0.0;
        for (int s = a.ns - 1; s >= 0; --s)
        {
          mbar_wait(bar0 + 8 * st, (unsigned)(use & 1));
          const double *Xs = tiles + (size_t)st * C::TR * C::LDT;
          if (s == 0 && a.identity0)
          {
            // lane (g, k) holds Out(rows 2k, 2k + 1; column 8 jb + g) of its row blocks
#pragma unroll
            for (int b = 0; b < C::BPW; ++b)
            {
              const double *x0 = Xs + ((b * NCW + cw) * 8 + 2 * k) * C::LDT + g;
#pragma unroll
              for (int jb = 0; jb < C::NB; ++jb)
              {
                c[b][jb][0] = fma(a.alpha, c[b][jb][0], x0[8 * jb]);
                c[b][jb][1] = fma(a.alpha, c[b][jb][1], x0[C::LDT + 8 * jb]);
              }
            }
          }
          else
          {
            // A(i = g, kk = k) of (ks, jb) = C_s(4 ks + k, 8 jb + g); B(kk = k, n = g) = S_s(row 8 rb + g, column 4 ks + k)
            const double *cs = Cs + ((size_t)s * M + k) * C::LDT + g;
            const double *xr = Xs + (cw * 8 + g) * C::LDT + k;
#pragma unroll
            for (int ks = 0; ks < C::KS; ++ks)
            {
              double xv[C::BPW];
#pragma unroll
              for (int b = 0; b < C::BPW; ++b)
                xv[b] = xr[(size_t)b * NCW * 8 * C::LDT + 4 * ks];
#pragma unroll
              for (int jb = 0; jb < C::NB; ++jb)
              {
                const double f = cs[4 * ks * C::LDT + 8 * jb];
#pragma unroll
                for (int b = 0; b < C::BPW; ++b)
                  dmma884_sp(c[b][jb][0], c[b][jb][1], f, xv[b]);
              }
            }
          }
          __syncwarp();
          if (lane == 0)
            mbar_arrive(bar0 + 8 * (ST + st));
          if (++st == ST)
          {
            st = 0;
            ++use;
          }
          double *dst = (s == 0) ? a.out : ((s == 1 && a.out2 != nullptr) ? a.out2 : nullptr);
          if (dst != nullptr)
          {
#pragma unroll
            for (int b = 0; b < C::BPW; ++b)
            {
              const long long row0 = r0 + (b * NCW + cw) * 8 + 2 * k;
#pragma unroll
              for (int jb = 0; jb < C::NB; ++jb)
              {
                if (row0 < a.n)
                  dst[(size_t)row0 * M + 8 * jb + g] = c[b][jb][0];
                if (row0 + 1 < a.n)
                  dst[(size_t)(row0 + 1) * M + 8 * jb + g] = c[b][jb][1];
              }
            }
          }
        }
      }
    }
  }

} // namespace de
