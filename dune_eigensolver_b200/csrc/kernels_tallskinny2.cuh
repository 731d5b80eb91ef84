// Second-generation block update (+ Gram) kernel:  Out = X R  and, in the same pass,  G = Out^T Out.
// This is the second sweep of CholQR2 (X <- X R1^-1 fused with the Gram matrix of the result) -- after the SpMM the
// most expensive kernel of a StandardLargest iteration.
//
// ncu of the first version (tall_skinny_kernel<32,1,1,1,1>, profiles/r01_ncu_kernels_brb.csv): 136 us for 456 MB
// (0.52 of HBM peak), 12.5 % of the warps active, 170 registers. Per 128-row tile it ran: update from shared memory ->
// barrier -> result back into shared memory -> barrier -> coalesced write-out -> Gram from shared memory -> barrier:
// five CTA-wide barriers and a shared-memory round trip of the whole tile, with 8 warps per SM.
//
// Here the product is formed TRANSPOSED on the FP64 tensor pipe,  Out^T(8 cols x 8 rows) += R^T(8 x 4) X^T(4 x 8):
//   A operand = R^T fragment (registers for the whole kernel; a triangular factor needs 20 of 32 for M = 32. Reading them
//               from shared memory instead frees 40 registers and allows 12 consumer warps, but measured no faster),
//   B operand = X(8 rows x 4 cols) fragment straight from the staged tile (stride M + 4 doubles: conflict-free),
//   C         = lane (g, k) holds Out(rows 2k, 2k+1; column 8 jb + g).
// C goes to global memory from registers (a warp store covers 4 rows x 64 contiguous bytes), and -- because rows 2kk,
// 2kk+1 of a column sit in lane kk of the same quad -- one shuffle inside the quad turns C into the operand fragment
// f = Out(row 4s + k, column 8 jb + g) of the Gram product G(jb, jb') += f(jb)^T f(jb'). Nothing goes back to shared
// memory and nothing in the loop is a CTA-wide barrier: producer warps fill a ring of tile buffers with cp.async
// copies that arrive on mbarriers, consumer warps own whole 8-row blocks (same scheme as spmm_brb_kernel).
//
// Bounds for M = 32: 16 n M bytes of HBM traffic (70 us on the 100^3 block) and n M^2 (update, triangular)
// + n M^2 (Gram, symmetric) tensor flops (55 us at the measured 37 TFLOP/s).
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

#include "kernels_spmm_blocked.cuh" // mbarrier / cp.async helpers, dmma884_sp
#include "kernels_tallskinny.cuh"   // TsArgs

namespace de
{

  constexpr int kTs2ProducerWarps = 4;
  constexpr int kTs2Stages = 3;

  template <int M>
  struct Ts2Cfg
  {
    static constexpr int NB = M / 8;                           // 8-column blocks
    static constexpr int KS = M / 4;                           // k steps of the update
    static constexpr int LDT = M + 4;                          // staged row stride (doubles)
    static constexpr bool RSMEM = (M == 64);                    // factor fragments from shared memory instead of registers
    static constexpr int NCW = (M == 32) ? 8 : 12;              // consumer warps (M = 32: up to 167 registers per thread; M = 64
                                                                // has no fused Gram and its factor in shared memory: 12 fit)
    static constexpr int THREADS = 32 * (kTs2ProducerWarps + NCW);
    static constexpr int TR = 8 * NCW * (M == 64 ? 1 : (M == 32 ? 2 : (M == 16 ? 4 : 8))); // rows per tile
    static constexpr int NBLK = TR / 8;                        // 8-row blocks per tile
    static constexpr int NT = NB * (NB + 1) / 2;               // Gram tiles jb <= jb'
    static constexpr size_t STAGE_BYTES = (size_t)TR * LDT * sizeof(double);
    static constexpr size_t RBYTES = RSMEM ? (size_t)M * LDT * sizeof(double) : 0; // staged factor, row stride LDT
    static constexpr size_t SMEM = 128 + RBYTES + kTs2Stages * STAGE_BYTES;
  };

  /** a.X (n x M, ld a.ldx) -> a.Out (may alias X), a.R row-major M x M (a.upper: upper triangular);
   *  DO_GRAM: per-CTA partial of Out^T Out at a.partials[cta * M * M + i * M + j] (full symmetric matrix).
   *  M = 64: the 128 factor fragments do not fit the register file next to the accumulators; the factor is staged in
   *  shared memory (row stride M + 4: the fragment load of a half warp is conflict-free) and DO_GRAM is not offered. */
  template <int M, bool DO_GRAM, bool PUSH = false>
  __global__ void __launch_bounds__(Ts2Cfg<M>::THREADS, 1) ts2_update_kernel(const TsArgs a)
  {
    using C = Ts2Cfg<M>;
    static_assert(!(M == 64 && DO_GRAM), "no fused Gram at M = 64");
    constexpr int NPW = kTs2ProducerWarps, NCW = C::NCW;
    extern __shared__ __align__(128) unsigned char dyn2[];
    pdl_prologue();
    if (a.done != nullptr && *a.done != 0)
      return;
    if (a.skip_flag != nullptr && *a.skip_flag != 0)
    {
      // nothing changes: the rows the previous update launch stored into the neighbours' windows are final (its stores
      // are complete: the launch has ended), so the flags can go up at once
      if (PUSH && a.push.release && blockIdx.x == 0 && (int)threadIdx.x < a.push.n)
      {
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(a.push.flag[threadIdx.x]), "l"(a.push.epoch) : "memory");
      }
      return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bar0 = smem_u32(dyn2); // full[s] at +8 s, empty[s] at +8 (kTs2Stages + s)
    double *Rs = reinterpret_cast<double *>(dyn2 + 128);                  // RSMEM: M x LDT
    double *tiles = reinterpret_cast<double *>(dyn2 + 128 + C::RBYTES);
    if (tid == 0)
    {
      for (int s = 0; s < kTs2Stages; ++s)
      {
        mbar_init(bar0 + 8 * s, 32 * NPW);
        mbar_init(bar0 + 8 * (kTs2Stages + s), NCW);
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (C::RSMEM)
      for (int e = tid; e < M * M; e += C::THREADS)
        Rs[(e / M) * C::LDT + e % M] = __ldg(a.R + e);
    __syncthreads();

    const long long ntiles = (a.n + C::TR - 1) / C::TR;
    const long long trot = PUSH ? a.push.first_row / C::TR : 0; // sweep starts here (PushRanges::first_row)
    double gacc[DO_GRAM ? C::NT : 1][2];
#pragma unroll
    for (int i = 0; i < (DO_GRAM ? C::NT : 1); ++i)
      gacc[i][0] = gacc[i][1] = 0.0;

    if (warp < NPW)
    {
      // ---------------- producers: rows of the tile -> staged rows (zero fill past the end of the block) ----------
      constexpr int CPR = M / 2;
      const int ptid = warp * 32 + lane;
      int s = 0, use = 0;
      for (long long tq = blockIdx.x; tq < ntiles; tq += gridDim.x)
      {
        const long long t = (PUSH && tq + trot >= ntiles) ? tq + trot - ntiles : tq + trot;
        if (use > 0)
          mbar_wait(bar0 + 8 * (kTs2Stages + s), (unsigned)((use - 1) & 1));
        double *dst = tiles + (size_t)s * C::TR * C::LDT;
        const long long r0 = t * C::TR;
#pragma unroll 4
        for (int e = ptid; e < C::TR * CPR; e += 32 * NPW)
        {
          const int r = e / CPR, c = 2 * (e % CPR);
          const bool in = r0 + r < a.n;
          const long long row = in ? r0 + r : 0;
          cp_async16(dst + r * C::LDT + c, a.X + (size_t)row * a.ldx + c, in);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar0 + 8 * s) : "memory");
        if (++s == kTs2Stages)
        {
          s = 0;
          ++use;
        }
      }
    }
    else
    {
      // ---------------- consumers ----------------
      const int cw = warp - NPW;
      const int g = lane >> 2, k = lane & 3;
      // R^T fragments: A(i = g, kk = k) of (ks, jb) = R(4 ks + k, 8 jb + g)
      double rfrag[C::RSMEM ? 1 : C::KS][C::RSMEM ? 1 : C::NB];
      if (!C::RSMEM)
      {
#pragma unroll
        for (int ks = 0; ks < (C::RSMEM ? 1 : C::KS); ++ks)
#pragma unroll
          for (int jb = 0; jb < (C::RSMEM ? 1 : C::NB); ++jb)
            rfrag[ks][jb] = (!a.upper || ks <= 2 * jb + 1) ? __ldg(a.R + (4 * ks + k) * M + 8 * jb + g) : 0.0;
      }
      const double *rs = Rs + k * C::LDT + g; // RSMEM: fragment (ks, jb) = rs[4 ks LDT + 8 jb]

      int s = 0, use = 0;
      for (long long tq = blockIdx.x; tq < ntiles; tq += gridDim.x)
      {
        const long long t = (PUSH && tq + trot >= ntiles) ? tq + trot - ntiles : tq + trot;
        mbar_wait(bar0 + 8 * s, (unsigned)(use & 1));
        const double *Xs = tiles + (size_t)s * C::TR * C::LDT;
        const long long r0 = t * C::TR;
        for (int rb = cw; rb < C::NBLK; rb += NCW)
        {
          // B(kk = k, n = g) of step ks = X(row 8 rb + g, column 4 ks + k)
          const double *xr = Xs + (rb * 8 + g) * C::LDT + k;
          double c[C::NB][2];
#pragma unroll
          for (int jb = 0; jb < C::NB; ++jb)
            c[jb][0] = c[jb][1] = 0.0;
#pragma unroll
          for (int ks = 0; ks < C::KS; ++ks)
          {
            const double xv = xr[4 * ks];
#pragma unroll
            for (int jb = 0; jb < C::NB; ++jb)
              if (!a.upper || ks <= 2 * jb + 1) // triangular factor: column block jb only sees k < 8 jb + 8 (uniform)
                dmma884_sp(c[jb][0], c[jb][1], C::RSMEM ? rs[4 * ks * C::LDT + 8 * jb] : rfrag[C::RSMEM ? 0 : ks][C::RSMEM ? 0 : jb], xv);
          }
          // c[jb] = Out(rows 8 rb + 2k, + 2k + 1; column 8 jb + g)
          const long long row0 = r0 + rb * 8 + 2 * k;
#pragma unroll
          for (int jb = 0; jb < C::NB; ++jb)
          {
            if (row0 < a.n)
              a.Out[(size_t)row0 * a.ldo + 8 * jb + g] = c[jb][0];
            if (row0 + 1 < a.n)
              a.Out[(size_t)(row0 + 1) * a.ldo + 8 * jb + g] = c[jb][1];
          }
          // halo rows of the neighbours: stored into their windows from the same registers (NVLink peer stores)
          // (a template switch: with the test in the loop the single-GPU kernel was 18 % slower, 101 -> 120 us at m = 32)
          for (int p = 0; PUSH && p < a.push.n; ++p)
          {
            if (row0 + 1 < a.push.lo[p] || row0 >= a.push.hi[p] || row0 >= a.n)
              continue;
            double *d0 = a.push.dst[p] + (size_t)(row0 - a.push.lo[p]) * M + g;
#pragma unroll
            for (int jb = 0; jb < C::NB; ++jb)
            {
              if (row0 >= a.push.lo[p])
                d0[8 * jb] = c[jb][0];
              if (row0 + 1 < a.push.hi[p] && row0 + 1 < a.n)
                d0[M + 8 * jb] = c[jb][1];
            }
          }
          if (DO_GRAM)
          {
            // rows past the end of the block were staged as zeros: they add nothing
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
            {
              const int src = (lane & ~3) | (2 * sl + (k >> 1));
              double f[C::NB];
#pragma unroll
              for (int jb = 0; jb < C::NB; ++jb)
              {
                const double v0 = __shfl_sync(0xffffffffu, c[jb][0], src);
                const double v1 = __shfl_sync(0xffffffffu, c[jb][1], src);
                f[jb] = (k & 1) ? v1 : v0; // Out(row 4 sl + k, column 8 jb + g)
              }
              int ti = 0;
#pragma unroll
              for (int bi = 0; bi < C::NB; ++bi)
#pragma unroll
                for (int bj = bi; bj < C::NB; ++bj)
                {
                  dmma884_sp(gacc[ti][0], gacc[ti][1], f[bi], f[bj]); // D(i = g, j = 2k, 2k+1) of tile (bi, bj)
                  ++ti;
                }
            }
          }
        }
        __syncwarp();
        if (lane == 0)
          mbar_arrive(bar0 + 8 * (kTs2Stages + s));
        if (++s == kTs2Stages)
        {
          s = 0;
          ++use;
        }
      }
    }

    if (PUSH && a.push.release)
    {
      // the CTA that finishes last releases the halo flags: every CTA's peer stores are fenced before its ticket
      __threadfence_system();
      __syncthreads();
      __shared__ int last_cta;
      if (tid == 0)
        last_cta = (atomicAdd(a.push.ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
      __syncthreads();
      if (last_cta)
      {
        __threadfence_system();
        if (tid < a.push.n)
          asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(a.push.flag[tid]), "l"(a.push.epoch) : "memory");
        if (tid == 0)
          *a.push.ticket = 0;
      }
    }
    if (DO_GRAM)
    {
      // fold the consumer warps in fixed order into one M x M matrix, mirror the strict lower block triangle
      __syncthreads();
      double *G = tiles;
      const int g = lane >> 2, k = lane & 3;
      for (int turn = 0; turn < NCW; ++turn)
      {
        if (warp - NPW == turn)
        {
          int ti = 0;
#pragma unroll
          for (int bi = 0; bi < C::NB; ++bi)
#pragma unroll
            for (int bj = bi; bj < C::NB; ++bj)
            {
#pragma unroll
              for (int e = 0; e < 2; ++e)
              {
                const int gi = 8 * bi + g, gj = 8 * bj + 2 * k + e;
                G[gi * M + gj] = gacc[ti][e] + (turn == 0 ? 0.0 : G[gi * M + gj]);
              }
              ++ti;
            }
        }
        __syncthreads();
      }
      double *outp = a.partials + (size_t)blockIdx.x * M * M;
      for (int e = tid; e < M * M; e += C::THREADS)
      {
        const int i = e / M, j = e % M;
        outp[e] = ((i >> 3) <= (j >> 3)) ? G[e] : G[j * M + i];
      }
    }
  }

} // namespace de

namespace de
{
  // ------------------------------------------------------------------------------------------------------------
  // Gram matrix of one block, G = X^T X (upper block triangle computed, mirrored on output), M = 8/16/32/64.
  // Same ring of staged tiles and the same producer warps as ts2_update_kernel. A consumer warp takes 4-row slabs; the
  // operand fragment of column block jb, f(jb) = X(row 4 s + k, column 8 jb + g), is ONE conflict-free 64-bit shared load
  // (it serves as A and as B operand). At M = 64 the 36 tiles of the upper triangle (72 accumulator registers each way)
  // are split over two tile groups: warps 2w and 2w+1 work on the same slabs with half of the tiles each.
  // The first version ran 8 warps per SM with a CTA barrier per tile: 0.15 of HBM peak at M = 64 (DMMA latency-bound).
  // ------------------------------------------------------------------------------------------------------------
  constexpr int kTg2ConsumerWarps = 12;
  constexpr int kTg2Threads = 32 * (kTs2ProducerWarps + kTg2ConsumerWarps);

  template <int M>
  struct Tg2Cfg
  {
    static constexpr int NB = M / 8;
    static constexpr int LDT = M + 4;
    static constexpr int TG = (M == 64) ? 2 : 1;                       // tile groups
    static constexpr int SW = kTg2ConsumerWarps / TG;                  // slab workers per tile group
    static constexpr int TR = (M == 64) ? 96 : (M == 32 ? 192 : (M == 16 ? 384 : 768)); // rows per tile (multiple of 4 SW)
    static constexpr int NT = NB * (NB + 1) / 2;
    static constexpr int NTW = (NT + TG - 1) / TG;                     // tiles per warp
    static constexpr size_t STAGE_BYTES = (size_t)TR * LDT * sizeof(double);
    static constexpr size_t SMEM = 128 + kTs2Stages * STAGE_BYTES;
  };

  template <int M>
  __global__ void __launch_bounds__(kTg2Threads, 1) ts2_gram_kernel(const TsArgs a)
  {
    using C = Tg2Cfg<M>;
    constexpr int NPW = kTs2ProducerWarps, NCW = kTg2ConsumerWarps;
    extern __shared__ __align__(128) unsigned char dyn3[];
    pdl_prologue();
    if (a.skip_flag != nullptr && *a.skip_flag != 0)
      return; // (one-sweep CholQR: the second sweep's Gram matrix is not needed)
    if (a.done != nullptr && *a.done != 0)
      return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bar0 = smem_u32(dyn3);
    double *tiles = reinterpret_cast<double *>(dyn3 + 128);
    if (tid == 0)
    {
      for (int s = 0; s < kTs2Stages; ++s)
      {
        mbar_init(bar0 + 8 * s, 32 * NPW);
        mbar_init(bar0 + 8 * (kTs2Stages + s), NCW);
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (a.n + C::TR - 1) / C::TR;
    double gacc[C::NTW][2];
#pragma unroll
    for (int i = 0; i < C::NTW; ++i)
      gacc[i][0] = gacc[i][1] = 0.0;
    const int cw = warp - NPW;
    const int tgi = cw % C::TG, swi = cw / C::TG;

    if (warp < NPW)
    {
      constexpr int CPR = M / 2;
      const int ptid = warp * 32 + lane;
      int s = 0, use = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      {
        if (use > 0)
          mbar_wait(bar0 + 8 * (kTs2Stages + s), (unsigned)((use - 1) & 1));
        double *dst = tiles + (size_t)s * C::TR * C::LDT;
        const long long r0 = t * C::TR;
#pragma unroll 4
        for (int e = ptid; e < C::TR * CPR; e += 32 * NPW)
        {
          const int r = e / CPR, c = 2 * (e % CPR);
          const bool in = r0 + r < a.n;
          const long long row = in ? r0 + r : 0;
          cp_async16(dst + r * C::LDT + c, a.X + (size_t)row * a.ldx + c, in);
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar0 + 8 * s) : "memory");
        if (++s == kTs2Stages)
        {
          s = 0;
          ++use;
        }
      }
    }
    else
    {
      const int g = lane >> 2, k = lane & 3;
      int s = 0, use = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      {
        mbar_wait(bar0 + 8 * s, (unsigned)(use & 1));
        const double *Xs = tiles + (size_t)s * C::TR * C::LDT;
#pragma unroll 2
        for (int slab = swi; slab < C::TR / 4; slab += C::SW)
        {
          const double *xs = Xs + (slab * 4 + k) * C::LDT + g;
          double f[C::NB];
#pragma unroll
          for (int jb = 0; jb < C::NB; ++jb)
            f[jb] = xs[8 * jb];
          int idx = 0;
#pragma unroll
          for (int bi = 0; bi < C::NB; ++bi)
#pragma unroll
            for (int bj = bi; bj < C::NB; ++bj)
            {
              if (idx % C::TG == tgi) // warp-uniform; idx is a compile-time constant after unrolling
                dmma884_sp(gacc[idx / C::TG][0], gacc[idx / C::TG][1], f[bi], f[bj]);
              ++idx;
            }
        }
        __syncwarp();
        if (lane == 0)
          mbar_arrive(bar0 + 8 * (kTs2Stages + s));
        if (++s == kTs2Stages)
        {
          s = 0;
          ++use;
        }
      }
    }

    // fold the slab workers in fixed order, mirror, emit the CTA partial
    __syncthreads();
    double *G = tiles;
    {
      const int g = lane >> 2, k = lane & 3;
      for (int turn = 0; turn < C::SW; ++turn)
      {
        if (warp >= NPW && swi == turn)
        {
          int idx = 0;
#pragma unroll
          for (int bi = 0; bi < C::NB; ++bi)
#pragma unroll
            for (int bj = bi; bj < C::NB; ++bj)
            {
              if (idx % C::TG == tgi)
              {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                {
                  const int gi = 8 * bi + g, gj = 8 * bj + 2 * k + e;
                  G[gi * M + gj] = gacc[idx / C::TG][e] + (turn == 0 ? 0.0 : G[gi * M + gj]);
                }
              }
              ++idx;
            }
        }
        __syncthreads();
      }
    }
    double *outp = a.partials + (size_t)blockIdx.x * M * M;
    for (int e = tid; e < M * M; e += kTg2Threads)
    {
      const int i = e / M, j = e % M;
      outp[e] = ((i >> 3) <= (j >> 3)) ? G[e] : G[j * M + i];
    }
  }

} // namespace de
