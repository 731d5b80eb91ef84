// Two-operand tall-skinny Gram matrix, G = X^T Y (replaces dot_products_all_blocked, reference kernels_cpp.hh:58-96, and
// the X^T (B X) Grams of B_orthonormalize_blocked, :451-460, :553-565) for wide blocks.
//
// At M = 64 this product is not HBM-bound: 2 n M^2 flops against 16 n M bytes = 8 flop/byte, above the FP64 ridge of a
// B200 (37 TFLOP/s DMMA / 6.5 TB/s = 5.7). The first-generation kernel (tall_skinny_kernel<64,0,1,*,0>: CTA barriers
// between the phases, 8 warps) ran at 0.41 of the tensor roof = 0.29 of HBM peak (profiles/r02_c5_*).
//
// Same scheme as ts2_gram_kernel (kernels_tallskinny2.cuh): 4 producer warps fill a ring of staged tiles with cp.async
// copies arriving on mbarriers -- here a tile holds the same rows of X and of Y -- and 12 consumer warps take 4-row
// slabs. The operand fragments are single conflict-free 64-bit shared loads: fx(bi) = X(row 4 s + k, column 8 bi + g) is
// the A operand (X^T), fy(bj) = Y(row 4 s + k, column 8 bj + g) the B operand of tile (bi, bj) of G. The NB x NB tiles
// are split over TG tile groups of warps as quadrants (M = 64: four groups of 4 x 4 tiles, 8 loads per 16 DMMA, 32
// accumulator registers); the SW warps of a group share the slabs of a tile. Per-CTA partials, reduced in fixed order.
#pragma once

#include "kernels_tallskinny2.cuh"

namespace de
{

  template <int M>
  struct Tg3Cfg
  {
    static constexpr int NB = M / 8;
    static constexpr int LDT = M + 4;
    static constexpr int QB = (M == 64) ? 4 : NB;                       // column blocks per quadrant side
    static constexpr int QS = NB / QB;                                  // quadrants per side
    static constexpr int TG = QS * QS;                                  // tile groups
    static constexpr int SW = kTg2ConsumerWarps / TG;                   // slab workers per tile group
    static constexpr int TR = (M == 64) ? 60 : (M == 32 ? 96 : (M == 16 ? 192 : 288)); // rows per tile (multiple of 4 SW)
    static constexpr size_t OP_BYTES = (size_t)TR * LDT * sizeof(double); // one operand of a stage
    static constexpr size_t STAGE_BYTES = 2 * OP_BYTES;
    static constexpr size_t SMEM = 128 + kTs2Stages * STAGE_BYTES;
    static_assert(TR % (4 * SW) == 0, "slabs must divide evenly over the slab workers");
    static_assert((size_t)M * M * sizeof(double) <= kTs2Stages * STAGE_BYTES, "the fold buffer reuses the tile ring");
  };

  /** a.X, a.Y: n x M views (ld a.ldx / a.ldy); a.partials[cta * M * M + i * M + j] = the CTA's part of (X^T Y)(i, j) */
  template <int M>
  __global__ void __launch_bounds__(kTg2Threads, 1) ts2_gram2_kernel(const TsArgs a)
  {
    using C = Tg3Cfg<M>;
    constexpr int NPW = kTs2ProducerWarps, NCW = kTg2ConsumerWarps;
    extern __shared__ __align__(128) unsigned char dyn4[];
    pdl_prologue();
    if (a.done != nullptr && *a.done != 0)
      return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bar0 = smem_u32(dyn4);
    double *tiles = reinterpret_cast<double *>(dyn4 + 128);
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
    double gacc[C::QB * C::QB][2];
#pragma unroll
    for (int i = 0; i < C::QB * C::QB; ++i)
      gacc[i][0] = gacc[i][1] = 0.0;
    const int cw = warp - NPW;
    const int tgi = cw % C::TG, swi = cw / C::TG;
    const int qi = tgi / C::QS, qj = tgi % C::QS;

    if (warp < NPW)
    {
      constexpr int CPR = M / 2; // 16-byte chunks per row
      const int ptid = warp * 32 + lane;
      int s = 0, use = 0;
      for (long long t = blockIdx.x; t < ntiles; t += gridDim.x)
      {
        if (use > 0)
          mbar_wait(bar0 + 8 * (kTs2Stages + s), (unsigned)((use - 1) & 1));
        double *dx = tiles + (size_t)s * 2 * C::TR * C::LDT;
        double *dy = dx + (size_t)C::TR * C::LDT;
        const long long r0 = t * C::TR;
#pragma unroll 4
        for (int e = ptid; e < C::TR * CPR; e += 32 * NPW)
        {
          const int r = e / CPR, c = 2 * (e % CPR);
          const bool in = r0 + r < a.n; // rows past the end are staged as zeros: they add nothing
          const long long row = in ? r0 + r : 0;
          cp_async16(dx + r * C::LDT + c, a.X + (size_t)row * a.ldx + c, in);
          cp_async16(dy + r * C::LDT + c, a.Y + (size_t)row * a.ldy + c, in);
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
        const double *Xs = tiles + (size_t)s * 2 * C::TR * C::LDT;
        const double *Ys = Xs + (size_t)C::TR * C::LDT;
#pragma unroll 2
        for (int slab = swi; slab < C::TR / 4; slab += C::SW)
        {
          const int off = (slab * 4 + k) * C::LDT + g;
          double fx[C::QB], fy[C::QB];
#pragma unroll
          for (int b = 0; b < C::QB; ++b)
          {
            fx[b] = Xs[off + 8 * (qi * C::QB + b)];
            fy[b] = Ys[off + 8 * (qj * C::QB + b)];
          }
#pragma unroll
          for (int bi = 0; bi < C::QB; ++bi)
#pragma unroll
            for (int bj = 0; bj < C::QB; ++bj)
              dmma884_sp(gacc[bi * C::QB + bj][0], gacc[bi * C::QB + bj][1], fx[bi], fy[bj]); // D(i = g; j = 2k, 2k + 1)
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

    // fold the slab workers of every tile group in fixed order (deterministic), emit the CTA partial
    __syncthreads();
    double *G = tiles;
    {
      const int g = lane >> 2, k = lane & 3;
      for (int turn = 0; turn < C::SW; ++turn)
      {
        if (warp >= NPW && swi == turn)
        {
#pragma unroll
          for (int bi = 0; bi < C::QB; ++bi)
#pragma unroll
            for (int bj = 0; bj < C::QB; ++bj)
#pragma unroll
              for (int e = 0; e < 2; ++e)
              {
                const int gi = 8 * (qi * C::QB + bi) + g, gj = 8 * (qj * C::QB + bj) + 2 * k + e;
                G[gi * M + gj] = gacc[bi * C::QB + bj][e] + (turn == 0 ? 0.0 : G[gi * M + gj]);
              }
        }
        __syncthreads();
      }
    }
    double *outp = a.partials + (size_t)blockIdx.x * M * M;
    for (int e = tid; e < M * M; e += kTg2Threads)
      outp[e] = G[e];
  }

} // namespace de
