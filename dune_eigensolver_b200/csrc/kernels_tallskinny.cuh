// Pipelined tall-skinny kernel for sm_100a (second generation of K3/K4/K5, used for m = 8/16/32/64):
// one persistent kernel streams row tiles of an n x M block through a 3-stage cp.async ring in shared memory and,
// per tile, optionally (a) right-multiplies it by a small M x M factor (block update X <- X R) and (b) accumulates
// a Gram matrix in registers. Doing (a) and (b) in one pass is what makes CholQR2 cost
//      8nm (Gram) + 16nm (update fused with the Gram of the result) [+ 16nm second update, skipped when the
//      fused Gram already equals I to working precision]
// instead of 48nm bytes. ncu of the first-generation kernels (profiles/r01_ncu_kernels_baseline.csv) showed them
// bound by shared-memory wavefronts (l1tex 57-72 %) at 30 % occupancy with single-buffered tiles; this version
// uses 8-wide register tiles (4x8 Gram blocks, 2x8 update blocks), 128-bit shared loads that are conflict-free
// for the padded row stride M+2, and keeps two tiles in flight per SM.
//
// Roofline: HBM-bound for M <= 32 (16nm or 8nm bytes); at M = 64 the FP64 pipe (2 n M^2 flops, halved by
// the triangular / symmetric structure) is the second limiter.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_sparse.cuh"

namespace de
{

  __device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, bool valid)
  {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int bytes = valid ? 16 : 0; // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(bytes));
  }
  __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
  template <int N>
  __device__ __forceinline__ void cp_async_wait()
  {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
  }

  template <int M, bool UPPER, int NOPS>
  struct TsCfg
  {
    static constexpr int THREADS = 256;
    static constexpr int STAGES = 3;
    static constexpr int TR = (NOPS == 2 ? 2048 : 4096) / M; // rows per tile: ~32 KB of operand data per stage, so
                                                             // two stages in flight on 148 SMs cover HBM latency
    static constexpr int LDT = M + 2;                     // padded row stride (doubles); rows stay 16-byte aligned
    static constexpr int TILE = TR * LDT;                 // doubles per staged tile
    // update phase: a warp unit is (32*RT rows) x (8 columns); 16/RT units per tile for every M
    static constexpr int RT = 2;
    static constexpr int NCG = M / 8;
    static constexpr int NRB = TR / (32 * RT);
    static constexpr int UNITS = NCG * NRB;
    // Gram phase: 4 x 8 register blocks of G
    static constexpr int NBI = M / 4, NBJ = M / 8;
    static constexpr int count_tiles()
    {
      int c = 0;
      for (int bi = 0; bi < NBI; ++bi)
        for (int bj = 0; bj < NBJ; ++bj)
          if (!UPPER || 4 * bi <= 8 * bj + 7)
            ++c;
      return c;
    }
    static constexpr int NT = count_tiles();              // threads that tile G once
    static constexpr int RG = THREADS / NT;               // row groups
  };

  struct TsArgs
  {
    long long n;
    const double *X;      // input block (n x M view)
    int ldx;
    const double *Y;      // second Gram operand when !SAME (n x M view)
    int ldy;
    const double *R;      // M x M row-major factor (DO_UPDATE)
    double *Out;          // updated block (DO_UPDATE); may alias X
    int ldo;
    int upper;            // R is upper triangular: column group c only needs k < 8c+8
    const int *skip_flag; // optional: if *skip_flag != 0 the kernel returns immediately (second CholQR sweep)
    double *partials;     // DO_GRAM: [gridDim.x][M*M]
  };

  /** see file comment. Template switches:
   *   DO_UPDATE  Out = X R (tile staged in shared memory first, so Out may alias X)
   *   DO_GRAM    accumulate G = A^T B over all rows, A = (DO_UPDATE ? updated X : X), B = (SAME ? A : Y)
   *   UPPER      G is symmetric: only the blocks meeting the upper triangle are computed, mirrored on output
   *   SAME       single Gram operand (B aliases A) */
  template <int M, bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  __global__ void __launch_bounds__(256, 1) tall_skinny_kernel(const TsArgs a)
  {
    constexpr int NOPS = (DO_GRAM && !SAME) ? 2 : 1;
    using C = TsCfg<M, UPPER, NOPS>;
    static_assert(!(DO_UPDATE && DO_GRAM) || SAME, "fused update+Gram works on one operand");
    extern __shared__ __align__(16) double smem[];
    double *tiles = smem;                                            // STAGES x NOPS x TILE
    double *Rs = smem + (size_t)C::STAGES * NOPS * C::TILE;          // M x M (DO_UPDATE)

    if (a.skip_flag != nullptr && *a.skip_flag != 0)
      return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (DO_UPDATE)
      for (int e = tid; e < M * M; e += C::THREADS)
        Rs[e] = __ldg(a.R + e);

    // Gram block of this thread
    int bi = 0, bj = 0, grp = 0;
    bool gram_thread = false;
    if (DO_GRAM)
    {
      grp = tid / C::NT;
      gram_thread = grp < C::RG;
      int want = tid % C::NT, seen = 0;
      for (int i = 0; i < C::NBI; ++i)
        for (int j = 0; j < C::NBJ; ++j)
          if (!UPPER || 4 * i <= 8 * j + 7)
          {
            if (seen == want)
            {
              bi = i;
              bj = j;
            }
            ++seen;
          }
    }
    double acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 8; ++q)
        acc[p][q] = 0.0;

    const long long ntiles = (a.n + C::TR - 1) / C::TR;
    constexpr int HP = M / 2;

    auto issue_tile = [&](long long tile, int stage)
    {
      if (tile < ntiles)
      {
        const long long r0 = tile * C::TR;
        double *dst = tiles + (size_t)stage * NOPS * C::TILE;
        for (int e = tid; e < C::TR * HP; e += C::THREADS)
        {
          const int r = e / HP, c = 2 * (e % HP);
          const bool in = r0 + r < a.n;
          const long long row = in ? r0 + r : 0;
          cp_async16(dst + r * C::LDT + c, a.X + (size_t)row * a.ldx + c, in);
          if (NOPS == 2)
            cp_async16(dst + C::TILE + r * C::LDT + c, a.Y + (size_t)row * a.ldy + c, in);
        }
      }
      cp_async_commit(); // always commit (possibly empty) so that the group accounting stays uniform
    };

    // prologue: two tiles in flight
    long long tile = blockIdx.x;
    issue_tile(tile, 0);
    issue_tile(tile + gridDim.x, 1);
    int stage = 0;
    for (; tile < ntiles; tile += gridDim.x)
    {
      // the buffer two stages ahead was released by the barrier that ended the previous iteration
      issue_tile(tile + 2 * (long long)gridDim.x, (stage + 2) % C::STAGES);
      cp_async_wait<2>();
      __syncthreads();
      double *Xs = tiles + (size_t)stage * NOPS * C::TILE;
      const long long r0 = tile * C::TR;

      if (DO_UPDATE)
      {
        // ---- Y = X R on the staged tile. Each warp owns one (32*RT rows) x (8 columns) patch; lane = row, so the
        //      R(k, 8 cols) operand is a warp-wide broadcast and X(row, k..k+1) a conflict-free 128-bit load.
        static_assert(C::UNITS <= 8, "one update unit per warp");
        const bool has_unit = warp < C::UNITS;
        const int cg = warp % C::NCG, rb = warp / C::NCG;
        const int c0 = cg * 8;
        double out[C::RT][8];
        if (has_unit)
        {
          const int kmax = a.upper ? c0 + 8 : M;
#pragma unroll
          for (int q = 0; q < C::RT; ++q)
#pragma unroll
            for (int c = 0; c < 8; ++c)
              out[q][c] = 0.0;
          const double *xrow = Xs + (rb * 32 * C::RT + lane) * C::LDT;
#pragma unroll 2
          for (int k = 0; k < kmax; k += 2)
          {
            double2 xv[C::RT];
#pragma unroll
            for (int q = 0; q < C::RT; ++q)
              xv[q] = ld2(xrow + q * 32 * C::LDT + k);
            double r0v[8], r1v[8];
#pragma unroll
            for (int c = 0; c < 8; c += 2)
            {
              const double2 u = ld2(Rs + k * M + c0 + c), w = ld2(Rs + (k + 1) * M + c0 + c);
              r0v[c] = u.x;
              r0v[c + 1] = u.y;
              r1v[c] = w.x;
              r1v[c + 1] = w.y;
            }
#pragma unroll
            for (int q = 0; q < C::RT; ++q)
#pragma unroll
              for (int c = 0; c < 8; ++c)
              {
                out[q][c] = fma(xv[q].x, r0v[c], out[q][c]);
                out[q][c] = fma(xv[q].y, r1v[c], out[q][c]);
              }
          }
        }
        __syncthreads(); // every warp has finished READING the tile: the result may now overwrite it in place
        if (has_unit)
        {
#pragma unroll
          for (int q = 0; q < C::RT; ++q)
          {
            double *y = Xs + (rb * 32 * C::RT + q * 32 + lane) * C::LDT + c0;
#pragma unroll
            for (int c = 0; c < 8; c += 2)
              st2(y + c, make_double2(out[q][c], out[q][c + 1]));
          }
        }
        __syncthreads();
        // ---- coalesced write-out of the updated tile ----
        for (int e = tid; e < C::TR * HP; e += C::THREADS)
        {
          const int r = e / HP, c = 2 * (e % HP);
          if (r0 + r < a.n)
            st2(a.Out + (size_t)(r0 + r) * a.ldo + c, ld2(Xs + r * C::LDT + c));
        }
      }

      if (DO_GRAM)
      {
        const double *Ys = SAME ? Xs : Xs + C::TILE;
        if (gram_thread)
        {
#pragma unroll 2
          for (int r = grp; r < C::TR; r += C::RG)
          {
            const double2 x01 = ld2(Xs + r * C::LDT + 4 * bi), x23 = ld2(Xs + r * C::LDT + 4 * bi + 2);
            const double xv[4] = {x01.x, x01.y, x23.x, x23.y};
            double yv[8];
#pragma unroll
            for (int q = 0; q < 8; q += 2)
            {
              const double2 t = ld2(Ys + r * C::LDT + 8 * bj + q);
              yv[q] = t.x;
              yv[q + 1] = t.y;
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
              for (int q = 0; q < 8; ++q)
                acc[p][q] = fma(xv[p], yv[q], acc[p][q]);
          }
        }
      }
      __syncthreads(); // all reads of this stage are done: it may be refilled by the next iteration's prefetch
      stage = (stage + 1) % C::STAGES;
    }
    cp_async_wait<0>();

    if (DO_GRAM)
    {
      // combine the row groups (fixed order) through shared memory, 8 values per thread at a time
      __syncthreads();
      double *red = smem; // THREADS x 8 doubles
      double *outp = a.partials + (size_t)blockIdx.x * M * M;
      for (int p = 0; p < 4; ++p)
      {
#pragma unroll
        for (int q = 0; q < 8; ++q)
          red[q * C::THREADS + tid] = acc[p][q];
        __syncthreads();
        if (tid < C::NT)
        {
#pragma unroll
          for (int q = 0; q < 8; ++q)
          {
            double s = 0.0;
            for (int g = 0; g < C::RG; ++g)
              s += red[q * C::THREADS + g * C::NT + tid];
            const int gi = 4 * bi + p, gj = 8 * bj + q;
            if (!UPPER)
              outp[gi * M + gj] = s;
            else if (gi <= gj)
            {
              outp[gi * M + gj] = s;
              outp[gj * M + gi] = s;
            }
          }
        }
        __syncthreads();
      }
    }
  }

  template <int M, bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  constexpr size_t tall_skinny_smem_bytes()
  {
    constexpr int NOPS = (DO_GRAM && !SAME) ? 2 : 1;
    using C = TsCfg<M, UPPER, NOPS>;
    size_t d = (size_t)C::STAGES * NOPS * C::TILE + (DO_UPDATE ? (size_t)M * M : 0);
    const size_t red = (size_t)C::THREADS * 8;
    if (d < red)
      d = red;
    return d * sizeof(double);
  }

} // namespace de
