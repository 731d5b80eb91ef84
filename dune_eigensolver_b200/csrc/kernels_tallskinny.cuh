// Pipelined tall-skinny kernel for sm_100a (second generation of K3/K4/K5, used for m = 8/16/32/64):
// one persistent kernel streams row tiles of an n x M block through a 3-stage cp.async ring in shared memory and,
// per tile, optionally (a) right-multiplies it by a small M x M factor (block update X <- X R) and (b) accumulates
// a Gram matrix in registers. Doing (a) and (b) in one pass is what makes CholQR2 cost
//      8nm (Gram) + 16nm (update fused with the Gram of the result) [+ 16nm second update, skipped when the
//      fused Gram already equals I to working precision]
// instead of 48nm bytes.
//
// The small dense products run on the FP64 tensor path, mma.sync.m8n8k4.f64 (SASS DMMA): measured on B200 it has
// the same peak as the FP64 FMA pipe (37.0 vs 34.8 TFLOP/s, tools/micro/fp64_peak.cu) -- tcgen05 has no FP64
// kind -- but one DMMA does 256 FMAs from two operand registers per lane. ncu of the FMA formulation
// (profiles/r01_ncu_kernels_v2.csv) showed it bound by shared-memory wavefronts (l1tex 65-80 %, FP64 pipe 30 %):
// a 4x8 register block needs 6 128-bit shared loads per 32 FMAs. With DMMA a 4-row slab of the tile needs M/8
// 64-bit fragment loads for M/8*(M/8+1)/2 Gram tiles (the same fragment serves as A and as B operand when Y = X),
// and the update's B operand (the M x M factor) lives in registers for the whole kernel.
//
// Shared-memory layout: row stride M+4 doubles. Both fragment patterns -- (4 rows x 8 columns) for the Gram
// operands and (8 rows x 4 columns) for the update's A operand -- then touch 16 distinct 8-byte bank pairs per
// half-warp, i.e. every fragment load is conflict-free; rows stay 16-byte aligned for cp.async.
//
// Roofline: HBM-bound for M <= 32 (16nm or 8nm bytes); at M = 64 the FP64 pipe (2 n M^2 flops, halved by
// the triangular / symmetric structure) is the second limiter.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "de_types.hpp"
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

  /** D(8x8) += A(8x4) * B(4x8), fp64. Lane l holds a = A[l/4][l%4], b = B[l%4][l/4], c0 = C[l/4][2(l%4)], c1 = C[l/4][2(l%4)+1]. */
  __device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
  {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
  }

  template <int M, bool UPPER, int NOPS>
  struct TsCfg
  {
    static constexpr int THREADS = 256;
    static constexpr int STAGES = 3;
    static constexpr int TR = (NOPS == 2 ? 2048 : 4096) / M; // rows per tile: ~32 KB of operand data per stage, so
                                                             // two stages in flight on 148 SMs cover HBM latency
    static constexpr int LDT = M + 4;                     // padded row stride (doubles), see file comment
    static constexpr int TILE = TR * LDT;                 // doubles per staged tile
    static constexpr int NB = M / 8;                      // 8-column blocks
    // update phase: a warp owns NJ column blocks (their factor fragments live in registers) and every RBG-th 8-row block
    static constexpr int NJ = (M == 64) ? 2 : NB;
    static constexpr int JG = NB / NJ;                    // column-block groups
    static constexpr int RBG = 8 / JG;                    // row-block groups
    static constexpr int KS = M / 4;                      // k steps of the update
    // Gram phase: 8x8 tiles of G split over TG tile groups, 4-row slabs of the tile split over KG slab groups
    static constexpr int NTILES = UPPER ? NB * (NB + 1) / 2 : NB * NB;
    static constexpr int TG = (M == 64) ? 2 : 1;
    static constexpr int KG = 8 / TG;
    static constexpr int NTW = (NTILES + TG - 1) / TG;    // tiles per warp
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
    int upper;            // R is upper triangular: column block c only needs k < 8c+8
    const int *skip_flag; // optional: if *skip_flag != 0 the kernel returns immediately (second CholQR sweep)
    const int *done;      // optional: a driver loop has converged, the launch is a no-op
    double *partials;     // DO_GRAM: [gridDim.x][M*M]
    PushRanges push;      // ts2_update_kernel: rows also stored into the neighbours' halo buffers (n = 0: none)
  };

  /** see file comment. Template switches:
   *   DO_UPDATE  Out = X R (tile staged in shared memory first, so Out may alias X)
   *   DO_GRAM    accumulate G = A^T B over all rows, A = (DO_UPDATE ? updated X : X), B = (SAME ? A : Y)
   *   UPPER      G is symmetric: only the 8x8 tiles on or above the diagonal are computed, mirrored on output
   *   SAME       single Gram operand (B aliases A) */
  template <int M, bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  __global__ void __launch_bounds__(256, (M <= 32) ? 2 : 1) tall_skinny_kernel(const TsArgs a)
  {
    constexpr int NOPS = (DO_GRAM && !SAME) ? 2 : 1;
    using C = TsCfg<M, UPPER, NOPS>;
    static_assert(!(DO_UPDATE && DO_GRAM) || SAME, "fused update+Gram works on one operand");
    extern __shared__ __align__(16) double smem[];
    double *tiles = smem; // STAGES x NOPS x TILE

    if (a.skip_flag != nullptr && *a.skip_flag != 0)
      return;
    if (a.done != nullptr && *a.done != 0)
      return;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lq = lane >> 2, lr = lane & 3; // lane / 4, lane % 4

    // ---- update: factor fragments of this warp's column blocks, B[k][j] = R[k0+k][j0+j] -> lane holds R[k0+lr][j0+lq]
    const int jg = warp % C::JG, rbg = warp / C::JG;
    double rfrag[DO_UPDATE ? C::KS : 1][DO_UPDATE ? C::NJ : 1];
    if (DO_UPDATE)
    {
#pragma unroll
      for (int ks = 0; ks < C::KS; ++ks)
#pragma unroll
        for (int q = 0; q < C::NJ; ++q)
          rfrag[ks][q] = __ldg(a.R + (4 * ks + lr) * M + 8 * (jg * C::NJ + q) + lq);
    }

    // ---- Gram: this warp's tiles (bi <= bj when UPPER), enumerated row-major, dealt round-robin to the tile groups
    const int tgi = warp % C::TG, kgi = warp / C::TG;
    double gacc[DO_GRAM ? C::NTW : 1][2];
#pragma unroll
    for (int w = 0; w < (DO_GRAM ? C::NTW : 1); ++w)
      gacc[w][0] = gacc[w][1] = 0.0;

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
        // ---- Y = X R on the staged tile, 8-row blocks: A[i][k] = X[rb*8+i][k0+k] -> lane loads X[rb*8+lq][k0+lr]
        constexpr int NRB = C::TR / 8;
        constexpr int MAXRB = (NRB + C::RBG - 1) / C::RBG;
        double out[MAXRB][C::NJ][2];
#pragma unroll
        for (int i = 0; i < MAXRB; ++i)
        {
          const int rb = rbg + i * C::RBG;
#pragma unroll
          for (int q = 0; q < C::NJ; ++q)
            out[i][q][0] = out[i][q][1] = 0.0;
          if (rb < NRB)
          {
            const double *xr = Xs + (rb * 8 + lq) * C::LDT + lr;
#pragma unroll
            for (int ks = 0; ks < C::KS; ++ks)
            {
              const double av = xr[4 * ks];
#pragma unroll
              for (int q = 0; q < C::NJ; ++q)
              {
                // triangular factor: column block jb only sees k < 8 jb + 8, i.e. k steps ks <= 2 jb + 1 (warp-uniform)
                const int jb = jg * C::NJ + q;
                if (!a.upper || ks <= 2 * jb + 1)
                  dmma884(out[i][q][0], out[i][q][1], av, rfrag[ks][q]);
              }
            }
          }
        }
        __syncthreads(); // every warp has finished READING the tile: the result may now overwrite it in place
#pragma unroll
        for (int i = 0; i < MAXRB; ++i)
        {
          const int rb = rbg + i * C::RBG;
          if (rb < NRB)
#pragma unroll
            for (int q = 0; q < C::NJ; ++q)
              st2(Xs + (rb * 8 + lq) * C::LDT + 8 * (jg * C::NJ + q) + 2 * lr, make_double2(out[i][q][0], out[i][q][1]));
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
        // ---- G += A^T B over 4-row slabs: fragment of column block c = X[slab*4 + lr][8c + lq] (A and B alike)
        const double *Ys = SAME ? Xs : Xs + C::TILE;
        for (int slab = kgi; slab < C::TR / 4; slab += C::KG)
        {
          double fa[C::NB], fb[SAME ? 1 : C::NB];
          const double *xs = Xs + (slab * 4 + lr) * C::LDT + lq;
          const double *ys = Ys + (slab * 4 + lr) * C::LDT + lq;
#pragma unroll
          for (int c = 0; c < C::NB; ++c)
          {
            fa[c] = xs[8 * c];
            if (!SAME)
              fb[c] = ys[8 * c];
          }
          // tile idx (compile-time after unrolling) belongs to tile group idx % TG and is slot idx / TG of that group
          int idx = 0;
#pragma unroll
          for (int bi = 0; bi < C::NB; ++bi)
#pragma unroll
            for (int bj = (UPPER ? bi : 0); bj < C::NB; ++bj)
            {
              if (idx % C::TG == tgi)
                dmma884(gacc[idx / C::TG][0], gacc[idx / C::TG][1], fa[bi], SAME ? fa[bj] : fb[bj]);
              ++idx;
            }
        }
      }
      __syncthreads(); // all reads of this stage are done: it may be refilled by the next iteration's prefetch
      stage = (stage + 1) % C::STAGES;
    }
    cp_async_wait<0>();

    if (DO_GRAM)
    {
      // combine the slab groups in fixed order into one M x M matrix in shared memory, then emit the CTA partial
      __syncthreads();
      double *G = smem; // M x M
      for (int turn = 0; turn < C::KG; ++turn)
      {
        if (kgi == turn)
        {
          int idx = 0;
#pragma unroll
          for (int bi = 0; bi < C::NB; ++bi)
#pragma unroll
            for (int bj = (UPPER ? bi : 0); bj < C::NB; ++bj)
            {
              if (idx % C::TG == tgi)
              {
                double *g = G + (8 * bi + lq) * M + 8 * bj + 2 * lr;
                const double v0 = gacc[idx / C::TG][0], v1 = gacc[idx / C::TG][1];
                if (turn == 0)
                  st2(g, make_double2(v0, v1));
                else
                {
                  const double2 o = ld2(g);
                  st2(g, make_double2(o.x + v0, o.y + v1));
                }
              }
              ++idx;
            }
        }
        __syncthreads();
      }
      double *outp = a.partials + (size_t)blockIdx.x * M * M;
      for (int e = tid; e < M * M; e += C::THREADS)
      {
        const int i = e / M, j = e % M;
        if (!UPPER)
          outp[e] = G[e];
        else if ((i >> 3) <= (j >> 3))
        {
          // tiles on or above the block diagonal were computed in full; mirror them below it
          outp[e] = G[e];
          if ((i >> 3) < (j >> 3))
            outp[j * M + i] = G[e];
        }
      }
    }
  }

  template <int M, bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  constexpr size_t tall_skinny_smem_bytes()
  {
    constexpr int NOPS = (DO_GRAM && !SAME) ? 2 : 1;
    using C = TsCfg<M, UPPER, NOPS>;
    size_t d = (size_t)C::STAGES * NOPS * C::TILE;
    const size_t g = (size_t)M * M;
    if (d < g)
      d = g;
    return d * sizeof(double);
  }

} // namespace de
