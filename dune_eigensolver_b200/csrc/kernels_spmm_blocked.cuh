// BRB SpMM for sm_100a: 8-row blocks x 4-column steps on the FP64 tensor path, X rows staged per tile in shared
// memory by a warp-specialised TMA / mbarrier pipeline. Format: brb_format.hpp.
//
// Why this shape (all measured on B200, logs under profiles/):
//  * tools/micro/gather_probe.cu: gathers of whole 8m-byte rows out of L2 top out near 10 TB/s for the chip -- 1.5x
//    HBM. The CSR kernels (kernels_sparse.cuh) fetch most X rows of a stencil row from L2 again and again (ncu:
//    3.0 GB of L2->L1 traffic for 0.83 GB of algorithmic bytes on the 27-point 100^3 matrix) and sit on that
//    ceiling and on the L1 data pipe (77 %), not on HBM.
//  * A TILE (6x4x4 grid points for a 27-point stencil) stages the union of its X rows ONCE in shared memory:
//    2.9 X rows per matrix row cross the fabric instead of ~11; DRAM traffic is at the algorithmic minimum.
//  * A ROW BLOCK is 8 matrix rows (a 2x2x2 sub-box); the sorted union of its columns is cut into steps of 4
//    columns; a step is C(8 rows x 8 cols) += A(8 x 4) * B(4 x 8) per 8-column panel of X on the FP64 tensor pipe
//    (mma.sync.m8n8k4.f64, SASS DMMA; tcgen05 has no FP64 kind). A = the 8 x 4 slice of the sparse matrix (one
//    value per lane, zero where the pattern bit is clear), B = 4 staged X rows. An X value is read from shared
//    memory once per 8 matrix rows and the broadcast of the matrix value happens inside the tensor pipe -- the CSR
//    kernels spent one shared-memory wavefront per nonzero just on that broadcast.
//  * Producer warps keep a ring of tile buffers full: the tile's packed matrix stream is one cp.async.bulk (SASS
//    UBLKCP) completing on an mbarrier, its X rows are 16-byte cp.async copies that arrive on the same mbarrier
//    (cp.async.mbarrier.arrive.noinc); a bulk copy per row was 3x slower (UBLKCP issues from the uniform datapath,
//    one lane at a time). Consumer warps wait on the full barrier, run their row blocks out of shared memory and
//    release the buffer through the empty barrier; nothing in the loop is a CTA-wide barrier.
//
// Roofline: HBM. Algorithmic bytes 12 nnz + 4 (n+1) + 16 n m; the BRB stream itself is ~9.4 bytes per nonzero.
// Second limiter: the FP64 tensor pipe (zero fill: 43 % of the DMMA lanes carry nonzeros for a 27-point stencil).
//
// Operand order of the GRAM variant: the product is formed TRANSPOSED, C(8 X-columns x 8 rows) += Xt(8 x 4) * At(4 x 8):
// the same two registers as for C = A * X, swapped in the instruction. A lane then owns a 2-row x NP-column patch of Y (rows 2k,
// 2k+1 of the block, columns NP g .. NP g + NP - 1): its stores are contiguous per row, and -- the point -- the
// accumulators of one quad of lanes already contain, up to a shuffle inside the quad, the operand fragments of the
// next tensor product over the rows: the Gram matrix Y^T Y of the result (GRAM epilogue) is accumulated in the same
// kernel, 20 extra DMMA per row block, and an orthonormalisation of Y can start without its own pass over Y. Measured:
// the epilogue adds 0.12 ms to a 0.19 ms SpMM (serialised accumulator chains, 40 more registers) where the separate Gram
// pass costs 0.05 ms -- de_spmm_gram offers it, the drivers keep the separate pass (kFuseGramIntoSpmm in de_capi.cu).
//
// Arithmetic: a row's products are summed in ascending column order in groups of four inside the DMMA instead of
// one FMA chain: results agree with the CSR order to rounding (bit-identical in the lab runs). A zero pattern entry
// multiplies an X value by 0.0, so a NaN/Inf in X reaches all 8 rows of every block that stages that X row.
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

#include "kernels_sparse.cuh"

namespace de
{
  __device__ __forceinline__ void dmma884_sp(double &c0, double &c1, double a, double b)
  {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
  }

  /** NP doubles from p (NP = 1, 2, 4, 8; 8 NP-byte aligned) if pred, else zeros. */
  template <int NP>
  __device__ __forceinline__ void ldg_row_if(double (&b)[NP], const double *p, bool pred)
  {
#pragma unroll
    for (int i = 0; i < NP; ++i)
      b[i] = 0.0;
    if (pred)
    {
      if constexpr (NP == 1)
        b[0] = __ldg(p);
      else if constexpr (NP == 2)
      {
        const double2 w = __ldg(reinterpret_cast<const double2 *>(p));
        b[0] = w.x;
        b[1] = w.y;
      }
      else
      {
#pragma unroll
        for (int q = 0; q < NP / 4; ++q)
          asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];\n"
                       : "=d"(b[4 * q]), "=d"(b[4 * q + 1]), "=d"(b[4 * q + 2]), "=d"(b[4 * q + 3])
                       : "l"(p + 4 * q));
      }
    }
  }

  template <int NP>
  __device__ __forceinline__ void stg_row(double *p, const double (&v)[NP])
  {
    if constexpr (NP == 1)
      p[0] = v[0];
    else if constexpr (NP == 2)
      *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    else
    {
#pragma unroll
      for (int q = 0; q < NP / 4; ++q)
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};\n" ::"l"(p + 4 * q), "d"(v[4 * q]), "d"(v[4 * q + 1]),
                     "d"(v[4 * q + 2]), "d"(v[4 * q + 3])
                     : "memory");
    }
  }

  // (An XOR swizzle of the staged chunks -- chunk c of row i stored at c ^ (i & 1) -- gives the same conflict-free
  //  bank pattern without the selects below, but measured 12-20 % SLOWER on B200: the swizzle bit depends on the
  //  step record and lengthens the address -> load -> DMMA chain. Kept as the rotation.)
  template <int NP>
  __device__ __forceinline__ void lds_frag(double (&b)[NP], const double *p, int k)
  {
    // p: this lane's NP doubles in the staged row. 128-bit loads are issued in an order rotated by k so that the
    // 8 lanes of a quarter warp (2 g x 4 k, rows of stride M+4 doubles) hit 8 distinct 16-byte bank groups.
    if constexpr (NP == 1)
      b[0] = p[0];
    else
    {
      constexpr int NV = NP / 2;
#pragma unroll
      for (int j = 0; j < NV; ++j)
      {
        const int jj = (j + k) % NV;
        const double2 w = *reinterpret_cast<const double2 *>(p + 2 * jj);
        // b must be indexed with compile-time constants: select into place
#pragma unroll
        for (int q = 0; q < NV; ++q)
          if (q == jj)
          {
            b[2 * q] = w.x;
            b[2 * q + 1] = w.y;
          }
      }
    }
  }

  __device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
  __device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
  }
  __device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void mbar_arrive(unsigned bar)
  {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
  }
  __device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
  {
    asm volatile("{\n .reg .pred p;\nWAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(bar),
                 "r"(parity)
                 : "memory");
  }
  __device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
  {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
  }


  constexpr int kBrbProducerWarps = 4;      // measured: 2 producer warps starve the pipeline (0.210 ms), 4: 0.181 ms, 8: 0.183 ms
  constexpr int kBrbConsumerWarps = 12;     // plain / DOT variants
  constexpr int kBrbConsumerWarpsGram = 8;  // GRAM variants carry 40 more accumulator registers per thread (168 in all)
  constexpr int brb_threads(bool gram) { return 32 * (kBrbProducerWarps + (gram ? kBrbConsumerWarpsGram : kBrbConsumerWarps)); }
  constexpr int kBrbMaxStages = 4;
  constexpr int kBrbBarrierBytes = 128; // mbarriers in front of the stage buffers

  struct BrbArgs
  {
    int ntiles;               // tiles of this launch
    long long n;              // rows of Y (bounds the output row ids)
    const int4 *tile;         // [ntiles] {blob16, len16, ucol0, nu} (brb::TileDesc)
    const int4 *blob;         // tile blobs (16-byte units)
    const int *ucol;          // union column ids of all tiles
    const double *X;          // owned rows of the input block (already offset to the first column of this pass)
    const double *H;          // halo rows (columns >= n_owned), same column offset
    int n_owned;
    int ldx;                  // row stride of X, H and Y in doubles
    double *Y;
    double *partials;         // DOT: per-CTA partial dot products, partials[cta * pstride + column]
    int pstride;              //      GRAM: followed by the CTA's partial Gram matrix at partials[cta * pstride + gram_off + i * M + j]
    int gram_off;
    int blob_cap16;           // shared-memory capacity of one stage: blob (16-byte units) ...
    int xs_cap;               // ... and staged X rows
    int stages;               // pipeline depth (2..kBrbMaxStages)
    const int *done;          // optional device flag: a driver loop has converged, the launch is a no-op
    // EPI = 1 (Chebyshev epilogue, LOBPCG's polynomial preconditioner): instead of storing Y = A X the kernel updates
    //   E0(i,:) <- X(i,:) + ealpha (X(i,:) - E0(i,:)) + ebeta edinv[i] (E1(i,:) - Y(i,:))
    // (cheb_step_kernel, kernels_lobpcg.cuh, with Z = X, Zold = E0, R = E1): A z never travels to HBM and back.
    double *E0;
    const double *E1;
    const double *edinv;
    double ealpha, ebeta;
  };

  /** bytes of dynamic shared memory for a pass of 8 NP columns */
  inline size_t spmm_brb_smem_bytes(int np, int blob_cap16, int xs_cap, int stages)
  {
    const size_t ring = (size_t)stages * ((size_t)blob_cap16 * 16 + (size_t)xs_cap * (8 * np + 4) * sizeof(double));
    // the stage buffers double as scratch of the final reductions (consumer warps x M dot partials, M x M Gram): tiny
    // matrices have tiles smaller than that
    const size_t scratch = ((size_t)kBrbConsumerWarps * 8 * np + (size_t)64 * np * np) * sizeof(double);
    return (size_t)kBrbBarrierBytes + (ring > scratch ? ring : scratch);
  }

  /** Y(:, 0 : 8 NP) = A X(:, 0 : 8 NP) for the tiles of one launch (+ per-CTA partials of diag(X^T Y) when DOT).
   *  Lane l = 4 g + k of a consumer warp:
   *    A fragment   a      = A(block row g, step column k)                       (zero if the pattern bit is clear)
   *    B fragment   b[p]   = X(step column k, NP g + p), p < NP                   (NP/2 128-bit shared loads)
   *    accumulators c[p]   = Y(row g, 2 k NP + p), Y(row g, (2 k + 1) NP + p)     -> the lane owns 2 NP contiguous columns
   *  i.e. panel p of the tensor product covers the X columns {NP j + p : j < 8}. */
  template <int NP, bool DOT, bool HALO, bool GRAM, int EPI = 0>
  __global__ void __launch_bounds__(brb_threads(GRAM), 1) spmm_brb_kernel(const BrbArgs a)
  {
    static_assert(!GRAM || DOT, "the Gram epilogue is only built together with the dot epilogue");
    static_assert(EPI == 0 || (!DOT && !GRAM), "the Chebyshev epilogue replaces the plain store");
    constexpr int NPW = kBrbProducerWarps, NCW = GRAM ? kBrbConsumerWarpsGram : kBrbConsumerWarps;
    constexpr int NT = NP * (NP + 1) / 2; // Gram tiles (p <= p') of the column groups {NP j + p}
    constexpr int M = 8 * NP;
    constexpr int LDR = M + 4; // staged row stride (doubles): see lds_frag
    extern __shared__ __align__(128) unsigned char dynq[];
    pdl_prologue();
    if (a.done != nullptr && *a.done != 0)
      return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stages = a.stages;
    const size_t buf_bytes = (size_t)a.blob_cap16 * 16 + (size_t)a.xs_cap * LDR * sizeof(double);
    unsigned char *bufs = dynq + kBrbBarrierBytes;
    const unsigned bar0 = smem_u32(dynq); // full[s] at bar0 + 8 s, empty[s] at bar0 + 8 (kBrbMaxStages + s)

    if (tid == 0)
    {
      for (int s = 0; s < stages; ++s)
      {
        mbar_init(bar0 + 8 * s, 1 + 32 * NPW);              // blob (expect_tx arrival) + one arrival per producer thread
        mbar_init(bar0 + 8 * (kBrbMaxStages + s), NCW);     // one arrival per consumer warp
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    double dacc[DOT ? 2 * NP : 1]; // direct form: 2 NP columns of one row; transposed (GRAM) form: NP columns of two rows
#pragma unroll
    for (int i = 0; i < (DOT ? 2 * NP : 1); ++i)
      dacc[i] = 0.0;
    double gacc[GRAM ? NT : 1][2];
#pragma unroll
    for (int i = 0; i < (GRAM ? NT : 1); ++i)
      gacc[i][0] = gacc[i][1] = 0.0;

    if (warp < NPW)
    {
      // ---------------- producers ----------------
      // Column ids: lane l keeps ucol[32 q + l] in registers (fetched one tile ahead); a row gets its id by shuffle.
      constexpr int CPR = M / 2;          // 16-byte chunks per row
      constexpr int RPR = 32 * NPW / CPR; // rows per round of all producer threads
      constexpr int MAXQ = 16;            // union rows <= 512 (brb::kMaxUnion)
      static_assert(CPR <= 32 && 32 % CPR == 0, "a row's chunks must fit one warp");
      const int ptid = warp * 32 + lane;
      const int c = ptid % CPR, rsub = ptid / CPR; // this thread's chunk, and its row within a round
      int creg[MAXQ], nreg[MAXQ];
      const double *Hm = HALO ? a.H - (size_t)a.n_owned * a.ldx : a.X;
      int4 d = make_int4(0, 0, 0, 0), dn = make_int4(0, 0, 0, 0);
      auto load_tile_meta = [&](int t, int4 &dd, int(&rr)[MAXQ])
      {
        if (t < a.ntiles)
        {
          dd = __ldg(a.tile + t);
#pragma unroll
          for (int q = 0; q < MAXQ; ++q)
            rr[q] = (32 * q + lane < dd.w) ? __ldg(a.ucol + dd.z + 32 * q + lane) : 0;
        }
      };
      load_tile_meta(blockIdx.x, d, creg);
      int s = 0, use = 0;
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x)
      {
        load_tile_meta(t + gridDim.x, dn, nreg); // in flight while this tile's copies are issued
        if (use > 0)
          mbar_wait(bar0 + 8 * (kBrbMaxStages + s), (unsigned)((use - 1) & 1));
        unsigned char *base = bufs + s * buf_bytes;
        const unsigned full = bar0 + 8 * s;
        if (ptid == 0)
        {
          mbar_expect_tx(full, (unsigned)d.y * 16u);
          bulk_g2s(smem_u32(base), a.blob + d.x, (unsigned)d.y * 16u, full);
        }
        double *xs = reinterpret_cast<double *>(base + (size_t)a.blob_cap16 * 16);
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
        {
          if (32 * q < d.w) // warp-uniform
          {
#pragma unroll
            for (int r = 0; r < 32 / RPR; ++r)
            {
              const int il = r * RPR + rsub; // row within this group of 32
              const int col = __shfl_sync(0xffffffffu, creg[q], il);
              const int i = 32 * q + il;
              if (i < d.w)
              {
                // halo rows: Hm + col * ldx == H + (col - n_owned) * ldx
                const double *base = (HALO && col >= a.n_owned) ? Hm : a.X;
                cp_async16_sparse(xs + i * LDR + 2 * c, base + (size_t)col * a.ldx + 2 * c);
              }
            }
          }
        }
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(full) : "memory");
        d = dn;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
          creg[q] = nreg[q];
        if (++s == stages)
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
      const unsigned lt = (1u << lane) - 1u;
      const int lsh = 16 * (k & 1);
      int s = 0, use = 0;
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x)
      {
        mbar_wait(bar0 + 8 * s, (unsigned)(use & 1));
        const unsigned char *base = bufs + s * buf_bytes;
        const int *hdr = reinterpret_cast<const int *>(base);
        const int nb = hdr[0];
        const int *blkstep = hdr + 4;
        const int *blkrows = blkstep + nb + 1;
        const unsigned short *blkself = reinterpret_cast<const unsigned short *>(blkrows + 8 * nb);
        const int o_step = (4 + (nb + 1) + 8 * nb + 4 * nb + 3) & ~3;
        const int4 *steps = reinterpret_cast<const int4 *>(hdr + o_step);
        const double *sval = reinterpret_cast<const double *>(hdr + o_step + 4 * hdr[1]);
        const double *xs = reinterpret_cast<const double *>(base + (size_t)a.blob_cap16 * 16);

        for (int blk = cw; blk < nb; blk += NCW)
        {
          const int s0 = blkstep[blk], s1 = blkstep[blk + 1];
          double c[NP][2];
#pragma unroll
          for (int p = 0; p < NP; ++p)
            c[p][0] = c[p][1] = 0.0;
          if (EPI == 1)
          {
            // the epilogue's operands of this lane's row: on their way while the steps run (no registers held)
            const long long prow = (long long)blkrows[8 * blk + g];
            if (prow >= 0 && prow < a.n)
            {
              asm volatile("prefetch.global.L2 [%0];\n" ::"l"(a.E0 + (size_t)prow * a.ldx + 2 * k * NP));
              asm volatile("prefetch.global.L2 [%0];\n" ::"l"(a.E1 + (size_t)prow * a.ldx + 2 * k * NP));
            }
          }
#pragma unroll 4
          for (int st = s0; st < s1; ++st)
          {
            const int4 r = steps[st]; // {lc0 | lc1 << 16, lc2 | lc3 << 16, mask, first value}
            const unsigned mask = (unsigned)r.z;
            const int lc = (int)(((unsigned)((k & 2) ? r.y : r.x) >> lsh) & 0xffffu);
            const bool has = (mask >> lane) & 1u;
            const double av = has ? sval[r.w + __popc(mask & lt)] : 0.0;
            double bv[NP];
            lds_frag<NP>(bv, xs + lc * LDR + NP * g, k);
#pragma unroll
            for (int p = 0; p < NP; ++p)
            {
              if (GRAM)
                dmma884_sp(c[p][0], c[p][1], bv[p], av); // transposed: c[p] = Y(rows 2k, 2k+1; column NP g + p)
              else
                dmma884_sp(c[p][0], c[p][1], av, bv[p]); // c[p] = Y(row g; columns 2 k NP + p, (2 k + 1) NP + p)
            }
          }
          if (!GRAM)
          {
            // the lane owns 2 NP contiguous columns of row g (measured 5 % faster than the transposed epilogue)
            const long long row = (long long)blkrows[8 * blk + g];
            if (row >= 0 && row < a.n)
            {
              double lo[NP], hi[NP];
#pragma unroll
              for (int p = 0; p < NP; ++p)
              {
                lo[p] = c[p][0];
                hi[p] = c[p][1];
              }
              if (EPI == 1)
              {
                // z = X(row, :) from the staged tile (the row's own column is among the tile's columns for any matrix with a
                // stored diagonal), else from global memory; zo, r from global memory (prefetched at the top of the block)
                double zl[NP], zh[NP], ol[NP], oh[NP], rl[NP], rh[NP];
                const unsigned self = blkself[8 * blk + g];
                if (self != 0xffffu)
                {
                  const double *zr = xs + self * LDR + 2 * k * NP;
#pragma unroll
                  for (int p = 0; p < NP; ++p)
                  {
                    zl[p] = zr[p];
                    zh[p] = zr[NP + p];
                  }
                }
                else
                {
                  const double *xr = a.X + (size_t)row * a.ldx + 2 * k * NP;
                  ldg_row_if<NP>(zl, xr, true);
                  ldg_row_if<NP>(zh, xr + NP, true);
                }
                double *er = a.E0 + (size_t)row * a.ldx + 2 * k * NP;
                const double *rr = a.E1 + (size_t)row * a.ldx + 2 * k * NP;
                ldg_row_if<NP>(ol, er, true);
                ldg_row_if<NP>(oh, er + NP, true);
                ldg_row_if<NP>(rl, rr, true);
                ldg_row_if<NP>(rh, rr + NP, true);
                const double bd = a.ebeta * __ldg(a.edinv + row);
#pragma unroll
                for (int p = 0; p < NP; ++p)
                {
                  lo[p] = zl[p] + a.ealpha * (zl[p] - ol[p]) + bd * (rl[p] - lo[p]);
                  hi[p] = zh[p] + a.ealpha * (zh[p] - oh[p]) + bd * (rh[p] - hi[p]);
                }
                stg_row<NP>(er, lo);
                stg_row<NP>(er + NP, hi);
              }
              else
              {
                double *yr = a.Y + (size_t)row * a.ldx + 2 * k * NP;
                stg_row<NP>(yr, lo);
                stg_row<NP>(yr + NP, hi);
              }
              if (DOT)
              {
                // X(row, :) -- from the staged tile when the row's own column is among the tile's columns (any matrix with
                // a stored diagonal), else from global memory
                double zl[NP], zh[NP];
                const unsigned self = blkself[8 * blk + g];
                if (self != 0xffffu)
                {
                  const double *zr = xs + self * LDR + 2 * k * NP;
#pragma unroll
                  for (int p = 0; p < NP; ++p)
                  {
                    zl[p] = zr[p];
                    zh[p] = zr[NP + p];
                  }
                }
                else
                {
                  const double *xr = a.X + (size_t)row * a.ldx + 2 * k * NP;
                  ldg_row_if<NP>(zl, xr, true);
                  ldg_row_if<NP>(zh, xr + NP, true);
                }
#pragma unroll
                for (int p = 0; p < NP; ++p)
                {
                  dacc[p] = fma(zl[p], lo[p], dacc[p]);
                  dacc[NP + p] = fma(zh[p], hi[p], dacc[NP + p]);
                }
              }
            }
            continue;
          }
          const long long row0 = (long long)blkrows[8 * blk + 2 * k], row1 = (long long)blkrows[8 * blk + 2 * k + 1];
          double y0[NP], y1[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p)
          {
            y0[p] = c[p][0];
            y1[p] = c[p][1];
          }
          if (row0 >= 0 && row0 < a.n)
          {
            stg_row<NP>(a.Y + (size_t)row0 * a.ldx + NP * g, y0);
            if (DOT)
            {
              double z[NP];
              ldg_row_if<NP>(z, a.X + (size_t)row0 * a.ldx + NP * g, true);
#pragma unroll
              for (int p = 0; p < NP; ++p)
                dacc[p] = fma(z[p], y0[p], dacc[p]);
            }
          }
          if (row1 >= 0 && row1 < a.n)
          {
            stg_row<NP>(a.Y + (size_t)row1 * a.ldx + NP * g, y1);
            if (DOT)
            {
              double z[NP];
              ldg_row_if<NP>(z, a.X + (size_t)row1 * a.ldx + NP * g, true);
#pragma unroll
              for (int p = 0; p < NP; ++p)
                dacc[p] = fma(z[p], y1[p], dacc[p]);
            }
          }
          if (GRAM)
          {
            // G += Yb^T Yb for the 8 rows of the block (empty slots hold zeros), as two 4-row slabs. The operand
            // fragment of column group p is  f[p] = Y(row 4 s + k, column NP g + p): rows 2 kk, 2 kk + 1 sit in lane kk
            // of this quad, so lane k fetches from lane 2 s + (k >> 1), accumulator (k & 1).
#pragma unroll
            for (int sl = 0; sl < 2; ++sl)
            {
              const int src = (lane & ~3) | (2 * sl + (k >> 1));
              double f[NP];
#pragma unroll
              for (int p = 0; p < NP; ++p)
              {
                const double v0 = __shfl_sync(0xffffffffu, c[p][0], src);
                const double v1 = __shfl_sync(0xffffffffu, c[p][1], src);
                f[p] = (k & 1) ? v1 : v0;
              }
              int t = 0;
#pragma unroll
              for (int p = 0; p < NP; ++p)
#pragma unroll
                for (int q = p; q < NP; ++q)
                {
                  dmma884_sp(gacc[t][0], gacc[t][1], f[p], f[q]); // tile (p, q): G(NP i + p, NP j + q)
                  ++t;
                }
            }
          }
        }
        __syncwarp();
        if (lane == 0)
          mbar_arrive(bar0 + 8 * (kBrbMaxStages + s));
        if (++s == stages)
        {
          s = 0;
          ++use;
        }
      }
    }

    if (DOT)
    {
      // fold the lanes that hold the same columns, then the consumer warps, in fixed order; the Gram accumulators of the warps are folded the same way into the natural m x m layout
      __syncthreads(); // every tile of this CTA has been consumed: the stage buffers are free
      double *red = reinterpret_cast<double *>(bufs); // NCW x M
      double *gred = red + NCW * M;                   // GRAM: M x M
      const int g = lane >> 2, k = lane & 3;
      if (warp >= NPW)
      {
        if (GRAM)
        {
          // transposed form: the 4 lanes of a quad hold the same NP columns
#pragma unroll
          for (int i = 0; i < NP; ++i)
          {
            double v = dacc[i];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (k == 0)
              red[(warp - NPW) * M + NP * g + i] = v;
          }
        }
        else
        {
          // direct form: lanes with equal k hold the same 2 NP columns
#pragma unroll
          for (int i = 0; i < 2 * NP; ++i)
          {
            double v = dacc[i];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0)
              red[(warp - NPW) * M + 2 * k * NP + i] = v;
          }
        }
      }
      if (GRAM)
      {
        for (int turn = 0; turn < NCW; ++turn)
        {
          if (warp - NPW == turn)
          {
            int t = 0;
#pragma unroll
            for (int p = 0; p < NP; ++p)
#pragma unroll
              for (int q = p; q < NP; ++q)
              {
                // accumulator tile (p, q): D(i = g, j = 2k, 2k+1) = G(NP g + p, NP j + q)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                {
                  const int gi = NP * g + p, gj = NP * (2 * k + e) + q;
                  const double v = gacc[t][e] + (turn == 0 ? 0.0 : gred[gi * M + gj]);
                  gred[gi * M + gj] = v;
                }
                ++t;
              }
          }
          __syncthreads();
        }
      }
      else
        __syncthreads();
      if (tid < M)
      {
        double sum = 0.0;
        for (int w = 0; w < NCW; ++w)
          sum += red[w * M + tid];
        a.partials[(size_t)blockIdx.x * a.pstride + tid] = sum;
      }
      if (GRAM)
      {
        // tiles with p < q were computed for (row % NP, column % NP) = (p, q) only: their mirror images complete G
        double *out = a.partials + (size_t)blockIdx.x * a.pstride + a.gram_off;
        for (int e = tid; e < M * M; e += blockDim.x)
        {
          const int i = e / M, j = e % M;
          out[e] = (i % NP <= j % NP) ? gred[e] : gred[j * M + i];
        }
      }
    }
  }

} // namespace de
