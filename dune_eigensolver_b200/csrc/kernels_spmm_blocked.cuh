// Third-generation SpMM kernels for sm_100a.
//
// What bounded the earlier kernels (ncu, profiles/r01_ncu_kernels_v3.csv): the L1/shared-memory data pipe
// (128 B per clock per SM). Every nonzero moved its 8*m-byte X row through that pipe once per MATRIX ROW that
// references it (27 times for a Q1 stencil), and the warp-broadcast reads of the column index and the value cost
// almost as many wavefronts again. Two answers:
//
//  spmm_csr4_kernel   plain CSR; indices and values are read from the staged stream four at a time with 128-bit
//                     shared loads, and the gathers of a chunk are predicated instead of padded, so the pipe only
//                     carries useful X rows.
//
//  spmm_brb8_kernel   "BRB8": rows are grouped in blocks of 8; the union of a block's columns is cut into steps of
//                     4 columns; a step is one FP64 tensor-core product  C(8 rows x 8 cols) += A(8 x 4) * B(4 x 8)
//                     per 8-column panel of X (mma.sync.m8n8k4.f64, SASS DMMA -- tcgen05 has no FP64 kind). The
//                     A fragment is the 8 x 4 slice of the sparse matrix (one value per lane, zero where the
//                     pattern is empty), the B fragment is a slice of 4 X rows that every lane loads with one
//                     256-bit load. An X row is therefore loaded ONCE per 8 matrix rows and the value broadcast
//                     happens inside the tensor pipe, not through shared memory.
#pragma once

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "kernels_sparse.cuh"

namespace de
{

  // ================================================================================================
  // VA: staged CSR, vectorised metadata, predicated gathers
  // ================================================================================================
  constexpr int kCsr4CapNnz = 2048;
  constexpr int kCsr4MaxRows = 256;
  constexpr int kCsr4Pad = 16; // slack (elements) so that aligned 8-element reads never leave the buffers

  constexpr size_t spmm_csr4_smem_bytes()
  {
    return 2 * ((kCsr4MaxRows + 8) * sizeof(int) + (kCsr4CapNnz + kCsr4Pad) * sizeof(int) +
                (kCsr4CapNnz + kCsr4Pad) * sizeof(double));
  }

  __device__ __forceinline__ double2 ldg2_if(const double2 *p, bool pred)
  {
    double2 r;
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %3, 0;\n mov.f64 %0, 0d0000000000000000;\n mov.f64 %1, 0d0000000000000000;\n"
                 " @p ld.global.nc.v2.f64 {%0,%1}, [%2];\n}\n"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p), "r"((int)pred));
    return r;
  }

  /** Same contract as spmm_staged_kernel (row blocks of <= 256 rows / <= 2048 nonzeros cut on the host, CSR stream
   *  double-buffered through shared memory, TPR lanes own the m = 2 TPR columns of a row, CSR-order FMA). */
  template <int TPR, bool DOT, bool HALO>
  __global__ void __launch_bounds__(256, 3) spmm_csr4_kernel(const StagedArgs a)
  {
    constexpr int RPB = 256 / TPR;
    constexpr int CAPI = kCsr4CapNnz + kCsr4Pad;
    extern __shared__ __align__(16) unsigned char dyn[];
    double *s_val0 = reinterpret_cast<double *>(dyn);     // 2 x CAPI doubles
    int *s_col0 = reinterpret_cast<int *>(s_val0 + 2 * CAPI); // 2 x CAPI ints
    int *s_ptr0 = s_col0 + 2 * CAPI;                      // 2 x (MAXROWS+8) ints

    const int tid = threadIdx.x;
    const int t = tid % TPR, gslot = tid / TPR;
    const unsigned ldh = (unsigned)TPR;
    const double2 *__restrict__ Xv = reinterpret_cast<const double2 *>(a.X) + t;
    const double2 *__restrict__ Hv = reinterpret_cast<const double2 *>(a.H) + t;
    double2 *__restrict__ Yv = reinterpret_cast<double2 *>(a.Y) + t;
    double2 dacc = make_double2(0.0, 0.0);

    auto stage_block = [&](int4 mt, int s)
    {
      if (mt.x < mt.y)
      {
        const int ra = mt.x & ~3;
        const int nptr = (mt.y + 1 - ra + 3) >> 2;
        int *sp = s_ptr0 + s * (kCsr4MaxRows + 8);
        for (int c = tid; c < nptr; c += 256)
          cp_async16_sparse(sp + 4 * c, a.rowptr + ra + 4 * c);
        if (mt.w - mt.z <= kCsr4CapNnz)
        {
          const int ka = mt.z & ~3;
          const int ncol = (mt.w - ka + 3) >> 2, nval = (mt.w - ka + 1) >> 1;
          int *sc = s_col0 + s * CAPI;
          double *sv = s_val0 + s * CAPI;
          for (int c = tid; c < ncol; c += 256)
            cp_async16_sparse(sc + 4 * c, a.col + ka + 4 * c);
          for (int c = tid; c < nval; c += 256)
            cp_async16_sparse(sv + 2 * c, a.val + ka + 2 * c);
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };
    auto load_meta = [&](int b) { return (b < a.nblocks) ? __ldg(a.blk_meta + b) : make_int4(0, 0, 0, 0); };

    int b = blockIdx.x;
    int4 cur = load_meta(b);
    int4 nxt = load_meta(b + gridDim.x);
    stage_block(cur, 0);
    int s = 0;
    for (; b < a.nblocks; b += gridDim.x)
    {
      stage_block(nxt, s ^ 1);
      const int4 nxt2 = load_meta(b + 2 * gridDim.x);
      asm volatile("cp.async.wait_group 1;\n" ::);
      __syncthreads();

      const bool direct = (cur.w - cur.z) > kCsr4CapNnz;
      const int *sp = s_ptr0 + s * (kCsr4MaxRows + 8) - (cur.x & ~3);
      // element k of the stream lives at cp[k] / vp[k]; k % 4 == 0 is 16-byte aligned in both
      const int *cp = direct ? a.col : s_col0 + s * CAPI - (cur.z & ~3);
      const double *vp = direct ? a.val : s_val0 + s * CAPI - (cur.z & ~3);

      for (int r = cur.x + gslot; r < cur.y; r += RPB)
      {
        const int kbeg = sp[r], kend = sp[r + 1];
        double2 acc = make_double2(0.0, 0.0);
        for (int k = kbeg & ~3; k < kend; k += 8)
        {
          const int4 ja = *reinterpret_cast<const int4 *>(cp + k);
          const int4 jb = *reinterpret_cast<const int4 *>(cp + k + 4);
          const int j[8] = {ja.x, ja.y, ja.z, ja.w, jb.x, jb.y, jb.z, jb.w};
          double av[8];
#pragma unroll
          for (int q = 0; q < 4; ++q)
          {
            const double2 w = *reinterpret_cast<const double2 *>(vp + k + 2 * q);
            av[2 * q] = w.x;
            av[2 * q + 1] = w.y;
          }
          double2 xv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
          {
            const bool in = (k + u >= kbeg) && (k + u < kend);
            const double2 *p = HALO ? ((j[u] < a.n_owned) ? Xv + (unsigned)j[u] * ldh : Hv + (unsigned)(j[u] - a.n_owned) * ldh)
                                    : Xv + (unsigned)j[u] * ldh;
            xv[u] = ldg2_if(p, in);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
          {
            const bool in = (k + u >= kbeg) && (k + u < kend);
            fma2(acc, in ? av[u] : 0.0, xv[u]);
          }
        }
        const int orow = a.rowmap ? __ldg(a.rowmap + r) : r;
        Yv[(unsigned)orow * ldh] = acc;
        if (DOT)
        {
          const double2 z = __ldg(Xv + (unsigned)orow * ldh);
          dacc.x = fma(z.x, acc.x, dacc.x);
          dacc.y = fma(z.y, acc.y, dacc.y);
        }
      }
      __syncthreads();
      cur = nxt;
      nxt = nxt2;
      s ^= 1;
    }
    asm volatile("cp.async.wait_group 0;\n" ::);

    if (DOT)
    {
      __syncthreads();
      double2 *red = reinterpret_cast<double2 *>(dyn);
      red[tid] = dacc;
      __syncthreads();
      if (gslot == 0)
      {
        double2 sum = make_double2(0.0, 0.0);
        for (int q = 0; q < RPB; ++q)
        {
          sum.x += red[q * TPR + t].x;
          sum.y += red[q * TPR + t].y;
        }
        st2(a.partials + (size_t)blockIdx.x * a.m + 2 * t, sum);
      }
    }
  }

  // ================================================================================================
  // VB: BRB8 -- 8-row blocks x 4-column steps on the FP64 tensor path
  // ================================================================================================
  struct Brb8Args
  {
    int nblocks;
    long long n;              // rows of Y (bounds the identity row map)
    const int *blkstep;       // [nblocks + 1] first step of each row block
    const int *blkval;        // [nblocks + 1] first packed value of each row block
    const int *stepcol;       // [4 * nsteps] the 4 (local) column indices of a step (short steps repeat the last one)
    const unsigned *stepmask; // [nsteps] bit 4 g + k set: row g of the block has an entry in column k of the step
    const double *val;        // packed values: step-major, within a step ascending bit position
    const int *blkrows;       // optional [8 * nblocks] output row of each block row (-1: none); null: rows 8 b + g
    const double *X;
    const double *H;          // halo rows (columns >= n_owned)
    int n_owned;
    double *Y;
    double *partials;         // DOT: [gridDim.x][m]
  };

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

  /** Y = A X with A in BRB8 form, m = 8 NP columns. One warp per row block; lane l = 4 g + k:
   *    A fragment   a        = A(row g, step column k)                          (zero if the pattern bit is clear)
   *    B fragment   b[p]     = X(step column k, NP g + p),  p < NP               (one NP*8-byte load per lane)
   *    accumulators c[p]     = Y(row g, 2 k NP + p), Y(row g, (2 k + 1) NP + p)  -> lane owns Y(row g, 2 k NP .. 2 k NP + 2 NP - 1)
   *  i.e. panel p of the tensor product covers the X columns {NP j + p : j < 8}.
   *  U steps are in flight per warp (metadata -> values and X rows -> DMMA). */
  template <int NP, bool DOT, bool HALO>
  __global__ void __launch_bounds__(256, 3) spmm_brb8_kernel(const Brb8Args a)
  {
    constexpr int U = 4;
    constexpr int M = 8 * NP;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, k = lane & 3;
    const unsigned lt = (1u << lane) - 1u;
    const int wpb = blockDim.x >> 5;
    double dacc[DOT ? 2 * NP : 1];
#pragma unroll
    for (int i = 0; i < (DOT ? 2 * NP : 1); ++i)
      dacc[i] = 0.0;

    for (int blk = blockIdx.x * wpb + warp; blk < a.nblocks; blk += gridDim.x * wpb)
    {
      const int s0 = __ldg(a.blkstep + blk), s1 = __ldg(a.blkstep + blk + 1);
      int voff = __ldg(a.blkval + blk);
      double c[NP][2];
#pragma unroll
      for (int p = 0; p < NP; ++p)
        c[p][0] = c[p][1] = 0.0;

      for (int s = s0; s < s1; s += U)
      {
        int col[U];
        unsigned mask[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
          // the arrays carry U steps of tail padding; steps past the block end are masked out
          col[u] = __ldg(a.stepcol + 4 * (s + u) + k);
          mask[u] = (s + u < s1) ? __ldg(a.stepmask + s + u) : 0u;
        }
        double av[U];
        double bv[U][NP];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
          const bool has = (mask[u] >> lane) & 1u;
          av[u] = has ? __ldg(a.val + voff + __popc(mask[u] & lt)) : 0.0;
          voff += __popc(mask[u]);
          const double *xr = HALO ? ((col[u] < a.n_owned) ? a.X + (size_t)col[u] * M : a.H + (size_t)(col[u] - a.n_owned) * M)
                                  : a.X + (size_t)col[u] * M;
          ldg_row_if<NP>(bv[u], xr + NP * g, mask[u] != 0u);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int p = 0; p < NP; ++p)
            dmma884_sp(c[p][0], c[p][1], av[u], bv[u][p]);
      }

      const long long row = a.blkrows ? (long long)__ldg(a.blkrows + 8 * blk + g) : (long long)8 * blk + g;
      if (row >= 0 && row < a.n)
      {
        double lo[NP], hi[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p)
        {
          lo[p] = c[p][0];
          hi[p] = c[p][1];
        }
        double *yr = a.Y + (size_t)row * M + 2 * k * NP;
        stg_row<NP>(yr, lo);
        stg_row<NP>(yr + NP, hi);
        if (DOT)
        {
          double zl[NP], zh[NP];
          const double *xr = a.X + (size_t)row * M + 2 * k * NP;
          ldg_row_if<NP>(zl, xr, true);
          ldg_row_if<NP>(zh, xr + NP, true);
#pragma unroll
          for (int p = 0; p < NP; ++p)
          {
            dacc[p] = fma(zl[p], lo[p], dacc[p]);
            dacc[NP + p] = fma(zh[p], hi[p], dacc[NP + p]);
          }
        }
      }
    }

    if (DOT)
    {
      // lanes with equal k hold the same 2 NP columns: fold the 8 row lanes, then the warps, in fixed order
      __shared__ double red[8][64];
#pragma unroll
      for (int i = 0; i < 2 * NP; ++i)
      {
        double v = dacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0)
          red[warp][2 * k * NP + i] = v;
      }
      __syncthreads();
      if (threadIdx.x < M)
      {
        double sum = 0.0;
        for (int w = 0; w < wpb; ++w)
          sum += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * M + threadIdx.x] = sum;
      }
    }
  }

  // ---- host-side construction of the BRB8 arrays (setup; one pass over the CSR rows) ------------------------
  struct Brb8Host
  {
    int nblocks = 0;
    std::vector<int> blkstep, blkval, stepcol, blkrows;
    std::vector<unsigned> stepmask;
    std::vector<double> val;
  };

  /** rows: optional list of the matrix rows to convert (block b holds rows[8b .. 8b+7]); null = all rows in order. */
  template <class Ptr, class Idx>
  void brb8_build_host(long long nrows, const Ptr *rowptr, const Idx *col, const double *val, const int *rows, Brb8Host &H)
  {
    const long long nb = (nrows + 7) / 8;
    H.nblocks = (int)nb;
    H.blkstep.assign(nb + 1, 0);
    H.blkval.assign(nb + 1, 0);
    H.stepcol.clear();
    H.stepmask.clear();
    H.val.clear();
    struct Ent
    {
      long long c;
      int g;
      double v;
    };
    std::vector<Ent> ent;
    std::vector<long long> ucol;
    for (long long b = 0; b < nb; ++b)
    {
      ent.clear();
      for (int g = 0; g < 8 && 8 * b + g < nrows; ++g)
      {
        const long long r = rows ? rows[8 * b + g] : 8 * b + g;
        for (long long q = rowptr[r]; q < rowptr[r + 1]; ++q)
          ent.push_back({(long long)col[q], g, val[q]});
      }
      std::sort(ent.begin(), ent.end(), [](const Ent &x, const Ent &y) { return x.c != y.c ? x.c < y.c : x.g < y.g; });
      ucol.clear();
      for (const Ent &e : ent)
        if (ucol.empty() || ucol.back() != e.c)
          ucol.push_back(e.c);
      size_t e0 = 0;
      for (size_t u0 = 0; u0 < ucol.size(); u0 += 4)
      {
        const size_t u1 = std::min(u0 + 4, ucol.size());
        unsigned mask = 0;
        double slot[32];
        size_t e = e0;
        for (size_t u = u0; u < u1; ++u)
          for (; e < ent.size() && ent[e].c == ucol[u]; ++e)
          {
            const int bit = 4 * ent[e].g + (int)(u - u0);
            if (mask & (1u << bit)) // duplicate entry of the CSR row: accumulate
              slot[bit] += ent[e].v;
            else
            {
              mask |= 1u << bit;
              slot[bit] = ent[e].v;
            }
          }
        e0 = e;
        for (int q = 0; q < 4; ++q)
          H.stepcol.push_back((int)ucol[std::min(u0 + q, u1 - 1)]);
        H.stepmask.push_back(mask);
        for (int bit = 0; bit < 32; ++bit)
          if (mask & (1u << bit))
            H.val.push_back(slot[bit]);
      }
      H.blkstep[b + 1] = (int)H.stepmask.size();
      H.blkval[b + 1] = (int)H.val.size();
    }
  }

} // namespace de

namespace de
{
  // ================================================================================================
  // VC: BRB8T -- BRB8 row blocks grouped into CTA tiles whose X rows are staged ONCE in shared memory
  // ================================================================================================
  // Measured on B200 (tools/micro/gather_probe.cu): gathers of whole rows out of L2 top out near 10 TB/s for the
  // whole chip, only 1.5x the HBM rate. A 27-point row block re-fetches most of its X rows from L2 (ncu: 3.0 GB of
  // L2->L1 traffic for 0.83 GB of algorithmic bytes), so the kernels above are bound by the L2->SM fabric, not by
  // HBM. Here the row blocks of one tile share one staged copy of the union of their X rows: a tile of w x h x d
  // grid points needs (w+2)(h+2)(d+2) X rows instead of 27 (or 90/8) per row.
  struct Brb8TArgs
  {
    int ntiles;
    long long n;
    const int4 *tile;          // {first union entry, end union entry, first row block, end row block}
    const int *ucol;           // union column ids of all tiles, concatenated
    const int *blkstep;        // [nblocks + 1]
    const int *blkval;         // [nblocks + 1]
    const unsigned short *steplc; // [4 * nsteps] tile-local index of the 4 columns of a step
    const unsigned *stepmask;  // [nsteps]
    const double *val;
    const int *blkrows;        // [8 * nblocks] output rows (-1: none)
    const double *X;
    const double *H;
    int n_owned;
    double *Y;
    double *partials;
  };

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

  template <int NP, bool DOT, bool HALO>
  __global__ void __launch_bounds__(256, 2) spmm_brb8t_kernel(const Brb8TArgs a)
  {
    constexpr int U = 4;
    constexpr int M = 8 * NP;
    constexpr int LDR = M + 4; // staged row stride (doubles)
    extern __shared__ __align__(16) double xs[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, k = lane & 3;
    const unsigned lt = (1u << lane) - 1u;
    double dacc[DOT ? 2 * NP : 1];
#pragma unroll
    for (int i = 0; i < (DOT ? 2 * NP : 1); ++i)
      dacc[i] = 0.0;

    for (int tl = blockIdx.x; tl < a.ntiles; tl += gridDim.x)
    {
      const int4 t = __ldg(a.tile + tl);
      const int nu = t.y - t.x;
      constexpr int CPR = M / 2; // 16-byte chunks per row
      for (int e = tid; e < nu * CPR; e += 256)
      {
        const int i = e / CPR, c = e % CPR;
        const int col = __ldg(a.ucol + t.x + i);
        const double *src = HALO ? ((col < a.n_owned) ? a.X + (size_t)col * M : a.H + (size_t)(col - a.n_owned) * M)
                                 : a.X + (size_t)col * M;
        cp_async16_sparse(xs + i * LDR + 2 * c, src + 2 * c);
      }
      asm volatile("cp.async.commit_group;\n" ::);
      asm volatile("cp.async.wait_group 0;\n" ::);
      __syncthreads();

      for (int blk = t.z + warp; blk < t.w; blk += 8)
      {
        const int s0 = __ldg(a.blkstep + blk), s1 = __ldg(a.blkstep + blk + 1);
        int voff = __ldg(a.blkval + blk);
        double c[NP][2];
#pragma unroll
        for (int p = 0; p < NP; ++p)
          c[p][0] = c[p][1] = 0.0;

        int lc[U];
        unsigned mask[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
        {
          lc[u] = __ldg(a.steplc + 4 * (s0 + u) + k);
          mask[u] = (s0 + u < s1) ? __ldg(a.stepmask + s0 + u) : 0u;
        }
        for (int s = s0; s < s1; s += U)
        {
          int nlc[U];
          unsigned nmask[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
          {
            nlc[u] = __ldg(a.steplc + 4 * (s + U + u) + k);
            nmask[u] = (s + U + u < s1) ? __ldg(a.stepmask + s + U + u) : 0u;
          }
          double av[U];
          double bv[U][NP];
#pragma unroll
          for (int u = 0; u < U; ++u)
          {
            const bool has = (mask[u] >> lane) & 1u;
            av[u] = has ? __ldg(a.val + voff + __popc(mask[u] & lt)) : 0.0;
            voff += __popc(mask[u]);
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
          {
            if (mask[u] != 0u)
              lds_frag<NP>(bv[u], xs + lc[u] * LDR + NP * g, k);
            else
            {
#pragma unroll
              for (int p = 0; p < NP; ++p)
                bv[u][p] = 0.0;
            }
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int p = 0; p < NP; ++p)
              dmma884_sp(c[p][0], c[p][1], av[u], bv[u][p]);
#pragma unroll
          for (int u = 0; u < U; ++u)
          {
            lc[u] = nlc[u];
            mask[u] = nmask[u];
          }
        }

        const long long row = (long long)__ldg(a.blkrows + 8 * blk + g);
        if (row >= 0 && row < a.n)
        {
          double lo[NP], hi[NP];
#pragma unroll
          for (int p = 0; p < NP; ++p)
          {
            lo[p] = c[p][0];
            hi[p] = c[p][1];
          }
          double *yr = a.Y + (size_t)row * M + 2 * k * NP;
          stg_row<NP>(yr, lo);
          stg_row<NP>(yr + NP, hi);
          if (DOT)
          {
            double zl[NP], zh[NP];
            const double *xr = a.X + (size_t)row * M + 2 * k * NP;
            ldg_row_if<NP>(zl, xr, true);
            ldg_row_if<NP>(zh, xr + NP, true);
#pragma unroll
            for (int p = 0; p < NP; ++p)
            {
              dacc[p] = fma(zl[p], lo[p], dacc[p]);
              dacc[NP + p] = fma(zh[p], hi[p], dacc[NP + p]);
            }
          }
        }
      }
      __syncthreads(); // the staged rows may be overwritten by the next tile
    }

    if (DOT)
    {
      __shared__ double red[8][64];
#pragma unroll
      for (int i = 0; i < 2 * NP; ++i)
      {
        double v = dacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0)
          red[warp][2 * k * NP + i] = v;
      }
      __syncthreads();
      if (threadIdx.x < M)
      {
        double sum = 0.0;
        for (int w = 0; w < 8; ++w)
          sum += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * M + threadIdx.x] = sum;
      }
    }
  }

  struct Brb8THost
  {
    int ntiles = 0, nblocks = 0, max_u = 0;
    std::vector<int4> tile;
    std::vector<int> ucol, blkstep, blkval, blkrows;
    std::vector<unsigned short> steplc;
    std::vector<unsigned> stepmask;
    std::vector<double> val;
  };

  /** rows: processing order of the matrix rows in 8-row blocks (-1 = empty slot), nblk blocks;
   *  tilecut: [ntiles + 1] block index where each tile starts. */
  template <class Ptr, class Idx>
  void brb8t_build_host(const Ptr *rowptr, const Idx *col, const double *val, const std::vector<int> &rows,
                        const std::vector<int> &tilecut, Brb8THost &H)
  {
    const int nblk = (int)(rows.size() / 8);
    H.ntiles = (int)tilecut.size() - 1;
    H.nblocks = nblk;
    H.blkrows = rows;
    H.blkstep.assign(nblk + 1, 0);
    H.blkval.assign(nblk + 1, 0);
    struct Ent
    {
      int lc, g;
      double v;
    };
    std::vector<int> uc;
    std::vector<Ent> ent;
    std::vector<int> ul;
    for (int t = 0; t < H.ntiles; ++t)
    {
      uc.clear();
      for (int b = tilecut[t]; b < tilecut[t + 1]; ++b)
        for (int g = 0; g < 8; ++g)
        {
          const int r = rows[8 * b + g];
          if (r < 0)
            continue;
          for (long long q = rowptr[r]; q < rowptr[r + 1]; ++q)
            uc.push_back((int)col[q]);
        }
      std::sort(uc.begin(), uc.end());
      uc.erase(std::unique(uc.begin(), uc.end()), uc.end());
      const int u0 = (int)H.ucol.size();
      H.ucol.insert(H.ucol.end(), uc.begin(), uc.end());
      H.tile.push_back(make_int4(u0, (int)H.ucol.size(), tilecut[t], tilecut[t + 1]));
      H.max_u = std::max(H.max_u, (int)uc.size());
      for (int b = tilecut[t]; b < tilecut[t + 1]; ++b)
      {
        ent.clear();
        for (int g = 0; g < 8; ++g)
        {
          const int r = rows[8 * b + g];
          if (r < 0)
            continue;
          for (long long q = rowptr[r]; q < rowptr[r + 1]; ++q)
          {
            const int lc = (int)(std::lower_bound(uc.begin(), uc.end(), (int)col[q]) - uc.begin());
            ent.push_back({lc, g, val[q]});
          }
        }
        std::sort(ent.begin(), ent.end(), [](const Ent &x, const Ent &y) { return x.lc != y.lc ? x.lc < y.lc : x.g < y.g; });
        ul.clear();
        for (const Ent &e : ent)
          if (ul.empty() || ul.back() != e.lc)
            ul.push_back(e.lc);
        size_t e0 = 0;
        for (size_t a0 = 0; a0 < ul.size(); a0 += 4)
        {
          const size_t a1 = std::min(a0 + 4, ul.size());
          unsigned mask = 0;
          double slot[32];
          size_t e = e0;
          for (size_t u = a0; u < a1; ++u)
            for (; e < ent.size() && ent[e].lc == ul[u]; ++e)
            {
              const int bit = 4 * ent[e].g + (int)(u - a0);
              if (mask & (1u << bit))
                slot[bit] += ent[e].v;
              else
              {
                mask |= 1u << bit;
                slot[bit] = ent[e].v;
              }
            }
          e0 = e;
          for (int q = 0; q < 4; ++q)
            H.steplc.push_back((unsigned short)ul[std::min(a0 + q, a1 - 1)]);
          H.stepmask.push_back(mask);
          for (int bit = 0; bit < 32; ++bit)
            if (mask & (1u << bit))
              H.val.push_back(slot[bit]);
        }
        H.blkstep[b + 1] = (int)H.stepmask.size();
        H.blkval[b + 1] = (int)H.val.size();
      }
    }
  }

  /** processing order for a grid-like matrix with strides (1, S1, S2): tiles of tw x th x td points, cut into
   *  8-row blocks of bw x bh x bd points (bw*bh*bd == 8). n need not be a full box. */
  inline void brb8t_grid_order(long long n, long long S1, long long S2, int tw, int th, int td, int bw, int bh, int bd,
                               std::vector<int> &rows, std::vector<int> &tilecut)
  {
    const long long nx = S1, ny = S2 / S1, nz = (n + S2 - 1) / S2;
    rows.clear();
    tilecut.assign(1, 0);
    for (long long z0 = 0; z0 < nz; z0 += td)
      for (long long y0 = 0; y0 < ny; y0 += th)
        for (long long x0 = 0; x0 < nx; x0 += tw)
        {
          for (long long zb = z0; zb < std::min<long long>(z0 + td, nz); zb += bd)
            for (long long yb = y0; yb < std::min<long long>(y0 + th, ny); yb += bh)
              for (long long xb = x0; xb < std::min<long long>(x0 + tw, nx); xb += bw)
              {
                int cnt = 0;
                int slot[8];
                for (int dz = 0; dz < bd; ++dz)
                  for (int dy = 0; dy < bh; ++dy)
                    for (int dx = 0; dx < bw; ++dx)
                    {
                      const long long x = xb + dx, y = yb + dy, z = zb + dz;
                      const bool in = x < std::min<long long>(x0 + tw, nx) && y < std::min<long long>(y0 + th, ny) &&
                                      z < std::min<long long>(z0 + td, nz);
                      const long long r = (z * ny + y) * nx + x;
                      slot[cnt++] = (in && r < n) ? (int)r : -1;
                    }
                bool any = false;
                for (int q = 0; q < 8; ++q)
                  any = any || slot[q] >= 0;
                if (any)
                  rows.insert(rows.end(), slot, slot + 8);
              }
          if ((int)(rows.size() / 8) > tilecut.back())
            tilecut.push_back((int)(rows.size() / 8));
        }
  }

} // namespace de

namespace de
{
  // ================================================================================================
  // VD: BRB8T with a double-buffered tile pipeline (X rows + the tile's packed matrix stream in shared memory)
  // ================================================================================================
  // One persistent CTA per SM. While the warps run the tensor-core steps of tile i out of shared memory, the
  // cp.async copies of tile i+1 (its packed matrix stream, one contiguous blob, and the union of its X rows) are in
  // flight; the column ids needed to ISSUE those copies were fetched into registers one tile earlier. The compute
  // phase touches global memory only to store Y (and to read the X rows of the DOT epilogue).
  //
  // Tile blob (16-byte aligned sections, 32-bit words):
  //   [0..3]   nb (row blocks), ns (steps), nv (values), nu (union rows)
  //   blkstep[nb+1] | blkval[nb+1] | blkrows[8 nb] | pad4 | mask[ns] | pad4 | lc[4 ns] (uint16) | pad4 | val[nv] (double)
  struct TileDesc
  {
    int blob16;  // offset of the blob in 16-byte units
    int len16;   // blob length in 16-byte units
    int ucol0;   // first entry of the tile's union column list
    int nu;      // union rows
  };

  struct Brb8PArgs
  {
    int ntiles;
    long long n;
    const TileDesc *tile;
    const int4 *blob;
    const int *ucol;
    const double *X;
    const double *H;
    int n_owned;
    int ldx;       // row stride of X / H / Y in doubles (>= 8 NP; the kernel works on columns [0, 8 NP) of the view)
    double *Y;
    double *partials;
    int blob_cap16; // shared-memory capacity reserved for one blob (16-byte units)
    int xs_cap;     // ... and for the staged X rows (rows)
  };

  constexpr int kBrbThreads = 512;

  template <int NP, bool DOT, bool HALO>
  __global__ void __launch_bounds__(kBrbThreads, 1) spmm_brb8p_kernel(const Brb8PArgs a)
  {
    constexpr int M = 8 * NP;
    constexpr int LDR = M + 4;
    constexpr int CPR = M / 2;
    constexpr int NW = kBrbThreads / 32;
    constexpr int MAXC = 16; // register-prefetched column ids per thread: supports xs_cap * CPR <= 16 * 512
    extern __shared__ __align__(16) unsigned char dynp[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, k = lane & 3;
    const unsigned lt = (1u << lane) - 1u;
    const size_t buf_bytes = (size_t)a.blob_cap16 * 16 + (size_t)a.xs_cap * LDR * sizeof(double);

    double dacc[DOT ? 2 * NP : 1];
#pragma unroll
    for (int i = 0; i < (DOT ? 2 * NP : 1); ++i)
      dacc[i] = 0.0;

    auto load_desc = [&](int t)
    {
      TileDesc d;
      if (t < a.ntiles)
      {
        const int4 w = __ldg(reinterpret_cast<const int4 *>(a.tile) + t);
        d.blob16 = w.x;
        d.len16 = w.y;
        d.ucol0 = w.z;
        d.nu = w.w;
      }
      else
        d.blob16 = d.len16 = d.ucol0 = d.nu = 0;
      return d;
    };
    int pcol[MAXC];
    auto prefetch_cols = [&](const TileDesc &d)
    {
#pragma unroll
      for (int j = 0; j < MAXC; ++j)
      {
        const int e = tid + j * kBrbThreads;
        pcol[j] = (e < d.nu * CPR) ? __ldg(a.ucol + d.ucol0 + e / CPR) : 0;
      }
    };
    auto issue = [&](const TileDesc &d, int b)
    {
      unsigned char *base = dynp + b * buf_bytes;
      int4 *sb = reinterpret_cast<int4 *>(base);
      const int4 *gb = a.blob + d.blob16;
      for (int e = tid; e < d.len16; e += kBrbThreads)
        cp_async16_sparse(sb + e, gb + e);
      double *xs = reinterpret_cast<double *>(base + (size_t)a.blob_cap16 * 16);
#pragma unroll
      for (int j = 0; j < MAXC; ++j)
      {
        const int e = tid + j * kBrbThreads;
        if (e < d.nu * CPR)
        {
          const int i = e / CPR, c = e % CPR;
          const int col = pcol[j];
          const double *src = HALO ? ((col < a.n_owned) ? a.X + (size_t)col * a.ldx : a.H + (size_t)(col - a.n_owned) * a.ldx)
                                   : a.X + (size_t)col * a.ldx;
          cp_async16_sparse(xs + i * LDR + 2 * c, src + 2 * c);
        }
      }
      asm volatile("cp.async.commit_group;\n" ::);
    };

    int t = blockIdx.x;
    TileDesc cur = load_desc(t);
    TileDesc nxt = load_desc(t + gridDim.x);
    prefetch_cols(cur);
    issue(cur, 0);
    prefetch_cols(nxt);
    int b = 0;
    for (; t < a.ntiles; t += gridDim.x)
    {
      issue(nxt, b ^ 1); // (empty group past the last tile)
      const TileDesc nxt2 = load_desc(t + 2 * gridDim.x);
      prefetch_cols(nxt2);
      asm volatile("cp.async.wait_group 1;\n" ::);
      __syncthreads();

      const unsigned char *base = dynp + b * buf_bytes;
      const int *hdr = reinterpret_cast<const int *>(base);
      const int nb = hdr[0], ns = hdr[1];
      const int *blkstep = hdr + 4;
      const int *blkval = blkstep + nb + 1;
      const int *blkrows = blkval + nb + 1;
      const int o_mask = (4 + 2 * (nb + 1) + 8 * nb + 3) & ~3;
      const unsigned *smask = reinterpret_cast<const unsigned *>(hdr) + o_mask;
      const int o_lc = (o_mask + ns + 3) & ~3;
      const unsigned short *slc = reinterpret_cast<const unsigned short *>(hdr + o_lc);
      const int o_val = (o_lc + 2 * ns + 3) & ~3;
      const double *sval = reinterpret_cast<const double *>(hdr + o_val);
      const double *xs = reinterpret_cast<const double *>(base + (size_t)a.blob_cap16 * 16);

      for (int blk = warp; blk < nb; blk += NW)
      {
        const int s0 = blkstep[blk], s1 = blkstep[blk + 1];
        int voff = blkval[blk];
        double c[NP][2];
#pragma unroll
        for (int p = 0; p < NP; ++p)
          c[p][0] = c[p][1] = 0.0;
#pragma unroll 2
        for (int s = s0; s < s1; ++s)
        {
          const unsigned mask = smask[s];
          const int lc = slc[4 * s + k];
          const bool has = (mask >> lane) & 1u;
          const double av = has ? sval[voff + __popc(mask & lt)] : 0.0;
          voff += __popc(mask);
          double bv[NP];
          lds_frag<NP>(bv, xs + lc * LDR + NP * g, k);
#pragma unroll
          for (int p = 0; p < NP; ++p)
            dmma884_sp(c[p][0], c[p][1], av, bv[p]);
        }
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
          double *yr = a.Y + (size_t)row * a.ldx + 2 * k * NP;
          stg_row<NP>(yr, lo);
          stg_row<NP>(yr + NP, hi);
          if (DOT)
          {
            double zl[NP], zh[NP];
            const double *xr = a.X + (size_t)row * a.ldx + 2 * k * NP;
            ldg_row_if<NP>(zl, xr, true);
            ldg_row_if<NP>(zh, xr + NP, true);
#pragma unroll
            for (int p = 0; p < NP; ++p)
            {
              dacc[p] = fma(zl[p], lo[p], dacc[p]);
              dacc[NP + p] = fma(zh[p], hi[p], dacc[NP + p]);
            }
          }
        }
      }
      __syncthreads(); // buffer b is refilled by the issue() of the next iteration
      cur = nxt;
      nxt = nxt2;
      b ^= 1;
    }
    asm volatile("cp.async.wait_group 0;\n" ::);

    if (DOT)
    {
      __syncthreads();
      double *red = reinterpret_cast<double *>(dynp); // NW x M
#pragma unroll
      for (int i = 0; i < 2 * NP; ++i)
      {
        double v = dacc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0)
          red[warp * M + 2 * k * NP + i] = v;
      }
      __syncthreads();
      if (tid < M)
      {
        double sum = 0.0;
        for (int w = 0; w < NW; ++w)
          sum += red[w * M + tid];
        a.partials[(size_t)blockIdx.x * M + tid] = sum;
      }
    }
  }

  struct Brb8PHost
  {
    std::vector<TileDesc> tile;
    std::vector<int> blob; // 32-bit words
    int max_len16 = 0, max_u = 0;
  };

  inline void brb8p_pack_host(const Brb8THost &H, Brb8PHost &P)
  {
    P.tile.clear();
    P.blob.clear();
    P.max_len16 = 0;
    P.max_u = H.max_u;
    auto pad4 = [&]() { while (P.blob.size() % 4) P.blob.push_back(0); };
    for (int t = 0; t < H.ntiles; ++t)
    {
      const int4 ti = H.tile[t];
      const int b0 = ti.z, b1 = ti.w, nb = b1 - b0;
      const int s0 = H.blkstep[b0], s1 = H.blkstep[b1], ns = s1 - s0;
      const int v0 = H.blkval[b0], v1 = H.blkval[b1], nv = v1 - v0;
      pad4();
      const size_t w0 = P.blob.size();
      P.blob.push_back(nb);
      P.blob.push_back(ns);
      P.blob.push_back(nv);
      P.blob.push_back(ti.y - ti.x);
      for (int b = b0; b <= b1; ++b)
        P.blob.push_back(H.blkstep[b] - s0);
      for (int b = b0; b <= b1; ++b)
        P.blob.push_back(H.blkval[b] - v0);
      for (int q = 8 * b0; q < 8 * b1; ++q)
        P.blob.push_back(H.blkrows[q]);
      pad4();
      for (int s = s0; s < s1; ++s)
        P.blob.push_back((int)H.stepmask[s]);
      pad4();
      for (int s = s0; s < s1; ++s)
      {
        P.blob.push_back((int)((unsigned)H.steplc[4 * s] | ((unsigned)H.steplc[4 * s + 1] << 16)));
        P.blob.push_back((int)((unsigned)H.steplc[4 * s + 2] | ((unsigned)H.steplc[4 * s + 3] << 16)));
      }
      pad4();
      for (int q = v0; q < v1; ++q)
      {
        int w[2];
        std::memcpy(w, &H.val[q], 8);
        P.blob.push_back(w[0]);
        P.blob.push_back(w[1]);
      }
      pad4();
      TileDesc d;
      d.blob16 = (int)(w0 / 4);
      d.len16 = (int)((P.blob.size() - w0) / 4);
      d.ucol0 = ti.x;
      d.nu = ti.y - ti.x;
      P.tile.push_back(d);
      P.max_len16 = std::max(P.max_len16, d.len16);
    }
  }

} // namespace de

namespace de
{
  // ================================================================================================
  // VE: BRB8T, warp-specialised: a producer warp feeds a ring of tile buffers with TMA bulk copies
  // (cp.async.bulk -> SASS UBLKCP) that complete on mbarriers; consumer warps run the tensor-core steps.
  // ================================================================================================
  // Tile blob v2 (16-byte aligned sections, 32-bit words):
  //   [0..3]  nb, ns, nv, nu | blkstep[nb+1] | blkrows[8 nb] | pad4 | step[ns] x {lc0|lc1<<16, lc2|lc3<<16, mask, voff} | val[nv]
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

  constexpr int kBrbProducerWarps = 2;

  template <int NP, int NCW, int STAGES, bool DOT, bool HALO>
  __global__ void __launch_bounds__(32 * (NCW + kBrbProducerWarps), 1) spmm_brb8q_kernel(const Brb8PArgs a)
  {
    constexpr int NPW = kBrbProducerWarps;
    constexpr int M = 8 * NP;
    constexpr int LDR = M + 4;
    extern __shared__ __align__(128) unsigned char dynq[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t buf_bytes = (size_t)a.blob_cap16 * 16 + (size_t)a.xs_cap * LDR * sizeof(double);
    unsigned char *bufs = dynq + 128;
    const unsigned bar0 = smem_u32(dynq); // full[s] at bar0 + 8 s, empty[s] at bar0 + 8 (STAGES + s)

    if (tid == 0)
    {
      for (int s = 0; s < STAGES; ++s)
      {
        mbar_init(bar0 + 8 * s, 1 + 32 * NPW);
        mbar_init(bar0 + 8 * (STAGES + s), NCW);
      }
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();

    double dacc[DOT ? 2 * NP : 1];
#pragma unroll
    for (int i = 0; i < (DOT ? 2 * NP : 1); ++i)
      dacc[i] = 0.0;

    if (warp < NPW)
    {
      // ---------------- producers ----------------
      // The blob is one TMA bulk copy (lane 0 of warp 0). The X rows are 16-byte cp.async copies spread over the
      // 32 NPW producer threads (a bulk copy per 8M-byte row serialises in the uniform datapath: measured 3x slower);
      // each thread's copies arrive on the stage's full barrier through cp.async.mbarrier.arrive.noinc.
      // Column ids: lane l keeps ucol[32 q + l] in registers (loaded one tile ahead) and rows fetch theirs by shuffle.
      constexpr int CPR = M / 2;            // 16-byte chunks per row
      constexpr int RPR = 32 * NPW / CPR;   // rows per round of all producer threads
      constexpr int MAXQ = 16;              // supports xs_cap <= 512 rows
      static_assert(CPR <= 32, "row chunks must fit one warp");
      const int ptid = warp * 32 + lane;
      const int c = ptid % CPR, rsub = ptid / CPR;       // this thread's chunk and its row within a round
      const int rsub_w = (warp * 32) / CPR;              // first row-in-round handled by this warp
      int creg[MAXQ], nreg[MAXQ];
      int4 d = make_int4(0, 0, 0, 0), dn = make_int4(0, 0, 0, 0);
      auto load_tile_meta = [&](int t, int4 &dd, int(&rr)[MAXQ])
      {
        if (t < a.ntiles)
        {
          dd = __ldg(reinterpret_cast<const int4 *>(a.tile) + t); // {blob16, len16, ucol0, nu}
#pragma unroll
          for (int q = 0; q < MAXQ; ++q)
            rr[q] = (32 * q + lane < dd.w) ? __ldg(a.ucol + dd.z + 32 * q + lane) : 0;
        }
      };
      load_tile_meta(blockIdx.x, d, creg);
      int it = 0;
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it)
      {
        load_tile_meta(t + gridDim.x, dn, nreg); // in flight while this tile's copies are issued
        const int s = it % STAGES, use = it / STAGES;
        if (use > 0)
          mbar_wait(bar0 + 8 * (STAGES + s), (unsigned)((use - 1) & 1));
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
            constexpr int ROUNDS = 32 / RPR;
#pragma unroll
            for (int r = 0; r < ROUNDS; ++r)
            {
              const int il = r * RPR + rsub; // row within this group of 32
              const int col = __shfl_sync(0xffffffffu, creg[q], il);
              const int i = 32 * q + il;
              if (i < d.w)
              {
                const double *src = HALO ? ((col < a.n_owned) ? a.X + (size_t)col * a.ldx : a.H + (size_t)(col - a.n_owned) * a.ldx)
                                         : a.X + (size_t)col * a.ldx;
                cp_async16_sparse(xs + i * LDR + 2 * c, src + 2 * c);
              }
            }
          }
        }
        (void)rsub_w;
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(full) : "memory");
        d = dn;
#pragma unroll
        for (int q = 0; q < MAXQ; ++q)
          creg[q] = nreg[q];
      }
    }
    else
    {
      // ---------------- consumers ----------------
      const int cw = warp - NPW;
      const int g = lane >> 2, k = lane & 3;
      const unsigned lt = (1u << lane) - 1u;
      const int lsh = 16 * (k & 1);
      int it = 0;
      for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x, ++it)
      {
        const int s = it % STAGES, use = it / STAGES;
        mbar_wait(bar0 + 8 * s, (unsigned)(use & 1));
        const unsigned char *base = bufs + s * buf_bytes;
        const int *hdr = reinterpret_cast<const int *>(base);
        const int nb = hdr[0];
        const int *blkstep = hdr + 4;
        const int *blkrows = blkstep + nb + 1;
        const int o_step = (4 + (nb + 1) + 8 * nb + 3) & ~3;
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
#pragma unroll 4
          for (int st = s0; st < s1; ++st)
          {
            const int4 r = steps[st];
            const unsigned mask = (unsigned)r.z;
            const int lc = (int)(((unsigned)((k & 2) ? r.y : r.x) >> lsh) & 0xffffu);
            const bool has = (mask >> lane) & 1u;
            const double av = has ? sval[r.w + __popc(mask & lt)] : 0.0;
            double bv[NP];
            lds_frag<NP>(bv, xs + lc * LDR + NP * g, k);
#pragma unroll
            for (int p = 0; p < NP; ++p)
              dmma884_sp(c[p][0], c[p][1], av, bv[p]);
          }
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
            double *yr = a.Y + (size_t)row * a.ldx + 2 * k * NP;
            stg_row<NP>(yr, lo);
            stg_row<NP>(yr + NP, hi);
            if (DOT)
            {
              double zl[NP], zh[NP];
              const double *xr = a.X + (size_t)row * a.ldx + 2 * k * NP;
              ldg_row_if<NP>(zl, xr, true);
              ldg_row_if<NP>(zh, xr + NP, true);
#pragma unroll
              for (int p = 0; p < NP; ++p)
              {
                dacc[p] = fma(zl[p], lo[p], dacc[p]);
                dacc[NP + p] = fma(zh[p], hi[p], dacc[NP + p]);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0)
          mbar_arrive(bar0 + 8 * (STAGES + s));
      }
    }

    if (DOT)
    {
      __syncthreads(); // every tile has been consumed: the buffers are free
      double *red = reinterpret_cast<double *>(bufs); // NCW x M
      if (warp >= NPW)
      {
        const int g = lane >> 2, k = lane & 3;
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
      __syncthreads();
      if (tid < M)
      {
        double sum = 0.0;
        for (int w = 0; w < NCW; ++w)
          sum += red[w * M + tid];
        a.partials[(size_t)blockIdx.x * M + tid] = sum;
      }
    }
  }

  inline void brb8q_pack_host(const Brb8THost &H, Brb8PHost &P)
  {
    P.tile.clear();
    P.blob.clear();
    P.max_len16 = 0;
    P.max_u = H.max_u;
    auto pad4 = [&]() { while (P.blob.size() % 4) P.blob.push_back(0); };
    for (int t = 0; t < H.ntiles; ++t)
    {
      const int4 ti = H.tile[t];
      const int b0 = ti.z, b1 = ti.w, nb = b1 - b0;
      const int s0 = H.blkstep[b0], s1 = H.blkstep[b1], ns = s1 - s0;
      const int v0 = H.blkval[b0], v1 = H.blkval[b1], nv = v1 - v0;
      pad4();
      const size_t w0 = P.blob.size();
      P.blob.push_back(nb);
      P.blob.push_back(ns);
      P.blob.push_back(nv);
      P.blob.push_back(ti.y - ti.x);
      for (int b = b0; b <= b1; ++b)
        P.blob.push_back(H.blkstep[b] - s0);
      for (int q = 8 * b0; q < 8 * b1; ++q)
        P.blob.push_back(H.blkrows[q]);
      pad4();
      int voff = 0;
      for (int s = s0; s < s1; ++s)
      {
        P.blob.push_back((int)((unsigned)H.steplc[4 * s] | ((unsigned)H.steplc[4 * s + 1] << 16)));
        P.blob.push_back((int)((unsigned)H.steplc[4 * s + 2] | ((unsigned)H.steplc[4 * s + 3] << 16)));
        P.blob.push_back((int)H.stepmask[s]);
        P.blob.push_back(voff);
        voff += __builtin_popcount(H.stepmask[s]);
      }
      for (int q = v0; q < v1; ++q)
      {
        int w[2];
        std::memcpy(w, &H.val[q], 8);
        P.blob.push_back(w[0]);
        P.blob.push_back(w[1]);
      }
      pad4();
      TileDesc d;
      d.blob16 = (int)(w0 / 4);
      d.len16 = (int)((P.blob.size() - w0) / 4);
      d.ucol0 = ti.x;
      d.nu = ti.y - ti.x;
      P.tile.push_back(d);
      P.max_len16 = std::max(P.max_len16, d.len16);
    }
  }

} // namespace de
