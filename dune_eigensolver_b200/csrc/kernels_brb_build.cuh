// Device-side construction of the BRB SpMM format (layout: brb_format.hpp) from the CSR arrays already on the GPU.
//
// Setup-time kernels, one CTA per tile, run twice: a COUNT pass sizes every tile blob (steps, values, union rows), the
// host turns the sizes into offsets (one small copy each way), and a FILL pass writes the blobs and union lists. The
// rows of every tile (`rows`, `tilecut`) come from the host (brb::grid_order / brb::linear_order: O(n) index
// arithmetic); everything that touches the nonzeros happens here, at HBM speed, instead of ~45 ns per nonzero on a
// host core. Produces bit-identical arrays to the host builder (brb::build_tiles) for matrices without duplicate
// entries; duplicates are accumulated with atomicAdd (their order is then not fixed).
//
// Per tile: (1) the union of the column ids of its rows is collected in a shared-memory hash set and sorted
// (bitonic, <= 512 keys); a column's tile-local id is its rank, found by binary search. (2) Per row block (one warp):
// the block's local ids are marked in a 512-bit map; the rank of a set bit is its (step, column-in-step); a second
// sweep over the entries builds the 32-bit pattern mask of every step; prefix sums of the mask populations place the
// values; a third sweep scatters the values.
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

namespace de
{

  constexpr int kBldThreads = 256;
  constexpr int kBldWarps = kBldThreads / 32;
  constexpr int kBldHash = 2048;        // hash-set slots (union <= 512)
  constexpr int kBldMaxUnion = 512;     // == brb::kMaxUnion
  constexpr int kBldMaxBlocks = 64;     // row blocks per tile
  constexpr int kBldMaxSteps = kBldMaxUnion / 4;

  struct BrbBuildArgs
  {
    int ntiles;
    const int *tilecut; // [ntiles + 1] first row block of each tile
    const int *rows;    // [8 * nblocks] matrix row of each block slot (-1: empty)
    const int *rowptr;
    const int *col;
    const double *val;
    int n_owned;
    int4 *info;         // COUNT: {steps, values, union rows, flags: bit 0 halo columns, bit 1 not representable}
    const int4 *place;  // FILL: {blob16, len16, ucol0, nu} of tile t
    int *blob;          // FILL: zero-initialised by the host
    int *ucol;
  };

  struct BrbBuildShared
  {
    int keys[kBldHash];
    int ulist[kBldMaxUnion];
    unsigned bitmap[kBldWarps][kBldMaxUnion / 32];
    unsigned masks[kBldWarps][kBldMaxSteps];
    int prefix[kBldWarps][kBldMaxUnion / 32 + 1]; // set bits of the block's bitmap before word w
    int voff[kBldWarps][kBldMaxSteps];
    int blk_ns[kBldMaxBlocks + 1];
    int blk_nv[kBldMaxBlocks + 1];
    int nu;
    int fail;
  };

  __device__ __forceinline__ int brb_lid(const int *ulist, int nu, int c)
  {
    int lo = 0, hi = nu - 1;
    while (lo < hi)
    {
      const int mid = (lo + hi) >> 1;
      if (ulist[mid] < c)
        lo = mid + 1;
      else
        hi = mid;
    }
    return lo;
  }

  /** marks the block's local ids, builds the step masks; returns (steps, values) of the block. Warp-collective. */
  __device__ __forceinline__ void brb_block_masks(const BrbBuildArgs &a, BrbBuildShared &S, int warp, int lane, const int *brows,
                                                  int &nsb, int &nvb)
  {
    unsigned *bm = S.bitmap[warp];
    unsigned *mk = S.masks[warp];
    int *prefix = S.prefix[warp];
    if (lane < kBldMaxUnion / 32)
      bm[lane] = 0u;
    __syncwarp();
    for (int g = 0; g < 8; ++g)
    {
      const int r = brows[g];
      if (r < 0)
        continue;
      for (int p = a.rowptr[r] + lane; p < a.rowptr[r + 1]; p += 32)
      {
        const int l = brb_lid(S.ulist, S.nu, a.col[p]);
        atomicOr(&bm[l >> 5], 1u << (l & 31));
      }
    }
    __syncwarp();
    if (lane <= kBldMaxUnion / 32)
    {
      int sum = 0;
      for (int w = 0; w < lane; ++w)
        sum += __popc(bm[w]);
      prefix[lane] = sum;
    }
    __syncwarp();
    const int cnt = prefix[kBldMaxUnion / 32];
    nsb = (cnt + 3) >> 2;
    for (int s = lane; s < nsb; s += 32)
      mk[s] = 0u;
    __syncwarp();
    for (int g = 0; g < 8; ++g)
    {
      const int r = brows[g];
      if (r < 0)
        continue;
      for (int p = a.rowptr[r] + lane; p < a.rowptr[r + 1]; p += 32)
      {
        const int l = brb_lid(S.ulist, S.nu, a.col[p]);
        const int rank = prefix[l >> 5] + __popc(bm[l >> 5] & ((1u << (l & 31)) - 1u));
        atomicOr(&mk[rank >> 2], 1u << (4 * g + (rank & 3)));
      }
    }
    __syncwarp();
    int nv = 0;
    for (int s = lane; s < nsb; s += 32)
      nv += __popc(mk[s]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      nv += __shfl_xor_sync(0xffffffffu, nv, o);
    nvb = nv;
  }

  template <bool FILL>
  __global__ void __launch_bounds__(kBldThreads) brb_build_kernel(const BrbBuildArgs a)
  {
    __shared__ BrbBuildShared S;
    const int t = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = a.tilecut[t], nb = a.tilecut[t + 1] - b0;
    const int *trows = a.rows + 8 * (size_t)b0;

    for (int i = tid; i < kBldHash; i += kBldThreads)
      S.keys[i] = -1;
    for (int i = tid; i < kBldMaxUnion; i += kBldThreads)
      S.ulist[i] = 0x7fffffff;
    if (tid == 0)
    {
      S.nu = 0;
      S.fail = nb > kBldMaxBlocks ? 1 : 0;
    }
    __syncthreads();

    // ---- (1) union of the tile's columns: hash set, then sort ----
    for (int q = warp; q < 8 * nb && q < 8 * kBldMaxBlocks; q += kBldWarps)
    {
      const int r = trows[q];
      if (r < 0)
        continue;
      for (int p = a.rowptr[r] + lane; p < a.rowptr[r + 1]; p += 32)
      {
        const int c = a.col[p];
        unsigned h = ((unsigned)c * 2654435761u) >> 21; // 11 bits
        while (true)
        {
          const int old = atomicCAS(&S.keys[h], -1, c);
          if (old == -1)
          {
            const int idx = atomicAdd(&S.nu, 1);
            if (idx < kBldMaxUnion)
              S.ulist[idx] = c;
            else
              S.fail = 1;
            break;
          }
          if (old == c)
            break;
          h = (h + 1) & (kBldHash - 1);
          if (*reinterpret_cast<volatile int *>(&S.fail)) // the table only fills up when the union is far beyond the limit
            break;
        }
      }
    }
    __syncthreads();
    if (S.fail)
    {
      if (!FILL && tid == 0)
        a.info[t] = make_int4(0, 0, 0, 2);
      return;
    }
    const int nu = S.nu;
    // bitonic sort of the 512 (padded) keys, two elements per thread
    for (int k = 2; k <= kBldMaxUnion; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1)
      {
        for (int i = tid; i < kBldMaxUnion; i += kBldThreads)
        {
          const int ixj = i ^ j;
          if (ixj > i)
          {
            const int x = S.ulist[i], y = S.ulist[ixj];
            const bool up = (i & k) == 0;
            if ((x > y) == up)
            {
              S.ulist[i] = y;
              S.ulist[ixj] = x;
            }
          }
        }
        __syncthreads();
      }

    // ---- (2) sizes of every row block ----
    for (int b = warp; b < nb; b += kBldWarps)
    {
      int nsb, nvb;
      brb_block_masks(a, S, warp, lane, trows + 8 * b, nsb, nvb);
      if (lane == 0)
      {
        S.blk_ns[b] = nsb;
        S.blk_nv[b] = nvb;
      }
    }
    __syncthreads();
    if (tid == 0)
    {
      int ns = 0, nv = 0;
      for (int b = 0; b < nb; ++b)
      {
        const int x = S.blk_ns[b], y = S.blk_nv[b];
        S.blk_ns[b] = ns;
        S.blk_nv[b] = nv;
        ns += x;
        nv += y;
      }
      S.blk_ns[nb] = ns;
      S.blk_nv[nb] = nv;
      if (!FILL)
        a.info[t] = make_int4(ns, nv, nu, (nu > 0 && S.ulist[nu - 1] >= a.n_owned) ? 1 : 0);
    }
    if (!FILL)
      return;
    __syncthreads();

    // ---- (3) write the blob ----
    const int4 pl = a.place[t];
    int *hdr = a.blob + 4 * (size_t)pl.x;
    const int ns = S.blk_ns[nb], nv = S.blk_nv[nb];
    const int o_step = (4 + (nb + 1) + 8 * nb + 4 * nb + 3) & ~3;
    int *steps = hdr + o_step;
    double *vals = reinterpret_cast<double *>(hdr + o_step + 4 * ns);
    if (tid == 0)
    {
      hdr[0] = nb;
      hdr[1] = ns;
      hdr[2] = nv;
      hdr[3] = nu;
    }
    for (int b = tid; b <= nb; b += kBldThreads)
      hdr[4 + b] = S.blk_ns[b];
    for (int q = tid; q < 8 * nb; q += kBldThreads)
      hdr[4 + (nb + 1) + q] = trows[q];
    for (int q2 = tid; q2 < 4 * nb; q2 += kBldThreads)
    {
      // tile-local id of each row's own column (0xffff: not among the tile's columns), two per word
      unsigned w = 0;
      for (int h = 0; h < 2; ++h)
      {
        const int r = trows[2 * q2 + h];
        unsigned self = 0xffffu;
        if (r >= 0 && nu > 0)
        {
          const int l = brb_lid(S.ulist, nu, r);
          if (S.ulist[l] == r)
            self = (unsigned)l;
        }
        w |= self << (16 * h);
      }
      hdr[4 + (nb + 1) + 8 * nb + q2] = (int)w;
    }
    for (int i = tid; i < nu; i += kBldThreads)
      a.ucol[pl.z + i] = S.ulist[i];

    for (int b = warp; b < nb; b += kBldWarps)
    {
      int nsb, nvb;
      const int *brows = trows + 8 * b;
      brb_block_masks(a, S, warp, lane, brows, nsb, nvb);
      const int *prefix = S.prefix[warp];
      const unsigned *bm = S.bitmap[warp];
      const unsigned *mk = S.masks[warp];
      int *vo = S.voff[warp];
      // first value of every step (tile-relative): running sum of the mask populations
      if (lane == 0)
      {
        int run = S.blk_nv[b];
        for (int s = 0; s < nsb; ++s)
        {
          vo[s] = run;
          run += __popc(mk[s]);
        }
      }
      __syncwarp();
      const int cnt = prefix[kBldMaxUnion / 32];
      const int sbase = S.blk_ns[b];
      for (int s = lane; s < nsb; s += 32)
      {
        unsigned lc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
          int rank = 4 * s + j;
          if (rank >= cnt)
            rank = cnt - 1; // short last step: repeat the last valid column
          int w = 0;
          while (prefix[w + 1] <= rank)
            ++w;
          lc[j] = 32u * w + __fns(bm[w], 0u, rank - prefix[w] + 1);
        }
        int4 rec;
        rec.x = (int)(lc[0] | (lc[1] << 16));
        rec.y = (int)(lc[2] | (lc[3] << 16));
        rec.z = (int)mk[s];
        rec.w = vo[s];
        reinterpret_cast<int4 *>(steps)[sbase + s] = rec;
      }
      for (int g = 0; g < 8; ++g)
      {
        const int r = brows[g];
        if (r < 0)
          continue;
        for (int p = a.rowptr[r] + lane; p < a.rowptr[r + 1]; p += 32)
        {
          const int l = brb_lid(S.ulist, nu, a.col[p]);
          const int rank = prefix[l >> 5] + __popc(bm[l >> 5] & ((1u << (l & 31)) - 1u));
          const int s = rank >> 2, bit = 4 * g + (rank & 3);
          const int pos = vo[s] + __popc(mk[s] & ((1u << bit) - 1u));
          atomicAdd(&vals[pos], a.val[p]); // the blob is zero-initialised; duplicates of a CSR row accumulate
        }
      }
      __syncwarp();
    }
  }

} // namespace de
