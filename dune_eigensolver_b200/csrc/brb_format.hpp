// Host-side construction of the BRB ("blocked row-block") SpMM format consumed by spmm_brb_kernel
// (kernels_spmm_blocked.cuh). Plain C++ (no CUDA), setup-time only, multi-threaded over tiles.
//
//   row block : 8 matrix rows handled by one warp; the sorted union of their columns is cut into STEPS of 4
//               columns; a step is one 8 x 4 slice of the sparse matrix = one FP64 tensor-core operand.
//   tile      : a group of row blocks whose X rows (the union of all their columns) are staged once in shared
//               memory. A tile is stored as one contiguous, 16-byte aligned BLOB of 32-bit words
//                   [0..3]  nb (row blocks), ns (steps), nv (values), nu (union rows)
//                   blkstep[nb+1]            first step of each row block (tile-relative)
//                   blkrows[8 nb]            output row of each block row (-1: empty slot)
//                   blkself[8 nb] (uint16)   tile-local id of the column equal to that row (its own X row is then in the staged
//                                            tile: the dot-product epilogue reads it from shared memory), 0xffff if absent
//                   pad to 4 words
//                   step[ns] x 4 words       lc0 | lc1 << 16, lc2 | lc3 << 16, pattern mask, first value (tile-relative)
//                   val[nv] doubles          packed values: step-major, ascending pattern bit (bit = 4 * block row + step column)
//               plus its list of union column ids (`ucol`, ascending; lc = position in that list).
//
// Which rows form a tile decides how often an X row crosses the L2 -> SM fabric (measured ceiling ~10 TB/s on B200,
// tools/micro/gather_probe.cu, only 1.5x HBM). For matrices whose pattern is a set of diagonals with strides
// (1, S1, S2) -- structured-grid discretisations in lexicographic order, every configuration of BASELINE.json --
// tiles are tw x th x td boxes of grid points and row blocks are 2 x 2 x 2 (or 4 x 2 x 1) sub-boxes: a 27-point
// tile of 6 x 4 x 4 points needs 288 X rows for 96 matrix rows instead of 27 per row. Anything else falls back to
// consecutive rows (tiles cut when the union outgrows the shared-memory budget), which is what a banded / RCM-ordered
// matrix wants anyway; a matrix whose 8-row blocks do not fit at all reports valid = false and the CSR kernel is used.
#pragma once

#include <algorithm>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace de
{
  namespace brb
  {

    constexpr int kMaxUnion = 512;                          // union rows per tile the kernel can stage (16 x 32)
    constexpr int kStageBudget = (227 * 1024 - 256) / 2;    // bytes of one pipeline stage when two stages must fit
    constexpr int kRowBytesWide = (32 + 4) * 8;             // staged X row of the widest pass (32 columns + padding)

    struct TileDesc
    {
      int blob16; // offset of the blob in 16-byte units
      int len16;  // blob length in 16-byte units
      int ucol0;  // first entry of the tile's union list
      int nu;     // union rows
    };

    struct Order
    {
      std::vector<int> rows;    // processing order in 8-row blocks, -1 = empty slot
      std::vector<int> tilecut; // [ntiles + 1] first block of each tile
    };

    struct Format
    {
      bool valid = false;
      bool grid = false;
      long long S1 = 0, S2 = 0;
      int tw = 0, th = 0, td = 0;
      int ntiles = 0, n_interior = 0; // tiles [0, n_interior) reference owned columns only
      int max_len16 = 0, max_u = 0;
      long long nblocks = 0, nsteps = 0, nvals = 0;
      std::vector<TileDesc> tile;
      std::vector<int> blob;
      std::vector<int> ucol;
    };

    /** per-thread scratch of the tile builder */
    template <class Ptr, class Idx>
    struct TileBuilder
    {
      const Ptr *rowptr;
      const Idx *col;
      const double *val;
      long long n_owned;
      std::vector<int> stamp;
      std::vector<unsigned short> lid;
      std::vector<int> ulist;
      std::vector<double> tvals;
      int tag = 0;

      TileBuilder(long long ncols, const Ptr *rp, const Idx *c, const double *v, long long owned)
          : rowptr(rp), col(c), val(v), n_owned(owned), stamp((size_t)ncols, -1), lid((size_t)ncols, 0)
      {
      }

      /** appends one tile (blocks b0..b1 of `rows`) to blob / ucol; returns false if it cannot be represented */
      bool build(const int *rows, int nb, std::vector<int> &blob, std::vector<int> &ucol, TileDesc &d, bool &has_halo,
                 long long &nsteps, long long &nvals)
      {
        ++tag;
        ulist.clear();
        for (int q = 0; q < 8 * nb; ++q)
        {
          const int r = rows[q];
          if (r < 0)
            continue;
          for (Ptr p = rowptr[r]; p < rowptr[r + 1]; ++p)
          {
            const int c = (int)col[p];
            if (stamp[c] != tag)
            {
              stamp[c] = tag;
              ulist.push_back(c);
            }
          }
        }
        const int nu = (int)ulist.size();
        if (nu > kMaxUnion)
          return false;
        std::sort(ulist.begin(), ulist.end());
        for (int i = 0; i < nu; ++i)
          lid[ulist[i]] = (unsigned short)i;
        has_halo = nu > 0 && ulist.back() >= n_owned;

        while (blob.size() % 4)
          blob.push_back(0);
        const size_t w0 = blob.size();
        blob.resize(w0 + 4 + (nb + 1) + 8 * nb + 4 * nb, 0);
        for (int q = 0; q < 8 * nb; ++q)
          blob[w0 + 4 + (nb + 1) + q] = rows[q];
        for (int q = 0; q < 8 * nb; ++q)
        {
          const int r = rows[q];
          const unsigned self = (r >= 0 && (size_t)r < stamp.size() && stamp[r] == tag) ? (unsigned)lid[r] : 0xffffu;
          int &w = blob[w0 + 4 + (nb + 1) + 8 * nb + q / 2];
          w = (int)((unsigned)w | (self << (16 * (q & 1))));
        }
        while (blob.size() % 4)
          blob.push_back(0);

        tvals.clear();
        int ns = 0;
        for (int b = 0; b < nb; ++b)
        {
          blob[w0 + 4 + b] = ns;
          uint64_t bits[kMaxUnion / 64] = {0};
          for (int g = 0; g < 8; ++g)
          {
            const int r = rows[8 * b + g];
            if (r < 0)
              continue;
            for (Ptr p = rowptr[r]; p < rowptr[r + 1]; ++p)
            {
              const int l = lid[(int)col[p]];
              bits[l >> 6] |= 1ull << (l & 63);
            }
          }
          int prefix[kMaxUnion / 64 + 1];
          prefix[0] = 0;
          for (int w = 0; w < kMaxUnion / 64; ++w)
            prefix[w + 1] = prefix[w] + __builtin_popcountll(bits[w]);
          const int cnt = prefix[kMaxUnion / 64];
          const int nsb = (cnt + 3) / 4;
          // step records (lc filled below), masks and value slots of this block
          const size_t s0 = blob.size();
          blob.resize(s0 + 4 * (size_t)nsb, 0);
          {
            int rank = 0;
            unsigned short lcs[4] = {0, 0, 0, 0};
            auto flush = [&](int st)
            {
              blob[s0 + 4 * st] = (int)((unsigned)lcs[0] | ((unsigned)lcs[1] << 16));
              blob[s0 + 4 * st + 1] = (int)((unsigned)lcs[2] | ((unsigned)lcs[3] << 16));
            };
            for (int w = 0; w < kMaxUnion / 64; ++w)
            {
              uint64_t x = bits[w];
              while (x)
              {
                const int l = 64 * w + __builtin_ctzll(x);
                x &= x - 1;
                const int k = rank & 3;
                lcs[k] = (unsigned short)l;
                if (k == 3)
                  flush(rank >> 2);
                ++rank;
              }
            }
            if (rank & 3)
            {
              // short last step: repeat the last valid column (its pattern bits stay clear)
              for (int k = rank & 3; k < 4; ++k)
                lcs[k] = lcs[(rank & 3) - 1];
              flush(rank >> 2);
            }
          }
          slots.assign((size_t)nsb * 32, 0.0);
          for (int g = 0; g < 8; ++g)
          {
            const int r = rows[8 * b + g];
            if (r < 0)
              continue;
            for (Ptr p = rowptr[r]; p < rowptr[r + 1]; ++p)
            {
              const int l = lid[(int)col[p]];
              const int rank = prefix[l >> 6] + __builtin_popcountll(bits[l >> 6] & ((1ull << (l & 63)) - 1ull));
              const int st = rank >> 2, bit = 4 * g + (rank & 3);
              blob[s0 + 4 * st + 2] |= (int)(1u << bit);
              slots[(size_t)st * 32 + bit] += val[p]; // duplicate entries of a CSR row accumulate
            }
          }
          for (int st = 0; st < nsb; ++st)
          {
            const unsigned mask = (unsigned)blob[s0 + 4 * st + 2];
            blob[s0 + 4 * st + 3] = (int)tvals.size();
            for (int bit = 0; bit < 32; ++bit)
              if (mask & (1u << bit))
                tvals.push_back(slots[(size_t)st * 32 + bit]);
          }
          ns += nsb;
        }
        blob[w0 + 4 + nb] = ns;
        const int nv = (int)tvals.size();
        blob[w0] = nb;
        blob[w0 + 1] = ns;
        blob[w0 + 2] = nv;
        blob[w0 + 3] = nu;
        const size_t v0 = blob.size();
        blob.resize(v0 + 2 * (size_t)nv);
        if (nv > 0)
          std::memcpy(&blob[v0], tvals.data(), sizeof(double) * nv);
        while (blob.size() % 4)
          blob.push_back(0);
        d.blob16 = (int)(w0 / 4);
        d.len16 = (int)((blob.size() - w0) / 4);
        d.ucol0 = (int)ucol.size();
        d.nu = nu;
        ucol.insert(ucol.end(), ulist.begin(), ulist.end());
        nsteps += ns;
        nvals += nv;
        return true;
      }

    private:
      std::vector<double> slots;
    };

    inline size_t stage_bytes(int len16, int nu, int row_bytes) { return (size_t)len16 * 16 + (size_t)nu * row_bytes; }

    /** strides of a multi-diagonal pattern, read off the offsets (column - row) of a sample row.
     *  Returns false if the row has no off-diagonal structure to exploit. S2 = 0 means two-dimensional. */
    template <class Ptr, class Idx>
    bool detect_grid(long long n, const Ptr *rowptr, const Idx *col, long long n_owned, long long &S1, long long &S2)
    {
      S1 = S2 = 0;
      if (n < 64)
        return false;
      // the longest of a set of sample rows is taken to be an interior point of the grid
      long long best = -1;
      long long bestlen = 0;
      for (int q = 0; q < 509; ++q)
      {
        const long long cand = std::min<long long>(n - 1, (n * q) / 509 + (q % 7));
        const long long len = (long long)(rowptr[cand + 1] - rowptr[cand]);
        if (len > bestlen)
        {
          bestlen = len;
          best = cand;
        }
      }
      if (best < 0 || bestlen < 3 || bestlen > 256)
        return false;
      std::vector<long long> off; // |column - row| of the model row (patterns are structurally symmetric)
      for (Ptr p = rowptr[best]; p < rowptr[best + 1]; ++p)
      {
        const long long c = (long long)col[p];
        if (c != best && c < n_owned)
          off.push_back(c > best ? c - best : best - c);
      }
      std::sort(off.begin(), off.end());
      off.erase(std::unique(off.begin(), off.end()), off.end());
      // the pattern must be (nearly) the same set of diagonals everywhere: offsets of evenly spaced sample rows
      // have to come from the model row's offsets (boundary rows simply miss some)
      {
        long long total = 0, hit = 0;
        for (int q = 0; q < 64; ++q)
        {
          const long long r = (n - 1) * q / 63;
          for (Ptr p = rowptr[r]; p < rowptr[r + 1]; ++p)
          {
            const long long c = (long long)col[p];
            if (c >= n_owned || c == r)
              continue;
            ++total;
            hit += std::binary_search(off.begin(), off.end(), c > r ? c - r : r - c);
          }
        }
        if (total == 0 || hit * 10 < total * 9)
          return false;
      }
      // runs of consecutive offsets -> centres
      std::vector<long long> centre;
      for (size_t i = 0; i < off.size();)
      {
        size_t j = i;
        while (j + 1 < off.size() && off[j + 1] == off[j] + 1)
          ++j;
        centre.push_back((off[i] + off[j]) / 2);
        i = j + 1;
      }
      // drop the run that contains the x neighbours (starts at 1)
      if (!centre.empty() && !off.empty() && off[0] == 1)
        centre.erase(centre.begin());
      if (centre.empty())
        return false;
      S1 = centre[0];
      if (S1 < 2)
        return false;
      if (centre.size() >= 2)
      {
        auto has = [&](long long c) { return std::find(centre.begin(), centre.end(), c) != centre.end(); };
        long long pick = 0;
        for (size_t i = 1; i < centre.size() && !pick; ++i)
          if (has(centre[i] - S1) && has(centre[i] + S1))
            pick = centre[i];
        if (!pick)
          pick = centre[1];
        if (pick % S1 == 0 && pick >= 2 * S1)
          S2 = pick;
      }
      return true;
    }

    constexpr long long kPlanePointsInL2 = 16384; // grid planes larger than this are swept in y chunks (grid_order)
    inline long long &plane_points_setting()
    {
      static long long v = kPlanePointsInL2;
      return v;
    }

    /** tiles of tw x th x td points cut into row blocks of bw x bh x bd points (bw bh bd == 8) */
    inline void grid_order(long long n, long long S1, long long S2, int tw, int th, int td, int bw, int bh, int bd, Order &o)
    {
      const long long nx = S1;
      const long long ny = S2 > 0 ? S2 / S1 : (n + S1 - 1) / S1;
      const long long nz = S2 > 0 ? (n + S2 - 1) / S2 : 1;
      o.rows.clear();
      o.tilecut.assign(1, 0);
      // Large planes: the tiles of one sweep over (x, y) at fixed z0 touch td + 2 planes of X, and the two halo planes are
      // needed again one sweep later -- after td planes' worth of X rows, Y rows and matrix stream (~3 KB per grid point of
      // the plane at m = 32) went through the L2. Beyond ~16 k points per plane that exceeds what the L2 keeps (measured:
      // SpMM at 0.74 of HBM peak on 100^3, 0.70 on 200^3, 0.60 on 256^3), so y is cut into chunks and the sweep over z runs
      // inside a chunk: the rows between two uses of a halo plane shrink to nx * ychunk * td.
      long long ychunk = ny;
      const long long limit = plane_points_setting(); // (de_context_set_option "brb_plane_points"; tools/ychunk_probe.py)
      if (nz > 1 && nx * ny > limit)
      {
        const long long want = std::max<long long>(th, limit * 3 / 4 / nx);
        const long long nchunks = (ny + want - 1) / want;
        ychunk = ((ny + nchunks - 1) / nchunks + th - 1) / th * th;
      }
      for (long long yc = 0; yc < ny; yc += ychunk)
      for (long long z0 = 0; z0 < nz; z0 += td)
        for (long long y0 = yc; y0 < std::min<long long>(yc + ychunk, ny); y0 += th)
          for (long long x0 = 0; x0 < nx; x0 += tw)
          {
            const long long x1 = std::min<long long>(x0 + tw, nx), y1 = std::min<long long>(y0 + th, ny),
                            z1 = std::min<long long>(z0 + td, nz);
            for (long long zb = z0; zb < z1; zb += bd)
              for (long long yb = y0; yb < y1; yb += bh)
                for (long long xb = x0; xb < x1; xb += bw)
                {
                  int slot[8];
                  int cnt = 0;
                  bool any = false;
                  for (int dz = 0; dz < bd; ++dz)
                    for (int dy = 0; dy < bh; ++dy)
                      for (int dx = 0; dx < bw; ++dx)
                      {
                        const long long x = xb + dx, y = yb + dy, z = zb + dz;
                        const long long r = (z * ny + y) * nx + x;
                        const bool in = x < x1 && y < y1 && z < z1 && r < n;
                        slot[cnt++] = in ? (int)r : -1;
                        any = any || in;
                      }
                  if (any)
                    o.rows.insert(o.rows.end(), slot, slot + 8);
                }
            if ((int)(o.rows.size() / 8) > o.tilecut.back())
              o.tilecut.push_back((int)(o.rows.size() / 8));
          }
    }

    /** consecutive rows; a tile ends when its union or its blob would outgrow the stage budget */
    template <class Ptr, class Idx>
    bool linear_order(long long n, long long ncols, const Ptr *rowptr, const Idx *col, Order &o)
    {
      const long long nb = (n + 7) / 8;
      o.rows.assign((size_t)nb * 8, -1);
      for (long long r = 0; r < n; ++r)
        o.rows[(size_t)r] = (int)r;
      o.tilecut.assign(1, 0);
      std::vector<int> stamp((size_t)ncols, -1), bstamp((size_t)ncols, -1);
      std::vector<int> bc;
      int tag = 0;
      long long tile_u = 0, tile_nnz = 0, tile_blocks = 0;
      constexpr int max_blocks = 32;
      auto bytes = [](long long u, long long nnz, long long blocks)
      { return (size_t)u * kRowBytesWide + (size_t)nnz * 12 + 64 * (size_t)blocks + 128; };
      for (long long b = 0; b < nb; ++b)
      {
        bc.clear();
        long long add_nnz = 0;
        for (int g = 0; g < 8; ++g)
        {
          const long long r = 8 * b + g;
          if (r >= n)
            break;
          for (Ptr p = rowptr[r]; p < rowptr[r + 1]; ++p)
          {
            const int c = (int)col[p];
            ++add_nnz;
            if (bstamp[c] != (int)b)
            {
              bstamp[c] = (int)b;
              bc.push_back(c);
            }
          }
        }
        long long add_u = 0;
        for (int c : bc)
          add_u += stamp[c] != tag;
        if (tile_blocks > 0 && (tile_u + add_u > 320 || bytes(tile_u + add_u, tile_nnz + add_nnz, tile_blocks + 1) > (size_t)kStageBudget ||
                                tile_blocks >= max_blocks))
        {
          o.tilecut.push_back((int)b);
          ++tag;
          tile_u = tile_nnz = tile_blocks = 0;
          add_u = (long long)bc.size();
        }
        for (int c : bc)
          stamp[c] = tag;
        tile_u += add_u;
        tile_nnz += add_nnz;
        ++tile_blocks;
        if (tile_blocks == 1 && (tile_u > kMaxUnion || bytes(tile_u, tile_nnz, 1) > (size_t)kStageBudget))
          return false; // a single row block does not fit
      }
      o.tilecut.push_back((int)nb);
      return true;
    }

    /** build every tile of `o` (in parallel) and order the tiles interior-first */
    template <class Ptr, class Idx>
    bool build_tiles(long long ncols, const Ptr *rowptr, const Idx *col, const double *val, long long n_owned, const Order &o,
                     Format &F, int nthreads = 0)
    {
      const int ntiles = (int)o.tilecut.size() - 1;
      if (nthreads <= 0)
      {
        nthreads = (int)std::thread::hardware_concurrency();
        if (nthreads <= 0)
          nthreads = 4;
        nthreads = std::min(nthreads, 16);
      }
      nthreads = std::max(1, std::min(nthreads, (ntiles + 63) / 64));
      struct Part
      {
        std::vector<int> blob, ucol;
        std::vector<TileDesc> desc;
        std::vector<char> halo;
        long long nsteps = 0, nvals = 0;
        bool ok = true;
      };
      std::vector<Part> part(nthreads);
      auto work = [&](int tix)
      {
        Part &P = part[tix];
        const int t0 = (int)((long long)ntiles * tix / nthreads), t1 = (int)((long long)ntiles * (tix + 1) / nthreads);
        TileBuilder<Ptr, Idx> tb(ncols, rowptr, col, val, n_owned);
        for (int t = t0; t < t1 && P.ok; ++t)
        {
          TileDesc d;
          bool hh = false;
          const int b0 = o.tilecut[t], b1 = o.tilecut[t + 1];
          if (!tb.build(o.rows.data() + 8 * (size_t)b0, b1 - b0, P.blob, P.ucol, d, hh, P.nsteps, P.nvals))
            P.ok = false;
          P.desc.push_back(d);
          P.halo.push_back(hh ? 1 : 0);
        }
      };
      if (nthreads == 1)
        work(0);
      else
      {
        std::vector<std::thread> th;
        for (int i = 0; i < nthreads; ++i)
          th.emplace_back(work, i);
        for (auto &t : th)
          t.join();
      }
      size_t blob_words = 0, ucols = 0;
      for (auto &P : part)
      {
        if (!P.ok)
          return false;
        while (P.blob.size() % 4)
          P.blob.push_back(0);
        blob_words += P.blob.size();
        ucols += P.ucol.size();
      }
      if (blob_words / 4 >= (size_t)1 << 31 || ucols >= (size_t)1 << 31)
        return false;
      F.blob.resize(blob_words);
      F.ucol.resize(ucols);
      std::vector<TileDesc> all;
      std::vector<char> halo;
      all.reserve(ntiles);
      size_t bw = 0, uw = 0;
      F.nsteps = F.nvals = 0;
      for (auto &P : part)
      {
        if (!P.blob.empty())
          std::memcpy(&F.blob[bw], P.blob.data(), P.blob.size() * sizeof(int));
        if (!P.ucol.empty())
          std::memcpy(&F.ucol[uw], P.ucol.data(), P.ucol.size() * sizeof(int));
        for (size_t i = 0; i < P.desc.size(); ++i)
        {
          TileDesc d = P.desc[i];
          d.blob16 += (int)(bw / 4);
          d.ucol0 += (int)uw;
          all.push_back(d);
          halo.push_back(P.halo[i]);
        }
        bw += P.blob.size();
        uw += P.ucol.size();
        F.nsteps += P.nsteps;
        F.nvals += P.nvals;
        std::vector<int>().swap(P.blob);
        std::vector<int>().swap(P.ucol);
      }
      F.tile.clear();
      F.tile.reserve(ntiles);
      for (int pass = 0; pass < 2; ++pass)
      {
        for (int t = 0; t < ntiles; ++t)
          if ((halo[t] != 0) == (pass == 1))
            F.tile.push_back(all[t]);
        if (pass == 0)
          F.n_interior = (int)F.tile.size();
      }
      F.ntiles = ntiles;
      F.nblocks = (long long)(o.rows.size() / 8);
      F.max_len16 = F.max_u = 0;
      for (const TileDesc &d : F.tile)
      {
        F.max_len16 = std::max(F.max_len16, d.len16);
        F.max_u = std::max(F.max_u, d.nu);
      }
      // every tile must fit one pipeline stage of the widest pass (two stages per SM)
      size_t worst = 0;
      for (const TileDesc &d : F.tile)
        worst = std::max(worst, stage_bytes(d.len16, d.nu, kRowBytesWide));
      // the kernel sizes its stages by (max_len16, max_u), which need not come from the same tile
      if (stage_bytes(F.max_len16, F.max_u, kRowBytesWide) > (size_t)kStageBudget || worst > (size_t)kStageBudget)
        return false;
      F.valid = true;
      return true;
    }

    /** Which rows form the tiles. Candidates are tried in order: box tiles of decreasing size if the pattern is a
     *  structured grid (a candidate is accepted when the tile around the middle grid point fits a pipeline stage),
     *  then consecutive rows. `next` is where to resume if the chosen candidate turns out not to fit. */
    struct Plan
    {
      Order order;
      bool grid = false;
      long long S1 = 0, S2 = 0;
      int tw = 0, th = 0, td = 0;
      int next = 0;
    };

    struct Cand
    {
      int tw, th, td, bw, bh, bd;
    };
    inline const Cand *candidates(bool three_d, int &count)
    {
      static const Cand cand3[] = {{8, 4, 4, 2, 2, 2}, {6, 4, 4, 2, 2, 2}, {4, 4, 4, 2, 2, 2}, {8, 4, 2, 2, 2, 2},
                                   {4, 4, 2, 2, 2, 2}, {4, 2, 2, 2, 2, 2}, {2, 2, 2, 2, 2, 2}};
      static const Cand cand2[] = {{16, 16, 1, 4, 2, 1}, {16, 12, 1, 4, 2, 1}, {16, 8, 1, 4, 2, 1}, {8, 8, 1, 4, 2, 1},
                                   {8, 4, 1, 4, 2, 1},   {4, 2, 1, 4, 2, 1}};
      count = three_d ? 7 : 6;
      return three_d ? cand3 : cand2;
    }
    constexpr int kLinearCandidate = 100;

    template <class Ptr, class Idx>
    bool plan(long long n, long long ncols, const Ptr *rowptr, const Idx *col, const double *val, long long n_owned, int first,
              Plan &P)
    {
      P = Plan();
      if (n <= 0)
        return false;
      long long S1 = 0, S2 = 0;
      if (first < kLinearCandidate && detect_grid(n, rowptr, col, n_owned, S1, S2))
      {
        int ncand = 0;
        const Cand *cands = candidates(S2 > 0, ncand);
        const long long nx = S1, ny = S2 > 0 ? S2 / S1 : (n + S1 - 1) / S1, nz = S2 > 0 ? (n + S2 - 1) / S2 : 1;
        TileBuilder<Ptr, Idx> tb(ncols, rowptr, col, val, n_owned);
        for (int ci = first; ci < ncand; ++ci)
        {
          const Cand &c = cands[ci];
          // sample: the tile that contains the middle grid point
          std::vector<int> srows;
          const long long x0 = (nx / 2 / c.tw) * c.tw, y0 = (ny / 2 / c.th) * c.th, z0 = (nz / 2 / c.td) * c.td;
          for (long long zb = z0; zb < std::min<long long>(z0 + c.td, nz); zb += c.bd)
            for (long long yb = y0; yb < std::min<long long>(y0 + c.th, ny); yb += c.bh)
              for (long long xb = x0; xb < std::min<long long>(x0 + c.tw, nx); xb += c.bw)
                for (int dz = 0; dz < c.bd; ++dz)
                  for (int dy = 0; dy < c.bh; ++dy)
                    for (int dx = 0; dx < c.bw; ++dx)
                    {
                      const long long x = xb + dx, y = yb + dy, z = zb + dz;
                      const long long r = (z * ny + y) * nx + x;
                      srows.push_back((x < nx && y < ny && z < nz && r < n) ? (int)r : -1);
                    }
          std::vector<int> sblob, sucol;
          TileDesc sd;
          bool hh;
          long long a = 0, b = 0;
          if (srows.empty() || !tb.build(srows.data(), (int)(srows.size() / 8), sblob, sucol, sd, hh, a, b))
            continue;
          // 6 % head room: boundary tiles of a distributed matrix (halo columns) and irregular rows may be larger
          if (stage_bytes(sd.len16, sd.nu, kRowBytesWide) * 106 / 100 > (size_t)kStageBudget)
            continue;
          grid_order(n, S1, S2, c.tw, c.th, c.td, c.bw, c.bh, c.bd, P.order);
          P.grid = true;
          P.S1 = S1;
          P.S2 = S2;
          P.tw = c.tw;
          P.th = c.th;
          P.td = c.td;
          P.next = ci + 1;
          return true;
        }
      }
      if (first > kLinearCandidate)
        return false;
      P.next = kLinearCandidate + 1;
      return linear_order(n, ncols, rowptr, col, P.order);
    }

    /** tile sizes (len16, nu) -> do all tiles fit a stage of the widest pass? */
    inline bool fits_budget(int max_len16, int max_u) { return stage_bytes(max_len16, max_u, kRowBytesWide) <= (size_t)kStageBudget; }

    /** complete construction on the host: plan, build; a candidate whose tiles outgrow the budget is replaced by the next. */
    template <class Ptr, class Idx>
    bool build(long long n, long long ncols, const Ptr *rowptr, const Idx *col, const double *val, long long n_owned, Format &F,
               int nthreads = 0)
    {
      F = Format();
      Plan P;
      int first = 0;
      while (plan(n, ncols, rowptr, col, val, n_owned, first, P))
      {
        F = Format();
        if (build_tiles(ncols, rowptr, col, val, n_owned, P.order, F, nthreads))
        {
          F.grid = P.grid;
          F.S1 = P.S1;
          F.S2 = P.S2;
          F.tw = P.tw;
          F.th = P.th;
          F.td = P.td;
          return true;
        }
        first = P.next;
      }
      F = Format();
      return false;
    }

  } // namespace brb
} // namespace de
