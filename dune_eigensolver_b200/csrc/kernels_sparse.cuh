// SpMM (K1 / K1'), diag-dot (K2), partial-sum reduction and small data-movement kernels for sm_100a.
//
// Data layout in HBM: a multivector is a dense row-major n x m fp64 array (leading dimension ld = m doubles,
// rows 64-byte aligned because m % 8 == 0). One matrix row touches one contiguous 8*m-byte row of X per
// nonzero, so the gather is made of full 32-byte sectors and 128-bit loads; A is CSR with int32 indices.
//
// Roofline: all kernels here are HBM-bound. Algorithmic bytes per call:
//   SpMM      12*nnz + 4*(n+1) + 16*n*m        diag-dot  16*n*m
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace de
{

  /** Programmatic dependent launch (kernels of the asynchronous driver loop are launched with
   *  cudaLaunchAttributeProgrammaticStreamSerialization): wait until the preceding kernel of the stream has completed
   *  and its writes are visible -- everything this kernel reads comes from it -- then allow the NEXT kernel's CTAs to be
   *  scheduled as SMs become free (they block in their own prologue). Only launch latency is overlapped; without the
   *  launch attribute both instructions do nothing. */
  __device__ __forceinline__ void pdl_prologue()
  {
    asm volatile("griddepcontrol.wait;\n" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");
  }

  __device__ __forceinline__ double2 ldg2(const double *p) { return __ldg(reinterpret_cast<const double2 *>(p)); }
  __device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
  __device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }
  __device__ __forceinline__ void fma2(double2 &acc, double a, double2 x)
  {
    acc.x = fma(a, x.x, acc.x);
    acc.y = fma(a, x.y, acc.y);
  }

  __device__ __forceinline__ void cp_async16_sparse(void *smem_dst, const void *gmem_src)
  {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(gmem_src));
  }

  struct SpmmArgs
  {
    long long nrows;     // rows to process (length of rowlist if given, else rows 0..nrows-1)
    const int *rowlist;  // optional list of local row indices (interior / boundary split of a distributed matrix)
    const int *rowptr;   // CSR row pointer (int32: nnz < 2^31 per GPU)
    const int *col;      // local column indices: [0,n_owned) -> X, [n_owned, n_owned+n_halo) -> H
    const double *val;
    const double *X;     // owned rows of the input block
    const double *H;     // halo rows received from peers (may be null when n_halo == 0)
    long long n_owned;
    int ld;              // leading dimension (doubles) of X, H and Y
    int m;               // columns
    double *Y;
    double *partials;    // DOT only: [gridDim.x][m] per-CTA partial dot products
    const int *done;     // optional device flag: a driver loop has converged, the launch is a no-op
  };

  /** Y = A X for all m columns in ONE pass over A (the reference re-streams A once per 8-column panel,
   *  kernels_cpp.hh:640-656). TPR threads cooperate on a row, each owning VPT double2 column pairs
   *  (columns 2*(t + v*TPR)); a row's nonzeros are accumulated in CSR order with FMA, like the CPU loop
   *  kernels_cpp.hh:644-655. The (val,col) loads of a row are warp-broadcast; the X-row gather is a coalesced
   *  8*m-byte segment. Consecutive CTAs walk consecutive row blocks so X rows shared by neighbouring matrix
   *  rows (stencil reuse distance = one grid plane) stay in the 126 MB L2.
   *  DOT additionally accumulates dp[j] += X(i,j) * Y(i,j) while the Y row is still in registers (K1':
   *  eigensolver.hh:84-85) and leaves one partial vector per CTA (reduced in fixed order afterwards).
   */
  template <int TPR, int VPT, bool DOT>
  __global__ void __launch_bounds__(256) spmm_kernel(const SpmmArgs a)
  {
    if (a.done != nullptr && *a.done != 0)
      return;
    constexpr int RPB = 256 / TPR; // rows per CTA per step
    const int t = threadIdx.x % TPR;
    const int rslot = threadIdx.x / TPR;
    int cidx[VPT];
    bool act[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v)
    {
      cidx[v] = 2 * (t + v * TPR);
      act[v] = cidx[v] < a.m;
    }
    double2 dacc[VPT];
#pragma unroll
    for (int v = 0; v < VPT; ++v)
      dacc[v] = make_double2(0.0, 0.0);

    for (long long r = (long long)blockIdx.x * RPB + rslot; r < a.nrows; r += (long long)gridDim.x * RPB)
    {
      const long long row = a.rowlist ? (long long)a.rowlist[r] : r;
      int k = __ldg(a.rowptr + row);
      const int kend = __ldg(a.rowptr + row + 1);
      double2 acc[VPT];
#pragma unroll
      for (int v = 0; v < VPT; ++v)
        acc[v] = make_double2(0.0, 0.0);

      for (; k + 4 <= kend; k += 4)
      {
        int j[4];
        double av[4];
        const double *xr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
        {
          j[u] = __ldg(a.col + k + u);
          av[u] = __ldg(a.val + k + u);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          xr[u] = (j[u] < a.n_owned) ? a.X + (size_t)j[u] * a.ld : a.H + (size_t)(j[u] - a.n_owned) * a.ld;
        double2 xv[4][VPT];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < VPT; ++v)
            xv[u][v] = act[v] ? ldg2(xr[u] + cidx[v]) : make_double2(0.0, 0.0);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < VPT; ++v)
            fma2(acc[v], av[u], xv[u][v]);
      }
      for (; k < kend; ++k)
      {
        const int j = __ldg(a.col + k);
        const double av = __ldg(a.val + k);
        const double *xr = (j < a.n_owned) ? a.X + (size_t)j * a.ld : a.H + (size_t)(j - a.n_owned) * a.ld;
#pragma unroll
        for (int v = 0; v < VPT; ++v)
          if (act[v])
            fma2(acc[v], av, ldg2(xr + cidx[v]));
      }
#pragma unroll
      for (int v = 0; v < VPT; ++v)
        if (act[v])
        {
          st2(a.Y + (size_t)row * a.ld + cidx[v], acc[v]);
          if (DOT)
          {
            const double2 z = ldg2(a.X + (size_t)row * a.ld + cidx[v]);
            dacc[v].x = fma(z.x, acc[v].x, dacc[v].x);
            dacc[v].y = fma(z.y, acc[v].y, dacc[v].y);
          }
        }
    }

    if (DOT)
    {
      __shared__ double2 red[VPT][256];
#pragma unroll
      for (int v = 0; v < VPT; ++v)
        red[v][threadIdx.x] = dacc[v];
      __syncthreads();
      if (rslot == 0)
      {
#pragma unroll
        for (int v = 0; v < VPT; ++v)
          if (act[v])
          {
            double2 s = make_double2(0.0, 0.0);
            for (int q = 0; q < RPB; ++q) // fixed order: deterministic
            {
              s.x += red[v][q * TPR + t].x;
              s.y += red[v][q * TPR + t].y;
            }
            st2(a.partials + (size_t)blockIdx.x * a.m + cidx[v], s);
          }
      }
    }
  }

  /** Second-generation SpMM for m = 2*TPR exactly (m = 8/16/32/64). Same mapping as spmm_kernel (TPR lanes own the
   *  m columns of a row as double2 pairs; CSR-order FMA accumulation), but built around what ncu showed for the first
   *  version (profiles/r01_ncu_kernels_baseline.csv: 28 warp instructions per nonzero, issue slots 49 % busy,
   *  no memory unit saturated): 32-bit element offsets instead of 64-bit address arithmetic, the halo select
   *  compiled out for single-GPU matrices, no per-column predicates, and nonzeros consumed in predicated chunks
   *  of 8 so that 8 independent X-row gathers are in flight per thread (the whole 7-point row, a third of a
   *  27-point row) instead of 4 followed by a serial tail. */
  template <int TPR, bool DOT, bool HALO>
  __global__ void __launch_bounds__(256, 3) spmm_kernel_v2(const SpmmArgs a)
  {
    if (a.done != nullptr && *a.done != 0)
      return;
    constexpr int RPB = 256 / TPR;
    constexpr int CH = 8;
    const int t = threadIdx.x % TPR;
    const int rslot = threadIdx.x / TPR;
    const unsigned ldh = (unsigned)TPR; // double2 elements per row
    const double2 *__restrict__ Xv = reinterpret_cast<const double2 *>(a.X) + t;
    const double2 *__restrict__ Hv = reinterpret_cast<const double2 *>(a.H) + t;
    double2 *__restrict__ Yv = reinterpret_cast<double2 *>(a.Y) + t;
    const int n_owned = (int)a.n_owned;
    double2 dacc = make_double2(0.0, 0.0);

    for (long long r = (long long)blockIdx.x * RPB + rslot; r < a.nrows; r += (long long)gridDim.x * RPB)
    {
      const int row = a.rowlist ? a.rowlist[r] : (int)r;
      const int kbeg = __ldg(a.rowptr + row), kend = __ldg(a.rowptr + row + 1);
      double2 acc = make_double2(0.0, 0.0);
      const int klast = kend - 1;
      for (int k = kbeg; k < kend; k += CH)
      {
        // Slots past the end of the row re-read the row's last nonzero with a zero coefficient: every load is
        // unpredicated, so all 8 index/value loads and then all 8 X-row gathers issue back to back.
        int j[CH];
        double av[CH];
        double2 xv[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u)
        {
          const int kk = min(k + u, klast);
          j[u] = __ldg(a.col + kk);
          av[u] = __ldg(a.val + kk);
        }
#pragma unroll
        for (int u = 0; u < CH; ++u)
        {
          if (HALO)
            xv[u] = __ldg((j[u] < n_owned) ? Xv + (unsigned)j[u] * ldh : Hv + (unsigned)(j[u] - n_owned) * ldh);
          else
            xv[u] = __ldg(Xv + (unsigned)j[u] * ldh);
        }
#pragma unroll
        for (int u = 0; u < CH; ++u)
          fma2(acc, (k + u <= klast) ? av[u] : 0.0, xv[u]);
      }
      Yv[(unsigned)row * ldh] = acc;
      if (DOT)
      {
        const double2 z = __ldg(Xv + (unsigned)row * ldh);
        dacc.x = fma(z.x, acc.x, dacc.x);
        dacc.y = fma(z.y, acc.y, dacc.y);
      }
    }

    if (DOT)
    {
      __shared__ double2 red[256];
      red[threadIdx.x] = dacc;
      __syncthreads();
      if (rslot == 0)
      {
        double2 s = make_double2(0.0, 0.0);
        for (int q = 0; q < RPB; ++q) // fixed order: deterministic
        {
          s.x += red[q * TPR + t].x;
          s.y += red[q * TPR + t].y;
        }
        st2(a.partials + (size_t)blockIdx.x * a.m + 2 * t, s);
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // Staged SpMM: the CSR stream of a block of rows is brought into shared memory with coalesced cp.async
  // ------------------------------------------------------------------------------------------------
  // spmm_kernel_v2 pays three DEPENDENT memory latencies per row (row pointer -> column/value -> X row); for short
  // rows (7-point stencil: one chunk per row) that chain, not bandwidth, bounds the kernel (ncu: issue 30-38 %,
  // no unit saturated). Here a CTA owns contiguous row blocks (<= 256 rows, <= 2048 nonzeros, cut on the host at
  // matrix creation, `blk_meta` = {first row, end row, first nonzero, end nonzero}); the block's row pointers,
  // column indices and values are copied to shared memory with 16-byte cp.async, double-buffered one block ahead,
  // so each matrix byte is read from HBM exactly once, fully coalesced, and the per-row chain shrinks to
  // shared-memory read -> X gather. Rows are then processed exactly as in spmm_kernel_v2 (same lanes, same CSR
  // accumulation order, same DOT epilogue). A block whose single row exceeds the staging capacity is read from
  // global memory directly.
  constexpr int kStageCapNnz = 2048;
  constexpr int kStageMaxRows = 256;

  struct StagedArgs
  {
    int nblocks;
    const int4 *blk_meta;  // {r0, r1, k0, k1}
    const int *rowmap;     // optional: local output row of each matrix row (permuted sub-matrix); null = identity
    const int *rowptr;     // arrays are allocated with 16 bytes of tail padding
    const int *col;
    const double *val;
    const double *X;
    const double *H;
    int n_owned;
    int m;
    double *Y;
    double *partials;
    const int *done; // optional: see SpmmArgs
  };

  constexpr size_t spmm_staged_smem_bytes()
  {
    return 2 * ((kStageMaxRows + 8) * sizeof(int) + (kStageCapNnz + 8) * sizeof(int) + (kStageCapNnz + 4) * sizeof(double));
  }

  template <int TPR, bool DOT, bool HALO>
  __global__ void __launch_bounds__(256, 3) spmm_staged_kernel(const StagedArgs a)
  {
    if (a.done != nullptr && *a.done != 0)
      return;
    constexpr int RPB = 256 / TPR;
    constexpr int CH = 8;
    extern __shared__ __align__(16) unsigned char dyn[];
    double *s_val0 = reinterpret_cast<double *>(dyn);                    // 2 x (CAP+4) doubles
    int *s_col0 = reinterpret_cast<int *>(s_val0 + 2 * (kStageCapNnz + 4)); // 2 x (CAP+8) ints
    int *s_ptr0 = s_col0 + 2 * (kStageCapNnz + 8);                       // 2 x (MAXROWS+8) ints

    const int tid = threadIdx.x;
    const int t = tid % TPR, gslot = tid / TPR;
    const unsigned ldh = (unsigned)TPR;
    const double2 *__restrict__ Xv = reinterpret_cast<const double2 *>(a.X) + t;
    const double2 *__restrict__ Hv = reinterpret_cast<const double2 *>(a.H) + t;
    double2 *__restrict__ Yv = reinterpret_cast<double2 *>(a.Y) + t;
    double2 dacc = make_double2(0.0, 0.0);

    auto stage_block = [&](int4 mt, int s)
    {
      // mt.x >= mt.y marks "no block"
      if (mt.x < mt.y)
      {
        const int ra = mt.x & ~3;
        const int nptr = (mt.y + 1 - ra + 3) >> 2;
        int *sp = s_ptr0 + s * (kStageMaxRows + 8);
        for (int c = tid; c < nptr; c += 256)
          cp_async16_sparse(sp + 4 * c, a.rowptr + ra + 4 * c);
        if (mt.w - mt.z <= kStageCapNnz)
        {
          const int ka = mt.z & ~3, kva = mt.z & ~1;
          const int ncol = (mt.w - ka + 3) >> 2, nval = (mt.w - kva + 1) >> 1;
          int *sc = s_col0 + s * (kStageCapNnz + 8);
          double *sv = s_val0 + s * (kStageCapNnz + 4);
          for (int c = tid; c < ncol; c += 256)
            cp_async16_sparse(sc + 4 * c, a.col + ka + 4 * c);
          for (int c = tid; c < nval; c += 256)
            cp_async16_sparse(sv + 2 * c, a.val + kva + 2 * c);
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
      stage_block(nxt, s ^ 1);                       // prefetch one block ahead
      const int4 nxt2 = load_meta(b + 2 * gridDim.x); // metadata two blocks ahead (no dependent latency later)
      asm volatile("cp.async.wait_group 1;\n" ::);
      __syncthreads();

      const bool direct = (cur.w - cur.z) > kStageCapNnz;
      const int *sp = s_ptr0 + s * (kStageMaxRows + 8) - (cur.x & ~3);
      const int *cp = direct ? a.col : s_col0 + s * (kStageCapNnz + 8) - (cur.z & ~3);
      const double *vp = direct ? a.val : s_val0 + s * (kStageCapNnz + 4) - (cur.z & ~1);

      for (int r = cur.x + gslot; r < cur.y; r += RPB)
      {
        const int kbeg = sp[r], kend = sp[r + 1];
        const int klast = kend - 1;
        double2 acc = make_double2(0.0, 0.0);
        for (int k = kbeg; k < kend; k += CH)
        {
          int j[CH];
          double av[CH];
          double2 xv[CH];
#pragma unroll
          for (int u = 0; u < CH; ++u)
          {
            const int kk = min(k + u, klast);
            j[u] = cp[kk];
            av[u] = vp[kk];
          }
#pragma unroll
          for (int u = 0; u < CH; ++u)
          {
            if (HALO)
              xv[u] = __ldg((j[u] < a.n_owned) ? Xv + (unsigned)j[u] * ldh : Hv + (unsigned)(j[u] - a.n_owned) * ldh);
            else
              xv[u] = __ldg(Xv + (unsigned)j[u] * ldh);
          }
#pragma unroll
          for (int u = 0; u < CH; ++u)
            fma2(acc, (k + u <= klast) ? av[u] : 0.0, xv[u]);
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
      __syncthreads(); // stage s may be overwritten by the prefetch issued in the next iteration
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

  /** dp[j] = sum_i X(i,j) Y(i,j) (reference dot_products_diagonal_blocked, kernels_cpp.hh:24-55).
   *  blockDim = (m/2, 256/(m/2)): x indexes a column pair, y a row lane; rows are strided over the grid.
   *  Leaves one partial vector per CTA. */
  static __global__ void __launch_bounds__(256) diag_dot_kernel(long long n, const double *__restrict__ X, int ldx,
                                                         const double *__restrict__ Y, int ldy, int m,
                                                         double *__restrict__ partials)
  {
    const int c = 2 * threadIdx.x;
    const long long step = (long long)gridDim.x * blockDim.y;
    double2 acc = make_double2(0.0, 0.0);
    long long r = (long long)blockIdx.x * blockDim.y + threadIdx.y;
    for (; r + 3 * step < n; r += 4 * step)
    {
      double2 xv[4], yv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        xv[u] = ldg2(X + (size_t)(r + u * step) * ldx + c);
        yv[u] = ldg2(Y + (size_t)(r + u * step) * ldy + c);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
      {
        acc.x = fma(xv[u].x, yv[u].x, acc.x);
        acc.y = fma(xv[u].y, yv[u].y, acc.y);
      }
    }
    for (; r < n; r += step)
    {
      const double2 xv = ldg2(X + (size_t)r * ldx + c), yv = ldg2(Y + (size_t)r * ldy + c);
      acc.x = fma(xv.x, yv.x, acc.x);
      acc.y = fma(xv.y, yv.y, acc.y);
    }
    __shared__ double2 red[256];
    red[threadIdx.y * blockDim.x + threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0)
    {
      double2 s = make_double2(0.0, 0.0);
      for (int q = 0; q < (int)blockDim.y; ++q)
      {
        s.x += red[q * blockDim.x + threadIdx.x].x;
        s.y += red[q * blockDim.x + threadIdx.x].y;
      }
      st2(partials + (size_t)blockIdx.x * m + c, s);
    }
  }

  /** out[e] = sum_p partials[p*len + e], p ascending: the fixed-order (deterministic) second stage of every
   *  reduction. blockDim = (32,32); each CTA owns 32 consecutive outputs. */
  static __global__ void __launch_bounds__(1024) reduce_partials_kernel(const double *__restrict__ partials, int nparts,
                                                                 int len, double *__restrict__ out,
                                                                 const int *__restrict__ done = nullptr)
  {
    if (done != nullptr && *done != 0)
      return;
    __shared__ double red[32][33];
    const int e = blockIdx.x * 32 + threadIdx.x;
    double s = 0.0;
    if (e < len)
      for (int p = threadIdx.y; p < nparts; p += 32)
        s += partials[(size_t)p * len + e];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < len)
    {
      double tot = 0.0;
#pragma unroll
      for (int q = 0; q < 32; ++q)
        tot += red[q][threadIdx.x];
      out[e] = tot;
    }
  }

  /** Device-side convergence test of the driver loops (reference eigensolver.hh:86-102, :176-189):
   *  s_j = dp_j - shift, distance = max_j |s_j - s_prev_j|, s_prev <- s; iteration k > 1 with distance < tol raises
   *  the `done` flag, after which every kernel of the iterations already enqueued returns at once.
   *  state: flags[1] = done, flags[2] = last completed iteration; hist[k] = distance of iteration k. One CTA of 64. */
  constexpr int kConvHistory = 1 << 20; // distances of the first 2^20 iterations are kept for the verbose listing

  __device__ __forceinline__ void convergence_body(int tid, int nthreads, int k, int m, double shift, double tol,
                                                   const double *__restrict__ dp, double *__restrict__ s_prev,
                                                   double *__restrict__ hist, int *__restrict__ flags)
  {
    __shared__ double red[64];
    double d = 0.0;
    if (tid < m)
    {
      const double s = __ldcg(dp + tid) - shift;
      d = fabs(s - s_prev[tid]);
      if (!(d == d))
        d = 1.0e300; // NaN never converges
      s_prev[tid] = s;
    }
    if (tid < 64)
      red[tid] = d;
    __syncthreads();
    if (tid == 0)
    {
      double mx = 0.0;
      for (int q = 0; q < 64; ++q)
        mx = fmax(mx, red[q]);
      if (k < 0) // the iteration number lives on the device (flags[3]): launches replayed from a CUDA graph carry no number
      {
        k = flags[3] + 1;
        flags[3] = k;
      }
      if (k < kConvHistory) // a "run until converged" maxiter of 1e9 must not size a buffer (de_drivers.cu)
        hist[k] = mx;
      flags[2] = k;
      if (k > 1 && mx < tol)
        flags[1] = 1;
    }
    (void)nthreads;
  }

  static __global__ void __launch_bounds__(64) convergence_kernel(int k, int m, double shift, double tol, const double *__restrict__ dp,
                                                           double *__restrict__ s_prev, double *__restrict__ hist,
                                                           int *__restrict__ flags)
  {
    if (flags[1] != 0)
      return;
    convergence_body(threadIdx.x, 64, k, m, shift, tol, dp, s_prev, hist, flags);
  }

  // ---- layout conversion at the boundary (reference MultiVector layout <-> row-major) ------------------
  /** to_rowmajor: dst[i*m + j] = src[((j/8)*n + i)*8 + j%8]; else the inverse. One thread per (row, 8-col panel). */
  static __global__ void __launch_bounds__(256) panel8_convert_kernel(long long n, int m, const double *__restrict__ src,
                                                               double *__restrict__ dst, int to_rowmajor)
  {
    const int np = m / 8;
    const long long total = n * np;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x)
    {
      const long long i = e / np;
      const int p = (int)(e % np);
      const size_t pan = ((size_t)p * n + i) * 8, row = (size_t)i * m + 8 * p;
      const double *s = to_rowmajor ? src + pan : src + row;
      double *d = to_rowmajor ? dst + row : dst + pan;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        st2(d + 2 * q, ldg2(s + 2 * q));
    }
  }

  /** evec[j*n + i] = X(i,j), j < nev (copy-out of eigensolver.hh:109-111): tiled transpose through smem. */
  static __global__ void __launch_bounds__(256) extract_columns_kernel(long long n, int m, int nev,
                                                                const double *__restrict__ X,
                                                                double *__restrict__ out)
  {
    __shared__ double tile[32][65];
    const long long i0 = (long long)blockIdx.x * 32;
    // load 32 rows x m columns (m <= 64), coalesced along columns
    for (int e = threadIdx.x; e < 32 * m; e += blockDim.x)
    {
      const int r = e / m, c = e % m;
      tile[r][c] = (i0 + r < n) ? X[(size_t)(i0 + r) * m + c] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * nev; e += blockDim.x)
    {
      const int c = e / 32, r = e % 32;
      if (i0 + r < n)
        out[(size_t)c * n + i0 + r] = tile[r][c];
    }
  }

  /** halo pack: buf[s*m + c] = X[rows[s]*m + c] */
  static __global__ void __launch_bounds__(256) pack_rows_kernel(long long count, const int *__restrict__ rows, int m,
                                                          const double *__restrict__ X, double *__restrict__ buf)
  {
    const int hp = m / 2;
    const long long total = count * hp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x)
    {
      const long long s = e / hp;
      const int c = 2 * (int)(e % hp);
      st2(buf + (size_t)s * m + c, ldg2(X + (size_t)rows[s] * m + c));
    }
  }

} // namespace de
