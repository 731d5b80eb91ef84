// Partial-sum reduction with a fused TAIL: what used to be three or four dependent launches of tiny kernels after
// every tall-skinny reduction (reduce the per-CTA partials -> all-reduce across GPUs -> Cholesky + inverse of the Gram
// matrix, or the convergence test on the Rayleigh quotients) is one launch. Every CTA reduces its 32 outputs as before
// (fixed order: deterministic); the CTA that finishes last -- a ticket counter tells -- then runs, with its 1024
// threads, the one-shot peer all-reduce (kernels_peer.cuh) and the Cholesky / convergence step on the complete vector.
// On a 27-point 100^3, m = 32 solve the small kernels were 12 % of the time on one GPU and 46 % on eight, almost all
// of it launch and dependency latency.
#pragma once

#include "de_types.hpp"
#include "kernels_dense.cuh"
#include "kernels_peer.cuh"
#include "kernels_sparse.cuh"

namespace de
{


  /** blockDim = (32, 32); gridDim.x = ceil((len + t.len2) / 32).
   *  out = [ second segment (t.len2 values reduced from t.partials2) | first segment (len values from partials) ]:
   *  the second segment carries the Rayleigh-quotient partials the preceding SpMM left behind, so that their
   *  reduction, the Gram reduction, ONE all-reduce over both, the convergence test and the Cholesky share a launch
   *  (one all-reduce wait per iteration less on several GPUs). */
  static __global__ void __launch_bounds__(1024) reduce_tail_kernel(const double *__restrict__ partials, int nparts, int len,
                                                             double *__restrict__ out, const int *__restrict__ done_in,
                                                             const TailArgs t)
  {
    pdl_prologue();
    if (done_in != nullptr && *done_in != 0)
      return;
    if (t.skip != nullptr && *t.skip != 0)
      return; // one CholQR sweep was enough: the kernel that would have written the partials skipped itself too
    __shared__ double red[32][33];
    __shared__ int last;
    const int tid = threadIdx.y * 32 + threadIdx.x;
    const int e = blockIdx.x * 32 + threadIdx.x;
    const int total = len + t.len2;
    double s = 0.0;
    if (e < t.len2)
    {
      for (int p = threadIdx.y; p < t.nparts2; p += 32)
        s += t.partials2[(size_t)p * t.len2 + e];
    }
    else if (e < total)
      for (int p = threadIdx.y; p < nparts; p += 32)
        s += partials[(size_t)p * len + (e - t.len2)];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && e < total)
    {
      double tot = 0.0;
#pragma unroll
      for (int q = 0; q < 32; ++q)
        tot += red[q][threadIdx.x];
      out[e] = tot;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0)
      last = (atomicAdd(t.ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!last)
      return;
    if (tid == 0)
      *t.ticket = 0;
    __threadfence();
    if (t.do_allreduce)
      peer_allreduce_body(t.pa, tid, out, total);
    if (t.kind & kTailConv)
    {
      convergence_body(tid, 1024, t.k, t.m, t.shift, t.tol, out, t.s_prev, t.hist, t.flags);
      __syncthreads();
      if ((t.kind & kTailChol) && *reinterpret_cast<volatile int *>(t.flags + 1) != 0)
        return; // converged: the blocks stay as they are
    }
    if (t.kind & kTailChol)
    {
      if (t.m <= 32)
        chol_inverse2_body<32>(tid, t.m, out + t.len2, t.Rinv, t.status, t.info, t.identity_flag, t.done, t.wellcond,
                               t.wellcond ? t.flags_identity : nullptr);
      else
        chol_inverse2_body<64>(tid, t.m, out + t.len2, t.Rinv, t.status, t.info, t.identity_flag, t.done, t.wellcond,
                               t.wellcond ? t.flags_identity : nullptr);
    }
  }

} // namespace de
