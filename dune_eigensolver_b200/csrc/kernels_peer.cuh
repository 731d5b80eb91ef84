// Multi-GPU data path over NVLink peer memory (one process per GPU; the windows are exchanged as CUDA IPC handles).
//
// Every rank owns a WINDOW in its HBM that all peers map:
//     [0, 4 KB)            flags      ar_flag[2 channels][2][8] | halo_flag[2][8] (64-bit epochs, written by the peers)
//     [4 KB, +2*2*8 slots) all-reduce slots, one per (channel, parity, sender rank), kPeerSlotDoubles doubles each
//     [halo_off, +2*cap)   two halo buffers (parity of the SpMM call), laid out like the matrix's halo block
//
//  * peer_allreduce_kernel: one-shot all-reduce of a short vector (the m Rayleigh quotients and the m x m Gram
//    matrices, <= 33 KB): every rank stores its vector into its slot in EVERY window, releases a flag there, waits for
//    the flags of all ranks in its own window and sums the slots in rank order -- the same order on every rank, so
//    all ranks hold bit-identical results (the replicated Cholesky and the convergence flag depend on that).
//    One launch, ~2 NVLink latencies; an NCCL all-reduce of the same vector costs 25-40 us and a second kernel.
//  * halo_push_kernel: the rows of X that neighbours need are stored straight into the neighbours' halo buffers
//    (instead of pack -> ncclSend/ncclRecv on a second stream, whose point-to-point kernel competes with the
//    persistent SpMM kernel for SMs); the last CTA to finish releases the flags. halo_wait_kernel (one warp) is the
//    only thing the boundary tiles wait for.
//
// Flow control needs no credits: consecutive SpMM calls alternate between the two halo buffers, and a rank cannot be
// two calls ahead of a neighbour because its boundary rows need that neighbour's rows of the call in between; an
// all-reduce is itself a barrier, so two slot sets suffice. Spins give up after ~30 s and raise an error flag (a neighbour that died must not hang the GPU).
#pragma once

#include <cstdint>

#include <cuda_runtime.h>

#include "de_types.hpp"

namespace de
{


  __device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
  {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
  }
  __device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
  {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
  }
  /** what = 1: all-reduce contribution, 2: halo rows; the error word records what, from which rank and the epoch */
  __device__ __forceinline__ bool peer_wait(const unsigned long long *flag, unsigned long long epoch, int *err,
                                            long long timeout, int what = 1, int from = 0)
  {
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < epoch)
    {
      if (clock64() - t0 > timeout) // default ~30 s: ranks of a multi-process job can be seconds apart on the host
      {
        if (err) // keep the FIRST failure: later waits of the same rank fail as a consequence
          atomicCAS(err, 0, what | (from << 4) | (int)((epoch & 0xfffull) << 8) | (int)((ld_acquire_sys(flag) & 0x7ffull) << 20));
        return false;
      }
      __nanosleep(64);
    }
    return true;
  }

  // set = 2 * channel + parity
  __host__ __device__ __forceinline__ unsigned long long *peer_ar_flag(unsigned char *base, int set, int sender)
  {
    return reinterpret_cast<unsigned long long *>(base) + set * kPeerMaxRanks + sender;
  }
  __host__ __device__ __forceinline__ unsigned long long *peer_halo_flag(unsigned char *base, int parity, int sender)
  {
    return reinterpret_cast<unsigned long long *>(base) + (2 * kPeerArChannels + parity) * kPeerMaxRanks + sender;
  }
  __device__ __forceinline__ double *peer_ar_slot(unsigned char *base, int set, int sender)
  {
    return reinterpret_cast<double *>(base + kPeerArOff) + (size_t)(set * kPeerMaxRanks + sender) * kPeerSlotDoubles;
  }

  /** buf[0..len) <- sum over ranks, identical bits on every rank; executed by one CTA of 1024 threads (tid).
   *  buf may have been written by other CTAs of the same launch: it is read through L2. */
  __device__ __forceinline__ void peer_allreduce_body(const PeerArgs &pa, int tid, double *__restrict__ buf, int len)
  {
    const int parity = 2 * pa.channel + (int)(pa.epoch & 1ull); // the slot / flag set
    for (int q = 0; q < pa.nranks; ++q)
    {
      double *dst = peer_ar_slot(pa.base[q], parity, pa.rank);
      for (int i = tid; i < len; i += 1024)
        dst[i] = __ldcg(buf + i);
    }
    __threadfence_system();
    __syncthreads();
    if (tid < pa.nranks)
    {
      st_release_sys(peer_ar_flag(pa.base[tid], parity, pa.rank), pa.epoch);
      peer_wait(peer_ar_flag(pa.base[pa.rank], parity, tid), pa.epoch, pa.err, pa.timeout, 1, tid);
    }
    __threadfence_system();
    __syncthreads();
    for (int i = tid; i < len; i += 1024)
    {
      double s = 0.0;
      for (int q = 0; q < pa.nranks; ++q)
        s += __ldcg(peer_ar_slot(pa.base[pa.rank], parity, q) + i);
      buf[i] = s;
    }
    __threadfence();
    __syncthreads();
  }

  static __global__ void __launch_bounds__(1024) peer_allreduce_kernel(const PeerArgs pa, double *__restrict__ buf, int len)
  {
    if (pa.done != nullptr && *pa.done != 0)
      return;
    peer_allreduce_body(pa, threadIdx.x, buf, len);
  }


  static __global__ void __launch_bounds__(256) halo_push_kernel(const PeerArgs pa, const HaloPushArgs h)
  {
    if (pa.done != nullptr && *pa.done != 0)
      return;
    const int parity = (int)(pa.epoch & 1ull);
    const int hp = h.m / 2;
    const long long total = h.send_off[h.npeers] * hp;
    const long long stride = (long long)gridDim.x * blockDim.x;
    auto locate = [&](long long e, const double2 *&src) -> double2 * {
      const long long s = e / hp;
      const int c = 2 * (int)(e % hp);
      int p = 0;
      while (s >= h.send_off[p + 1])
        ++p;
      src = reinterpret_cast<const double2 *>(h.X + (size_t)h.send_rows[s] * h.m + c);
      return reinterpret_cast<double2 *>(reinterpret_cast<double *>(pa.base[h.peer_rank[p]] + kPeerHaloOff + (size_t)parity * h.halo_cap_bytes) +
                                         (size_t)(h.deposit[p] + (s - h.send_off[p])) * h.m + c);
    };
    long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // four independent 16-byte loads in flight per thread before the peer stores. Measured: no effect (256^3, one neighbour,
    // 16.8 MB: 48 us = 350 GB/s with and without) -- the launch is bound by the NVLink store path, not by load latency.
    for (; e + 3 * stride < total; e += 4 * stride)
    {
      const double2 *src[4];
      double2 *dst[4];
      double2 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        dst[u] = locate(e + u * stride, src[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = __ldg(src[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *dst[u] = v[u];
    }
    for (; e < total; e += stride)
    {
      const double2 *src;
      double2 *dst = locate(e, src);
      *dst = __ldg(src);
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0)
      last = (atomicAdd(h.ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (last)
    {
      __threadfence_system();
      if (threadIdx.x < h.npeers)
        st_release_sys(peer_halo_flag(pa.base[h.peer_rank[threadIdx.x]], parity, pa.rank), pa.epoch);
      if (threadIdx.x == 0)
        *h.ticket = 0;
    }
  }


  /** returns when the halo rows of this epoch from every listed peer have landed in this rank's window */
  static __global__ void __launch_bounds__(32) halo_wait_kernel(const PeerArgs pa, const PeerList peers)
  {
    if (pa.done != nullptr && *pa.done != 0)
      return;
    const int parity = (int)(pa.epoch & 1ull);
    if ((int)threadIdx.x < peers.n)
      peer_wait(peer_halo_flag(pa.base[pa.rank], parity, peers.rank[threadIdx.x]), pa.epoch, pa.err, pa.timeout, 2, peers.rank[threadIdx.x]);
    __threadfence_system();
  }

} // namespace de
