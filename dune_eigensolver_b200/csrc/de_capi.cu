// libdune_eigensolver_b200.so -- C ABI (include/dune_eigensolver_b200.h) over the sm_100a kernels.
//
// Host-side runtime of the block-eigensolver hot path: contexts (stream, workspaces, NCCL communicator),
// device-resident matrices / multivectors / factor schedules, the kernel launch logic and the three
// device-resident driver loops (reference eigensolver.hh:28-112, :116-198, :204-351).
// No CPU fallback exists: every compute entry point needs a CUDA device and fails with DE_ERR_CUDA otherwise.

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <dlfcn.h>
#include <limits>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <unordered_map>
#include <random>
#include <string>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h> // types only; the library is bound at run time with dlopen (see NcclApi)

#define DE_KERNEL_MAX_M 64

#include "../../include/dune_eigensolver_b200.h"
#include "../../include/dune/eigensolver/sparse_lu.hh"
#include "kernels_dense.cuh"
#include "kernels_sparse.cuh"
#include "kernels_spmm_blocked.cuh"
#include "brb_format.hpp"
#include "kernels_brb_build.cuh"
#include "kernels_peer.cuh"
#include "kernels_tail.cuh"
#include "kernels_tallskinny.cuh"
#include "kernels_tallskinny2.cuh"
#include "kernels_trsv.cuh"
#include "kernels_lobpcg.cuh"
#include "lobpcg_core.hpp"

// ====================================================================================================
// error plumbing
// ====================================================================================================
namespace
{
  thread_local std::string g_thread_error;

  constexpr int kMaxPartials = 592;               // CTAs of a reduction kernel (4 per SM on 148 SMs)
  constexpr size_t kPartialDoubles = (size_t)2 * kMaxPartials * DE_KERNEL_MAX_M * DE_KERNEL_MAX_M;
  constexpr int kSmall = 5 * DE_KERNEL_MAX_M * DE_KERNEL_MAX_M + 1024; // doubles of small device / pinned scratch
}

struct NcclApi
{
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

static NcclApi &nccl_api()
{
  static NcclApi api;
  static bool tried = false;
  if (!tried)
  {
    tried = true;
    // a process that already loaded NCCL (torch) resolves to that copy through the soname
    api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (api.handle)
    {
      auto sym = [&](const char *n) { return dlsym(api.handle, n); };
      api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
      api.Send = (decltype(api.Send))sym("ncclSend");
      api.Recv = (decltype(api.Recv))sym("ncclRecv");
      api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
      api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
      api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.Send && api.Recv &&
               api.GroupStart && api.GroupEnd && api.GetErrorString;
    }
  }
  return api;
}

struct de_context
{
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_pack = nullptr, ev_halo = nullptr;
  int sm_count = 148;
  int rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  mutable std::string err;
  double *partials = nullptr; // kPartialDoubles
  double *dsmall = nullptr;   // kSmall doubles: G | Rinv | dp | info
  int *dstatus = nullptr;     // sticky Cholesky status
  int *dflags = nullptr;      // device flags: [0] second CholQR sweep not needed, [1] driver loop converged, [2] last iteration
  const int *done_ptr = nullptr; // = dflags + 1 while an asynchronous driver loop is enqueuing, else null
  int *hflags = nullptr;      // pinned: 2 slots x 4 ints, polled copies of dflags
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};
  double *dconv = nullptr;    // s_prev[64] | hist[dconv_cap]
  void *xfer = nullptr;       // XferEngine: pinned staging buffers and copy streams (created on first use)
  // NVLink peer window (kernels_peer.cuh); peer_ready once every rank's window is mapped
  bool peer_ready = false;
  unsigned char *window = nullptr;
  size_t window_bytes = 0, halo_cap = 0;
  unsigned char *peer_base[de::kPeerMaxRanks] = {};
  unsigned long long ar_epoch = 0, halo_epoch = 0;
  int *dticket = nullptr; // [0] ticket of halo_push_kernel, [1] peer error flag
  // fused tail of the NEXT partial-sum reduction (kernels_tail.cuh): set by the caller, consumed by reduce_partials
  de::TailArgs tail{};
  mutable bool tail_armed = false; // cleared by any error return (set_error), so a failed call cannot leave it behind
  mutable bool tail_did_allreduce = false, tail_did_op = false; // one-shot: the following allreduce_sum / chol / convergence is skipped
  int *dtail_ticket = nullptr;
  size_t dconv_cap = 0;
  double *hsmall = nullptr;   // pinned mirror of dsmall
  int *hstatus = nullptr;     // pinned
  double *stage = nullptr;    // layout-conversion staging
  size_t stage_bytes = 0;
  long long launches = 0;
  // optional per-kernel CUDA-event timing (bench.py's roofline numbers)
  bool profiling = false;
  unsigned prof_mask = ~0u; // categories that are timed while profiling is on
  struct ProfRecord
  {
    int cat;
    cudaEvent_t e0, e1;
  };
  std::vector<ProfRecord> prof_records;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[DE_PROF_CATEGORIES] = {0};
  long long prof_count[DE_PROF_CATEGORIES] = {0};

  double *dG() const { return dsmall; }
  double *dR() const { return dsmall + DE_KERNEL_MAX_M * DE_KERNEL_MAX_M; }
  double *dDP() const { return dsmall + 3 * DE_KERNEL_MAX_M * DE_KERNEL_MAX_M; }
  double *dInfo() const { return dsmall + 2 * DE_KERNEL_MAX_M * DE_KERNEL_MAX_M + 512; }
  // [dp (m) | G = Y^T Y (m x m)] of the last SpMM with dot (+ Gram) epilogue; dDP() aliases its head
  double *dDG() const { return dsmall + 3 * DE_KERNEL_MAX_M * DE_KERNEL_MAX_M; }
};

struct de_mv
{
  de_context *ctx;
  long long n;
  int m;
  double *d;
};

namespace
{
  void dev_free(void *p);
}

/** a set of rows prepared for spmm_staged_kernel: CSR (possibly a row-permuted copy) + row-block metadata */
struct StagedRows
{
  bool valid = false;
  bool owns_csr = false;
  int nblocks = 0;
  int *rowptr = nullptr, *col = nullptr, *rowmap = nullptr;
  double *val = nullptr;
  int4 *blk_meta = nullptr;
  void release()
  {
    if (owns_csr)
    {
      dev_free(rowptr);
      dev_free(col);
      dev_free(val);
    }
    dev_free(rowmap);
    dev_free(blk_meta);
    rowptr = col = rowmap = nullptr;
    val = nullptr;
    blk_meta = nullptr;
    valid = false;
  }
};

/** BRB form of a matrix on the device (brb_format.hpp): tiles [0, n_interior) touch owned columns only */
struct BrbDevice
{
  bool valid = false;
  int ntiles = 0, n_interior = 0, max_len16 = 0, max_u = 0;
  long long nblocks = 0, nsteps = 0, nvals = 0;
  bool grid = false;
  int tw = 0, th = 0, td = 0;
  int4 *tile = nullptr, *blob = nullptr;
  int *ucol = nullptr;
  size_t blob16 = 0, nucol = 0; // sizes of blob (16-byte units) and ucol
  void release()
  {
    dev_free(tile);
    dev_free(blob);
    dev_free(ucol);
    tile = blob = nullptr;
    ucol = nullptr;
    valid = false;
  }
};

struct de_matrix
{
  de_context *ctx;
  long long n = 0, n_halo = 0, nnz = 0;
  int *rowptr = nullptr, *col = nullptr;
  double *val = nullptr;
  // distributed part
  int npeers = 0;
  std::vector<int> peer;
  std::vector<long long> recv_count, recv_off, send_count, send_off;
  long long n_send = 0;
  int *send_rows = nullptr;
  int *interior = nullptr, *boundary = nullptr;
  long long n_interior = 0, n_boundary = 0;
  double *send_buf = nullptr, *halo_buf = nullptr;
  double *halo_view = nullptr; // where the kernels read halo rows of the current SpMM: halo_buf (NCCL) or the peer window
  int buf_m = 0;
  StagedRows st_all, st_interior, st_boundary;
  BrbDevice brb;
  bool peer_halo = false;            // halo rows travel as peer stores into the neighbours' windows
  std::vector<long long> deposit;    // [npeers] first row of this rank's rows in peer p's halo block
  long long halo_rows_max = 0;       // largest halo block over ALL ranks: peer path or NCCL must be the same decision everywhere
  int spmm_format = DE_SPMM_AUTO; // which SpMM kernel family to use (de_matrix_set_spmm_format)
};

struct TrsvSegment
{
  int chain;     // 1: chain kernel over levels [a,b) ; 0: single wide level a
  int a, b;
};

struct TrsvSchedule
{
  int *rows = nullptr, *rowptr = nullptr, *col = nullptr, *level_ptr = nullptr;
  double *val = nullptr, *invdiag = nullptr;
  std::vector<int> h_level_ptr;
  std::vector<TrsvSegment> segments;
  int nlevels = 0;
  long long nnz = 0;
};

struct de_factor
{
  de_context *ctx;
  long long n = 0, lnz = 0, unz = 0;
  TrsvSchedule L, U;
  int *P = nullptr, *Q = nullptr;
  double *rowscale = nullptr;
  double *W = nullptr;
  int W_m = 0;
  // the two triangular sweeps (hundreds of dependent launches on the fixed work block W) captured once per width
  cudaGraphExec_t sweep_graph = nullptr;
  int sweep_graph_m = 0;
  long long sweep_graph_nodes = 0;
};

struct de_host_factor
{
  de_b200::FactorArrays F;
};

namespace
{
  int set_error(const de_context *ctx, int code, const std::string &msg)
  {
    g_thread_error = msg;
    if (ctx)
    {
      ctx->err = msg;
      ctx->tail_armed = ctx->tail_did_allreduce = ctx->tail_did_op = false;
    }
    return code;
  }

  /** brackets one kernel launch with CUDA events on the launching stream when profiling is on */
  struct ProfScope
  {
    de_context *c;
    int cat;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    static cudaEvent_t take(de_context *c)
    {
      cudaEvent_t e = nullptr;
      if (!c->prof_pool.empty())
      {
        e = c->prof_pool.back();
        c->prof_pool.pop_back();
      }
      else
        cudaEventCreate(&e);
      return e;
    }
    ProfScope(de_context *ctx, int category) : c(ctx), cat(category)
    {
      if (c->profiling && ((c->prof_mask >> category) & 1u))
      {
        e0 = take(c);
        e1 = take(c);
        cudaEventRecord(e0, c->stream);
      }
    }
    ~ProfScope()
    {
      if (e0)
      {
        cudaEventRecord(e1, c->stream);
        c->prof_records.push_back(de_context::ProfRecord{cat, e0, e1});
      }
    }
  };

#define DE_CUDA(ctx, call)                                                                                   \
  do                                                                                                         \
  {                                                                                                          \
    cudaError_t e__ = (call);                                                                                \
    if (e__ != cudaSuccess)                                                                                  \
      return set_error(ctx, DE_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));               \
  } while (0)

#define DE_NCCL(ctx, call)                                                                                   \
  do                                                                                                         \
  {                                                                                                          \
    ncclResult_t r__ = (call);                                                                               \
    if (r__ != ncclSuccess)                                                                                  \
      return set_error(ctx, DE_ERR_NCCL, std::string(#call) + ": " + nccl_api().GetErrorString(r__));        \
  } while (0)

#define DE_TRY(call)                                                                                         \
  do                                                                                                         \
  {                                                                                                          \
    int s__ = (call);                                                                                        \
    if (s__ != DE_OK)                                                                                        \
      return s__;                                                                                            \
  } while (0)

#define DE_LAUNCH_CHECK(ctx)                                                                                 \
  do                                                                                                         \
  {                                                                                                          \
    (ctx)->launches++;                                                                                       \
    DE_CUDA(ctx, cudaGetLastError());                                                                        \
  } while (0)

  bool valid_cols(int m) { return m > 0 && m % 8 == 0 && m <= DE_MAX_COLS; }

  /** launch with programmatic stream serialization: the kernel's CTAs may be scheduled while the preceding kernel of the
   *  stream drains; the kernel itself waits for that kernel's completion in pdl_prologue() (kernels_sparse.cuh) */
  template <class... KArgs, class... Args>
  cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args)
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
  }

  // ---- device memory: a caching allocator ------------------------------------------------------------------
  // cudaMalloc / cudaFree take driver-wide locks; on the shared B200 boxes single calls were seen to stall for
  // 0.3-1 s (a 256 MB block allocated and freed per solve made one step in ten take 800 ms instead of 28). Blocks
  // released by the library are therefore kept, keyed by (device, stream, size), and handed out again; reuse on the
  // same stream is ordered after the previous user's kernels. Freed for real when the owning context is destroyed.
  struct DevBlockInfo
  {
    size_t bytes;
    int device;
    cudaStream_t stream;
  };
  struct DevCache
  {
    std::mutex mu;
    std::unordered_map<void *, DevBlockInfo> live;
    std::map<std::tuple<int, cudaStream_t, size_t>, std::vector<void *>> idle;
    size_t idle_bytes = 0;
  };
  DevCache &dev_cache()
  {
    static DevCache c;
    return c;
  }
  constexpr size_t kDevCacheMaxIdle = (size_t)96 << 30;

  int dev_alloc_bytes(de_context *ctx, void **p, size_t bytes)
  {
    *p = nullptr;
    bytes = std::max<size_t>((bytes + 511) & ~(size_t)511, 512);
    DevCache &C = dev_cache();
    {
      std::lock_guard<std::mutex> lock(C.mu);
      auto it = C.idle.find(std::make_tuple(ctx->device, ctx->stream, bytes));
      if (it != C.idle.end() && !it->second.empty())
      {
        *p = it->second.back();
        it->second.pop_back();
        C.idle_bytes -= bytes;
        C.live[*p] = DevBlockInfo{bytes, ctx->device, ctx->stream};
        return DE_OK;
      }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess)
    {
      // give the idle blocks back to the driver and try once more
      std::vector<void *> drop;
      {
        std::lock_guard<std::mutex> lock(C.mu);
        for (auto &kv : C.idle)
          if (std::get<0>(kv.first) == ctx->device)
          {
            for (void *q : kv.second)
              drop.push_back(q);
            C.idle_bytes -= std::get<2>(kv.first) * kv.second.size();
            kv.second.clear();
          }
      }
      cudaGetLastError();
      for (void *q : drop)
        cudaFree(q);
      e = cudaMalloc(p, bytes);
    }
    if (e != cudaSuccess)
      return set_error(ctx, e == cudaErrorMemoryAllocation ? DE_ERR_ALLOC : DE_ERR_CUDA,
                       std::string("cudaMalloc: ") + cudaGetErrorString(e));
    std::lock_guard<std::mutex> lock(C.mu);
    C.live[*p] = DevBlockInfo{bytes, ctx->device, ctx->stream};
    return DE_OK;
  }

  template <class T>
  int dev_alloc(de_context *ctx, T **p, size_t count)
  {
    return dev_alloc_bytes(ctx, (void **)p, std::max<size_t>(count, 1) * sizeof(T));
  }

  /** release a device pointer: blocks of this library go back to the cache, anything else to cudaFree */
  void dev_free(void *p)
  {
    if (!p)
      return;
    DevCache &C = dev_cache();
    {
      std::lock_guard<std::mutex> lock(C.mu);
      auto it = C.live.find(p);
      if (it != C.live.end())
      {
        const DevBlockInfo b = it->second;
        C.live.erase(it);
        if (C.idle_bytes + b.bytes <= kDevCacheMaxIdle)
        {
          C.idle[std::make_tuple(b.device, b.stream, b.bytes)].push_back(p);
          C.idle_bytes += b.bytes;
          return;
        }
      }
    }
    cudaFree(p);
  }

  /** really free the idle blocks of one (device, stream): context destruction */
  void dev_cache_trim(int device, cudaStream_t stream)
  {
    DevCache &C = dev_cache();
    std::vector<void *> drop;
    {
      std::lock_guard<std::mutex> lock(C.mu);
      for (auto &kv : C.idle)
        if (std::get<0>(kv.first) == device && std::get<1>(kv.first) == stream)
        {
          for (void *q : kv.second)
            drop.push_back(q);
          C.idle_bytes -= std::get<2>(kv.first) * kv.second.size();
          kv.second.clear();
        }
    }
    for (void *q : drop)
      cudaFree(q);
  }

  // ---- host <-> device transfers of caller (pageable) memory -------------------------------------------------
  // A plain cudaMemcpy from pageable memory is staged by the driver through one pinned buffer on one thread
  // (~10 GB/s); here kXferThreads (8) workers convert / copy 8 MB chunks into their own pinned buffers and issue
  // asynchronous copies on their own streams, so the PCIe link and several host cores work at the same time.
  constexpr int kXferThreads = 8;
  constexpr size_t kXferChunk = (size_t)8 << 20; // bytes per pinned buffer

  struct XferEngine
  {
    bool ready = false;
    unsigned char *pinned[kXferThreads][2] = {};
    cudaStream_t stream[kXferThreads] = {};
    cudaEvent_t ev[kXferThreads][2] = {};
  };

  int xfer_init(de_context *ctx);

  template <class T, class S>
  int upload_converted(de_context *ctx, T **dst, const S *src, size_t count)
  {
    std::vector<T> tmp(count);
    for (size_t i = 0; i < count; ++i)
      tmp[i] = (T)src[i];
    DE_TRY(dev_alloc(ctx, dst, count + 16 / sizeof(T))); // 16 bytes of tail padding: staged kernels copy 16-byte chunks
    DE_CUDA(ctx, cudaMemcpyAsync(*dst, tmp.data(), count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // tmp dies here
    return DE_OK;
  }

  int xfer_init(de_context *ctx)
  {
    if (ctx->xfer)
      return DE_OK;
    XferEngine *X = new XferEngine();
    for (int t = 0; t < kXferThreads; ++t)
    {
      DE_CUDA(ctx, cudaStreamCreateWithFlags(&X->stream[t], cudaStreamNonBlocking));
      for (int b = 0; b < 2; ++b)
      {
        DE_CUDA(ctx, cudaMallocHost((void **)&X->pinned[t][b], kXferChunk));
        DE_CUDA(ctx, cudaEventCreateWithFlags(&X->ev[t][b], cudaEventDisableTiming));
      }
    }
    X->ready = true;
    ctx->xfer = X;
    return DE_OK;
  }

  void xfer_destroy(de_context *ctx)
  {
    XferEngine *X = static_cast<XferEngine *>(ctx->xfer);
    if (!X)
      return;
    for (int t = 0; t < kXferThreads; ++t)
    {
      for (int b = 0; b < 2; ++b)
      {
        if (X->pinned[t][b])
          cudaFreeHost(X->pinned[t][b]);
        if (X->ev[t][b])
          cudaEventDestroy(X->ev[t][b]);
      }
      if (X->stream[t])
        cudaStreamDestroy(X->stream[t]);
    }
    delete X;
    ctx->xfer = nullptr;
  }

  /** dst[i] = (T) src[i], i < count, host -> device. `range` (optional) receives min and max of the source values.
   *  Work that ctx->stream has already queued on dst must be complete (callers upload into fresh allocations);
   *  on return the data is on the device. */
  template <class T, class S>
  int upload_parallel(de_context *ctx, T *dst, const S *src, size_t count, long long *range = nullptr)
  {
    if (range)
    {
      range[0] = 0;
      range[1] = -1;
    }
    if (count == 0)
      return DE_OK;
    DE_TRY(xfer_init(ctx));
    XferEngine *X = static_cast<XferEngine *>(ctx->xfer);
    const size_t per = kXferChunk / sizeof(T);
    const size_t nchunks = (count + per - 1) / per;
    const int nthreads = (int)std::min<size_t>(kXferThreads, nchunks);
    cudaError_t err[kXferThreads];
    long long lo[kXferThreads], hi[kXferThreads];
    auto work = [&](int t)
    {
      cudaSetDevice(ctx->device);
      err[t] = cudaSuccess;
      lo[t] = std::numeric_limits<long long>::max();
      hi[t] = std::numeric_limits<long long>::min();
      int b = 0;
      for (size_t c = (size_t)t; c < nchunks && err[t] == cudaSuccess; c += (size_t)nthreads, b ^= 1)
      {
        const size_t i0 = c * per, i1 = std::min(count, i0 + per);
        T *buf = reinterpret_cast<T *>(X->pinned[t][b]);
        cudaError_t e = cudaEventSynchronize(X->ev[t][b]); // the copy that last used this buffer
        if (e != cudaSuccess)
        {
          err[t] = e;
          break;
        }
        if (std::is_same<T, S>::value && !range)
          std::memcpy(buf, src + i0, (i1 - i0) * sizeof(T));
        else if (range)
        {
          long long l = lo[t], h = hi[t];
          for (size_t i = i0; i < i1; ++i)
          {
            const long long v = (long long)src[i];
            l = std::min(l, v);
            h = std::max(h, v);
            buf[i - i0] = (T)src[i];
          }
          lo[t] = l;
          hi[t] = h;
        }
        else
          for (size_t i = i0; i < i1; ++i)
            buf[i - i0] = (T)src[i];
        e = cudaMemcpyAsync(dst + i0, buf, (i1 - i0) * sizeof(T), cudaMemcpyHostToDevice, X->stream[t]);
        if (e == cudaSuccess)
          e = cudaEventRecord(X->ev[t][b], X->stream[t]);
        err[t] = e;
      }
      if (err[t] == cudaSuccess)
        err[t] = cudaStreamSynchronize(X->stream[t]);
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t)
      th.emplace_back(work, t);
    work(0);
    for (auto &x : th)
      x.join();
    for (int t = 0; t < nthreads; ++t)
    {
      if (err[t] != cudaSuccess)
        return set_error(ctx, DE_ERR_CUDA, std::string("host-to-device transfer: ") + cudaGetErrorString(err[t]));
      if (range && lo[t] <= hi[t])
      {
        if (range[0] > range[1])
        {
          range[0] = lo[t];
          range[1] = hi[t];
        }
        else
        {
          range[0] = std::min(range[0], lo[t]);
          range[1] = std::max(range[1], hi[t]);
        }
      }
    }
    return DE_OK;
  }

  /** device -> caller memory, `bytes` bytes; src must be complete on ctx->stream (the function synchronises it first) */
  int download_parallel(de_context *ctx, void *dst, const void *src, size_t bytes)
  {
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bytes == 0)
      return DE_OK;
    DE_TRY(xfer_init(ctx));
    XferEngine *X = static_cast<XferEngine *>(ctx->xfer);
    const size_t nchunks = (bytes + kXferChunk - 1) / kXferChunk;
    const int nthreads = (int)std::min<size_t>(kXferThreads, nchunks);
    cudaError_t err[kXferThreads];
    auto work = [&](int t)
    {
      cudaSetDevice(ctx->device);
      err[t] = cudaSuccess;
      // two chunks in flight per worker: issue c, then drain the previous one
      size_t prev = (size_t)-1;
      int b = 0;
      for (size_t c = (size_t)t; err[t] == cudaSuccess; c += (size_t)nthreads, b ^= 1)
      {
        if (c < nchunks)
        {
          const size_t o = c * kXferChunk, len = std::min(bytes - o, kXferChunk);
          cudaError_t e = cudaMemcpyAsync(X->pinned[t][b], (const unsigned char *)src + o, len, cudaMemcpyDeviceToHost, X->stream[t]);
          if (e == cudaSuccess)
            e = cudaEventRecord(X->ev[t][b], X->stream[t]);
          err[t] = e;
        }
        if (prev != (size_t)-1 && err[t] == cudaSuccess)
        {
          const size_t o = prev * kXferChunk, len = std::min(bytes - o, kXferChunk);
          err[t] = cudaEventSynchronize(X->ev[t][b ^ 1]);
          if (err[t] == cudaSuccess)
            std::memcpy((unsigned char *)dst + o, X->pinned[t][b ^ 1], len);
        }
        if (c >= nchunks)
          break;
        prev = c;
      }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t)
      th.emplace_back(work, t);
    work(0);
    for (auto &x : th)
      x.join();
    for (int t = 0; t < nthreads; ++t)
      if (err[t] != cudaSuccess)
        return set_error(ctx, DE_ERR_CUDA, std::string("device-to-host transfer: ") + cudaGetErrorString(err[t]));
    return DE_OK;
  }

  int bind_device(const de_context *ctx)
  {
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
    return DE_OK;
  }

  // ---- reductions -----------------------------------------------------------------------------------
  de::PeerArgs peer_args(de_context *ctx, unsigned long long epoch);

  int reduce_partials(de_context *ctx, const double *partials, int nparts, int len, double *out)
  {
    dim3 block(32, 32);
    ProfScope prof(ctx, DE_PROF_SMALL);
    const bool multi = ctx->nranks > 1;
    if (ctx->tail_armed && ctx->dtail_ticket && (!multi || (ctx->peer_ready && len <= de::kPeerSlotDoubles)))
    {
      // reduce -> (all-reduce) -> Cholesky / convergence test in ONE launch
      de::TailArgs t = ctx->tail;
      t.do_allreduce = multi ? 1 : 0;
      if (multi)
        t.pa = peer_args(ctx, ++ctx->ar_epoch);
      t.ticket = ctx->dtail_ticket;
      ctx->tail_armed = false;
      ctx->tail_did_allreduce = true;
      ctx->tail_did_op = true;
      DE_CUDA(ctx, launch_pdl(de::reduce_tail_kernel, dim3((len + 31) / 32), block, 0, ctx->stream, partials, nparts, len, out,
                              ctx->done_ptr, t));
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    ctx->tail_armed = false;
    de::reduce_partials_kernel<<<(len + 31) / 32, block, 0, ctx->stream>>>(partials, nparts, len, out, ctx->done_ptr);
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  de::PeerArgs peer_args(de_context *ctx, unsigned long long epoch)
  {
    de::PeerArgs pa;
    pa.rank = ctx->rank;
    pa.nranks = ctx->nranks;
    for (int q = 0; q < de::kPeerMaxRanks; ++q)
      pa.base[q] = ctx->peer_base[q];
    pa.epoch = epoch;
    pa.done = ctx->done_ptr;
    pa.err = ctx->dticket + 1;
    return pa;
  }

  int allreduce_sum(de_context *ctx, double *buf, size_t count)
  {
    if (ctx->tail_did_allreduce)
    {
      ctx->tail_did_allreduce = false; // the fused tail of the reduction already did it
      return DE_OK;
    }
    if (ctx->nranks <= 1)
      return DE_OK;
    if (ctx->peer_ready && count <= (size_t)de::kPeerSlotDoubles)
    {
      ProfScope prof(ctx, DE_PROF_SMALL);
      de::peer_allreduce_kernel<<<1, 1024, 0, ctx->stream>>>(peer_args(ctx, ++ctx->ar_epoch), buf, (int)count);
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    DE_NCCL(ctx, nccl_api().AllReduce(buf, buf, count, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    return DE_OK;
  }

  // ---- SpMM -----------------------------------------------------------------------------------------
  template <bool DOT>
  int launch_spmm_rows(de_context *ctx, const de_matrix *A, const double *X, double *Y, int m, const int *rowlist,
                       long long nrows, double *partials, int *grid_out)
  {
    *grid_out = 0;
    if (nrows <= 0)
      return DE_OK;
    de::SpmmArgs a;
    a.nrows = nrows;
    a.rowlist = rowlist;
    a.rowptr = A->rowptr;
    a.col = A->col;
    a.val = A->val;
    a.X = X;
    a.H = A->halo_view;
    a.n_owned = A->n;
    a.ld = m;
    a.m = m;
    a.Y = Y;
    a.partials = partials;
    a.done = ctx->done_ptr;
    const int hp = m / 2;
    const int tpr = hp <= 4 ? 4 : (hp <= 8 ? 8 : (hp <= 16 ? 16 : 32));
    const int rpb = 256 / tpr;
    const long long need = (nrows + rpb - 1) / rpb;
    const int cap = DOT ? kMaxPartials : ctx->sm_count * 8;
    int grid = (int)std::min<long long>(need, cap);
    ProfScope prof(ctx, DE_PROF_SPMM);
    const bool exact = (m == 2 * tpr) && ((A->n + A->n_halo) * (long long)(m / 2) < (1LL << 31));
    const bool halo = A->n_halo > 0;
    if (exact)
    {
      // spmm_kernel_v2 is compiled for 3 resident CTAs per SM: launch exactly one wave of the grid-stride loop
      grid = (int)std::min<long long>(need, (long long)ctx->sm_count * 3);
#define DE_SPMM_V2(T)                                                                                        \
  if (halo)                                                                                                  \
    de::spmm_kernel_v2<T, DOT, true><<<grid, 256, 0, ctx->stream>>>(a);                                      \
  else                                                                                                       \
    de::spmm_kernel_v2<T, DOT, false><<<grid, 256, 0, ctx->stream>>>(a);
      switch (tpr)
      {
      case 4:
        DE_SPMM_V2(4)
        break;
      case 8:
        DE_SPMM_V2(8)
        break;
      case 16:
        DE_SPMM_V2(16)
        break;
      default:
        DE_SPMM_V2(32)
        break;
      }
#undef DE_SPMM_V2
    }
    else
      switch (tpr)
      {
      case 4:
        de::spmm_kernel<4, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      case 8:
        de::spmm_kernel<8, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      case 16:
        de::spmm_kernel<16, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      default:
        de::spmm_kernel<32, 1, DOT><<<grid, 256, 0, ctx->stream>>>(a);
        break;
      }
    DE_LAUNCH_CHECK(ctx);
    *grid_out = grid;
    return DE_OK;
  }

  /** cut rows 0..nrows of a (host) row pointer into blocks of <= 256 rows and <= 2048 nonzeros */
  template <class Ptr>
  std::vector<int4> cut_row_blocks(long long nrows, const Ptr *rowptr)
  {
    std::vector<int4> meta;
    long long r0 = 0;
    while (r0 < nrows)
    {
      long long r1 = r0;
      while (r1 < nrows && r1 - r0 < de::kStageMaxRows && (long long)(rowptr[r1 + 1] - rowptr[r0]) <= de::kStageCapNnz)
        ++r1;
      if (r1 == r0)
        r1 = r0 + 1; // a single row longer than the staging capacity: read directly from global memory
      meta.push_back(make_int4((int)r0, (int)r1, (int)rowptr[r0], (int)rowptr[r1]));
      r0 = r1;
    }
    return meta;
  }

  /** staged view of the whole matrix (shares the CSR arrays already on the device) */
  int build_staged_all(de_context *ctx, de_matrix *A, const int64_t *rowptr)
  {
    std::vector<int4> meta = cut_row_blocks(A->n, rowptr);
    StagedRows &S = A->st_all;
    S.rowptr = A->rowptr;
    S.col = A->col;
    S.val = A->val;
    S.owns_csr = false;
    S.nblocks = (int)meta.size();
    DE_TRY(dev_alloc(ctx, &S.blk_meta, meta.size()));
    DE_CUDA(ctx, cudaMemcpyAsync(S.blk_meta, meta.data(), meta.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S.valid = true;
    return DE_OK;
  }

  /** staged copy of the rows in `list` (interior or boundary rows of a distributed matrix), permuted to be contiguous */
  int build_staged_subset(de_context *ctx, const std::vector<int> &list, const int64_t *rowptr, const int64_t *col,
                          const double *val, StagedRows &S)
  {
    S.valid = false;
    if (list.empty())
      return DE_OK;
    std::vector<int> ptr(list.size() + 1, 0), c;
    std::vector<double> v;
    for (size_t i = 0; i < list.size(); ++i)
    {
      const int r = list[i];
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
      {
        c.push_back((int)col[k]);
        v.push_back(val[k]);
      }
      ptr[i + 1] = (int)c.size();
    }
    std::vector<int4> meta = cut_row_blocks((long long)list.size(), ptr.data());
    S.owns_csr = true;
    S.nblocks = (int)meta.size();
    DE_TRY(upload_converted(ctx, &S.rowptr, ptr.data(), ptr.size()));
    DE_TRY(upload_converted(ctx, &S.col, c.data(), c.size()));
    DE_TRY(upload_converted(ctx, &S.val, v.data(), v.size()));
    DE_TRY(upload_converted(ctx, &S.rowmap, list.data(), list.size()));
    DE_TRY(dev_alloc(ctx, &S.blk_meta, meta.size()));
    DE_CUDA(ctx, cudaMemcpyAsync(S.blk_meta, meta.data(), meta.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    S.valid = true;
    return DE_OK;
  }

  inline bool staged_usable(const de_matrix *A, int m)
  {
    return (m == 8 || m == 16 || m == 32 || m == 64) && (A->n + A->n_halo) * (long long)(m / 2) < (1LL << 31);
  }

  template <bool DOT>
  int launch_spmm_staged(de_context *ctx, const de_matrix *A, const StagedRows &S, const double *X, double *Y, int m,
                         double *partials, int *grid_out)
  {
    *grid_out = 0;
    if (!S.valid || S.nblocks <= 0)
      return DE_OK;
    de::StagedArgs a;
    a.nblocks = S.nblocks;
    a.blk_meta = S.blk_meta;
    a.rowmap = S.rowmap;
    a.rowptr = S.rowptr;
    a.col = S.col;
    a.val = S.val;
    a.X = X;
    a.H = A->halo_view;
    a.n_owned = (int)A->n;
    a.m = m;
    a.Y = Y;
    a.partials = partials;
    a.done = ctx->done_ptr;
    constexpr size_t smem = de::spmm_staged_smem_bytes();
    const int grid = std::min(S.nblocks, ctx->sm_count * 3); // 3 resident CTAs per SM
    const bool halo = A->n_halo > 0;
    const int tpr = m / 2;
    ProfScope prof(ctx, DE_PROF_SPMM);
#define DE_SPMM_ST(T)                                                                                        \
  {                                                                                                          \
    static bool cfg = false;                                                                                 \
    if (!cfg)                                                                                                \
    {                                                                                                        \
      DE_CUDA(ctx, cudaFuncSetAttribute(de::spmm_staged_kernel<T, DOT, true>,                                \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
      DE_CUDA(ctx, cudaFuncSetAttribute(de::spmm_staged_kernel<T, DOT, false>,                               \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
      cfg = true;                                                                                            \
    }                                                                                                        \
    if (halo)                                                                                                \
      de::spmm_staged_kernel<T, DOT, true><<<grid, 256, smem, ctx->stream>>>(a);                             \
    else                                                                                                     \
      de::spmm_staged_kernel<T, DOT, false><<<grid, 256, smem, ctx->stream>>>(a);                            \
  }
    switch (tpr)
    {
    case 4:
      DE_SPMM_ST(4)
      break;
    case 8:
      DE_SPMM_ST(8)
      break;
    case 16:
      DE_SPMM_ST(16)
      break;
    default:
      DE_SPMM_ST(32)
      break;
    }
#undef DE_SPMM_ST
    DE_LAUNCH_CHECK(ctx);
    *grid_out = grid;
    return DE_OK;
  }

  /** BRB form of the matrix. The host only plans which rows form the tiles (brb::plan: pattern detection on a few
   *  sample rows + O(n) index arithmetic); the blobs are built on the device from the CSR arrays already uploaded
   *  (kernels_brb_build.cuh: COUNT pass -> offsets on the host -> FILL pass). A matrix the format cannot represent
   *  (or an empty one) simply keeps the CSR kernels: not an error. */
  int build_brb(de_context *ctx, de_matrix *A, long long n, long long ncols, const int64_t *rowptr, const int64_t *col,
                const double *val)
  {
    BrbDevice &B = A->brb;
    B.release();
    if (n <= 0 || rowptr[n] <= 0)
      return DE_OK;
    static_assert(sizeof(de::brb::TileDesc) == sizeof(int4), "tile descriptors are loaded as int4");
    de::brb::Plan P;
    int first = 0;
    int *d_rows = nullptr, *d_cut = nullptr;
    int4 *d_info = nullptr, *d_place = nullptr;
    auto cleanup = [&]()
    {
      dev_free(d_rows);
      dev_free(d_cut);
      dev_free(d_info);
      dev_free(d_place);
      d_rows = d_cut = nullptr;
      d_info = d_place = nullptr;
    };
    struct Scope
    {
      decltype(cleanup) &f;
      ~Scope() { f(); }
    } scope{cleanup};
    while (de::brb::plan(n, ncols, rowptr, col, val, n, first, P))
    {
      first = P.next;
      cleanup();
      const int ntiles = (int)P.order.tilecut.size() - 1;
      if (ntiles <= 0)
        continue;
      DE_TRY(dev_alloc(ctx, &d_rows, P.order.rows.size()));
      DE_TRY(dev_alloc(ctx, &d_cut, P.order.tilecut.size()));
      DE_TRY(dev_alloc(ctx, &d_info, (size_t)ntiles));
      DE_TRY(dev_alloc(ctx, &d_place, (size_t)ntiles));
      DE_CUDA(ctx, cudaMemcpyAsync(d_rows, P.order.rows.data(), P.order.rows.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      DE_CUDA(ctx, cudaMemcpyAsync(d_cut, P.order.tilecut.data(), P.order.tilecut.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
      de::BrbBuildArgs a{};
      a.ntiles = ntiles;
      a.tilecut = d_cut;
      a.rows = d_rows;
      a.rowptr = A->rowptr;
      a.col = A->col;
      a.val = A->val;
      a.n_owned = (int)n;
      a.info = d_info;
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        de::brb_build_kernel<false><<<ntiles, de::kBldThreads, 0, ctx->stream>>>(a);
      }
      DE_LAUNCH_CHECK(ctx);
      std::vector<int4> info((size_t)ntiles);
      DE_CUDA(ctx, cudaMemcpyAsync(info.data(), d_info, info.size() * sizeof(int4), cudaMemcpyDeviceToHost, ctx->stream));
      DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      // sizes -> placement (tile order of the plan); the SpMM kernel sees the tiles interior-first
      std::vector<int4> place((size_t)ntiles);
      size_t blob16 = 0, ucols = 0;
      int max_len16 = 0, max_u = 0;
      long long nsteps = 0, nvals = 0;
      bool ok = true;
      for (int t = 0; t < ntiles && ok; ++t)
      {
        const int nb = P.order.tilecut[t + 1] - P.order.tilecut[t];
        const int4 f = info[t];
        if (f.w & 2)
        {
          ok = false;
          break;
        }
        const size_t words = (((size_t)4 + (nb + 1) + 12 * (size_t)nb + 3) & ~(size_t)3) + 4 * (size_t)f.x + (((size_t)2 * f.y + 3) & ~(size_t)3);
        const int len16 = (int)(words / 4);
        place[t] = make_int4((int)blob16, len16, (int)ucols, f.z);
        blob16 += (size_t)len16;
        ucols += (size_t)f.z;
        max_len16 = std::max(max_len16, len16);
        max_u = std::max(max_u, f.z);
        nsteps += f.x;
        nvals += f.y;
        if (blob16 >= ((size_t)1 << 31) || ucols >= ((size_t)1 << 31))
          ok = false;
      }
      if (!ok || !de::brb::fits_budget(max_len16, max_u))
        continue; // next candidate (smaller tiles / consecutive rows)
      DE_TRY(dev_alloc(ctx, &B.blob, blob16 + 4));
      DE_TRY(dev_alloc(ctx, &B.ucol, ucols + 64));
      DE_TRY(dev_alloc(ctx, &B.tile, (size_t)ntiles));
      DE_CUDA(ctx, cudaMemsetAsync(B.blob, 0, (blob16 + 4) * sizeof(int4), ctx->stream));
      DE_CUDA(ctx, cudaMemcpyAsync(d_place, place.data(), place.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
      a.place = d_place;
      a.blob = reinterpret_cast<int *>(B.blob);
      a.ucol = B.ucol;
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        de::brb_build_kernel<true><<<ntiles, de::kBldThreads, 0, ctx->stream>>>(a);
      }
      DE_LAUNCH_CHECK(ctx);
      std::vector<int4> tiles;
      tiles.reserve((size_t)ntiles);
      int n_interior = 0;
      for (int pass = 0; pass < 2; ++pass)
      {
        for (int t = 0; t < ntiles; ++t)
          if (((info[t].w & 1) != 0) == (pass == 1))
            tiles.push_back(place[t]);
        if (pass == 0)
          n_interior = (int)tiles.size();
      }
      DE_CUDA(ctx, cudaMemcpyAsync(B.tile, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, ctx->stream));
      DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // host vectors die here
      B.ntiles = ntiles;
      B.n_interior = n_interior;
      B.max_len16 = max_len16;
      B.max_u = max_u;
      B.nblocks = (long long)(P.order.rows.size() / 8);
      B.nsteps = nsteps;
      B.nvals = nvals;
      B.grid = P.grid;
      B.tw = P.tw;
      B.th = P.th;
      B.td = P.td;
      B.blob16 = blob16;
      B.nucol = ucols;
      B.valid = true;
      return DE_OK;
    }
    return DE_OK;
  }

  /** AUTO policy (measured, profiles/r01_spmm_lab_*.log): the BRB kernel wins for long rows at every width (27-point:
   *  1.5-2.4x) and for short rows from m = 32 up; narrow blocks on a 7-point matrix leave the tensor-core steps mostly
   *  zero-filled and the CSR kernel is faster there. */
  inline bool brb_usable(const de_matrix *A, int m)
  {
    if (!A->brb.valid || A->spmm_format == DE_SPMM_CSR || m % 8 != 0 || m < 8 || m > DE_MAX_COLS)
      return false;
    if (A->spmm_format == DE_SPMM_BRB)
      return true;
    return m >= 32 || A->nnz >= 12 * A->n;
  }

  constexpr int kBrbMaxSmem = 227 * 1024;

  template <int NP, bool DOT, bool HALO, bool GRAM>
  int launch_brb_pass(de_context *ctx, const de::BrbArgs &a, int grid)
  {
    static bool cfg = false; // one per instantiation
    if (!cfg)
    {
      DE_CUDA(ctx, cudaFuncSetAttribute(de::spmm_brb_kernel<NP, DOT, HALO, GRAM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kBrbMaxSmem));
      cfg = true;
    }
    const size_t smem = de::spmm_brb_smem_bytes(NP, a.blob_cap16, a.xs_cap, a.stages);
    DE_CUDA(ctx, launch_pdl(de::spmm_brb_kernel<NP, DOT, HALO, GRAM>, dim3(grid), dim3(de::brb_threads(GRAM)), smem, ctx->stream, a));
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  /** The drivers could take G = Y^T Y from the SpMM epilogue instead of a separate Gram pass (de_spmm_gram does).
   *  Measured on B200 (100^3 Q1, m = 32): the epilogue adds 0.12 ms to a 0.19 ms SpMM (20 more DMMA per row block
   *  on accumulators that serialise, 40 more registers, 10 instead of 12 consumer warps), the separate pass costs
   *  0.056 ms + one launch: the solve went from 22.4 to 25.9 ms. Off until the epilogue is cheaper. */
  constexpr bool kFuseGramIntoSpmm = false;

  /** can the Gram matrix Y^T Y be accumulated in the SpMM epilogue? (single pass over the columns) */
  inline bool brb_gram_epilogue(int m) { return m == 8 || m == 16 || m == 32; }

  /** Y = A X on tiles [t0, t0 + nt) of the BRB form. m columns are covered by passes of 32 / 16 / 8 columns (one
   *  kernel launch each; every pass re-streams the matrix blobs, so m = 64 costs two passes).
   *  DOT: per-CTA partials of diag(X^T Y) at partials[cta * pstride + j]; GRAM (needs brb_gram_epilogue(m)):
   *  additionally the CTA's partial of Y^T Y at partials[cta * pstride + m + i * m + j], pstride = m + m * m. */
  template <bool DOT, bool GRAM>
  int launch_spmm_brb(de_context *ctx, const de_matrix *A, int t0, int nt, const double *X, double *Y, int m, double *partials,
                      int *grid_out, int category = DE_PROF_SPMM)
  {
    *grid_out = 0;
    if (nt <= 0)
      return DE_OK;
    const BrbDevice &B = A->brb;
    const int grid = std::min(nt, ctx->sm_count);
    const bool halo = A->n_halo > 0;
    ProfScope prof(ctx, category);
    for (int c0 = 0; c0 < m;)
    {
      const int left = (m - c0) / 8;
      const int np = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
      de::BrbArgs a;
      a.ntiles = nt;
      a.n = A->n;
      a.tile = B.tile + t0;
      a.blob = B.blob;
      a.ucol = B.ucol;
      a.X = X + c0;
      a.H = A->halo_view ? A->halo_view + c0 : nullptr;
      a.n_owned = (int)A->n;
      a.ldx = m;
      a.Y = Y + c0;
      a.partials = partials ? partials + c0 : nullptr;
      a.pstride = GRAM ? m + m * m : m;
      a.gram_off = m;
      a.blob_cap16 = B.max_len16;
      a.xs_cap = B.max_u;
      a.done = ctx->done_ptr;
      int stages = de::kBrbMaxStages;
      while (stages > 2 && de::spmm_brb_smem_bytes(np, a.blob_cap16, a.xs_cap, stages) > (size_t)kBrbMaxSmem)
        --stages;
      a.stages = stages;
      if (de::spmm_brb_smem_bytes(np, a.blob_cap16, a.xs_cap, stages) > (size_t)kBrbMaxSmem)
        return set_error(ctx, DE_ERR_UNSUPPORTED, "BRB tile does not fit in shared memory");
#define DE_BRB(NPV)                                                             \
  {                                                                             \
    if (halo)                                                                   \
      DE_TRY((launch_brb_pass<NPV, DOT, true, GRAM>(ctx, a, grid)));            \
    else                                                                        \
      DE_TRY((launch_brb_pass<NPV, DOT, false, GRAM>(ctx, a, grid)));           \
  }
      if (np == 4)
        DE_BRB(4)
      else if (np == 2)
        DE_BRB(2)
      else
        DE_BRB(1)
#undef DE_BRB
      c0 += 8 * np;
    }
    *grid_out = grid;
    return DE_OK;
  }

  int ensure_halo_buffers(de_context *ctx, de_matrix *A, int m)
  {
    if (A->buf_m >= m)
      return DE_OK;
    if (A->send_buf)
      dev_free(A->send_buf);
    if (A->halo_buf)
      dev_free(A->halo_buf);
    A->send_buf = A->halo_buf = nullptr;
    DE_TRY(dev_alloc(ctx, &A->send_buf, (size_t)A->n_send * m));
    DE_TRY(dev_alloc(ctx, &A->halo_buf, (size_t)A->n_halo * m));
    A->buf_m = m;
    return DE_OK;
  }

  /** Y = A X (+ dp = diag(X^T Y) into ctx->dDP when DOT). Distributed matrices first start the halo exchange
   *  (pack -> NCCL send/recv over NVLink on the communication stream), run the interior rows meanwhile, then the
   *  boundary rows once the halo rows have landed. */
  template <bool DOT>
  int spmm_device(de_context *ctx, const de_matrix *Ac, const double *X, double *Y, int m, bool *gram_out = nullptr)
  {
    de_matrix *A = const_cast<de_matrix *>(Ac);
    int g1 = 0, g2 = 0;
    const bool dist = ctx->nranks > 1 && (A->n_halo > 0 || A->n_send > 0);
    // peer-store halo exchange? decided from quantities that are equal on all ranks; the epoch advances on every rank
    // of the job, also on one that has neither halo rows nor rows to send for this matrix
    const bool peer_path = ctx->nranks > 1 && ctx->peer_ready && A->peer_halo &&
                           (size_t)A->halo_rows_max * m * sizeof(double) <= ctx->halo_cap && A->npeers <= de::kPeerMaxRanks;
    const unsigned long long halo_epoch = peer_path ? ++ctx->halo_epoch : 0ull;
    // GRAM epilogue: the caller can use G = Y^T Y (in ctx->dDG() + m); only the tensor-core kernel has it
    const bool gram = DOT && gram_out != nullptr && brb_usable(A, m) && brb_gram_epilogue(m);
    if (gram_out)
      *gram_out = gram;
    if (!dist)
    {
      if (brb_usable(A, m))
      {
        if (gram)
          DE_TRY((launch_spmm_brb<DOT, DOT>(ctx, A, 0, A->brb.ntiles, X, Y, m, ctx->partials, &g1)));
        else
          DE_TRY((launch_spmm_brb<DOT, false>(ctx, A, 0, A->brb.ntiles, X, Y, m, ctx->partials, &g1)));
      }
      else if (A->st_all.valid && staged_usable(A, m))
        DE_TRY(launch_spmm_staged<DOT>(ctx, A, A->st_all, X, Y, m, ctx->partials, &g1));
      else
        DE_TRY(launch_spmm_rows<DOT>(ctx, A, X, Y, m, nullptr, A->n, ctx->partials, &g1));
    }
    else
    {
      const bool peer = peer_path;
      de::PeerArgs pa{};
      if (peer)
      {
        // halo rows go straight into the neighbours' windows; the boundary tiles wait for this epoch's flags
        pa = peer_args(ctx, halo_epoch);
        A->halo_view = reinterpret_cast<double *>(ctx->window + de::kPeerHaloOff + (size_t)(pa.epoch & 1ull) * ctx->halo_cap);
        if (A->n_send > 0)
        {
          de::HaloPushArgs h{};
          h.npeers = A->npeers;
          for (int p = 0; p < A->npeers; ++p)
          {
            h.peer_rank[p] = A->peer[p];
            h.send_off[p] = A->send_off[p];
            h.deposit[p] = A->deposit[p];
          }
          h.send_off[A->npeers] = A->n_send;
          h.send_rows = A->send_rows;
          h.X = X;
          h.m = m;
          h.halo_cap_bytes = ctx->halo_cap;
          h.ticket = ctx->dticket;
          const long long total = A->n_send * (m / 2);
          const int grid = (int)std::min<long long>((total + 255) / 256, ctx->sm_count * 4);
          ProfScope prof(ctx, DE_PROF_MISC);
          de::halo_push_kernel<<<grid, 256, 0, ctx->stream>>>(pa, h);
          DE_LAUNCH_CHECK(ctx);
        }
      }
      else
      {
        DE_TRY(ensure_halo_buffers(ctx, A, m));
        A->halo_view = A->halo_buf;
        if (A->n_send > 0)
        {
          const long long total = A->n_send * (m / 2);
          const int grid = (int)std::min<long long>((total + 255) / 256, ctx->sm_count * 8);
          ProfScope prof(ctx, DE_PROF_MISC);
          de::pack_rows_kernel<<<grid, 256, 0, ctx->stream>>>(A->n_send, A->send_rows, m, X, A->send_buf);
          DE_LAUNCH_CHECK(ctx);
        }
        DE_CUDA(ctx, cudaEventRecord(ctx->ev_pack, ctx->stream));
        DE_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_pack, 0));
        NcclApi &nc = nccl_api();
        DE_NCCL(ctx, nc.GroupStart());
        for (int p = 0; p < A->npeers; ++p)
        {
          if (A->send_count[p] > 0)
            DE_NCCL(ctx, nc.Send(A->send_buf + (size_t)A->send_off[p] * m, (size_t)A->send_count[p] * m, ncclDouble,
                                 A->peer[p], ctx->comm, ctx->comm_stream));
          if (A->recv_count[p] > 0)
            DE_NCCL(ctx, nc.Recv(A->halo_buf + (size_t)A->recv_off[p] * m, (size_t)A->recv_count[p] * m, ncclDouble,
                                 A->peer[p], ctx->comm, ctx->comm_stream));
        }
        DE_NCCL(ctx, nc.GroupEnd());
        DE_CUDA(ctx, cudaEventRecord(ctx->ev_halo, ctx->comm_stream));
      }
      const bool staged = staged_usable(A, m) && (A->st_interior.valid || A->n_interior == 0) &&
                          (A->st_boundary.valid || A->n_boundary == 0);
      const bool brb = brb_usable(A, m);
      const size_t pstride = gram ? (size_t)m + (size_t)m * m : (size_t)m;
      if (brb)
      {
        if (gram)
          DE_TRY((launch_spmm_brb<DOT, DOT>(ctx, A, 0, A->brb.n_interior, X, Y, m, ctx->partials, &g1)));
        else
          DE_TRY((launch_spmm_brb<DOT, false>(ctx, A, 0, A->brb.n_interior, X, Y, m, ctx->partials, &g1)));
      }
      else if (staged)
        DE_TRY(launch_spmm_staged<DOT>(ctx, A, A->st_interior, X, Y, m, ctx->partials, &g1));
      else
        DE_TRY(launch_spmm_rows<DOT>(ctx, A, X, Y, m, A->interior, A->n_interior, ctx->partials, &g1));
      if (peer)
      {
        de::PeerList pl{};
        for (int p = 0; p < A->npeers; ++p)
          if (A->recv_count[p] > 0)
            pl.rank[pl.n++] = A->peer[p];
        ProfScope prof(ctx, DE_PROF_MISC);
        de::halo_wait_kernel<<<1, 32, 0, ctx->stream>>>(pa, pl);
        DE_LAUNCH_CHECK(ctx);
      }
      else
        DE_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
      if (brb)
      {
        if (gram)
          DE_TRY((launch_spmm_brb<DOT, DOT>(ctx, A, A->brb.n_interior, A->brb.ntiles - A->brb.n_interior, X, Y, m,
                                            ctx->partials + (size_t)g1 * pstride, &g2, DE_PROF_SPMM_BOUNDARY)));
        else
          DE_TRY((launch_spmm_brb<DOT, false>(ctx, A, A->brb.n_interior, A->brb.ntiles - A->brb.n_interior, X, Y, m,
                                              ctx->partials + (size_t)g1 * pstride, &g2, DE_PROF_SPMM_BOUNDARY)));
      }
      else if (staged)
        DE_TRY(launch_spmm_staged<DOT>(ctx, A, A->st_boundary, X, Y, m, ctx->partials + (size_t)g1 * m, &g2));
      else
        DE_TRY(launch_spmm_rows<DOT>(ctx, A, X, Y, m, A->boundary, A->n_boundary, ctx->partials + (size_t)g1 * m, &g2));
    }
    if (DOT)
    {
      // dp (and G = Y^T Y when the Gram epilogue ran) are reduced, and all-reduced, as ONE vector dDG = [dp | G]
      const int len = gram ? m + m * m : m;
      if (g1 + g2 > 0)
        DE_TRY(reduce_partials(ctx, ctx->partials, g1 + g2, len, ctx->dDG()));
      else
        DE_CUDA(ctx, cudaMemsetAsync(ctx->dDG(), 0, sizeof(double) * len, ctx->stream));
      DE_TRY(allreduce_sum(ctx, ctx->dDG(), (size_t)len));
    }
    return DE_OK;
  }

  // ---- diag-dot ---------------------------------------------------------------------------------------
  int diag_dot_device(de_context *ctx, long long n, int m, const double *X, const double *Y, double *out)
  {
    const int hp = m / 2;
    dim3 block(hp, 256 / hp);
    const long long need = (n + block.y - 1) / block.y;
    const int grid = (int)std::max<long long>(1, std::min<long long>(need, kMaxPartials));
    {
      ProfScope prof(ctx, DE_PROF_DOT);
      de::diag_dot_kernel<<<grid, block, 0, ctx->stream>>>(n, X, m, Y, m, m, ctx->partials);
    }
    DE_LAUNCH_CHECK(ctx);
    DE_TRY(reduce_partials(ctx, ctx->partials, grid, m, out));
    return allreduce_sum(ctx, out, m);
  }

  // ---- pipelined tall-skinny kernel (m = 8/16/32/64) -------------------------------------------------------
  inline bool ts_supported(int w) { return w == 8 || w == 16 || w == 32 || w == 64; }

  template <int M, bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  int launch_ts_t(de_context *ctx, de::TsArgs a, double *gram_out)
  {
    if constexpr (DO_UPDATE && M == 64 && DO_GRAM && UPPER && SAME)
    {
      // no fused kernel at this width: block update, then the Gram matrix of the result (two passes, both on the
      // warp-specialised kernels; the one-pass first-generation kernel at M = 64 is slower than the two together)
      de::TsArgs u = a;
      DE_TRY((launch_ts_t<M, true, false, false, true>(ctx, u, nullptr)));
      de::TsArgs g = a;
      g.X = a.Out;
      g.ldx = a.ldo;
      g.skip_flag = nullptr;
      return launch_ts_t<M, false, true, true, true>(ctx, g, gram_out);
    }
    if constexpr (DO_UPDATE && (M <= 32 || !DO_GRAM) && (!DO_GRAM || (UPPER && SAME)))
    {
      // block update (+ Gram of the result): the register-to-register tensor-core kernel (kernels_tallskinny2.cuh)
      using C2 = de::Ts2Cfg<M>;
      static bool cfg2 = false;
      if (!cfg2)
      {
        DE_CUDA(ctx, cudaFuncSetAttribute(de::ts2_update_kernel<M, DO_GRAM>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)C2::SMEM));
        cfg2 = true;
      }
      const long long nt2 = (a.n + C2::TR - 1) / C2::TR;
      const int grid2 = (int)std::max<long long>(1, std::min<long long>(nt2, (long long)ctx->sm_count));
      a.partials = ctx->partials;
      a.done = ctx->done_ptr;
      {
        ProfScope prof(ctx, DE_PROF_UPDATE);
        DE_CUDA(ctx, launch_pdl(de::ts2_update_kernel<M, DO_GRAM>, dim3(grid2), dim3(C2::THREADS), C2::SMEM, ctx->stream, a));
      }
      DE_LAUNCH_CHECK(ctx);
      if (DO_GRAM)
        DE_TRY(reduce_partials(ctx, ctx->partials, grid2, M * M, gram_out));
      return DE_OK;
    }
    if constexpr (!DO_UPDATE && DO_GRAM && UPPER && SAME)
    {
      // G = X^T X of one block: warp-specialised tensor-core kernel (kernels_tallskinny2.cuh)
      using C3 = de::Tg2Cfg<M>;
      static bool cfg3 = false;
      if (!cfg3)
      {
        DE_CUDA(ctx, cudaFuncSetAttribute(de::ts2_gram_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C3::SMEM));
        cfg3 = true;
      }
      const long long nt3 = (a.n + C3::TR - 1) / C3::TR;
      const int grid3 = (int)std::max<long long>(1, std::min<long long>(nt3, (long long)ctx->sm_count));
      a.partials = ctx->partials;
      a.done = ctx->done_ptr;
      {
        ProfScope prof(ctx, DE_PROF_GRAM);
        DE_CUDA(ctx, launch_pdl(de::ts2_gram_kernel<M>, dim3(grid3), dim3(de::kTg2Threads), C3::SMEM, ctx->stream, a));
      }
      DE_LAUNCH_CHECK(ctx);
      return reduce_partials(ctx, ctx->partials, grid3, M * M, gram_out);
    }
    constexpr int NOPS = (DO_GRAM && !SAME) ? 2 : 1;
    using C = de::TsCfg<M, UPPER, NOPS>;
    constexpr size_t smem = de::tall_skinny_smem_bytes<M, DO_UPDATE, DO_GRAM, UPPER, SAME>();
    static bool configured = false;
    if (!configured)
    {
      DE_CUDA(ctx, cudaFuncSetAttribute(de::tall_skinny_kernel<M, DO_UPDATE, DO_GRAM, UPPER, SAME>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = true;
    }
    static int ctas_per_sm = 0; // resident CTAs per SM of this instantiation (registers / shared memory decide)
    if (ctas_per_sm == 0)
    {
      int occ = 1;
      DE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(
                       &occ, de::tall_skinny_kernel<M, DO_UPDATE, DO_GRAM, UPPER, SAME>, C::THREADS, smem));
      ctas_per_sm = std::max(1, std::min(occ, 2));
    }
    const long long ntiles = (a.n + C::TR - 1) / C::TR;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * ctas_per_sm));
    a.partials = ctx->partials;
    a.done = ctx->done_ptr;
    {
      ProfScope prof(ctx, DO_UPDATE ? DE_PROF_UPDATE : DE_PROF_GRAM);
      de::tall_skinny_kernel<M, DO_UPDATE, DO_GRAM, UPPER, SAME><<<grid, C::THREADS, smem, ctx->stream>>>(a);
    }
    DE_LAUNCH_CHECK(ctx);
    if (DO_GRAM)
      DE_TRY(reduce_partials(ctx, ctx->partials, grid, M * M, gram_out));
    return DE_OK;
  }

  template <bool DO_UPDATE, bool DO_GRAM, bool UPPER, bool SAME>
  int launch_ts(de_context *ctx, int w, const de::TsArgs &a, double *gram_out)
  {
    switch (w)
    {
    case 8:
      return launch_ts_t<8, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    case 16:
      return launch_ts_t<16, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    case 32:
      return launch_ts_t<32, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    case 64:
      return launch_ts_t<64, DO_UPDATE, DO_GRAM, UPPER, SAME>(ctx, a, gram_out);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "pipelined tall-skinny kernel: unsupported width");
  }

  // ---- Gram -------------------------------------------------------------------------------------------
  template <int M, bool UPPER, bool SAME>
  int launch_gram_t(de_context *ctx, long long n, const double *X, int ldx, const double *Y, int ldy, double *out)
  {
    using C = de::GramCfg<M, UPPER, SAME>;
    const long long ntiles = (n + C::TR - 1) / C::TR;
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, kMaxPartials));
    {
      ProfScope prof(ctx, DE_PROF_GRAM);
      de::gram_kernel<M, UPPER, SAME><<<grid, C::THREADS, 0, ctx->stream>>>(n, X, ldx, Y, ldy, ctx->partials);
    }
    DE_LAUNCH_CHECK(ctx);
    return reduce_partials(ctx, ctx->partials, grid, M * M, out);
  }

  template <bool UPPER, bool SAME>
  int launch_gram_m(de_context *ctx, int w, long long n, const double *X, int ldx, const double *Y, int ldy, double *out)
  {
    switch (w)
    {
    case 8:
      return launch_gram_t<8, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 16:
      return launch_gram_t<16, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 24:
      return launch_gram_t<24, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 32:
      return launch_gram_t<32, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 40:
      return launch_gram_t<40, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 48:
      return launch_gram_t<48, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 56:
      return launch_gram_t<56, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    case 64:
      return launch_gram_t<64, UPPER, SAME>(ctx, n, X, ldx, Y, ldy, out);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "gram: column count must be a multiple of 8 in [8,64]");
  }

  /** out (device, w*w) = X^T Y over the local rows, all-reduced over the ranks.
   *  symmetric: the result is known to be symmetric (only upper blocks are computed and mirrored). */
  int gram_device(de_context *ctx, int w, long long n, const double *X, int ldx, const double *Y, int ldy,
                  bool symmetric, double *out)
  {
    const bool same = (X == Y && ldx == ldy);
    if (ts_supported(w))
    {
      de::TsArgs a{};
      a.n = n;
      a.X = X;
      a.ldx = ldx;
      a.Y = Y;
      a.ldy = ldy;
      if (symmetric && same)
        DE_TRY((launch_ts<false, true, true, true>(ctx, w, a, out)));
      else if (symmetric)
        DE_TRY((launch_ts<false, true, true, false>(ctx, w, a, out)));
      else if (same)
        DE_TRY((launch_ts<false, true, false, true>(ctx, w, a, out)));
      else
        DE_TRY((launch_ts<false, true, false, false>(ctx, w, a, out)));
      return allreduce_sum(ctx, out, (size_t)w * w);
    }
    if (symmetric && same)
      DE_TRY((launch_gram_m<true, true>(ctx, w, n, X, ldx, Y, ldy, out)));
    else if (symmetric)
      DE_TRY((launch_gram_m<true, false>(ctx, w, n, X, ldx, Y, ldy, out)));
    else if (same)
      DE_TRY((launch_gram_m<false, true>(ctx, w, n, X, ldx, Y, ldy, out)));
    else
      DE_TRY((launch_gram_m<false, false>(ctx, w, n, X, ldx, Y, ldy, out)));
    return allreduce_sum(ctx, out, (size_t)w * w);
  }

  // ---- block update -----------------------------------------------------------------------------------
  template <int M, int MODE>
  int launch_update_t(de_context *ctx, long long n, const double *X, int ldx, const double *R, double *Y, int ldy,
                      int upper)
  {
    using C = de::UpdCfg<M>;
    static bool configured = false; // per instantiation; same attribute for every device of this process
    if (!configured)
    {
      DE_CUDA(ctx, cudaFuncSetAttribute(de::update_kernel<M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)C::SMEM_BYTES));
      configured = true;
    }
    const long long ntiles = (n + C::TR - 1) / C::TR;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)(200 * 1024) / C::SMEM_BYTES));
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    ProfScope prof(ctx, DE_PROF_UPDATE);
    de::update_kernel<M, MODE><<<grid, C::THREADS, C::SMEM_BYTES, ctx->stream>>>(n, X, ldx, R, Y, ldy, upper);
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  template <int MODE>
  int update_device(de_context *ctx, int w, long long n, const double *X, int ldx, const double *R, double *Y, int ldy,
                    int upper, const int *skip_flag = nullptr)
  {
    if (MODE == 0 && ts_supported(w))
    {
      de::TsArgs a{};
      a.n = n;
      a.X = X;
      a.ldx = ldx;
      a.R = R;
      a.Out = Y;
      a.ldo = ldy;
      a.upper = upper;
      a.skip_flag = skip_flag;
      return launch_ts<true, false, false, true>(ctx, w, a, nullptr);
    }
    switch (w)
    {
    case 8:
      return launch_update_t<8, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 16:
      return launch_update_t<16, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 24:
      return launch_update_t<24, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 32:
      return launch_update_t<32, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 40:
      return launch_update_t<40, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 48:
      return launch_update_t<48, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 56:
      return launch_update_t<56, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    case 64:
      return launch_update_t<64, MODE>(ctx, n, X, ldx, R, Y, ldy, upper);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "block update: column count must be a multiple of 8 in [8,64]");
  }

  // ---- (B-)orthonormalisation: CholQR2 ------------------------------------------------------------------
  int chol_inverse(de_context *ctx, int m, const double *G, double *Rinv, double *info, int *identity_flag = nullptr)
  {
    if (ctx->tail_did_op)
    {
      ctx->tail_did_op = false; // done by the fused tail of the reduction that produced G
      return DE_OK;
    }
    ProfScope prof(ctx, DE_PROF_SMALL);
    if (m <= 32)
      de::chol_inverse2_kernel<32><<<1, 1024, 0, ctx->stream>>>(m, G, Rinv, ctx->dstatus, info, identity_flag,
                                                                const_cast<int *>(ctx->done_ptr));
    else
      de::chol_inverse2_kernel<64><<<1, 1024, 0, ctx->stream>>>(m, G, Rinv, ctx->dstatus, info, identity_flag,
                                                                const_cast<int *>(ctx->done_ptr));
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  /** the next partial-sum reduction (of a Gram matrix into G) is followed, in the same launch, by the all-reduce and by
   *  Rinv = chol(G)^-1 -- exactly what chol_inverse(ctx, m, G, Rinv, info, identity_flag) would do afterwards */
  void arm_chol_tail(de_context *ctx, int m, double *Rinv, double *info, int *identity_flag)
  {
    ctx->tail = de::TailArgs{};
    ctx->tail.kind = de::kTailChol;
    ctx->tail.m = m;
    ctx->tail.Rinv = Rinv;
    ctx->tail.status = ctx->dstatus;
    ctx->tail.info = info;
    ctx->tail.identity_flag = identity_flag;
    ctx->tail.done = const_cast<int *>(ctx->done_ptr);
    ctx->tail_armed = true;
    ctx->tail_did_allreduce = ctx->tail_did_op = false;
  }

  /** X <- X R^-1 (thin QR with positive-diagonal triangular R; reference orthonormalize_blocked,
   *  kernels_cpp.hh:180-351). Two CholQR sweeps over the WHOLE block: G = X^T X, R = chol(G), X <- X R^-1.
   *  The triangular factor of a full-rank block is unique, so the result equals the reference's block
   *  Gram-Schmidt up to round-off; the second sweep restores orthogonality to O(eps) for cond(X) < ~1e7. */
  int orthonormalize_device(de_context *ctx, long long n, int m, double *X, const double *G_ready = nullptr)
  {
    if (ts_supported(m))
    {
      // sweep 1: G = X^T X (already known if the SpMM that produced X ran its Gram epilogue) ; R1 = chol(G) ;
      // X <- X R1^-1 fused with G2 = X^T X of the result
      if (G_ready == nullptr)
      {
        arm_chol_tail(ctx, m, ctx->dR(), nullptr, nullptr);
        DE_TRY(gram_device(ctx, m, n, X, m, X, m, true, ctx->dG()));
      }
      DE_TRY(chol_inverse(ctx, m, G_ready ? G_ready : ctx->dG(), ctx->dR(), nullptr, nullptr));
      de::TsArgs a{};
      a.n = n;
      a.X = X;
      a.ldx = m;
      a.R = ctx->dR();
      a.Out = X;
      a.ldo = m;
      a.upper = 1;
      // sweep 2: R2 = chol(G2) ; X <- X R2^-1, skipped on the device when G2 = I to working precision. The factor of
      // sweep 1 is read at the start of the update kernel and overwritten by the tail of its reduction: the factor
      // fragments are in registers long before the last CTA of the reduction runs (it is a later launch).
      arm_chol_tail(ctx, m, ctx->dR(), nullptr, ctx->dflags);
      DE_TRY((launch_ts<true, true, true, true>(ctx, m, a, ctx->dG())));
      DE_TRY(allreduce_sum(ctx, ctx->dG(), (size_t)m * m));
      DE_TRY(chol_inverse(ctx, m, ctx->dG(), ctx->dR(), nullptr, ctx->dflags));
      return update_device<0>(ctx, m, n, X, m, ctx->dR(), X, m, 1, ctx->dflags);
    }
    for (int sweep = 0; sweep < 2; ++sweep)
    {
      DE_TRY(gram_device(ctx, m, n, X, m, X, m, true, ctx->dG()));
      DE_TRY(chol_inverse(ctx, m, ctx->dG(), ctx->dR(), nullptr));
      DE_TRY(update_device<0>(ctx, m, n, X, m, ctx->dR(), X, m, 1));
    }
    return DE_OK;
  }

  /** X^T B X = I (reference B_orthonormalize_blocked, kernels_cpp.hh:356-591). BX = B X is formed once with the
   *  SpMM kernel and then carried through both sweeps with the same triangular factor (the reference keeps
   *  P = B V_k updated the same way, :527-539), so on return BX = B X for the new X. */
  int b_orthonormalize_device(de_context *ctx, const de_matrix *B, long long n, int m, double *X, double *BX,
                              bool want_info)
  {
    DE_TRY(spmm_device<false>(ctx, B, X, BX, m));
    for (int sweep = 0; sweep < 2; ++sweep)
    {
      double *info = (want_info && sweep == 0) ? ctx->dInfo() : nullptr;
      arm_chol_tail(ctx, m, ctx->dR(), info, nullptr); // reduce -> all-reduce -> Cholesky in one launch (kernels_tail.cuh)
      DE_TRY(gram_device(ctx, m, n, X, m, BX, m, true, ctx->dG()));
      DE_TRY(chol_inverse(ctx, m, ctx->dG(), ctx->dR(), info));
      DE_TRY(update_device<0>(ctx, m, n, X, m, ctx->dR(), X, m, 1));
      DE_TRY(update_device<0>(ctx, m, n, BX, m, ctx->dR(), BX, m, 1));
    }
    return DE_OK;
  }

  int reset_status(de_context *ctx)
  {
    DE_CUDA(ctx, cudaMemsetAsync(ctx->dstatus, 0, sizeof(int), ctx->stream));
    return DE_OK;
  }

  /** copy `count` doubles of device scratch and the sticky status to the host and wait for them */
  int fetch_small(de_context *ctx, const double *dsrc, double *hdst, size_t count)
  {
    if (count > 0)
      DE_CUDA(ctx, cudaMemcpyAsync(ctx->hsmall, dsrc, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->hstatus, ctx->dstatus, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->hflags[7] = 0;
    if (ctx->peer_ready)
      DE_CUDA(ctx, cudaMemcpyAsync(ctx->hflags + 7, ctx->dticket + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->hflags[7] != 0)
      return set_error(ctx, DE_ERR_NCCL, "NVLink peer window: a neighbour rank did not arrive within 30 s");
    if (count > 0 && hdst != ctx->hsmall)
      std::memcpy(hdst, ctx->hsmall, count * sizeof(double));
    if (*ctx->hstatus != 0)
      return set_error(ctx, DE_ERR_SINGULAR,
                       "orthonormalize: Gram matrix is not positive definite (pivot " + std::to_string(*ctx->hstatus - 1) +
                           "); the block is numerically rank deficient");
    return DE_OK;
  }

  // ---- triangular solves --------------------------------------------------------------------------------
  constexpr int kChainMaxRows = 32;

  /** rows sorted by level for a strictly-triangular CSR whose dependencies point to already-solved rows */
  int build_schedule(de_context *ctx, long long n, const std::vector<int> &ptr, const std::vector<int> &col,
                     const std::vector<double> &val, const std::vector<double> *invdiag, bool lower, TrsvSchedule &S)
  {
    std::vector<int> level(n, 0);
    int nlev = 0;
    if (lower)
      for (long long i = 0; i < n; ++i)
      {
        int l = 0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k)
          l = std::max(l, level[col[k]] + 1);
        level[i] = l;
        nlev = std::max(nlev, l + 1);
      }
    else
      for (long long i = n - 1; i >= 0; --i)
      {
        int l = 0;
        for (int k = ptr[i]; k < ptr[i + 1]; ++k)
          l = std::max(l, level[col[k]] + 1);
        level[i] = l;
        nlev = std::max(nlev, l + 1);
      }
    if (n == 0)
      nlev = 0;
    S.nlevels = nlev;
    S.h_level_ptr.assign(nlev + 1, 0);
    for (long long i = 0; i < n; ++i)
      S.h_level_ptr[level[i] + 1]++;
    for (int l = 0; l < nlev; ++l)
      S.h_level_ptr[l + 1] += S.h_level_ptr[l];
    std::vector<int> rows(n), fill(S.h_level_ptr.begin(), S.h_level_ptr.end() - (nlev >= 0 ? 1 : 0));
    for (long long i = 0; i < n; ++i)
      rows[fill[level[i]]++] = (int)i;
    // segments: runs of narrow levels are chained in one CTA, wide levels get their own launch
    S.segments.clear();
    for (int l = 0; l < nlev;)
    {
      const int width = S.h_level_ptr[l + 1] - S.h_level_ptr[l];
      if (width <= kChainMaxRows)
      {
        int e = l + 1;
        while (e < nlev && S.h_level_ptr[e + 1] - S.h_level_ptr[e] <= kChainMaxRows)
          ++e;
        S.segments.push_back(TrsvSegment{1, l, e});
        l = e;
      }
      else
      {
        S.segments.push_back(TrsvSegment{0, l, l + 1});
        ++l;
      }
    }
    S.nnz = (long long)col.size();
    DE_TRY(upload_converted(ctx, &S.rows, rows.data(), rows.size()));
    DE_TRY(upload_converted(ctx, &S.rowptr, ptr.data(), ptr.size()));
    DE_TRY(upload_converted(ctx, &S.col, col.data(), col.size()));
    DE_TRY(upload_converted(ctx, &S.val, val.data(), val.size()));
    DE_TRY(upload_converted(ctx, &S.level_ptr, S.h_level_ptr.data(), S.h_level_ptr.size()));
    if (invdiag)
      DE_TRY(upload_converted(ctx, &S.invdiag, invdiag->data(), invdiag->size()));
    return DE_OK;
  }

  void free_schedule(TrsvSchedule &S)
  {
    dev_free(S.rows);
    dev_free(S.rowptr);
    dev_free(S.col);
    dev_free(S.val);
    dev_free(S.level_ptr);
    dev_free(S.invdiag);
  }

  template <int LC>
  int run_schedule_t(de_context *ctx, const TrsvSchedule &S, double *W, int m)
  {
    de::TrsvArgs a{S.rows, S.rowptr, S.col, S.val, S.invdiag, W, m};
    for (const TrsvSegment &seg : S.segments)
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      if (seg.chain)
        de::trsv_chain_kernel<LC><<<1, 1024, 0, ctx->stream>>>(a, S.level_ptr, seg.a, seg.b);
      else
      {
        const int first = S.h_level_ptr[seg.a], count = S.h_level_ptr[seg.a + 1] - first;
        de::trsv_level_kernel<LC><<<(count + 7) / 8, 256, 0, ctx->stream>>>(a, first, count);
      }
      DE_LAUNCH_CHECK(ctx);
    }
    return DE_OK;
  }

  int run_schedule(de_context *ctx, const TrsvSchedule &S, double *W, int m)
  {
    const int hp = m / 2;
    if (hp <= 4)
      return run_schedule_t<4>(ctx, S, W, m);
    if (hp <= 8)
      return run_schedule_t<8>(ctx, S, W, m);
    if (hp <= 16)
      return run_schedule_t<16>(ctx, S, W, m);
    return run_schedule_t<32>(ctx, S, W, m);
  }

  int ensure_factor_work(de_context *ctx, de_factor *F, int m)
  {
    if (F->W_m >= m)
      return DE_OK;
    if (F->W)
      dev_free(F->W);
    F->W = nullptr;
    if (F->sweep_graph) // captured on the old work block
    {
      cudaGraphExecDestroy(F->sweep_graph);
      F->sweep_graph = nullptr;
      F->sweep_graph_m = 0;
    }
    DE_TRY(dev_alloc(ctx, &F->W, (size_t)F->n * m));
    F->W_m = m;
    return DE_OK;
  }

  /** Y = (factored A)^-1 X (reference matmul_inverse_tallskinny_blocked, kernels_cpp.hh:660-755) */
  int factor_apply_device(de_context *ctx, const de_factor *Fc, const double *X, double *Y, int m)
  {
    de_factor *F = const_cast<de_factor *>(Fc);
    if (ctx->nranks > 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "factored apply is single-GPU (triangular solves do not row-shard)");
    DE_TRY(ensure_factor_work(ctx, F, m));
    const long long total = F->n * (m / 2);
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->sm_count * 8));
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      de::permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(F->n, m, F->P, F->rowscale, X, F->W, 0);
    }
    DE_LAUNCH_CHECK(ctx);
    // forward sweep over the levels of L, backward sweep over the levels of U: a fixed sequence of launches on the
    // fixed block W -> a CUDA graph, replayed with one call (per-launch CPU cost and front-end latency dominate these
    // sweeps for 2D problems: hundreds of levels of a few hundred rows). With per-kernel timers on, launch one by one.
    bool replayed = false;
    if (!ctx->profiling)
    {
      if (F->sweep_graph == nullptr || F->sweep_graph_m != m)
      {
        if (F->sweep_graph)
          cudaGraphExecDestroy(F->sweep_graph);
        F->sweep_graph = nullptr;
        F->sweep_graph_m = 0;
        cudaGraph_t graph = nullptr;
        const long long before = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess)
        {
          const int s1 = run_schedule(ctx, F->L, F->W, m);
          const int s2 = (s1 == DE_OK) ? run_schedule(ctx, F->U, F->W, m) : s1;
          const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &graph);
          F->sweep_graph_nodes = ctx->launches - before;
          ctx->launches = before; // counted again at every replay
          if (s2 == DE_OK && ce == cudaSuccess && graph != nullptr &&
              cudaGraphInstantiate(&F->sweep_graph, graph, 0) == cudaSuccess)
            F->sweep_graph_m = m;
          else
            F->sweep_graph = nullptr;
          if (graph)
            cudaGraphDestroy(graph);
          cudaGetLastError();
        }
      }
      if (F->sweep_graph != nullptr && F->sweep_graph_m == m)
      {
        DE_CUDA(ctx, cudaGraphLaunch(F->sweep_graph, ctx->stream));
        ctx->launches += F->sweep_graph_nodes;
        replayed = true;
      }
    }
    if (!replayed)
    {
      DE_TRY(run_schedule(ctx, F->L, F->W, m));
      DE_TRY(run_schedule(ctx, F->U, F->W, m));
    }
    {
      ProfScope prof(ctx, DE_PROF_TRSV);
      de::permute_rows_kernel<<<grid, 256, 0, ctx->stream>>>(F->n, m, F->Q, nullptr, F->W, Y, 1);
    }
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  // ---- layout helpers -----------------------------------------------------------------------------------
  int ensure_stage(de_context *ctx, size_t bytes)
  {
    if (ctx->stage_bytes >= bytes)
      return DE_OK;
    if (ctx->stage)
      dev_free(ctx->stage);
    ctx->stage = nullptr;
    ctx->stage_bytes = 0;
    DE_TRY(dev_alloc(ctx, (char **)&ctx->stage, bytes));
    ctx->stage_bytes = bytes;
    return DE_OK;
  }

  int convert_layout(de_context *ctx, long long n, int m, const double *src, double *dst, int to_rowmajor)
  {
    const long long total = n * (m / 8);
    if (total == 0)
      return DE_OK;
    const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, ctx->sm_count * 8));
    ProfScope prof(ctx, DE_PROF_MISC);
    de::panel8_convert_kernel<<<grid, 256, 0, ctx->stream>>>(n, m, src, dst, to_rowmajor);
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  int upload_panel8_device(de_context *ctx, long long n, int m, const double *host, double *dst)
  {
    const size_t bytes = sizeof(double) * (size_t)n * m;
    if (bytes == 0)
      return DE_OK;
    DE_TRY(ensure_stage(ctx, bytes));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the staging block may still be read by earlier work
    DE_TRY(upload_parallel(ctx, reinterpret_cast<double *>(ctx->stage), host, (size_t)n * m));
    return convert_layout(ctx, n, m, ctx->stage, dst, 1);
  }

  /** eval / evec copy-out of the drivers (eigensolver.hh:105-111, :328-341) */
  int copy_out(de_context *ctx, long long n, int m, int nev, const double *Q, const std::vector<double> &s,
               double *eval, double *evec)
  {
    for (int j = 0; j < nev; ++j)
      eval[j] = s[j];
    if (n == 0 || nev == 0)
      return DE_OK;
    const size_t bytes = sizeof(double) * (size_t)n * nev;
    DE_TRY(ensure_stage(ctx, bytes));
    {
      ProfScope prof(ctx, DE_PROF_MISC);
      de::extract_columns_kernel<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(n, m, nev, Q, ctx->stage);
    }
    DE_LAUNCH_CHECK(ctx);
    return download_parallel(ctx, evec, ctx->stage, bytes);
  }

  struct ScopedBlocks
  {
    std::vector<double *> p;
    ~ScopedBlocks()
    {
      for (double *q : p)
        dev_free(q);
    }
    int alloc(de_context *ctx, double **out, size_t count)
    {
      DE_TRY(dev_alloc(ctx, out, count));
      p.push_back(*out);
      return DE_OK;
    }
  };

  inline int padded_cols(int nev) { return (nev / 8 + std::min(nev % 8, 1)) * 8; } // eigensolver.hh:43

  // ---- LOBPCG: device implementation of the Ops interface of lobpcg_core.hpp ----------------------------------
  template <int M>
  int launch_lincomb_t(de_context *ctx, long long n, int ns, const double *const *S, const double *C, double *out,
                       double *out2)
  {
    using K = de::LinCfg<M>;
    static bool configured = false; // per instantiation; same attribute for every device of this process
    if (!configured)
    {
      DE_CUDA(ctx, cudaFuncSetAttribute(de::lincomb_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)K::SMEM_BYTES));
      configured = true;
    }
    const long long ntiles = (n + K::TR - 1) / K::TR;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (size_t)(200 * 1024) / K::SMEM_BYTES));
    const int grid = (int)std::max<long long>(1, std::min<long long>(ntiles, (long long)ctx->sm_count * per_sm));
    ProfScope prof(ctx, DE_PROF_UPDATE);
    de::lincomb_kernel<M><<<grid, K::THREADS, K::SMEM_BYTES, ctx->stream>>>(n, ns, S[0], ns > 1 ? S[1] : nullptr,
                                                                           ns > 2 ? S[2] : nullptr, C, out, out2);
    DE_LAUNCH_CHECK(ctx);
    return DE_OK;
  }

  int lincomb_device(de_context *ctx, int m, long long n, int ns, const double *const *S, const double *C, double *out,
                     double *out2)
  {
    switch (m)
    {
    case 8:
      return launch_lincomb_t<8>(ctx, n, ns, S, C, out, out2);
    case 16:
      return launch_lincomb_t<16>(ctx, n, ns, S, C, out, out2);
    case 24:
      return launch_lincomb_t<24>(ctx, n, ns, S, C, out, out2);
    case 32:
      return launch_lincomb_t<32>(ctx, n, ns, S, C, out, out2);
    case 40:
      return launch_lincomb_t<40>(ctx, n, ns, S, C, out, out2);
    case 48:
      return launch_lincomb_t<48>(ctx, n, ns, S, C, out, out2);
    case 56:
      return launch_lincomb_t<56>(ctx, n, ns, S, C, out, out2);
    case 64:
      return launch_lincomb_t<64>(ctx, n, ns, S, C, out, out2);
    }
    return set_error(ctx, DE_ERR_UNSUPPORTED, "lincomb: column count must be a multiple of 8 in [8,64]");
  }

  struct LobpcgDeviceOps
  {
    using Blk = double *;
    de_context *ctx = nullptr;
    const de_matrix *A = nullptr, *B = nullptr;
    const de_factor *T = nullptr; // optional preconditioner: W <- T^-1 W (factored apply, kernels_trsv.cuh)
    long long n = 0;
    int m = 0;
    ScopedBlocks blocks;
    double *dcoef = nullptr;  // 3 m^2 coefficients | m Ritz values
    double *dgrams = nullptr; // 12 m^2

    int init()
    {
      DE_TRY(blocks.alloc(ctx, &dcoef, (size_t)3 * m * m + m));
      DE_TRY(blocks.alloc(ctx, &dgrams, (size_t)12 * m * m));
      return DE_OK;
    }
    int alloc(Blk *b) { return blocks.alloc(ctx, b, (size_t)n * m); }
    /** wait for the stream; surfaces the sticky Cholesky status and the peer-window error flag */
    int sync_check() { return fetch_small(ctx, nullptr, nullptr, 0); }
    int orthonormalize(Blk X, Blk BX)
    {
      DE_TRY(reset_status(ctx));
      if (B)
        DE_TRY(b_orthonormalize_device(ctx, B, n, m, X, BX, false));
      else
        DE_TRY(orthonormalize_device(ctx, n, m, X));
      return sync_check();
    }
    int apply_A(Blk Y, Blk X) { return spmm_device<false>(ctx, A, X, Y, m); }
    int apply_B(Blk Y, Blk X) { return spmm_device<false>(ctx, B, X, Y, m); }
    int residual(Blk W, Blk AX, Blk BX, const double *theta, double *norm2)
    {
      double *dtheta = dcoef + (size_t)3 * m * m;
      DE_CUDA(ctx, cudaMemcpyAsync(dtheta, theta, sizeof(double) * m, cudaMemcpyHostToDevice, ctx->stream));
      const long long pairs = n * m / 2;
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        de::residual_kernel<<<elementwise_grid(pairs), 256, 0, ctx->stream>>>(pairs, m, AX, BX, dtheta, W);
      }
      DE_LAUNCH_CHECK(ctx);
      DE_TRY(diag_dot_device(ctx, n, m, W, W, ctx->dDP()));
      return fetch_small(ctx, ctx->dDP(), norm2, m);
    }
    int precondition(Blk W)
    {
      if (!T)
        return DE_OK;
      // in place: the apply permutes W into the factor's own work block before anything is written back
      return factor_apply_device(ctx, T, W, W, m);
    }
    int elementwise_grid(long long pairs) const
    {
      return (int)std::max<long long>(1, std::min<long long>((pairs + 255) / 256, (long long)ctx->sm_count * 8));
    }
    /** Jacobi scale dinv = 1 / diag(A) and the Gershgorin bound of D^-1 A over all ranks' rows. The all-reduce of this
     *  library sums, so every rank deposits its local maximum in its own slot of a zeroed vector and the maximum is
     *  taken on the host. */
    double *ddinv = nullptr;
    int spectral_bound(double *b)
    {
      const int nr = std::max(1, ctx->nranks);
      double *slots = ctx->dDP(); // nr <= 64 doubles of scratch
      DE_TRY(blocks.alloc(ctx, &ddinv, (size_t)std::max<long long>(n, 1)));
      DE_CUDA(ctx, cudaMemsetAsync(slots, 0, sizeof(double) * nr, ctx->stream));
      {
        ProfScope prof(ctx, DE_PROF_MISC);
        de::gershgorin_kernel<<<elementwise_grid(A->n), 256, 0, ctx->stream>>>(
            A->n, A->rowptr, A->col, A->val, ddinv, reinterpret_cast<unsigned long long *>(slots + ctx->rank));
      }
      DE_LAUNCH_CHECK(ctx);
      DE_TRY(allreduce_sum(ctx, slots, nr));
      std::vector<double> h(nr, 0.0);
      DE_TRY(fetch_small(ctx, slots, h.data(), nr));
      *b = *std::max_element(h.begin(), h.end()); // +inf if some row has no positive diagonal entry: the caller then
      return DE_OK;                               // runs without the preconditioner (lobpcg_core.hpp)
    }
    /** (m/2, 256/(m/2)) thread blocks of the row-wise streaming kernels */
    dim3 row_block() const { return dim3((unsigned)(m / 2), (unsigned)(256 / (m / 2))); }
    int row_grid() const
    {
      const long long rpb = 256 / (m / 2);
      return (int)std::max<long long>(1, std::min<long long>((n + rpb - 1) / rpb, (long long)ctx->sm_count * 8));
    }
    int cheb_start(Blk Z, Blk Zold, Blk R, double s)
    {
      ProfScope prof(ctx, DE_PROF_MISC);
      de::cheb_start_kernel<<<row_grid(), row_block(), 0, ctx->stream>>>(n, m / 2, s, ddinv, R, Z, Zold);
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    int cheb_step(Blk Zold, Blk Z, Blk R, Blk AZ, double alpha, double beta)
    {
      ProfScope prof(ctx, DE_PROF_MISC);
      de::cheb_step_kernel<<<row_grid(), row_block(), 0, ctx->stream>>>(n, m / 2, alpha, beta, ddinv, Z, R, AZ, Zold);
      DE_LAUNCH_CHECK(ctx);
      return DE_OK;
    }
    int project(Blk W, Blk X, Blk BX)
    {
      DE_TRY(gram_device(ctx, m, n, BX, m, W, m, false, ctx->dG()));
      return update_device<1>(ctx, m, n, X, m, ctx->dG(), W, m, 0); // W -= X G
    }
    int grams(int count, const Blk *L, const Blk *R, const char *sym, double *out)
    {
      for (int g = 0; g < count; ++g)
        DE_TRY(gram_device(ctx, m, n, L[g], m, R[g], m, sym[g] != 0, dgrams + (size_t)g * m * m));
      DE_CUDA(ctx, cudaMemcpyAsync(out, dgrams, sizeof(double) * (size_t)count * m * m, cudaMemcpyDeviceToHost,
                                   ctx->stream));
      return sync_check();
    }
    int rotate(Blk X, const double *C)
    {
      DE_CUDA(ctx, cudaMemcpyAsync(dcoef, C, sizeof(double) * (size_t)m * m, cudaMemcpyHostToDevice, ctx->stream));
      const double *S[1] = {X};
      return lincomb_device(ctx, m, n, 1, S, dcoef, X, nullptr);
    }
    int lincomb(int ns, const Blk *S, const double *C, Blk out, Blk out2)
    {
      DE_CUDA(ctx, cudaMemcpyAsync(dcoef, C, sizeof(double) * (size_t)ns * m * m, cudaMemcpyHostToDevice, ctx->stream));
      const double *src[3] = {S[0], ns > 1 ? S[1] : nullptr, ns > 2 ? S[2] : nullptr};
      return lincomb_device(ctx, m, n, ns, src, dcoef, out, out2);
    }
  };

  /** LOBPCG on the device block X (n x m, start block on entry, Ritz vectors on return) */
  int lobpcg_device(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *T, bool largest,
                    int cheb_degree, double tol, int maxiter, int nev, int m, double *X, de::LobpcgResult &res,
                    int verbose)
  {
    LobpcgDeviceOps ops;
    ops.ctx = ctx;
    ops.A = A;
    ops.B = B;
    ops.T = T;
    ops.n = A->n;
    ops.m = m;
    DE_TRY(reset_status(ctx));
    DE_TRY(ops.init());
    de::LobpcgParams prm;
    prm.m = m;
    prm.nev = nev;
    prm.tol = tol;
    prm.maxiter = maxiter;
    prm.verbose = verbose;
    prm.has_B = B != nullptr;
    prm.largest = largest;
    prm.cheb_degree = (T || largest) ? 0 : std::max(0, cheb_degree); // a factored preconditioner takes precedence
    prm.name = B ? "GeneralizedLOBPCG" : "StandardLOBPCG";
    const int rc = de::lobpcg_run(ops, prm, X, res);
    if (rc == de::kLobpcgRitzFailed)
      return set_error(ctx, DE_ERR_SINGULAR,
                       "LOBPCG: the Rayleigh-Ritz problem on [X W] is numerically singular (or residuals are not finite)");
    if (rc != DE_OK)
      return rc;
    return ops.sync_check();
  }

  int lobpcg_check_args(de_context *ctx, const char *who, const de_matrix *A, const de_matrix *B, const de_factor *T,
                        int nev, int m)
  {
    if (!valid_cols(m))
      return set_error(ctx, DE_ERR_UNSUPPORTED, std::string(who) + ": nev exceeds DE_MAX_COLS (64)");
    if (nev <= 0 || nev > m)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": nev must be in [1, number of columns]");
    if ((B && B->n != A->n) || (T && T->n != A->n))
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": A, B and the preconditioner must have the same size");
    if (T && ctx->nranks > 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, std::string(who) + ": the factored preconditioner is single-GPU");
    return DE_OK;
  }

  int lobpcg_driver(de_context *ctx, const char *who, const de_matrix *A, const de_matrix *B, const de_factor *T,
                    double tol, int maxiter, int nev, const double *start_panel8, double *eval, double *evec, int verbose,
                    int *iterations)
  {
    if (!ctx || !A || !start_panel8 || !eval || !evec || nev <= 0)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": bad arguments");
    const int m = padded_cols(nev);
    DE_TRY(lobpcg_check_args(ctx, who, A, B, T, nev, m));
    DE_TRY(bind_device(ctx));
    const long long n = A->n;
    ScopedBlocks blk;
    double *X;
    DE_TRY(blk.alloc(ctx, &X, (size_t)n * m));
    DE_TRY(upload_panel8_device(ctx, n, m, start_panel8, X));
    de::LobpcgResult res;
    const auto t0 = std::chrono::steady_clock::now();
    DE_TRY(lobpcg_device(ctx, A, B, T, false, DE_LOBPCG_DEFAULT_CHEB_DEGREE, tol, maxiter, nev, m, X, res, verbose));
    if (iterations)
      *iterations = res.iterations;
    if (verbose > 0) // one summary line in the style of eigensolver.hh:345-350
    {
      double worst = 0.0;
      for (int j = 0; j < nev; ++j)
        worst = std::max(worst, res.resnorm[j] / std::max(std::abs(res.theta[j]), std::numeric_limits<double>::min()));
      std::printf("%s:  time_total=%g iterations=%d restarts=%d relres=%g\n", who,
                  std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), res.iterations,
                  res.restarts, worst);
    }
    return copy_out(ctx, n, m, nev, X, res.theta, eval, evec);
  }

} // namespace

// ====================================================================================================
// C ABI
// ====================================================================================================
extern "C"
{

  int de_version(void) { return 100; }

  const char *de_last_error_string(const de_context *ctx) { return ctx ? ctx->err.c_str() : g_thread_error.c_str(); }

  int de_context_create(int device, void *stream, de_context **out)
  {
    if (!out)
      return set_error(nullptr, DE_ERR_INVALID, "de_context_create: out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
      return set_error(nullptr, DE_ERR_CUDA,
                       std::string("de_context_create: no CUDA device available (") +
                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                           "); this library has no CPU fallback");
    if (device < 0 || device >= count)
      return set_error(nullptr, DE_ERR_INVALID, "de_context_create: device ordinal out of range");
    de_context *ctx = new (std::nothrow) de_context();
    if (!ctx)
      return set_error(nullptr, DE_ERR_ALLOC, "de_context_create: out of host memory");
    ctx->device = device;
    auto bail = [&](int code) {
      std::string msg = ctx->err;
      de_context_destroy(ctx);
      return set_error(nullptr, code, msg);
    };
    if (bind_device(ctx) != DE_OK)
      return bail(DE_ERR_CUDA);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess)
      ctx->sm_count = prop.multiProcessorCount;
    if (stream)
      ctx->stream = (cudaStream_t)stream;
    else
    {
      if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess)
      {
        ctx->err = "cudaStreamCreate failed";
        return bail(DE_ERR_CUDA);
      }
      ctx->own_stream = true;
    }
    bool ok = cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming) == cudaSuccess &&
              cudaMalloc((void **)&ctx->partials, kPartialDoubles * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dsmall, kSmall * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dstatus, sizeof(int)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dflags, 4 * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dtail_ticket, sizeof(int)) == cudaSuccess &&
              cudaMemset(ctx->dtail_ticket, 0, sizeof(int)) == cudaSuccess &&
              cudaMemset(ctx->dflags, 0, 4 * sizeof(int)) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->hsmall, kSmall * sizeof(double)) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->hstatus, sizeof(int)) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->hflags, 8 * sizeof(int)) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_poll[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_poll[1], cudaEventDisableTiming) == cudaSuccess &&
              cudaMemset(ctx->dstatus, 0, sizeof(int)) == cudaSuccess;
    if (!ok)
    {
      ctx->err = std::string("de_context_create: workspace allocation failed: ") + cudaGetErrorString(cudaGetLastError());
      return bail(DE_ERR_ALLOC);
    }
    *out = ctx;
    return DE_OK;
  }

  int de_context_destroy(de_context *ctx)
  {
    if (!ctx)
      return DE_OK;
    cudaSetDevice(ctx->device);
    if (ctx->comm && nccl_api().ok)
      nccl_api().CommDestroy(ctx->comm);
    for (const de_context::ProfRecord &r : ctx->prof_records)
    {
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    for (cudaEvent_t e : ctx->prof_pool)
      cudaEventDestroy(e);
    dev_free(ctx->partials);
    dev_free(ctx->dsmall);
    dev_free(ctx->dstatus);
    dev_free(ctx->dflags);
    dev_free(ctx->dtail_ticket);
    dev_free(ctx->stage);
    if (ctx->hsmall)
      cudaFreeHost(ctx->hsmall);
    if (ctx->hstatus)
      cudaFreeHost(ctx->hstatus);
    if (ctx->hflags)
      cudaFreeHost(ctx->hflags);
    for (cudaEvent_t e : ctx->ev_poll)
      if (e)
        cudaEventDestroy(e);
    dev_free(ctx->dconv);
    for (int q = 0; q < de::kPeerMaxRanks; ++q)
      if (ctx->peer_base[q] && q != ctx->rank)
        cudaIpcCloseMemHandle(ctx->peer_base[q]);
    if (ctx->window)
      cudaFree(ctx->window);
    if (ctx->dticket)
      cudaFree(ctx->dticket);
    xfer_destroy(ctx);
    dev_cache_trim(ctx->device, ctx->stream);
    if (ctx->ev_pack)
      cudaEventDestroy(ctx->ev_pack);
    if (ctx->ev_halo)
      cudaEventDestroy(ctx->ev_halo);
    if (ctx->comm_stream)
      cudaStreamDestroy(ctx->comm_stream);
    if (ctx->own_stream && ctx->stream)
      cudaStreamDestroy(ctx->stream);
    delete ctx;
    return DE_OK;
  }

  int de_context_synchronize(de_context *ctx)
  {
    if (!ctx)
      return set_error(nullptr, DE_ERR_INVALID, "null context");
    DE_TRY(bind_device(ctx));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_context_set_profiling(de_context *ctx, int enable)
  {
    if (!ctx)
      return set_error(nullptr, DE_ERR_INVALID, "null context");
    // enable: 0 = off, 1 = every category, otherwise a bit mask of categories shifted left by one (2 << DE_PROF_SPMM ...)
    ctx->profiling = enable != 0;
    ctx->prof_mask = (enable == 0 || enable == 1) ? ~0u : ((unsigned)enable >> 1);
    if (ctx->profiling)
    {
      // events are created up front so that no cudaEventCreate happens inside a timed region
      DE_TRY(bind_device(ctx));
      while (ctx->prof_pool.size() < 16384)
      {
        cudaEvent_t e = nullptr;
        DE_CUDA(ctx, cudaEventCreate(&e));
        ctx->prof_pool.push_back(e);
      }
    }
    return DE_OK;
  }

  int de_context_profile(de_context *ctx, int category, double *total_ms, int64_t *launches, int reset)
  {
    if (!ctx || category < 0 || category >= DE_PROF_CATEGORIES)
      return set_error(ctx, DE_ERR_INVALID, "de_context_profile: bad arguments");
    DE_TRY(bind_device(ctx));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (const de_context::ProfRecord &r : ctx->prof_records)
    {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess)
      {
        ctx->prof_ms[r.cat] += ms;
        ctx->prof_count[r.cat] += 1;
      }
      ctx->prof_pool.push_back(r.e0);
      ctx->prof_pool.push_back(r.e1);
    }
    ctx->prof_records.clear();
    if (total_ms)
      *total_ms = ctx->prof_ms[category];
    if (launches)
      *launches = ctx->prof_count[category];
    if (reset)
      for (int c = 0; c < DE_PROF_CATEGORIES; ++c)
      {
        ctx->prof_ms[c] = 0.0;
        ctx->prof_count[c] = 0;
      }
    return DE_OK;
  }

  int de_context_launch_count(const de_context *ctx, int64_t *count)
  {
    if (!ctx || !count)
      return set_error(ctx, DE_ERR_INVALID, "null argument");
    *count = ctx->launches;
    return DE_OK;
  }

  int de_comm_unique_id(void *id128)
  {
    if (!id128)
      return set_error(nullptr, DE_ERR_INVALID, "null id");
    if (!nccl_api().ok)
      return set_error(nullptr, DE_ERR_NCCL, "libnccl.so.2 could not be loaded");
    ncclUniqueId id;
    DE_NCCL(nullptr, nccl_api().GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "unexpected ncclUniqueId size");
    std::memcpy(id128, &id, 128);
    return DE_OK;
  }

  int de_context_init_comm(de_context *ctx, int rank, int nranks, const void *id128)
  {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks)
      return set_error(ctx, DE_ERR_INVALID, "de_context_init_comm: bad arguments");
    if (!nccl_api().ok)
      return set_error(ctx, DE_ERR_NCCL, "libnccl.so.2 could not be loaded");
    DE_TRY(bind_device(ctx));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    DE_NCCL(ctx, nccl_api().CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->rank = rank;
    ctx->nranks = nranks;
    return DE_OK;
  }

  int de_context_rank(const de_context *ctx, int *rank, int *nranks)
  {
    if (!ctx)
      return set_error(nullptr, DE_ERR_INVALID, "null context");
    if (rank)
      *rank = ctx->rank;
    if (nranks)
      *nranks = ctx->nranks;
    return DE_OK;
  }

  int de_context_peer_window_create(de_context *ctx, int64_t halo_bytes, void *ipc_handle64)
  {
    if (!ctx || !ipc_handle64 || halo_bytes < 0)
      return set_error(ctx, DE_ERR_INVALID, "de_context_peer_window_create: bad arguments");
    if (ctx->nranks < 2 || ctx->nranks > de::kPeerMaxRanks)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "de_context_peer_window_create: needs 2..8 ranks (call de_context_init_comm first)");
    if (ctx->window)
      return set_error(ctx, DE_ERR_INVALID, "de_context_peer_window_create: window exists");
    DE_TRY(bind_device(ctx));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
    ctx->halo_cap = ((size_t)halo_bytes + 255) & ~(size_t)255;
    ctx->window_bytes = de::kPeerHaloOff + 2 * ctx->halo_cap;
    // a dedicated cudaMalloc allocation: an IPC handle exports the whole allocation
    DE_CUDA(ctx, cudaMalloc((void **)&ctx->window, ctx->window_bytes));
    DE_CUDA(ctx, cudaMemset(ctx->window, 0, ctx->window_bytes));
    DE_CUDA(ctx, cudaMalloc((void **)&ctx->dticket, 2 * sizeof(int)));
    DE_CUDA(ctx, cudaMemset(ctx->dticket, 0, 2 * sizeof(int)));
    DE_CUDA(ctx, cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    DE_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->window));
    std::memcpy(ipc_handle64, &h, 64);
    return DE_OK;
  }

  int de_context_peer_window_open(de_context *ctx, const void *ipc_handles)
  {
    if (!ctx || !ipc_handles || !ctx->window)
      return set_error(ctx, DE_ERR_INVALID, "de_context_peer_window_open: bad arguments (create the window first)");
    DE_TRY(bind_device(ctx));
    for (int q = 0; q < ctx->nranks; ++q)
    {
      if (q == ctx->rank)
      {
        ctx->peer_base[q] = ctx->window;
        continue;
      }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, (const unsigned char *)ipc_handles + 64 * (size_t)q, 64);
      void *p = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
      {
        cudaGetLastError();
        for (int r = 0; r < q; ++r)
          if (r != ctx->rank && ctx->peer_base[r])
          {
            cudaIpcCloseMemHandle(ctx->peer_base[r]);
            ctx->peer_base[r] = nullptr;
          }
        return set_error(ctx, DE_ERR_UNSUPPORTED,
                         std::string("de_context_peer_window_open: cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
      }
      ctx->peer_base[q] = (unsigned char *)p;
    }
    ctx->ar_epoch = ctx->halo_epoch = 0;
    ctx->peer_ready = true; // the caller runs a barrier before the first collective (every window must be zeroed)
    return DE_OK;
  }

  int de_context_peer_ready(const de_context *ctx, int *ready)
  {
    if (!ctx || !ready)
      return set_error(nullptr, DE_ERR_INVALID, "de_context_peer_ready: bad arguments");
    *ready = ctx->peer_ready ? 1 : 0;
    return DE_OK;
  }

  int de_matrix_set_peer_deposit(de_matrix *A, const int64_t *deposit_rows, int64_t max_halo_rows_all_ranks)
  {
    if (!A || (A->npeers > 0 && !deposit_rows) || max_halo_rows_all_ranks < 0)
      return set_error(A ? A->ctx : nullptr, DE_ERR_INVALID, "de_matrix_set_peer_deposit: bad arguments");
    if (!A->ctx->peer_ready)
      return set_error(A->ctx, DE_ERR_UNSUPPORTED, "de_matrix_set_peer_deposit: the context has no peer window");
    if (A->npeers > de::kPeerMaxRanks)
      return set_error(A->ctx, DE_ERR_UNSUPPORTED, "de_matrix_set_peer_deposit: too many peers");
    A->deposit.assign(deposit_rows, deposit_rows + A->npeers);
    for (int p = 0; p < A->npeers; ++p)
      if (A->deposit[p] < 0)
        return set_error(A->ctx, DE_ERR_INVALID, "de_matrix_set_peer_deposit: negative offset");
    A->halo_rows_max = std::max<long long>(max_halo_rows_all_ranks, A->n_halo);
    A->peer_halo = true;
    return DE_OK;
  }

  // ---- matrices ---------------------------------------------------------------------------------------
  static int matrix_upload(de_context *ctx, long long n, long long ncols, long long nnz, const int64_t *rowptr,
                           const int64_t *col, const double *val, de_matrix *A)
  {
    if (nnz >= (1LL << 31) || n >= (1LL << 31) - 1 || ncols >= (1LL << 31) - 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "matrix too large for 32-bit indices on one GPU");
    if (rowptr[0] != 0 || rowptr[n] != nnz)
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create: rowptr does not match nnz");
    A->n = n;
    A->nnz = nnz;
    // 16 bytes of tail padding: the staged kernels copy 16-byte chunks
    DE_TRY(dev_alloc(ctx, &A->rowptr, (size_t)n + 1 + 4));
    DE_TRY(dev_alloc(ctx, &A->col, (size_t)nnz + 4));
    DE_TRY(dev_alloc(ctx, &A->val, (size_t)nnz + 2));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // recycled blocks: their previous users are done
    long long range[2];
    DE_TRY(upload_parallel(ctx, A->rowptr, rowptr, (size_t)n + 1));
    DE_TRY(upload_parallel(ctx, A->col, col, (size_t)nnz, range));
    if (nnz > 0 && (range[0] < 0 || range[1] >= ncols))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create: column index out of range");
    DE_TRY(upload_parallel(ctx, A->val, val, (size_t)nnz));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_matrix_create_csr(de_context *ctx, int64_t n, int64_t nnz, const int64_t *rowptr, const int64_t *col,
                           const double *val, de_matrix **out)
  {
    if (!ctx || !out || n < 0 || nnz < 0 || !rowptr || (nnz > 0 && (!col || !val)))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_csr: bad arguments");
    *out = nullptr;
    DE_TRY(bind_device(ctx));
    de_matrix *A = new de_matrix();
    A->ctx = ctx;
    int s = matrix_upload(ctx, n, n, nnz, rowptr, col, val, A);
    if (s == DE_OK)
      s = build_staged_all(ctx, A, rowptr);
    if (s == DE_OK)
      s = build_brb(ctx, A, n, n, rowptr, col, val);
    if (s != DE_OK)
    {
      de_matrix_destroy(A);
      return s;
    }
    *out = A;
    return DE_OK;
  }

  int de_matrix_create_distributed(de_context *ctx, int64_t n_owned, int64_t n_halo, int64_t nnz,
                                   const int64_t *rowptr, const int64_t *col_local, const double *val, int npeers,
                                   const int *peer_ranks, const int64_t *recv_counts, const int64_t *send_offsets,
                                   const int64_t *send_rows, de_matrix **out)
  {
    if (!ctx || !out || n_owned < 0 || n_halo < 0 || nnz < 0 || !rowptr || npeers < 0 ||
        (npeers > 0 && (!peer_ranks || !recv_counts || !send_offsets)))
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: bad arguments");
    *out = nullptr;
    DE_TRY(bind_device(ctx));
    de_matrix *A = new de_matrix();
    A->ctx = ctx;
    auto fail = [&](int s) {
      de_matrix_destroy(A);
      return s;
    };
    int s = matrix_upload(ctx, n_owned, n_owned + n_halo, nnz, rowptr, col_local, val, A);
    if (s != DE_OK)
      return fail(s);
    A->n_halo = n_halo;
    A->npeers = npeers;
    long long roff = 0;
    for (int p = 0; p < npeers; ++p)
    {
      if (peer_ranks[p] < 0 || peer_ranks[p] >= ctx->nranks || peer_ranks[p] == ctx->rank ||
          (p > 0 && peer_ranks[p] <= peer_ranks[p - 1]))
        return fail(set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: peers must be ascending ranks != own"));
      A->peer.push_back(peer_ranks[p]);
      A->recv_count.push_back(recv_counts[p]);
      A->recv_off.push_back(roff);
      roff += recv_counts[p];
      A->send_count.push_back(send_offsets[p + 1] - send_offsets[p]);
      A->send_off.push_back(send_offsets[p]);
    }
    if (roff != n_halo)
      return fail(set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: recv_counts do not sum to n_halo"));
    A->n_send = npeers > 0 ? send_offsets[npeers] : 0;
    for (long long k = 0; k < A->n_send; ++k)
      if (send_rows[k] < 0 || send_rows[k] >= n_owned)
        return fail(set_error(ctx, DE_ERR_INVALID, "de_matrix_create_distributed: send row out of range"));
    if ((s = upload_converted(ctx, &A->send_rows, send_rows, (size_t)A->n_send)) != DE_OK)
      return fail(s);
    // interior rows touch owned columns only and can run while the halo is in flight
    std::vector<int> in, bd;
    for (long long i = 0; i < n_owned; ++i)
    {
      bool halo = false;
      for (int64_t k = rowptr[i]; k < rowptr[i + 1] && !halo; ++k)
        halo = col_local[k] >= n_owned;
      (halo ? bd : in).push_back((int)i);
    }
    A->n_interior = (long long)in.size();
    A->n_boundary = (long long)bd.size();
    if ((s = upload_converted(ctx, &A->interior, in.data(), in.size())) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &A->boundary, bd.data(), bd.size())) != DE_OK)
      return fail(s);
    if ((s = build_brb(ctx, A, n_owned, n_owned + n_halo, rowptr, col_local, val)) != DE_OK)
      return fail(s);
    // the row-permuted CSR copies for the staged kernel are only built when the CSR family can be chosen by the AUTO
    // policy (brb_usable); a forced DE_SPMM_CSR then still works through the row-list kernel
    if (!(A->brb.valid && nnz >= 12 * n_owned))
    {
      if ((s = build_staged_subset(ctx, in, rowptr, col_local, val, A->st_interior)) != DE_OK)
        return fail(s);
      if ((s = build_staged_subset(ctx, bd, rowptr, col_local, val, A->st_boundary)) != DE_OK)
        return fail(s);
    }
    *out = A;
    return DE_OK;
  }

  int de_matrix_destroy(de_matrix *A)
  {
    if (!A)
      return DE_OK;
    cudaSetDevice(A->ctx->device);
    dev_free(A->rowptr);
    dev_free(A->col);
    dev_free(A->val);
    dev_free(A->send_rows);
    dev_free(A->interior);
    dev_free(A->boundary);
    dev_free(A->send_buf);
    dev_free(A->halo_buf);
    A->st_all.release();
    A->st_interior.release();
    A->st_boundary.release();
    A->brb.release();
    delete A;
    return DE_OK;
  }

  int de_matrix_rows(const de_matrix *A, int64_t *n_owned, int64_t *nnz)
  {
    if (!A)
      return set_error(nullptr, DE_ERR_INVALID, "null matrix");
    if (n_owned)
      *n_owned = A->n;
    if (nnz)
      *nnz = A->nnz;
    return DE_OK;
  }

  int de_brb_format_check(int64_t n, int64_t ncols, int64_t n_owned, const int64_t *rowptr, const int64_t *col,
                          const double *val, int nthreads, int64_t *info8, double *max_abs_diff)
  {
    if (n < 0 || ncols < 0 || !rowptr || !info8 || !max_abs_diff)
      return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: bad arguments");
    for (int i = 0; i < 8; ++i)
      info8[i] = 0;
    *max_abs_diff = 0.0;
    de::brb::Format F;
    if (n == 0 || !de::brb::build(n, ncols, rowptr, col, val, n_owned, F, nthreads) || !F.valid)
      return DE_OK; // info8[0] == 0: no BRB form
    info8[0] = 1;
    info8[1] = F.grid ? 1 : 0;
    info8[2] = F.ntiles;
    info8[3] = F.n_interior;
    info8[4] = F.nblocks;
    info8[5] = F.nsteps;
    info8[6] = F.max_u;
    info8[7] = (int64_t)F.tw | ((int64_t)F.th << 16) | ((int64_t)F.td << 32);
    // decode every tile exactly as the kernel does and apply it to a probe vector
    auto probe = [](int64_t c) { return 1.0 + (double)((c * 2654435761ull) % 1021) / 1021.0; };
    std::vector<double> y((size_t)n, 0.0);
    std::vector<char> seen((size_t)n, 0);
    for (const de::brb::TileDesc &d : F.tile)
    {
      const int *h = &F.blob[(size_t)d.blob16 * 4];
      const int nb = h[0], ns = h[1];
      const int *blkstep = h + 4, *blkrows = blkstep + nb + 1;
      const int o_step = (4 + (nb + 1) + 8 * nb + 4 * nb + 3) & ~3;
      const unsigned short *self = reinterpret_cast<const unsigned short *>(blkrows + 8 * nb);
      const int *st = h + o_step;
      const double *v = reinterpret_cast<const double *>(h + o_step + 4 * ns);
      if (d.nu != h[3] || (size_t)d.len16 * 4 < (size_t)o_step + 4 * (size_t)ns + 2 * (size_t)h[2])
        return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: inconsistent tile header");
      for (int b = 0; b < nb; ++b)
      {
        for (int g = 0; g < 8; ++g)
          if (blkrows[8 * b + g] >= 0)
          {
            seen[blkrows[8 * b + g]]++;
            // the recorded position of the row's own column must hold exactly that column
            const unsigned sl = self[8 * b + g];
            if (sl != 0xffffu && ((int)sl >= d.nu || F.ucol[(size_t)d.ucol0 + sl] != blkrows[8 * b + g]))
              return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: wrong self column id");
          }
        for (int q = blkstep[b]; q < blkstep[b + 1]; ++q)
        {
          const unsigned lc[4] = {(unsigned)st[4 * q] & 0xffffu, (unsigned)st[4 * q] >> 16, (unsigned)st[4 * q + 1] & 0xffffu,
                                  (unsigned)st[4 * q + 1] >> 16};
          const unsigned mask = (unsigned)st[4 * q + 2];
          int k = st[4 * q + 3];
          for (int bit = 0; bit < 32; ++bit)
            if ((mask >> bit) & 1u)
            {
              const int row = blkrows[8 * b + bit / 4];
              if (row < 0 || row >= n || (int)lc[bit % 4] >= d.nu)
                return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: entry outside its tile");
              y[row] += v[k++] * probe(F.ucol[(size_t)d.ucol0 + lc[bit % 4]]);
            }
        }
      }
    }
    double mx = 0.0;
    for (int64_t r = 0; r < n; ++r)
    {
      if (seen[r] != 1)
        return set_error(nullptr, DE_ERR_INVALID, "de_brb_format_check: a row is not covered exactly once");
      double ref = 0.0;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k)
        ref += val[k] * probe(col[k]);
      mx = std::max(mx, std::fabs(ref - y[r]));
    }
    *max_abs_diff = mx;
    return DE_OK;
  }

  int de_matrix_brb_selfcheck(const de_matrix *A, int64_t n, int64_t ncols, const int64_t *rowptr, const int64_t *col,
                              const double *val, int64_t *mismatches)
  {
    if (!A || !rowptr || !mismatches)
      return set_error(A ? A->ctx : nullptr, DE_ERR_INVALID, "de_matrix_brb_selfcheck: bad arguments");
    de_context *ctx = A->ctx;
    *mismatches = -1;
    if (!A->brb.valid)
      return DE_OK;
    if (n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "de_matrix_brb_selfcheck: matrix size does not match");
    DE_TRY(bind_device(ctx));
    de::brb::Format F;
    if (!de::brb::build(n, ncols, rowptr, col, val, n, F))
    {
      *mismatches = -2; // the host builder found no BRB form although the device did
      return DE_OK;
    }
    const BrbDevice &B = A->brb;
    int64_t bad = 0;
    bad += (B.ntiles != F.ntiles) + (B.n_interior != F.n_interior) + (B.max_len16 != F.max_len16) + (B.max_u != F.max_u) +
           (B.nsteps != F.nsteps) + (B.nvals != F.nvals) + (B.blob16 * 4 != F.blob.size()) + (B.nucol != F.ucol.size());
    if (bad == 0)
    {
      std::vector<int> blob(F.blob.size()), ucol(F.ucol.size());
      std::vector<de::brb::TileDesc> tile(F.tile.size());
      DE_CUDA(ctx, cudaMemcpy(blob.data(), B.blob, blob.size() * sizeof(int), cudaMemcpyDeviceToHost));
      DE_CUDA(ctx, cudaMemcpy(ucol.data(), B.ucol, ucol.size() * sizeof(int), cudaMemcpyDeviceToHost));
      DE_CUDA(ctx, cudaMemcpy(tile.data(), B.tile, tile.size() * sizeof(int4), cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < blob.size(); ++i)
        bad += blob[i] != F.blob[i];
      for (size_t i = 0; i < ucol.size(); ++i)
        bad += ucol[i] != F.ucol[i];
      bad += std::memcmp(tile.data(), F.tile.data(), tile.size() * sizeof(int4)) != 0;
    }
    else
      bad += 1000000;
    *mismatches = bad;
    return DE_OK;
  }

  int de_matrix_set_spmm_format(de_matrix *A, int format)
  {
    if (!A)
      return set_error(nullptr, DE_ERR_INVALID, "null matrix");
    if (format != DE_SPMM_AUTO && format != DE_SPMM_CSR && format != DE_SPMM_BRB)
      return set_error(A->ctx, DE_ERR_INVALID, "de_matrix_set_spmm_format: unknown format");
    if (format == DE_SPMM_BRB && !A->brb.valid)
      return set_error(A->ctx, DE_ERR_UNSUPPORTED, "de_matrix_set_spmm_format: this matrix has no BRB form");
    A->spmm_format = format;
    return DE_OK;
  }

  int de_matrix_spmm_info(const de_matrix *A, int *format, int64_t *tiles, int64_t *row_blocks, int64_t *steps,
                          int64_t *union_rows_max, int *tile_shape3)
  {
    if (!A)
      return set_error(nullptr, DE_ERR_INVALID, "null matrix");
    const bool brb = A->brb.valid && A->spmm_format != DE_SPMM_CSR;
    if (format)
      *format = brb ? DE_SPMM_BRB : DE_SPMM_CSR;
    if (tiles)
      *tiles = A->brb.valid ? A->brb.ntiles : 0;
    if (row_blocks)
      *row_blocks = A->brb.valid ? A->brb.nblocks : 0;
    if (steps)
      *steps = A->brb.valid ? A->brb.nsteps : 0;
    if (union_rows_max)
      *union_rows_max = A->brb.valid ? A->brb.max_u : 0;
    if (tile_shape3)
    {
      tile_shape3[0] = A->brb.grid ? A->brb.tw : 0;
      tile_shape3[1] = A->brb.grid ? A->brb.th : 0;
      tile_shape3[2] = A->brb.grid ? A->brb.td : 0;
    }
    return DE_OK;
  }

  int de_halo_plan_local(int64_t n_owned, const int64_t *rowptr, const int64_t *col_global, int nranks, int rank,
                         const int64_t *part, int64_t *col_local, int64_t *halo_global, int64_t *n_halo,
                         int64_t *recv_counts)
  {
    if (n_owned < 0 || !rowptr || !part || nranks < 1 || rank < 0 || rank >= nranks || !n_halo || !recv_counts)
      return set_error(nullptr, DE_ERR_INVALID, "de_halo_plan_local: bad arguments");
    if (part[rank + 1] - part[rank] != n_owned)
      return set_error(nullptr, DE_ERR_INVALID, "de_halo_plan_local: partition does not match n_owned");
    const int64_t lo = part[rank], hi = part[rank + 1], nglob = part[nranks];
    const int64_t nnz = rowptr[n_owned];
    std::vector<int64_t> ext;
    for (int64_t k = 0; k < nnz; ++k)
    {
      const int64_t g = col_global[k];
      if (g < 0 || g >= nglob)
        return set_error(nullptr, DE_ERR_INVALID, "de_halo_plan_local: column index out of range");
      if (g < lo || g >= hi)
        ext.push_back(g);
    }
    std::sort(ext.begin(), ext.end());
    ext.erase(std::unique(ext.begin(), ext.end()), ext.end());
    for (int p = 0; p < nranks; ++p)
      recv_counts[p] = 0;
    {
      int p = 0;
      for (int64_t g : ext)
      {
        while (g >= part[p + 1])
          ++p;
        recv_counts[p]++;
      }
    }
    *n_halo = (int64_t)ext.size();
    for (size_t h = 0; h < ext.size(); ++h)
      halo_global[h] = ext[h];
    {
      const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(8, nnz >> 20));
      auto work = [&](int t)
      {
        const int64_t k0 = nnz * t / nth, k1 = nnz * (t + 1) / nth;
        for (int64_t k = k0; k < k1; ++k)
        {
          const int64_t g = col_global[k];
          if (g >= lo && g < hi)
            col_local[k] = g - lo;
          else
            col_local[k] = n_owned + (std::lower_bound(ext.begin(), ext.end(), g) - ext.begin());
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nth; ++t)
        th.emplace_back(work, t);
      work(0);
      for (auto &x : th)
        x.join();
    }
    return DE_OK;
  }

  // ---- multivectors -------------------------------------------------------------------------------------
  int de_mv_create(de_context *ctx, int64_t n, int m, de_mv **out)
  {
    if (!ctx || !out || n < 0)
      return set_error(ctx, DE_ERR_INVALID, "de_mv_create: bad arguments");
    *out = nullptr;
    if (m <= 0 || m % 8 != 0)
      return set_error(ctx, DE_ERR_INVALID, "number of cols must be a multiple of block size"); // multivector.hh:49
    if (m > DE_MAX_COLS)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "de_mv_create: more than DE_MAX_COLS (64) columns");
    DE_TRY(bind_device(ctx));
    de_mv *X = new de_mv{ctx, n, m, nullptr};
    int s = dev_alloc(ctx, &X->d, (size_t)n * m);
    if (s != DE_OK)
    {
      delete X;
      return s;
    }
    cudaError_t e = cudaMemsetAsync(X->d, 0, sizeof(double) * (size_t)n * m, ctx->stream);
    if (e != cudaSuccess)
    {
      dev_free(X->d);
      delete X;
      return set_error(ctx, DE_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = X;
    return DE_OK;
  }

  int de_mv_destroy(de_mv *X)
  {
    if (!X)
      return DE_OK;
    cudaSetDevice(X->ctx->device);
    dev_free(X->d);
    delete X;
    return DE_OK;
  }

  int de_mv_shape(const de_mv *X, int64_t *n, int *m)
  {
    if (!X)
      return set_error(nullptr, DE_ERR_INVALID, "null multivector");
    if (n)
      *n = X->n;
    if (m)
      *m = X->m;
    return DE_OK;
  }

  int de_mv_upload_panel8(de_mv *X, const double *host)
  {
    if (!X || !host)
      return set_error(X ? X->ctx : nullptr, DE_ERR_INVALID, "de_mv_upload_panel8: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    DE_TRY(upload_panel8_device(ctx, X->n, X->m, host, X->d));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_mv_download_panel8(const de_mv *X, double *host)
  {
    if (!X || !host)
      return set_error(X ? X->ctx : nullptr, DE_ERR_INVALID, "de_mv_download_panel8: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    const size_t bytes = sizeof(double) * (size_t)X->n * X->m;
    if (bytes == 0)
      return DE_OK;
    DE_TRY(ensure_stage(ctx, bytes));
    DE_TRY(convert_layout(ctx, X->n, X->m, X->d, ctx->stage, 0));
    DE_CUDA(ctx, cudaMemcpyAsync(host, ctx->stage, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_mv_upload_rowmajor(de_mv *X, const double *host)
  {
    if (!X || !host)
      return set_error(X ? X->ctx : nullptr, DE_ERR_INVALID, "de_mv_upload_rowmajor: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    DE_CUDA(ctx, cudaMemcpyAsync(X->d, host, sizeof(double) * (size_t)X->n * X->m, cudaMemcpyHostToDevice, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_mv_download_rowmajor(const de_mv *X, double *host)
  {
    if (!X || !host)
      return set_error(X ? X->ctx : nullptr, DE_ERR_INVALID, "de_mv_download_rowmajor: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    DE_CUDA(ctx, cudaMemcpyAsync(host, X->d, sizeof(double) * (size_t)X->n * X->m, cudaMemcpyDeviceToHost, ctx->stream));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_mv_copy(de_mv *dst, const de_mv *src)
  {
    if (!dst || !src)
      return set_error(nullptr, DE_ERR_INVALID, "de_mv_copy: null argument");
    de_context *ctx = dst->ctx;
    if (dst->n != src->n || dst->m != src->m)
      return set_error(ctx, DE_ERR_INVALID, "de_mv_copy: shape mismatch");
    DE_TRY(bind_device(ctx));
    DE_CUDA(ctx, cudaMemcpyAsync(dst->d, src->d, sizeof(double) * (size_t)src->n * src->m, cudaMemcpyDeviceToDevice,
                                 ctx->stream));
    return DE_OK;
  }

  int de_mv_device_ptr(de_mv *X, void **dptr)
  {
    if (!X || !dptr)
      return set_error(nullptr, DE_ERR_INVALID, "null argument");
    *dptr = X->d;
    return DE_OK;
  }

  // ---- kernels --------------------------------------------------------------------------------------------
  static int check_spmm_shapes(de_context *ctx, const char *who, const de_mv *Y, const de_matrix *A, const de_mv *X)
  {
    if (Y->n != A->n || X->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": number of rows does not match the matrix");
    if (Y->m != X->m)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": number of columns does not match");
    if (Y->d == X->d)
      return set_error(ctx, DE_ERR_INVALID, std::string(who) + ": output must not alias input");
    return DE_OK;
  }

  int de_spmm(de_mv *Y, const de_matrix *A, const de_mv *X)
  {
    if (!Y || !A || !X)
      return set_error(nullptr, DE_ERR_INVALID, "de_spmm: null argument");
    de_context *ctx = A->ctx;
    DE_TRY(check_spmm_shapes(ctx, "matmul_sparse_tallskinny", Y, A, X));
    DE_TRY(bind_device(ctx));
    return spmm_device<false>(ctx, A, X->d, Y->d, X->m);
  }

  int de_spmm_diag_dot(de_mv *Y, const de_matrix *A, const de_mv *X, double *dp_host)
  {
    if (!Y || !A || !X || !dp_host)
      return set_error(nullptr, DE_ERR_INVALID, "de_spmm_diag_dot: null argument");
    de_context *ctx = A->ctx;
    DE_TRY(check_spmm_shapes(ctx, "matmul_sparse_tallskinny", Y, A, X));
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    DE_TRY(spmm_device<true>(ctx, A, X->d, Y->d, X->m));
    return fetch_small(ctx, ctx->dDP(), dp_host, X->m);
  }

  int de_spmm_gram(de_mv *Y, const de_matrix *A, const de_mv *X, double *dp_host, double *G_host)
  {
    DE_TRY(check_spmm_shapes(Y ? Y->ctx : nullptr, "de_spmm_gram", Y, A, X));
    if (!dp_host || !G_host)
      return set_error(Y->ctx, DE_ERR_INVALID, "de_spmm_gram: null output");
    de_context *ctx = Y->ctx;
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    const int m = X->m;
    bool fused = false;
    DE_TRY(spmm_device<true>(ctx, A, X->d, Y->d, m, &fused));
    const double *G = ctx->dDG() + m;
    if (!fused)
    {
      // no Gram epilogue for this matrix / width: a separate pass over Y
      DE_TRY(gram_device(ctx, m, Y->n, Y->d, m, Y->d, m, true, ctx->dG()));
      G = ctx->dG();
    }
    DE_TRY(fetch_small(ctx, ctx->dDP(), dp_host, (size_t)m));
    DE_CUDA(ctx, cudaMemcpy(G_host, G, sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    return DE_OK;
  }

  int de_diag_dot(double *dp_host, const de_mv *X, const de_mv *Y)
  {
    if (!dp_host || !X || !Y)
      return set_error(nullptr, DE_ERR_INVALID, "de_diag_dot: null argument");
    de_context *ctx = X->ctx;
    if (X->n != Y->n)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of rows does not match"); // kernels_cpp.hh:30
    if (X->m != Y->m)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of columns does not match"); // :32
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    DE_TRY(diag_dot_device(ctx, X->n, X->m, X->d, Y->d, ctx->dDP()));
    return fetch_small(ctx, ctx->dDP(), dp_host, X->m);
  }

  int de_gram(double *G_host, const de_mv *X, const de_mv *Y)
  {
    if (!G_host || !X || !Y)
      return set_error(nullptr, DE_ERR_INVALID, "de_gram: null argument");
    de_context *ctx = X->ctx;
    if (X->n != Y->n)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of rows does not match"); // kernels_cpp.hh:62
    if (X->m != Y->m)
      return set_error(ctx, DE_ERR_INVALID, "dot_products_blocked: number of columns does not match"); // :64
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    // X^T X is symmetric: only the upper block triangle is computed (and mirrored) when both operands are the same block
    DE_TRY(gram_device(ctx, X->m, X->n, X->d, X->m, Y->d, Y->m, X->d == Y->d, ctx->dG()));
    return fetch_small(ctx, ctx->dG(), G_host, (size_t)X->m * X->m);
  }

  int de_block_update(de_mv *X, const double *Q_host)
  {
    if (!X || !Q_host)
      return set_error(nullptr, DE_ERR_INVALID, "de_block_update: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    const size_t cnt = (size_t)X->m * X->m;
    std::memcpy(ctx->hsmall, Q_host, cnt * sizeof(double));
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->dR(), ctx->hsmall, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DE_TRY(update_device<0>(ctx, X->m, X->n, X->d, X->m, ctx->dR(), X->d, X->m, 0));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // hsmall may be reused by the next call
    return DE_OK;
  }

  int de_block_project(de_mv *X, int j0, int k0, int w, const double *S_host)
  {
    if (!X || !S_host)
      return set_error(nullptr, DE_ERR_INVALID, "de_block_project: null argument");
    de_context *ctx = X->ctx;
    if (w <= 0 || w % 8 != 0 || j0 % 8 != 0 || k0 % 8 != 0 || j0 < 0 || k0 < 0 || j0 + w > X->m || k0 + w > X->m ||
        (j0 < k0 + w && k0 < j0 + w))
      return set_error(ctx, DE_ERR_INVALID, "de_block_project: panels must be disjoint, 8-aligned and inside the block");
    DE_TRY(bind_device(ctx));
    const size_t cnt = (size_t)w * w;
    std::memcpy(ctx->hsmall, S_host, cnt * sizeof(double));
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->dR(), ctx->hsmall, cnt * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    DE_TRY(update_device<1>(ctx, w, X->n, X->d + k0, X->m, ctx->dR(), X->d + j0, X->m, 0));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return DE_OK;
  }

  int de_orthonormalize(de_mv *X)
  {
    if (!X)
      return set_error(nullptr, DE_ERR_INVALID, "de_orthonormalize: null argument");
    de_context *ctx = X->ctx;
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    DE_TRY(orthonormalize_device(ctx, X->n, X->m, X->d));
    return fetch_small(ctx, nullptr, nullptr, 0);
  }

  int de_b_orthonormalize(const de_matrix *B, de_mv *X, de_mv *BX, double *norm)
  {
    if (!B || !X)
      return set_error(nullptr, DE_ERR_INVALID, "de_b_orthonormalize: null argument");
    de_context *ctx = X->ctx;
    if (B->n != X->n || (BX && (BX->n != X->n || BX->m != X->m)))
      return set_error(ctx, DE_ERR_INVALID, "B_orthonormalize: shape mismatch");
    DE_TRY(bind_device(ctx));
    DE_TRY(reset_status(ctx));
    ScopedBlocks tmp;
    double *bx = BX ? BX->d : nullptr;
    if (!bx)
      DE_TRY(tmp.alloc(ctx, &bx, (size_t)X->n * X->m));
    DE_TRY(b_orthonormalize_device(ctx, B, X->n, X->m, X->d, bx, norm != nullptr));
    double info = 0.0;
    DE_TRY(fetch_small(ctx, ctx->dInfo(), &info, norm ? 1 : 0));
    if (norm)
      *norm = info;
    return DE_OK;
  }

  // ---- factored apply -----------------------------------------------------------------------------------------
  int de_factor_upload(de_context *ctx, int64_t n, const long *Lp, const long *Lj, const double *Lx, const long *Up,
                       const long *Ui, const double *Ux, const long *P, const long *Q, const double *Rs, long do_recip,
                       de_factor **out)
  {
    if (!ctx || !out || n < 0 || !Lp || !Up || !P || !Q || !Rs)
      return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: bad arguments");
    *out = nullptr;
    DE_TRY(bind_device(ctx));
    const long lnz = Lp[n], unz = Up[n];
    if (lnz >= (1L << 31) || unz >= (1L << 31) || n >= (1L << 31) - 1)
      return set_error(ctx, DE_ERR_UNSUPPORTED, "factor too large for 32-bit indices");
    // L: CSR with the (unit) diagonal stored last in each row -> strip it (kernels_cpp.hh:717 skips it the same way)
    std::vector<int> lptr(n + 1, 0), lcol;
    std::vector<double> lval;
    lcol.reserve(lnz > n ? lnz - n : 0);
    lval.reserve(lnz > n ? lnz - n : 0);
    for (int64_t i = 0; i < n; ++i)
    {
      if (Lp[i + 1] - Lp[i] < 1 || Lj[Lp[i + 1] - 1] != i)
        return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: L rows must end with their diagonal entry");
      for (long k = Lp[i]; k < Lp[i + 1] - 1; ++k)
      {
        if (Lj[k] < 0 || Lj[k] >= i)
          return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: L is not strictly lower triangular");
        lcol.push_back((int)Lj[k]);
        lval.push_back(Lx[k]);
      }
      lptr[i + 1] = (int)lcol.size();
    }
    // U: CSC with the diagonal last in each column -> CSR of the strictly upper part + inverse diagonal
    std::vector<int> uptr(n + 1, 0);
    std::vector<double> invd(n, 0.0);
    for (int64_t j = 0; j < n; ++j)
    {
      if (Up[j + 1] - Up[j] < 1 || Ui[Up[j + 1] - 1] != j)
        return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: U columns must end with their diagonal entry");
      const double d = Ux[Up[j + 1] - 1];
      if (d == 0.0 || !std::isfinite(d))
        return set_error(ctx, DE_ERR_SINGULAR, "UMFPackFactorizedMatrix: input matrix is singular"); // umfpacktools.hh:163
      invd[j] = 1.0 / d;
      for (long k = Up[j]; k < Up[j + 1] - 1; ++k)
      {
        if (Ui[k] < 0 || Ui[k] >= j)
          return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: U is not strictly upper triangular");
        uptr[Ui[k] + 1]++;
      }
    }
    for (int64_t i = 0; i < n; ++i)
      uptr[i + 1] += uptr[i];
    std::vector<int> ucol(uptr[n]);
    std::vector<double> uval(uptr[n]);
    {
      std::vector<int> w(uptr.begin(), uptr.end() - 1);
      for (int64_t j = 0; j < n; ++j) // ascending j => ascending columns inside every row
        for (long k = Up[j]; k < Up[j + 1] - 1; ++k)
        {
          const int dst = w[Ui[k]]++;
          ucol[dst] = (int)j;
          uval[dst] = Ux[k];
        }
    }
    std::vector<double> rowscale(n);
    std::vector<char> seenp(n, 0), seenq(n, 0);
    for (int64_t k = 0; k < n; ++k)
    {
      if (P[k] < 0 || P[k] >= n || Q[k] < 0 || Q[k] >= n || seenp[P[k]] || seenq[Q[k]])
        return set_error(ctx, DE_ERR_INVALID, "de_factor_upload: P / Q are not permutations");
      seenp[P[k]] = seenq[Q[k]] = 1;
      rowscale[k] = do_recip ? Rs[P[k]] : 1.0 / Rs[P[k]]; // kernels_cpp.hh:687, :699
    }
    de_factor *F = new de_factor();
    F->ctx = ctx;
    F->n = n;
    F->lnz = lnz;
    F->unz = unz;
    auto fail = [&](int s) {
      de_factor_destroy(F);
      return s;
    };
    int s;
    if ((s = build_schedule(ctx, n, lptr, lcol, lval, nullptr, true, F->L)) != DE_OK)
      return fail(s);
    if ((s = build_schedule(ctx, n, uptr, ucol, uval, &invd, false, F->U)) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &F->P, P, (size_t)n)) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &F->Q, Q, (size_t)n)) != DE_OK)
      return fail(s);
    if ((s = upload_converted(ctx, &F->rowscale, rowscale.data(), (size_t)n)) != DE_OK)
      return fail(s);
    *out = F;
    return DE_OK;
  }

  int de_factor_destroy(de_factor *F)
  {
    if (!F)
      return DE_OK;
    cudaSetDevice(F->ctx->device);
    free_schedule(F->L);
    free_schedule(F->U);
    if (F->sweep_graph)
      cudaGraphExecDestroy(F->sweep_graph);
    dev_free(F->P);
    dev_free(F->Q);
    dev_free(F->rowscale);
    dev_free(F->W);
    delete F;
    return DE_OK;
  }

  int de_factor_apply(de_mv *Y, const de_factor *F, de_mv *X)
  {
    if (!Y || !F || !X)
      return set_error(nullptr, DE_ERR_INVALID, "de_factor_apply: null argument");
    de_context *ctx = F->ctx;
    if (Y->n != X->n || Y->m != X->m)
      return set_error(ctx, DE_ERR_INVALID, "matmul_inverse_tallskinny_blocked: Qout/Qin size mismatch"); // kernels_cpp.hh:665
    if (F->n != X->n)
      return set_error(ctx, DE_ERR_INVALID,
                       "matmul_inverse_tallskinny_blocked: Factorization does not match size of Qout/Qin"); // :667
    DE_TRY(bind_device(ctx));
    return factor_apply_device(ctx, F, X->d, Y->d, X->m);
  }

  int de_factor_info(const de_factor *F, int64_t *n, int64_t *lnz, int64_t *unz, int *levels_L, int *levels_U)
  {
    if (!F)
      return set_error(nullptr, DE_ERR_INVALID, "null factor");
    if (n)
      *n = F->n;
    if (lnz)
      *lnz = F->lnz;
    if (unz)
      *unz = F->unz;
    if (levels_L)
      *levels_L = F->L.nlevels;
    if (levels_U)
      *levels_U = F->U.nlevels;
    return DE_OK;
  }

  // ---- drivers --------------------------------------------------------------------------------------------------
  /** shared skeleton of StandardLargest (eigensolver.hh:28-112) and StandardInverse (:116-198) on device blocks.
   *  On entry Qa holds the start block; on exit Qa holds the orthonormal iterate (reference Q1 after the swap)
   *  and Qb the block it was mapped to (reference Q2). For the largest-eigenvalue variant the product A*Qa that
   *  the reference recomputes at the top of the loop (:78) is the one it already formed for the Rayleigh
   *  quotients (:84) -- it is reused, bit-identically. */
  /** Asynchronous form of the StandardLargest loop (no factor): the convergence test runs on the device
   *  (convergence_kernel) and raises a flag that turns every kernel of the iterations already enqueued into a no-op,
   *  so the host never waits for the GPU inside the loop -- it enqueues kPollEvery iterations, requests a copy of the
   *  flags, and only looks at the copy requested one batch earlier. The blocks are frozen in the state of the
   *  converged iteration; which buffer is which follows from the parity of the iteration count. */
  static int standard_core_async(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, int m,
                                 double *&Qa, double *&Qb, std::vector<double> &s2, int verbose, int *k_exit_out)
  {
    constexpr int kPollEvery = 4;
    const long long n = A->n;
    struct Guard
    {
      de_context *c;
      ~Guard() { c->done_ptr = nullptr; }
    } guard{ctx};
    const size_t need = 64 + (size_t)std::max(maxiter, 2) + 1;
    if (ctx->dconv_cap < need)
    {
      dev_free(ctx->dconv);
      ctx->dconv = nullptr;
      ctx->dconv_cap = 0;
      DE_TRY(dev_alloc(ctx, &ctx->dconv, need));
      ctx->dconv_cap = need;
    }
    double *s_prev = ctx->dconv, *hist = ctx->dconv + 64;
    DE_TRY(reset_status(ctx));
    DE_CUDA(ctx, cudaMemsetAsync(ctx->dflags, 0, 4 * sizeof(int), ctx->stream));
    DE_CUDA(ctx, cudaMemsetAsync(ctx->dconv, 0, need * sizeof(double), ctx->stream));
    ctx->done_ptr = ctx->dflags + 1;
    DE_TRY(orthonormalize_device(ctx, n, m, Qa)); // (:69)
    s2.assign(m, 0.0);
    int enqueued = 0;
    bool have_product = false, finished = false;
    bool have_gram = false; // dDG() + m holds Qb^T Qb of the block the next orthonormalisation works on
    int pending[2] = {0, 0}; // poll slot in flight?
    const auto t_enq0 = std::chrono::steady_clock::now();
    double t_wait = 0.0;
    int slot = 0;
    for (int k = 1; k < maxiter && !finished; ++k)
    {
      if (!have_product)
        DE_TRY(spmm_device<false>(ctx, A, Qa, Qb, m)); // Qb = A Qa (:78)
      DE_TRY(orthonormalize_device(ctx, n, m, Qb, have_gram ? ctx->dDG() + m : nullptr)); // (:81)
      // Qa = A Qb, dp = diag(Qb^T Qa) (:84-85) and, in the same pass, G = Qa^T Qa for the next orthonormalisation
      // ... and the convergence test as the tail of the reduction of the dot-product partials
      ctx->tail = de::TailArgs{};
      ctx->tail.kind = de::kTailConv;
      ctx->tail.m = m;
      ctx->tail.k = k;
      ctx->tail.shift = shift;
      ctx->tail.tol = tol;
      ctx->tail.s_prev = s_prev;
      ctx->tail.hist = hist;
      ctx->tail.flags = ctx->dflags;
      ctx->tail_armed = true;
      ctx->tail_did_allreduce = ctx->tail_did_op = false;
      DE_TRY(spmm_device<true>(ctx, A, Qb, Qa, m, kFuseGramIntoSpmm ? &have_gram : nullptr));
      if (ctx->tail_did_op)
        ctx->tail_did_op = false;
      else
      {
        de::convergence_kernel<<<1, 64, 0, ctx->stream>>>(k, m, shift, tol, ctx->dDP(), s_prev, hist, ctx->dflags);
        DE_LAUNCH_CHECK(ctx);
      }
      ctx->tail_armed = false;
      std::swap(Qa, Qb); // now Qa orthonormal, Qb = A*Qa
      have_product = true;
      ++enqueued;
      if (enqueued % kPollEvery == 0)
      {
        // look at the copy requested one batch ago (it has almost always landed), then request a new one
        const int prev = slot ^ 1;
        if (pending[prev])
        {
          const auto tw0 = std::chrono::steady_clock::now();
          DE_CUDA(ctx, cudaEventSynchronize(ctx->ev_poll[prev]));
          t_wait += std::chrono::duration<double>(std::chrono::steady_clock::now() - tw0).count();
          pending[prev] = 0;
          if (ctx->hflags[4 * prev + 1] != 0)
            finished = true;
        }
        DE_CUDA(ctx, cudaMemcpyAsync(ctx->hflags + 4 * slot, ctx->dflags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        DE_CUDA(ctx, cudaEventRecord(ctx->ev_poll[slot], ctx->stream));
        pending[slot] = 1;
        slot ^= 1;
      }
    }
    ctx->done_ptr = nullptr;
    const double t_enq = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_enq0).count();
    // final state
    DE_CUDA(ctx, cudaMemcpyAsync(ctx->hflags, ctx->dflags, 4 * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    DE_TRY(fetch_small(ctx, s_prev, s2.data(), (size_t)m)); // synchronises; reports a failed Cholesky
    const int done = ctx->hflags[1];
    int k_exit = ctx->hflags[2];
    if (!done)
      k_exit = enqueued; // ran to maxiter - 1
    if (maxiter <= 1)
      k_exit = std::min(1, maxiter - 1);
    if ((enqueued - k_exit) % 2 != 0)
      std::swap(Qa, Qb); // iterations enqueued after convergence did nothing: undo their pointer swaps
    if (verbose > 1)
      std::printf("async loop: %d iterations enqueued in %.3f ms of host time (%.3f ms of it waiting for poll copies), "
                  "drained %.3f ms later\n",
                  enqueued, t_enq * 1e3, t_wait * 1e3,
                  std::chrono::duration<double>(std::chrono::steady_clock::now() - t_enq0).count() * 1e3 - t_enq * 1e3);
    if (verbose > 0 && k_exit > 1)
    {
      std::vector<double> h((size_t)k_exit + 1);
      DE_CUDA(ctx, cudaMemcpy(h.data(), hist, h.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (int k = 2; k <= k_exit; ++k)
        std::printf("Iter=%d %g\n", k, h[k]);
    }
    if (k_exit_out)
      *k_exit_out = k_exit;
    return DE_OK;
  }

  static int standard_core(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                           int maxiter, int m, double *&Qa, double *&Qb, std::vector<double> &s2, int verbose,
                           int *k_exit_out)
  {
    const long long n = A->n;
    if (!F && ts_supported(m) && maxiter > 1)
      return standard_core_async(ctx, A, shift, tol, maxiter, m, Qa, Qb, s2, verbose, k_exit_out);
    DE_TRY(reset_status(ctx));
    DE_TRY(orthonormalize_device(ctx, n, m, Qa)); // (:69, :159)
    std::vector<double> s1(m, 0.0);
    s2.assign(m, 0.0);
    int k_exit = std::min(1, maxiter - 1);
    bool have_product = false; // Qb == A*Qa already?
    for (int k = 1; k < maxiter; ++k)
    {
      k_exit = k;
      if (F)
        DE_TRY(factor_apply_device(ctx, F, Qa, Qb, m)); // Qb = A^-1 Qa (:168)
      else if (!have_product)
        DE_TRY(spmm_device<false>(ctx, A, Qa, Qb, m)); // Qb = A Qa (:78)
      DE_TRY(orthonormalize_device(ctx, n, m, Qb));    // (:81, :171)
      DE_TRY(spmm_device<true>(ctx, A, Qb, Qa, m));    // Qa = A Qb and s1 = diag(Qb^T Qa) (:84-85, :174-175)
      DE_TRY(fetch_small(ctx, ctx->dDP(), s1.data(), m));
      double distance = 0.0;
      for (int i = 0; i < m; ++i)
      {
        s1[i] -= shift;
        distance = std::max(distance, std::abs(s1[i] - s2[i]));
      }
      if (verbose > 0 && k > 1)
        std::printf("%s=%d %g\n", F ? "iter" : "Iter", k, distance);
      std::swap(s1, s2);
      std::swap(Qa, Qb); // now Qa orthonormal, Qb = A*Qa
      have_product = true;
      if (k > 1 && distance < tol) // absolute change of the Rayleigh quotients (:101-102, :188-189)
        break;
    }
    if (k_exit_out)
      *k_exit_out = k_exit;
    return DE_OK;
  }

  static int standard_driver(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                             int maxiter, int nev, const double *start_panel8, double *eval, double *evec, int verbose,
                             int *iterations)
  {
    if (!ctx || !A || !start_panel8 || !eval || !evec || nev <= 0)
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: bad arguments");
    const int m = padded_cols(nev);
    if (!valid_cols(m))
      return set_error(ctx, DE_ERR_UNSUPPORTED, "standard eigensolver driver: nev exceeds DE_MAX_COLS (64)");
    if (F && F->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: factorisation does not match the matrix");
    DE_TRY(bind_device(ctx));
    const long long n = A->n;
    ScopedBlocks blk;
    double *Qa, *Qb;
    DE_TRY(blk.alloc(ctx, &Qa, (size_t)n * m));
    DE_TRY(blk.alloc(ctx, &Qb, (size_t)n * m));
    DE_TRY(upload_panel8_device(ctx, n, m, start_panel8, Qa));
    std::vector<double> s2;
    DE_TRY(standard_core(ctx, A, F, shift, tol, maxiter, m, Qa, Qb, s2, verbose, iterations));
    return copy_out(ctx, n, m, nev, Qa, s2, eval, evec);
  }

  /** device-resident variant: Q holds the start block on entry and the eigenvector block on return */
  static int standard_driver_mv(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                                int maxiter, de_mv *Q, double *eval_m, int verbose, int *iterations)
  {
    if (!ctx || !A || !Q || !eval_m)
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: bad arguments");
    if (Q->n != A->n || (F && F->n != A->n))
      return set_error(ctx, DE_ERR_INVALID, "standard eigensolver driver: block / factorisation do not match the matrix");
    DE_TRY(bind_device(ctx));
    double *Qa = Q->d, *Qb = nullptr;
    DE_TRY(dev_alloc(ctx, &Qb, (size_t)Q->n * Q->m));
    std::vector<double> s2;
    int s = standard_core(ctx, A, F, shift, tol, maxiter, Q->m, Qa, Qb, s2, verbose, iterations);
    Q->d = Qa; // the buffers may have swapped roles; Q keeps the one with the result
    dev_free(Qb);
    if (s != DE_OK)
      return s;
    for (int j = 0; j < Q->m; ++j)
      eval_m[j] = s2[j];
    return DE_OK;
  }

  int de_standard_largest_mv(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, de_mv *Q,
                             double *eval_m, int verbose, int *iterations)
  {
    return standard_driver_mv(ctx, A, nullptr, shift, tol, maxiter, Q, eval_m, verbose, iterations);
  }

  int de_standard_inverse_mv(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                             int maxiter, de_mv *Q, double *eval_m, int verbose, int *iterations)
  {
    if (!F)
      return set_error(ctx, DE_ERR_INVALID, "de_standard_inverse_mv: factorisation is null");
    return standard_driver_mv(ctx, A, F, shift, tol, maxiter, Q, eval_m, verbose, iterations);
  }

  int de_standard_largest(de_context *ctx, const de_matrix *A, double shift, double tol, int maxiter, int nev,
                          const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    return standard_driver(ctx, A, nullptr, shift, tol, maxiter, nev, start_panel8, eval, evec, verbose, iterations);
  }

  int de_standard_inverse(de_context *ctx, const de_matrix *A, const de_factor *F, double shift, double tol,
                          int maxiter, int nev, const double *start_panel8, double *eval, double *evec, int verbose,
                          int *iterations)
  {
    if (!F)
      return set_error(ctx, DE_ERR_INVALID, "de_standard_inverse: factorisation is null");
    return standard_driver(ctx, A, F, shift, tol, maxiter, nev, start_panel8, eval, evec, verbose, iterations);
  }

  int de_generalized_inverse(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *F, double shift,
                             double tol, int maxiter, int nev, const double *start_panel8, double *eval, double *evec,
                             int verbose, int *iterations, double *relerror_out)
  {
    if (!ctx || !A || !B || !F || !start_panel8 || !eval || !evec || nev <= 0)
      return set_error(ctx, DE_ERR_INVALID, "de_generalized_inverse: bad arguments");
    const int m = padded_cols(nev);
    if (!valid_cols(m))
      return set_error(ctx, DE_ERR_UNSUPPORTED, "de_generalized_inverse: nev exceeds DE_MAX_COLS (64)");
    if (A->n != B->n || F->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "de_generalized_inverse: A, B and the factorisation must have the same size");
    DE_TRY(bind_device(ctx));
    const long long n = A->n;
    ScopedBlocks blk;
    double *Q1, *Q2, *BQ;
    DE_TRY(blk.alloc(ctx, &Q1, (size_t)n * m));
    DE_TRY(blk.alloc(ctx, &Q2, (size_t)n * m));
    DE_TRY(blk.alloc(ctx, &BQ, (size_t)n * m));
    DE_TRY(reset_status(ctx));
    DE_TRY(upload_panel8_device(ctx, n, m, start_panel8, Q1));
    std::vector<double> ra1(m, 0.0), ra2(m, 0.0), sA(m, 0.0);
    DE_TRY(b_orthonormalize_device(ctx, B, n, m, Q1, BQ, false)); // (:270) ; BQ = B*Q1
    DE_TRY(spmm_device<true>(ctx, A, Q1, Q2, m));                 // (:271-272)
    DE_TRY(fetch_small(ctx, ctx->dDP(), sA.data(), m));
    for (int i = 0; i < m; ++i)
      ra2[i] = sA[i] - shift;
    int iter = 0;
    double relerror = 0.0;
    while (iter < maxiter)
    {
      // Q2 = B*Q1 (:295) is the product BQ that B-orthonormalisation carried along; Q1 = A^-1 * (B*Q1) (:296)
      DE_TRY(factor_apply_device(ctx, F, BQ, Q1, m));
      DE_TRY(b_orthonormalize_device(ctx, B, n, m, Q1, BQ, false)); // (:297)
      iter += 1;
      DE_TRY(spmm_device<true>(ctx, A, Q1, Q2, m)); // (:308-309)
      DE_TRY(fetch_small(ctx, ctx->dDP(), sA.data(), m));
      relerror = 0.0;
      for (int i = 0; i < m; ++i)
      {
        ra1[i] = sA[i] - shift;
        relerror = std::max(relerror, std::abs(ra1[i] - ra2[i]));
      }
      relerror /= *std::max_element(ra1.begin(), ra1.end());
      if (verbose > 2)
        std::printf("iter=%d relerror=%g\n", iter, relerror);
      std::swap(ra1, ra2);
      if (iter > 10 && relerror < tol) // (:323)
        break;
    }
    if (iterations)
      *iterations = iter;
    if (relerror_out)
      *relerror_out = relerror;
    return copy_out(ctx, n, m, nev, Q1, ra2, eval, evec);
  }


  // ---- LOBPCG drivers (new; no reference counterpart) ----------------------------------------
  int de_standard_lobpcg(de_context *ctx, const de_matrix *A, double tol, int maxiter, int nev,
                         const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    return lobpcg_driver(ctx, "StandardLOBPCG", A, nullptr, nullptr, tol, maxiter, nev, start_panel8, eval, evec, verbose,
                         iterations);
  }

  int de_generalized_lobpcg(de_context *ctx, const de_matrix *A, const de_matrix *B, double tol, int maxiter, int nev,
                            const double *start_panel8, double *eval, double *evec, int verbose, int *iterations)
  {
    if (!B)
      return set_error(ctx, DE_ERR_INVALID, "GeneralizedLOBPCG: bad arguments");
    return lobpcg_driver(ctx, "GeneralizedLOBPCG", A, B, nullptr, tol, maxiter, nev, start_panel8, eval, evec, verbose,
                         iterations);
  }

  int de_lobpcg_mv(de_context *ctx, const de_matrix *A, const de_matrix *B, const de_factor *T, int largest,
                   int cheb_degree, double tol, int maxiter, int nev, de_mv *Q, double *eval_m, double *resnorm_m, int verbose, int *iterations,
                   int *restarts, int *converged)
  {
    if (!ctx || !A || !Q || !eval_m)
      return set_error(ctx, DE_ERR_INVALID, "de_lobpcg_mv: bad arguments");
    if (Q->n != A->n)
      return set_error(ctx, DE_ERR_INVALID, "de_lobpcg_mv: the block does not match the matrix");
    DE_TRY(lobpcg_check_args(ctx, "de_lobpcg_mv", A, B, T, nev, Q->m));
    DE_TRY(bind_device(ctx));
    de::LobpcgResult res;
    DE_TRY(lobpcg_device(ctx, A, B, T, largest != 0, cheb_degree, tol, maxiter, nev, Q->m, Q->d, res, verbose));
    for (int j = 0; j < Q->m; ++j)
    {
      eval_m[j] = res.theta[j];
      if (resnorm_m)
        resnorm_m[j] = res.resnorm[j];
    }
    if (iterations)
      *iterations = res.iterations;
    if (restarts)
      *restarts = res.restarts;
    if (converged)
      *converged = res.converged ? 1 : 0;
    return DE_OK;
  }

  int de_block_lincomb(de_mv *out, de_mv *out2, int ns, const de_mv *const *S, const double *C_host)
  {
    if (!out || !S || !C_host || ns < 1 || ns > 3)
      return set_error(nullptr, DE_ERR_INVALID, "de_block_lincomb: bad arguments");
    de_context *ctx = out->ctx;
    for (int s = 0; s < ns; ++s)
      if (!S[s] || S[s]->n != out->n || S[s]->m != out->m)
        return set_error(ctx, DE_ERR_INVALID, "de_block_lincomb: blocks must have the same shape");
    if (out2 && (out2->n != out->n || out2->m != out->m || out2->d == S[0]->d || out2->d == out->d))
      return set_error(ctx, DE_ERR_INVALID, "de_block_lincomb: out2 must have the same shape and alias neither out nor S[0]");
    for (int s = 1; s < ns; ++s)
      if (out->d == S[s]->d)
        return set_error(ctx, DE_ERR_INVALID, "de_block_lincomb: out may alias S[0] only");
    DE_TRY(bind_device(ctx));
    const int m = out->m;
    ScopedBlocks tmp;
    double *dC = nullptr;
    DE_TRY(tmp.alloc(ctx, &dC, (size_t)3 * m * m));
    DE_CUDA(ctx, cudaMemcpyAsync(dC, C_host, sizeof(double) * (size_t)ns * m * m, cudaMemcpyHostToDevice, ctx->stream));
    const double *src[3] = {S[0]->d, ns > 1 ? S[1]->d : nullptr, ns > 2 ? S[2]->d : nullptr};
    DE_TRY(lincomb_device(ctx, m, out->n, ns, src, dC, out->d, (out2 && ns > 1) ? out2->d : nullptr));
    DE_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // dC returns to the allocator, C_host to the caller
    return DE_OK;
  }

  int de_host_sym_eig(int n, const double *A, double *w, double *V)
  {
    if (n < 0 || (n > 0 && (!A || !w || !V)))
      return set_error(nullptr, DE_ERR_INVALID, "de_host_sym_eig: bad arguments");
    return de::hosteig::sym_eig(n, A, w, V) == 0 ? DE_OK
                                                  : set_error(nullptr, DE_ERR_SINGULAR, "de_host_sym_eig: QL iteration failed");
  }

  int de_host_sym_gen_eig(int n, const double *GA, const double *GB, double *w, double *C, double *min_pivot)
  {
    if (n < 0 || (n > 0 && (!GA || !GB || !w || !C)))
      return set_error(nullptr, DE_ERR_INVALID, "de_host_sym_gen_eig: bad arguments");
    const int rc = de::hosteig::sym_gen_eig(n, GA, GB, w, C, 0.0, min_pivot);
    return rc == 0 ? DE_OK : set_error(nullptr, DE_ERR_SINGULAR, "de_host_sym_gen_eig: GB is not positive definite");
  }

  // ---- host-side helpers ---------------------------------------------------------------------------------------
  int de_start_block(int64_t n, int m, unsigned seed, double *out)
  {
    if (n < 0 || m <= 0 || m % 8 != 0 || !out)
      return set_error(nullptr, DE_ERR_INVALID, "de_start_block: bad arguments");
    std::mt19937 urbg{seed};
    std::normal_distribution<double> gen{0.0, 1.0};
    for (int64_t bj = 0; bj < m; bj += 8)
      for (int64_t i = 0; i < n; ++i)
        for (int j = 0; j < 8; ++j)
          out[(bj / 8 * n + i) * 8 + j] = gen(urbg);
    return DE_OK;
  }

  int de_host_factorize(int64_t n, const int64_t *rowptr, const int64_t *col, const double *val, int ordering,
                        int scale_rows, de_host_factor **out)
  {
    if (!out || n < 0 || !rowptr)
      return set_error(nullptr, DE_ERR_INVALID, "de_host_factorize: bad arguments");
    *out = nullptr;
    de_host_factor *F = new de_host_factor();
    try
    {
      de_b200::factorize_csr((long)n, rowptr, col, val, F->F, (de_b200::Ordering)ordering, scale_rows != 0);
    }
    catch (const std::exception &e)
    {
      delete F;
      const std::string msg = e.what();
      return set_error(nullptr, msg.find("singular") != std::string::npos ? DE_ERR_SINGULAR : DE_ERR_INVALID, msg);
    }
    *out = F;
    return DE_OK;
  }

  int de_host_factor_arrays(const de_host_factor *F, int64_t *n, int64_t *lnz, int64_t *unz, const long **Lp,
                            const long **Lj, const double **Lx, const long **Up, const long **Ui, const double **Ux,
                            const long **P, const long **Q, const double **Rs, long *do_recip)
  {
    if (!F)
      return set_error(nullptr, DE_ERR_INVALID, "null host factor");
    const de_b200::FactorArrays &A = F->F;
    if (n)
      *n = A.n;
    if (lnz)
      *lnz = A.lnz;
    if (unz)
      *unz = A.unz;
    if (Lp)
      *Lp = A.Lp.data();
    if (Lj)
      *Lj = A.Lj.data();
    if (Lx)
      *Lx = A.Lx.data();
    if (Up)
      *Up = A.Up.data();
    if (Ui)
      *Ui = A.Ui.data();
    if (Ux)
      *Ux = A.Ux.data();
    if (P)
      *P = A.P.data();
    if (Q)
      *Q = A.Q.data();
    if (Rs)
      *Rs = A.Rs.data();
    if (do_recip)
      *do_recip = A.do_recip;
    return DE_OK;
  }

  int de_host_factor_destroy(de_host_factor *F)
  {
    delete F;
    return DE_OK;
  }

} // extern "C"
