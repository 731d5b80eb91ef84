// de_runtime.cu -- contexts, the caching device allocator, the pinned transfer engine, multivectors and the host-side
// helpers of libdune_eigensolver_b200.so (C ABI: include/dune_eigensolver_b200.h).
#include <dlfcn.h>

#include "de_internal.hpp"
#include "kernels_sparse.cuh" // panel8_convert_kernel, extract_columns_kernel

using namespace dei;

NcclApi &nccl_api()
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


namespace dei
{
  thread_local std::string g_thread_error;
  const std::string &thread_error() { return g_thread_error; }

  int set_error(const de_context *ctx, int code, const std::string &msg)
  {
    g_thread_error = msg;
    if (ctx)
    {
      ctx->err = msg;
      ctx->tail_armed = ctx->tail_did_allreduce = ctx->tail_did_op = false;
      ctx->pending_dot.valid = false;
    }
    return code;
  }


  static std::vector<const void *> &kernel_registry()
  {
    static std::vector<const void *> *v = new std::vector<const void *>(); // used during static initialisation of every unit
    return *v;
  }

  void register_kernel(const void *func) { kernel_registry().push_back(func); }

  int preload_kernels(de_context *ctx)
  {
    DE_TRY(bind_device(ctx));
    for (const void *f : kernel_registry())
    {
      cudaFuncAttributes a;
      DE_CUDA(ctx, cudaFuncGetAttributes(&a, f)); // loads the function's module on this device
    }
    return DE_OK;
  }

  int ensure_func_smem(de_context *ctx, const void *func, size_t bytes)
  {
    auto it = ctx->func_smem.find(func);
    if (it != ctx->func_smem.end() && it->second >= (int)bytes)
      return DE_OK;
    DE_CUDA(ctx, cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    ctx->func_smem[func] = (int)bytes;
    return DE_OK;
  }

  int func_occupancy(de_context *ctx, const void *func, int threads, size_t smem, int *out)
  {
    auto it = ctx->func_occ.find(func);
    if (it == ctx->func_occ.end())
    {
      int occ = 1;
      DE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, func, threads, smem));
      it = ctx->func_occ.emplace(func, std::max(1, occ)).first;
    }
    *out = it->second;
    return DE_OK;
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
    // never destroyed: contexts owned by objects with static storage duration (the drop-in headers' Parallel singleton)
    // are torn down during exit, possibly after the destructors of this library's own statics have run
    static DevCache *c = new DevCache();
    return *c;
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
  int upload_parallel(de_context *ctx, T *dst, const S *src, size_t count, long long *range)
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

  template int upload_parallel<int, int64_t>(de_context *, int *, const int64_t *, size_t, long long *);
  template int upload_parallel<double, double>(de_context *, double *, const double *, size_t, long long *);

  int bind_device(const de_context *ctx)
  {
    DE_CUDA(ctx, cudaSetDevice(ctx->device));
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
    DE_REG(de::panel8_convert_kernel), de::panel8_convert_kernel<<<grid, 256, 0, ctx->stream>>>(n, m, src, dst, to_rowmajor);
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
      DE_REG(de::extract_columns_kernel), de::extract_columns_kernel<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(n, m, nev, Q, ctx->stage);
    }
    DE_LAUNCH_CHECK(ctx);
    return download_parallel(ctx, evec, ctx->stage, bytes);
  }


  void context_retain(de_context *ctx)
  {
    if (ctx)
      ctx->children++;
  }

  void context_release(de_context *ctx)
  {
    if (!ctx)
      return;
    ctx->children--;
    if (ctx->children <= 0 && ctx->zombie)
      de_context_destroy(ctx);
  }

} // namespace dei

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
    // (the second stream of the NCCL halo path is created by de_context_init_comm: one stream per context otherwise)
    bool ok = cudaEventCreateWithFlags(&ctx->ev_pack, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming) == cudaSuccess &&
              cudaMalloc((void **)&ctx->partials, kPartialDoubles * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dsmall, kSmall * sizeof(double)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dstatus, sizeof(int)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dflags, 4 * sizeof(int)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dtail_ticket, sizeof(int)) == cudaSuccess &&
              cudaMalloc((void **)&ctx->dwell, 2 * sizeof(int)) == cudaSuccess &&
              cudaMemset(ctx->dwell, 0, 2 * sizeof(int)) == cudaSuccess &&
              cudaMemset(ctx->dtail_ticket, 0, sizeof(int)) == cudaSuccess &&
              cudaMemset(ctx->dflags, 0, 4 * sizeof(int)) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->hsmall, kSmall * sizeof(double)) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->hstatus, sizeof(int)) == cudaSuccess &&
              cudaMallocHost((void **)&ctx->hflags, 12 * sizeof(int)) == cudaSuccess &&
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
    if (ctx->children > 0)
    {
      // matrices / multivectors / factors created on this context are still alive (e.g. Python objects finalised in
      // arbitrary order): keep the context until the last of them is destroyed (context_release)
      ctx->zombie = true;
      return DE_OK;
    }
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
    if (ctx->dwell)
      cudaFree(ctx->dwell);
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
      if (ctx->peer_base[q] && q != ctx->rank && ctx->peer_ipc)
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

  int de_context_set_option(de_context *ctx, const char *name, int64_t value)
  {
    if (!ctx || !name)
      return set_error(ctx, DE_ERR_INVALID, "de_context_set_option: bad arguments");
    const std::string key(name);
    if (key == "one_sweep")
      ctx->use_one_sweep = value != 0;
    else if (key == "cheb_epilogue")
      ctx->use_cheb_epilogue = value != 0;
    else if (key == "lincomb2")
      ctx->use_lincomb2 = value != 0;
    else if (key == "loop_graph")
      ctx->use_loop_graph = value != 0;
    else if (key == "fused_push")
      ctx->fused_push = value != 0;
    else if (key == "brb_plane_points")
    {
      if (value < 64)
        return set_error(ctx, DE_ERR_INVALID, "de_context_set_option: brb_plane_points must be at least 64");
      brb_set_plane_points((long long)value);
    }
    else
      return set_error(ctx, DE_ERR_INVALID, "de_context_set_option: unknown option '" + key + "'");
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
    if (!ctx->comm_stream)
      DE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
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
    ctx->ar_epoch = ctx->ar_epoch_b = ctx->halo_epoch = 0;
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
    // pass 1 (parallel): the columns outside [lo, hi), per thread; stencil matrices have few of them
    const int nth = (int)std::max<int64_t>(1, std::min<int64_t>(8, nnz >> 20));
    std::vector<int64_t> ext;
    {
      std::vector<std::vector<int64_t>> part_ext((size_t)nth);
      std::vector<int> bad((size_t)nth, 0);
      auto scan = [&](int t)
      {
        const int64_t k0 = nnz * t / nth, k1 = nnz * (t + 1) / nth;
        std::vector<int64_t> &e = part_ext[t];
        int64_t last = -1;
        for (int64_t k = k0; k < k1; ++k)
        {
          const int64_t g = col_global[k];
          if (g < 0 || g >= nglob)
          {
            bad[t] = 1;
            return;
          }
          if ((g < lo || g >= hi) && g != last) // (consecutive duplicates are common: skip them early)
          {
            e.push_back(g);
            last = g;
          }
        }
      };
      std::vector<std::thread> th;
      for (int t = 1; t < nth; ++t)
        th.emplace_back(scan, t);
      scan(0);
      for (auto &x : th)
        x.join();
      for (int t = 0; t < nth; ++t)
      {
        if (bad[t])
          return set_error(nullptr, DE_ERR_INVALID, "de_halo_plan_local: column index out of range");
        ext.insert(ext.end(), part_ext[t].begin(), part_ext[t].end());
      }
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
    context_retain(ctx);
    *out = X;
    return DE_OK;
  }

  int de_mv_destroy(de_mv *X)
  {
    if (!X)
      return DE_OK;
    de_context *ctx = X->ctx;
    cudaSetDevice(ctx->device);
    dev_free(X->d);
    delete X;
    context_release(ctx);
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
    X->d_user = X->d; // the drivers keep results in THIS block (see standard_driver_mv)
    return DE_OK;
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
    if (F->sn && F->F.n == 0 && sn_expand_contract(const_cast<de_host_factor *>(F)) != DE_OK)
      return DE_ERR_UNSUPPORTED;
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
