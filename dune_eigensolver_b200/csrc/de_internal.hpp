// Internal header of libdune_eigensolver_b200.so: the opaque types of the C ABI (include/dune_eigensolver_b200.h),
// the error / launch plumbing and the functions the translation units call in each other.
//
//   de_runtime.cu   contexts, caching device allocator, pinned transfer engine, multivectors, host helpers
//   de_spmm.cu      matrices (CSR + BRB forms, distributed parts) and the SpMM launch logic
//   de_dense.cu     reductions with fused tails, Gram / block update / CholQR2, LOBPCG combination kernels
//   de_trsv.cu      factor schedules and the factored apply
//   de_drivers.cu   the device-resident driver loops (reference eigensolver.hh:28-112, :116-198, :204-351) and LOBPCG
//   de_multi.cu     single-process multi-GPU front end
// No CPU fallback exists: every compute entry point needs a CUDA device and fails with DE_ERR_CUDA otherwise.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
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
#include "de_types.hpp"

namespace dei
{
  constexpr int kMaxPartials = 592;               // CTAs of a reduction kernel (4 per SM on 148 SMs)
  constexpr size_t kPartialDoubles = (size_t)2 * kMaxPartials * DE_KERNEL_MAX_M * DE_KERNEL_MAX_M;
  constexpr int kSmall = 5 * DE_KERNEL_MAX_M * DE_KERNEL_MAX_M + 1024; // doubles of small device / pinned scratch
  void dev_free(void *p);
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

NcclApi &nccl_api();

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
  int *hflags = nullptr;      // pinned: 2 slots x 4 ints, polled copies of dflags; [8] = copy of the peer error flag (fetch_small)
  cudaEvent_t ev_poll[2] = {nullptr, nullptr};
  double *dconv = nullptr;    // s_prev[64] | hist[dconv_cap]
  void *xfer = nullptr;       // XferEngine: pinned staging buffers and copy streams (created on first use)
  // NVLink peer window (kernels_peer.cuh); peer_ready once every rank's window is mapped
  bool peer_ready = false;
  bool pdl = true;      // programmatic dependent launch of the loop kernels (off when ranks share a device)
  // Chebyshev epilogue of the next BRB SpMM launch (spmm_cheb_device, de_spmm.cu)
  struct SpmmEpilogue
  {
    bool valid = false;
    double *zold = nullptr;
    const double *r = nullptr, *dinv = nullptr;
    double alpha = 0.0, beta = 0.0;
  } epi;
  bool use_cheb_epilogue = true; // option "cheb_epilogue" 0: SpMM + cheb_step_kernel as two passes (A/B measurements)
  bool use_lincomb2 = true;   // LOBPCG combination / projection on the tensor-core kernel (kernels_lincomb2.cuh); option "lincomb2" 0: first version
  bool use_loop_graph = true; // StandardLargest: replay the steady-state iterations from a CUDA graph (one GPU)
  long long peer_timeout_cycles = 60000000000LL; // spins on peer flags give up after this many clocks (~30 s)
  bool peer_ipc = true; // peer windows were opened from CUDA IPC handles (else: same-process allocations, de_multi.cu)
  unsigned char *window = nullptr;
  size_t window_bytes = 0, halo_cap = 0;
  unsigned char *peer_base[de::kPeerMaxRanks] = {};
  unsigned long long ar_epoch = 0, halo_epoch = 0;
  unsigned long long ar_epoch_b = 0; // all-reduce channel 1 (de::PeerArgs::channel)
  int *dwell = nullptr;              // device int[2]: one CholQR sweep is enough / is not (chol_inverse2_body)
  bool use_one_sweep = true;         // option "one_sweep" 0: always two sweeps with the a-posteriori test (A/B measurements)
  // halo rows stored by the block-update kernels of an orthonormalisation (plan_fused_push, de_spmm.cu): push_pending is
  // what the next ts2_update launches add to their arguments; prepushed_* say which SpMM call finds its halo rows already
  // in the neighbours' windows (that call only releases the flags)
  de::PushRanges push_pending{};
  // OFF by default -- measured on B200 (profiles/README.md, round 2): the 64-byte row segments a warp of the update kernel
  // stores from its registers cross NVLink far less efficiently than halo_push_kernel's coalesced 16-byte-per-thread rows:
  // 100^3 on 2 GPUs 13.17 -> 13.04 ms (-1 %), 256^3 on 2 GPUs 0.1065 -> 0.111 s (+4 %). de_context_set_option("fused_push", 1) enables it.
  bool fused_push = false;
  const double *prepushed_X = nullptr;
  const de_matrix *prepushed_A = nullptr;
  unsigned long long prepushed_epoch = 0;
  int prepushed_m = 0;
  bool prepushed_released = false; // ... and the flags were raised by the last update launch (PushRanges::release)
  int *dticket = nullptr; // [0] ticket of halo_push_kernel, [1] peer error flag
  // fused tail of the NEXT partial-sum reduction (kernels_tail.cuh): set by the caller, consumed by reduce_partials
  de::TailArgs tail{};
  mutable bool tail_armed = false; // cleared by any error return (set_error), so a failed call cannot leave it behind
  mutable bool tail_did_allreduce = false, tail_did_op = false; // one-shot: the following allreduce_sum / chol / convergence is skipped
  // Rayleigh-quotient partials of the last SpMM whose reduction + convergence test were deferred into the tail of the
  // NEXT Gram reduction (asynchronous driver loop: one all-reduce per iteration less); see reduce_partials
  bool defer_dot = false;
  struct PendingDot
  {
    bool valid = false;
    int nparts = 0, m = 0;
    de::TailArgs conv{};
  };
  mutable PendingDot pending_dot;
  int *dtail_ticket = nullptr;
  size_t dconv_cap = 0;
  double *hsmall = nullptr;   // pinned mirror of dsmall
  int *hstatus = nullptr;     // pinned
  double *stage = nullptr;    // layout-conversion staging
  size_t stage_bytes = 0;
  long long launches = 0;
  // kernels that already have their dynamic-shared-memory opt-in / occupancy on THIS context's device (the attribute
  // is per device, so the state cannot be a function-local static)
  std::unordered_map<const void *, int> func_smem, func_occ;
  // children (matrices, multivectors, factors) keep the context alive: de_context_destroy on a context that still
  // has children only marks it; the last child to go really destroys it
  int children = 0;
  bool zombie = false;
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
  double *d_user = nullptr; // block handed out by de_mv_device_ptr (results are copied back into it if a driver swapped buffers)
};


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
      dei::dev_free(rowptr);
      dei::dev_free(col);
      dei::dev_free(val);
    }
    dei::dev_free(rowmap);
    dei::dev_free(blk_meta);
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
    dei::dev_free(tile);
    dei::dev_free(blob);
    dei::dev_free(ucol);
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
  std::vector<long long> send_first; // [npeers] first of the CONSECUTIVE rows sent to peer p, -1 if they are not consecutive
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
  struct de_sn_device *sn = nullptr; // supernodal form (de_snode.cu); the level schedules L / U are then empty
  double *W2 = nullptr;              // second work block of the supernodal apply
  // the two triangular sweeps (hundreds of dependent launches on the fixed work block W) captured once per width
  cudaGraphExec_t sweep_graph = nullptr;
  int sweep_graph_m = 0;
  long long sweep_graph_nodes = 0;
};

namespace de_b200
{
  struct SupernodalFactor; // supernodal_cholesky.hh (only de_snode.cu needs the definition)
}
struct de_host_factor
{
  de_b200::FactorArrays F;                       // the UMFPACK field contract (sparse_lu.hh), or ...
  std::shared_ptr<de_b200::SupernodalFactor> sn; // ... a supernodal Cholesky factor (F is then filled on demand)
};

namespace dei
{
  int set_error(const de_context *ctx, int code, const std::string &msg);
  const std::string &thread_error();

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

  inline bool valid_cols(int m) { return m > 0 && m % 8 == 0 && m <= DE_MAX_COLS; }

  /** launch with programmatic stream serialization: the kernel's CTAs may be scheduled while the preceding kernel of the
   *  stream drains; the kernel itself waits for that kernel's completion in pdl_prologue() (kernels_sparse.cuh) */
  template <class... KArgs, class... Args>
  cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                         Args &&...args)
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
    // pdl == false (several ranks sharing ONE device, de_multi.cu): CTAs that sit resident in griddepcontrol.wait behind a
    // reduction tail spinning on a peer's flag would keep that peer's kernels off the SMs -- a deadlock
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
  }



  // ---- de_runtime.cu ------------------------------------------------------------------------------------------
  int bind_device(const de_context *ctx);
  /** caching device allocator (blocks keyed by device, stream and size) */
  int dev_alloc_bytes(de_context *ctx, void **p, size_t bytes);
  template <class T>
  int dev_alloc(de_context *ctx, T **p, size_t count)
  {
    return dev_alloc_bytes(ctx, (void **)p, std::max<size_t>(count, 1) * sizeof(T));
  }
  void dev_cache_trim(int device, cudaStream_t stream);
  /** dst[i] = (T) src[i] host -> device through the pinned transfer engine; instantiated for (int <- int64_t) and
   *  (double <- double). `range` (optional) receives min and max of the source values. */
  template <class T, class S>
  int upload_parallel(de_context *ctx, T *dst, const S *src, size_t count, long long *range = nullptr);
  int download_parallel(de_context *ctx, void *dst, const void *src, size_t bytes);
  /** small arrays: convert on the host, allocate, copy, wait */
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
  int ensure_stage(de_context *ctx, size_t bytes);
  int convert_layout(de_context *ctx, long long n, int m, const double *src, double *dst, int to_rowmajor);
  int upload_panel8_device(de_context *ctx, long long n, int m, const double *host, double *dst);
  int copy_out(de_context *ctx, long long n, int m, int nev, const double *Q, const std::vector<double> &s, double *eval,
               double *evec);
  /** children keep their context alive (see de_context::children) */
  void context_retain(de_context *ctx);
  void context_release(de_context *ctx);

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

  // ---- de_dense.cu --------------------------------------------------------------------------------------------
  inline bool ts_supported(int w) { return w == 8 || w == 16 || w == 32 || w == 64; }
  de::PeerArgs peer_args(de_context *ctx, unsigned long long epoch);
  /** doubles at the head of ctx->partials reserved for deferred Rayleigh-quotient partials */
  constexpr size_t kDotPartialsReserve = (size_t)kMaxPartials * DE_KERNEL_MAX_M;
  /** where a reduction kernel launched NOW must leave its per-CTA partials */
  inline double *reduction_partials(de_context *ctx)
  {
    return ctx->partials + (ctx->pending_dot.valid ? kDotPartialsReserve : 0);
  }
  int reduce_partials(de_context *ctx, const double *partials, int nparts, int len, double *out);
  /** reduce + all-reduce + convergence test of deferred Rayleigh-quotient partials in a launch of their own */
  int flush_pending_dot(de_context *ctx);
  int allreduce_sum(de_context *ctx, double *buf, size_t count);
  int diag_dot_device(de_context *ctx, long long n, int m, const double *X, const double *Y, double *out);
  int gram_device(de_context *ctx, int w, long long n, const double *X, int ldx, const double *Y, int ldy, bool symmetric,
                  double *out);
  /** de_dense64.cu: G = X^T Y on the warp-specialised tensor-core kernel (widths where it beats the first-generation one) */
  int lincomb2_device(de_context *ctx, int w, long long n, int ns, const double *const *S, const double *const *Cm, double *out,
                      double *out2, bool identity0, double alpha);
  bool gram2_supported(int w);
  int gram2_device(de_context *ctx, int w, long long n, const double *X, int ldx, const double *Y, int ldy, double *out);
  /** mode 0: Y = X R ; mode 1: Y -= X R (projection) */
  int update_device(de_context *ctx, int mode, int w, long long n, const double *X, int ldx, const double *R, double *Y,
                    int ldy, int upper, const int *skip_flag = nullptr);
  int chol_inverse(de_context *ctx, int m, const double *G, double *Rinv, double *info, int *identity_flag = nullptr);
  void arm_chol_tail(de_context *ctx, int m, double *Rinv, double *info, int *identity_flag);
  int orthonormalize_device(de_context *ctx, long long n, int m, double *X, const double *G_ready = nullptr,
                            const de_matrix *next_spmm = nullptr);
  bool plan_fused_push(de_context *ctx, const de_matrix *A, int m);
  int b_orthonormalize_device(de_context *ctx, const de_matrix *B, long long n, int m, double *X, double *BX, bool want_info);
  int reset_status(de_context *ctx);
  int fetch_small(de_context *ctx, const double *dsrc, double *hdst, size_t count);

  // ---- de_spmm.cu ---------------------------------------------------------------------------------------------
  /** Y = A X; dot: also dp = diag(X^T Y) into ctx->dDP(); gram_out (dot only): may receive G = Y^T Y, see spmm_device */
  void brb_set_plane_points(long long points); // process-wide: brb::plane_points_setting() (brb_format.hpp)
  int spmm_device(de_context *ctx, const de_matrix *A, const double *X, double *Y, int m, bool dot, bool *gram_out = nullptr);
  int spmm_cheb_device(de_context *ctx, const de_matrix *A, const double *Z, double *Zold, const double *R, const double *dinv,
                       double alpha, double beta, int m, bool *fused);

  // ---- de_snode.cu --------------------------------------------------------------------------------------------
  int sn_apply_device(de_context *ctx, de_factor *F, const double *X, double *Y, int m);
  void sn_release(struct de_sn_device *S);
  /** fill H->F (UMFPACK field contract) from the supernodal factor H->sn */
  int sn_expand_contract(de_host_factor *H);

  // ---- de_trsv.cu ---------------------------------------------------------------------------------------------
  int factor_apply_device(de_context *ctx, const de_factor *F, const double *X, double *Y, int m);

  /** debugging aid (DE_TRACE_SETUP=1): where the time of a matrix setup goes. No behaviour depends on it. */
  struct SetupTrace
  {
    const char *who;
    int rank;
    bool on;
    std::chrono::steady_clock::time_point t0;
    SetupTrace(const char *w, int r) : who(w), rank(r), on(std::getenv("DE_TRACE_SETUP") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void lap(const char *what)
    {
      if (on)
        std::fprintf(stderr, "[de setup] rank %d %s: %-26s at %8.2f ms\n", rank, who, what,
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() * 1e3);
    }
  };

  /** Every kernel this library can launch, collected at LOAD time: each launch site names its kernel through DE_REG /
   *  DE_KERNEL, which odr-uses a static member of KernelReg<&kernel> whose initialiser adds the kernel to the registry.
   *  preload_kernels() then forces the (lazily loaded, CUDA_MODULE_LOADING) device code of all of them onto the current
   *  device. Needed where several ranks live in ONE process (de_multi.cu): the first launch of a not-yet-loaded kernel
   *  blocks while any kernel of the process is resident on the device (measured: tools/micro/spin_probe.cu), and a
   *  reduction tail spinning on a peer's flag would wait for exactly that peer forever. */
  void register_kernel(const void *func);
  int preload_kernels(de_context *ctx);
  template <auto K>
  struct KernelReg
  {
    static inline const bool reg = (register_kernel((const void *)K), true);
    static auto get()
    {
      (void)reg;
      return K;
    }
  };
#define DE_REG(...) ((void)dei::KernelReg<&__VA_ARGS__>::reg)
#define DE_KERNEL(...) (dei::KernelReg<&__VA_ARGS__>::get())

  /** dynamic shared memory opt-in of a kernel, once per (context, kernel): the attribute is per DEVICE */
  int ensure_func_smem(de_context *ctx, const void *func, size_t bytes);
  /** resident CTAs per SM of a kernel on this context's device (cached per context) */
  int func_occupancy(de_context *ctx, const void *func, int threads, size_t smem, int *out);
}
