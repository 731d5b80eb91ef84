#ifndef DUNE_EIGENSOLVER_B200_RUNTIME_HH
#define DUNE_EIGENSOLVER_B200_RUNTIME_HH

/** \file
 *  Thin C++ RAII layer between the drop-in header templates and the C ABI (dune_eigensolver_b200.h).
 *  Nothing here computes: it flattens an ISTL-style matrix to CSR through the same iterator surface the
 *  reference uses (eigensolver.hh:61-65, kernels_cpp.hh:644-653), owns the opaque handles, and turns status
 *  codes back into the exception types the reference throws (std::invalid_argument for shape / block-size /
 *  singular-matrix violations).
 */

#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../dune_eigensolver_b200.h"
#include "multivector.hh"

namespace de_b200
{
  inline void check(int status, const de_context *ctx = nullptr)
  {
    if (status == DE_OK)
      return;
    const char *msg = de_last_error_string(ctx);
    const std::string text = (msg && *msg) ? msg : ("dune-eigensolver-b200: status " + std::to_string(status));
    if (status == DE_ERR_INVALID || status == DE_ERR_SINGULAR)
      throw std::invalid_argument(text);
    throw std::runtime_error(text);
  }

  //! one GPU context per host thread (the reference's functions are re-entrant per thread)
  class Context
  {
    de_context *h_ = nullptr;

  public:
    explicit Context(int device = 0) { check(de_context_create(device, nullptr, &h_)); }
    ~Context() { de_context_destroy(h_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    de_context *get() const { return h_; }

    static Context &thread_default()
    {
      thread_local Context ctx(default_device());
      return ctx;
    }
    static int &default_device()
    {
      static int dev = 0;
      return dev;
    }
  };

  /** Process-wide multi-GPU configuration of the drop-in drivers (new key `parallel.numgpus` of the ini, read by the
   *  driver program; the reference has `parallel.numthreads` for its replica harness only, src/dune-eigensolver.cc:757).
   *  With more than one device StandardLargest / StandardLOBPCG / GeneralizedLOBPCG run row-partitioned through the
   *  single-process front end of the C ABI (de_multi_*): one library thread per GPU, halo rows and the small reductions
   *  over NVLink peer memory. The factored drivers stay on one GPU (triangular solves do not row-shard). */
  class Parallel
  {
    std::vector<int> devices_{0};
    std::int64_t row_align_ = 1;
    std::int64_t halo_bytes_ = 0;
    de_multi *multi_ = nullptr;

    Parallel() = default;

  public:
    ~Parallel() { de_multi_destroy(multi_); }
    Parallel(const Parallel &) = delete;
    Parallel &operator=(const Parallel &) = delete;
    static Parallel &instance()
    {
      static Parallel p;
      return p;
    }
    //! CUDA ordinals to use, e.g. {0,1,2,3}; one entry = the single-GPU path (Context::default_device follows it)
    void set_devices(const std::vector<int> &devices)
    {
      if (devices.empty())
        throw std::invalid_argument("Parallel::set_devices: empty device list");
      de_multi_destroy(multi_);
      multi_ = nullptr;
      devices_ = devices;
    }
    //! use the first n devices of the machine
    void set_num_gpus(int n)
    {
      std::vector<int> d;
      for (int i = 0; i < n; ++i)
        d.push_back(i);
      set_devices(d);
    }
    //! cut the rows only at multiples of this many rows (one grid plane of a lexicographic 3D grid keeps halos thin)
    void set_row_align(std::int64_t rows) { row_align_ = rows < 1 ? 1 : rows; }
    void set_halo_bytes(std::int64_t bytes)
    {
      halo_bytes_ = bytes;
      de_multi_destroy(multi_);
      multi_ = nullptr;
    }
    int num_gpus() const { return (int)devices_.size(); }
    int first_device() const { return devices_[0]; }
    std::int64_t row_align() const { return row_align_; }
    de_multi *multi()
    {
      if (!multi_)
        check(de_multi_create(devices_.data(), (int)devices_.size(), halo_bytes_, &multi_));
      return multi_;
    }
    void check_multi(int status) const
    {
      if (status == DE_OK)
        return;
      const char *msg = de_multi_last_error(multi_);
      const std::string text = (msg && *msg) ? msg : ("dune-eigensolver-b200: status " + std::to_string(status));
      if (status == DE_ERR_INVALID || status == DE_ERR_SINGULAR)
        throw std::invalid_argument(text);
      throw std::runtime_error(text);
    }
  };

  //! (B)CSR copy of an ISTL-style matrix with k x k blocks, taken through its row / column iterators. k = 1: plain CSR.
  //! For k > 1 the block arrays are kept (C ABI de_matrix_create_bcsr) and scalar() gives the (N k) x (N k) matrix the
  //! blocks denote -- row i*k + r, column j*k + c -- for the host-side consumers (factorisation, multi-GPU front end).
  struct HostCsr
  {
    std::vector<std::int64_t> rowptr, col;
    std::vector<double> val; // k*k values per stored block, row-major inside a block
    std::int64_t n = 0;      // SCALAR size: block rows * k
    int k = 1;

    template <class ISTLM>
    explicit HostCsr(const ISTLM &A)
    {
      using block_type = typename ISTLM::block_type;
      k = block_type::rows;
      const std::int64_t nb = static_cast<std::int64_t>(A.N());
      n = nb * k;
      rowptr.assign(nb + 1, 0);
      col.reserve(A.nonzeroes());
      val.reserve(A.nonzeroes() * (std::size_t)k * k);
      for (auto row = A.begin(); row != A.end(); ++row)
      {
        for (auto entry = row->begin(); entry != row->end(); ++entry)
        {
          col.push_back(static_cast<std::int64_t>(entry.index()));
          for (int r = 0; r < block_type::rows; ++r)
            for (int c = 0; c < block_type::cols; ++c)
              val.push_back(static_cast<double>((*entry)[r][c]));
        }
        rowptr[row.index() + 1] = static_cast<std::int64_t>(col.size());
      }
    }
    HostCsr() = default;

    std::int64_t block_rows() const { return (std::int64_t)rowptr.size() - 1; }

    //! the scalar CSR matrix (a copy for k > 1, *this for k = 1)
    HostCsr scalar() const
    {
      if (k == 1)
        return *this;
      HostCsr S;
      S.k = 1;
      S.n = n;
      const std::int64_t nb = block_rows();
      S.rowptr.assign(n + 1, 0);
      S.col.resize(col.size() * (std::size_t)k * k);
      S.val.resize(S.col.size());
      std::int64_t pos = 0;
      for (std::int64_t ib = 0; ib < nb; ++ib)
        for (int r = 0; r < k; ++r)
        {
          for (std::int64_t e = rowptr[ib]; e < rowptr[ib + 1]; ++e)
            for (int c = 0; c < k; ++c)
            {
              S.col[pos] = col[e] * k + c;
              S.val[pos++] = val[(e * k + r) * k + c];
            }
          S.rowptr[ib * k + r + 1] = pos;
        }
      return S;
    }
  };

  class DeviceMatrix
  {
    de_matrix *h_ = nullptr;

  public:
    DeviceMatrix(Context &ctx, const HostCsr &A)
    {
      check(de_matrix_create_bcsr(ctx.get(), A.block_rows(), (std::int64_t)A.col.size(), A.k, A.rowptr.data(), A.col.data(),
                                  A.val.data(), &h_),
            ctx.get());
    }
    template <class ISTLM>
    DeviceMatrix(Context &ctx, const ISTLM &A) : DeviceMatrix(ctx, HostCsr(A))
    {
    }
    ~DeviceMatrix() { de_matrix_destroy(h_); }
    DeviceMatrix(const DeviceMatrix &) = delete;
    DeviceMatrix &operator=(const DeviceMatrix &) = delete;
    de_matrix *get() const { return h_; }
  };

  class DeviceMV
  {
    de_mv *h_ = nullptr;
    Context &ctx_;

  public:
    DeviceMV(Context &ctx, std::size_t n, std::size_t m) : ctx_(ctx)
    {
      check(de_mv_create(ctx.get(), (std::int64_t)n, (int)m, &h_), ctx.get());
    }
    //! upload a host MultiVector<double,8>
    DeviceMV(Context &ctx, const MultiVector<double, 8> &Q) : DeviceMV(ctx, Q.rows(), Q.cols())
    {
      if (Q.rows() * Q.cols() > 0)
        check(de_mv_upload_panel8(h_, Q.data()), ctx.get());
    }
    ~DeviceMV() { de_mv_destroy(h_); }
    DeviceMV(const DeviceMV &) = delete;
    DeviceMV &operator=(const DeviceMV &) = delete;
    void download(MultiVector<double, 8> &Q) const
    {
      if (Q.rows() * Q.cols() > 0)
        check(de_mv_download_panel8(h_, Q.data()), ctx_.get());
    }
    de_mv *get() const { return h_; }
  };

  template <class MV>
  inline void require_block8(const char *who)
  {
    if (MV::blocksize != 8)
      throw std::invalid_argument(std::string(who) + ": blocksize must be 8");
  }

  //! The reference's kernels throw "only implemented for FieldMatrix<..,1,1>" for every k != 1 (kernels_cpp.hh:362-363,
  //! :632-633). Here square blocks of any size are accepted (BCSR): the matrix acts on vector
  //! blocks with N*k rows. Non-square blocks keep the reference's exception.
  template <class ISTLM>
  inline void require_scalar_blocks(const char *who)
  {
    using block_type = typename ISTLM::block_type;
    if (block_type::rows != block_type::cols)
      throw std::invalid_argument(std::string(who) + ": only implemented for FieldMatrix<..,1,1>");
  }

  //! scalar size of the problem: block rows times block size
  template <class ISTLM>
  inline std::size_t scalar_rows(const ISTLM &A)
  {
    return (std::size_t)A.N() * (std::size_t)ISTLM::block_type::rows;
  }
} // namespace de_b200

#endif
