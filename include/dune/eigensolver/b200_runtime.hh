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

  //! CSR copy of an ISTL-style matrix with 1x1 blocks, taken through its row / column iterators
  struct HostCsr
  {
    std::vector<std::int64_t> rowptr, col;
    std::vector<double> val;
    std::int64_t n = 0;

    template <class ISTLM>
    explicit HostCsr(const ISTLM &A)
    {
      n = static_cast<std::int64_t>(A.N());
      rowptr.assign(n + 1, 0);
      col.reserve(A.nonzeroes());
      val.reserve(A.nonzeroes());
      for (auto row = A.begin(); row != A.end(); ++row)
      {
        for (auto entry = row->begin(); entry != row->end(); ++entry)
        {
          col.push_back(static_cast<std::int64_t>(entry.index()));
          val.push_back(static_cast<double>((*entry)[0][0]));
        }
        rowptr[row.index() + 1] = static_cast<std::int64_t>(col.size());
      }
    }
  };

  class DeviceMatrix
  {
    de_matrix *h_ = nullptr;

  public:
    DeviceMatrix(Context &ctx, const HostCsr &A)
    {
      check(de_matrix_create_csr(ctx.get(), A.n, (std::int64_t)A.col.size(), A.rowptr.data(), A.col.data(),
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

  template <class ISTLM>
  inline void require_scalar_blocks(const char *who)
  {
    using block_type = typename ISTLM::block_type;
    if (block_type::rows != 1 || block_type::cols != 1)
      throw std::invalid_argument(std::string(who) + ": only implemented for FieldMatrix<..,1,1>");
  }
} // namespace de_b200

#endif
