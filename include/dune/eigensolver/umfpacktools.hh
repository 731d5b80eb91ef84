#ifndef DUNE_EIGENSOLVER_B200_UMFPACKTOOLS_HH
#define DUNE_EIGENSOLVER_B200_UMFPACKTOOLS_HH

/** \file
 *  UMFPackFactorizedMatrix<ISTLM>: the reference's factor-array holder (reference umfpacktools.hh:16-220) with the
 *  same public data members -- n, lnz, unz, n_row, n_col, nz_udiag, Lp, Lj, Lx, Up, Ui, Ux, P, Q, do_recip, Rs --
 *  and the same meaning (L compressed-row with the diagonal last in each row, U compressed-column with the
 *  diagonal last in each column, P[k] = original row of pivot row k, Q[k] = original column of pivot column k,
 *  row i scaled by *Rs[i] if do_recip else /Rs[i]).
 *
 *  The reference fills those arrays through SuiteSparse UMFPACK. That is one-time host setup and stays on the
 *  host. This header fills them with the repo's own sparse LU (sparse_lu.hh) so that the path works where
 *  UMFPACK is absent; a site that has UMFPACK can populate the same members from umfpack_dl_get_numeric and hand
 *  the object to matmul_inverse_tallskinny_blocked / de_factor_upload unchanged.
 *
 *  Large symmetric positive definite matrices (3D pencils) go through the library's second provider, a supernodal
 *  multifrontal Cholesky (de_host_factorize_spd): the factor then lives in supernodal form inside the library, the public
 *  arrays are an expansion of it (L unit lower, U = D L^T, P = Q, Rs = 1) that is only materialised for factors of
 *  moderate size -- beyond that they stay null and upload() / the drivers hand the supernodal factor to the GPU directly.
 */

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../dune_eigensolver_b200.h"
#include "sparse_lu.hh"

namespace de_b200
{
  //! which host provider fills the factor: the scalar LU of sparse_lu.hh, the supernodal Cholesky of the library
  //! (supernodal_cholesky.hh through de_host_factorize_spd; symmetric positive definite matrices only), or -- automatic --
  //! Cholesky for large symmetric matrices with a fall-back to LU when a pivot is not positive
  enum class Provider
  {
    automatic,
    lu,
    cholesky
  };
  constexpr long kCholeskyFromRows = 10000;         // automatic: below this size the LU provider is used
  constexpr long kContractMaxEntries = 150000000L;  // explicit L / U arrays are only materialised up to this many entries
} // namespace de_b200

template <typename ISTLM>
class UMFPackFactorizedMatrix
{
  de_b200::FactorArrays store_;
  de_host_factor *host_ = nullptr; // supernodal Cholesky factor held by the library (Provider::cholesky)

  void publish()
  {
    n = n_row = n_col = store_.n;
    lnz = store_.lnz;
    unz = store_.unz;
    nz_udiag = store_.nz_udiag;
    Lp = store_.Lp.data();
    Lj = store_.Lj.data();
    Lx = store_.Lx.data();
    Up = store_.Up.data();
    Ui = store_.Ui.data();
    Ux = store_.Ux.data();
    P = store_.P.data();
    Q = store_.Q.data();
    Rs = store_.Rs.data();
    do_recip = store_.do_recip;
  }

  //! point the public members at the library's expansion of the supernodal factor (if it is small enough for that)
  void publish_host()
  {
    std::int64_t hn = 0, hl = 0, hs = 0;
    int sn = 0;
    de_host_factor_info(host_, &sn, &hn, &hl, &hs, nullptr, nullptr);
    n = n_row = n_col = (long)hn;
    lnz = unz = (long)hl;
    nz_udiag = (long)hn;
    if (hl > de_b200::kContractMaxEntries)
      return; // the arrays stay null: the factor exists in supernodal form only (upload() / the drivers use it directly)
    const long *lp, *lj, *up, *ui, *p, *q;
    const double *lx, *ux, *rs;
    long rec = 1;
    std::int64_t a, b, c;
    if (de_host_factor_arrays(host_, &a, &b, &c, &lp, &lj, &lx, &up, &ui, &ux, &p, &q, &rs, &rec) != DE_OK)
      return;
    Lp = const_cast<long *>(lp);
    Lj = const_cast<long *>(lj);
    Lx = const_cast<double *>(lx);
    Up = const_cast<long *>(up);
    Ui = const_cast<long *>(ui);
    Ux = const_cast<double *>(ux);
    P = const_cast<long *>(p);
    Q = const_cast<long *>(q);
    Rs = const_cast<double *>(rs);
    do_recip = rec;
    lnz = (long)b;
    unz = (long)c;
  }

public:
  using IntType = long;

  IntType n = 0, lnz = 0, unz = 0, n_row = 0, n_col = 0, nz_udiag = 0;
  IntType *Lp = nullptr, *Lj = nullptr;
  double *Lx = nullptr;
  IntType *Up = nullptr, *Ui = nullptr;
  double *Ux = nullptr;
  IntType *P = nullptr, *Q = nullptr;
  IntType do_recip = 1;
  double *Rs = nullptr;

  //! factorise A (square blocks of any size). `verbose` is accepted for signature compatibility.
  explicit UMFPackFactorizedMatrix(const ISTLM &A, int verbose = 0,
                                   de_b200::Ordering ordering = de_b200::Ordering::nested_dissection,
                                   de_b200::Provider provider = de_b200::Provider::automatic)
  {
    using block_type = typename ISTLM::block_type;
    if (A.N() != A.M() || block_type::rows != block_type::cols)
      throw std::invalid_argument("UMFPackFactorizedMatrix: input matrix must be square");
    // k x k blocks: the scalar matrix they denote (the reference flattens blocks the same way, umfpacktools.hh:62-95)
    const int k = block_type::rows;
    const long ns = static_cast<long>(A.N()) * k;
    std::vector<std::int64_t> rowptr(ns + 1, 0), col;
    std::vector<double> val;
    col.reserve(A.nonzeroes() * (std::size_t)k * k);
    val.reserve(A.nonzeroes() * (std::size_t)k * k);
    for (auto row = A.begin(); row != A.end(); ++row)
      for (int r = 0; r < k; ++r)
      {
        for (auto entry = row->begin(); entry != row->end(); ++entry)
          for (int c = 0; c < k; ++c)
          {
            col.push_back(static_cast<std::int64_t>(entry.index()) * k + c);
            val.push_back(static_cast<double>((*entry)[r][c]));
          }
        rowptr[row.index() * k + r + 1] = static_cast<std::int64_t>(col.size());
      }
    bool try_cholesky = provider == de_b200::Provider::cholesky;
    if (provider == de_b200::Provider::automatic && ns >= de_b200::kCholeskyFromRows)
      try_cholesky = is_symmetric(ns, rowptr, col, val);
    if (try_cholesky)
    {
      const int rc = de_host_factorize_spd(ns, rowptr.data(), col.data(), val.data(), (int)ordering, 0, &host_);
      if (rc == DE_OK)
      {
        publish_host();
        return;
      }
      host_ = nullptr;
      if (provider == de_b200::Provider::cholesky)
        throw std::invalid_argument(std::string("UMFPackFactorizedMatrix: ") + de_last_error_string(nullptr));
      // automatic: not positive definite -> the LU provider
    }
    de_b200::factorize_csr(ns, rowptr.data(), col.data(), val.data(), store_, ordering);
    publish();
    (void)verbose;
  }

  //! adopt factor arrays computed elsewhere (e.g. by UMFPACK at a site that has it)
  explicit UMFPackFactorizedMatrix(de_b200::FactorArrays &&factors) : store_(std::move(factors)) { publish(); }

  ~UMFPackFactorizedMatrix() { de_host_factor_destroy(host_); }

  UMFPackFactorizedMatrix(const UMFPackFactorizedMatrix &) = delete;
  UMFPackFactorizedMatrix &operator=(const UMFPackFactorizedMatrix &) = delete;

  //! true if the factor is a supernodal Cholesky factor held by the library
  bool supernodal() const { return host_ != nullptr; }

  //! device copy of the factor, ready for de_factor_apply and the drivers (caller destroys it with de_factor_destroy)
  int upload(de_context *ctx, de_factor **out) const
  {
    if (host_)
      return de_factor_upload_host(ctx, host_, out);
    return de_factor_upload(ctx, n, Lp, Lj, Lx, Up, Ui, Ux, P, Q, Rs, do_recip, out);
  }

private:
  static bool is_symmetric(long ns, const std::vector<std::int64_t> &rowptr, const std::vector<std::int64_t> &col,
                           const std::vector<double> &val)
  {
    // pattern and values, by binary search in the (sorted) rows; unsorted rows simply fail the test -> LU
    for (long i = 0; i < ns; ++i)
      for (std::int64_t q = rowptr[i]; q < rowptr[i + 1]; ++q)
      {
        const std::int64_t j = col[q];
        if (j <= i)
          continue;
        const std::int64_t *b = col.data() + rowptr[j], *e = col.data() + rowptr[j + 1];
        const std::int64_t *p = std::lower_bound(b, e, (std::int64_t)i);
        if (p == e || *p != i)
          return false;
        const double t = val[p - col.data()];
        if (std::abs(t - val[q]) > 1e-13 * (std::abs(t) + std::abs(val[q])))
          return false;
      }
    return true;
  }
};

#endif
