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
 */

#include <stdexcept>
#include <vector>

#include "sparse_lu.hh"

template <typename ISTLM>
class UMFPackFactorizedMatrix
{
  de_b200::FactorArrays store_;

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
                                   de_b200::Ordering ordering = de_b200::Ordering::nested_dissection)
  {
    using block_type = typename ISTLM::block_type;
    if (A.N() != A.M() || block_type::rows != block_type::cols)
      throw std::invalid_argument("UMFPackFactorizedMatrix: input matrix must be square");
    // k x k blocks: the scalar matrix they denote (the reference flattens blocks the same way, umfpacktools.hh:62-95)
    const int k = block_type::rows;
    const long ns = static_cast<long>(A.N()) * k;
    std::vector<long> rowptr(ns + 1, 0), col;
    std::vector<double> val;
    col.reserve(A.nonzeroes() * (std::size_t)k * k);
    val.reserve(A.nonzeroes() * (std::size_t)k * k);
    for (auto row = A.begin(); row != A.end(); ++row)
      for (int r = 0; r < k; ++r)
      {
        for (auto entry = row->begin(); entry != row->end(); ++entry)
          for (int c = 0; c < k; ++c)
          {
            col.push_back(static_cast<long>(entry.index()) * k + c);
            val.push_back(static_cast<double>((*entry)[r][c]));
          }
        rowptr[row.index() * k + r + 1] = static_cast<long>(col.size());
      }
    de_b200::factorize_csr(ns, rowptr.data(), col.data(), val.data(), store_, ordering);
    publish();
    (void)verbose;
  }

  //! adopt factor arrays computed elsewhere (e.g. by UMFPACK at a site that has it)
  explicit UMFPackFactorizedMatrix(de_b200::FactorArrays &&factors) : store_(std::move(factors)) { publish(); }

  UMFPackFactorizedMatrix(const UMFPackFactorizedMatrix &) = delete;
  UMFPackFactorizedMatrix &operator=(const UMFPackFactorizedMatrix &) = delete;
};

#endif
