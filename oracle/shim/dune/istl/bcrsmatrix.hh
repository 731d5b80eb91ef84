// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product.
//
// Minimal stand-in for dune-istl's BCRSMatrix with the iterator surface the reference's
// hot path uses:
//   A.begin()/end(), row->begin()/end(), row.index(), col.index(), *col,
//   A.N(), A.M(), A.nonzeroes(), A.axpy(s,B), copy construction, block_type
//   (reference eigensolver.hh:32-66,208-252; kernels_cpp.hh:383-392,644-653; umfpacktools.hh:49-95)
// Storage is plain CSR (row pointer, ascending column indices, one block per entry).
// Written from scratch for this repo; dune-istl is not available in the image.
#ifndef DE_ORACLE_SHIM_BCRSMATRIX_HH
#define DE_ORACLE_SHIM_BCRSMATRIX_HH

#include <cstddef>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace Dune
{
  template <class B>
  class BCRSMatrix
  {
  public:
    using block_type = B;
    using size_type = std::size_t;

  private:
    size_type n_ = 0, m_ = 0;
    std::vector<size_type> ptr_; // n_+1 row starts
    std::vector<size_type> col_; // column index per stored block
    std::vector<B> val_;         // the blocks

    // one class for const and mutable traversal
    template <bool is_const>
    struct Walk
    {
      using Mat = typename std::conditional<is_const, const BCRSMatrix, BCRSMatrix>::type;
      using Blk = typename std::conditional<is_const, const B, B>::type;

      struct ColIt
      {
        Mat *a;
        size_type k;
        size_type index() const { return a->col_[k]; }
        Blk &operator*() const { return a->val_[k]; }
        Blk *operator->() const { return &a->val_[k]; }
        ColIt &operator++()
        {
          ++k;
          return *this;
        }
        bool operator!=(const ColIt &o) const { return k != o.k; }
        bool operator==(const ColIt &o) const { return k == o.k; }
      };

      struct Row
      {
        Mat *a;
        size_type i;
        ColIt begin() const { return ColIt{a, a->ptr_[i]}; }
        ColIt end() const { return ColIt{a, a->ptr_[i + 1]}; }
        size_type size() const { return a->ptr_[i + 1] - a->ptr_[i]; }
        Row *operator->() { return this; } // lets RowIt::operator-> chain through a temporary
      };

      struct RowIt
      {
        Mat *a;
        size_type i;
        size_type index() const { return i; }
        Row operator*() const { return Row{a, i}; }
        Row operator->() const { return Row{a, i}; }
        RowIt &operator++()
        {
          ++i;
          return *this;
        }
        bool operator!=(const RowIt &o) const { return i != o.i; }
        bool operator==(const RowIt &o) const { return i == o.i; }
      };
    };

  public:
    using RowIterator = typename Walk<false>::RowIt;
    using ConstRowIterator = typename Walk<true>::RowIt;
    using ColIterator = typename Walk<false>::ColIt;
    using ConstColIterator = typename Walk<true>::ColIt;

    BCRSMatrix() = default;

    //! build from CSR arrays (scalar values become diagonal blocks; only 1x1 is used in practice)
    template <class I, class V>
    BCRSMatrix(size_type n, size_type m, const I *rowptr, const I *col, const V *val)
        : n_(n), m_(m), ptr_(n + 1)
    {
      for (size_type i = 0; i <= n; ++i)
        ptr_[i] = static_cast<size_type>(rowptr[i]);
      const size_type nnz = ptr_[n];
      col_.resize(nnz);
      val_.resize(nnz);
      for (size_type k = 0; k < nnz; ++k)
      {
        col_[k] = static_cast<size_type>(col[k]);
        val_[k] = val[k];
      }
    }

    size_type N() const { return n_; }
    size_type M() const { return m_; }
    size_type nonzeroes() const { return col_.size(); }

    RowIterator begin() { return RowIterator{this, 0}; }
    RowIterator end() { return RowIterator{this, n_}; }
    ConstRowIterator begin() const { return ConstRowIterator{this, 0}; }
    ConstRowIterator end() const { return ConstRowIterator{this, n_}; }

    //! this += s * o ; every entry of o must exist in this (dune-istl semantics)
    BCRSMatrix &axpy(double s, const BCRSMatrix &o)
    {
      if (o.n_ != n_)
        throw std::invalid_argument("BCRSMatrix shim: axpy size mismatch");
      for (size_type i = 0; i < n_; ++i)
      {
        size_type k = ptr_[i];
        for (size_type ko = o.ptr_[i]; ko < o.ptr_[i + 1]; ++ko)
        {
          while (k < ptr_[i + 1] && col_[k] < o.col_[ko])
            ++k;
          if (k == ptr_[i + 1] || col_[k] != o.col_[ko])
            throw std::invalid_argument("BCRSMatrix shim: axpy pattern mismatch");
          val_[k].axpy(s, o.val_[ko]);
        }
      }
      return *this;
    }

    // raw access for the C wrappers around the reference (not used by reference code)
    const std::vector<size_type> &raw_ptr() const { return ptr_; }
    const std::vector<size_type> &raw_col() const { return col_; }
    const std::vector<B> &raw_val() const { return val_; }
    std::vector<B> &raw_val() { return val_; }
  };
} // namespace Dune

#endif
