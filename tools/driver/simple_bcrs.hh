#ifndef DE_DRIVER_SIMPLE_BCRS_HH
#define DE_DRIVER_SIMPLE_BCRS_HH

// A stand-in for dune-istl's BCRSMatrix<FieldMatrix<K,R,C>> for the driver program, with exactly the surface the
// drop-in headers use (the same surface the reference uses, SURVEY.md §8b "Matrix access"): N(), M(), nonzeroes(),
// begin()/end() over rows, row->begin()/end() over stored blocks, index() on both iterators, *entry -> block,
// block[i][j], block_type::rows/cols, axpy(s, B), copy construction. At a DUNE site the real types are used instead.

#include <cstddef>
#include <stdexcept>
#include <vector>

namespace Dune
{
  template <class K, int R, int C>
  class FieldMatrix
  {
    K a_[R][C] = {};

  public:
    static constexpr int rows = R;
    static constexpr int cols = C;
    FieldMatrix() = default;
    FieldMatrix(const K &v)
    {
      for (int i = 0; i < R && i < C; ++i)
        a_[i][i] = v;
    }
    K *operator[](std::size_t i) { return a_[i]; }
    const K *operator[](std::size_t i) const { return a_[i]; }
  };

  template <class B>
  class BCRSMatrix
  {
    std::size_t n_ = 0, m_ = 0;
    std::vector<std::size_t> ptr_, col_;
    std::vector<B> val_;

    template <class Mat, class Blk>
    struct EntryIt
    {
      Mat *a;
      std::size_t k;
      std::size_t index() const { return a->col_[k]; }
      Blk &operator*() const { return a->val_[k]; }
      EntryIt &operator++()
      {
        ++k;
        return *this;
      }
      bool operator!=(const EntryIt &o) const { return k != o.k; }
      bool operator==(const EntryIt &o) const { return k == o.k; }
    };
    template <class Mat, class Blk>
    struct RowView
    {
      Mat *a;
      std::size_t i;
      EntryIt<Mat, Blk> begin() const { return {a, a->ptr_[i]}; }
      EntryIt<Mat, Blk> end() const { return {a, a->ptr_[i + 1]}; }
    };
    template <class Mat, class Blk>
    struct RowIt
    {
      RowView<Mat, Blk> row;
      std::size_t index() const { return row.i; }
      const RowView<Mat, Blk> *operator->() const { return &row; }
      const RowView<Mat, Blk> &operator*() const { return row; }
      RowIt &operator++()
      {
        ++row.i;
        return *this;
      }
      bool operator!=(const RowIt &o) const { return row.i != o.row.i; }
      bool operator==(const RowIt &o) const { return row.i == o.row.i; }
    };

  public:
    using block_type = B;
    using size_type = std::size_t;

    BCRSMatrix() = default;
    template <class I, class V>
    BCRSMatrix(size_type n, size_type m, const I *rowptr, const I *col, const V *val) : n_(n), m_(m), ptr_(n + 1)
    {
      for (size_type i = 0; i <= n; ++i)
        ptr_[i] = (size_type)rowptr[i];
      col_.resize(ptr_[n]);
      val_.resize(ptr_[n]);
      for (size_type k = 0; k < ptr_[n]; ++k)
      {
        col_[k] = (size_type)col[k];
        val_[k] = B(val[k]);
      }
    }
    size_type N() const { return n_; }
    size_type M() const { return m_; }
    size_type nonzeroes() const { return col_.size(); }
    RowIt<BCRSMatrix, B> begin() { return {{this, 0}}; }
    RowIt<BCRSMatrix, B> end() { return {{this, n_}}; }
    RowIt<const BCRSMatrix, const B> begin() const { return {{this, 0}}; }
    RowIt<const BCRSMatrix, const B> end() const { return {{this, n_}}; }

    //! this += s * o; every entry of o must exist in this (dune-istl semantics)
    BCRSMatrix &axpy(double s, const BCRSMatrix &o)
    {
      if (o.n_ != n_)
        throw std::invalid_argument("BCRSMatrix::axpy: size mismatch");
      for (size_type i = 0; i < n_; ++i)
      {
        size_type k = ptr_[i];
        for (size_type ko = o.ptr_[i]; ko < o.ptr_[i + 1]; ++ko)
        {
          while (k < ptr_[i + 1] && col_[k] < o.col_[ko])
            ++k;
          if (k == ptr_[i + 1] || col_[k] != o.col_[ko])
            throw std::invalid_argument("BCRSMatrix::axpy: pattern mismatch");
          for (int r = 0; r < B::rows; ++r)
            for (int c = 0; c < B::cols; ++c)
              val_[k][r][c] += s * o.val_[ko][r][c];
        }
      }
      return *this;
    }
  };
} // namespace Dune

#endif
