#ifndef DUNE_EIGENSOLVER_B200_MULTIVECTOR_HH
#define DUNE_EIGENSOLVER_B200_MULTIVECTOR_HH

/** \file
 *  Host-side tall-skinny container with the reference's interface and storage contract
 *  (reference dune/eigensolver/multivector.hh:17-146): an n x m block kept as m/b panels of b columns,
 *  each panel an n x b row-major slab, i.e. element (i,j) lives at ((j/b)*n + i)*b + j%b. 64-byte aligned,
 *  zero-initialised, deep copy, movable; m % b != 0 throws std::invalid_argument with the reference's text.
 *
 *  This is the layout at the BOUNDARY: de_mv_upload_panel8 / de_mv_download_panel8 of the C ABI consume and
 *  produce exactly this buffer for b = 8. On the device the block is row-major n x m (one contiguous 8*m-byte
 *  row per matrix row), which is what the sm_100a kernels want; the conversion is a device kernel.
 *  Unlike the reference, constructors do not print the buffer address.
 */

#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <new>
#include <stdexcept>
#include <utility>

template <typename T, std::size_t b = 8>
class MultiVector
{
public:
  static const std::size_t blocksize = b;
  using value_type = T;

  MultiVector() = default;

  MultiVector(std::size_t rows, std::size_t cols) : n_(rows), m_(cols)
  {
    if (cols % b != 0)
      throw std::invalid_argument("number of cols must be a multiple of block size");
    data_ = allocate(n_ * m_);
  }

  MultiVector(const MultiVector &o) : n_(o.n_), m_(o.m_), data_(allocate(o.n_ * o.m_))
  {
    std::copy(o.data_, o.data_ + n_ * m_, data_);
  }

  MultiVector(MultiVector &&o) noexcept : n_(o.n_), m_(o.m_), data_(o.data_)
  {
    o.n_ = o.m_ = 0;
    o.data_ = nullptr;
  }

  ~MultiVector() { release(data_); }

  MultiVector &operator=(const MultiVector &o)
  {
    if (this != &o)
    {
      MultiVector tmp(o);
      swap(tmp);
    }
    return *this;
  }

  MultiVector &operator=(MultiVector &&o) noexcept
  {
    if (this != &o)
    {
      release(data_);
      n_ = o.n_;
      m_ = o.m_;
      data_ = o.data_;
      o.n_ = o.m_ = 0;
      o.data_ = nullptr;
    }
    return *this;
  }

  void swap(MultiVector &o) noexcept
  {
    std::swap(n_, o.n_);
    std::swap(m_, o.m_);
    std::swap(data_, o.data_);
  }

  T &operator()(std::size_t i, std::size_t j) { return data_[offset(i, j)]; }
  const T &operator()(std::size_t i, std::size_t j) const { return data_[offset(i, j)]; }

  std::size_t rows() const { return n_; }
  std::size_t cols() const { return m_; }

  //! raw panel storage (what the C ABI's *_panel8 entry points take for b = 8)
  T *data() { return data_; }
  const T *data() const { return data_; }

private:
  std::size_t n_ = 0, m_ = 0;
  T *data_ = nullptr;

  std::size_t offset(std::size_t i, std::size_t j) const { return ((j / b) * n_ + i) * b + (j % b); }

  static T *allocate(std::size_t count)
  {
    if (count == 0)
      return nullptr;
    const std::size_t bytes = ((count * sizeof(T) + 63) / 64) * 64;
    void *p = std::aligned_alloc(64, bytes);
    if (!p)
      throw std::bad_alloc();
    T *t = static_cast<T *>(p);
    std::fill(t, t + count, T(0));
    return t;
  }
  static void release(T *p) { std::free(p); }
};

namespace std
{
  template <typename T, std::size_t b>
  inline void swap(MultiVector<T, b> &x, MultiVector<T, b> &y) noexcept
  {
    x.swap(y);
  }
} // namespace std

#endif
