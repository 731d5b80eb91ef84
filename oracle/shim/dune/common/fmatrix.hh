// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product.
//
// Minimal stand-in for dune-common's FieldMatrix, exposing exactly the surface the
// reference's hot path touches (SURVEY.md §8b "Matrix access"):
//   block_type::rows / ::cols                      (reference eigensolver.hh:36-37)
//   implicit conversion of a 1x1 block to double   (reference kernels_cpp.hh:391, :652)
//   (*col)[i][j] element access                    (reference eigensolver.hh:65, src/dune-eigensolver.cc:117)
// Written from scratch for this repo; dune-common itself is not available in the image.
#ifndef DE_ORACLE_SHIM_FMATRIX_HH
#define DE_ORACLE_SHIM_FMATRIX_HH

#include <cstddef>

namespace Dune
{
  template <class K, int R, int C>
  class FieldMatrix
  {
    K a_[R][C];

  public:
    static constexpr int rows = R;
    static constexpr int cols = C;
    using field_type = K;

    FieldMatrix()
    {
      for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j)
          a_[i][j] = K(0);
    }
    FieldMatrix(const K &v)
    {
      for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j)
          a_[i][j] = (i == j) ? v : K(0);
    }

    K *operator[](std::size_t i) { return a_[i]; }
    const K *operator[](std::size_t i) const { return a_[i]; }

    // only meaningful for 1x1 blocks, which is all the reference kernels accept
    operator K &() { return a_[0][0]; }
    operator const K &() const { return a_[0][0]; }

    FieldMatrix &operator=(const K &v)
    {
      for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j)
          a_[i][j] = (i == j) ? v : K(0);
      return *this;
    }

    FieldMatrix &axpy(const K &s, const FieldMatrix &o)
    {
      for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j)
          a_[i][j] += s * o.a_[i][j];
      return *this;
    }
  };
} // namespace Dune

#endif
