// graphblast_b200 frontend mirror — dimension checks used by operations.hpp.
// Same helper names and return codes as reference graphblas/dimension.hpp:13-115
// (NULL operands pass; a mismatch prints the message and returns
// GrB_DIMENSION_MISMATCH), generated from one comparison helper.
#ifndef GRAPHBLAS_DIMENSION_HPP_
#define GRAPHBLAS_DIMENSION_HPP_

#include <string>
#include <iostream>

namespace graphblas {
template <typename T> class Vector;
template <typename T> class Matrix;

namespace dimension_detail {
inline Info compare(Index lhs, Index rhs, const std::string& str) {
  if (lhs == rhs) return GrB_SUCCESS;
  std::cout << str << std::endl;
  return GrB_DIMENSION_MISMATCH;
}
}  // namespace dimension_detail

template <typename U>
inline Info checkDimVecNvals(const Vector<U>* u, const std::string& str) {
  if (u == NULL) return GrB_INVALID_OBJECT;
  Index u_nvals;
  CHECK(u->nvals(&u_nvals));
  if (u_nvals == 0) {
    std::cout << str << std::endl;
    return GrB_INVALID_OBJECT;
  }
  return GrB_SUCCESS;
}

#define GB_DIM_CHECK(NAME, TA, TB, GET_A, GET_B)                              \
template <typename a, typename b>                                             \
inline Info NAME(const TA<a>* A, const TB<b>* B, const std::string& str) {    \
  if (A == NULL || B == NULL) return GrB_SUCCESS;                             \
  Index lhs, rhs;                                                             \
  CHECK(A->GET_A(&lhs));                                                      \
  CHECK(B->GET_B(&rhs));                                                      \
  return dimension_detail::compare(lhs, rhs, str);                            \
}

GB_DIM_CHECK(checkDimRowCol,   Matrix, Matrix, nrows, ncols)
GB_DIM_CHECK(checkDimRowRow,   Matrix, Matrix, nrows, nrows)
GB_DIM_CHECK(checkDimColCol,   Matrix, Matrix, ncols, ncols)
GB_DIM_CHECK(checkDimRowSize,  Matrix, Vector, nrows, size)
GB_DIM_CHECK(checkDimColSize,  Matrix, Vector, ncols, size)
GB_DIM_CHECK(checkDimSizeSize, Vector, Vector, size,  size)

#undef GB_DIM_CHECK
}  // namespace graphblas

#endif  // GRAPHBLAS_DIMENSION_HPP_
