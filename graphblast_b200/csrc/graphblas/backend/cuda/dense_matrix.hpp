// graphblast_b200 backend — DenseMatrix<T> placeholder.
//
// The reference's dense-matrix paths (gemm/gemv/spmm) are stubs that print
// "not implemented" (reference gemv.hpp:16-43, operations.hpp:55-57) and no
// algorithm on the hot path reaches them (out of scope).
// The class exists so backend::Matrix<T> keeps its two-storage shape.
#ifndef GRAPHBLAS_BACKEND_CUDA_DENSE_MATRIX_HPP_
#define GRAPHBLAS_BACKEND_CUDA_DENSE_MATRIX_HPP_

#include <vector>
#include <iostream>

namespace graphblas {
namespace backend {

template <typename T>
class DenseMatrix {
 public:
  DenseMatrix() : nrows_(0), ncols_(0), nvals_(0) {}
  DenseMatrix(Index nrows, Index ncols)
      : nrows_(nrows), ncols_(ncols), nvals_(0) {}
  ~DenseMatrix() {}

  Info nnew(Index nrows, Index ncols) {
    nrows_ = nrows;
    ncols_ = ncols;
    return GrB_SUCCESS;
  }
  Info dup(const DenseMatrix* rhs) { return GrB_NOT_IMPLEMENTED; }
  Info clear() { nvals_ = 0; return GrB_SUCCESS; }
  Info nrows(Index* n) const { *n = nrows_; return GrB_SUCCESS; }
  Info ncols(Index* n) const { *n = ncols_; return GrB_SUCCESS; }
  Info nvals(Index* n) const { *n = nvals_; return GrB_SUCCESS; }
  Info build(const std::vector<T>* values, Index nvals) {
    std::cout << "DeMat Build\nError: Feature not implemented yet!\n";
    return GrB_NOT_IMPLEMENTED;
  }
  Info print(bool force_update = false) { return GrB_SUCCESS; }
  Info setNrows(Index nrows) { nrows_ = nrows; return GrB_SUCCESS; }
  Info setNcols(Index ncols) { ncols_ = ncols; return GrB_SUCCESS; }
  Info resize(Index nrows, Index ncols) {
    nrows_ = nrows;
    ncols_ = ncols;
    return GrB_SUCCESS;
  }

  Index nrows_;
  Index ncols_;
  Index nvals_;
};

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_DENSE_MATRIX_HPP_
