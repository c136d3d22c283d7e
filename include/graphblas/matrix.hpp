// graphblast_b200 frontend mirror — graphblas::Matrix<T>.
// Method set and argument checks of reference graphblas/matrix.hpp:14-252; calls
// forward to backend::Matrix<T> held by value as `matrix_` (the CPU verifiers
// read matrix_.sparse_.h_csr*, reference algorithm/bfs.hpp:101-107).
#ifndef GRAPHBLAS_MATRIX_HPP_
#define GRAPHBLAS_MATRIX_HPP_

#include <vector>

#include <graphblas/backend/cuda/matrix.hpp>

namespace graphblas {
template <typename T>
class Matrix {
 public:
  Matrix() : matrix_() {}
  Matrix(Index nrows, Index ncols) : matrix_(nrows, ncols) {}
  ~Matrix() {}

  // C API Methods
  Info nnew(Index nrows, Index ncols) {
    if (nrows == 0 || ncols == 0) return GrB_INVALID_VALUE;
    return matrix_.nnew(nrows, ncols);
  }
  Info dup(const Matrix* rhs) {
    if (rhs == NULL) return GrB_NULL_POINTER;
    return matrix_.dup(&rhs->matrix_);
  }
  Info clear() { return matrix_.clear(); }
  Info nrows(Index* nrows) const {
    if (nrows == NULL) return GrB_NULL_POINTER;
    return mutableBackend()->nrows(nrows);
  }
  Info ncols(Index* ncols) const {
    if (ncols == NULL) return GrB_NULL_POINTER;
    return mutableBackend()->ncols(ncols);
  }
  Info nvals(Index* nvals) const {
    if (nvals == NULL) return GrB_NULL_POINTER;
    return mutableBackend()->nvals(nvals);
  }
  // Host COO triples; empty triples + dat_name means "load the binary cache".
  template <typename BinaryOpT>
  Info build(const std::vector<Index>* row_indices,
      const std::vector<Index>* col_indices, const std::vector<T>* values, Index nvals,
      BinaryOpT dup, char* dat_name = NULL) {
    if (row_indices == NULL || col_indices == NULL || values == NULL)
      return GrB_NULL_POINTER;
    const bool empty = row_indices->empty() && col_indices->empty() &&
                       values->empty();
    if (empty && dat_name == NULL) return GrB_NO_VALUE;
    if (dat_name == NULL || !row_indices->empty())
      return matrix_.build(row_indices, col_indices, values, nvals, dup,
          dat_name);
    return matrix_.build(dat_name);
  }
  Info build(const std::vector<T>* values, Index nvals) {
    return matrix_.build(values, nvals);
  }
  // DEVICE CSR arrays: row_ptr (nrows+1), col_ind (nvals), values (nvals).
  Info build(Index* row_ptr, Index* col_ind, T* values, Index nvals) {
    if (row_ptr == NULL || col_ind == NULL || values == NULL)
      return GrB_NULL_POINTER;
    if (nvals == 0) return GrB_INVALID_VALUE;
    return matrix_.build(row_ptr, col_ind, values, nvals);
  }
  Info setElement(Index row_index, Index col_index) {
    return matrix_.setElement(row_index, col_index);
  }
  Info extractElement(T* val, Index row_index, Index col_index) {
    if (val == NULL) return GrB_NULL_POINTER;
    return matrix_.extractElement(val, row_index, col_index);
  }
  Info extractTuples(std::vector<Index>* row_indices, std::vector<Index>* col_indices,
      std::vector<T>* values, Index* n) {
    if (row_indices == NULL || col_indices == NULL || values == NULL ||
        n == NULL)
      return GrB_NULL_POINTER;
    return matrix_.extractTuples(row_indices, col_indices, values, n);
  }
  Info extractTuples(std::vector<T>* values, Index* n) {
    if (values == NULL || n == NULL) return GrB_NULL_POINTER;
    return matrix_.extractTuples(values, n);
  }

  // Handy methods
  void operator=(const Matrix& rhs) { matrix_.dup(&rhs.matrix_); }
  const T operator[](Index ind) { return matrix_[ind]; }
  Info print(bool force_update = false) { return matrix_.print(force_update); }
  Info check() { return matrix_.check(); }
  Info setNrows(Index nrows) { return matrix_.setNrows(nrows); }
  Info setNcols(Index ncols) { return matrix_.setNcols(ncols); }
  Info resize(Index nrows, Index ncols) { return matrix_.resize(nrows, ncols); }
  Info setStorage(Storage mat_type) { return matrix_.setStorage(mat_type); }
  Info getStorage(Storage* mat_type) const {
    if (mat_type == NULL) return GrB_NULL_POINTER;
    return matrix_.getStorage(mat_type);
  }
  template <typename U>
  Info fill(Index axis, Index nvals, U start) {
    return matrix_.fill(axis, nvals, start);
  }
  template <typename U>
  Info fillAscending(Index axis, Index nvals, U start) {
    return matrix_.fillAscending(axis, nvals, start);
  }

  backend::Matrix<T> matrix_;

 private:
  backend::Matrix<T>* mutableBackend() const {
    return const_cast<backend::Matrix<T>*>(&matrix_);
  }
};
}  // namespace graphblas

#endif  // GRAPHBLAS_MATRIX_HPP_
