// graphblast_b200 frontend mirror — graphblas::Vector<T>.
// Method set and NULL-argument behaviour of reference graphblas/vector.hpp:13-264;
// every call forwards to backend::Vector<T> held by value as `vector_`.
#ifndef GRAPHBLAS_VECTOR_HPP_
#define GRAPHBLAS_VECTOR_HPP_

#include <vector>

#include <graphblas/backend/cuda/vector.hpp>

namespace graphblas {
template <typename T>
class Vector {
 public:
  Vector() : vector_() {}
  explicit Vector(Index nsize) : vector_(nsize) {}
  ~Vector() {}

  // C API Methods
  Info nnew(Index nsize) { return vector_.nnew(nsize); }
  Info dup(const Vector* rhs) { return vector_.dup(&rhs->vector_); }
  Info clear() { return vector_.clear(); }
  Info size(Index* nsize) const {
    if (nsize == NULL) return GrB_NULL_POINTER;
    return mutableBackend()->size(nsize);
  }
  Info nvals(Index* nvals) const {
    if (nvals == NULL) return GrB_NULL_POINTER;
    return mutableBackend()->nvals(nvals);
  }
  template <typename BinaryOpT>
  Info build(const std::vector<Index>* indices, const std::vector<T>* values,
      Index nvals, BinaryOpT dup) {
    if (indices == NULL || values == NULL) return GrB_NULL_POINTER;
    return vector_.build(indices, values, nvals, dup);
  }
  Info build(const std::vector<T>* values, Index nvals) {
    if (values == NULL) return GrB_NULL_POINTER;
    return vector_.build(values, nvals);
  }
  // Device pointers, adopted
  Info build(Index* indices, T* values, Index nvals) {
    if (indices == NULL || values == NULL) return GrB_NULL_POINTER;
    if (nvals == 0) return GrB_INVALID_VALUE;
    return vector_.build(indices, values, nvals);
  }
  Info build(T* values, Index nvals) {
    if (values == NULL) return GrB_NULL_POINTER;
    if (nvals == 0) return GrB_INVALID_VALUE;
    return vector_.build(values, nvals);
  }
  Info setElement(T val, Index index) { return vector_.setElement(val, index); }
  Info extractElement(T* val, Index index) {
    if (val == NULL) return GrB_NULL_POINTER;
    return vector_.extractElement(val, index);
  }
  Info extractTuples(std::vector<Index>* indices, std::vector<T>* values, Index* n) {
    if (indices == NULL || values == NULL || n == NULL) return GrB_NULL_POINTER;
    return vector_.extractTuples(indices, values, n);
  }
  Info extractTuples(std::vector<T>* values, Index* n) {
    if (values == NULL || n == NULL) return GrB_NULL_POINTER;
    return vector_.extractTuples(values, n);
  }

  // Handy methods
  void operator=(const Vector& rhs) { vector_.dup(&rhs.vector_); }
  const T& operator[](Index ind) { return vector_[ind]; }
  Info resize(Index nvals) { return vector_.resize(nvals); }
  Info fill(T val) { return vector_.fill(val); }
  Info fillAscending(Index nvals) { return vector_.fillAscending(nvals); }
  Info print(bool force_update = false) { return vector_.print(force_update); }
  Info countUnique(Index* count) {
    if (count == NULL) return GrB_NULL_POINTER;
    return vector_.countUnique(count);
  }
  Info setStorage(Storage vec_type) { return vector_.setStorage(vec_type); }
  Info getStorage(Storage* vec_type) const {
    if (vec_type == NULL) return GrB_NULL_POINTER;
    return vector_.getStorage(vec_type);
  }
  Info sparse2dense(T identity, Descriptor* desc = NULL) {
    return vector_.sparse2dense(identity,
        desc == NULL ? NULL : &desc->descriptor_);
  }
  Info dense2sparse(T identity, Descriptor* desc) {
    return vector_.dense2sparse(identity, &desc->descriptor_);
  }
  Info swap(Vector* rhs) {  // NOLINT(build/include_what_you_use)
    if (rhs == NULL) return GrB_NULL_POINTER;
    return vector_.swap(&rhs->vector_);
  }

  backend::Vector<T> vector_;

 private:
  backend::Vector<T>* mutableBackend() const {
    return const_cast<backend::Vector<T>*>(&vector_);
  }
};
}  // namespace graphblas

#endif  // GRAPHBLAS_VECTOR_HPP_
