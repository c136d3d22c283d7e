// graphblast_b200 backend — SparseVector<T>: (index, value) lists with capacity
// nsize_ resident in HBM, host mirror materialised on demand.
//
// Replaces reference graphblas/backend/cuda/sparse_vector.hpp:22-417 (same
// methods; member names d_ind_, d_val_, nvals_, nsize_, need_update_ are the ones
// reference tests read, test/gvxm.cu:73).  Storage is allocated on first use
// and adopted device pointers are not freed here.
#ifndef GRAPHBLAS_BACKEND_CUDA_SPARSE_VECTOR_HPP_
#define GRAPHBLAS_BACKEND_CUDA_SPARSE_VECTOR_HPP_

#include <vector>
#include <iostream>
#include <unordered_set>

#include "graphblas/backend/cuda/util.hpp"

namespace graphblas {
namespace backend {

template <typename T>
class DenseVector;

template <typename T>
class SparseVector {
 public:
  SparseVector()
      : nsize_(0), nvals_(0), h_ind_(NULL), h_val_(NULL),
        d_ind_(NULL), d_val_(NULL), need_update_(0), owns_device_(true) {}

  explicit SparseVector(Index nsize)
      : nsize_(nsize), nvals_(0), h_ind_(NULL), h_val_(NULL),
        d_ind_(NULL), d_val_(NULL), need_update_(0), owns_device_(true) {}

  ~SparseVector();

  // C API Methods
  Info nnew(Index nsize);
  Info dup(const SparseVector* rhs);
  Info clear();
  inline Info size(Index* nsize_t) const;
  inline Info nvals(Index* nvals_t) const;
  template <typename BinaryOpT>
  Info build(const std::vector<Index>* indices, const std::vector<T>* values,
      Index nvals, BinaryOpT dup);
  Info build(const std::vector<T>* values, Index nvals);
  Info build(Index* indices, T* values, Index nvals);
  Info setElement(T val, Index index);
  Info extractElement(T* val, Index index);
  Info extractTuples(std::vector<Index>* indices, std::vector<T>* values, Index* n);

  // Handy methods
  const T& operator[](Index ind);
  Info resize(Index nsize);
  Info fill(Index vals);
  Info print(bool force_update = false);
  Info countUnique(Index* count);
  Info allocateCpu();
  Info allocateGpu();
  Info allocate();
  Info cpuToGpu();
  Info gpuToCpu(bool force_update = false);
  Info swap(SparseVector* rhs);

 public:  // (private in the reference; its drivers `#define private public`)
  Index  nsize_;  // capacity == logical length of the vector
  Index  nvals_;  // stored entries
  Index* h_ind_;
  T*     h_val_;
  Index* d_ind_;
  T*     d_val_;

  bool  need_update_;  // device copy newer than host copy
  bool  owns_device_;
};

template <typename T>
SparseVector<T>::~SparseVector() {
  if (h_ind_ != NULL) free(h_ind_);
  if (h_val_ != NULL) free(h_val_);
  if (owns_device_) {
    if (d_ind_ != NULL) gbFree(d_ind_);
    if (d_val_ != NULL) gbFree(d_val_);
  }
}

template <typename T>
Info SparseVector<T>::nnew(Index nsize) {
  if (nsize != nsize_) {
    if (h_ind_ != NULL) { free(h_ind_); h_ind_ = NULL; }
    if (h_val_ != NULL) { free(h_val_); h_val_ = NULL; }
    if (owns_device_) {
      if (d_ind_ != NULL) gbFree(d_ind_);
      if (d_val_ != NULL) gbFree(d_val_);
    }
    d_ind_ = NULL;
    d_val_ = NULL;
    owns_device_ = true;
  }
  nsize_ = nsize;
  nvals_ = 0;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::dup(const SparseVector* rhs) {
  if (nsize_ != rhs->nsize_) CHECK(nnew(rhs->nsize_));
  nvals_ = rhs->nvals_;
  CHECK(allocateGpu());
  if (rhs->d_ind_ != NULL && nvals_ > 0) {
    CUDA_CALL(cudaMemcpyAsync(d_ind_, rhs->d_ind_, nvals_*sizeof(Index),
        cudaMemcpyDeviceToDevice, gbStream()));
    CUDA_CALL(cudaMemcpyAsync(d_val_, rhs->d_val_, nvals_*sizeof(T),
        cudaMemcpyDeviceToDevice, gbStream()));
  }
  need_update_ = true;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::clear() {
  nvals_ = 0;
  return GrB_SUCCESS;
}

template <typename T>
inline Info SparseVector<T>::size(Index* nsize_t) const {
  *nsize_t = nsize_;
  return GrB_SUCCESS;
}

template <typename T>
inline Info SparseVector<T>::nvals(Index* nvals_t) const {
  *nvals_t = nvals_;
  return GrB_SUCCESS;
}

template <typename T>
template <typename BinaryOpT>
Info SparseVector<T>::build(const std::vector<Index>* indices,
    const std::vector<T>* values, Index nvals, BinaryOpT dup) {
  if (nvals > nsize_) {
    std::cout << "SpVec Build with indices greater than nsize_\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_PANIC;
  }
  if (nvals_ > 0) return GrB_OUTPUT_NOT_EMPTY;
  CHECK(allocate());
  nvals_ = nvals;
  for (Index i = 0; i < nvals; i++) {
    h_ind_[i] = (*indices)[i];
    h_val_[i] = (*values) [i];
  }
  CHECK(cpuToGpu());
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::build(const std::vector<T>* values, Index nvals) {
  std::cout << "Sparse Build with dense input\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

// Adopts device pointers; ownership stays with the caller.
template <typename T>
Info SparseVector<T>::build(Index* indices, T* values, Index nvals) {
  if (owns_device_) {
    if (d_ind_ != NULL) gbFree(d_ind_);
    if (d_val_ != NULL) gbFree(d_val_);
  }
  d_ind_ = indices;
  d_val_ = values;
  nvals_ = nvals;
  owns_device_ = false;
  need_update_ = true;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::setElement(T val, Index index) {
  if (nvals_ >= nsize_) return GrB_INSUFFICIENT_SPACE;
  CHECK(gpuToCpu());
  h_ind_[nvals_] = index;
  h_val_[nvals_] = val;
  nvals_++;
  CHECK(cpuToGpu());
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::extractElement(T* val, Index index) {
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::extractTuples(std::vector<Index>* indices, std::vector<T>* values,
    Index* n) {
  indices->clear();
  values->clear();
  if (*n > nvals_) {
    std::cout << *n << " > " << nvals_ << std::endl;
    std::cout << "Error: *n > nvals!\n";
    return GrB_UNINITIALIZED_OBJECT;
  } else if (*n < nvals_) {
    std::cout << *n << " < " << nvals_ << std::endl;
    std::cout << "Error: *n < nvals!\n";
    return GrB_INSUFFICIENT_SPACE;
  }
  CHECK(gpuToCpu());
  indices->assign(h_ind_, h_ind_ + *n);
  values->assign(h_val_, h_val_ + *n);
  return GrB_SUCCESS;
}

// Value stored at index `ind`, or 0 when absent (reference :226-240).
template <typename T>
const T& SparseVector<T>::operator[](Index ind) {
  static T zero = T();
  gpuToCpu();
  for (Index i = 0; i < nvals_; ++i)
    if (h_ind_[i] == ind) return h_val_[i];
  return zero;
}

template <typename T>
Info SparseVector<T>::resize(Index nsize) {
  CHECK(gpuToCpu());
  Index* h_ind_old = h_ind_;
  T*     h_val_old = h_val_;
  Index* d_ind_old = d_ind_;
  T*     d_val_old = d_val_;
  bool   old_owned = owns_device_;
  Index  to_copy   = std::min(nsize, nvals_);
  h_ind_ = NULL; h_val_ = NULL; d_ind_ = NULL; d_val_ = NULL;
  owns_device_ = true;
  nsize_ = nsize;
  CHECK(allocate());
  if (h_ind_old != NULL) memcpy(h_ind_, h_ind_old, to_copy*sizeof(Index));
  if (h_val_old != NULL) memcpy(h_val_, h_val_old, to_copy*sizeof(T));
  if (d_ind_old != NULL)
    CUDA_CALL(cudaMemcpyAsync(d_ind_, d_ind_old, to_copy*sizeof(Index), cudaMemcpyDeviceToDevice, gbStream()));
  if (d_val_old != NULL)
    CUDA_CALL(cudaMemcpyAsync(d_val_, d_val_old, to_copy*sizeof(T), cudaMemcpyDeviceToDevice, gbStream()));
  nvals_ = to_copy;
  if (h_ind_old != NULL) free(h_ind_old);
  if (h_val_old != NULL) free(h_val_old);
  if (old_owned) {
    if (d_ind_old != NULL) gbFree(d_ind_old);
    if (d_val_old != NULL) gbFree(d_val_old);
  }
  return GrB_SUCCESS;
}

// Entries 0..nvals-1 all present with value 0 (reference :280-291).
template <typename T>
Info SparseVector<T>::fill(Index nvals) {
  if (nvals > nsize_) return GrB_INDEX_OUT_OF_BOUNDS;
  CHECK(allocate());
  for (Index i = 0; i < nvals; i++) {
    h_ind_[i] = i;
    h_val_[i] = T();
  }
  nvals_ = nvals;
  CHECK(cpuToGpu());
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::print(bool force_update) {
  CHECK(gpuToCpu(force_update));
  if (nvals_ == 0) {
    std::cout << "Error: SpVec is empty!\n";
    return GrB_SUCCESS;
  }
  printArray("ind", h_ind_, std::min(nvals_, 40));
  printArray("val", h_val_, std::min(nvals_, 40));
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::countUnique(Index* count) {
  CHECK(gpuToCpu());
  std::unordered_set<Index> unique;
  for (Index i = 0; i < nvals_; i++) unique.insert(h_val_[i]);
  *count = unique.size();
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::allocateCpu() {
  if (nsize_ > 0 && h_ind_ == NULL) {
    h_ind_ = reinterpret_cast<Index*>(
        malloc(static_cast<size_t>(nsize_)*sizeof(Index)));
    h_val_ = reinterpret_cast<T*>(malloc(static_cast<size_t>(nsize_)*sizeof(T)));
    if (h_ind_ == NULL || h_val_ == NULL) {
      std::cout << "Error: CPU SpVec Out of memory!\n";
      return GrB_OUT_OF_MEMORY;
    }
    if (d_ind_ != NULL) need_update_ = true;
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::allocateGpu() {
  if (nsize_ > 0 && d_ind_ == NULL) {
    d_ind_ = reinterpret_cast<Index*>(
        gbMalloc(static_cast<size_t>(nsize_)*sizeof(Index)));
    d_val_ = reinterpret_cast<T*>(
        gbMalloc(static_cast<size_t>(nsize_)*sizeof(T)));
    owns_device_ = true;
    printMemory("SpVec");
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::allocate() {
  CHECK(allocateCpu());
  CHECK(allocateGpu());
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::cpuToGpu() {
  CHECK(allocate());
  if (nvals_ > 0) {
    CUDA_CALL(cudaMemcpyAsync(d_ind_, h_ind_, nvals_*sizeof(Index), cudaMemcpyHostToDevice, gbStream()));
    CUDA_CALL(cudaMemcpyAsync(d_val_, h_val_, nvals_*sizeof(T), cudaMemcpyHostToDevice, gbStream()));
    runtime().sync();
  }
  need_update_ = false;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::gpuToCpu(bool force_update) {
  bool fresh_host = (h_ind_ == NULL);
  CHECK(allocate());
  if ((need_update_ || force_update || fresh_host) && nvals_ > 0) {
    CUDA_CALL(cudaMemcpyAsync(h_ind_, d_ind_, nvals_*sizeof(Index), cudaMemcpyDeviceToHost, gbStream()));
    CUDA_CALL(cudaMemcpyAsync(h_val_, d_val_, nvals_*sizeof(T), cudaMemcpyDeviceToHost, gbStream()));
    runtime().sync();
  }
  need_update_ = false;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseVector<T>::swap(SparseVector* rhs) {  // NOLINT(build/include_what_you_use)
  std::swap(nsize_,       rhs->nsize_);
  std::swap(nvals_,       rhs->nvals_);
  std::swap(h_ind_,       rhs->h_ind_);
  std::swap(h_val_,       rhs->h_val_);
  std::swap(d_ind_,       rhs->d_ind_);
  std::swap(d_val_,       rhs->d_val_);
  std::swap(need_update_, rhs->need_update_);
  std::swap(owns_device_, rhs->owns_device_);
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_SPARSE_VECTOR_HPP_
