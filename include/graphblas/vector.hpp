// graphblast_b200 frontend mirror — graphblas::Vector<T>.
// The method set of reference graphblas/vector.hpp:13-264 over a backend::Vector<T>
// held by value as `vector_`.  Every method is the same two steps — a pointer argument
// that is missing answers GrB_NULL_POINTER, otherwise the backend object does the work
// — so they are written once (`given`) and each method only names its pointers and its
// backend call.
#ifndef GRAPHBLAS_VECTOR_HPP_
#define GRAPHBLAS_VECTOR_HPP_

#include <vector>

#include <graphblas/backend/cuda/vector.hpp>

namespace graphblas {
template <typename T>
class Vector {
  typedef backend::Vector<T> Impl;

 public:
  Vector() {}
  explicit Vector(Index nsize) : vector_(nsize) {}

  // ---- size, contents -------------------------------------------------------------------
  Info nnew(Index nsize)      { return vector_.nnew(nsize); }
  Info clear()                { return vector_.clear(); }
  Info dup(const Vector* rhs) { return vector_.dup(&rhs->vector_); }
  void operator=(const Vector& rhs) { vector_.dup(&rhs.vector_); }
  Info size(Index* out) const  { return given([&](Impl& v) { return v.size(out); }, out); }
  Info nvals(Index* out) const { return given([&](Impl& v) { return v.nvals(out); }, out); }
  Info swap(Vector* other) {  // NOLINT(build/include_what_you_use)
    return given([&](Impl& v) { return v.swap(&other->vector_); }, other);
  }

  // ---- build: host tuples, host values, adopted device arrays ---------------------------
  template <typename BinaryOpT>
  Info build(const std::vector<Index>* indices, const std::vector<T>* values, Index nvals,
             BinaryOpT dup) {
    return given([&](Impl& v) { return v.build(indices, values, nvals, dup); }, indices,
                 values);
  }
  Info build(const std::vector<T>* values, Index nvals) {
    return given([&](Impl& v) { return v.build(values, nvals); }, values);
  }
  Info build(Index* d_indices, T* d_values, Index nvals) {
    return given([&](Impl& v) {
      return nvals == 0 ? GrB_INVALID_VALUE : v.build(d_indices, d_values, nvals);
    }, d_indices, d_values);
  }
  Info build(T* d_values, Index nvals) {
    return given([&](Impl& v) {
      return nvals == 0 ? GrB_INVALID_VALUE : v.build(d_values, nvals);
    }, d_values);
  }

  // ---- element and tuple access ----------------------------------------------------------
  Info setElement(T val, Index index) { return vector_.setElement(val, index); }
  Info extractElement(T* out, Index index) {
    return given([&](Impl& v) { return v.extractElement(out, index); }, out);
  }
  Info extractTuples(std::vector<Index>* indices, std::vector<T>* values, Index* n) {
    return given([&](Impl& v) { return v.extractTuples(indices, values, n); }, indices,
                 values, n);
  }
  Info extractTuples(std::vector<T>* values, Index* n) {
    return given([&](Impl& v) { return v.extractTuples(values, n); }, values, n);
  }
  const T& operator[](Index ind) { return vector_[ind]; }

  // ---- handy methods ------------------------------------------------------------------------
  Info resize(Index nvals)               { return vector_.resize(nvals); }
  Info fill(T val)                       { return vector_.fill(val); }
  Info fillAscending(Index nvals)        { return vector_.fillAscending(nvals); }
  Info print(bool force_update = false)  { return vector_.print(force_update); }
  Info countUnique(Index* out) {
    return given([&](Impl& v) { return v.countUnique(out); }, out);
  }

  // ---- storage (sparse <-> dense) ---------------------------------------------------------
  Info setStorage(Storage kind) { return vector_.setStorage(kind); }
  Info getStorage(Storage* out) const {
    return given([&](Impl& v) { return v.getStorage(out); }, out);
  }
  Info sparse2dense(T identity, Descriptor* desc = NULL) {
    return vector_.sparse2dense(identity, desc ? &desc->descriptor_ : NULL);
  }
  Info dense2sparse(T identity, Descriptor* desc) {
    return vector_.dense2sparse(identity, &desc->descriptor_);
  }

  Impl vector_;

 private:
  // `work(backend object)` once none of `needed` is NULL.  const methods of the
  // reference call non-const backend methods: the backend object is handed out mutable.
  template <typename Work, typename... Pointers>
  Info given(Work&& work, const Pointers*... needed) const {
    const bool missing = ((needed == NULL) || ...);
    if (missing) return GrB_NULL_POINTER;
    return work(const_cast<Impl&>(vector_));
  }
};
}  // namespace graphblas

#endif  // GRAPHBLAS_VECTOR_HPP_
