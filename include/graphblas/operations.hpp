// graphblast_b200 frontend mirror — the GraphBLAS operation templates.
//
// Signature-for-signature mirror of reference graphblas/operations.hpp:13-889:
// each entry validates its pointers and dimensions, unwraps the backend objects
// and forwards to backend::<op>.  Return codes follow the reference:
// GrB_UNINITIALIZED_OBJECT for a NULL required operand — and, for vxm/mxv, also
// for an input vector with no stored values (:71-74, :111-114) —
// GrB_DIMENSION_MISMATCH from the dimension checks, GrB_NOT_IMPLEMENTED for the
// variants the reference only declares.
#ifndef GRAPHBLAS_OPERATIONS_HPP_
#define GRAPHBLAS_OPERATIONS_HPP_

#include <vector>
#include <iostream>

#include <graphblas/backend/cuda/operations.hpp>

namespace graphblas {

namespace ops_detail {
template <typename T>
inline const backend::Vector<T>* unwrap(const Vector<T>* v) {
  return v == NULL ? NULL : &v->vector_;
}
template <typename T>
inline backend::Vector<T>* unwrap(Vector<T>* v) {
  return v == NULL ? NULL : &v->vector_;
}
template <typename T>
inline const backend::Matrix<T>* unwrap(const Matrix<T>* m) {
  return m == NULL ? NULL : &m->matrix_;
}
template <typename T>
inline backend::Matrix<T>* unwrap(Matrix<T>* m) {
  return m == NULL ? NULL : &m->matrix_;
}
inline backend::Descriptor* unwrap(Descriptor* d) {
  return d == NULL ? NULL : &d->descriptor_;
}
inline Info notImplemented(const char* what) {
  std::cout << "Error: " << what << " not implemented yet!\n";
  return GrB_NOT_IMPLEMENTED;
}
}  // namespace ops_detail

// C<mask> = accum(C, A (+.x) B)
template <typename c, typename m, typename a, typename b,
          typename BinaryOpT,     typename SemiringT>
Info mxm(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Matrix<b>* B, Descriptor* desc) {
  if (C == NULL || A == NULL || B == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowCol(B, A,    "B.nrows != A.ncols"));
  CHECK(checkDimRowRow(A, C,    "A.nrows != C.nrows"));
  CHECK(checkDimColCol(B, C,    "B.ncols != C.ncols"));
  CHECK(checkDimRowRow(C, mask, "C.nrows != mask.nrows"));
  CHECK(checkDimColCol(C, mask, "C.ncols != mask.ncols"));

  return backend::mxm<c, a, b, m>(ops_detail::unwrap(C),
      ops_detail::unwrap(mask), accum, op, ops_detail::unwrap(A),
      ops_detail::unwrap(B), ops_detail::unwrap(desc));
}

// w<mask> = accum(w, u (+.x) A)
template <typename W, typename M, typename U, typename a,
          typename BinaryOpT, typename SemiringT>
Info vxm(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Matrix<a>* A, Descriptor* desc) {
  if (w == NULL || u == NULL || A == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  Index u_nvals = 0;
  CHECK(u->nvals(&u_nvals));
  if (u_nvals == 0) return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowSize(A,  u,    "A.nrows != u.size"));
  CHECK(checkDimColSize(A,  w,    "A.ncols != w.size"));
  CHECK(checkDimSizeSize(w, mask, "w.size  != mask.size"));

  return backend::vxm<W, U, a, M>(ops_detail::unwrap(w),
      ops_detail::unwrap(mask), accum, op, ops_detail::unwrap(u),
      ops_detail::unwrap(A), ops_detail::unwrap(desc));
}

// w<mask> = accum(w, A (+.x) u)
template <typename W, typename M, typename a, typename U,
          typename BinaryOpT, typename SemiringT>
Info mxv(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Vector<U>* u, Descriptor* desc) {
  if (w == NULL || u == NULL || A == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  Index u_nvals = 0;
  CHECK(u->nvals(&u_nvals));
  if (u_nvals == 0) return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimColSize(A,  u,    "A.ncols != u.size"));
  CHECK(checkDimRowSize(A,  w,    "A.nrows != w.size"));
  CHECK(checkDimSizeSize(w, mask, "w.size  != mask.size"));

  return backend::mxv<W, U, a, M>(ops_detail::unwrap(w),
      ops_detail::unwrap(mask), accum, op, ops_detail::unwrap(A),
      ops_detail::unwrap(u), ops_detail::unwrap(desc));
}

// w<mask> = accum(w, u .* v)
template <typename W, typename M, typename U, typename V,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Vector<V>* v, Descriptor* desc) {
  if (w == NULL || u == NULL || v == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimSizeSize(u, v,    "u.size != v.size"));
  CHECK(checkDimSizeSize(u, w,    "u.size != mask.size"));
  CHECK(checkDimSizeSize(u, mask, "v.size != mask.size"));

  return backend::eWiseMult(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(u), ops_detail::unwrap(v), ops_detail::unwrap(desc));
}

// C<mask> = accum(C, A .* B)
template <typename c, typename m, typename a, typename b,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Matrix<b>* B, Descriptor* desc) {
  if (C == NULL || A == NULL || B == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowRow(B, A,    "B.nrows != A.nrows"));
  CHECK(checkDimColCol(B, A,    "B.ncols != A.ncols"));
  CHECK(checkDimRowRow(A, C,    "A.nrows != C.nrows"));
  CHECK(checkDimColCol(A, C,    "A.ncols != C.ncols"));
  CHECK(checkDimRowRow(C, mask, "C.nrows != mask.nrows"));
  CHECK(checkDimColCol(C, mask, "C.ncols != mask.ncols"));

  return backend::eWiseMult(ops_detail::unwrap(C), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(A), ops_detail::unwrap(B), ops_detail::unwrap(desc));
}

// Extension: C = A .* val (scalar broadcast)
template <typename c, typename m, typename a, typename b,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, b val, Descriptor* desc) {
  if (C == NULL || A == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowRow(A, C,    "A.nrows != C.nrows"));
  CHECK(checkDimColCol(A, C,    "A.ncols != C.ncols"));
  CHECK(checkDimRowRow(A, mask, "A.nrows != mask.nrows"));
  CHECK(checkDimColCol(A, mask, "A.ncols != mask.ncols"));

  return backend::eWiseMult(ops_detail::unwrap(C), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(A), val, ops_detail::unwrap(desc));
}

// Extension: C = A .* B with a vector B broadcast along rows (or along columns
// when GrB_INP1 is GrB_TRAN)
template <typename c, typename m, typename a, typename b,
          typename BinaryOpT,     typename SemiringT>
Info eWiseMult(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Vector<b>* B, Descriptor* desc) {
  if (C == NULL || A == NULL || B == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowRow(A, C,    "A.nrows != C.nrows"));
  CHECK(checkDimColCol(A, C,    "A.ncols != C.ncols"));
  CHECK(checkDimRowRow(A, mask, "A.nrows != mask.nrows"));
  CHECK(checkDimColCol(A, mask, "A.ncols != mask.ncols"));

  return backend::eWiseMult(ops_detail::unwrap(C), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(A), ops_detail::unwrap(B), ops_detail::unwrap(desc));
}

// w<mask> = accum(w, u + v)
template <typename W, typename M, typename U, typename V,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAdd(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Vector<V>* v, Descriptor* desc) {
  if (w == NULL || u == NULL || v == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimSizeSize(u, v,    "u.size != v.size"));
  CHECK(checkDimSizeSize(u, mask, "u.size != mask.size"));
  CHECK(checkDimSizeSize(v, mask, "v.size != mask.size"));
  CHECK(checkDimSizeSize(w, mask, "w.size != mask.size"));

  return backend::eWiseAdd(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(u), ops_detail::unwrap(v), ops_detail::unwrap(desc));
}

template <typename c, typename m, typename a, typename b,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAdd(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, SemiringT op,
    const Matrix<a>* A, const Matrix<b>* B, Descriptor* desc) {
  return ops_detail::notImplemented("eWiseAdd matrix variant");
}

// Extension: w = u + val (scalar broadcast)
template <typename W, typename M, typename U, typename V,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAdd(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, V val, Descriptor* desc) {
  if (w == NULL || u == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimSizeSize(u, w,    "u.size != mask.size"));
  CHECK(checkDimSizeSize(u, mask, "v.size != mask.size"));

  return backend::eWiseAdd(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(u), val, ops_detail::unwrap(desc));
}

template <typename W, typename M, typename U,
          typename BinaryOpT>
Info extract(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, const Vector<U>* u,
    const std::vector<Index>* indices, Index nindices, Descriptor* desc) {
  return ops_detail::notImplemented("extract vector variant");
}

template <typename c, typename m, typename a,
          typename BinaryOpT>
Info extract(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, const Matrix<a>* A,
    const std::vector<Index>* row_indices, Index nrows,
    const std::vector<Index>* col_indices, Index ncols, Descriptor* desc) {
  return ops_detail::notImplemented("extract matrix variant");
}

template <typename W, typename M, typename a,
          typename BinaryOpT>
Info extract(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, const Matrix<a>* A,
    const std::vector<Index>* row_indices, Index nrows, Index col_index,
    Descriptor* desc) {
  return ops_detail::notImplemented("extract matrix variant");
}

template <typename W, typename M, typename U,
          typename BinaryOpT>
Info assignIndexed(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    const Vector<U>* u, int* indices, Index nindices, Descriptor* desc) {
  if (w == NULL || u == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  CHECK(checkDimSizeSize(w, mask, "w.size  != mask.size"));
  return backend::assignIndexed(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum,
      ops_detail::unwrap(u), indices, nindices, ops_detail::unwrap(desc));
}

template <typename c, typename m, typename a,
          typename BinaryOpT>
Info assign(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, const Matrix<a>* A,
    const std::vector<Index>* row_indices, Index nrows,
    const std::vector<Index>* col_indices, Index ncols, Descriptor* desc) {
  return ops_detail::notImplemented("assign matrix variant");
}

template <typename c, typename M, typename U,
          typename BinaryOpT>
Info assign(Matrix<c>* C, const Vector<M>* mask, BinaryOpT accum, const Vector<U>* u,
    const std::vector<Index>* row_indices, Index nrows, Index col_index,
    Descriptor* desc) {
  return ops_detail::notImplemented("assign matrix variant");
}

template <typename c, typename M, typename U,
          typename BinaryOpT>
Info assign(Matrix<c>* C, const Vector<M>* mask, BinaryOpT accum, const Vector<U>* u,
    Index row_index, const std::vector<Index>* col_indices, Index ncols,
    Descriptor* desc) {
  return ops_detail::notImplemented("assign matrix variant");
}

// w<mask>[indices] = val (constant assign)
template <typename W, typename M, typename T, typename I,
          typename BinaryOpT>
Info assign(Vector<W>* w, Vector<M>* mask, BinaryOpT accum, T val,
    const Vector<I>* indices, Index nindices, Descriptor* desc) {
  if (w == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  CHECK(checkDimSizeSize(w, mask, "w.size  != mask.size"));
  return backend::assign(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum, val,
      ops_detail::unwrap(indices), nindices, ops_detail::unwrap(desc));
}

template <typename c, typename m, typename T,
          typename BinaryOpT>
Info assign(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, T val,
    const std::vector<Index>* row_indices, Index nrows,
    const std::vector<Index>* col_indices, Index ncols, Descriptor* desc) {
  return ops_detail::notImplemented("assign matrix variant");
}

// w<mask> = accum(w, op(u))
template <typename W, typename M, typename U,
          typename BinaryOpT,     typename UnaryOpT>
Info apply(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, UnaryOpT op,
    const Vector<U>* u, Descriptor* desc) {
  if (w == NULL || u == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimSizeSize(u, w,    "u.size != w.size"));
  CHECK(checkDimSizeSize(u, mask, "u.size != mask.size"));
  CHECK(checkDimSizeSize(w, mask, "w.size != mask.size"));

  return backend::apply(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(u), ops_detail::unwrap(desc));
}

// C<mask> = accum(C, op(A))
template <typename c, typename m, typename a,
          typename BinaryOpT,     typename UnaryOpT>
Info apply(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, UnaryOpT op,
    const Matrix<a>* A, Descriptor* desc) {
  if (C == NULL || A == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowRow(A, C,    "A.nrows != C.nrows"));
  CHECK(checkDimColCol(A, C,    "A.ncols != C.ncols"));
  CHECK(checkDimRowRow(A, mask, "A.nrows != mask.nrows"));
  CHECK(checkDimColCol(A, mask, "A.ncols != mask.ncols"));

  return backend::apply(ops_detail::unwrap(C), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(A), ops_detail::unwrap(desc));
}

// w<mask> = accum(w, reduce rows of A)
template <typename W, typename M, typename a,
          typename BinaryOpT,     typename MonoidT>
Info reduce(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, MonoidT op,
    const Matrix<a>* A, Descriptor* desc) {
  if (w == NULL || A == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  return backend::reduce(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum, op,
      ops_detail::unwrap(A), ops_detail::unwrap(desc));
}

// val = accum(val, reduce(u))
template <typename T, typename U,
          typename BinaryOpT, typename MonoidT>
Info reduce(T* val, BinaryOpT accum, MonoidT op, const Vector<U>* u, Descriptor* desc) {
  if (val == NULL || u == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  return backend::reduce(val, accum, op, ops_detail::unwrap(u),
      ops_detail::unwrap(desc));
}

// val = accum(val, reduce(A))
template <typename T, typename a,
          typename BinaryOpT, typename MonoidT>
Info reduce(T* val, BinaryOpT accum, MonoidT op, const Matrix<a>* A, Descriptor* desc) {
  if (val == NULL || A == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  return backend::reduce(val, accum, op, ops_detail::unwrap(A),
      ops_detail::unwrap(desc));
}

template <typename c, typename m, typename a,
          typename BinaryOpT>
Info transpose(Matrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, const Matrix<a>* A,
    Descriptor* desc) {
  return ops_detail::notImplemented("transpose");
}

// Extension: val = trace(A * B^T)
template <typename T, typename a, typename b,
          typename SemiringT>
Info traceMxmTranspose(T* val, SemiringT op, const Matrix<a>* A, const Matrix<b>* B,
    Descriptor* desc) {
  if (val == NULL || A == NULL || B == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  return backend::traceMxmTranspose(val, op, ops_detail::unwrap(A),
      ops_detail::unwrap(B), ops_detail::unwrap(desc));
}

template <typename b, typename a, typename T,
          typename MonoidT>
Info scale(Matrix<b>* B, MonoidT op, const Matrix<a>* A, T val, Descriptor* desc) {
  return ops_detail::notImplemented("scale matrix variant");
}

template <typename W, typename U, typename T,
          typename MonoidT>
Info scale(Vector<W>* w, MonoidT op, const Vector<U>* u, T val, Descriptor* desc) {
  return ops_detail::notImplemented("scale vector variant");
}

// Extension: w[indices[i]] = mask .* val
template <typename W, typename M, typename I, typename T>
Info scatter(Vector<W>* w, const Vector<M>* mask, const Vector<I>* indices, T val,
    Descriptor* desc) {
  if (indices == NULL || w == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  return backend::scatter(ops_detail::unwrap(w), ops_detail::unwrap(mask),
      ops_detail::unwrap(indices), val, ops_detail::unwrap(desc));
}

// Extension: w[indices[i]] = u[i]
template <typename W, typename M, typename U, typename I,
          typename BinaryOpT>
Info assignScatter(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    const Vector<U>* u, const Vector<I>* indices, Descriptor* desc) {
  if (w == NULL || u == NULL || indices == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  CHECK(checkDimSizeSize(w, mask, "w.size  != mask.size"));
  return backend::assignScatter(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum,
      ops_detail::unwrap(u), ops_detail::unwrap(indices), ops_detail::unwrap(desc));
}

// Extension: w[i] = u[indices[i]]
template <typename W, typename M, typename U, typename I,
          typename BinaryOpT>
Info extractGather(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    const Vector<U>* u, const Vector<I>* indices, Descriptor* desc) {
  if (u == NULL || w == NULL || indices == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  return backend::extractGather(ops_detail::unwrap(w), ops_detail::unwrap(mask), accum,
      ops_detail::unwrap(u), ops_detail::unwrap(indices), ops_detail::unwrap(desc));
}

template <typename W, typename a>
Info graphColor(Vector<W>* w, const Matrix<a>* A, Descriptor* desc) {
  if (w == NULL || A == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;
  return backend::graphColor(ops_detail::unwrap(w), ops_detail::unwrap(A),
      ops_detail::unwrap(desc));
}

// Extension: vxm fused with an apply on the input
template <typename W, typename M, typename U, typename a,
          typename BinaryOpT, typename SemiringT>
Info applyVxm(Vector<W>* w, const Vector<M>* mask, BinaryOpT accum, SemiringT op,
    const Vector<U>* u, const Matrix<a>* A, Descriptor* desc) {
  if (w == NULL || u == NULL || A == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  Index u_nvals = 0;
  CHECK(u->nvals(&u_nvals));
  if (u_nvals == 0) return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowSize(A,  u,    "A.nrows != u.size"));
  CHECK(checkDimColSize(A,  w,    "A.ncols != w.size"));
  CHECK(checkDimSizeSize(w, mask, "w.size  != mask.size"));

  return backend::applyVxm<W, U, a, M>(ops_detail::unwrap(w),
      ops_detail::unwrap(mask), accum, op, ops_detail::unwrap(u),
      ops_detail::unwrap(A), ops_detail::unwrap(desc));
}

// Extension: C = lower triangle of A (row >= col)
template <typename c, typename a>
Info tril(Matrix<c>* C, Matrix<a>* A, Descriptor* desc) {
  if (C == NULL || A == NULL || desc == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  CHECK(checkDimRowRow(A, C, "A.nrows != C.nrows"));
  CHECK(checkDimColCol(A, C, "A.ncols != C.ncols"));

  return backend::tril(ops_detail::unwrap(C), ops_detail::unwrap(A),
      ops_detail::unwrap(desc));
}

}  // namespace graphblas

#endif  // GRAPHBLAS_OPERATIONS_HPP_
