// graphblast_b200 backend — eWiseAdd hosts (vector (+) vector, vector (+) scalar).
//
// Replaces reference graphblas/backend/cuda/ewiseadd.hpp:18-280; every variant
// applies the semiring's ADD like the reference (the accum argument is only
// inspected for presence).  Quirks kept:
//  * sparse (+) dense first rewrites EVERY element of the dense operand with
//    op(x, identity) (or op(identity, x) when the dense operand came first,
//    `reverse`), reference :147-148, then applies op(sparse, dense) at the sparse
//    positions — always in that argument order (kernels/ewiseadd.hpp:29-45);
//    the reference's dup + constant pass is fused into one kernel here;
//  * dense (+) scalar computes add(u[i], val) (reference :268-276).
#ifndef GRAPHBLAS_BACKEND_CUDA_EWISEADD_HPP_
#define GRAPHBLAS_BACKEND_CUDA_EWISEADD_HPP_

#include <iostream>
#include <string>

#include "graphblas/backend/cuda/kernels/kernels.hpp"

namespace graphblas {
namespace backend {

// sparse (+) sparse -> dense: not implemented in the reference either (:18-27).
template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAddInner(DenseVector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    SemiringT op, const SparseVector<U>* u, const SparseVector<V>* v,
    Descriptor* desc) {
  std::cout << "Error: eWiseAdd sparse-sparse not implemented yet!\n";
  return GrB_SUCCESS;
}

// dense (+) dense
template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAddInner(DenseVector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    SemiringT op, const DenseVector<U>* u, const DenseVector<V>* v, Descriptor* desc) {
  if (mask != NULL) {
    std::cout << "Error: Masked eWiseAdd dense-dense not implemented yet!\n";
    return GrB_SUCCESS;
  }
  Index n;
  u->nvals(&n);
  CHECK(w->allocateGpu());
  if (n > 0) {
    ewiseBinaryDenseKernel<<<gridFor(n, 256), 256, 0, gbStream()>>>(w->d_val_,
        extractAdd(op), u->d_val_, v->d_val_, n);
    GB_KERNEL_CHECK();
  }
  w->touched();
  return GrB_SUCCESS;
}

// sparse (+) dense; reverse == true when the dense operand was the first argument
template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAddInner(DenseVector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    SemiringT op, const SparseVector<U>* u, const DenseVector<V>* v, bool reverse,
    Descriptor* desc) {
  if (mask != NULL) {
    std::cout << "Error: Masked eWiseAdd sparse-dense not implemented yet!\n";
    return GrB_SUCCESS;
  }
  Index u_nvals, v_nvals;
  u->nvals(&u_nvals);
  v->nvals(&v_nvals);
  CHECK(w->allocateGpu());
  cudaStream_t s = gbStream();

  // When w is a different vector the sparse pass must read the ORIGINAL dense
  // operand (the reference reads v->d_val_, not w), which still holds it.  When
  // w aliases v the constant pass rewrites it in place first, as the reference
  // does.
  if (v_nvals > 0) {
    ewiseConstantFromKernel<<<gridFor(v_nvals, 256), 256, 0, s>>>(w->d_val_,
        v->d_val_, extractAdd(op), op.identity(), reverse, v_nvals);
    GB_KERNEL_CHECK();
  }
  if (u_nvals > 0) {
    ewiseSparseDenseKernel<<<gridFor(u_nvals, 256), 256, 0, s>>>(w->d_val_,
        extractAdd(op), u->d_ind_, u->d_val_, v->d_val_, u_nvals);
    GB_KERNEL_CHECK();
  }
  w->touched();
  return GrB_SUCCESS;
}

// sparse (+) scalar
template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAddInner(DenseVector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    SemiringT op, const SparseVector<U>* u, V val, Descriptor* desc) {
  if (mask != NULL) {
    std::cout << "eWiseAdd Sparse Vector Broadcast Scalar with Mask\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_SUCCESS;
  }
  Index u_nvals;
  u->nvals(&u_nvals);
  auto add_op = extractAdd(op);
  CHECK(w->fill(add_op(op.identity(), val)));
  if (u_nvals > 0) {
    ewiseSparseDenseKernel<<<gridFor(u_nvals, 256), 256, 0, gbStream()>>>(
        w->d_val_, extractAdd(op), u->d_ind_, u->d_val_, w->d_val_, u_nvals);
    GB_KERNEL_CHECK();
  }
  w->touched();
  return GrB_SUCCESS;
}

// dense (+) scalar
template <typename W, typename U, typename V, typename M,
          typename BinaryOpT,     typename SemiringT>
Info eWiseAddInner(DenseVector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    SemiringT op, const DenseVector<U>* u, V val, Descriptor* desc) {
  if (mask != NULL) {
    std::cout << "eWiseAdd Dense Vector Broadcast Scalar with Mask\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_SUCCESS;
  }
  Index n;
  u->size(&n);
  CHECK(w->allocateGpu());
  if (n > 0) {
    ewiseScalarKernel<<<gridFor(n, 256), 256, 0, gbStream()>>>(w->d_val_,
        extractAdd(op), u->d_val_, n, val);
    GB_KERNEL_CHECK();
  }
  w->touched();
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_EWISEADD_HPP_
