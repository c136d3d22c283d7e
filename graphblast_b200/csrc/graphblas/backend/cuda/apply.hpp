// graphblast_b200 backend — apply (unary op over stored values).
//
// Replaces reference graphblas/backend/cuda/apply.hpp:14-117.  Only the
// host-side sparse-matrix variant is functional in the reference (used by
// example/gsssp.cu:79-84 to draw random edge weights in CSR order under
// GrB_BACKEND = GrB_SEQUENTIAL, outside every timed region): values are
// rewritten in CSR order with a stateful functor, the CSC is rebuilt from the
// CSR and both are uploaded.  The same is done here.
#ifndef GRAPHBLAS_BACKEND_CUDA_APPLY_HPP_
#define GRAPHBLAS_BACKEND_CUDA_APPLY_HPP_

#include <iostream>

namespace graphblas {
namespace backend {

template <typename U, typename W, typename M,
          typename BinaryOpT, typename UnaryOpT>
Info applyDense(DenseVector<W>* w, const Vector<M>* mask, BinaryOpT accum, UnaryOpT op,
    DenseVector<U>* u, Descriptor* desc) {
  std::cout << "DeVec Apply\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

template <typename U, typename W, typename M,
          typename BinaryOpT, typename UnaryOpT>
Info applySparse(SparseVector<W>* w, const Vector<M>* mask, BinaryOpT accum,
    UnaryOpT op, SparseVector<U>* u, Descriptor* desc) {
  std::cout << "SpVec Apply\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

template <typename a, typename c, typename m,
          typename BinaryOpT, typename UnaryOpT>
Info applyDense(DenseMatrix<c>* C, const Matrix<m>* mask, BinaryOpT accum, UnaryOpT op,
    DenseMatrix<a>* A, Descriptor* desc) {
  std::cout << "DeMat Apply\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

template <typename a, typename c, typename m,
          typename BinaryOpT, typename UnaryOpT>
Info applySparse(SparseMatrix<c>* C, const Matrix<m>* mask, BinaryOpT accum,
    UnaryOpT op, SparseMatrix<a>* A, Descriptor* desc) {
  Desc_value backend;
  CHECK(desc->get(GrB_BACKEND, &backend));

  if (desc->debug())
    std::cout << "Executing applySparse\n";

  if (backend == GrB_SEQUENTIAL) {
    if (mask != NULL) {
      std::cout << "Error: SpMat apply masked not implemented yet!\n";
    } else {
      CHECK(A->gpuToCpu());
      if (reinterpret_cast<void*>(C) != reinterpret_cast<void*>(A))
        CHECK(C->gpuToCpu());
      for (Index i = 0; i < A->nvals_; ++i)
        C->h_csrVal_[i] = op(A->h_csrVal_[i]);
      CHECK(C->syncCpu());
      CHECK(C->cpuToGpu());
    }
  } else {
    std::cout << "SpMat apply GPU\n";
    std::cout << "Error: Feature not implemented yet!\n";
  }
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_APPLY_HPP_
