// graphblast_b200 backend — tril: keep entries with row >= col.
//
// Replaces reference graphblas/backend/cuda/tri.hpp:21-48.  Like the reference it
// runs on the HOST when the descriptor's GrB_BACKEND is GrB_SEQUENTIAL (the only
// mode example/gtc.cu:80-82 uses; setup, outside every timed region) and
// re-uploads; the GPU variant is "not implemented" in the reference too.
#ifndef GRAPHBLAS_BACKEND_CUDA_TRI_HPP_
#define GRAPHBLAS_BACKEND_CUDA_TRI_HPP_

#include <iostream>

namespace graphblas {
namespace backend {

template <typename a, typename c>
Info trilSparse(SparseMatrix<c>* C,
                SparseMatrix<a>* A,
                Descriptor*      desc) {
  Desc_value backend;
  CHECK(desc->get(GrB_BACKEND, &backend));

  if (desc->debug())
    std::cout << "Executing trilSparse\n";

  if (backend == GrB_SEQUENTIAL) {
    CHECK(A->gpuToCpu());
    Index kept = 0;
    Index read = 0;
    for (Index row = 0; row < A->nrows_; ++row) {
      const Index row_end = A->h_csrRowPtr_[row+1];
      A->h_csrRowPtr_[row] = kept;
      for (; read < row_end; ++read) {
        const Index col = A->h_csrColInd_[read];
        if (col <= row) {
          A->h_csrColInd_[kept] = col;
          A->h_csrVal_[kept]    = A->h_csrVal_[read];
          ++kept;
        }
      }
    }
    A->h_csrRowPtr_[A->nrows_] = kept;
    A->nvals_ = kept;

    CHECK(C->syncCpu());
    CHECK(C->cpuToGpu());
  } else {
    std::cout << "trilSparse GPU\n";
    std::cout << "Error: Feature not implemented yet!\n";
  }
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_TRI_HPP_
