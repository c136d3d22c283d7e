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

// In-place on A's host CSR (C is the same object in the only caller, reference
// example/gtc.cu:80-82): entries above the diagonal are squeezed out row by row,
// then the CSC is rebuilt from the CSR and both are uploaded again.
template <typename a, typename c>
Info trilSparse(SparseMatrix<c>* C, SparseMatrix<a>* A, Descriptor* desc) {
  Desc_value where;
  CHECK(desc->get(GrB_BACKEND, &where));
  if (desc->debug()) std::cout << "Executing trilSparse\n";

  if (where != GrB_SEQUENTIAL) {
    std::cout << "trilSparse GPU\n";
    std::cout << "Error: Feature not implemented yet!\n";
    return GrB_SUCCESS;
  }

  CHECK(A->gpuToCpu());
  Index* const rowptr = A->h_csrRowPtr_;
  Index* const colind = A->h_csrColInd_;
  a* const     values = A->h_csrVal_;
  Index out = 0;                       // next free slot of the squeezed arrays
  Index in  = 0;                       // next entry to look at
  for (Index r = 0; r < A->nrows_; ++r) {
    const Index stop = rowptr[r + 1];  // read before rowptr[r + 1] is overwritten
    rowptr[r] = out;
    while (in < stop) {
      if (colind[in] <= r) {
        colind[out] = colind[in];
        values[out] = values[in];
        ++out;
      }
      ++in;
    }
  }
  rowptr[A->nrows_] = out;
  A->nvals_ = out;

  // A triangle of a symmetric pattern is not symmetric: the column-major side can
  // no longer borrow the CSR index arrays (the masked mxm walks mask columns).
  C->dropSymmetry();

  CHECK(C->syncCpu());
  CHECK(C->cpuToGpu());
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_TRI_HPP_
