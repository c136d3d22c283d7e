// graphblast_b200 backend — masked SpGEMM host (triangle counting path).
//
// Replaces reference graphblas/backend/cuda/spgemm.hpp:22-110 (spgemmMasked).
// C takes the mask's pattern (C->dup(mask)) and one value per mask entry is
// computed as the dot product A(i,:) . B(:,j).  The reference's unmasked
// cusparse_spgemm/cusparse_spgemm2 (:114-512) call cuSPARSE csrgemm2 entry
// points that no longer exist in CUDA 12 and are out of scope.
#ifndef GRAPHBLAS_BACKEND_CUDA_SPGEMM_HPP_
#define GRAPHBLAS_BACKEND_CUDA_SPGEMM_HPP_

#include <iostream>

#include "graphblas/backend/cuda/kernels/kernels.hpp"

namespace graphblas {
namespace backend {

template <typename c, typename a, typename b, typename m,
          typename BinaryOpT,     typename SemiringT>
Info spgemmMasked(SparseMatrix<c>* C, const Matrix<m>* mask, BinaryOpT accum,
    SemiringT op, const SparseMatrix<a>* A, const SparseMatrix<b>* B,
    Descriptor* desc) {
  Desc_value scmp_mode, inp0_mode, inp1_mode;
  CHECK(desc->get(GrB_MASK, &scmp_mode));
  CHECK(desc->get(GrB_INP0, &inp0_mode));
  CHECK(desc->get(GrB_INP1, &inp1_mode));

  const bool use_mask   = (mask != NULL);
  const bool use_tran_A = inp0_mode == GrB_TRAN;
  const bool use_tran_B = inp1_mode == GrB_TRAN;

  const Index* A_csrRowPtr = (use_tran_A) ? A->d_cscColPtr_ : A->d_csrRowPtr_;
  const Index* A_csrColInd = (use_tran_A) ? A->d_cscRowInd_ : A->d_csrColInd_;
  const a*     A_csrVal    = (use_tran_A) ? A->d_cscVal_    : A->d_csrVal_;
  const Index  A_nrows     = (use_tran_A) ? A->ncols_       : A->nrows_;

  const Index* B_cscColPtr = (use_tran_B) ? B->d_csrRowPtr_ : B->d_cscColPtr_;
  const Index* B_cscRowInd = (use_tran_B) ? B->d_csrColInd_ : B->d_cscRowInd_;
  const b*     B_cscVal    = (use_tran_B) ? B->d_csrVal_    : B->d_cscVal_;

  if (A_csrRowPtr == NULL || B_cscColPtr == NULL)
    return GrB_UNINITIALIZED_OBJECT;

  if (use_mask) {
    Storage mask_mat_type;
    CHECK(mask->getStorage(&mask_mat_type));
    if (mask_mat_type == GrB_DENSE) {
      std::cout << "SpGEMM with dense mask\n";
      std::cout << "Error: Feature not implemented yet!\n";
    } else {
      if (reinterpret_cast<const void*>(C) != reinterpret_cast<const void*>(A) &&
          reinterpret_cast<const void*>(C) != reinterpret_cast<const void*>(B))
        CHECK(C->dup(&mask->sparse_));

      const SparseMatrix<m>* sparse_mask = &mask->sparse_;
      unsigned long long* work = desc->counters() + 3;
      cudaStream_t s = gbStream();
      CUDA_CALL(cudaMemsetAsync(work, 0, sizeof(unsigned long long), s));
      const int grid = runtime().sm_count*8;
      unsigned long long* prof_cell = NULL;
      if (profiler().enabled) {
        profiler().ensureCells();
        prof_cell = profiler().d_cells + GB_PROF_SPGEMM;
      }
      profiler().begin(GB_PROF_SPGEMM, s);
      // GB200_SPGEMM_ROWS=1 selects the warp-per-mask-row form (the reference's
      // work decomposition) for comparison.
      static const bool row_form = getEnv("GB200_SPGEMM_ROWS", 0) != 0;
      if (row_form)
        spgemmMaskedKernel<<<grid, GB_SPGEMM_NT, 0, s>>>(C->d_csrVal_,
            sparse_mask->d_csrRowPtr_, sparse_mask->d_csrColInd_,
            sparse_mask->d_csrVal_, extractMul(op), extractAdd(op),
            static_cast<c>(op.identity()), A_csrRowPtr, A_csrColInd, A_csrVal,
            B_cscColPtr, B_cscRowInd, B_cscVal, A_nrows, work, prof_cell);
      else {
        // thread per mask entry; entries whose lists are both long are deferred
        // to a warp-per-entry kernel through a device-side list (`work` counts it)
        Index* heavy = reinterpret_cast<Index*>(desc->scratch(GB_SCRATCH_VEC_A,
            2*static_cast<size_t>(sparse_mask->nvals_ + 1)*sizeof(Index)));
        spgemmMaskedEdgeKernel<<<grid, GB_SPGEMM_NT, 0, s>>>(C->d_csrVal_,
            sparse_mask->d_csrRowPtr_, sparse_mask->d_csrColInd_,
            sparse_mask->d_csrVal_, extractMul(op), extractAdd(op),
            static_cast<c>(op.identity()), A_csrRowPtr, A_csrColInd, A_csrVal,
            B_cscColPtr, B_cscRowInd, B_cscVal, A_nrows, sparse_mask->nvals_,
            heavy, work, prof_cell);
        GB_KERNEL_CHECK();
        spgemmMaskedHeavyKernel<<<grid, GB_SPGEMM_NT, 0, s>>>(C->d_csrVal_,
            sparse_mask->d_csrColInd_, extractMul(op), extractAdd(op),
            static_cast<c>(op.identity()), A_csrRowPtr, A_csrColInd, A_csrVal,
            B_cscColPtr, B_cscRowInd, B_cscVal, heavy, work);
      }
      GB_KERNEL_CHECK();
      profiler().end(GB_PROF_SPGEMM, s, 8.0*(A_nrows + 1) +
          8.0*sparse_mask->nvals_);
    }
  }
  C->need_update_ = true;
  C->csr_initialized_ = true;
  C->csc_initialized_ = false;
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_SPGEMM_HPP_
