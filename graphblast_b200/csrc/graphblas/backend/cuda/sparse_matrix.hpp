// graphblast_b200 backend — SparseMatrix<T>: CSR + CSC, host and device mirrors.
//
// Replaces reference graphblas/backend/cuda/sparse_matrix.hpp:24-853.  Member
// names are the reference's because the CPU verifiers read them directly
// (reference algorithm/bfs.hpp:101-107: matrix_.sparse_.h_csrRowPtr_ ...).
// Layout in HBM: int32 rowptr[nrows+1], int32 colind[nvals], T val[nvals] for
// CSR, and the same triple for CSC; when the matrix is structurally symmetric
// (".ud." cache name, reference :300-306) the CSC index arrays ALIAS the CSR ones
// and only cscVal is separate (reference cpuToGpu :789-797).  All device arrays
// come from the stream-ordered pool and are 256-byte aligned, which the pull
// kernel's 256-bit loads rely on.
//
// Extra entry points for the C-ABI / large graphs (no reference counterpart):
// adoptCsc() and the ownership flags let a caller hand over CSR and CSC arrays
// that already live in device memory (the reference can only adopt CSR, :418-435).
#ifndef GRAPHBLAS_BACKEND_CUDA_SPARSE_MATRIX_HPP_
#define GRAPHBLAS_BACKEND_CUDA_SPARSE_MATRIX_HPP_

#include <vector>
#include <iostream>
#include <cassert>
#include <algorithm>

#include "graphblas/backend/cuda/util.hpp"
#include "graphblas/backend/cuda/hub_index.hpp"

namespace graphblas {
namespace backend {

template <typename T>
class DenseMatrix;

template <typename T>
class Vector;

template <typename T>
class SparseMatrix {
 public:
  SparseMatrix() { init(0, 0); }
  explicit SparseMatrix(Index nrows, Index ncols) { init(nrows, ncols); }

  ~SparseMatrix();

  // C API Methods
  Info nnew(Index nrows, Index ncols);
  Info dup(const SparseMatrix* rhs);
  Info clear();     // 1 way to free: (1) clear
  Info nrows(Index* nrows_t) const;
  Info ncols(Index* ncols_t) const;
  Info nvals(Index* nvals_t) const;
  template <typename BinaryOpT>
  Info build(const std::vector<Index>* row_indices,
      const std::vector<Index>* col_indices, const std::vector<T>* values, Index nvals,
      BinaryOpT dup, char* dat_name);
  Info build(char* dat_name);
  Info build(const std::vector<T>* values, Index nvals);
  Info build(Index* row_ptr, Index* col_ind, T* values, Index nvals);
  // Device-resident CSC (or aliasing request) supplied by the caller.
  Info adoptCsc(Index* col_ptr, Index* row_ind, T* values, bool symmetric);
  Info setElement(Index row_index, Index col_index);
  Info extractElement(T* val, Index row_index, Index col_index);
  Info extractTuples(std::vector<Index>* row_indices, std::vector<Index>* col_indices,
      std::vector<T>* values, Index* n);
  Info extractTuples(std::vector<T>* values, Index* n);

  // Handy methods
  const T operator[](Index ind);
  Info print(bool force_update);
  Info check();
  Info setNrows(Index nrows);
  Info setNcols(Index ncols);
  Info setNvals(Index nvals);
  Info getFormat(SparseMatrixFormat* format) const;
  Info getSymmetry(bool* symmetry) const;
  Info resize(Index nrows, Index ncols);
  template <typename U>
  Info fill(Index axis, Index nvals, U start);
  template <typename U>
  Info fillAscending(Index axis, Index nvals, U start);

 public:  // (private in the reference; its drivers `#define private public`)
  void init(Index nrows, Index ncols);
  void freeHost();
  void freeDevice();
  Info allocateCpu();
  Info allocateGpu();
  Info allocate();  // 3 ways to allocate: (1) dup, (2) build, (3) spgemm
  Info printCSR(const char* str);  // private method for pretty printing
  Info printCSC(const char* str);
  Info cpuToGpu();
  Info gpuToCpu(bool force_update = false);

  Info syncCpu();   // synchronizes CSR and CSC representations

  Index nrows_;
  Index ncols_;
  Index nvals_;     // 3 ways to set: (1) dup (2) build (3) nnew
  Index ncapacity_;
  Index nempty_;

  Index* h_csrRowPtr_;  // CSR format
  Index* h_csrColInd_;
  T*     h_csrVal_;
  Index* h_cscColPtr_;  // CSC format
  Index* h_cscRowInd_;
  T*     h_cscVal_;

  Index* d_csrRowPtr_;  // GPU CSR format
  Index* d_csrColInd_;
  T*     d_csrVal_;
  Index* d_cscColPtr_;  // GPU CSC format
  Index* d_cscRowInd_;
  T*     d_cscVal_;

  bool need_update_;
  bool csr_initialized_;
  bool csc_initialized_;
  bool csr_ownership_;   // device CSR arrays owned by this object
  bool csc_ownership_;   // device CSC index arrays owned by this object
  bool cscval_ownership_;  // device CSC value array owned by this object
  bool symmetric_;

  SparseMatrixFormat format_;

  // Cached merge-path tile partitions of the pull SpMV (one per traversed
  // structure: 0 = CSR rows, 1 = CSC columns), valid for the rowptr they were
  // computed from.
  Index*       d_spmv_tiles_[2];
  const Index* spmv_tiles_key_[2];
  Index        spmv_tiles_nvals_[2];
  int          spmv_tiles_count_[2];
  // Cached "first neighbour" summary of the Boolean pull (same indexing):
  // entry i = -1 for an empty row, colind[rowptr[i]] for a longer row, and that
  // value with the top bit set when it is the row's only entry.
  Index*       d_pull_first_[2];
  const Index* pull_first_key_[2];
  Index        pull_first_nvals_[2];
  // Hub index of the hub-cached pull SpMV (same indexing): which columns live in
  // shared memory, encoded column array, compact non-empty rows, tile records.
  // hub_state_: 0 = not built, 1 = built and used, 2 = built and rejected (too
  // little of the matrix references the hub columns).
  HubIndex     hub_[2];
  int          hub_state_[2] = {0, 0};
  void dropSpmvTiles() {
    for (int k = 0; k < 2; ++k) {
      hub_[k].release();
      hub_state_[k] = 0;
      if (d_pull_first_[k] != NULL) gbFree(d_pull_first_[k]);
      d_pull_first_[k] = NULL;
      pull_first_key_[k] = NULL;
      pull_first_nvals_[k] = -1;
      if (d_spmv_tiles_[k] != NULL) gbFree(d_spmv_tiles_[k]);
      d_spmv_tiles_[k] = NULL;
      spmv_tiles_key_[k] = NULL;
      spmv_tiles_nvals_[k] = -1;
      spmv_tiles_count_[k] = 0;
    }
  }
};

template <typename T>
void SparseMatrix<T>::init(Index nrows, Index ncols) {
  nrows_ = nrows; ncols_ = ncols; nvals_ = 0; ncapacity_ = 0; nempty_ = 0;
  h_csrRowPtr_ = NULL; h_csrColInd_ = NULL; h_csrVal_ = NULL;
  h_cscColPtr_ = NULL; h_cscRowInd_ = NULL; h_cscVal_ = NULL;
  d_csrRowPtr_ = NULL; d_csrColInd_ = NULL; d_csrVal_ = NULL;
  d_cscColPtr_ = NULL; d_cscRowInd_ = NULL; d_cscVal_ = NULL;
  need_update_ = false;
  csr_initialized_ = false; csc_initialized_ = false;
  csr_ownership_ = false;   csc_ownership_ = false;
  cscval_ownership_ = false;
  symmetric_ = false;
  format_ = getEnv("GRB_SPARSE_MATRIX_FORMAT", GrB_SPARSE_MATRIX_CSRCSC);
  for (int k = 0; k < 2; ++k) {
    d_pull_first_[k] = NULL;
    pull_first_key_[k] = NULL;
    pull_first_nvals_[k] = -1;
    d_spmv_tiles_[k] = NULL;
    spmv_tiles_key_[k] = NULL;
    spmv_tiles_nvals_[k] = -1;
    spmv_tiles_count_[k] = 0;
  }
}

template <typename T>
void SparseMatrix<T>::freeHost() {
  bool csc_is_alias = (h_cscColPtr_ == h_csrRowPtr_);
  if (h_csrRowPtr_) free(h_csrRowPtr_);
  if (h_csrColInd_) free(h_csrColInd_);
  if (h_csrVal_   ) free(h_csrVal_);
  if (!csc_is_alias) {
    if (h_cscColPtr_) free(h_cscColPtr_);
    if (h_cscRowInd_) free(h_cscRowInd_);
    if (h_cscVal_   ) free(h_cscVal_);
  }
  h_csrRowPtr_ = NULL; h_csrColInd_ = NULL; h_csrVal_ = NULL;
  h_cscColPtr_ = NULL; h_cscRowInd_ = NULL; h_cscVal_ = NULL;
}

template <typename T>
void SparseMatrix<T>::freeDevice() {
  dropSpmvTiles();
  if (csc_ownership_) {
    if (d_cscColPtr_ && d_cscColPtr_ != d_csrRowPtr_) gbFree(d_cscColPtr_);
    if (d_cscRowInd_ && d_cscRowInd_ != d_csrColInd_) gbFree(d_cscRowInd_);
  }
  if (cscval_ownership_ && d_cscVal_ && d_cscVal_ != d_csrVal_)
    gbFree(d_cscVal_);
  if (csr_ownership_) {
    if (d_csrRowPtr_) gbFree(d_csrRowPtr_);
    if (d_csrColInd_) gbFree(d_csrColInd_);
    if (d_csrVal_   ) gbFree(d_csrVal_);
  }
  d_csrRowPtr_ = NULL; d_csrColInd_ = NULL; d_csrVal_ = NULL;
  d_cscColPtr_ = NULL; d_cscRowInd_ = NULL; d_cscVal_ = NULL;
  csr_ownership_ = false; csc_ownership_ = false;
  cscval_ownership_ = false;
}

template <typename T>
SparseMatrix<T>::~SparseMatrix() {
  freeHost();
  freeDevice();
}

template <typename T>
Info SparseMatrix<T>::nnew(Index nrows, Index ncols) {
  nrows_ = nrows;
  ncols_ = ncols;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::dup(const SparseMatrix* rhs) {
  if (nrows_ != rhs->nrows_) return GrB_DIMENSION_MISMATCH;
  if (ncols_ != rhs->ncols_) return GrB_DIMENSION_MISMATCH;
  if (nvals_ != rhs->nvals_ || symmetric_ != rhs->symmetric_ ||
      !csr_ownership_) {
    freeDevice();
    freeHost();
  }
  nvals_     = rhs->nvals_;
  symmetric_ = rhs->symmetric_;
  format_    = rhs->format_;

  CHECK(allocateGpu());
  cudaStream_t s = gbStream();
  CUDA_CALL(cudaMemcpyAsync(d_csrRowPtr_, rhs->d_csrRowPtr_, (nrows_+1)*sizeof(Index),
      cudaMemcpyDeviceToDevice, s));
  if (nvals_ > 0) {
    CUDA_CALL(cudaMemcpyAsync(d_csrColInd_, rhs->d_csrColInd_,
        static_cast<size_t>(nvals_)*sizeof(Index), cudaMemcpyDeviceToDevice, s));
    CUDA_CALL(cudaMemcpyAsync(d_csrVal_, rhs->d_csrVal_,
        static_cast<size_t>(nvals_)*sizeof(T), cudaMemcpyDeviceToDevice, s));
  }
  if (format_ == GrB_SPARSE_MATRIX_CSRCSC && rhs->d_cscVal_ != NULL) {
    if (nvals_ > 0)
      CUDA_CALL(cudaMemcpyAsync(d_cscVal_, rhs->d_cscVal_,
          static_cast<size_t>(nvals_)*sizeof(T), cudaMemcpyDeviceToDevice, s));
    if (!symmetric_ && rhs->d_cscColPtr_ != NULL && rhs->d_cscRowInd_ != NULL) {
      CUDA_CALL(cudaMemcpyAsync(d_cscColPtr_, rhs->d_cscColPtr_,
          (ncols_+1)*sizeof(Index), cudaMemcpyDeviceToDevice, s));
      if (nvals_ > 0)
        CUDA_CALL(cudaMemcpyAsync(d_cscRowInd_, rhs->d_cscRowInd_,
            static_cast<size_t>(nvals_)*sizeof(Index), cudaMemcpyDeviceToDevice, s));
    }
    csc_initialized_ = true;
  }
  need_update_ = true;
  csr_initialized_ = true;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::clear() {
  nvals_     = 0;
  ncapacity_ = 0;
  freeHost();
  freeDevice();
  csr_initialized_ = false;
  csc_initialized_ = false;
  return GrB_SUCCESS;
}

template <typename T>
inline Info SparseMatrix<T>::nrows(Index* nrows_t) const {
  *nrows_t = nrows_;
  return GrB_SUCCESS;
}

template <typename T>
inline Info SparseMatrix<T>::ncols(Index* ncols_t) const {
  *ncols_t = ncols_;
  return GrB_SUCCESS;
}

template <typename T>
inline Info SparseMatrix<T>::nvals(Index* nvals_t) const {
  *nvals_t = nvals_;
  return GrB_SUCCESS;
}

// Host COO -> CSR and CSC, optional ".bin" cache (reference :291-351; cache
// layout int32 nrows, int32 nvals, rowptr[nrows+1], colind[nvals]), then upload.
template <typename T>
template <typename BinaryOpT>
Info SparseMatrix<T>::build(const std::vector<Index>* row_indices,
    const std::vector<Index>* col_indices, const std::vector<T>* values, Index nvals,
    BinaryOpT dup, char* dat_name) {
  freeHost();
  freeDevice();
  nvals_ = nvals;
  CHECK(allocateCpu());

  if (dat_name != NULL)
    symmetric_ = (strstr(dat_name, ".ud.") != NULL);

  coo2csr(h_csrRowPtr_, h_csrColInd_, h_csrVal_, *row_indices, *col_indices, *values,
      nrows_, ncols_);

  if (format_ == GrB_SPARSE_MATRIX_CSRONLY) {
    if (h_cscColPtr_ != NULL) free(h_cscColPtr_);
    if (h_cscRowInd_ != NULL) free(h_cscRowInd_);
    if (h_cscVal_    != NULL) free(h_cscVal_);
    h_cscColPtr_ = h_csrRowPtr_;
    h_cscRowInd_ = h_csrColInd_;
    h_cscVal_    = h_csrVal_;
  } else {
    coo2csc(h_cscColPtr_, h_cscRowInd_, h_cscVal_, *row_indices, *col_indices, *values,
        nrows_, ncols_);
    csc_initialized_ = true;
  }
  csr_initialized_ = true;

  if (dat_name != NULL) {
    if (!exists(dat_name)) {
      std::ofstream ofs(dat_name, std::ios::out | std::ios::binary);
      if (ofs.fail()) {
        std::cout << "Error: Unable to open file for writing!\n";
      } else {
        printf("Writing %s\n", dat_name);
        ofs.write(reinterpret_cast<char*>(&nrows_), sizeof(Index));
        if (ncols_ != nrows_)
          std::cout << "Error: nrows not equal to ncols!\n";
        ofs.write(reinterpret_cast<char*>(&nvals_), sizeof(Index));
        ofs.write(reinterpret_cast<char*>(h_csrRowPtr_),
            (nrows_+1)*sizeof(Index));
        ofs.write(reinterpret_cast<char*>(h_csrColInd_),
            static_cast<size_t>(nvals_)*sizeof(Index));
        ofs.close();
      }
    }
    free(dat_name);
  }

  CHECK(cpuToGpu());
  return GrB_SUCCESS;
}

// Reload from the ".bin" cache; values become 1 (reference :354-407).
template <typename T>
Info SparseMatrix<T>::build(char* dat_name) {
  if (dat_name != NULL && exists(dat_name)) {
    std::ifstream ifs(dat_name, std::ios::in | std::ios::binary);
    if (ifs.fail()) {
      std::cout << "Error: Unable to open file for reading!\n";
    } else {
      printf("Reading %s\n", dat_name);
      freeHost();
      freeDevice();
      symmetric_ = (strstr(dat_name, ".ud.") != NULL);

      ifs.read(reinterpret_cast<char*>(&nrows_), sizeof(Index));
      if (ncols_ != nrows_)
        std::cout << "Error: nrows not equal to ncols!\n";
      ifs.read(reinterpret_cast<char*>(&nvals_), sizeof(Index));
      CHECK(allocateCpu());

      ifs.read(reinterpret_cast<char*>(h_csrRowPtr_),
          (nrows_+1)*sizeof(Index));
      ifs.read(reinterpret_cast<char*>(h_csrColInd_),
          static_cast<size_t>(nvals_)*sizeof(Index));

      for (Index i = 0; i < nvals_; i++)
        h_csrVal_[i] = static_cast<T>(1);

      if (format_ == GrB_SPARSE_MATRIX_CSRONLY) {
        if (h_cscColPtr_ != NULL) free(h_cscColPtr_);
        if (h_cscRowInd_ != NULL) free(h_cscRowInd_);
        if (h_cscVal_    != NULL) free(h_cscVal_);
        h_cscColPtr_ = h_csrRowPtr_;
        h_cscRowInd_ = h_csrColInd_;
        h_cscVal_    = h_csrVal_;
      } else {
        csr2csc(h_cscColPtr_, h_cscRowInd_, h_cscVal_, h_csrRowPtr_, h_csrColInd_,
            h_csrVal_, nrows_, ncols_);
        csc_initialized_ = true;
      }
      csr_initialized_ = true;

      CHECK(cpuToGpu());
    }
    free(dat_name);
  } else {
    std::cout << "Error: Unable to read file!\n";
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::build(const std::vector<T>* values, Index nvals) {
  std::cout << "SparseMatrix Build from dense input\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

// Adopts DEVICE CSR arrays without taking ownership (reference :418-435).
template <typename T>
Info SparseMatrix<T>::build(Index* row_ptr, Index* col_ind, T* values, Index nvals) {
  freeDevice();
  freeHost();
  d_csrRowPtr_ = row_ptr;
  d_csrColInd_ = col_ind;
  d_csrVal_    = values;

  nvals_ = nvals;
  need_update_ = true;
  csr_initialized_ = true;
  csr_ownership_ = false;
  return GrB_SUCCESS;
}

// symmetric == true: CSC index arrays alias the CSR ones (col_ptr/row_ind may be
// NULL).  values == NULL asks for an owned copy of the CSR values (only correct
// when the values are symmetric, e.g. a pattern matrix).
template <typename T>
Info SparseMatrix<T>::adoptCsc(Index* col_ptr, Index* row_ind, T* values,
    bool symmetric) {
  if (d_csrRowPtr_ == NULL) return GrB_UNINITIALIZED_OBJECT;
  dropSpmvTiles();
  symmetric_ = symmetric;
  if (symmetric && (col_ptr == NULL || row_ind == NULL)) {
    d_cscColPtr_ = d_csrRowPtr_;
    d_cscRowInd_ = d_csrColInd_;
  } else {
    d_cscColPtr_ = col_ptr;
    d_cscRowInd_ = row_ind;
  }
  csc_ownership_ = false;
  if (values != NULL) {
    d_cscVal_ = values;
    cscval_ownership_ = false;
  } else {
    // Independent copy: operations that rescale values per row (PageRank
    // normalisation) must be able to make CSR and CSC values differ.
    const size_t nv = nvals_ > 0 ? nvals_ : 1;
    d_cscVal_ = reinterpret_cast<T*>(gbMalloc(nv*sizeof(T)));
    CUDA_CALL(cudaMemcpyAsync(d_cscVal_, d_csrVal_, nv*sizeof(T), cudaMemcpyDeviceToDevice, gbStream()));
    cscval_ownership_ = true;
  }
  csc_initialized_ = true;
  need_update_     = true;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::setElement(Index row_index, Index col_index) {
  std::cout << "SparseMatrix setElement\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::extractElement(T* val, Index row_index, Index col_index) {
  std::cout << "SparseMatrix extractElement\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::extractTuples(std::vector<Index>* row_indices,
    std::vector<Index>* col_indices, std::vector<T>* values, Index* n) {
  CHECK(gpuToCpu());
  row_indices->clear();
  col_indices->clear();
  values->clear();

  if (*n > nvals_) {
    std::cout << "Error: Too many tuples requested!\n";
    return GrB_UNINITIALIZED_OBJECT;
  }
  if (*n < nvals_) {
    std::cout << "Error: Insufficient space!\n";
    return GrB_INSUFFICIENT_SPACE;
  }

  for (Index row = 0; row < nrows_; row++) {
    for (Index k = h_csrRowPtr_[row]; k < h_csrRowPtr_[row+1]; k++) {
      if (h_csrColInd_[k] >= 0 && static_cast<Index>(values->size()) < *n) {
        row_indices->push_back(row);
        col_indices->push_back(h_csrColInd_[k]);
        values->push_back(h_csrVal_[k]);
      }
    }
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::extractTuples(std::vector<T>* values, Index* n) {
  std::cout << "SparseMatrix extractTuples into dense\n";
  std::cout << "Error: Feature not implemented yet!\n";
  return GrB_SUCCESS;
}

template <typename T>
const T SparseMatrix<T>::operator[](Index ind) {
  gpuToCpu(true);
  if (ind >= nvals_) std::cout << "Error: index out of bounds!\n";
  return h_csrColInd_[ind];
}

template <typename T>
Info SparseMatrix<T>::print(bool force_update) {
  CHECK(gpuToCpu(force_update));
  printArray("csrColInd", h_csrColInd_, std::min(nvals_, 40));
  printArray("csrRowPtr", h_csrRowPtr_, std::min(nrows_+1, 40));
  printArray("csrVal",    h_csrVal_,    std::min(nvals_, 40));
  CHECK(printCSR("pretty print"));
  if (format_ == GrB_SPARSE_MATRIX_CSRCSC && h_cscColPtr_ != NULL) {
    printArray("cscRowInd", h_cscRowInd_, std::min(nvals_, 40));
    printArray("cscColPtr", h_cscColPtr_, std::min(ncols_+1, 40));
    printArray("cscVal",    h_cscVal_,    std::min(nvals_, 40));
    CHECK(printCSC("pretty print"));
  }
  return GrB_SUCCESS;
}

// Row pointers monotone, column indices strictly increasing inside a row.
template <typename T>
Info SparseMatrix<T>::check() {
  CHECK(gpuToCpu());
  std::cout << "Begin check:\n";
  for (Index row = 0; row < nrows_; row++)
    assert(h_csrRowPtr_[row+1] >= h_csrRowPtr_[row]);
  for (Index row = 0; row < nrows_; row++) {
    for (Index k = h_csrRowPtr_[row]; k + 1 < h_csrRowPtr_[row+1]; k++) {
      assert(h_csrColInd_[k] != -1);
      assert(h_csrColInd_[k+1] > h_csrColInd_[k]);
    }
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::setNrows(Index nrows) {
  nrows_ = nrows;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::setNcols(Index ncols) {
  ncols_ = ncols;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::setNvals(Index nvals) {
  nvals_ = nvals;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::getFormat(SparseMatrixFormat* format) const {
  *format = format_;
  return GrB_SUCCESS;
}

// Always reports false, as the reference does (:578-582); symmetric_ still
// drives the CSR/CSC aliasing on the device.
template <typename T>
Info SparseMatrix<T>::getSymmetry(bool* symmetry) const {
  *symmetry = false;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::resize(Index nrows, Index ncols) {
  if (nrows <= nrows_) nrows_ = nrows;
  else return GrB_PANIC;
  if (ncols <= ncols_) ncols_ = ncols;
  else return GrB_PANIC;
  return GrB_SUCCESS;
}

template <typename T>
template <typename U>
Info SparseMatrix<T>::fill(Index axis, Index nvals, U start) {
  CHECK(setNvals(nvals));
  CHECK(allocate());
  if (axis == 0) {
    for (Index i = 0; i < nvals; i++) h_csrRowPtr_[i] = static_cast<Index>(start);
  } else if (axis == 1) {
    for (Index i = 0; i < nvals; i++) h_csrColInd_[i] = static_cast<Index>(start);
  } else if (axis == 2) {
    for (Index i = 0; i < nvals; i++) h_csrVal_[i] = static_cast<T>(start);
  }
  CHECK(cpuToGpu());
  return GrB_SUCCESS;
}

template <typename T>
template <typename U>
Info SparseMatrix<T>::fillAscending(Index axis, Index nvals, U start) {
  CHECK(setNvals(nvals));
  CHECK(allocate());
  if (axis == 0) {
    for (Index i = 0; i < nvals; i++)
      h_csrRowPtr_[i] = i + static_cast<Index>(start);
  } else if (axis == 1) {
    for (Index i = 0; i < nvals; i++)
      h_csrColInd_[i] = i + static_cast<Index>(start);
  } else if (axis == 2) {
    for (Index i = 0; i < nvals; i++)
      h_csrVal_[i] = static_cast<T>(i) + static_cast<T>(start);
  }
  CHECK(cpuToGpu());
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::allocateCpu() {
  ncapacity_ = nvals_;
  const size_t nv = nvals_ > 0 ? nvals_ : 1;
  if (h_csrRowPtr_ == NULL)
    h_csrRowPtr_ = reinterpret_cast<Index*>(malloc((nrows_+1)*sizeof(Index)));
  if (h_csrColInd_ == NULL)
    h_csrColInd_ = reinterpret_cast<Index*>(malloc(nv*sizeof(Index)));
  if (h_csrVal_ == NULL)
    h_csrVal_ = reinterpret_cast<T*>(malloc(nv*sizeof(T)));
  if (format_ != GrB_SPARSE_MATRIX_CSRONLY) {
    if (h_cscColPtr_ == NULL)
      h_cscColPtr_ = reinterpret_cast<Index*>(malloc((ncols_+1)*sizeof(Index)));
    if (h_cscRowInd_ == NULL)
      h_cscRowInd_ = reinterpret_cast<Index*>(malloc(nv*sizeof(Index)));
    if (h_cscVal_ == NULL)
      h_cscVal_ = reinterpret_cast<T*>(malloc(nv*sizeof(T)));
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::allocateGpu() {
  const size_t nv = nvals_ > 0 ? nvals_ : 1;
  if (d_csrRowPtr_ == NULL) {
    d_csrRowPtr_ = reinterpret_cast<Index*>(gbMalloc((nrows_+1)*sizeof(Index)));
    d_csrColInd_ = reinterpret_cast<Index*>(gbMalloc(nv*sizeof(Index)));
    d_csrVal_    = reinterpret_cast<T*>(gbMalloc(nv*sizeof(T)));
    csr_ownership_ = true;
    printMemory("csrVal");
  }
  if (format_ == GrB_SPARSE_MATRIX_CSRCSC && d_cscVal_ == NULL) {
    d_cscVal_ = reinterpret_cast<T*>(gbMalloc(nv*sizeof(T)));
    if (!symmetric_) {
      d_cscColPtr_ = reinterpret_cast<Index*>(
          gbMalloc((ncols_+1)*sizeof(Index)));
      d_cscRowInd_ = reinterpret_cast<Index*>(gbMalloc(nv*sizeof(Index)));
    } else {
      d_cscColPtr_ = d_csrRowPtr_;
      d_cscRowInd_ = d_csrColInd_;
    }
    csc_ownership_ = true;
    cscval_ownership_ = true;
    printMemory("cscVal");
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::allocate() {
  CHECK(allocateCpu());
  CHECK(allocateGpu());
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::printCSR(const char* str) {
  Index row_length = std::min(20, nrows_);
  Index col_length = std::min(20, ncols_);
  std::cout << str << ":\n";
  for (Index row = 0; row < row_length; row++) {
    Index k   = h_csrRowPtr_[row];
    Index end = h_csrRowPtr_[row+1];
    for (Index col = 0; col < col_length; col++) {
      if (k < end && h_csrColInd_[k] == col && h_csrVal_[k] > 0) {
        std::cout << "x ";
        k++;
      } else {
        std::cout << "0 ";
      }
    }
    std::cout << std::endl;
  }
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::printCSC(const char* str) {
  Index row_length = std::min(20, nrows_);
  Index col_length = std::min(20, ncols_);
  std::cout << str << ":\n";
  for (Index col = 0; col < col_length; col++) {
    Index k   = h_cscColPtr_[col];
    Index end = h_cscColPtr_[col+1];
    for (Index row = 0; row < row_length; row++) {
      if (k < end && h_cscRowInd_[k] == row && h_cscVal_[k] > 0) {
        std::cout << "x ";
        k++;
      } else {
        std::cout << "0 ";
      }
    }
    std::cout << std::endl;
  }
  return GrB_SUCCESS;
}

// Host -> device.  If the stored entry count changed (tril) or the device arrays
// are not ours, the device side is re-created first.
template <typename T>
Info SparseMatrix<T>::cpuToGpu() {
  if (!csr_ownership_ || ncapacity_ != nvals_ || d_csrRowPtr_ == NULL) {
    freeDevice();
    ncapacity_ = nvals_;
  }
  CHECK(allocateGpu());
  dropSpmvTiles();   // derived caches describe the previous contents
  cudaStream_t s = gbStream();
  const size_t nv = nvals_;

  CUDA_CALL(cudaMemcpyAsync(d_csrRowPtr_, h_csrRowPtr_, (nrows_+1)*sizeof(Index), cudaMemcpyHostToDevice, s));
  if (nv > 0) {
    CUDA_CALL(cudaMemcpyAsync(d_csrColInd_, h_csrColInd_, nv*sizeof(Index), cudaMemcpyHostToDevice, s));
    CUDA_CALL(cudaMemcpyAsync(d_csrVal_, h_csrVal_, nv*sizeof(T), cudaMemcpyHostToDevice, s));
  }

  if (format_ == GrB_SPARSE_MATRIX_CSRCSC) {
    if (nv > 0)
      CUDA_CALL(cudaMemcpyAsync(d_cscVal_, h_cscVal_, nv*sizeof(T), cudaMemcpyHostToDevice, s));
    if (!symmetric_) {
      CUDA_CALL(cudaMemcpyAsync(d_cscColPtr_, h_cscColPtr_, (ncols_+1)*sizeof(Index), cudaMemcpyHostToDevice, s));
      if (nv > 0)
        CUDA_CALL(cudaMemcpyAsync(d_cscRowInd_, h_cscRowInd_, nv*sizeof(Index), cudaMemcpyHostToDevice, s));
    } else {
      d_cscColPtr_ = d_csrRowPtr_;
      d_cscRowInd_ = d_csrColInd_;
    }
  }
  runtime().sync();
  need_update_ = false;
  return GrB_SUCCESS;
}

template <typename T>
Info SparseMatrix<T>::gpuToCpu(bool force_update) {
  bool fresh_host = (h_csrRowPtr_ == NULL);
  if (fresh_host) CHECK(allocateCpu());
  if ((need_update_ || force_update || fresh_host) && d_csrRowPtr_ != NULL) {
    cudaStream_t s = gbStream();
    const size_t nv = nvals_;
    CUDA_CALL(cudaMemcpyAsync(h_csrRowPtr_, d_csrRowPtr_, (nrows_+1)*sizeof(Index), cudaMemcpyDeviceToHost, s));
    if (nv > 0) {
      CUDA_CALL(cudaMemcpyAsync(h_csrColInd_, d_csrColInd_, nv*sizeof(Index), cudaMemcpyDeviceToHost, s));
      CUDA_CALL(cudaMemcpyAsync(h_csrVal_, d_csrVal_, nv*sizeof(T), cudaMemcpyDeviceToHost, s));
    }
    if (format_ == GrB_SPARSE_MATRIX_CSRCSC && d_cscVal_ && d_cscColPtr_ &&
        d_cscRowInd_ && h_cscVal_ && h_cscColPtr_ && h_cscRowInd_) {
      if (nv > 0)
        CUDA_CALL(cudaMemcpyAsync(h_cscVal_, d_cscVal_, nv*sizeof(T), cudaMemcpyDeviceToHost, s));
      if (!symmetric_ || fresh_host) {
        CUDA_CALL(cudaMemcpyAsync(h_cscColPtr_, d_cscColPtr_, (ncols_+1)*sizeof(Index), cudaMemcpyDeviceToHost, s));
        if (nv > 0)
          CUDA_CALL(cudaMemcpyAsync(h_cscRowInd_, d_cscRowInd_, nv*sizeof(Index), cudaMemcpyDeviceToHost, s));
      }
    }
    runtime().sync();
  }
  need_update_ = false;
  return GrB_SUCCESS;
}

// Rebuilds the host CSC from the host CSR (reference :836-848).
template <typename T>
Info SparseMatrix<T>::syncCpu() {
  CHECK(allocateCpu());
  if (h_csrRowPtr_ && h_csrColInd_ && h_csrVal_ &&
      h_cscColPtr_ && h_cscRowInd_ && h_cscVal_)
    csr2csc(h_cscColPtr_, h_cscRowInd_, h_cscVal_, h_csrRowPtr_, h_csrColInd_,
        h_csrVal_, nrows_, ncols_);
  else
    return GrB_INVALID_OBJECT;
  return GrB_SUCCESS;
}
}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_SPARSE_MATRIX_HPP_
