// graphblast_b200 backend — apply: a unary operator over the stored values of a
// vector or matrix.
//
// Stands in for reference graphblas/backend/cuda/apply.hpp:14-117.  The operators
// the algorithms pass are STATEFUL host functors (set_uniform_random for the SSSP
// weights, example/gsssp.cu:79-84; set_random for the MIS / colouring priorities,
// algorithm/mis.hpp:132, gc.hpp:69): each call advances a generator, so values must
// be produced one after the other in storage order.  That is why apply runs on the
// host when the descriptor's GrB_BACKEND is GrB_SEQUENTIAL — the only mode any
// caller uses — and mirrors the result to the device; outside every timed region.
// One routine serves all containers (dense vector, sparse vector, sparse matrix; dense
// matrices are outside the hot path and answered by operations.hpp): they only differ
// in where their value array lives and how they are mirrored back.
#ifndef GRAPHBLAS_BACKEND_CUDA_APPLY_HPP_
#define GRAPHBLAS_BACKEND_CUDA_APPLY_HPP_

#include <iostream>

namespace graphblas {
namespace backend {

namespace apply_detail {
// Host value arrays and their length, per container kind.
template <typename T> T*    values(DenseVector<T>* x)  { return x->h_val_; }
template <typename T> Index count(DenseVector<T>* x)   { return x->nvals_; }
template <typename T> T*    values(SparseVector<T>* x) { return x->h_val_; }
template <typename T> Index count(SparseVector<T>* x)  { return x->nvals_; }
template <typename T> T*    values(SparseMatrix<T>* x) { return x->h_csrVal_; }
template <typename T> Index count(SparseMatrix<T>* x)  { return x->nvals_; }
// What has to happen after the host values changed.
template <typename T> Info publish(DenseVector<T>* x)  { return x->cpuToGpu(); }
template <typename T> Info publish(SparseVector<T>* x) { return x->cpuToGpu(); }
template <typename T> Info publish(SparseMatrix<T>* x) {
  CHECK(x->syncCpu());                 // CSC values follow the CSR values
  return x->cpuToGpu();
}
}  // namespace apply_detail

// out[k] = op(in[k]) over the stored values, in storage order, on the host.
template <typename Out, typename In, typename MaskT, typename UnaryOpT>
Info applyStored(Out* out, const MaskT* mask, UnaryOpT op, In* in, Descriptor* desc,
                 const char* what) {
  Desc_value where;
  CHECK(desc->get(GrB_BACKEND, &where));
  if (desc->debug()) std::cout << "Executing apply on " << what << "\n";
  if (where != GrB_SEQUENTIAL) {
    std::cout << "Error: " << what << " apply needs GrB_BACKEND = GrB_SEQUENTIAL "
              << "(stateful operators run in storage order on the host)\n";
    return GrB_NOT_IMPLEMENTED;
  }
  if (mask != NULL) {
    std::cout << "Error: masked apply on " << what << " is not implemented\n";
    return GrB_NOT_IMPLEMENTED;
  }
  CHECK(in->gpuToCpu());
  if (reinterpret_cast<void*>(out) != reinterpret_cast<void*>(in)) CHECK(out->gpuToCpu());
  const Index n = apply_detail::count(in);
  for (Index k = 0; k < n; ++k)
    apply_detail::values(out)[k] = op(apply_detail::values(in)[k]);
  return apply_detail::publish(out);
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_APPLY_HPP_
