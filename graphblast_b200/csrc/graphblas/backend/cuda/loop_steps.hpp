// graphblast_b200 backend — fused element-wise tails of the SSSP and PageRank loops.
//
// Between two mxv calls the reference's loops run four to six GraphBLAS operations
// over n-vectors (algorithm/sssp.hpp:60-75: eWiseAdd less, eWiseAdd min, masked
// assign, reduce; algorithm/pr.hpp:49-66: copy, eWiseAdd, eWiseMult, eWiseAdd,
// reduce): ~40 us of launches and passes per iteration next to a 0.4 ms SpMV.  When
// every vector involved is dense (the pull-only configuration, and PageRank always)
// this project's algorithm headers call the two routines below instead: one pass
// plus the reduction's fold.  They return GrB_NOT_IMPLEMENTED without touching
// anything when a vector is not dense, and the caller takes the operation-by-
// operation route.  GB200_LOOP_STEPS=0 disables them.
#ifndef GRAPHBLAS_BACKEND_CUDA_LOOP_STEPS_HPP_
#define GRAPHBLAS_BACKEND_CUDA_LOOP_STEPS_HPP_

#include "graphblas/backend/cuda/kernels/loop_steps.cuh"

namespace graphblas {
namespace backend {

inline bool loopStepsEnabled() {
  static const bool on = getEnv("GB200_LOOP_STEPS", 1) != 0;
  return on;
}

template <typename T>
bool denseAndSized(Vector<T>* x, Index n) {
  return x->vec_type_ == GrB_DENSE && x->dense_.nvals_ == n && x->dense_.d_val_ != NULL;
}

// dist and relaxed are updated in place; *improved_count = number of distances
// that went down.
template <typename T>
Info ssspRelaxStep(Vector<T>* dist, Vector<T>* relaxed, T inf, float* improved_count,
                   Descriptor* desc) {
  const Index n = dist->nsize_;
  if (!loopStepsEnabled() || n == 0 || !denseAndSized(dist, n) ||
      !denseAndSized(relaxed, n))
    return GrB_NOT_IMPLEMENTED;
  CHECK(dist->dense_.materialize());
  CHECK(relaxed->dense_.materialize());
  int grid;
  T* partials = reducePartials<T>(n, desc, &grid);
  ssspRelaxKernel<<<grid, GB_REDUCE_NT, 0, gbStream()>>>(dist->dense_.d_val_,
      relaxed->dense_.d_val_, n, inf, partials);
  GB_KERNEL_CHECK();
  dist->dense_.touched();
  relaxed->dense_.touched();
  T count;
  CHECK(reduceFold(&count, PlusMonoid<T>(), partials, grid));
  *improved_count = static_cast<float>(count);
  return GrB_SUCCESS;
}

// rank = contrib + jump; *error2 = sum((rank - rank_before)^2).
template <typename T>
Info prUpdateStep(Vector<T>* rank, Vector<T>* contrib, Vector<T>* rank_before, T jump,
                  float* error2, Descriptor* desc) {
  const Index n = rank_before->nsize_;
  if (!loopStepsEnabled() || n == 0 || !denseAndSized(contrib, n) ||
      !denseAndSized(rank_before, n) || !denseAndSized(rank, n))
    return GrB_NOT_IMPLEMENTED;
  CHECK(contrib->dense_.materialize());
  CHECK(rank_before->dense_.materialize());
  int grid;
  T* partials = reducePartials<T>(n, desc, &grid);
  prUpdateKernel<<<grid, GB_REDUCE_NT, 0, gbStream()>>>(rank->dense_.d_val_,
      contrib->dense_.d_val_, rank_before->dense_.d_val_, n, jump, partials);
  GB_KERNEL_CHECK();
  rank->dense_.touched();
  T total;
  CHECK(reduceFold(&total, PlusMonoid<T>(), partials, grid));
  *error2 = static_cast<float>(total);
  return GrB_SUCCESS;
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_LOOP_STEPS_HPP_
