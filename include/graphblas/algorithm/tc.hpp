// graphblast_b200 — triangle counting on a lower-triangular matrix L:
// ntris = sum( (L * L^T) .* L ), computed as one masked mxm + one reduce, as
// reference graphblas/algorithm/tc.hpp:15-54.  The reference toggles GrB_INP1 at
// entry and never restores it (tc.hpp:24), so consecutive calls alternate
// between L*L^T and L*L; both give the same masked sum.  Here the transpose flag
// is set for the call and restored afterwards.  Returns the time in ms, or -1
// with the failing status in algorithm::lastStatus().
#ifndef GRAPHBLAS_ALGORITHM_TC_HPP_
#define GRAPHBLAS_ALGORITHM_TC_HPP_

#include <limits>
#include <vector>
#include <string>

#include "graphblas/algorithm/common.hpp"

namespace graphblas {
namespace algorithm {

inline float tc(int*               ntris,
                const Matrix<int>* A,     // lower triangular matrix
                Matrix<int>*       B,     // buffer matrix (receives (A*A^T).*A)
                Descriptor*        desc) {
  Desc_value inp1_before;
  GB_ALGO_STEP(desc->get(GrB_INP1, &inp1_before));
  GB_ALGO_STEP(desc->set(GrB_INP1, GrB_TRAN));

  LoopTimer clock(false);
  clock.begin();
  Info err = mxm<int, int, int, int>(B, A, GrB_NULL,
      PlusMultipliesSemiring<int>(), A, A, desc);
  if (err == GrB_SUCCESS)
    err = reduce<int, int>(ntris, GrB_NULL, PlusMonoid<int>(), B, desc);
  float ms = clock.finish();

  GB_ALGO_STEP(desc->set(GrB_INP1, inp1_before));
  if (err != GrB_SUCCESS) {            // not a time: report through lastStatus()
    lastStatus() = err;
    return -1.f;
  }
  if (desc->descriptor_.timing_ > 0)
    std::cout << "tc, " << *ntris << " triangles, " << ms << "\n";
  return ms;
}

}  // namespace algorithm
}  // namespace graphblas

#endif  // GRAPHBLAS_ALGORITHM_TC_HPP_
