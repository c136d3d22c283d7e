// graphblast_b200 — single-source shortest paths (Bellman-Ford style relaxation)
// as a loop of GraphBLAS operations.
//
// Operation sequence per round is the reference's
// (graphblas/algorithm/sssp.hpp:53-94):
//   f2 = f1 (min.+) A ; m = (f2 < v) ; v = min(v, f2) ; f2<!m> = inf (prune
//   vertices that did not improve) ; swap(f1,f2) ; stop when f1 has no entries
//   or reduce(+, m) == 0.
// Unreached vertices keep FLT_MAX.  Returns the device time of the loop in ms.
#ifndef GRAPHBLAS_ALGORITHM_SSSP_HPP_
#define GRAPHBLAS_ALGORITHM_SSSP_HPP_

#include <limits>
#include <vector>
#include <string>

#include "graphblas/algorithm/common.hpp"

namespace graphblas {
namespace algorithm {

inline float sssp(Vector<float>* v, const Matrix<float>* A, Index s, Descriptor* desc) {
  const float kInf = std::numeric_limits<float>::max();
  Index n;
  GB_ALGO_STEP(A->nrows(&n));

  GB_ALGO_STEP(v->fill(kInf));
  GB_ALGO_STEP(v->setElement(0.f, s));

  Vector<float> frontier(n);
  Vector<float> relaxed(n);
  Vector<float> improved(n);

  Desc_value mxv_mode;
  GB_ALGO_STEP(desc->get(GrB_MXVMODE, &mxv_mode));
  if (mxv_mode == GrB_PULLONLY) {
    GB_ALGO_STEP(frontier.fill(kInf));
    GB_ALGO_STEP(frontier.setElement(0.f, s));
  } else {
    std::vector<Index> src_ind(1, s);
    std::vector<float> src_val(1, 0.f);
    GB_ALGO_STEP(frontier.build(&src_ind, &src_val, 1, GrB_NULL));
  }

  backend::Descriptor& d = desc->descriptor_;
  const bool verbose = (d.timing_ == 1);
  LoopTimer clock(verbose);
  Index frontier_nvals = 1;
  float succ = 1.f;
  clock.begin();

  for (int round = 1; round <= d.max_niter_; ++round) {
    vxm<float, float, float, float>(&relaxed, GrB_NULL, GrB_NULL,
        MinimumPlusSemiring<float>(), &frontier, A, desc);
    // improved = relaxed < v ; v = min(v, relaxed) ; relaxed<!improved> = inf ;
    // succ = |improved| — one pass when everything is dense (backend loop_steps.hpp),
    // else the reference's four operations (algorithm/sssp.hpp:60-75).
    if (backend::ssspRelaxStep(&v->vector_, &relaxed.vector_, kInf, &succ,
                               &desc->descriptor_) == GrB_SUCCESS) {
      GB_ALGO_STEP(relaxed.swap(&frontier));
      GB_ALGO_STEP(frontier.nvals(&frontier_nvals));
    } else {
      eWiseAdd<float, float, float, float>(&improved, GrB_NULL, GrB_NULL,
          CustomLessPlusSemiring<float>(), &relaxed, v, desc);
      eWiseAdd<float, float, float, float>(v, GrB_NULL, GrB_NULL,
          MinimumPlusSemiring<float>(), v, &relaxed, desc);

      GB_ALGO_STEP(desc->toggle(GrB_MASK));
      assign<float, float, float, Index>(&relaxed, &improved, GrB_NULL, kInf,
          GrB_ALL, n, desc);
      GB_ALGO_STEP(desc->toggle(GrB_MASK));

      GB_ALGO_STEP(relaxed.swap(&frontier));
      GB_ALGO_STEP(frontier.nvals(&frontier_nvals));
      reduce<float, float>(&succ, GrB_NULL, PlusMonoid<float>(), &improved, desc);
    }

    if (verbose) {
      float ms = clock.lap();
      std::cout << round << ", " << frontier_nvals << "/" << n << ", "
                << (d.lastmxv_ == GrB_PUSHONLY ? "push" : "pull") << ", "
                << ms << "\n";
    }
    if (frontier_nvals == 0 || succ == 0) break;
  }
  return clock.finish();
}

}  // namespace algorithm
}  // namespace graphblas

#endif  // GRAPHBLAS_ALGORITHM_SSSP_HPP_
