// graphblast_b200 backend — umbrella header (same role as reference
// graphblas/backend/cuda/cuda.hpp:1-34; included last by graphblas/graphblas.hpp).
#ifndef GRAPHBLAS_BACKEND_CUDA_CUDA_HPP_
#define GRAPHBLAS_BACKEND_CUDA_CUDA_HPP_

#include "graphblas/backend/cuda/types.hpp"
#include "graphblas/backend/cuda/util.hpp"
#include "graphblas/backend/cuda/descriptor.hpp"
#include "graphblas/backend/cuda/kernels/kernels.hpp"
#include "graphblas/backend/cuda/compact.hpp"
#include "graphblas/backend/cuda/sparse_vector.hpp"
#include "graphblas/backend/cuda/dense_vector.hpp"
#include "graphblas/backend/cuda/vector.hpp"
#include "graphblas/backend/cuda/sparse_matrix.hpp"
#include "graphblas/backend/cuda/dense_matrix.hpp"
#include "graphblas/backend/cuda/matrix.hpp"
#include "graphblas/backend/cuda/pull_summary.hpp"
#include "graphblas/backend/cuda/spmv.hpp"
#include "graphblas/backend/cuda/spmspv.hpp"
#include "graphblas/backend/cuda/spgemm.hpp"
#include "graphblas/backend/cuda/ewiseadd.hpp"
#include "graphblas/backend/cuda/ewisemult.hpp"
#include "graphblas/backend/cuda/assign.hpp"
#include "graphblas/backend/cuda/reduce.hpp"
#include "graphblas/backend/cuda/apply.hpp"
#include "graphblas/backend/cuda/tri.hpp"
#include "graphblas/backend/cuda/bfs_fused.hpp"
#include "graphblas/backend/cuda/loop_steps.hpp"
#include "graphblas/backend/cuda/operations.hpp"

#endif  // GRAPHBLAS_BACKEND_CUDA_CUDA_HPP_
