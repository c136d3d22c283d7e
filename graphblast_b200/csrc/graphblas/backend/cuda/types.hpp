// graphblast_b200 — sm_100a backend for the GraphBLAS header-only dispatch.
//
// This is the first backend header the frontend pulls in
// (reference: graphblas/types.hpp:13-15 includes <graphblas/backend/cuda/types.hpp>
// before it opens namespace graphblas).  Every toolkit header the backend needs
// (CUDA runtime, CCCL/cub) is included HERE, before the frontend defines the
// templates `first`/`second` and does `using namespace graphblas` at global
// scope (reference graphblas/stddef.hpp:78,85 and graphblas/util.hpp:499) —
// libcu++ headers parsed after those break (`__x.first < __y...`).
//
// Replaces reference graphblas/backend/cuda/types.hpp:7-17 (same two enums).
#ifndef GRAPHBLAS_BACKEND_CUDA_TYPES_HPP_
#define GRAPHBLAS_BACKEND_CUDA_TYPES_HPP_

#include <cuda_runtime.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <cassert>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <random>
#include <string>
#include <type_traits>
#include <typeinfo>
#include <unordered_set>
#include <vector>

namespace graphblas {
namespace backend {

enum SparseMatrixFormat {
  GrB_SPARSE_MATRIX_CSRCSC,
  GrB_SPARSE_MATRIX_CSRONLY,
  GrB_SPARSE_MATRIX_CSCONLY
};

enum LoadBalanceMode {
  GrB_LOAD_BALANCE_SIMPLE,
  GrB_LOAD_BALANCE_TWC,
  GrB_LOAD_BALANCE_MERGE
};

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_TYPES_HPP_
