// graphblast_b200 frontend mirror — umbrella header; include order follows
// reference graphblas/graphblas.hpp:4-17 (backend selector, mmio, types, stddef,
// util, dimension, descriptor, vector, matrix, operations, backend umbrella).
#ifndef GRAPHBLAS_GRAPHBLAS_HPP_
#define GRAPHBLAS_GRAPHBLAS_HPP_

#include "graphblas/backend.hpp"
#include "graphblas/mmio.hpp"
#include "graphblas/types.hpp"
#include "graphblas/stddef.hpp"
#include "graphblas/util.hpp"
#include "graphblas/dimension.hpp"
#include "graphblas/descriptor.hpp"
#include "graphblas/vector.hpp"
#include "graphblas/matrix.hpp"
#include "graphblas/operations.hpp"

#include <graphblas/backend/cuda/cuda.hpp>

#endif  // GRAPHBLAS_GRAPHBLAS_HPP_
