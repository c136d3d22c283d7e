// graphblast_b200 frontend mirror — the one header applications include.
// Pulls in, in dependency order: the backend selector and the vocabulary (enums,
// operator/monoid/semiring definitions), the host utilities (Matrix Market reader,
// CLI flags, COO/CSR helpers), the objects (Descriptor, Vector,
// Matrix), the operations, and finally the implementation of the selected backend.
// Same contents as reference graphblas/graphblas.hpp:4-17.
#ifndef GRAPHBLAS_GRAPHBLAS_HPP_
#define GRAPHBLAS_GRAPHBLAS_HPP_

// selector + vocabulary
#include <graphblas/backend.hpp>
#include <graphblas/mmio.hpp>
#include <graphblas/types.hpp>
#include <graphblas/stddef.hpp>

// host utilities
#include <graphblas/util.hpp>

// objects and operations
#include <graphblas/descriptor.hpp>
#include <graphblas/vector.hpp>
#include <graphblas/matrix.hpp>
#include <graphblas/operations.hpp>

// the backend behind them (GRB_USE_CUDA: graphblas/backend/cuda/)
#include <graphblas/backend/cuda/cuda.hpp>

#endif  // GRAPHBLAS_GRAPHBLAS_HPP_
