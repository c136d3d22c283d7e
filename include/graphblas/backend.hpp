// graphblast_b200 frontend mirror — backend selector.
// Same contract as reference graphblas/backend.hpp:4-15: GRB_USE_CUDA selects
// graphblas/backend/cuda/ and makes functors __host__ __device__.  The
// reference's GRB_USE_SEQUENTIAL branch points at a backend that does not
// compile and is not offered here.
#ifndef GRAPHBLAS_BACKEND_HPP_
#define GRAPHBLAS_BACKEND_HPP_

#if defined(GRB_USE_CUDA)
  #define __GRB_BACKEND_ROOT cuda
  #define GRB_HOST_DEVICE __host__ __device__
#else
  #error "graphblast_b200: define GRB_USE_CUDA before including graphblas headers"
#endif

#endif  // GRAPHBLAS_BACKEND_HPP_
