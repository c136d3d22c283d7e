// graphblast_b200 frontend mirror — scalar typedefs and the public enums.
// The enumerator VALUES are the API (Descriptor::toggle relies on
// GrB_SCMP=0, GrB_REPLACE=1, GrB_TRAN=2) and equal reference
// graphblas/types.hpp:18-78 one for one.
#ifndef GRAPHBLAS_TYPES_HPP_
#define GRAPHBLAS_TYPES_HPP_

#define GrB_NULL   NULL
#define GrB_ALL    NULL
#define GrB_MEMORY false  // print memory usage info

#include <cstddef>
#include <cstdint>

#include <graphblas/backend/cuda/types.hpp>

namespace graphblas {
typedef int   Index;
typedef float T;

enum Storage { GrB_UNKNOWN, GrB_SPARSE, GrB_DENSE };

enum Major { GrB_ROWMAJOR, GrB_COLMAJOR };

enum Info {
  GrB_SUCCESS,
  // API errors
  GrB_UNINITIALIZED_OBJECT, GrB_NULL_POINTER, GrB_INVALID_VALUE,
  GrB_INVALID_INDEX, GrB_DOMAIN_MISMATCH, GrB_DIMENSION_MISMATCH,
  GrB_OUTPUT_NOT_EMPTY, GrB_NO_VALUE, GrB_NOT_IMPLEMENTED,
  // Execution errors
  GrB_OUT_OF_MEMORY, GrB_INSUFFICIENT_SPACE, GrB_INVALID_OBJECT,
  GrB_INDEX_OUT_OF_BOUNDS, GrB_PANIC
};

enum Desc_field {
  GrB_MASK, GrB_OUTP, GrB_INP0, GrB_INP1, GrB_MODE, GrB_TA, GrB_TB, GrB_NT,
  GrB_MXVMODE, GrB_TOL, GrB_BACKEND, GrB_NDESCFIELD
};

enum Desc_value {
  GrB_SCMP       =    0,  // GrB_MASK
  GrB_REPLACE    =    1,  // GrB_OUTP
  GrB_TRAN       =    2,  // GrB_INP0, GrB_INP1
  GrB_DEFAULT    =    3,
  GrB_CUSPARSE   =    4,  // GrB_MODE
  GrB_CUSPARSE2  =    5,
  GrB_FIXEDROW   =    6,
  GrB_FIXEDCOL   =    7,
  GrB_MERGEPATH  =    9,
  GrB_PUSHPULL   =   10,  // GrB_MXVMODE
  GrB_PUSHONLY   =   11,
  GrB_PULLONLY   =   12,
  GrB_SEQUENTIAL =   13,  // GrB_BACKEND
  GrB_CUDA       =   14,
  GrB_8          =    8,  // GrB_TA, GrB_TB, GrB_NT, GrB_TOL
  GrB_16         =   16,
  GrB_32         =   32,
  GrB_64         =   64,
  GrB_128        =  128,
  GrB_256        =  256,
  GrB_512        =  512,
  GrB_1024       = 1024
};
}  // namespace graphblas

#endif  // GRAPHBLAS_TYPES_HPP_
