// graphblast_b200 frontend mirror — scalar typedefs and the public enums.
// The enumerator VALUES are the API: they travel through the C ABI as plain ints and
// Descriptor::toggle relies on GrB_SCMP = 0, GrB_REPLACE = 1, GrB_TRAN = 2.  They equal
// reference graphblas/types.hpp:18-78 one for one.  Each enum is written once as a
// list, from which both the enum and its printable names (infoName, descValueName:
// what error messages and the Python host show) are generated.
#ifndef GRAPHBLAS_TYPES_HPP_
#define GRAPHBLAS_TYPES_HPP_

#define GrB_NULL   NULL
#define GrB_ALL    NULL
#define GrB_MEMORY false  // print memory usage info

#include <cstddef>
#include <cstdint>

#include <graphblas/backend/cuda/types.hpp>

// Return codes, in value order (0 = success; API errors; execution errors).
#define GB_INFO_LIST(X)                                                        \
  X(GrB_SUCCESS)                                                               \
  X(GrB_UNINITIALIZED_OBJECT) X(GrB_NULL_POINTER)       X(GrB_INVALID_VALUE)   \
  X(GrB_INVALID_INDEX)        X(GrB_DOMAIN_MISMATCH)    X(GrB_DIMENSION_MISMATCH) \
  X(GrB_OUTPUT_NOT_EMPTY)     X(GrB_NO_VALUE)           X(GrB_NOT_IMPLEMENTED) \
  X(GrB_OUT_OF_MEMORY)        X(GrB_INSUFFICIENT_SPACE) X(GrB_INVALID_OBJECT)  \
  X(GrB_INDEX_OUT_OF_BOUNDS)  X(GrB_PANIC)

// Descriptor values with the field they belong to.  8 sits between the mode values.
#define GB_DESC_VALUE_LIST(X)                                                  \
  X(GrB_SCMP, 0)      /* GrB_MASK */                                           \
  X(GrB_REPLACE, 1)   /* GrB_OUTP */                                           \
  X(GrB_TRAN, 2)      /* GrB_INP0, GrB_INP1 */                                 \
  X(GrB_DEFAULT, 3)                                                            \
  X(GrB_CUSPARSE, 4)  X(GrB_CUSPARSE2, 5) X(GrB_FIXEDROW, 6) X(GrB_FIXEDCOL, 7) \
  X(GrB_MERGEPATH, 9) /* GrB_MODE */                                           \
  X(GrB_PUSHPULL, 10) X(GrB_PUSHONLY, 11) X(GrB_PULLONLY, 12) /* GrB_MXVMODE */ \
  X(GrB_SEQUENTIAL, 13) X(GrB_CUDA, 14)   /* GrB_BACKEND */                    \
  X(GrB_8, 8) X(GrB_16, 16) X(GrB_32, 32) X(GrB_64, 64) X(GrB_128, 128)        \
  X(GrB_256, 256) X(GrB_512, 512) X(GrB_1024, 1024) /* GrB_TA, GrB_TB, GrB_NT, GrB_TOL */

namespace graphblas {
typedef int   Index;
typedef float T;

enum Storage { GrB_UNKNOWN, GrB_SPARSE, GrB_DENSE };
enum Major   { GrB_ROWMAJOR, GrB_COLMAJOR };

#define GB_PLAIN(name) name,
enum Info { GB_INFO_LIST(GB_PLAIN) GrB_NINFO };
#undef GB_PLAIN

enum Desc_field {
  GrB_MASK, GrB_OUTP, GrB_INP0, GrB_INP1,          // toggled by Descriptor::toggle
  GrB_MODE, GrB_TA, GrB_TB, GrB_NT, GrB_MXVMODE, GrB_TOL, GrB_BACKEND,
  GrB_NDESCFIELD
};

#define GB_VALUED(name, value) name = value,
enum Desc_value { GB_DESC_VALUE_LIST(GB_VALUED) };
#undef GB_VALUED

inline const char* infoName(int code) {
#define GB_CASE(name) case name: return #name;
  switch (code) { GB_INFO_LIST(GB_CASE) default: return "unknown Info"; }
#undef GB_CASE
}
inline const char* descValueName(int value) {
#define GB_CASE(name, v) case v: return #name;
  switch (value) { GB_DESC_VALUE_LIST(GB_CASE) default: return "unknown Desc_value"; }
#undef GB_CASE
}
}  // namespace graphblas

#endif  // GRAPHBLAS_TYPES_HPP_
