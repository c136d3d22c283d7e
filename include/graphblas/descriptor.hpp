// graphblast_b200 frontend mirror — graphblas::Descriptor.
// Public surface of reference graphblas/descriptor.hpp:17-62: set / get / toggle /
// loadArgs forwarding to backend::Descriptor held by value as `descriptor_`
// (algorithms read desc->descriptor_.max_niter_ etc., reference
// algorithm/bfs.hpp:46).
#ifndef GRAPHBLAS_DESCRIPTOR_HPP_
#define GRAPHBLAS_DESCRIPTOR_HPP_

#include <vector>

#include "graphblas/types.hpp"
#include <graphblas/backend/cuda/descriptor.hpp>

namespace graphblas {
template <typename T>
class Matrix;

class Descriptor {
 public:
  Descriptor() : descriptor_() {}
  ~Descriptor() {}

  Info set(Desc_field field, Desc_value value) {
    return descriptor_.set(field, value);
  }
  Info set(Desc_field field, int value) {
    return descriptor_.set(field, static_cast<Desc_value>(value));
  }
  Info get(Desc_field field, Desc_value* value) const {
    if (value == NULL) return GrB_NULL_POINTER;
    return descriptor_.get(field, value);
  }
  Info toggle(Desc_field field) { return descriptor_.toggle(field); }
  Info loadArgs(const po::variables_map& vm) {
    return descriptor_.loadArgs(vm);
  }

  backend::Descriptor descriptor_;
};
}  // namespace graphblas

#endif  // GRAPHBLAS_DESCRIPTOR_HPP_
