// graphblast_b200 frontend mirror — binary operators, monoids and semirings.
//
// Same names, template parameters, identities and host/device qualifiers as
// reference graphblas/stddef.hpp:14-277, because user code and the algorithm
// headers instantiate them by name (LogicalOrAndSemiring<float>() ...).  Notable
// identities kept as the reference defines them (:160-173): LogicalAndMonoid ->
// false, MaximumMonoid -> 0, GreaterMonoid -> numeric_limits::min().
// identity() is host-only in the reference; the backend passes it by value.
#ifndef GRAPHBLAS_STDDEF_HPP_
#define GRAPHBLAS_STDDEF_HPP_

#include <cstddef>
#include <cstdint>
#include <limits>
#include <algorithm>

// name, default (T_in2, T_out) types, expression in lhs/rhs
#define GB_DEFINE_BINARY_OP(NAME, DEF_IN1, DEF_IN2, DEF_OUT, EXPR)            \
template <typename T_in1 DEF_IN1, typename T_in2 DEF_IN2,                     \
          typename T_out DEF_OUT>                                             \
struct NAME {                                                                 \
  inline GRB_HOST_DEVICE T_out operator()(T_in1 lhs, T_in2 rhs) {             \
    return EXPR;                                                              \
  }                                                                           \
};

#define GB_SAME  = T_in1
#define GB_BOOL  = bool
#define GB_NONE

namespace graphblas {
GB_DEFINE_BINARY_OP(logical_or,    GB_BOOL, GB_BOOL, GB_BOOL, lhs || rhs)
GB_DEFINE_BINARY_OP(logical_and,   GB_BOOL, GB_BOOL, GB_BOOL, lhs && rhs)
GB_DEFINE_BINARY_OP(logical_xor,   GB_BOOL, GB_BOOL, GB_BOOL,
                    (lhs && !rhs) || (!lhs && rhs))
GB_DEFINE_BINARY_OP(equal,         GB_NONE, GB_SAME, GB_SAME, lhs == rhs)
GB_DEFINE_BINARY_OP(not_equal_to,  GB_NONE, GB_SAME, GB_SAME, lhs != rhs)
GB_DEFINE_BINARY_OP(greater,       GB_NONE, GB_SAME, GB_BOOL, lhs > rhs)
GB_DEFINE_BINARY_OP(less,          GB_NONE, GB_SAME, GB_BOOL, lhs < rhs)
GB_DEFINE_BINARY_OP(greater_equal, GB_NONE, GB_SAME, GB_BOOL, lhs >= rhs)
GB_DEFINE_BINARY_OP(less_equal,    GB_NONE, GB_SAME, GB_BOOL, lhs <= rhs)
GB_DEFINE_BINARY_OP(first,         GB_NONE, GB_SAME, GB_SAME, lhs)
GB_DEFINE_BINARY_OP(second,        GB_NONE, GB_SAME, GB_SAME, rhs)
GB_DEFINE_BINARY_OP(minimum,       GB_NONE, GB_SAME, GB_SAME, min(lhs, rhs))
GB_DEFINE_BINARY_OP(maximum,       GB_NONE, GB_SAME, GB_SAME, max(lhs, rhs))
GB_DEFINE_BINARY_OP(plus,          GB_NONE, GB_SAME, GB_SAME, lhs + rhs)
GB_DEFINE_BINARY_OP(minus,         GB_NONE, GB_SAME, GB_SAME, lhs - rhs)
GB_DEFINE_BINARY_OP(multiplies,    GB_NONE, GB_SAME, GB_SAME, lhs * rhs)
GB_DEFINE_BINARY_OP(divides,       GB_NONE, GB_SAME, GB_SAME, lhs / rhs)
GB_DEFINE_BINARY_OP(select_second, GB_NONE, GB_SAME, GB_SAME, rhs)
}  // namespace graphblas

#undef GB_SAME
#undef GB_BOOL
#undef GB_NONE

#define REGISTER_MONOID(M_NAME, BINARYOP, IDENTITY)                          \
template <typename T_out>                                                    \
struct M_NAME {                                                              \
  inline T_out identity() const { return static_cast<T_out>(IDENTITY); }     \
  inline __host__ __device__ T_out operator()(T_out lhs, T_out rhs) const {  \
    return BINARYOP<T_out>()(lhs, rhs);                                      \
  }                                                                          \
};

namespace graphblas {
REGISTER_MONOID(PlusMonoid,       plus,         0)
REGISTER_MONOID(MultipliesMonoid, multiplies,   1)
REGISTER_MONOID(MinimumMonoid,    minimum,      std::numeric_limits<T_out>::max())
REGISTER_MONOID(MaximumMonoid,    maximum,      0)
REGISTER_MONOID(LogicalOrMonoid,  logical_or,   false)
REGISTER_MONOID(LogicalAndMonoid, logical_and,  false)
REGISTER_MONOID(GreaterMonoid,    greater,      std::numeric_limits<T_out>::min())
// "less" is not a monoid (two-sided identities differ, not associative); the
// reference registers it anyway for the SSSP improvement test.
REGISTER_MONOID(CustomLessMonoid, less,         std::numeric_limits<T_out>::max())
REGISTER_MONOID(NotEqualToMonoid, not_equal_to, std::numeric_limits<T_out>::max())
}  // namespace graphblas

#define REGISTER_SEMIRING(SR_NAME, ADD_MONOID, MULT_BINARYOP)                \
template <typename T_in1, typename T_in2 = T_in1, typename T_out = T_in1>    \
struct SR_NAME {                                                             \
  typedef T_out result_type;                                                 \
  typedef T_out T_out_type;                                                  \
  inline T_out identity() const { return ADD_MONOID<T_out>().identity(); }   \
  inline __host__ __device__ T_out add_op(T_out lhs, T_out rhs) {            \
    return ADD_MONOID<T_out>()(lhs, rhs);                                    \
  }                                                                          \
  inline __host__ __device__ T_out mul_op(T_in1 lhs, T_in2 rhs) {            \
    return MULT_BINARYOP<T_in1, T_in2, T_out>()(lhs, rhs);                   \
  }                                                                          \
};

namespace graphblas {
REGISTER_SEMIRING(LogicalOrAndSemiring,         LogicalOrMonoid,  logical_and)
REGISTER_SEMIRING(PlusMultipliesSemiring,       PlusMonoid,       multiplies)
REGISTER_SEMIRING(MinimumPlusSemiring,          MinimumMonoid,    plus)
REGISTER_SEMIRING(MaximumMultipliesSemiring,    MaximumMonoid,    multiplies)
REGISTER_SEMIRING(PlusDividesSemiring,          PlusMonoid,       divides)
REGISTER_SEMIRING(PlusGreaterSemiring,          PlusMonoid,       greater)
REGISTER_SEMIRING(GreaterPlusSemiring,          GreaterMonoid,    plus)
REGISTER_SEMIRING(PlusMinusSemiring,            PlusMonoid,       minus)
REGISTER_SEMIRING(PlusLessSemiring,             PlusMonoid,       less)
REGISTER_SEMIRING(CustomLessPlusSemiring,       CustomLessMonoid, plus)
REGISTER_SEMIRING(MinimumMultipliesSemiring,    MinimumMonoid,    multiplies)
REGISTER_SEMIRING(MultipliesMultipliesSemiring, MultipliesMonoid, multiplies)
REGISTER_SEMIRING(NotEqualToPlusSemiring,       NotEqualToMonoid, plus)
REGISTER_SEMIRING(MinimumSelectSecondSemiring,  MinimumMonoid,    select_second)
REGISTER_SEMIRING(PlusNotEqualToSemiring,       PlusMonoid,       not_equal_to)
REGISTER_SEMIRING(CustomLessLessSemiring,       CustomLessMonoid, less)
REGISTER_SEMIRING(MinimumNotEqualToSemiring,    MinimumMonoid,    not_equal_to)

// Functor views of a semiring's two operations (what kernels are templated on).
template <typename SemiringT, bool IsAdd>
struct SemiringOpView {
  typedef typename SemiringT::T_out_type T_out_type;
  typedef typename SemiringT::T_out_type result_type;
  typedef typename SemiringT::T_out_type first_argument_type;
  typedef typename SemiringT::T_out_type second_argument_type;

  SemiringOpView() : sr() {}
  explicit SemiringOpView(SemiringT const& s) : sr(s) {}

  inline GRB_HOST_DEVICE T_out_type identity() const { return sr.identity(); }

  template <typename T_in1, typename T_in2>
  inline GRB_HOST_DEVICE T_out_type operator()(T_in1 lhs, T_in2 rhs) {
    return IsAdd ? sr.add_op(lhs, rhs) : sr.mul_op(lhs, rhs);
  }

 private:
  SemiringT sr;
};

template <typename SemiringT>
using AdditiveMonoidFromSemiring = SemiringOpView<SemiringT, true>;
template <typename SemiringT>
using MultiplicativeMonoidFromSemiring = SemiringOpView<SemiringT, false>;

template <typename SemiringT>
AdditiveMonoidFromSemiring<SemiringT> extractAdd(SemiringT const& sr) {
  return AdditiveMonoidFromSemiring<SemiringT>(sr);
}

template <typename SemiringT>
MultiplicativeMonoidFromSemiring<SemiringT> extractMul(SemiringT const& sr) {
  return MultiplicativeMonoidFromSemiring<SemiringT>(sr);
}
}  // namespace graphblas

#endif  // GRAPHBLAS_STDDEF_HPP_
