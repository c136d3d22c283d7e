// graphblast_b200 — helpers shared by the algorithm drivers.
// set_uniform_random mirrors reference graphblas/algorithm/common.hpp:22-42:
// a stateful unary functor drawing uniform_int[start,end] from
// std::default_random_engine(GRB_SEED); applied in CSR order it produces the
// SSSP edge weights of reference example/gsssp.cu:79-84.
#ifndef GRAPHBLAS_ALGORITHM_COMMON_HPP_
#define GRAPHBLAS_ALGORITHM_COMMON_HPP_

#include <random>

namespace graphblas {

template <typename T_in1, typename T_out = T_in1>
struct set_uniform_random {
  set_uniform_random()
      : seed_(getEnv("GRB_SEED", 0)),
        start_(getEnv("GRB_UNIFORM_START", 0)),
        end_(getEnv("GRB_UNIFORM_END", 1)),
        gen_(seed_), dist_(start_, end_) {}

  set_uniform_random(int seed, int start, int end)
      : seed_(seed), start_(start), end_(end), gen_(seed), dist_(start, end) {}

  inline T_out operator()(T_in1 lhs) {
    return static_cast<T_out>(dist_(gen_));
  }

  int seed_;
  int start_;
  int end_;
  std::default_random_engine gen_;
  std::uniform_int_distribution<int> dist_;
};

namespace algorithm {
// The drivers return a time in milliseconds, as the reference's do
// (algorithm/bfs.hpp:18-21), so an Info cannot travel in the return value: a
// failing step records its status here and the driver returns -1.  Callers
// that want the code (the C ABI) read lastStatus() after the call.
inline Info& lastStatus() {
  static thread_local Info status = GrB_SUCCESS;
  return status;
}
#define GB_ALGO_STEP(x)                                               \
  do {                                                                \
    graphblas::Info gb_step__ = (x);                                  \
    if (gb_step__ != graphblas::GrB_SUCCESS) {                        \
      fprintf(stderr, "Runtime error: %s returned %s (%d) at %s:%d\n",     \
              #x, graphblas::infoName(gb_step__),                     \
              static_cast<int>(gb_step__), __FILE__, __LINE__);       \
      graphblas::algorithm::lastStatus() = gb_step__;                 \
      return -1.f;                                                    \
    }                                                                 \
  } while (0)

// Timer policy shared by the drivers: per-iteration lines need an event
// synchronisation per level (reference algorithm/bfs.hpp:51-63); with
// --timing 0 one event pair brackets the whole loop instead.
struct LoopTimer {
  backend::GpuTimer timer;
  float total_ms;
  bool  per_iteration;
  explicit LoopTimer(bool per_iter) : total_ms(0.f), per_iteration(per_iter) {}
  void begin() { timer.Start(); }
  // End of one iteration; returns its time when measured per iteration.
  float lap() {
    if (!per_iteration) return 0.f;
    timer.Stop();
    float ms = timer.ElapsedMillis();
    total_ms += ms;
    timer.Start();
    return ms;
  }
  float finish() {
    timer.Stop();
    total_ms += timer.ElapsedMillis();
    return total_ms;
  }
};
}  // namespace algorithm
}  // namespace graphblas

#endif  // GRAPHBLAS_ALGORITHM_COMMON_HPP_
