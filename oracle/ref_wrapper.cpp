// oracle/_ref/libgbref.so — the reference's OWN CPU implementations of the path,
// compiled from the sources where they lie under /root/reference (never copied):
//   graphblas/algorithm/test_bfs.hpp:11-61   SimpleReferenceBfs
//   graphblas/algorithm/test_sssp.hpp:15-79  SimpleReferenceSssp
//   graphblas/algorithm/test_pr.hpp:15-80    SimpleReferencePr
//   graphblas/algorithm/test_tc.hpp:41-85    SimpleReferenceTc
//   graphblas/util.hpp:364-430, 502-556      readMtx (+removeSelfloop, customSort), coo2csr
//   graphblas/algorithm/common.hpp:22-42     set_uniform_random (SSSP weights)
// These four functions are what every reference driver runs and compares the
// GPU result against (example/gbfs.cu:82,93 ...), i.e. the reference's
// "CPU sequential backend".  This file only adds extern "C" entry points.
//
// TEST INFRASTRUCTURE ONLY: used by tests/ to pin oracle/gb_oracle.c and by
// bench.py's cpu_baseline / --impl reference leg.  Nothing in graphblast_b200/
// links or loads it.
#define GRB_USE_CUDA
#define __host__
#define __device__

#include <fcntl.h>
#include <unistd.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <typeinfo>
#include <vector>

#include <boost/program_options.hpp>

#include "graphblas/backend.hpp"
#include "graphblas/mmio.hpp"
#include "graphblas/types.hpp"
#include "graphblas/util.hpp"
#include "graphblas/algorithm/test_bfs.hpp"
#include "graphblas/algorithm/test_sssp.hpp"
#include "graphblas/algorithm/test_pr.hpp"
#include "graphblas/algorithm/test_tc.hpp"
#include "graphblas/algorithm/common.hpp"

namespace {
// The reference functions print progress lines and array dumps to stdout;
// keep the caller's stdout clean (bench.py prints exactly one JSON line).
struct QuietStdout {
  int saved;
  QuietStdout() {
    fflush(stdout);
    std::cout.flush();
    saved = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    close(nul);
  }
  ~QuietStdout() {
    fflush(stdout);
    std::cout.flush();
    dup2(saved, 1);
    close(saved);
  }
};
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

int ref_bfs(int nrows, const int* rowptr, const int* colind, int* levels,
            int src, int stop) {
  QuietStdout q;
  return graphblas::algorithm::SimpleReferenceBfs<int>(nrows, rowptr, colind,
      levels, NULL, src, stop);
}

int ref_sssp(int nrows, const int* rowptr, const int* colind, float* val,
             float* dist, int src, int stop) {
  QuietStdout q;
  return graphblas::algorithm::SimpleReferenceSssp<float>(nrows, rowptr, colind,
      val, dist, src, stop);
}

int ref_pr(int nrows, const int* rowptr, const int* colind, float* val,
           float* pr, float alpha, float eps, int max_niter) {
  QuietStdout q;
  return graphblas::algorithm::SimpleReferencePr<float>(nrows, rowptr, colind,
      val, pr, alpha, eps, max_niter);
}

int ref_tc(int nrows, const int* rowptr, const int* colind, int* ntris) {
  QuietStdout q;
  return graphblas::algorithm::SimpleReferenceTc<int>(nrows, rowptr, colind,
      ntris);
}

// Reference loader: readMtx (symmetrise per `directed`, drop self-loops and
// duplicates, sort) followed by coo2csr.  Returns nvals; *nrows_out = nrows.
// Call once with rowptr == NULL to get the sizes.
int ref_load_mtx(const char* path, int directed, int* nrows_out, int* rowptr,
                 int* colind, float* val) {
  QuietStdout q;
  std::vector<graphblas::Index> rows, cols;
  std::vector<float> vals;
  graphblas::Index nrows, ncols, nvals;
  readMtx(path, &rows, &cols, &vals, &nrows, &ncols, &nvals, directed, false);
  *nrows_out = nrows;
  if (rowptr != NULL)
    coo2csr(rowptr, colind, val, rows, cols, vals, nrows, ncols);
  return nvals;
}

// The SSSP edge-weight stream of reference example/gsssp.cu:75-84: the
// set_uniform_random functor (graphblas/algorithm/common.hpp:22-42) configured
// through GRB_SEED / GRB_UNIFORM_START / GRB_UNIFORM_END, applied n times.
int ref_uniform_weights(int seed, int lo, int hi, long long n, float* out) {
  setenv("GRB_SEED", std::to_string(seed).c_str(), 1);
  setenv("GRB_UNIFORM_START", std::to_string(lo).c_str(), 1);
  setenv("GRB_UNIFORM_END", std::to_string(hi).c_str(), 1);
  graphblas::set_uniform_random<float> draw;
  for (long long i = 0; i < n; ++i) out[i] = draw(0.f);
  return 0;
}

}  // extern "C"
#pragma GCC visibility pop
