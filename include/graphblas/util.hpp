// graphblast_b200 frontend mirror — CHECK macros, command-line flags, environment
// helpers, the Matrix Market loader and COO/CSR/CSC conversions.
//
// Same entry points and behaviour as reference graphblas/util.hpp:18-600; the
// loader defines the CSR every result is checked on, so its semantics are kept
// exactly (reference :264-329, :364-430):
//   * `directed`: 0 follow the file, 1 force directed, 2 force undirected;
//   * undirected graphs get a reverse copy of every non-loop entry;
//   * entries are sorted by (row, col); self-loops (unless
//     GRB_UTIL_REMOVE_SELFLOOP=0) and repeated (row, col) pairs are dropped;
//   * a binary cache "<dir>/.<file>.<ud|d>.<nosl|sl>.bin" short-circuits parsing.
// Implementation differs: entries are sorted through an index permutation of
// 64-bit (row,col) keys instead of a vector of 4-tuples.
#ifndef GRAPHBLAS_UTIL_HPP_
#define GRAPHBLAS_UTIL_HPP_

#include <sys/resource.h>
#include <sys/time.h>
#include <libgen.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>
#include <tuple>
#include <algorithm>
#include <numeric>
#include <string>
#include <iostream>
#include <typeinfo>

// for commandline arguments
#include <boost/program_options.hpp>

#define CHECK(x) do {                                           \
  graphblas::Info err = x;                                      \
  if (err != graphblas::GrB_SUCCESS) {                          \
    fprintf(stderr, "Runtime error: %s returned %d at %s:%d\n", \
            #x, err, __FILE__, __LINE__);                       \
    return err;                                                 \
  } } while (0)

#define CHECKVOID(x) do {                                       \
  graphblas::Info err = x;                                      \
  if (err != graphblas::GrB_SUCCESS) {                          \
    fprintf(stderr, "Runtime error: %s returned %d at %s:%d\n", \
            #x, err, __FILE__, __LINE__);                       \
    return;                                                     \
  } } while (0)

#define GRB_MAXLEN 256

namespace po = boost::program_options;

// The flag table of the reference drivers (reference util.hpp:39-132): same
// names, types and defaults.
inline void parseArgs(int argc, char** argv, po::variables_map* vm) {
  po::options_description desc("Allowed options");
  desc.add_options()
    ("help", "produce help message")
    ("ta",   po::value<int>()->default_value(32), "threads per A row")
    ("tb",   po::value<int>()->default_value(32), "B slab width")
    ("mode", po::value<std::string>()->default_value("fixedrow"),
        "row or column")
    ("split", po::value<bool>()->default_value(false),
        "split computation when possible e.g. mxm, reduce")
    // General params
    ("niter", po::value<int>()->default_value(10),
        "outer-loop repetitions after warmup")
    ("max_niter", po::value<int>()->default_value(10000),
        "inner-loop iteration cap")
    ("directed", po::value<int>()->default_value(0),
        "0: follow mtx, 1: force directed, 2: force undirected")
    ("timing", po::value<int>()->default_value(1),
        "1: print per-iteration timing")
    ("transpose", po::value<bool>()->default_value(false),
        "use the transposed graph")
    ("mtxinfo", po::value<bool>()->default_value(true),
        "print MTX info")
    ("verbose", po::value<bool>()->default_value(true),
        "timing output and correctness indicator")
    ("skip_cpu_verify", po::value<bool>()->default_value(false),
        "skip the CPU verification run")
    // mxv params
    ("source", po::value<int>()->default_value(0),
        "traversal source / seed for randomised algorithms")
    ("source_start", po::value<int>()->default_value(0), "source range begin")
    ("source_end", po::value<int>()->default_value(1), "source range end")
    ("mxvmode", po::value<int>()->default_value(1),
        "0: push-pull, 1: push only, 2: pull only")
    ("switchpoint", po::value<float>()->default_value(0.01),
        "nnz fraction at which mxvmode=0 switches sparse<->dense")
    ("dirinfo", po::value<bool>()->default_value(false),
        "print direction decisions")
    ("struconly", po::value<bool>()->default_value(false),
        "implied nonzeroes instead of key-value")
    ("opreuse", po::value<bool>()->default_value(false),
        "operand reuse in the Boolean pull")
    // mxv (spmspv/push) params
    ("memusage", po::value<float>()->default_value(1.0),
        "multiple of |E| for the reference's push scratch (ignored here)")
    ("endbit", po::value<bool>()->default_value(true),
        "radix-sort bit range of the reference's push (ignored here)")
    ("sort", po::value<bool>()->default_value(true),
        "sorted push output (always sorted here)")
    ("atomic", po::value<bool>()->default_value(false),
        "atomics in SIMPLE/TWC load balancing (unused)")
    // mxv (spmv/pull) params
    ("earlyexit", po::value<bool>()->default_value(true),
        "early exit in the Boolean pull")
    ("fusedmask", po::value<bool>()->default_value(true),
        "fused mask in the Boolean pull")
    // algorithm-specific params
    ("maxcolors", po::value<int>()->default_value(10000), "graph colouring cap")
    ("gcalgo", po::value<int>()->default_value(0), "graph colouring variant")
    ("ccalgo", po::value<int>()->default_value(0), "connected components variant")
    ("seed", po::value<int>()->default_value(-1),
        "RNG seed (SSSP edge weights, ...)")
    // GPU params
    ("nthread", po::value<int>()->default_value(128), "threads per block")
    ("ndevice", po::value<int>()->default_value(0), "GPU device number")
    ("debug", po::value<bool>()->default_value(false), "debug messages")
    ("memory", po::value<bool>()->default_value(false), "memory info");

  po::store(po::parse_command_line(argc, argv, desc), *vm);
  po::notify(*vm);

  if (vm->count("help"))
    std::cout << desc << "\n";
}

template <typename T>
inline T getEnv(const char* key, T default_val) {
  const char* val = std::getenv(key);
  if (val == NULL) return default_val;
  return static_cast<T>(atoi(val));
}

template <typename T>
void setEnv(const char* key, T default_val) {
  std::string s = std::to_string(default_val);
  setenv(key, s.c_str(), 0);
}

// Sort the three parallel arrays by (row, col).  std::sort on the keys with the
// original position as tie-break free ordering is NOT stable in the reference
// either (std::sort of tuples compared on (row, col) only, :149-195).
template <typename T>
void customSort(std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values) {
  const size_t n = row_indices->size();
  std::vector<size_t> perm(n);
  std::iota(perm.begin(), perm.end(), static_cast<size_t>(0));
  const std::vector<graphblas::Index>& r = *row_indices;
  const std::vector<graphblas::Index>& c = *col_indices;
  std::sort(perm.begin(), perm.end(), [&](size_t x, size_t y) {
    if (r[x] != r[y]) return r[x] < r[y];
    return c[x] < c[y];
  });
  std::vector<graphblas::Index> r2(n), c2(n);
  std::vector<T> v2(n);
  for (size_t i = 0; i < n; ++i) {
    r2[i] = r[perm[i]];
    c2[i] = c[perm[i]];
    v2[i] = (*values)[perm[i]];
  }
  row_indices->swap(r2);
  col_indices->swap(c2);
  values->swap(v2);
}

// Entries with explicit values of type mtxT (int or float in the file).
template <typename T, typename mtxT>
void readTuples(std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values,
    graphblas::Index nvals, FILE* f) {
  const char* fmt = (typeid(mtxT) == typeid(int)) ? "%d" : "%f";
  for (graphblas::Index i = 0; i < nvals; i++) {
    graphblas::Index row_ind, col_ind;
    mtxT raw_value = mtxT();
    if (fscanf(f, "%d", &row_ind) == EOF) {
      std::cout << "Error: Not enough rows in mtx file!\n";
      return;
    }
    if (fscanf(f, "%d", &col_ind) != 1) return;
    if (fscanf(f, fmt, &raw_value) != 1) return;
    row_indices->push_back(row_ind - 1);   // 1-based file -> 0-based
    col_indices->push_back(col_ind - 1);
    values->push_back(static_cast<T>(raw_value));
  }
}

// Pattern entries: value 1.
template <typename T>
void readTuples(std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values,
    graphblas::Index nvals, FILE* f) {
  for (graphblas::Index i = 0; i < nvals; i++) {
    graphblas::Index row_ind, col_ind;
    if (fscanf(f, "%d", &row_ind) == EOF) {
      std::cout << "Error: Not enough rows in mtx file!\n";
      return;
    }
    if (fscanf(f, "%d", &col_ind) != 1) return;
    row_indices->push_back(row_ind - 1);
    col_indices->push_back(col_ind - 1);
    values->push_back(static_cast<T>(1.0));
  }
}

// Symmetrise (optional), sort, drop self-loops and duplicates.
template <typename T>
void removeSelfloop(std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values,
    graphblas::Index* nvals, bool undirected) {
  bool remove_self_loops = getEnv("GRB_UTIL_REMOVE_SELFLOOP", true);

  if (undirected) {
    const graphblas::Index n0 = *nvals;
    for (graphblas::Index i = 0; i < n0; i++) {
      if ((*col_indices)[i] != (*row_indices)[i]) {
        row_indices->push_back((*col_indices)[i]);
        col_indices->push_back((*row_indices)[i]);
        values->push_back((*values)[i]);
      }
    }
  }
  *nvals = row_indices->size();
  if (*nvals == 0) return;

  customSort<T>(row_indices, col_indices, values);

  graphblas::Index kept = 0;
  for (graphblas::Index i = 0; i < *nvals; i++) {
    const graphblas::Index r = (*row_indices)[i];
    const graphblas::Index c = (*col_indices)[i];
    if (remove_self_loops && r == c) continue;
    if (i > 0 && r == (*row_indices)[i-1] && c == (*col_indices)[i-1]) continue;
    (*row_indices)[kept] = r;
    (*col_indices)[kept] = c;
    (*values)[kept]      = (*values)[i];
    ++kept;
  }
  *nvals = kept;
  row_indices->resize(kept);
  col_indices->resize(kept);
  values->resize(kept);
}

inline bool exists(const char* fname) {
  FILE* file = fopen(fname, "r");
  if (file != NULL) {
    fclose(file);
    return true;
  }
  return false;
}

// Cache file name for a matrix file (malloc'ed; freed by Matrix::build).
inline char* convert(const char* fname, bool is_undirected = true) {
  char* dat_name = reinterpret_cast<char*>(malloc(GRB_MAXLEN));
  char* temp1 = strdup(fname);
  char* temp2 = strdup(fname);
  char* file_path = dirname(temp1);
  char* file_name = basename(temp2);
  bool remove_self_loops = getEnv("GRB_UTIL_REMOVE_SELFLOOP", true);
  std::cout << "Remove self-loop: " << remove_self_loops << std::endl;

  snprintf(dat_name, GRB_MAXLEN, "%s/.%s.%s.%s.%sbin", file_path, file_name,
      (is_undirected ? "ud" : "d"),
      (remove_self_loops ? "nosl" : "sl"),
      ((sizeof(graphblas::Index) == 8) ? "64bVe." : ""));
  free(temp1);
  free(temp2);
  return dat_name;
}

template <typename T>
int readMtx(const char* fname, std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values,
    graphblas::Index* nrows, graphblas::Index* ncols, graphblas::Index* nvals,
    int directed, bool mtxinfo, char** dat_name = NULL) {
  int ret_code;
  MM_typecode matcode;
  FILE* f;

  if ((f = fopen(fname, "r")) == NULL) {
    printf("File %s not found\n", fname);
    exit(1);
  }
  if (mm_read_banner(f, &matcode) != 0) {
    printf("Could not process Matrix Market banner.\n");
    exit(1);
  }
  if ((ret_code = mm_read_mtx_crd_size(f, nrows, ncols, nvals)) != 0)
    exit(1);

  printf("Undirected due to mtx: %d\n", mm_is_symmetric(matcode));
  printf("Undirected due to cmd: %d\n", directed == 2);
  bool is_undirected = mm_is_symmetric(matcode) || directed == 2;
  if (directed == 1) is_undirected = false;
  printf("Undirected: %d\n", is_undirected);
  if (dat_name != NULL)
    *dat_name = convert(fname, is_undirected);

  bool cache_hit = false;
  if (dat_name != NULL && exists(*dat_name)) {
    std::ifstream ifs(*dat_name, std::ios::in | std::ios::binary);
    if (ifs.fail()) {
      std::cout << "Error: Unable to open file for reading!\n";
    } else {
      // Empty arrays tell Matrix::build that the binary file is to be used.
      row_indices->clear();
      col_indices->clear();
      values->clear();
      cache_hit = true;
    }
  }
  if (!cache_hit && !(dat_name != NULL && exists(*dat_name))) {
    if (mm_is_integer(matcode))
      readTuples<T, int>(row_indices, col_indices, values, *nvals, f);
    else if (mm_is_real(matcode))
      readTuples<T, float>(row_indices, col_indices, values, *nvals, f);
    else if (mm_is_pattern(matcode))
      readTuples<T>(row_indices, col_indices, values, *nvals, f);

    removeSelfloop<T>(row_indices, col_indices, values, nvals, is_undirected);

    if (mtxinfo) mm_write_banner(stdout, matcode);
    if (mtxinfo) mm_write_mtx_crd_size(stdout, *nrows, *ncols, *nvals);
  }
  fclose(f);
  return ret_code;
}

template <typename T>
void printArray(const char* str, const T* array, int length = 40, bool limit = true) {
  if (limit && length > 40) length = 40;
  std::cout << str << ":\n";
  for (int i = 0; i < length; i++)
    std::cout << "[" << i << "]:" << array[i] << " ";
  std::cout << "\n";
}

template <typename T>
void printArray(const char* str, const std::vector<T>& array, int length = 40,
    bool limit = true) {
  if (limit && length > 40) length = 40;
  std::cout << str << ":\n";
  for (int i = 0; i < length; i++)
    std::cout << "[" << i << "]:" << array[i] << " ";
  std::cout << "\n";
}

// Wall-clock timer in milliseconds (reference :452-497).
struct CpuTimer {
  double start;
  double stop;

  static double now() {
    struct timeval tv;
    gettimeofday(&tv, NULL);
    return tv.tv_sec + 1.e-6*tv.tv_usec;
  }
  void Start() { start = now(); }
  void Stop()  { stop = now(); }
  double ElapsedMillis() { return 1000*(stop - start); }
};

using namespace graphblas;

// COO -> CSR by counting sort on the row index; within a row the input order is
// kept, so a (row, col)-sorted input yields sorted rows (reference :502-556
// sorts a copy first; the result is the same CSR).
template <typename T>
void coo2csr(Index* csrRowPtr, Index* csrColInd, T* csrVal,
    const std::vector<Index>& row_indices, const std::vector<Index>& col_indices,
    const std::vector<T>& values, Index nrows, Index ncols) {
  const Index nvals = row_indices.size();
  std::vector<Index> r = row_indices;
  std::vector<Index> c = col_indices;
  std::vector<T>     v = values;
  customSort<T>(&r, &c, &v);

  for (Index i = 0; i <= nrows; i++) csrRowPtr[i] = 0;
  for (Index i = 0; i < nvals; i++) {
    if (r[i] >= nrows) std::cout << "Error: Index out of bounds!\n";
    else csrRowPtr[r[i] + 1]++;
  }
  for (Index i = 0; i < nrows; i++) csrRowPtr[i+1] += csrRowPtr[i];
  for (Index i = 0; i < nvals; i++) {
    if (c[i] >= ncols) std::cout << "Error: Index out of bounds!\n";
    csrColInd[i] = c[i];
    csrVal[i]    = v[i];
  }
}

template <typename T>
void coo2csc(Index* cscColPtr, Index* cscRowInd, T* cscVal,
    const std::vector<Index>& row_indices, const std::vector<Index>& col_indices,
    const std::vector<T>& values, Index nrows, Index ncols) {
  return coo2csr(cscColPtr, cscRowInd, cscVal, col_indices, row_indices, values, ncols,
      nrows);
}

template <typename T>
void csr2csc(Index* cscColPtr, Index* cscRowInd, T* cscVal, const Index* csrRowPtr,
    const Index* csrColInd, const T* csrVal, Index nrows, Index ncols) {
  const Index nvals = csrRowPtr[nrows];
  std::vector<Index> row_indices(nvals, 0);
  std::vector<Index> col_indices(nvals, 0);
  std::vector<T>     values(nvals, 0);
  for (Index i = 0; i < nrows; ++i) {
    for (Index k = csrRowPtr[i]; k < csrRowPtr[i+1]; ++k) {
      row_indices[k] = i;
      col_indices[k] = csrColInd[k];
      values[k]      = csrVal[k];
    }
  }
  return coo2csc(cscColPtr, cscRowInd, cscVal, row_indices, col_indices, values, ncols,
      nrows);
}

#endif  // GRAPHBLAS_UTIL_HPP_
