// graphblast_b200 frontend mirror — error macros, command-line flags, environment
// helpers, timers, and the Matrix Market loader.
//
// The entry points keep the names and call shapes of reference graphblas/util.hpp
// (CHECK, parseArgs, getEnv, readMtx, convert, printArray, CpuTimer, coo2csr, ...)
// because algorithm headers and drivers written against the reference call them;
// the bodies are this project's own:
//   * the flag table is data (kFlags) walked by one loop;
//   * the loader reads the file in one gulp and tokenises it by hand (the
//     reference calls fscanf per number), keeps tuples as packed 64-bit keys and
//     orders them with one std::sort over (key, position) pairs;
//   * the loader's observable semantics are the reference's, since the CSR it
//     produces is what every result is checked on (reference :264-329, :364-430):
//       directed = 0 follow the file, 1 force directed, 2 force undirected;
//       undirected graphs get a reverse copy of every non-loop entry, appended
//       after all entries of the file;
//       entries are ordered by (row, col); self-loops (unless
//       GRB_UTIL_REMOVE_SELFLOOP=0) and repeated (row, col) pairs are dropped;
//       a binary cache "<dir>/.<file>.<ud|d>.<nosl|sl>.bin" short-circuits parsing.
// The C ABI does not go through readMtx: it parses with MtxFile and hands the raw
// tuples to the device ingest (backend/cuda/ingest.hpp).
#ifndef GRAPHBLAS_UTIL_HPP_
#define GRAPHBLAS_UTIL_HPP_

#include <libgen.h>
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <utility>
#include <vector>

// for commandline arguments
#include <boost/program_options.hpp>

#define GB_CHECK_BODY(x, on_error)                              \
  do {                                                          \
    graphblas::Info gb_status__ = (x);                          \
    if (gb_status__ != graphblas::GrB_SUCCESS) {                \
      fprintf(stderr, "Runtime error: %s returned %s (%d) at %s:%d\n", \
              #x, graphblas::infoName(gb_status__),             \
              static_cast<int>(gb_status__), __FILE__, __LINE__); \
      on_error;                                                 \
    }                                                           \
  } while (0)

#define CHECK(x)     GB_CHECK_BODY(x, return gb_status__)
#define CHECKVOID(x) GB_CHECK_BODY(x, return)

#define GRB_MAXLEN 256

namespace po = boost::program_options;

// ---------------------------------------------------------------------------
// Command-line flags of the drivers: name, kind (i int, f float, b bool,
// s string), default, help.  Names, kinds and defaults are the reference's
// (util.hpp:39-132) — scripts such as run_bfs.sh pass them.
// ---------------------------------------------------------------------------
struct GbFlag {
  const char* name;
  char        kind;
  double      number;
  const char* text;
  const char* help;
};

static const GbFlag kFlags[] = {
  {"ta",              'i', 32,    NULL,       "threads per A row"},
  {"tb",              'i', 32,    NULL,       "B slab width"},
  {"mode",            's', 0,     "fixedrow", "row or column"},
  {"split",           'b', 0,     NULL,       "split computation when possible"},
  {"niter",           'i', 10,    NULL,       "outer-loop repetitions after warmup"},
  {"max_niter",       'i', 10000, NULL,       "inner-loop iteration cap"},
  {"directed",        'i', 0,     NULL,       "0 follow mtx, 1 force directed, 2 force undirected"},
  {"timing",          'i', 1,     NULL,       "1 print per-iteration timing"},
  {"transpose",       'b', 0,     NULL,       "use the transposed graph"},
  {"mtxinfo",         'b', 1,     NULL,       "print MTX info"},
  {"verbose",         'b', 1,     NULL,       "timing output and correctness indicator"},
  {"skip_cpu_verify", 'b', 0,     NULL,       "skip the CPU verification run"},
  {"source",          'i', 0,     NULL,       "traversal source / seed vertex"},
  {"source_start",    'i', 0,     NULL,       "source range begin"},
  {"source_end",      'i', 1,     NULL,       "source range end"},
  {"mxvmode",         'i', 1,     NULL,       "0 push-pull, 1 push only, 2 pull only"},
  {"switchpoint",     'f', 0.01,  NULL,       "nnz fraction at which mxvmode 0 switches"},
  {"dirinfo",         'b', 0,     NULL,       "print direction decisions"},
  {"struconly",       'b', 0,     NULL,       "implied nonzeroes instead of key-value"},
  {"opreuse",         'b', 0,     NULL,       "operand reuse in the Boolean pull"},
  {"memusage",        'f', 1.0,   NULL,       "push scratch multiple of |E| (unused here)"},
  {"endbit",          'b', 1,     NULL,       "radix-sort bit range of the push (unused here)"},
  {"sort",            'b', 1,     NULL,       "sorted push output (always sorted here)"},
  {"atomic",          'b', 0,     NULL,       "atomics in SIMPLE/TWC load balancing (unused)"},
  {"earlyexit",       'b', 1,     NULL,       "early exit in the Boolean pull"},
  {"fusedmask",       'b', 1,     NULL,       "fused mask in the Boolean pull"},
  {"maxcolors",       'i', 10000, NULL,       "graph colouring cap"},
  {"gcalgo",          'i', 0,     NULL,       "graph colouring variant"},
  {"ccalgo",          'i', 0,     NULL,       "connected components variant"},
  {"seed",            'i', -1,    NULL,       "RNG seed (SSSP edge weights, ...)"},
  {"nthread",         'i', 128,   NULL,       "threads per block"},
  {"ndevice",         'i', 0,     NULL,       "GPU device number"},
  {"debug",           'b', 0,     NULL,       "debug messages"},
  {"memory",          'b', 0,     NULL,       "memory info"},
};

inline void parseArgs(int argc, char** argv, po::variables_map* vm) {
  po::options_description desc("Allowed options");
  desc.add_options()("help", "produce help message");
  for (size_t k = 0; k < sizeof(kFlags)/sizeof(kFlags[0]); ++k) {
    const GbFlag& f = kFlags[k];
    if (f.kind == 'i')
      desc.add_options()(f.name, po::value<int>()->default_value(static_cast<int>(f.number)), f.help);
    else if (f.kind == 'f')
      desc.add_options()(f.name, po::value<float>()->default_value(static_cast<float>(f.number)), f.help);
    else if (f.kind == 'b')
      desc.add_options()(f.name, po::value<bool>()->default_value(f.number != 0), f.help);
    else
      desc.add_options()(f.name, po::value<std::string>()->default_value(f.text), f.help);
  }
  po::store(po::parse_command_line(argc, argv, desc), *vm);
  po::notify(*vm);
  if (vm->count("help")) std::cout << desc << "\n";
}

template <typename T>
inline T getEnv(const char* key, T fallback) {
  const char* text = std::getenv(key);
  return text == NULL ? fallback : static_cast<T>(atoi(text));
}

template <typename T>
void setEnv(const char* key, T value) {
  setenv(key, std::to_string(value).c_str(), 0);     // does not overwrite
}

inline bool exists(const char* fname) {
  FILE* f = fopen(fname, "r");
  if (f == NULL) return false;
  fclose(f);
  return true;
}

template <typename T>
void printArray(const char* str, const T* array, int length = 40, bool limit = true) {
  const int shown = (limit && length > 40) ? 40 : length;
  std::cout << str << ":\n";
  for (int i = 0; i < shown; ++i) std::cout << "[" << i << "]:" << array[i] << " ";
  std::cout << "\n";
}

template <typename T>
void printArray(const char* str, const std::vector<T>& array, int length = 40,
    bool limit = true) {
  printArray(str, array.data(), length, limit);
}

// Wall-clock stopwatch in milliseconds (interface of reference :452-497).
struct CpuTimer {
  std::chrono::steady_clock::time_point begin_, end_;
  void Start() { begin_ = std::chrono::steady_clock::now(); }
  void Stop()  { end_ = std::chrono::steady_clock::now(); }
  double ElapsedMillis() {
    return std::chrono::duration<double, std::milli>(end_ - begin_).count();
  }
};

// ---------------------------------------------------------------------------
// Ordering of tuples.
// ---------------------------------------------------------------------------
namespace gbutil {

inline unsigned long long packKey(graphblas::Index row, graphblas::Index col) {
  return (static_cast<unsigned long long>(static_cast<unsigned int>(row)) << 32) |
         static_cast<unsigned int>(col);
}

// Permutation that orders the tuples by (row, col), ties by position (stable).
inline std::vector<unsigned int> orderByRowCol(
    const std::vector<graphblas::Index>& rows,
    const std::vector<graphblas::Index>& cols) {
  const size_t n = rows.size();
  std::vector<std::pair<unsigned long long, unsigned int> > keyed(n);
  for (size_t i = 0; i < n; ++i)
    keyed[i] = std::make_pair(packKey(rows[i], cols[i]), static_cast<unsigned int>(i));
  std::sort(keyed.begin(), keyed.end());
  std::vector<unsigned int> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = keyed[i].second;
  return order;
}

template <typename T>
void applyOrder(const std::vector<unsigned int>& order, std::vector<T>* data) {
  std::vector<T> moved(order.size());
  for (size_t i = 0; i < order.size(); ++i) moved[i] = (*data)[order[i]];
  data->swap(moved);
}

}  // namespace gbutil

// Sort three parallel arrays by (row, col).
template <typename T>
void customSort(std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values) {
  const std::vector<unsigned int> order = gbutil::orderByRowCol(*row_indices, *col_indices);
  gbutil::applyOrder(order, row_indices);
  gbutil::applyOrder(order, col_indices);
  gbutil::applyOrder(order, values);
}

// Symmetrise (optional), order, drop self-loops and repeated pairs.
template <typename T>
void removeSelfloop(std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values,
    graphblas::Index* nvals, bool undirected) {
  const bool drop_loops = getEnv("GRB_UTIL_REMOVE_SELFLOOP", true);
  std::vector<graphblas::Index>& r = *row_indices;
  std::vector<graphblas::Index>& c = *col_indices;
  std::vector<T>& v = *values;
  if (undirected) {
    const size_t original = r.size();
    for (size_t i = 0; i < original; ++i) {
      if (r[i] == c[i]) continue;
      const graphblas::Index from = r[i], to = c[i];
      const T weight = v[i];
      r.push_back(to); c.push_back(from); v.push_back(weight);
    }
  }
  if (r.empty()) { *nvals = 0; return; }
  customSort<T>(row_indices, col_indices, values);
  size_t kept = 0;
  for (size_t i = 0; i < r.size(); ++i) {
    const bool loop = drop_loops && r[i] == c[i];
    const bool repeat = i > 0 && r[i] == r[i - 1] && c[i] == c[i - 1];
    if (loop || repeat) continue;
    r[kept] = r[i]; c[kept] = c[i]; v[kept] = v[i];
    ++kept;
  }
  r.resize(kept); c.resize(kept); v.resize(kept);
  *nvals = static_cast<graphblas::Index>(kept);
}

// ---------------------------------------------------------------------------
// Matrix Market coordinate files.
// ---------------------------------------------------------------------------
class MtxFile {
 public:
  MM_typecode type;
  graphblas::Index nrows, ncols, declared;     // size line
  bool ok;

  explicit MtxFile(const char* path) : nrows(0), ncols(0), declared(0), ok(false) {
    FILE* f = fopen(path, "rb");
    if (f == NULL) return;
    if (mm_read_banner(f, &type) == 0 &&
        mm_read_mtx_crd_size(f, &nrows, &ncols, &declared) == 0) {
      const long body = ftell(f);
      fseek(f, 0, SEEK_END);
      const long end = ftell(f);
      fseek(f, body, SEEK_SET);
      text_.resize(static_cast<size_t>(end - body) + 1);
      const size_t got = fread(&text_[0], 1, static_cast<size_t>(end - body), f);
      text_[got] = '\0';
      ok = true;
    }
    fclose(f);
  }

  bool symmetric() const { return mm_is_symmetric(type); }

  // Zero-based tuples in file order; pattern files get value 1.  Stops at the
  // declared count or at the end of the text, whichever comes first.
  template <typename T>
  void tuples(std::vector<graphblas::Index>* rows, std::vector<graphblas::Index>* cols,
              std::vector<T>* vals) const {
    rows->clear(); cols->clear(); vals->clear();
    rows->reserve(declared); cols->reserve(declared); vals->reserve(declared);
    const bool has_value = !mm_is_pattern(type);
    const bool integer_value = mm_is_integer(type);
    const char* p = text_.c_str();
    for (graphblas::Index e = 0; e < declared; ++e) {
      char* next;
      const long r = strtol(p, &next, 10);
      if (next == p) {
        std::cout << "Error: Not enough rows in mtx file!\n";
        break;
      }
      p = next;
      const long c = strtol(p, &next, 10);
      if (next == p) break;
      p = next;
      T value = static_cast<T>(1);
      if (has_value) {
        // the reference reads integer fields with %d and real fields with %f
        if (integer_value) value = static_cast<T>(strtol(p, &next, 10));
        else               value = static_cast<T>(strtof(p, &next));
        if (next == p) break;
        p = next;
        if (mm_is_complex(type)) { strtof(p, &next); p = next; }
      }
      rows->push_back(static_cast<graphblas::Index>(r - 1));
      cols->push_back(static_cast<graphblas::Index>(c - 1));
      vals->push_back(value);
    }
  }

 private:
  std::string text_;
};

// Cache file name for a matrix file (malloc'ed; freed by Matrix::build).
inline char* convert(const char* fname, bool is_undirected = true) {
  const bool drop_loops = getEnv("GRB_UTIL_REMOVE_SELFLOOP", true);
  std::cout << "Remove self-loop: " << drop_loops << std::endl;
  std::string dir_copy(fname), base_copy(fname);
  const std::string dir = dirname(&dir_copy[0]);
  const std::string base = basename(&base_copy[0]);
  const std::string name = dir + "/." + base + (is_undirected ? ".ud." : ".d.") +
      (drop_loops ? "nosl." : "sl.") +
      (sizeof(graphblas::Index) == 8 ? "64bVe." : "") + "bin";
  char* out = reinterpret_cast<char*>(malloc(GRB_MAXLEN));
  snprintf(out, GRB_MAXLEN, "%s", name.c_str());
  return out;
}

// The reference's loader call: fills the tuple vectors (empty when the binary
// cache is to be used instead) and the dimensions.
template <typename T>
int readMtx(const char* fname, std::vector<graphblas::Index>* row_indices,
    std::vector<graphblas::Index>* col_indices, std::vector<T>* values,
    graphblas::Index* nrows, graphblas::Index* ncols, graphblas::Index* nvals,
    int directed, bool mtxinfo, char** dat_name = NULL) {
  MtxFile file(fname);
  if (!file.ok) {
    printf("File %s not found or not a Matrix Market coordinate file\n", fname);
    exit(1);
  }
  *nrows = file.nrows; *ncols = file.ncols; *nvals = file.declared;
  printf("Undirected due to mtx: %d\n", file.symmetric());
  printf("Undirected due to cmd: %d\n", directed == 2);
  const bool undirected = directed != 1 && (file.symmetric() || directed == 2);
  printf("Undirected: %d\n", undirected);
  row_indices->clear(); col_indices->clear(); values->clear();
  if (dat_name != NULL) {
    *dat_name = convert(fname, undirected);
    if (exists(*dat_name)) return 0;        // Matrix::build(dat_name) takes over
  }
  file.tuples<T>(row_indices, col_indices, values);
  *nvals = static_cast<graphblas::Index>(row_indices->size());
  removeSelfloop<T>(row_indices, col_indices, values, nvals, undirected);
  if (mtxinfo) {
    mm_write_banner(stdout, file.type);
    mm_write_mtx_crd_size(stdout, *nrows, *ncols, *nvals);
  }
  return 0;
}

using namespace graphblas;

// ---------------------------------------------------------------------------
// Host conversions (kept for callers written against the reference; the backend
// builds its matrices on the device and does not use them).
// ---------------------------------------------------------------------------
// Tuples -> compressed rows.  Any input order; rows come out sorted by column.
template <typename T>
void coo2csr(Index* csrRowPtr, Index* csrColInd, T* csrVal,
    const std::vector<Index>& row_indices, const std::vector<Index>& col_indices,
    const std::vector<T>& values, Index nrows, Index ncols) {
  const std::vector<unsigned int> order = gbutil::orderByRowCol(row_indices, col_indices);
  std::fill(csrRowPtr, csrRowPtr + nrows + 1, 0);
  for (size_t i = 0; i < order.size(); ++i) {
    const unsigned int src = order[i];
    if (row_indices[src] >= nrows || col_indices[src] >= ncols) {
      std::cout << "Error: Index out of bounds!\n";
      continue;
    }
    csrColInd[i] = col_indices[src];
    csrVal[i] = values[src];
    ++csrRowPtr[row_indices[src] + 1];
  }
  for (Index r = 0; r < nrows; ++r) csrRowPtr[r + 1] += csrRowPtr[r];
}

template <typename T>
void coo2csc(Index* cscColPtr, Index* cscRowInd, T* cscVal,
    const std::vector<Index>& row_indices, const std::vector<Index>& col_indices,
    const std::vector<T>& values, Index nrows, Index ncols) {
  coo2csr(cscColPtr, cscRowInd, cscVal, col_indices, row_indices, values, ncols, nrows);
}

// Compressed rows -> compressed columns by counting (no sort: walking the rows in
// order leaves every column's rows ordered).
template <typename T>
void csr2csc(Index* cscColPtr, Index* cscRowInd, T* cscVal, const Index* csrRowPtr,
    const Index* csrColInd, const T* csrVal, Index nrows, Index ncols) {
  const Index nvals = csrRowPtr[nrows];
  std::fill(cscColPtr, cscColPtr + ncols + 1, 0);
  for (Index k = 0; k < nvals; ++k) ++cscColPtr[csrColInd[k] + 1];
  for (Index c = 0; c < ncols; ++c) cscColPtr[c + 1] += cscColPtr[c];
  std::vector<Index> cursor(cscColPtr, cscColPtr + ncols);
  for (Index r = 0; r < nrows; ++r) {
    for (Index k = csrRowPtr[r]; k < csrRowPtr[r + 1]; ++k) {
      const Index at = cursor[csrColInd[k]]++;
      cscRowInd[at] = r;
      cscVal[at] = csrVal[k];
    }
  }
}

#endif  // GRAPHBLAS_UTIL_HPP_
