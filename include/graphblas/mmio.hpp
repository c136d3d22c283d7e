// graphblast_b200 frontend mirror — the slice of the Matrix Market I/O interface
// that the loader uses (reference graphblas/mmio.hpp is the NIST mmio library;
// this is an independent, smaller implementation with the same entry points:
// mm_read_banner, mm_read_mtx_crd_size, mm_write_banner, mm_write_mtx_crd_size
// and the mm_is_* predicates).
#ifndef GRAPHBLAS_MMIO_HPP_
#define GRAPHBLAS_MMIO_HPP_

#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define MM_MAX_LINE_LENGTH 1025
#define MM_PREMATURE_EOF   12
#define MM_NO_HEADER       14
#define MM_UNSUPPORTED_TYPE 15

// typecode[0] object (M), [1] format (C/A), [2] field (R/I/P/C), [3] symmetry
typedef char MM_typecode[4];

#define mm_is_matrix(t)     ((t)[0] == 'M')
#define mm_is_sparse(t)     ((t)[1] == 'C')
#define mm_is_coordinate(t) ((t)[1] == 'C')
#define mm_is_dense(t)      ((t)[1] == 'A')
#define mm_is_complex(t)    ((t)[2] == 'C')
#define mm_is_real(t)       ((t)[2] == 'R')
#define mm_is_pattern(t)    ((t)[2] == 'P')
#define mm_is_integer(t)    ((t)[2] == 'I')
#define mm_is_symmetric(t)  ((t)[3] == 'S')
#define mm_is_general(t)    ((t)[3] == 'G')
#define mm_is_skew(t)       ((t)[3] == 'K')
#define mm_is_hermitian(t)  ((t)[3] == 'H')

inline void mm_lowercase(char* s) {
  for (; *s; ++s) *s = static_cast<char>(tolower(*s));
}

inline int mm_read_banner(FILE* f, MM_typecode* matcode) {
  char line[MM_MAX_LINE_LENGTH];
  char banner[64], object[64], format[64], field[64], symmetry[64];
  (*matcode)[0] = (*matcode)[1] = (*matcode)[2] = ' ';
  (*matcode)[3] = 'G';
  if (fgets(line, MM_MAX_LINE_LENGTH, f) == NULL) return MM_PREMATURE_EOF;
  if (sscanf(line, "%63s %63s %63s %63s %63s", banner, object, format, field,
             symmetry) != 5)
    return MM_PREMATURE_EOF;
  mm_lowercase(object); mm_lowercase(format);
  mm_lowercase(field);  mm_lowercase(symmetry);
  if (strncmp(banner, "%%MatrixMarket", 14) != 0) return MM_NO_HEADER;
  if (strcmp(object, "matrix") != 0) return MM_UNSUPPORTED_TYPE;
  (*matcode)[0] = 'M';
  if      (strcmp(format, "coordinate") == 0) (*matcode)[1] = 'C';
  else if (strcmp(format, "array") == 0)      (*matcode)[1] = 'A';
  else return MM_UNSUPPORTED_TYPE;
  if      (strcmp(field, "real") == 0)    (*matcode)[2] = 'R';
  else if (strcmp(field, "complex") == 0) (*matcode)[2] = 'C';
  else if (strcmp(field, "pattern") == 0) (*matcode)[2] = 'P';
  else if (strcmp(field, "integer") == 0) (*matcode)[2] = 'I';
  else return MM_UNSUPPORTED_TYPE;
  if      (strcmp(symmetry, "general") == 0)        (*matcode)[3] = 'G';
  else if (strcmp(symmetry, "symmetric") == 0)      (*matcode)[3] = 'S';
  else if (strcmp(symmetry, "hermitian") == 0)      (*matcode)[3] = 'H';
  else if (strcmp(symmetry, "skew-symmetric") == 0) (*matcode)[3] = 'K';
  else return MM_UNSUPPORTED_TYPE;
  return 0;
}

// Skips comment lines, then reads "M N nz".
inline int mm_read_mtx_crd_size(FILE* f, int* M, int* N, int* nz) {
  char line[MM_MAX_LINE_LENGTH];
  *M = *N = *nz = 0;
  do {
    if (fgets(line, MM_MAX_LINE_LENGTH, f) == NULL) return MM_PREMATURE_EOF;
  } while (line[0] == '%');
  while (sscanf(line, "%d %d %d", M, N, nz) != 3) {
    if (fgets(line, MM_MAX_LINE_LENGTH, f) == NULL) return MM_PREMATURE_EOF;
  }
  return 0;
}

inline int mm_write_banner(FILE* f, MM_typecode matcode) {
  const char* format = mm_is_sparse(matcode) ? "coordinate" : "array";
  const char* field = mm_is_real(matcode) ? "real" :
                      mm_is_complex(matcode) ? "complex" :
                      mm_is_pattern(matcode) ? "pattern" : "integer";
  const char* symm = mm_is_general(matcode) ? "general" :
                     mm_is_symmetric(matcode) ? "symmetric" :
                     mm_is_hermitian(matcode) ? "hermitian" : "skew-symmetric";
  return fprintf(f, "%%%%MatrixMarket matrix %s %s %s\n", format, field, symm)
         < 0 ? 17 : 0;
}

inline int mm_write_mtx_crd_size(FILE* f, int M, int N, int nz) {
  return fprintf(f, "%d %d %d\n", M, N, nz) < 0 ? 17 : 0;
}

#endif  // GRAPHBLAS_MMIO_HPP_
