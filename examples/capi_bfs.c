/* examples/capi_bfs.c — the C ABI (include/graphblast_b200.h) from plain C: what
 * reference example/gbfs.cu does (load a .mtx, direction-optimised BFS from a source,
 * print the level histogram), without a C++ compiler on the caller's side.
 *
 *   gcc -std=c99 -I include examples/capi_bfs.c -L graphblast_b200/lib \
 *       -lgraphblast_b200 -Wl,-rpath,$PWD/graphblast_b200/lib -o capi_bfs
 *   ./capi_bfs tests/golden/chesapeake.mtx 0
 */
#include <stdio.h>
#include <stdlib.h>

#include "graphblast_b200.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int info_ = (call);                                                    \
    if (info_ != 0) {                                                      \
      fprintf(stderr, "%s failed with GrB info %d\n", #call, info_);       \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s graph.mtx [source]\n", argv[0]);
    return 2;
  }
  const int source = argc > 2 ? atoi(argv[2]) : 0;

  gb200_matrix_t A;
  gb200_vector_t v;
  gb200_desc_t desc;
  int n = 0, i, depth = 0;
  float tight_ms = 0.f;
  float* levels;

  CHECK(gb200_init(0));
  /* readMtx + Matrix::build as the reference drivers do; 2 = treat as undirected */
  CHECK(gb200_matrix_load_mtx(&A, GB200_FP32, argv[1], 2));
  CHECK(gb200_matrix_nrows(A, &n));
  CHECK(gb200_vector_new(&v, GB200_FP32, n));
  CHECK(gb200_desc_new(&desc));
  /* the flags of reference run_bfs.sh */
  CHECK(gb200_desc_set_knob(desc, "mxvmode", 0));
  CHECK(gb200_desc_set_knob(desc, "struconly", 1));
  CHECK(gb200_desc_set_knob(desc, "opreuse", 1));
  CHECK(gb200_desc_set_knob(desc, "earlyexit", 1));

  CHECK(gb200_bfs(v, A, source, desc, &tight_ms));

  levels = (float*)malloc((size_t)n * sizeof(float));
  if (levels == NULL) return 1;
  CHECK(gb200_vector_extract_dense(v, levels, n));
  for (i = 0; i < n; ++i)
    if ((int)levels[i] > depth) depth = (int)levels[i];
  printf("n = %d, source = %d, depth = %d, device loop %.3f ms\n", n, source, depth,
         tight_ms);
  {
    int d;
    for (d = 0; d <= depth; ++d) {
      int count = 0;
      for (i = 0; i < n; ++i) count += ((int)levels[i] == d);
      printf("  level %d: %d vertices%s\n", d, count, d == 0 ? " (unreached)" : "");
    }
  }
  free(levels);
  gb200_vector_free(v);
  gb200_desc_free(desc);
  gb200_matrix_free(A);
  return 0;
}
