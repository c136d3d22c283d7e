/* oracle/gb_oracle.h — CPU restatement of the reference's algorithms for the
 * direction-optimised mxv/vxm + masked-mxm path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this library, and only
 * as the checker / CPU baseline.  The product (graphblast_b200/) never links,
 * loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  Checked in tests/test_oracle.py against
 *   - oracle/_ref/libgbref.so (the reference's own SimpleReference* and loader,
 *     compiled from /root/reference) on the bundled graphs and on R-MAT graphs;
 *   - the known answers in BASELINE.md §2 (chesapeake BFS levels, 194 triangles,
 *     greduce row sums {1,1,3,2,2,3,3,0,1,2,2}).
 */
#ifndef GB_ORACLE_H_
#define GB_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Semiring ids (reference graphblas/stddef.hpp:194-213, REGISTER_SEMIRING order). */
enum {
  ORC_LOGICAL_OR_AND = 0,
  ORC_PLUS_MULTIPLIES,
  ORC_MINIMUM_PLUS,
  ORC_MAXIMUM_MULTIPLIES,
  ORC_PLUS_DIVIDES,
  ORC_PLUS_GREATER,
  ORC_GREATER_PLUS,
  ORC_PLUS_MINUS,
  ORC_PLUS_LESS,
  ORC_CUSTOM_LESS_PLUS,
  ORC_MINIMUM_MULTIPLIES,
  ORC_MULTIPLIES_MULTIPLIES,
  ORC_NOT_EQUAL_TO_PLUS,
  ORC_MINIMUM_SELECT_SECOND,
  ORC_PLUS_NOT_EQUAL_TO,
  ORC_CUSTOM_LESS_LESS,
  ORC_MINIMUM_NOT_EQUAL_TO,
  ORC_NSEMIRINGS
};

float orc_identity(int semiring);
float orc_add(int semiring, float a, float b);
float orc_mul(int semiring, float a, float b);

/* Level-synchronous BFS, levels 1-based, unreachable = 0; returns search depth.
 * reference graphblas/algorithm/test_bfs.hpp:11-61 */
int orc_bfs(int nrows, const int* rowptr, const int* colind, int* levels,
            int src, int stop);

/* Lazy Dijkstra with a binary min-heap, unreachable = FLT_MAX; returns depth.
 * reference graphblas/algorithm/test_sssp.hpp:15-79 */
int orc_sssp(int nrows, const int* rowptr, const int* colind, const float* val,
             float* dist, int src);

/* Push-style power iteration, p0 = 1/n, teleport (1-alpha)/n, out-degrees from
 * rowptr, stops when sum(diff^2) < eps (no sqrt) or after max_niter.
 * reference graphblas/algorithm/test_pr.hpp:15-80 */
int orc_pr(int nrows, const int* rowptr, const int* colind, float* pr,
           float alpha, float eps, int max_niter);

/* Sorted-list intersection triangle count over a (lower-triangular) CSR.
 * reference graphblas/algorithm/test_tc.hpp:15-85 */
long long orc_tc(int nrows, const int* rowptr, const int* colind);

/* w = u^T A over a semiring, push formulation w[col] (+)= A(row,col) (x) u[row]
 * for u[row] "present".  u_present may be NULL (all present).  mask may be NULL;
 * with a mask, positions failing the mask test are set to 0 afterwards:
 * scmp == 0 keeps mask != 0, scmp == 1 keeps mask == 0.
 * reference test/gvxm.cu:41-55, 103-119, 169-192 (inline expected-value loops) */
void orc_vxm(int semiring, int nrows, int ncols, const int* rowptr,
             const int* colind, const float* val, const float* u,
             const unsigned char* u_present, const float* mask, int scmp,
             float* w, unsigned char* w_present);

/* Row-wise monoid reduce of CSR values with (+): reference test/greduce.cu:63-75 */
void orc_reduce_rows(int nrows, const int* rowptr, const float* val, float* w);

/* Loader semantics: symmetrise (optional), drop self-loops and duplicates, sort
 * row-major, emit CSR with all values 1.  Input arrays are modified.  Returns
 * the number of stored entries.  colind must hold 2*nedges ints when
 * undirected.  reference graphblas/util.hpp:264-329 (removeSelfloop),
 * 170-195 (customSort), 502-556 (coo2csr) */
long long orc_build_csr(int nrows, long long nedges, const int* src,
                        const int* dst, int undirected, int* rowptr,
                        int* colind);

/* Keep entries with col <= row.  reference backend/cuda/tri.hpp:21-48 */
long long orc_tril(int nrows, int* rowptr, int* colind);

/* Graph500-style R-MAT edge list, (a,b,c,d) = (0.57,0.19,0.19,0.05), one
 * counter-based SplitMix64 draw per (edge, level): bit-identical to the device
 * generator behind gb200_rmat_edges (the reference has no
 * generator; this one is ours and documented in DESIGN.md). */
void orc_rmat_edges(int scale, long long nedges, unsigned long long seed,
                    long long first_edge, int* src, int* dst);

#ifdef __cplusplus
}
#endif

#endif  /* GB_ORACLE_H_ */
