// graphblast_b200 backend — host side of the hub-cached pull SpMV
// (kernels/spmv_hub.cuh): per-matrix hub index (which columns live in shared
// memory, encoded column array) and the launch.  No reference counterpart: the
// reference hands the generic SpMV to mgpu::SpmvCsrBinary (spmv.hpp:188-190).
#ifndef GRAPHBLAS_BACKEND_CUDA_SPMV_HUB_HPP_
#define GRAPHBLAS_BACKEND_CUDA_SPMV_HUB_HPP_

#include "graphblas/backend/cuda/kernels/kernels.hpp"
#include "graphblas/backend/cuda/kernels/spmv_hub.cuh"

namespace graphblas {
namespace backend {

#define GB_HUB_GROUPS   4
#define GB_HUB_CAPACITY 32768            // hub slots (x 4 bytes of shared memory)

// Per (matrix, direction): which columns are hubs and the encoded column array.
struct HubIndex {
  Index*       enc_ci;      // [nnz] column id, or GB_HUB_FLAG | slot
  Index*       hub_ids;     // [capacity] column id of every slot (first `count`)
  void*        hub_vals;    // [capacity] x 4 bytes, refreshed per call
  Index*       tile_rows;   // weighted merge-path partition: (rows, nonzeros) per boundary
  int          ntiles;
  int          count;       // slots in use
  double       coverage;    // share of the stored entries that reference a hub
  const Index* key;         // colind pointer this was built from
  Index        key_nvals;
  HubIndex() : enc_ci(NULL), hub_ids(NULL), hub_vals(NULL), tile_rows(NULL),
               ntiles(0), count(0), coverage(0.), key(NULL), key_nvals(-1) {}
  void release() {
    if (enc_ci    != NULL) gbFree(enc_ci);
    if (hub_ids   != NULL) gbFree(hub_ids);
    if (hub_vals  != NULL) gbFree(hub_vals);
    if (tile_rows != NULL) gbFree(tile_rows);
    enc_ci = NULL; hub_ids = NULL; hub_vals = NULL; tile_rows = NULL;
    ntiles = 0; count = 0; coverage = 0.; key = NULL; key_nvals = -1;
  }
};

// Chooses the `capacity` most referenced columns (ties by arrival), assigns them
// shared-memory slots and encodes the column array.  One-time cost per matrix:
// a counting pass over colind, ~30 threshold probes over the n counts, two
// assignment passes and the encoding pass.
inline void buildHubIndex(HubIndex* h, const Index* rowptr, const Index* colind,
                          Index nrows, Index ncols, Index nnz, int capacity) {
  h->release();
  cudaStream_t s = gbStream();
  Runtime& rt = runtime();
  int* cnt = reinterpret_cast<int*>(gbMalloc(static_cast<size_t>(ncols)*sizeof(int)));
  Index* slot = reinterpret_cast<Index*>(gbMalloc(static_cast<size_t>(ncols)*sizeof(Index)));
  unsigned long long* cells = reinterpret_cast<unsigned long long*>(
      gbMalloc(2*sizeof(unsigned long long)));
  h->enc_ci   = reinterpret_cast<Index*>(gbMalloc(static_cast<size_t>(nnz)*sizeof(Index)));
  h->hub_ids  = reinterpret_cast<Index*>(gbMalloc(static_cast<size_t>(capacity)*sizeof(Index)));
  h->hub_vals = gbMalloc(static_cast<size_t>(capacity)*4);
  CUDA_CALL(cudaMemsetAsync(cnt, 0, static_cast<size_t>(ncols)*sizeof(int), s));
  hubCountKernel<<<gridFor(nnz, 256, 8), 256, 0, s>>>(cnt, colind, nnz);
  GB_KERNEL_CHECK();
  // smallest t with #{cnt > t} <= capacity
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    CUDA_CALL(cudaMemsetAsync(cells, 0, 2*sizeof(unsigned long long), s));
    hubAboveKernel<<<gridFor(ncols, 256, 4), 256, 0, s>>>(cells, cnt, ncols,
        static_cast<int>(mid + 1));
    GB_KERNEL_CHECK();
    const unsigned long long above = rt.fetch(cells);
    if (above <= static_cast<unsigned long long>(capacity)) hi = mid; else lo = mid + 1;
  }
  const int t = static_cast<int>(lo);
  CUDA_CALL(cudaMemsetAsync(cells, 0, 2*sizeof(unsigned long long), s));
  hubAssignKernel<<<gridFor(ncols, 256, 4), 256, 0, s>>>(slot, h->hub_ids, cells,
      cnt, ncols, t, capacity, 0);
  GB_KERNEL_CHECK();
  hubAssignKernel<<<gridFor(ncols, 256, 4), 256, 0, s>>>(slot, h->hub_ids, cells,
      cnt, ncols, t, capacity, 1);
  GB_KERNEL_CHECK();
  unsigned long long taken = 0, covered = 0;
  rt.fetch2(cells, &taken, &covered);
  h->count = static_cast<int>(taken);
  h->coverage = nnz > 0 ? static_cast<double>(covered)/static_cast<double>(nnz) : 0.;
  hubEncodeKernel<<<gridFor(nnz, 256, 8), 256, 0, s>>>(h->enc_ci, colind, slot, nnz);
  GB_KERNEL_CHECK();
  // weighted merge-path partition (a row end weighs GB_HUB_RW items)
  const long long total = static_cast<long long>(GB_HUB_RW)*nrows + nnz;
  h->ntiles = static_cast<int>((total + GB_HUB_TILE - 1)/GB_HUB_TILE);
  if (h->ntiles < 1) h->ntiles = 1;
  h->tile_rows = reinterpret_cast<Index*>(
      gbMalloc(2*(static_cast<size_t>(h->ntiles) + 1)*sizeof(Index)));
  hubPartitionKernel<<<(h->ntiles + 256)/256, 256, 0, s>>>(h->tile_rows,
      rowptr, nrows, nnz, h->ntiles);
  GB_KERNEL_CHECK();
  gbFree(cells); gbFree(slot); gbFree(cnt);
  h->key = colind;
  h->key_nvals = nnz;
}

// w = A (+.x) u through the hub kernel.  carry_row / carry_val: ntiles entries.
template <int GROUPS, int HUB_K, int PF, typename W, typename a, typename U, typename SemiringT>
void spmvHubRun(W* out, const HubIndex& h, SemiringT op, const Index* rowptr,
                const a* val, const U* u, Index nrows, Index nnz,
                Index* carry_row, W* carry_val, cudaStream_t s) {
  typedef decltype(extractMul(op)) MulT;
  typedef decltype(extractAdd(op)) AddT;
  auto kern = spmvHubKernel<GROUPS, HUB_K, PF, W, a, U, MulT, AddT>;
  const HubSmemPlan plan = hubSmemPlan(GROUPS, HUB_K);
  static bool configured = false;        // once per instantiation
  if (!configured) {
    CUDA_CALL(cudaFuncSetAttribute(kern,
        cudaFuncAttributeMaxDynamicSharedMemorySize, plan.total));
    configured = true;
  }
  if (HUB_K > 0) {
    hubGatherKernel<<<(HUB_K + 255)/256, 256, 0, s>>>(
        reinterpret_cast<U*>(h.hub_vals), u, h.hub_ids, h.count, HUB_K,
        static_cast<U>(op.identity()));
    GB_KERNEL_CHECK();
  }
  int grid = runtime().sm_count;
  const int need = (h.ntiles + GROUPS - 1)/GROUPS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<grid, GROUPS*GB_HUB_GT, plan.total, s>>>(out, h.tile_rows, carry_row,
      carry_val, rowptr, h.enc_ci, val, u, reinterpret_cast<const U*>(h.hub_vals),
      nrows, nnz, h.ntiles, static_cast<W>(op.identity()), extractMul(op),
      extractAdd(op));
  GB_KERNEL_CHECK();
  spmvCarryFixupKernel<<<(h.ntiles + 255)/256, 256, 0, s>>>(out, carry_row,
      carry_val, h.ntiles, extractAdd(op));
  GB_KERNEL_CHECK();
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_SPMV_HUB_HPP_
