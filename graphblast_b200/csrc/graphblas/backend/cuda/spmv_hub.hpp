// graphblast_b200 backend — host side of the hub-cached pull SpMV
// (kernels/spmv_hub.cuh): per-matrix hub index (which columns live in shared
// memory, encoded column array) and the launch.  No reference counterpart: the
// reference hands the generic SpMV to mgpu::SpmvCsrBinary (spmv.hpp:188-190).
#ifndef GRAPHBLAS_BACKEND_CUDA_SPMV_HUB_HPP_
#define GRAPHBLAS_BACKEND_CUDA_SPMV_HUB_HPP_

#include "graphblas/backend/cuda/kernels/kernels.hpp"
#include "graphblas/backend/cuda/hub_index.hpp"
#include "graphblas/backend/cuda/kernels/spmv_hub.cuh"

namespace graphblas {
namespace backend {

#define GB_HUB_GROUPS   8                // 128-thread groups per CTA (one CTA per SM)
#define GB_HUB_CAPACITY 32768            // hub slots (x 4 bytes of shared memory)

// Chooses the `capacity` most referenced columns (ties by arrival), assigns them
// shared-memory slots, encodes the column array, compacts the non-empty rows and
// cuts the tiles.  One-time cost per matrix: a counting pass over colind, ~30
// threshold probes over the n counts, the encoding pass, and a few passes over
// rowptr.
inline void buildHubIndex(HubIndex* h, const Index* rowptr, const Index* colind,
                          Index nrows, Index ncols, Index nnz, int capacity) {
  h->release();
  cudaStream_t s = gbStream();
  Runtime& rt = runtime();
  h->capacity = capacity;
  int* cnt = reinterpret_cast<int*>(gbMalloc(static_cast<size_t>(ncols)*sizeof(int)));
  Index* slot = reinterpret_cast<Index*>(gbMalloc(static_cast<size_t>(ncols)*sizeof(Index)));
  unsigned long long* cells = reinterpret_cast<unsigned long long*>(
      gbMalloc(2*sizeof(unsigned long long)));
  h->enc_ci   = reinterpret_cast<Index*>(gbMalloc((static_cast<size_t>(nnz) + 8)*sizeof(Index)));
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
  gbFree(slot); gbFree(cnt);

  // compact list of the non-empty rows
  const int nblocks = static_cast<int>((static_cast<size_t>(nrows) + GB_HUB_CNT - 1)/GB_HUB_CNT);
  Index* block_off = reinterpret_cast<Index*>(gbMalloc((static_cast<size_t>(nblocks) + 1)*sizeof(Index)));
  Index* d_total = reinterpret_cast<Index*>(cells);
  hubRowCountKernel<<<nblocks, GB_HUB_CNT, 0, s>>>(block_off, rowptr, nrows);
  GB_KERNEL_CHECK();
  hubRowScanKernel<<<1, GB_HUB_CNT, 0, s>>>(block_off, nblocks, d_total);
  GB_KERNEL_CHECK();
  h->m = rt.fetch(d_total);
  h->nempty = nrows - h->m;
  h->ne_ptr  = reinterpret_cast<Index*>(gbMalloc((static_cast<size_t>(h->m) + GB_HUB_PAD)*sizeof(Index)));
  h->ne_rows = reinterpret_cast<Index*>(gbMalloc((static_cast<size_t>(h->m) + GB_HUB_PAD)*sizeof(Index)));
  h->empty_rows = reinterpret_cast<Index*>(gbMalloc((static_cast<size_t>(h->nempty) + 1)*sizeof(Index)));
  hubRowEmitKernel<<<nblocks, GB_HUB_CNT, 0, s>>>(h->ne_rows, h->ne_ptr, h->empty_rows,
      block_off, rowptr, nrows);
  GB_KERNEL_CHECK();
  hubRowPadKernel<<<1, 32, 0, s>>>(h->ne_rows, h->ne_ptr, h->m, nnz);
  GB_KERNEL_CHECK();
  gbFree(block_off);

  // weighted merge-path partition (a row end weighs GB_HUB_RW items)
  const long long total = static_cast<long long>(GB_HUB_RW)*h->m + nnz;
  h->ntiles = static_cast<int>((total + GB_HUB_TILE - 1)/GB_HUB_TILE);
  if (h->ntiles < 1) h->ntiles = 1;
  Index* bounds = reinterpret_cast<Index*>(
      gbMalloc(2*(static_cast<size_t>(h->ntiles) + 1)*sizeof(Index)));
  hubPartitionKernel<<<(h->ntiles + 256)/256, 256, 0, s>>>(bounds, h->ne_ptr, h->m,
      nnz, h->ntiles);
  GB_KERNEL_CHECK();
  h->desc = reinterpret_cast<int4*>(gbMalloc(static_cast<size_t>(h->ntiles)*sizeof(int4)));
  hubTileDescKernel<<<(h->ntiles + 255)/256, 256, 0, s>>>(h->desc, bounds, h->ne_ptr,
      h->m, h->ntiles);
  GB_KERNEL_CHECK();
  const size_t nchunks = (static_cast<size_t>(nnz) + 7)/8;
  h->chunk_rel = reinterpret_cast<unsigned char*>(gbMalloc(nchunks + 16));
  if (nnz > 0 && h->m > 0) {
    hubChunkRowKernel<<<gridFor(nchunks, 256, 8), 256, 0, s>>>(h->chunk_rel, bounds,
        h->ne_ptr, h->m, nnz, h->ntiles);
    GB_KERNEL_CHECK();
  }
  gbFree(bounds); gbFree(cells);
  h->key = colind;
  h->key_nvals = nnz;
}

// w = A (+.x) u through the hub kernel.  carry_row / carry_val: ntiles entries.
template <int GROUPS, int HUB_K, typename W, typename a, typename U, typename SemiringT>
void spmvHubRun(W* out, const HubIndex& h, SemiringT op, const a* val, const U* u,
                Index nnz, Index* carry_row, W* carry_val, cudaStream_t s) {
  typedef decltype(extractMul(op)) MulT;
  typedef decltype(extractAdd(op)) AddT;
  auto kern = spmvHubKernel<GROUPS, HUB_K, W, a, U, MulT, AddT>;
  const HubSmemPlan plan = hubSmemPlan(GROUPS, HUB_K);
  static bool configured = false;        // once per instantiation
  if (!configured) {
    CUDA_CALL(cudaFuncSetAttribute(kern,
        cudaFuncAttributeMaxDynamicSharedMemorySize, plan.total));
    configured = true;
  }
  const W identity = static_cast<W>(op.identity());
  {
    size_t work = static_cast<size_t>(h.nempty) > static_cast<size_t>(HUB_K)
                      ? static_cast<size_t>(h.nempty) : static_cast<size_t>(HUB_K);
    if (work < 1) work = 1;
    hubPrepassKernel<<<gridFor(work, 256, 8), 256, 0, s>>>(
        reinterpret_cast<U*>(h.hub_vals), u, h.hub_ids, h.count, HUB_K,
        static_cast<U>(op.identity()), out, h.empty_rows, h.nempty, identity);
    GB_KERNEL_CHECK();
  }
  int grid = runtime().sm_count;
  const int need = (h.ntiles + GROUPS - 1)/GROUPS;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  HubTiles tiles;
  tiles.desc = h.desc; tiles.chunk_rel = h.chunk_rel;
  tiles.ne_ptr = h.ne_ptr; tiles.ne_rows = h.ne_rows;
  tiles.m = h.m; tiles.ntiles = h.ntiles;
  kern<<<grid, GROUPS*GB_HUB_GT, plan.total, s>>>(out, tiles, carry_row, carry_val,
      h.enc_ci, val, u, reinterpret_cast<const U*>(h.hub_vals), nnz, identity,
      extractMul(op), extractAdd(op));
  GB_KERNEL_CHECK();
  spmvCarryFixupKernel<<<(h.ntiles + 255)/256, 256, 0, s>>>(out, carry_row,
      carry_val, h.ntiles, extractAdd(op));
  GB_KERNEL_CHECK();
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_SPMV_HUB_HPP_
