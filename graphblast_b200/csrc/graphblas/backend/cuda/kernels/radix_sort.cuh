// graphblast_b200 backend — device-wide exclusive scan and stable LSD radix sort of
// 64-bit keys with a 32-bit payload.  Building blocks of the device ingest
// (ingest.hpp): the reference sorts edge tuples on the HOST with std::sort
// (graphblas/util.hpp:170-195, minutes at R-MAT scale 24) and its GPU paths lean on
// cub/moderngpu; nothing here uses either.
//
// Scan: three launches (block-local scan + block totals, one-CTA scan of the
// totals, add-back).
// Sort: one pass per 8-bit digit, three launches per pass:
//   1. radixHistogramKernel  per-tile digit counts -> hist[digit][tile]
//   2. scan of hist (digit-major), giving every (digit, tile) its global offset
//   3. radixScatterKernel    the tile is ranked stably (warp match_any + per-warp
//                            counters), staged in shared memory in digit order and
//                            written out as runs, so stores are contiguous per digit
// Callers sort only the bits their keys use (ingest.hpp packs row and column
// into the fewest bits), so no pass is wasted on constant digits.
#ifndef GRAPHBLAS_BACKEND_CUDA_KERNELS_RADIX_SORT_CUH_
#define GRAPHBLAS_BACKEND_CUDA_KERNELS_RADIX_SORT_CUH_

#include "graphblas/backend/cuda/kernels/common.cuh"

namespace graphblas {
namespace backend {

// ---------------------------------------------------------------------------
// Exclusive scan of 32-bit counts, in place.
// ---------------------------------------------------------------------------
#define GB_SCAN_NT   1024
#define GB_SCAN_IPT  4
#define GB_SCAN_TILE (GB_SCAN_NT*GB_SCAN_IPT)

__global__ void __launch_bounds__(GB_SCAN_NT)
scanTileKernel(int* __restrict__ data, int* __restrict__ tile_total, long long n) {
  __shared__ int s_scan[GB_SCAN_NT/32 + 1];
  const long long base = static_cast<long long>(blockIdx.x)*GB_SCAN_TILE +
                         static_cast<long long>(threadIdx.x)*GB_SCAN_IPT;
  int v[GB_SCAN_IPT];
  int sum = 0;
#pragma unroll
  for (int j = 0; j < GB_SCAN_IPT; ++j) {
    v[j] = (base + j < n) ? data[base + j] : 0;
    sum += v[j];
  }
  int total;
  int run = blockExclusiveScan<GB_SCAN_NT>(sum, s_scan, &total);
#pragma unroll
  for (int j = 0; j < GB_SCAN_IPT; ++j) {
    if (base + j < n) data[base + j] = run;
    run += v[j];
  }
  if (threadIdx.x == 0) tile_total[blockIdx.x] = total;
}

// One CTA: exclusive scan of the tile totals in place; *grand = sum of everything.
__global__ void __launch_bounds__(GB_SCAN_NT)
scanTotalsKernel(int* __restrict__ tile_total, int ntiles,
                 unsigned long long* __restrict__ grand) {
  __shared__ int s_scan[GB_SCAN_NT/32 + 1];
  long long running = 0;
  for (int base = 0; base < ntiles; base += GB_SCAN_NT) {
    const int idx = base + threadIdx.x;
    const int v = idx < ntiles ? tile_total[idx] : 0;
    int sum;
    const int excl = blockExclusiveScan<GB_SCAN_NT>(v, s_scan, &sum);
    if (idx < ntiles) tile_total[idx] = static_cast<int>(running) + excl;
    running += sum;
  }
  if (threadIdx.x == 0 && grand != NULL)
    *grand = static_cast<unsigned long long>(running);
}

__global__ void __launch_bounds__(GB_SCAN_NT)
scanAddKernel(int* __restrict__ data, const int* __restrict__ tile_offset,
              long long n) {
  const int add = tile_offset[blockIdx.x];
  const long long base = static_cast<long long>(blockIdx.x)*GB_SCAN_TILE +
                         static_cast<long long>(threadIdx.x)*GB_SCAN_IPT;
#pragma unroll
  for (int j = 0; j < GB_SCAN_IPT; ++j)
    if (base + j < n) data[base + j] += add;
}

// ---------------------------------------------------------------------------
// Radix sort pass.
// ---------------------------------------------------------------------------
#define GB_RADIX_NT    256
#define GB_RADIX_IPT   8
#define GB_RADIX_TILE  (GB_RADIX_NT*GB_RADIX_IPT)      // 2048 keys per CTA (32 KB of static shared memory)
#define GB_RADIX_BINS  256

// `shift` carries the digit position in its low 8 bits and, above them, how many
// bits of the digit count (the last pass of a sort over a bit count that is not a
// multiple of 8 looks at fewer than 8).
__device__ __forceinline__ int radixDigit(unsigned long long key, int shift) {
  const int pos = shift & 0xff;
  const int width = shift >> 8;
  return static_cast<int>((key >> pos) & ((1ull << width) - 1ull));
}

// hist[d*ntiles + tile] = #keys of the tile whose digit is d
__global__ void __launch_bounds__(GB_RADIX_NT)
radixHistogramKernel(int* __restrict__ hist,
                     const unsigned long long* __restrict__ keys,
                     long long n, int shift, int ntiles) {
  __shared__ int s_hist[GB_RADIX_BINS];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const long long base = static_cast<long long>(blockIdx.x)*GB_RADIX_TILE;
#pragma unroll
  for (int r = 0; r < GB_RADIX_IPT; ++r) {
    const long long i = base + r*GB_RADIX_NT + threadIdx.x;
    if (i < n) atomicAdd(&s_hist[radixDigit(keys[i], shift)], 1);
  }
  __syncthreads();
  hist[static_cast<size_t>(threadIdx.x)*ntiles + blockIdx.x] = s_hist[threadIdx.x];
}

// Stable scatter of one tile.  Round r handles keys r*256 + tid of the tile, so
// input order = (round, thread); within a round a warp ranks its 32 keys with
// match_any, the warps are ordered through per-warp counters.
template <bool HasPayload>
__global__ void __launch_bounds__(GB_RADIX_NT)
radixScatterKernel(unsigned long long* __restrict__ keys_out,
                   unsigned int* __restrict__ pay_out,
                   const unsigned long long* __restrict__ keys_in,
                   const unsigned int* __restrict__ pay_in,
                   const int* __restrict__ offsets,     // scanned hist, digit-major
                   long long n, int shift, int ntiles) {
  __shared__ unsigned long long s_keys[GB_RADIX_TILE];
  __shared__ unsigned int s_pay[HasPayload ? GB_RADIX_TILE : 1];
  __shared__ int s_wcount[GB_RADIX_NT/32][GB_RADIX_BINS];
  __shared__ int s_run[GB_RADIX_BINS];      // keys of digit d seen so far in the tile
  __shared__ int s_start[GB_RADIX_BINS];    // first staged slot of digit d
  const int tid  = threadIdx.x;
  const int lane = tid & 31;
  const int wid  = tid >> 5;
  const long long base = static_cast<long long>(blockIdx.x)*GB_RADIX_TILE;
  long long left = n - base;
  const int count = left < GB_RADIX_TILE ? static_cast<int>(left) : GB_RADIX_TILE;

  unsigned long long key[GB_RADIX_IPT];
  unsigned int pay[HasPayload ? GB_RADIX_IPT : 1];
  int rank[GB_RADIX_IPT];                   // rank among the tile's keys of the same digit
#pragma unroll
  for (int r = 0; r < GB_RADIX_IPT; ++r) {
    const int j = r*GB_RADIX_NT + tid;
    key[r] = (j < count) ? keys_in[base + j] : ~0ull;
    if (HasPayload) pay[r] = (j < count) ? pay_in[base + j] : 0u;
  }
  s_run[tid] = 0;
#pragma unroll
  for (int r = 0; r < GB_RADIX_IPT; ++r) {
#pragma unroll
    for (int w = 0; w < GB_RADIX_NT/32; ++w) s_wcount[w][tid] = 0;
    __syncthreads();
    const int j = r*GB_RADIX_NT + tid;
    const bool valid = j < count;
    const int d = radixDigit(key[r], shift);
    // lanes with an invalid key form their own group (digit code 256)
    const unsigned peers = __match_any_sync(GB_FULL_MASK, valid ? d : GB_RADIX_BINS);
    const int in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && in_warp == 0) s_wcount[wid][d] = __popc(peers);
    __syncthreads();
    {                                        // thread tid owns digit tid
      int running = s_run[tid];
#pragma unroll
      for (int w = 0; w < GB_RADIX_NT/32; ++w) {
        const int c = s_wcount[w][tid];
        s_wcount[w][tid] = running;
        running += c;
      }
      s_run[tid] = running;
    }
    __syncthreads();
    rank[r] = valid ? (s_wcount[wid][d] + in_warp) : 0;
    __syncthreads();
  }
  // digit totals of the tile -> first staged slot per digit (one-CTA scan)
  {
    __shared__ int s_scan[GB_RADIX_NT/32 + 1];
    int total;
    const int excl = blockExclusiveScan<GB_RADIX_NT>(s_run[tid], s_scan, &total);
    s_start[tid] = excl;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < GB_RADIX_IPT; ++r) {
    const int j = r*GB_RADIX_NT + tid;
    if (j < count) {
      const int slot = s_start[radixDigit(key[r], shift)] + rank[r];
      s_keys[slot] = key[r];
      if (HasPayload) s_pay[slot] = pay[r];
    }
  }
  __syncthreads();
  // staged in digit order: slot s of digit d goes to offsets[d][tile] + (s - start[d])
  for (int s = tid; s < count; s += GB_RADIX_NT) {
    const unsigned long long k = s_keys[s];
    const int d = radixDigit(k, shift);
    const long long pos = static_cast<long long>(
        offsets[static_cast<size_t>(d)*ntiles + blockIdx.x]) + (s - s_start[d]);
    keys_out[pos] = k;
    if (HasPayload) pay_out[pos] = s_pay[s];
  }
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_KERNELS_RADIX_SORT_CUH_
