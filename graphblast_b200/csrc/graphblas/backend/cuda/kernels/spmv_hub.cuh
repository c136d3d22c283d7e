// graphblast_b200 backend — hub-cached, TMA-staged generic-semiring pull SpMV.
//
// Same contract as spmvMergeKernelT (kernels/spmv_pull.cuh; replaces
// mgpu::SpmvCsrBinary, reference spmv.hpp:188-190 and
// ext/moderngpu/include/kernels/spmvcsr.cuh:334-413,489-587), built for what
// bounds that kernel on B200: every gathered u[col] of a power-law graph is its
// own L1 wavefront and its own 32-byte L2 sector, 128 M of them on RMAT-22
// against 1.07 GB of streamed CSR.  Here
//
//  * the K most referenced columns ("hubs", chosen once per matrix) live in
//    SHARED memory: a per-call compact copy hub_vals[K] = u[hub_ids[K]] is
//    brought in with one TMA bulk copy per CTA, and the matrix carries an encoded
//    column array (bit 31 set: low bits = hub slot; clear: the column id), so a
//    hub reference costs a bank access instead of a sector;
//  * colind / val / rowptr windows of every tile are staged into shared memory
//    with cp.async.bulk + mbarrier (SASS: UBLKCP) and prefetched into L2 a few
//    tiles ahead (cp.async.bulk.prefetch.L2), so the streamed bytes in flight no
//    longer compete with the gathers for L1 lines (the old kernel lost 2.5x when
//    L1 shrank).  A group owns ONE colind/val buffer: its threads move their 8
//    entries to registers, meet at a named barrier, and the buffer is refilled
//    for the next tile while the gathers of this one are in flight — the
//    registers are the second buffer;
//  * tiles come from a WEIGHTED merge path (a row end weighs GB_HUB_RW nonzeros),
//    so a tile has at most 253 row ends and its rowptr window is 1 KB;
//  * one persistent CTA per SM; independent 128-thread groups per CTA, each with
//    its own buffers and named barriers, so one group's reduction overlaps the
//    other groups' gathers;
//  * threads own 8 consecutive nonzeros, reduce them in registers (no product
//    round trip through shared memory), find their first row with a binary
//    search over the staged row offsets; the cross-thread part is the same
//    shuffle-based segmented scan + per-tile carry as before
//    (spmvCarryFixupKernel folds the carries).
//
// Algorithmic bytes per launch: 4(n+1) + 8 nnz + 4n + 4n.
#ifndef GRAPHBLAS_BACKEND_CUDA_KERNELS_SPMV_HUB_CUH_
#define GRAPHBLAS_BACKEND_CUDA_KERNELS_SPMV_HUB_CUH_

#include "graphblas/backend/cuda/kernels/common.cuh"

namespace graphblas {
namespace backend {

#define GB_HUB_GT    128                 // threads per tile group
#define GB_HUB_RW    4                   // weight of a row end in merge items
#define GB_HUB_TILE  1008                // weighted merge items per tile
#define GB_HUB_WIN   1024                // staged colind / val elements (128 chunks)
#define GB_HUB_RPWIN 264                 // staged rowptr entries (<= 253 + 2 + 6)
#define GB_HUB_FLAG  0x80000000u         // encoded column: hub slot in the low bits

// ---------------------------------------------------------------------------
// mbarrier / bulk-copy (TMA, non-tensor form) primitives.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
               :: "r"(smemAddr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbarFenceInit() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbarTryWait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n"
               ".reg .pred p;\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
               "selp.u32 %0, 1, 0, p;\n"
               "}\n"
               : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
  return ok != 0;
}

__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
  while (!mbarTryWait(bar, parity)) { }
}

__device__ __forceinline__ uint64_t makeEvictFirstPolicy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;"
               : "=l"(pol));
  return pol;
}

// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulkLoad(void* dst, const void* src,
                                         uint32_t bytes, uint64_t* bar,
                                         uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
               ".L2::cache_hint [%0], [%1], %2, [%3], %4;"
               :: "r"(smemAddr(dst)), "l"(src), "r"(bytes), "r"(smemAddr(bar)),
                  "l"(pol) : "memory");
}

// L2 prefetch of a window that a later bulk copy will read.
__device__ __forceinline__ void bulkPrefetchL2(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
               :: "l"(src), "r"(bytes) : "memory");
}

__device__ __forceinline__ void groupBarrier(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
// Weighted merge-path partition.  With f(r) = RW*r + rowptr[r] (items consumed
// when exactly r rows are finished), tile boundary c sits on diagonal
// d = c*TILE: rows consumed = max{r : f(r) <= d}, nonzeros consumed =
// min(d - RW*r, rowptr[r+1]) (a row end cut by the diagonal goes to the next
// tile).  tile_rk[2c] = rows, tile_rk[2c+1] = nonzeros, for c in [0, ntiles].
// ---------------------------------------------------------------------------
__global__ void hubPartitionKernel(Index* __restrict__ tile_rk,
                                   const Index* __restrict__ rowptr,
                                   Index nrows, Index nnz, int ntiles) {
  const int c = blockIdx.x*blockDim.x + threadIdx.x;
  if (c > ntiles) return;
  const long long total = static_cast<long long>(GB_HUB_RW)*nrows + nnz;
  long long d = static_cast<long long>(c)*GB_HUB_TILE;
  if (d > total) d = total;
  long long lo = 0, hi = nrows;              // largest r with f(r) <= d
  while (lo < hi) {
    const long long mid = (lo + hi + 1) >> 1;
    if (GB_HUB_RW*mid + static_cast<long long>(__ldg(rowptr + mid)) <= d) lo = mid;
    else hi = mid - 1;
  }
  const Index r = static_cast<Index>(lo);
  long long k = d - static_cast<long long>(GB_HUB_RW)*r;
  const long long kmax = (r < nrows) ? static_cast<long long>(__ldg(rowptr + r + 1))
                                     : static_cast<long long>(nnz);
  if (k > kmax) k = kmax;
  tile_rk[2*c]     = r;
  tile_rk[2*c + 1] = static_cast<Index>(k);
}

// ---------------------------------------------------------------------------
// Hub selection (once per matrix structure).
// ---------------------------------------------------------------------------
// cnt[col] += 1 for every stored entry.
__global__ void hubCountKernel(int* __restrict__ cnt,
                               const Index* __restrict__ colind, Index nnz) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < nnz; k += stride) atomicAdd(cnt + __ldg(colind + k), 1);
}

// out[0] = #{c : cnt[c] >= t}, out[1] = sum of those counts.
__global__ void hubAboveKernel(unsigned long long* __restrict__ out,
                               const int* __restrict__ cnt, Index n, int t) {
  __shared__ int s_red[256/32];
  Index c = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  int num = 0;
  unsigned long long sum = 0ull;
  for (; c < n; c += stride) {
    const int v = cnt[c];
    if (v >= t) { ++num; sum += static_cast<unsigned long long>(v); }
  }
  const int total = blockSum<256>(num, s_red);
  if (threadIdx.x == 0 && total)
    atomicAdd(out, static_cast<unsigned long long>(total));
  // 64-bit sum: warp shuffle, then one atomic per warp
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    sum += __shfl_xor_sync(GB_FULL_MASK, sum, off);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(out + 1, sum);
}

// Columns with cnt > t take a slot unconditionally (the host chose t so that they
// fit), columns with cnt == t while slots remain.  slot[c] = hub slot or -1.
// taken[0] = slots handed out, taken[1] = references covered.
__global__ void hubAssignKernel(Index* __restrict__ slot,
                                Index* __restrict__ hub_ids,
                                unsigned long long* __restrict__ taken,
                                const int* __restrict__ cnt, Index n, int t,
                                int capacity, int pass) {
  Index c = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; c < n; c += stride) {
    const int v = cnt[c];
    if (pass == 0) {
      Index s = -1;
      if (v > t) {
        s = static_cast<Index>(atomicAdd(taken, 1ull));
        hub_ids[s] = c;
        atomicAdd(taken + 1, static_cast<unsigned long long>(v));
      }
      slot[c] = s;
    } else if (v == t && v > 0) {
      const unsigned long long s = atomicAdd(taken, 1ull);
      if (s < static_cast<unsigned long long>(capacity)) {
        hub_ids[s] = c;
        slot[c] = static_cast<Index>(s);
        atomicAdd(taken + 1, static_cast<unsigned long long>(v));
      } else {
        atomicAdd(taken, static_cast<unsigned long long>(-1ll));
      }
    }
  }
}

// enc[k] = hub slot | FLAG, or the column id itself.
__global__ void hubEncodeKernel(Index* __restrict__ enc,
                                const Index* __restrict__ colind,
                                const Index* __restrict__ slot, Index nnz) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < nnz; k += stride) {
    const Index c = __ldg(colind + k);
    const Index s = __ldg(slot + c);
    enc[k] = (s >= 0) ? static_cast<Index>(GB_HUB_FLAG | static_cast<unsigned>(s))
                      : c;
  }
}

// Per call: hub_vals[j] = u[hub_ids[j]]; slots past the count get `pad`.
template <typename U>
__global__ void hubGatherKernel(U* __restrict__ hub_vals,
                                const U* __restrict__ u,
                                const Index* __restrict__ hub_ids,
                                int count, int capacity, U pad) {
  int j = blockIdx.x*blockDim.x + threadIdx.x;
  if (j < count)         hub_vals[j] = __ldg(u + __ldg(hub_ids + j));
  else if (j < capacity) hub_vals[j] = pad;
}

// ---------------------------------------------------------------------------
// Shared-memory plan of the SpMV kernel (bytes), shared by host and device.
// ---------------------------------------------------------------------------
struct HubSmemPlan {
  int hub_bytes;        // K*4
  int group_bytes;      // colind + val window, 2 rowptr windows, meta, scan scratch
  int bar_offset;
  int total;
};

__host__ __device__ inline HubSmemPlan hubSmemPlan(int groups, int hub_k) {
  HubSmemPlan p;
  p.hub_bytes   = hub_k*4;
  // colind[WIN] val[WIN] rp[2][RPWIN] meta[2][4] wkey[2][4] wval[2][4]
  p.group_bytes = 2*GB_HUB_WIN*4 + 2*GB_HUB_RPWIN*4 + 3*32;
  p.bar_offset  = p.hub_bytes + groups*p.group_bytes;
  p.total       = p.bar_offset + 8*(groups + 1) + 8;
  return p;
}

template <int GROUPS, int HUB_K, int PF, typename W, typename a, typename U,
          typename MulOp, typename AddOp>
__global__ void __launch_bounds__(GROUPS*GB_HUB_GT, 1)
spmvHubKernel(W* __restrict__           w,
              const Index* __restrict__ tile_rk,
              Index* __restrict__       carry_row,
              W* __restrict__           carry_val,
              const Index* __restrict__ rowptr,
              const Index* __restrict__ enc_colind,
              const a* __restrict__     val,
              const U* __restrict__     u,
              const U* __restrict__     hub_vals,
              Index                     nrows,
              Index                     nnz,
              int                       ntiles,
              W                         identity,
              MulOp                     mul_op,
              AddOp                     add_op) {
  static_assert(sizeof(W) == 4 && sizeof(U) == 4 && sizeof(a) == 4 &&
                sizeof(Index) == 4, "32-bit values and indices");
  extern __shared__ __align__(128) unsigned char s_raw[];
  const HubSmemPlan plan = hubSmemPlan(GROUPS, HUB_K);

  const int tid  = threadIdx.x;
  const int g    = tid / GB_HUB_GT;            // group in CTA
  const int t    = tid % GB_HUB_GT;            // thread in group
  const int lane = t & 31;
  const int wid  = t >> 5;                     // warp in group (0..3)

  unsigned char* const s_group = s_raw + plan.hub_bytes + g*plan.group_bytes;
  Index* const st_ci  = reinterpret_cast<Index*>(s_group);
  a*     const st_va  = reinterpret_cast<a*>(s_group + GB_HUB_WIN*4);
  Index* const st_rp  = reinterpret_cast<Index*>(s_group + 2*GB_HUB_WIN*4);
  int*   const s_meta = reinterpret_cast<int*>(st_rp + 2*GB_HUB_RPWIN);
  Index* const s_wkey = reinterpret_cast<Index*>(s_meta + 8);
  W*     const s_wval = reinterpret_cast<W*>(s_wkey + 8);
  uint64_t* const s_bar = reinterpret_cast<uint64_t*>(s_raw + plan.bar_offset);
  uint64_t* const bar_hub  = s_bar + GROUPS;
  uint64_t* const bar_full = s_bar + g;

  if (tid == 0) {
    for (int i = 0; i < GROUPS + 1; ++i) mbarInit(s_bar + i, 1);
    mbarFenceInit();
  }
  // Entries of the colind window that a tile does not cover keep whatever an
  // earlier tile left there; they are gathered (and ignored), so they must be
  // valid column codes from the start.
  for (int i = t; i < GB_HUB_WIN; i += GB_HUB_GT) st_ci[i] = 0;
  __syncthreads();

  const uint64_t pol_stream = makeEvictFirstPolicy();
  const uint64_t pol_keep   = makeEvictLastPolicy();

  if (tid == 0 && HUB_K > 0) {
    mbarExpectTx(bar_hub, HUB_K*4);
    bulkLoad(s_raw, hub_vals, HUB_K*4, bar_hub, pol_keep);
  }

  const int gg      = blockIdx.x*GROUPS + g;   // global group id
  const int gstride = gridDim.x*GROUPS;

  // Producer step (thread 0 of the group): stage this group's tile number q whose
  // boundaries (r0,k0)-(r1,k1) the caller read from tile_rk.
  auto issue = [&](int q, Index r0, Index k0, Index r1, Index k1) {
    const int b = q & 1;
    Index* const rpw = st_rp + b*GB_HUB_RPWIN;
    s_meta[4*b + 0] = r0;
    s_meta[4*b + 1] = r1;
    s_meta[4*b + 2] = k0;
    s_meta[4*b + 3] = k1;
    uint32_t bytes = 0;
    // nonzero window [k0a, kend): bulk part up to the last whole 16 bytes of the
    // arrays, the (< 4 element) tail with plain loads.
    const Index k0a = k0 & ~7;
    Index kb_end = k0a;
    if (k1 > k0) {
      const Index kend = (k1 + 3) & ~3;
      const Index klim = nnz & ~3;
      kb_end = kend < klim ? kend : klim;
      if (kb_end < k0a) kb_end = k0a;
      for (Index k = kb_end; k < k1; ++k) {
        st_ci[k - k0a] = __ldg(enc_colind + k);
        st_va[k - k0a] = __ldg(val + k);
      }
      bytes += 2u*static_cast<uint32_t>(kb_end - k0a)*4u;
    }
    // row offsets rowptr[r0 .. r1+1] (window starts at ra); entries past the
    // array hold the sentinel nnz.
    const Index ra = r0 & ~3;
    const Index rneed = r1 + 2;                       // exclusive
    const Index rend = (rneed + 3) & ~3;
    const Index rlim = (nrows + 1) & ~3;
    Index rb_end = rend < rlim ? rend : rlim;
    if (rb_end < ra) rb_end = ra;
    for (Index e = rb_end; e < rneed; ++e)
      rpw[e - ra] = (e <= nrows) ? __ldg(rowptr + e) : nnz;
    bytes += static_cast<uint32_t>(rb_end - ra)*4u;
    mbarExpectTx(bar_full, bytes);
    if (kb_end > k0a) {
      const uint32_t nb = static_cast<uint32_t>(kb_end - k0a)*4u;
      bulkLoad(st_ci, enc_colind + k0a, nb, bar_full, pol_stream);
      bulkLoad(st_va, val + k0a, nb, bar_full, pol_stream);
    }
    if (rb_end > ra)
      bulkLoad(rpw, rowptr + ra, static_cast<uint32_t>(rb_end - ra)*4u,
               bar_full, pol_stream);
  };

  // Thread 0 keeps the boundaries of the next tile to stage in registers, so the
  // refill after the barrier does not wait for a global load.
  Index nr0 = 0, nk0 = 0, nr1 = 0, nk1 = 0;
  auto fetchBounds = [&](int q) {
    const long long tile = static_cast<long long>(gg) +
                           static_cast<long long>(q)*gstride;
    if (tile < ntiles) {
      const int4 b = make_int4(__ldg(tile_rk + 2*tile), __ldg(tile_rk + 2*tile + 1),
                               __ldg(tile_rk + 2*tile + 2), __ldg(tile_rk + 2*tile + 3));
      nr0 = b.x; nk0 = b.y; nr1 = b.z; nk1 = b.w;
    }
  };
  // L2 prefetch of a later tile's colind/val window; its bounds are loaded one
  // iteration before they are used, like the staging bounds.
  Index pf0 = 0, pf1 = 0;
  auto fetchPrefetchBounds = [&](int q) {
    pf0 = 0; pf1 = 0;
    if (PF <= 0) return;
    const long long tile = static_cast<long long>(gg) +
                           static_cast<long long>(q)*gstride;
    if (tile < ntiles) {
      pf0 = __ldg(tile_rk + 2*tile + 1);
      pf1 = __ldg(tile_rk + 2*tile + 3);
    }
  };
  auto prefetch = [&]() {
    if (PF <= 0) return;
    const Index a0 = pf0 & ~7;
    Index a1 = (pf1 + 3) & ~3;
    const Index klim = nnz & ~3;
    if (a1 > klim) a1 = klim;
    if (a1 > a0) {
      const uint32_t nb = static_cast<uint32_t>(a1 - a0)*4u;
      bulkPrefetchL2(enc_colind + a0, nb);
      bulkPrefetchL2(val + a0, nb);
    }
  };

  if (t == 0 && gg < ntiles) {
    fetchBounds(0);
    issue(0, nr0, nk0, nr1, nk1);
    for (int q = 1; q <= PF; ++q) { fetchPrefetchBounds(q); prefetch(); }
    fetchBounds(1);
    fetchPrefetchBounds(1 + PF);
  }
  if (HUB_K > 0) mbarWait(bar_hub, 0);
  const uint32_t hub_base = smemAddr(s_raw);

  for (int q = 0; static_cast<long long>(gg) +
                  static_cast<long long>(q)*gstride < ntiles; ++q) {
    const long long tile = static_cast<long long>(gg) +
                           static_cast<long long>(q)*gstride;
    const int b = q & 1;
    mbarWait(bar_full, b);

    const Index r0 = s_meta[4*b + 0];
    const Index r1 = s_meta[4*b + 1];
    const Index k0 = s_meta[4*b + 2];
    const Index k1 = s_meta[4*b + 3];
    const int   nr  = r1 - r0;                 // rows that END in this tile
    const Index k0a = k0 & ~7;
    // rowStart(i) = rp[i], rowEnd(i) = rp[i+1] for local row i in [0, nr]
    const Index* const rp = st_rp + b*GB_HUB_RPWIN + (r0 & 3);

    // ---- this thread's 8 nonzeros: staged words -> registers ------------------
    // Two 128-bit shared loads per array; lanes 4..7 of every eight take the
    // upper half first so that a quarter warp covers 32 distinct banks.
    const Index kbase = k0a + 8*t;
    Index ci[8];
    int   vb[8];
    {
      const int half = (t >> 2) & 1;
      const int4 c_first  = *reinterpret_cast<const int4*>(st_ci + 8*t + 4*half);
      const int4 c_second = *reinterpret_cast<const int4*>(st_ci + 8*t + 4*(half ^ 1));
      const int4 v_first  = *reinterpret_cast<const int4*>(
          reinterpret_cast<const int*>(st_va) + 8*t + 4*half);
      const int4 v_second = *reinterpret_cast<const int4*>(
          reinterpret_cast<const int*>(st_va) + 8*t + 4*(half ^ 1));
      const int4 c_lo = half ? c_second : c_first;
      const int4 c_hi = half ? c_first : c_second;
      const int4 v_lo = half ? v_second : v_first;
      const int4 v_hi = half ? v_first : v_second;
      ci[0] = c_lo.x; ci[1] = c_lo.y; ci[2] = c_lo.z; ci[3] = c_lo.w;
      ci[4] = c_hi.x; ci[5] = c_hi.y; ci[6] = c_hi.z; ci[7] = c_hi.w;
      vb[0] = v_lo.x; vb[1] = v_lo.y; vb[2] = v_lo.z; vb[3] = v_lo.w;
      vb[4] = v_hi.x; vb[5] = v_hi.y; vb[6] = v_hi.z; vb[7] = v_hi.w;
    }
    // Gathers: cold columns from global memory (long latency, issued first), hub
    // slots from shared memory.  Positions outside [k0, k1) hold valid codes of an
    // earlier tile: gathered and never used.
    U uv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (ci[j] >= 0) uv[j] = ldGatherCold(u + ci[j], pol_keep);
    // The colind/val window is in registers now: hand it back for the next tile.
    groupBarrier(1 + g, GB_HUB_GT);
    if (t == 0) {
      const long long next = tile + gstride;
      if (next < ntiles) issue(q + 1, nr0, nk0, nr1, nk1);
      prefetch();                              // tile q + 1 + PF
      fetchBounds(q + 2);
      fetchPrefetchBounds(q + 2 + PF);
    }
    if (HUB_K > 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (ci[j] < 0) {
          int bits;
          asm volatile("ld.shared.b32 %0, [%1];"
                       : "=r"(bits)
                       : "r"(hub_base + (static_cast<uint32_t>(ci[j]) << 2)));
          memcpy(&uv[j], &bits, 4);
        }
      }
    }

    // ---- rows of this tile without a nonzero in it: write identity -------------
    for (int i = t; i < nr; i += GB_HUB_GT) {
      Index rs = rp[i];
      const Index re = rp[i + 1];
      if (rs < k0) rs = k0;
      if (rs == re) w[r0 + i] = identity;
    }

    // ---- this thread's range [ka, kz) and its first row ---------------------------
    Index ka = kbase < k0 ? k0 : kbase;
    if (ka > k1) ka = k1;
    Index kz = kbase + 8 < k1 ? kbase + 8 : k1;       // exclusive end
    if (kz < ka) kz = ka;
    int i;
    {
      int lo = 0, hi = nr;                     // smallest i with rowEnd(i) > ka
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (rp[mid + 1] <= ka) lo = mid + 1; else hi = mid;
      }
      i = lo;
    }
    const int first_i = i;
    W acc  = identity;
    W head = identity;
    Index re = rp[i + 1];
    if (kz > ka) {
      const int jlo = ka - kbase;
      const int jhi = kz - kbase;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a av;
        memcpy(&av, &vb[j], 4);
        const W prod = mul_op(av, uv[j]);
        if (j >= jlo && j < jhi) {
          if (kbase + j >= re) {               // row i ended before this entry
            if (i == first_i) head = acc; else w[r0 + i] = acc;
            acc = identity;
            ++i;
            re = rp[i + 1];
            if (kbase + j >= re) {             // empty rows follow: search
              int lo = i + 1, hi = nr;
              while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (rp[mid + 1] <= kbase + j) lo = mid + 1; else hi = mid;
              }
              i = lo;
              re = rp[i + 1];
            }
          }
          acc = add_op(acc, prod);
        }
      }
      // the row ends exactly where this range ends: it is complete here
      if (i < nr && kz >= re) {
        if (i == first_i) head = acc; else w[r0 + i] = acc;
        acc = identity;
        int lo = i + 1, hi = nr;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (rp[mid + 1] <= kz) lo = mid + 1; else hi = mid;
        }
        i = lo;
      }
    }

    // ---- segmented scan of (open row, partial) over the group ---------------------
    Index key = i;
    W     v   = acc;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const Index pk = __shfl_up_sync(GB_FULL_MASK, key, off);
      const W     pv = __shfl_up_sync(GB_FULL_MASK, v, off);
      if (lane >= off && pk == key) v = add_op(pv, v);
    }
    Index* const wkey = s_wkey + 4*b;
    W*     const wval = s_wval + 4*b;
    if (lane == 31) { wkey[wid] = key; wval[wid] = v; }
    const Index key0 = __shfl_sync(GB_FULL_MASK, key, 0);
    const Index ekey = __shfl_up_sync(GB_FULL_MASK, key, 1);
    const W     eval = __shfl_up_sync(GB_FULL_MASK, v, 1);
    groupBarrier(1 + g, GB_HUB_GT);

    Index ck = -1;
    W     cv = identity;
#pragma unroll
    for (int ww = 0; ww < GB_HUB_GT/32 - 1; ++ww) {
      if (ww < wid) {
        const Index wk = wkey[ww];
        const W     wv = wval[ww];
        if (wk == ck) cv = add_op(cv, wv);
        else { ck = wk; cv = wv; }
      }
    }
    W carry_in;
    if (lane == 0) {
      carry_in = (ck == first_i) ? cv : identity;
    } else {
      carry_in = eval;                         // ekey == first_i always
      if (ekey == key0 && ck == ekey) carry_in = add_op(cv, eval);
    }
    if (i > first_i) w[r0 + first_i] = add_op(carry_in, head);
    if (t == GB_HUB_GT - 1) {
      const W out = (i > first_i) ? acc : add_op(carry_in, acc);
      carry_row[tile] = (r0 + i < nrows) ? (r0 + i) : -1;
      carry_val[tile] = out;
    }
  }
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_KERNELS_SPMV_HUB_CUH_
