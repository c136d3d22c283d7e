// graphblast_b200 backend — masked SpGEMM, hash formulation (triangle counting).
//
// C(i,j) = add_k mul(A(i,k), B(k,j)) for (i,j) in the mask is the size (or the
// semiring sum) of the intersection of two sorted index lists: row i of A and
// column j of B.  The search formulation (spgemm_masked.cuh; the reference's
// kernels/spgemm.hpp:17-79 is its one-warp-per-row ancestor) walks the shorter list
// and binary-searches the longer one in global memory: min(|a|,|b|) * log max(|a|,|b|)
// dependent probes per mask entry, 2.3e10 at RMAT-20.
//
// Here every mask entry is charged to the OWNER of its longer list:
//   * pass 1, owner = row i of A (|B(:,j)| <= |A(i,:)|): the owner's list goes into a
//     shared-memory hash table (key -> value), then every partner column j of mask
//     row i streams ITS (shorter) list through the table;
//   * pass 2, owner = column j of B (|A(i,:)| < |B(:,j)|): same with the roles
//     swapped; the partners are the rows of mask COLUMN j (mask CSC), and the slot
//     of (i,j) in the CSR-ordered output is found by one search of mask row i.
// Work per entry: min(|a|,|b|) shared-memory probes, all global loads sequential
// runs — 2.3e9 probes at RMAT-20 and none of them a dependent global load.
//
// Owners are binned by list length (classify kernel): up to 64 keys a warp owns the
// table (no CTA barriers), up to 1024 a 256-thread CTA, beyond that a 1024-thread
// CTA with a 128 KB table that holds 8192 keys at a time — longer lists go through
// the table in segments, every partner list streamed once per segment and the
// partial results combined in C.
// A work item is (owner, a chunk of its partners): a hub column of an R-MAT has
// 10^5 partners and a short list of its own, and one group walking all of them
// was the whole run time of the first version; the table is rebuilt per item,
// which costs one pass over a list that is short next to what streams through it.
// Inside a CTA the warps take partners from a shared list one at a time (partner
// lists of one owner differ by three orders of magnitude; with a fixed assignment
// two thirds of the issue slots were barrier waits), and the partners are described
// by all threads at once before any of them is streamed (three dependent loads
// and, in pass 2, a search per partner — serialised in front of every few lists
// they cost as much as the streaming).
#ifndef GRAPHBLAS_BACKEND_CUDA_KERNELS_SPGEMM_HASH_CUH_
#define GRAPHBLAS_BACKEND_CUDA_KERNELS_SPGEMM_HASH_CUH_

#include "graphblas/backend/cuda/kernels/common.cuh"

namespace graphblas {
namespace backend {

#define GB_HASH_EMPTY   (-1)
#define GB_HASH_CAP_S   64       // warp-owned tables: 256 slots, load <= 0.25
#ifndef GB_HASH_CAP_M
#define GB_HASH_CAP_M   1024     // CTA-owned tables
#endif
#define GB_HASH_SLOTS_S 256
#ifndef GB_HASH_SLOTS_M
#define GB_HASH_SLOTS_M 2048     // load <= 0.5: a probe reads four slots, chains stay short
#endif
#ifndef GB_HASH_SLOTS_L
#define GB_HASH_SLOTS_L 16384    // big-CTA tables
#endif
#ifndef GB_HASH_SEG_L
#define GB_HASH_SEG_L   8192     // keys of a long list that go into the table at a time
#endif
#define GB_HASH_NCLASS  3        // S, M, L
#define GB_HASH_CHUNK_S 256      // partners per work item
#define GB_HASH_CHUNK_M 1024
#ifndef GB_HASH_CHUNK_L
#define GB_HASH_CHUNK_L 2048
#endif
#ifndef GB_HASH_UNROLL_M
#define GB_HASH_UNROLL_M 4       // keys of a partner list in flight per lane
#endif
#ifndef GB_HASH_UNROLL_L
#define GB_HASH_UNROLL_L 8
#endif
#ifndef GB_HASH_CTAS_M
#define GB_HASH_CTAS_M 7         // resident CTAs per SM the launches are sized for
#endif
#ifndef GB_HASH_CTAS_L
#define GB_HASH_CTAS_L 1
#endif

struct HashItem { Index owner; Index first_partner; };   // index into the M_* arrays

// Bins the owners that have at least one partner by the length of their list and
// cuts their partner ranges into work items.  lists: GB_HASH_NCLASS arrays of
// `stride` items; counts[c] their fill.
__global__ void spgemmHashClassifyKernel(const Index* __restrict__ own_ptr,
                                         const Index* __restrict__ partner_ptr,
                                         Index                     nowners,
                                         bool                      skip_empty,
                                         HashItem* __restrict__    lists,
                                         size_t                    stride,
                                         unsigned int*             counts) {
  const int lane = threadIdx.x & 31;
  Index v = blockIdx.x*blockDim.x + threadIdx.x;
  const Index step = gridDim.x*blockDim.x;
  // whole warps iterate together (the shuffles below need every lane)
  for (Index base = v - lane; base < nowners; base += step) {
    v = base + lane;
    int cls = -1;
    Index p_beg = 0, p_end = 0;
    if (v < nowners) {
      p_beg = __ldg(partner_ptr + v);
      p_end = __ldg(partner_ptr + v + 1);
    }
    if (p_end > p_beg) {
      const Index len = __ldg(own_ptr + v + 1) - __ldg(own_ptr + v);
      if (len == 0 && skip_empty) cls = -1;
      else if (len <= GB_HASH_CAP_S) cls = 0;
      else if (len <= GB_HASH_CAP_M) cls = 1;
      else cls = 2;
    }
#pragma unroll
    for (int k = 0; k < GB_HASH_NCLASS; ++k) {
      const Index chunk = k == 0 ? GB_HASH_CHUNK_S : k == 1 ? GB_HASH_CHUNK_M
                                                           : GB_HASH_CHUNK_L;
      if (__ballot_sync(GB_FULL_MASK, cls == k) == 0) continue;
      unsigned int mine = 0;
      if (cls == k) mine = (p_end - p_beg + chunk - 1)/chunk;
      unsigned int before = mine;                   // inclusive warp scan
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const unsigned int up = __shfl_up_sync(GB_FULL_MASK, before, d);
        if (lane >= d) before += up;
      }
      unsigned int first = 0;
      if (lane == 31) first = atomicAdd(counts + k, before);
      first = __shfl_sync(GB_FULL_MASK, first, 31) + before - mine;
      if (cls == k) {
        HashItem* out = lists + k*stride + first;
        for (Index q = p_beg; q < p_end; q += chunk, ++out) {
          out->owner = v;
          out->first_partner = q;
        }
      }
    }
  }
}

__device__ __forceinline__ unsigned int hashSlot(Index key, int shift) {
  return (static_cast<unsigned int>(key)*0x9E3779B1u) >> shift;
}

// Shared memory of one group: the table and the item's partner list.
template <int SLOTS, int CHUNK, typename TV>
struct __align__(16) HashGroupSmem {
  Index keys[SLOTS];          // buckets of four slots, filled from the left
  TV    vals[SLOTS];
  Index list_beg[CHUNK];      // partners that stream: first entry, length, output slot
  Index list_len[CHUNK];
  Index list_out[CHUNK];
  unsigned int list_size;
  unsigned int next;          // next partner of the list to be taken by a warp
};

// One group (a warp when WARP_OWNER, else the CTA) per work item.
//   T_* : the owners' lists (table side)      P_* : the partners' lists (streamed)
//   M_* : mask adjacency by owner — CSR rows in pass 1, CSC columns in pass 2
//   SWAP: pass 2; the table side is B, so products are mul(P value, T value), and
//         the output slot is looked up in the CSR-ordered mask.
// Per item: (A) all threads describe the item's partners — list bounds, whether this
// owner is in charge of the pair, the output slot — and append the ones that stream
// to a shared list; (B) the owner's list goes into the table; (C) warps take
// partners from the list one at a time and stream them through the table.
template <int CT, bool WARP_OWNER, int SLOTS, int SEG, int CHUNK, int UNROLL, bool SWAP,
          typename c, typename TV, typename PV, typename m,
          typename MulOp, typename AddOp>
__global__ void __launch_bounds__(CT)
spgemmHashKernel(c* __restrict__             C_val,
                 const HashItem* __restrict__ items,
                 const unsigned int*         item_count,
                 unsigned int*               grab,
                 const Index* __restrict__   T_ptr,
                 const Index* __restrict__   T_ind,
                 const TV* __restrict__      T_val,
                 const Index* __restrict__   P_ptr,
                 const Index* __restrict__   P_ind,
                 const PV* __restrict__      P_val,
                 const Index* __restrict__   M_ptr,
                 const Index* __restrict__   M_ind,
                 const m* __restrict__       M_val,
                 const Index* __restrict__   mask_rowptr,
                 const Index* __restrict__   mask_colind,
                 MulOp                       mul_op,
                 AddOp                       add_op,
                 c                           identity,
                 unsigned long long*         list_bytes) {
  constexpr int GT = WARP_OWNER ? 32 : CT;              // threads per group
  typedef HashGroupSmem<SLOTS, CHUNK, TV> Smem;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned int grabbed;

  const int lane = threadIdx.x & 31;
  const int gtid = WARP_OWNER ? lane : threadIdx.x;
  Smem& sm = reinterpret_cast<Smem*>(smem_raw)[WARP_OWNER ? (threadIdx.x >> 5) : 0];
  const unsigned int nitems = *item_count;
  unsigned long long scanned = 0;

  while (true) {
    unsigned int idx;
    if (WARP_OWNER) {
      __syncwarp();
      idx = 0;
      if (lane == 0) { idx = atomicAdd(grab, 1u); sm.list_size = 0; }
      idx = __shfl_sync(GB_FULL_MASK, idx, 0);
    } else {
      __syncthreads();                      // previous item's table and list are done with
      if (threadIdx.x == 0) { grabbed = atomicAdd(grab, 1u); sm.list_size = 0; }
      __syncthreads();
      idx = grabbed;
    }
    if (idx >= nitems) break;
    const Index owner = items[idx].owner;
    const Index m_beg = items[idx].first_partner;
    const Index m_stop = __ldg(M_ptr + owner + 1);
    const Index m_end = (m_stop - m_beg > CHUNK) ? m_beg + CHUNK : m_stop;
    const Index t_beg = __ldg(T_ptr + owner);
    const Index t_len = __ldg(T_ptr + owner + 1) - t_beg;
    if (gtid == 0) scanned += t_len;

    // (A) the item's partners
    for (Index q = m_beg + gtid; q < m_end; q += GT) {
      const Index w     = __ldg(M_ind + q);
      const Index p_beg = __ldg(P_ptr + w);
      const Index p_len = __ldg(P_ptr + w + 1) - p_beg;
      const bool chosen = SWAP ? (p_len < t_len) : (p_len <= t_len);
      if (!chosen) continue;                // the partner owns this pair
      const Index out = SWAP ? findSorted(mask_colind, __ldg(mask_rowptr + w),
                                          __ldg(mask_rowptr + w + 1), owner)
                             : q;
      if (p_len > 0 && M_val[q] != 0) {
        const unsigned int at = atomicAdd(&sm.list_size, 1u);
        sm.list_beg[at] = p_beg;
        sm.list_len[at] = p_len;
        sm.list_out[at] = out;
        scanned += p_len;
      } else {
        C_val[out] = identity;
      }
    }

    // the owner's list goes through the table SEG keys at a time (one segment
    // unless it is longer than the largest table holds)
    for (Index seg = 0; seg == 0 || seg < t_len; seg += SEG) {
      const Index seg_len = (t_len - seg > SEG) ? SEG : t_len - seg;
      // table size: power of two >= 4*seg_len (load <= 0.25) where SLOTS allows
      int lg = 5;
      while ((1 << lg) < 4*seg_len && (1 << lg) < SLOTS) ++lg;
      const int nslots = 1 << lg;
      const int shift  = 32 - lg;
      const unsigned int smask = nslots - 1;
      const unsigned int bmask = (nslots >> 2) - 1;
      const int4* buckets = reinterpret_cast<const int4*>(sm.keys);

      // (B)
      if (!WARP_OWNER && seg > 0) __syncthreads();      // previous segment's probes
      for (int s = gtid; s < nslots; s += GT) sm.keys[s] = GB_HASH_EMPTY;
      if (gtid == 0) sm.next = 0;
      if (WARP_OWNER) __syncwarp(); else __syncthreads();
      for (Index p = gtid; p < seg_len; p += GT) {
        const Index key = __ldg(T_ind + t_beg + seg + p);
        unsigned int s = hashSlot(key, shift + 2) << 2;      // first slot of the bucket
        while (atomicCAS(sm.keys + s, GB_HASH_EMPTY, key) != GB_HASH_EMPTY)
          s = (s + 1) & smask;
        sm.vals[s] = T_val[t_beg + seg + p];
      }
      if (WARP_OWNER) __syncwarp(); else __syncthreads();

      // (C) the first 32 keys of the next list are requested before the current
      // one is probed
      const unsigned int nlist = sm.list_size;
      unsigned int cur = 0;
      if (lane == 0) cur = atomicAdd(&sm.next, 1u);
      cur = __shfl_sync(GB_FULL_MASK, cur, 0);
      Index cb = 0, cl = 0, key = GB_HASH_EMPTY;
      if (cur < nlist) {
        cb = sm.list_beg[cur];
        cl = sm.list_len[cur];
        if (lane < cl) key = __ldg(P_ind + cb + lane);
      }
      while (cur < nlist) {
        unsigned int nxt = 0;
        if (lane == 0) nxt = atomicAdd(&sm.next, 1u);
        nxt = __shfl_sync(GB_FULL_MASK, nxt, 0);
        Index nb = 0, nl = 0, next_key = GB_HASH_EMPTY;
        if (nxt < nlist) {
          nb = sm.list_beg[nxt];
          nl = sm.list_len[nxt];
          if (lane < nl) next_key = __ldg(P_ind + nb + lane);
        }
        c acc = identity;
        // UNROLL keys of the list in flight per lane
        for (Index e0 = 0; e0 < cl; e0 += 32*UNROLL) {
          Index k4[UNROLL];
          int   hit[UNROLL];
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            const Index e = e0 + 32*u + lane;
            k4[u] = GB_HASH_EMPTY;
            if (e < cl) k4[u] = (e < 32) ? key : __ldg(P_ind + cb + e);
          }
          // a probe reads a whole bucket of four slots; a bucket fills from the left,
          // so an empty last slot ends an unsuccessful search
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            hit[u] = -1;
            if (k4[u] == GB_HASH_EMPTY) continue;
            unsigned int bkt = hashSlot(k4[u], shift + 2);
            while (true) {
              const int4 w = buckets[bkt];
              if (w.x == k4[u] || w.y == k4[u] || w.z == k4[u] || w.w == k4[u]) {
                hit[u] = 4*bkt + ((w.x == k4[u]) ? 0 : (w.y == k4[u]) ? 1
                                                    : (w.z == k4[u]) ? 2 : 3);
                break;
              }
              if (w.w == GB_HASH_EMPTY) break;
              bkt = (bkt + 1) & bmask;
            }
          }
          // the values of the hits are requested together
          PV pv[UNROLL];
#pragma unroll
          for (int u = 0; u < UNROLL; ++u)
            if (hit[u] >= 0) pv[u] = P_val[cb + e0 + 32*u + lane];
#pragma unroll
          for (int u = 0; u < UNROLL; ++u) {
            if (hit[u] < 0) continue;
            if (SWAP) acc = add_op(mul_op(pv[u], sm.vals[hit[u]]), acc);
            else      acc = add_op(mul_op(sm.vals[hit[u]], pv[u]), acc);
          }
        }
        acc = warpReduce(acc, add_op);
        if (lane == 0) {
          const Index out = sm.list_out[cur];
          C_val[out] = (seg == 0) ? acc : add_op(C_val[out], acc);
        }
        cur = nxt; cb = nb; cl = nl; key = next_key;
      }
    }
  }
  if (list_bytes != NULL) {
    for (int d = 16; d > 0; d >>= 1)
      scanned += __shfl_down_sync(GB_FULL_MASK, scanned, d);
    if (lane == 0 && scanned) atomicAdd(list_bytes, 4ull*scanned);
  }
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_KERNELS_SPGEMM_HASH_CUH_
