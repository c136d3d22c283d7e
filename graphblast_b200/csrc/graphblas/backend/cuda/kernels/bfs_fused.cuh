// graphblast_b200 backend — the whole direction-optimised BFS as ONE persistent
// cooperative kernel (loop fusion behind the same API).
//
// The operation sequence of reference graphblas/algorithm/bfs.hpp:46-79 per level —
//   assign(v<f> = level); vxm(f' <!v> = f (||.&&) A); swap; succ = reduce(+, f') —
// with --struconly 1 --opreuse 1 --earlyexit 1 --fusedmask 1 costs 2 launches on a
// pull level and 6 on a push level through the generic operations, plus a host
// read of the frontier size; a traversal of R-MAT scale 24 (6 levels, 0.47 ms) is
// then half launch latency.  Here the level loop, the direction decision, the
// level assignment and the frontier count all live in the kernel; levels are
// separated by grid-wide barriers (cooperative launch, one CTA set resident for
// the whole traversal).
//
// State: visited bitmap (two copies, ping-pong on pull levels), frontier bitmap F,
// next-frontier bitmap N, float levels v (the result, 1-based, 0 = unreached).
//   push level (frontier small): warps scan F; a vertex of moderate degree is
//     expanded by its warp, lanes striding the adjacency; vertices with more than
//     GB_BFS_HEAVY neighbours go to a list that the WHOLE grid expands after a
//     barrier (an R-MAT source has 10^5..10^6 neighbours).  Discoveries set the
//     visited bit with atomicOr immediately (same level either way), the winner
//     writes v and the N bit.
//   pull level (frontier large): the fused Boolean pull of kernels/spmv_pull.cuh —
//     a warp owns 4 words of the bitmap per iteration, the first-neighbour summary
//     decides most rows, the rest walk their list with early exit, probing the
//     visited bitmap AS OF THE LEVEL'S START (operand reuse, reference
//     kernels/spmv.hpp:35-41) — and the owner of a word writes N, the merged
//     visited word of the other copy, v for the discovered rows and clears F.
// Direction: the reference's ratio rule with hysteresis (vector.hpp:318-342):
// sparse -> dense when |f|/n > switchpoint and growing, dense -> sparse when
// <= switchpoint and shrinking; results do not depend on it.
#ifndef GRAPHBLAS_BACKEND_CUDA_KERNELS_BFS_FUSED_CUH_
#define GRAPHBLAS_BACKEND_CUDA_KERNELS_BFS_FUSED_CUH_

#include <cooperative_groups.h>

#include "graphblas/backend/cuda/kernels/common.cuh"

namespace graphblas {
namespace backend {

// CTA shape, measured on RMAT-24: 768 x 2 per SM 0.275 ms, 512 x 3 0.281, 1024 x 2
// 0.305 (32 registers: the pull loop spills), 1024 x 1 0.297 (no spills, half the
// warps: the first pull level is latency-bound and takes 40 % longer).
#ifndef GB_BFS_NT
#define GB_BFS_NT     768
#endif
#ifndef GB_BFS_MINB
#define GB_BFS_MINB   2               // resident CTAs per SM the register budget allows
#endif
#define GB_BFS_HEAVY  2048            // adjacency longer than this: grid-wide expansion
#define GB_BFS_HEAVY_CAP 4096         // heavy vertices per level kept in the list
#define GB_BFS_PARK   4096            // rows a CTA parks for its walking phase per pull level

struct BfsFusedArgs {
  // structure: rows to expand when pushing, rows to inspect when pulling
  const Index* push_ptr;   const Index* push_ind;     // out-neighbours of a vertex
  const Index* pull_ptr;   const Index* pull_ind;     // in-neighbours of a vertex
  const Index* pull_first;                            // first-neighbour summary of pull_*
  const unsigned int* pull_empty;                     // bitmap of rows without in-neighbours
  Index n;
  Index source;
  int   max_levels;
  float switchpoint;
  int   mode;                // 0 push-pull, 1 push only, 2 pull only (reference --mxvmode)
  // state (device memory, sized for n)
  float*        levels;      // result
  unsigned int* visited[2];
  unsigned int* frontier;    // F
  unsigned int* next;        // N
  // small cells
  unsigned long long* counters;   // [0..2] next-frontier size, [3..5] heavy-list length
                                  // (both rotate with level % 3: the cell of level L
                                  // is zeroed during level L-1, filled during L and
                                  // read after L's barrier, when slow threads may still
                                  // be reading the cell of L-1),
                                  // [6] levels executed, [7..11] work counters (out),
                                  // [12..27] time at the end of the set-up and of every
                                  // level (ns << 1 | pulled), for GB200_BFS_TRACE
  Index*        heavy;            // [GB_BFS_HEAVY_CAP]
};

__device__ __forceinline__ unsigned long long bfsClockNs() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ bool bfsClaim(unsigned int* visited, Index vtx) {
  const unsigned int bit = 1u << (vtx & 31);
  unsigned int* word = visited + (vtx >> 5);
  if (*reinterpret_cast<volatile unsigned int*>(word) & bit) return false;
  return (atomicOr(word, bit) & bit) == 0;
}

// MINB: CTAs per SM the register allocation must allow (the pull levels are
// latency-bound: occupancy matters more than a few spilled pointers).
template <int MINB>
__global__ void __launch_bounds__(GB_BFS_NT, MINB)
bfsFusedKernel(BfsFusedArgs a) {
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  __shared__ int s_red[GB_BFS_NT/32];
  __shared__ int s_parked;                 // rows waiting in s_park (pull levels)
  __shared__ Index s_park[GB_BFS_PARK];
  if (threadIdx.x == 0) s_parked = 0;

  const Index n = a.n;
  const Index nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const Index gtid = blockIdx.x*blockDim.x + threadIdx.x;        // grid << 2^31 threads
  const Index gthreads = gridDim.x*blockDim.x;
  const Index gwarp = gtid >> 5;
  const Index gwarps = gthreads >> 5;

  if (gtid == 0) a.counters[28] = bfsClockNs();
  // ---- level 0: clear the state, seed the source ---------------------------------
  for (Index i = gtid; i < n; i += gthreads)
    a.levels[i] = (i == a.source) ? 1.f : 0.f;
  for (Index w = gtid; w < nwords; w += gthreads) {
    const unsigned int seed = (w == (a.source >> 5)) ? (1u << (a.source & 31)) : 0u;
    // rows nothing points at count as visited from the start: no level can discover
    // them, and the pull levels would look at them every time
    a.visited[0][w] = seed | a.pull_empty[w];
    a.visited[1][w] = 0u; a.frontier[w] = seed; a.next[w] = 0u;
  }
  if (gtid < 12) a.counters[gtid] = 0ull;
  grid.sync();
  if (gtid == 0) a.counters[12] = bfsClockNs() << 1;          // [12..27]: level clock

  unsigned int* vis = a.visited[0];       // visited as of the level's start
  unsigned int* vis_other = a.visited[1];
  unsigned int* F = a.frontier;
  unsigned int* N = a.next;
  unsigned long long fcount = 1ull;
  bool dense = (a.mode == 2);             // direction state (storage of the frontier)
  float prev_ratio = 0.f;
  int inspected = 0;                      // colind entries looked at by this thread
  int pushed_vertices = 0;                // frontier entries expanded (lane 0 counts)
  long long pushed_edges = 0;             // their adjacency lengths
  int discovered_pushing = 0;
  int pull_levels = 0;
  int level = 1;

  for (; level <= a.max_levels && fcount > 0ull; ++level) {
    // direction for this level (reference Vector::convert)
    if (a.mode == 0) {
      const float ratio = static_cast<float>(fcount)/static_cast<float>(n);
      if (!dense) {
        if (ratio > a.switchpoint && ratio > prev_ratio) dense = true; else prev_ratio = ratio;
      } else {
        if (ratio <= a.switchpoint && ratio < prev_ratio) dense = false; else prev_ratio = ratio;
      }
    }
    unsigned long long* const count_cell = a.counters + (level % 3);
    unsigned long long* const heavy_cell = a.counters + 3 + (level % 3);
    if (gtid == 0) {                        // next level's cells
      a.counters[(level + 1) % 3] = 0ull;
      a.counters[3 + (level + 1) % 3] = 0ull;
    }
    const float next_level = static_cast<float>(level + 1);
    int found_here = 0;

    if (!dense) {
      // ---------------- push: expand the frontier --------------------------------
      // A warp reads 32 frontier words at once (the frontier is sparse here: most
      // words are zero and a word-at-a-time scan is a chain of dependent loads),
      // then walks the non-empty ones.
      for (Index w0 = gwarp*32; w0 < nwords; w0 += gwarps*32) {
        const Index mine = w0 + lane;
        unsigned int my_bits = (mine < nwords) ? F[mine] : 0u;
        if (my_bits != 0u) F[mine] = 0u;          // this buffer is the next level's N
        unsigned int pending = __ballot_sync(GB_FULL_MASK, my_bits != 0u);
        while (pending != 0u) {
          const int src_lane = __ffs(pending) - 1;
          pending &= pending - 1u;
          unsigned int bits = __shfl_sync(GB_FULL_MASK, my_bits, src_lane);
          const Index w = w0 + src_lane;
          while (bits != 0u) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1u;
            const Index u = w*32 + b;
            const Index beg = __ldg(a.push_ptr + u);
            const Index deg = __ldg(a.push_ptr + u + 1) - beg;
            if (lane == 0) { ++pushed_vertices; pushed_edges += deg; }
            if (deg > GB_BFS_HEAVY) {
              // the whole grid expands it after the barrier; a full list falls back
              // to this warp (slow, still correct)
              unsigned long long slot = 0ull;
              if (lane == 0) slot = atomicAdd(heavy_cell, 1ull);
              slot = __shfl_sync(GB_FULL_MASK, slot, 0);
              if (slot < GB_BFS_HEAVY_CAP) {
                if (lane == 0) a.heavy[slot] = u;
                continue;
              }
            }
            for (Index k = lane; k < deg; k += 32) {
              const Index nbr = __ldg(a.push_ind + beg + k);
              if (bfsClaim(vis, nbr)) {
                a.levels[nbr] = next_level;
                atomicOr(N + (nbr >> 5), 1u << (nbr & 31));
                ++found_here;
              }
            }
          }
        }
      }
      grid.sync();
      unsigned long long nheavy = *reinterpret_cast<volatile unsigned long long*>(heavy_cell);
      if (nheavy > GB_BFS_HEAVY_CAP) nheavy = GB_BFS_HEAVY_CAP;
      if (nheavy > 0ull) {
        for (unsigned long long h = 0; h < nheavy; ++h) {
          const Index u = a.heavy[h];
          const Index beg = __ldg(a.push_ptr + u);
          const Index deg = __ldg(a.push_ptr + u + 1) - beg;
          for (Index k = gtid; k < deg; k += gthreads) {
            const Index nbr = __ldg(a.push_ind + beg + k);
            if (bfsClaim(vis, nbr)) {
              a.levels[nbr] = next_level;
              atomicOr(N + (nbr >> 5), 1u << (nbr & 31));
              ++found_here;
            }
          }
        }
      }
      discovered_pushing += found_here;
    } else {
      ++pull_levels;
      // ---------------- pull: every unvisited row looks for a visited neighbour ----
      const Index ngroups = (nwords + 3) >> 2;
      for (Index g = gwarp; g < ngroups; g += gwarps) {
        unsigned int mword[4];
        Index f[4];
        unsigned int pword[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const Index word = g*4 + j;
          mword[j] = (word < nwords) ? vis[word] : 0xffffffffu;
        }
        if ((mword[0] & mword[1] & mword[2] & mword[3]) == 0xffffffffu) {
          // nothing left to discover in these 128 rows (the common case after the
          // first pull level): only the bitmaps move on
          if (lane < 4 && g*4 + lane < nwords) {
            N[g*4 + lane] = 0u;
            vis_other[g*4 + lane] = 0xffffffffu;
            F[g*4 + lane] = 0u;
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const Index row = (g*4 + j)*32 + lane;
          const bool open = (row < n) && !((mword[j] >> lane) & 1u);
          f[j] = open ? __ldg(a.pull_first + row) : static_cast<Index>(-1);
          if (open) ++inspected;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pword[j] = 0u;
          if (f[j] != static_cast<Index>(-1)) pword[j] = vis[(f[j] & 0x7fffffff) >> 5];
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const Index word = g*4 + j;
          const Index row = word*32 + lane;
          bool found = (pword[j] >> (f[j] & 31)) & 1u;
          // more entries and the first one not visited: the row has to be walked.
          // Few lanes of a warp are in that position and a walk is a chain of
          // dependent loads, so the rows are parked in a CTA-wide list and walked by
          // all threads after the scan (inline only when the list is full).
          bool walk = (f[j] >= 0 && !found);
          const unsigned int walkers = __ballot_sync(GB_FULL_MASK, walk);
          if (walkers != 0u) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_parked, __popc(walkers));
            base = __shfl_sync(GB_FULL_MASK, base, 0);
            const int slot = base + __popc(walkers & ((1u << lane) - 1u));
            if (walk && slot < GB_BFS_PARK) { s_park[slot] = row; walk = false; }
          }
          if (walk) {
            Index k = __ldg(a.pull_ptr + row) + 1;
            const Index end = __ldg(a.pull_ptr + row + 1);
            for (; k < end; ++k) {
              const Index col = __ldg(a.pull_ind + k);
              ++inspected;
              if ((vis[col >> 5] >> (col & 31)) & 1u) { found = true; break; }
            }
          }
          const unsigned int out = __ballot_sync(GB_FULL_MASK, found);
          if (word < nwords) {
            if (found) a.levels[row] = next_level;
            if (lane == 0) {
              N[word] = out;
              vis_other[word] = mword[j] | out;
              F[word] = 0u;
            }
          }
          found_here += found ? 1 : 0;
        }
      }
      // the parked rows, one per thread at a time
      __syncthreads();
      if (gtid == 0 && level == 2) { a.counters[29] = bfsClockNs(); a.counters[31] = s_parked; }
      const int parked = (s_parked < GB_BFS_PARK) ? s_parked : GB_BFS_PARK;
      for (int i = threadIdx.x; i < parked; i += GB_BFS_NT) {
        const Index row = s_park[i];
        Index k = __ldg(a.pull_ptr + row) + 1;
        const Index end = __ldg(a.pull_ptr + row + 1);
        bool found = false;
        // four entries per step: their bitmap words are fetched together, then
        // examined in list order (the count stops at the first visited one)
        while (k < end && !found) {
          Index col[4];
          unsigned int word[4];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            col[u] = (k + u < end) ? __ldg(a.pull_ind + k + u) : static_cast<Index>(-1);
#pragma unroll
          for (int u = 0; u < 4; ++u)
            word[u] = (col[u] >= 0) ? vis[col[u] >> 5] : 0u;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (found || col[u] < 0) continue;
            ++inspected;
            found = (word[u] >> (col[u] & 31)) & 1u;
          }
          k += 4;
        }
        if (found) {
          const unsigned int bit = 1u << (row & 31);
          atomicOr(N + (row >> 5), bit);
          atomicOr(vis_other + (row >> 5), bit);
          a.levels[row] = next_level;
          ++found_here;
        }
      }
      __syncthreads();
      if (gtid == 0 && level == 2) a.counters[30] = bfsClockNs();
      if (threadIdx.x == 0) s_parked = 0;
    }
    // ---- frontier size of the next level ------------------------------------------
    const int block_found = blockSum<GB_BFS_NT>(found_here, s_red);
    if (threadIdx.x == 0 && block_found)
      atomicAdd(count_cell, static_cast<unsigned long long>(block_found));
    grid.sync();
    if (gtid == 0 && level < 16)
      a.counters[12 + level] = (bfsClockNs() << 1) | (dense ? 1ull : 0ull);
    fcount = *reinterpret_cast<volatile unsigned long long*>(count_cell);
    if (dense) { unsigned int* t = vis; vis = vis_other; vis_other = t; }
    { unsigned int* t = F; F = N; N = t; }
  }
  // ---- results -----------------------------------------------------------------
  // [7] entries inspected pulling, [8] pull levels, [9] vertices pushed, [10] edges
  // pushed, [11] vertices discovered pushing — the algorithmic bytes of SURVEY.md §8d
  const int block_insp = blockSum<GB_BFS_NT>(inspected, s_red);
  const int block_pv = blockSum<GB_BFS_NT>(pushed_vertices, s_red);
  const int block_dp = blockSum<GB_BFS_NT>(discovered_pushing, s_red);
  if (threadIdx.x == 0) {
    if (block_insp) atomicAdd(a.counters + 7, static_cast<unsigned long long>(block_insp));
    if (block_pv)   atomicAdd(a.counters + 9, static_cast<unsigned long long>(block_pv));
    if (block_dp)   atomicAdd(a.counters + 11, static_cast<unsigned long long>(block_dp));
  }
  if (pushed_edges) atomicAdd(a.counters + 10, static_cast<unsigned long long>(pushed_edges));
  if (gtid == 0) {
    a.counters[6] = static_cast<unsigned long long>(level - 1);
    a.counters[8] = static_cast<unsigned long long>(pull_levels);
  }
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_KERNELS_BFS_FUSED_CUH_
