// graphblast_b200 backend — PUSH direction kernel (sparse frontier).
//
// The reference push is: lengths -> scan -> expand -> gather colind/val ->
// multiply -> radix sort of E_f keys -> reduce-by-key (reference
// spmspv_inner.hpp:62-320), i.e. the E_f neighbour list is materialised and
// moved ~4 more times by the sort.  Here the neighbour list is never
// materialised: one load-balanced kernel walks the E_f edges of the frontier
// (equal edge share per CTA, found by searching the scanned frontier degrees)
// and combines each product straight into a dense accumulator cell with the
// semiring's add (atomic), marking the cell in a touched-bitmap; the ordered
// bitmap compaction (kernels/compact.cuh) then yields the sorted, duplicate-free
// sparse output the reference produces with its sort.
//
// Algorithmic bytes per launch:
//   12|f| (ind + rowptr pair) + 4 E_f colind [+ 4 E_f val] + 4 E_f mask lookup
//   + 8|f'| written by the compaction.
#ifndef GRAPHBLAS_BACKEND_CUDA_KERNELS_SPMSPV_PUSH_CUH_
#define GRAPHBLAS_BACKEND_CUDA_KERNELS_SPMSPV_PUSH_CUH_

#include "graphblas/backend/cuda/kernels/common.cuh"

namespace graphblas {
namespace backend {

#define GB_PUSH_NT   256
#define GB_PUSH_EPT  8                              // edges per thread per tile
#define GB_PUSH_TILE (GB_PUSH_NT*GB_PUSH_EPT)
#define GB_PUSH_SEG  (GB_PUSH_TILE + 2)                // frontier entries staged per tile

// MaskMode: 0 = no mask, 1 = keep where mask != 0, 2 = keep where mask == 0.
// (The reference applies the mask after its reduce-by-key by overwriting masked
//  entries with a sentinel and compacting, spmspv.hpp:111-243; filtering each
//  edge before the combine gives the same surviving set.)
//
// StructOnly: values are implied (reference --struconly 1): only the bitmap is
// written.  Otherwise prod = mul(A(k), u_val[i]) with the reference's identity
// short-circuit (kernels/ewisemult.hpp:22-25 as used at spmspv_inner.hpp:204).
template <bool StructOnly, int MaskMode,
          typename W, typename a, typename U, typename M,
          typename MulOp, typename AddOp>
__global__ void __launch_bounds__(GB_PUSH_NT)
spmspvPushKernel(unsigned int* __restrict__ bits,
                 W* __restrict__            acc,
                 const M* __restrict__      mask,
                 const unsigned int* __restrict__ mask_bits,  // or NULL
                 const Index* __restrict__  offs,      // nf+1 scanned degrees
                 const Index* __restrict__  f_ind,
                 const U* __restrict__      f_val,
                 Index                      nf,
                 const Index* __restrict__  rowptr,
                 const Index* __restrict__  colind,
                 const a* __restrict__      val,
                 W                          identity,
                 MulOp                      mul_op,
                 AddOp                      add_op,
                 int                        add_kind,  // add_op(3, 5)
                 unsigned long long*        edge_bytes) {
  __shared__ Index s_offs[GB_PUSH_SEG + 1];
  __shared__ Index s_base[GB_PUSH_SEG];
  __shared__ U     s_uval[GB_PUSH_SEG];
  __shared__ Index s_range[2];

  const Index total = offs[nf];
  const long long ntiles = (static_cast<long long>(total) + GB_PUSH_TILE - 1)
                           / GB_PUSH_TILE;
  // Algorithmic bytes per expanded edge: colind (+ val) (+ mask lookup).
  if (blockIdx.x == 0 && threadIdx.x == 0 && edge_bytes != NULL)
    atomicAdd(edge_bytes, static_cast<unsigned long long>(total) *
        (4ull + (StructOnly ? 0ull : 4ull) + (MaskMode ? 4ull : 0ull)));

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const Index e0 = static_cast<Index>(tile*GB_PUSH_TILE);
    const long long e1_ll = (tile + 1)*GB_PUSH_TILE;
    const Index e1 = (e1_ll > total) ? total : static_cast<Index>(e1_ll);

    __syncthreads();                       // smem reuse across tiles
    // Two warps find the frontier entries that own the first and the last edge
    // of the tile (32-ary cooperative search: ~5 dependent loads, not ~24).
    if (threadIdx.x < 32) {
      const int r = warpUpperBound(offs, nf + 1, e0) - 1;
      if (threadIdx.x == 0) s_range[0] = r;
    } else if (threadIdx.x < 64) {
      const int r = warpUpperBound(offs, nf + 1, e1 - 1) - 1;
      if (threadIdx.x == 32) s_range[1] = r;
    }
    __syncthreads();
    const Index i_lo = s_range[0];
    const Index i_hi = s_range[1];
    const int   cnt  = i_hi - i_lo + 1;
    const bool  staged = (cnt <= GB_PUSH_SEG);

    if (staged) {
      for (int j = threadIdx.x; j < cnt; j += GB_PUSH_NT) {
        const Index fi = i_lo + j;
        const Index r  = f_ind[fi];
        s_offs[j] = offs[fi];
        s_base[j] = rowptr[r];
        if (!StructOnly) s_uval[j] = f_val[fi];
      }
      if (threadIdx.x == 0) s_offs[cnt] = offs[i_hi + 1];
    }
    __syncthreads();

#pragma unroll
    for (int it = 0; it < GB_PUSH_EPT; ++it) {
      const Index e = e0 + it*GB_PUSH_NT + threadIdx.x;
      if (e < e1) {
        Index k;
        U     uv = U();
        if (staged) {
          const int j = upperBound(s_offs, cnt + 1, e) - 1;
          k = s_base[j] + (e - s_offs[j]);
          if (!StructOnly) uv = s_uval[j];
        } else {
          // Frontier segment too long for shared memory (only happens with
          // many zero-degree frontier entries): search the global array.
          const Index fi = upperBound(offs, nf + 1, e) - 1;
          k = rowptr[f_ind[fi]] + (e - offs[fi]);
          if (!StructOnly) uv = f_val[fi];
        }
        const Index col = ldStream(colind + k);
        bool keep = true;
        if (MaskMode != 0) {
          // The mask's bitmap shadow (bit == value != 0) when it is current:
          // a 32x denser gather target than the float array.
          const bool nonzero = (mask_bits != NULL)
              ? bitTest(mask_bits, col)
              : (__ldg(mask + col) != static_cast<M>(0));
          keep = (MaskMode == 1) ? nonzero : !nonzero;
        }
        if (keep) {
          if (!StructOnly) {
            const a av = ldStream(val + k);
            W prod;
            if (av == identity || uv == identity) prod = identity;
            else                                  prod = mul_op(av, uv);
            // min / max: a cell that already absorbs this product needs no
            // atomic at all — and no bit either, whoever moved it off the
            // identity has set that.  (The plain load may be stale; for a
            // monotone cell that only makes the filter weaker.)
            bool absorbed = false;
            bool touched_before = false;
            if (add_kind == 3 || add_kind == 5) {
              const W cur = acc[col];
              touched_before = (cur != identity);
              absorbed = touched_before && (add_op(cur, prod) == cur);
            }
            if (absorbed) {
              // nothing to do
            } else if (touched_before) {
              // the cell left the identity earlier (its bit is set or about to
              // be): result unused, so this compiles to a fire-and-forget RED
              (void)atomicCombineFetch(acc + col, prod, add_op, add_kind);
            } else {
              // the touched bit is set by the update that finds the identity
              const W old = atomicCombineFetch(acc + col, prod, add_op, add_kind);
              if (old == identity) bitSetAtomic(bits, col);
            }
          } else {
            bitSetAtomic(bits, col);
          }
        }
      }
    }
  }
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_KERNELS_SPMSPV_PUSH_CUH_
