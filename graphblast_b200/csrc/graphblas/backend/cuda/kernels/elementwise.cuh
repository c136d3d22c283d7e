// graphblast_b200 backend — element-wise / assign kernels that sit between two
// mxv calls in every algorithm loop.  Semantics follow the reference kernels
// one-for-one (file:line cited per kernel); launch shape is grid-stride sized
// to the SM count instead of ceil(n/128) CTAs.
#ifndef GRAPHBLAS_BACKEND_CUDA_KERNELS_ELEMENTWISE_CUH_
#define GRAPHBLAS_BACKEND_CUDA_KERNELS_ELEMENTWISE_CUH_

#include "graphblas/backend/cuda/kernels/common.cuh"

namespace graphblas {
namespace backend {

// w[i] = op(u[i], v[i])
// (reference eWiseAddDenseDenseKernel, kernels/ewiseadd.hpp:9-24)
template <typename W, typename U, typename V, typename Op>
__global__ void ewiseBinaryDenseKernel(W* w, Op op, const U* u, const V* v,
                                       Index n) {
  Index i = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; i < n; i += stride) {
    U a = u[i];
    V b = v[i];
    w[i] = op(a, b);
  }
}

// w[i] = reverse ? op(identity, w[i]) : op(w[i], identity)
// (reference eWiseAddDenseConstantKernel, kernels/ewiseadd.hpp:49-63)
template <typename W, typename T, typename Op>
__global__ void ewiseConstantKernel(W* w, Op op, T identity, bool reverse,
                                    Index n) {
  Index i = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; i < n; i += stride) {
    W x = w[i];
    w[i] = reverse ? op(identity, x) : op(x, identity);
  }
}

// Fused form of "w = dup(v); constant pass" for the sparse-dense eWiseAdd:
// w[i] = reverse ? op(identity, v[i]) : op(v[i], identity)  (w may alias v).
template <typename W, typename V, typename T, typename Op>
__global__ void ewiseConstantFromKernel(W* w, const V* v, Op op, T identity,
                                        bool reverse, Index n) {
  Index i = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; i < n; i += stride) {
    V x = v[i];
    w[i] = reverse ? op(identity, x) : op(x, identity);
  }
}

// w[ind] = op(u_val[k], v[ind]) for the k-th sparse entry
// (reference eWiseAddSparseDenseKernel, kernels/ewiseadd.hpp:29-45)
template <typename W, typename U, typename V, typename Op>
__global__ void ewiseSparseDenseKernel(W* w, Op op, const Index* u_ind,
                                       const U* u_val, const V* v,
                                       Index u_nvals) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < u_nvals; k += stride) {
    Index ind = u_ind[k];
    U a = u_val[k];
    V b = v[ind];
    w[ind] = op(a, b);
  }
}

// w[i] = op(u[i], val)
// (reference scalar eWiseMultKernel, kernels/ewisemult.hpp:161-174)
template <typename W, typename U, typename V, typename Op>
__global__ void ewiseScalarKernel(W* w, Op op, const U* u, Index n, V val) {
  Index i = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; i < n; i += stride) w[i] = op(u[i], val);
}

// Dense-dense eWiseMult with the identity short-circuit
// (reference eWiseMultKernel, kernels/ewisemult.hpp:9-28):
// w[i] = (u[i]==identity || v[i]==identity) ? identity : mul(u[i], v[i]).
template <typename W, typename T, typename U, typename V, typename MulOp>
__global__ void ewiseMultDenseKernel(W* w, T identity, MulOp mul_op,
                                     const U* u, const V* v, Index n) {
  Index i = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; i < n; i += stride) {
    U a = u[i];
    V b = v[i];
    if (a == identity || b == identity) w[i] = identity;
    else                                w[i] = mul_op(a, b);
  }
}

// Dense-dense eWiseMult under a dense mask
// (reference kernels/ewisemult.hpp:63-86).
template <typename W, typename M, typename T, typename U, typename V,
          typename MulOp>
__global__ void ewiseMultDenseMaskedKernel(W* w, const M* mask, T identity,
                                           MulOp mul_op, const U* u,
                                           const V* v, Index n) {
  Index i = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; i < n; i += stride) {
    U a = u[i];
    V b = v[i];
    M m = mask[i];
    if (m == static_cast<M>(0) || a == identity || b == identity)
      w[i] = identity;
    else
      w[i] = mul_op(a, b);
  }
}

// Dense-dense eWiseMult restricted to a sparse mask -> sparse output
// (reference kernels/ewisemult.hpp:33-58).
template <typename W, typename M, typename U, typename V, typename MulOp>
__global__ void ewiseMultSparseMaskKernel(Index* w_ind, W* w_val,
                                          const Index* mask_ind,
                                          const M* mask_val, Index mask_nvals,
                                          MulOp mul_op, const U* u,
                                          const V* v) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < mask_nvals; k += stride) {
    Index ind = mask_ind[k];
    M m = mask_val[k];
    W out = static_cast<W>(0);
    if (m != static_cast<M>(0)) out = mul_op(u[ind], v[ind]);
    w_ind[k] = ind;
    w_val[k] = out;
  }
}

// Sparse-dense eWiseMult -> sparse output with u's pattern
// (reference kernels/ewisemult.hpp:88-117): identity inputs produce 0.
template <typename W, typename T, typename U, typename V, typename MulOp>
__global__ void ewiseMultSparseDenseKernel(Index* w_ind, W* w_val, T identity,
                                           MulOp mul_op, const Index* u_ind,
                                           const U* u_val, Index u_nvals,
                                           const V* v, bool reverse) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < u_nvals; k += stride) {
    Index ind = u_ind[k];
    U a = u_val[k];
    if (a != identity) {
      V b = v[ind];
      w_val[k] = reverse ? mul_op(b, a) : mul_op(a, b);
    } else {
      w_val[k] = static_cast<W>(0);
    }
    w_ind[k] = ind;
  }
}

// Sparse-dense eWiseMult under a SPARSE mask -> sparse output with the mask's
// pattern (reference kernels/ewisemult.hpp:119-160): an entry is mul(u, v) where
// the mask value is non-zero, v is not the identity and u stores that index
// (u's indices are sorted: lower-bound search), 0 everywhere else.
template <typename W, typename M, typename T, typename U, typename V, typename MulOp>
__global__ void ewiseMultSparseDenseSparseMaskKernel(
    Index* w_ind, W* w_val, const Index* mask_ind, const M* mask_val, Index mask_nvals,
    T identity, MulOp mul_op, const Index* u_ind, const U* u_val, Index u_nvals,
    const V* v, bool reverse) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < mask_nvals; k += stride) {
    const Index ind = mask_ind[k];
    W out = static_cast<W>(0);
    if (mask_val[k] != static_cast<M>(0)) {
      const V b = v[ind];
      if (b != identity) {
        const Index at = findSorted(u_ind, 0, u_nvals, ind);
        if (at < u_nvals && u_ind[at] == ind) {
          const U a = u_val[at];
          out = reverse ? mul_op(b, a) : mul_op(a, b);
        }
      }
    }
    w_ind[k] = ind;
    w_val[k] = out;
  }
}

// CSR values scaled by a per-row vector entry: C(i,j) = mul(A(i,j), b[i])
// (reference eWiseMultCSRKernel, kernels/ewisemult.hpp:177-205); warp per row.
template <typename c, typename a, typename b, typename MulOp>
__global__ void ewiseMultRowBroadcastKernel(c* C_val, MulOp mul_op,
                                            const Index* rowptr,
                                            const a* A_val, Index nrows,
                                            const b* B_val) {
  const int lane = threadIdx.x & 31;
  Index row = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
  const Index nwarps = (gridDim.x*blockDim.x) >> 5;
  for (; row < nrows; row += nwarps) {
    Index beg = rowptr[row], end = rowptr[row+1];
    b s = B_val[row];
    for (Index k = beg + lane; k < end; k += 32)
      C_val[k] = mul_op(A_val[k], s);
  }
}

// Compressed values scaled by the vector entry of their minor index:
// C(k) = mul(A(k), b[minor_ind[k]])
// (reference eWiseMultCSCKernel, kernels/ewisemult.hpp:208-237).
template <typename c, typename a, typename b, typename MulOp>
__global__ void ewiseMultIndexBroadcastKernel(c* C_val, MulOp mul_op,
                                              const Index* minor_ind,
                                              const a* A_val, Index nvals,
                                              const b* B_val) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < nvals; k += stride)
    C_val[k] = mul_op(A_val[k], B_val[minor_ind[k]]);
}

// Masked constant assign, dense target, dense mask
// (reference assignDenseDenseMaskedKernel, kernels/assign_dense.hpp:8-39).
template <bool UseScmp, typename U, typename M>
__global__ void assignDenseDenseMaskKernel(U* u, Index n, const M* mask,
                                           U val) {
  Index i = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; i < n; i += stride) {
    M m = mask[i];
    if ((UseScmp && m == static_cast<M>(0)) ||
        (!UseScmp && m != static_cast<M>(0)))
      u[i] = val;
  }
}

// Masked constant assign, dense target, dense mask given as its bitmap shadow.
// One lane per row, one mask word per warp (broadcast load); optionally keeps
// the target's own bitmap shadow exact (val != 0 sets the selected bits, val == 0
// clears them).
template <bool UseScmp, typename U>
__global__ void assignDenseBitsMaskKernel(U* u, unsigned int* u_bits, Index n,
                                          const unsigned int* mask_bits,
                                          U val) {
  // Each lane fetches one mask word (one coalesced 128-byte request covers 1024
  // rows); the warp then walks the non-empty words, so the stores of different
  // words are not separated by a dependent load.
  const int lane = threadIdx.x & 31;
  const Index warp   = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
  const Index nwarps = (gridDim.x*blockDim.x) >> 5;
  const Index nwords = (n + 31) >> 5;
  for (Index wbase = warp*32; wbase < nwords; wbase += nwarps*32) {
    const Index word = wbase + lane;
    unsigned int sel = 0u;
    if (word < nwords) {
      sel = __ldg(mask_bits + word);
      if (UseScmp) sel = ~sel;
      const Index base = word*32;
      if (base + 32 > n) sel &= (1u << (n - base)) - 1u;
    }
    unsigned int pending = __ballot_sync(GB_FULL_MASK, sel != 0u);
    while (pending) {
      const int j = __ffs(pending) - 1;
      pending &= pending - 1;
      const unsigned int s = __shfl_sync(GB_FULL_MASK, sel, j);
      if ((s >> lane) & 1u) u[(wbase + j)*32 + lane] = val;
    }
    if (u_bits != NULL && sel != 0u) {
      if (val != static_cast<U>(0)) u_bits[word] |= sel;
      else                          u_bits[word] &= ~sel;
    }
  }
}

// Masked constant assign, dense target, sparse mask, keeping the bitmap shadow.
template <typename U>
__global__ void assignDenseSparseMaskBitsKernel(U* u, unsigned int* u_bits,
                                                const Index* mask_ind,
                                                Index mask_nvals, U val) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < mask_nvals; k += stride) {
    const Index ind = mask_ind[k];
    u[ind] = val;
    const unsigned int m = 1u << (ind & 31);
    if (val != static_cast<U>(0)) atomicOr(u_bits + (ind >> 5), m);
    else                          atomicAnd(u_bits + (ind >> 5), ~m);
  }
}

// bits[ind[k]] = 1 (bitmap shadow of a scatter of non-zero constants)
__global__ void scatterBitsKernel(unsigned int* bits, const Index* ind,
                                  Index nvals) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < nvals; k += stride)
    atomicOr(bits + (ind[k] >> 5), 1u << (ind[k] & 31));
}

// Masked constant assign, dense target, sparse mask (non-complemented only:
// the reference's SCMP variant is "not implemented", assign_dense.hpp:58-63).
template <typename U>
__global__ void assignDenseSparseMaskKernel(U* u, const Index* mask_ind,
                                            Index mask_nvals, U val) {
  Index k = blockIdx.x*blockDim.x + threadIdx.x;
  const Index stride = gridDim.x*blockDim.x;
  for (; k < mask_nvals; k += stride) u[mask_ind[k]] = val;
}

}  // namespace backend
}  // namespace graphblas

#endif  // GRAPHBLAS_BACKEND_CUDA_KERNELS_ELEMENTWISE_CUH_
