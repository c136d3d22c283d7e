"""Graph ingest on the device: R-MAT generation and
COO -> CSR/CSC construction with the reference loader's semantics
(graphblas/util.hpp:264-329: symmetrise, drop self-loops and duplicates, sort
row-major), using torch for sort/unique plumbing and the native library for the
edge generator.  Results are checked against oracle/ in tests/.
"""
import ctypes as C

import torch

from . import _lib
from .api import _check, Matrix, FP32, INT32


def rmat_edges(scale, edgefactor=16, seed=1, device="cuda"):
    """Directed R-MAT edge list (a,b,c,d)=(0.57,0.19,0.19,0.05) as two int32
    device tensors; bit-identical to oracle orc_rmat_edges."""
    nedges = edgefactor << scale
    src = torch.empty(nedges, dtype=torch.int32, device=device)
    dst = torch.empty(nedges, dtype=torch.int32, device=device)
    _check(_lib.load().gb200_rmat_edges(int(scale), int(nedges), int(seed), 0,
                                        C.c_void_p(src.data_ptr()),
                                        C.c_void_p(dst.data_ptr())),
           "gb200_rmat_edges")
    return src, dst


def build_csr(n, src, dst, undirected=True):
    """Loader semantics on the device.  Returns (rowptr int32[n+1],
    colind int32[nnz]) as device tensors, rows sorted, no loops, no duplicates."""
    s = src.to(torch.int64)
    d = dst.to(torch.int64)
    if undirected:
        keep = s != d
        keys = torch.cat([s * n + d, (d * n + s)[keep]])
    else:
        keys = s * n + d
    del s, d
    keys = keys[(keys // n) != (keys % n)]          # drop self-loops
    keys = torch.unique(keys)                       # sorted + deduplicated
    rows = keys // n
    colind = (keys % n).to(torch.int32)
    del keys
    counts = torch.bincount(rows, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=src.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr.to(torch.int32), colind.contiguous()


def transpose_values(n, rowptr, colind, val):
    """Values of the CSC of a structurally symmetric CSR: cscVal such that the
    entry stored at position k of "column j" (= row j of the symmetric pattern)
    holds A(colind[k], j)."""
    nnz = colind.numel()
    rows = torch.repeat_interleave(
        torch.arange(n, device=colind.device, dtype=torch.int64),
        (rowptr[1:] - rowptr[:-1]).to(torch.int64))
    # position of (i, j) in row-major order is k; (j, i) sits at perm[k]
    key_t = colind.to(torch.int64) * n + rows
    order = torch.argsort(key_t)
    del key_t, rows
    # sorted transposed keys enumerate the same pattern in row-major order, so
    # cscVal[pos] = val[order[pos]]
    out = val[order]
    assert out.numel() == nnz
    return out.contiguous()


def matrix_from_csr(n, rowptr, colind, val=None, dtype=FP32, symmetric=True,
                    cscval=None):
    """graphblas::Matrix over device-resident CSR (+ aliased CSC when symmetric)."""
    if val is None:
        tdtype = torch.float32 if dtype == FP32 else torch.int32
        val = torch.ones(colind.numel(), dtype=tdtype, device=colind.device)
    A = Matrix(n, n, dtype=dtype)
    A.build_device_csr(rowptr, colind, val, colind.numel(), None, None, cscval,
                       symmetric=symmetric)
    return A
