"""Graph ingest on the device: R-MAT generation and
tuples -> CSR / CSC construction with the reference loader's semantics
(graphblas/util.hpp:264-329: symmetrise, drop self-loops and duplicates, sort
row-major).  Everything here runs in the native library's own kernels
(csrc/graphblas/backend/cuda/ingest.hpp: radix sort of packed keys, scans,
compaction); torch only owns the buffers.  Results are checked against oracle/ in
tests/.
"""
import ctypes as C

import torch

from . import _lib
from .api import _check, Matrix, FP32, INT32

INGEST_SYMMETRIZE = 1
INGEST_DROP_LOOPS = 2
INGEST_DEDUP = 4
INGEST_SYMMETRIC_STRUCTURE = 8


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def rmat_edges(scale, edgefactor=16, seed=1, device="cuda"):
    """Directed R-MAT edge list (a,b,c,d)=(0.57,0.19,0.19,0.05) as two int32
    device tensors; bit-identical to oracle orc_rmat_edges."""
    nedges = edgefactor << scale
    src = torch.empty(nedges, dtype=torch.int32, device=device)
    dst = torch.empty(nedges, dtype=torch.int32, device=device)
    _check(_lib.load().gb200_rmat_edges(int(scale), int(nedges), int(seed), 0,
                                        _p(src), _p(dst)),
           "gb200_rmat_edges")
    return src, dst


def build_csr(n, src, dst, undirected=True, val=None, return_values=False):
    """Loader semantics on the device.  Returns (rowptr int32[n+1],
    colind int32[nnz]) as device tensors — rows sorted, no loops, no duplicates —
    plus the float32 values of the kept tuples when return_values is set."""
    lib = _lib.load()
    flags = INGEST_DROP_LOOPS | INGEST_DEDUP | (INGEST_SYMMETRIZE if undirected else 0)
    handle = C.c_void_p()
    nnz = C.c_longlong(0)
    _check(lib.gb200_ingest_coo(int(n), int(n), _p(src), _p(dst), _p(val),
                                int(src.numel()), flags, C.byref(handle),
                                C.byref(nnz)), "gb200_ingest_coo")
    rowptr = torch.empty(n + 1, dtype=torch.int32, device=src.device)
    colind = torch.empty(max(nnz.value, 1), dtype=torch.int32, device=src.device)
    values = (torch.empty(max(nnz.value, 1), dtype=torch.float32, device=src.device)
              if return_values else None)
    try:
        _check(lib.gb200_ingest_export(handle, _p(rowptr), _p(colind), _p(values)),
               "gb200_ingest_export")
    finally:
        lib.gb200_ingest_free(handle)
    colind = colind[:nnz.value]
    if return_values:
        return rowptr, colind, values[:nnz.value]
    return rowptr, colind


def transpose_values(n, rowptr, colind, val):
    """Values of the CSC of a structurally symmetric CSR: cscVal such that the
    entry stored at position k of "column j" (= row j of the symmetric pattern)
    holds A(colind[k], j)."""
    out = torch.empty_like(val)
    _check(_lib.load().gb200_csr_transpose_values(
        int(n), int(n), int(colind.numel()), _p(rowptr), _p(colind), _p(val),
        None, None, _p(out)), "gb200_csr_transpose_values")
    return out


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
