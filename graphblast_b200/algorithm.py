"""graphblas::algorithm — the GraphBLAS algorithm drivers of the reference
(graphblas/algorithm/{bfs,sssp,pr,tc}.hpp), executed inside the native library as
loops of backend operations (include/graphblas/algorithm/*.hpp).

Each function returns the device time of the operation loop in milliseconds
("tight" in the reference drivers, example/gbfs.cu:110-115).
"""
import ctypes as C

from . import _lib
from .api import _check


def bfs(v, A, s, desc):
    """v[i] = BFS level of i from source s (source = 1, unreached = 0).
    reference algorithm/bfs.hpp:14-89"""
    ms = C.c_float(0)
    _check(_lib.load().gb200_bfs(v._h, A._h, int(s), desc._h, C.byref(ms)),
           "algorithm::bfs")
    return ms.value


def sssp(v, A, s, desc):
    """v[i] = shortest distance from s (unreached = FLT_MAX).
    reference algorithm/sssp.hpp:15-103"""
    ms = C.c_float(0)
    _check(_lib.load().gb200_sssp(v._h, A._h, int(s), desc._h, C.byref(ms)),
           "algorithm::sssp")
    return ms.value


def pr(p, A, alpha, eps, desc):
    """PageRank power iteration on a pre-normalised A (Matrix.pr_normalize).
    reference algorithm/pr.hpp:15-94"""
    ms = C.c_float(0)
    _check(_lib.load().gb200_pr(p._h, A._h, float(alpha), float(eps), desc._h,
                                C.byref(ms)), "algorithm::pr")
    return ms.value


def tc(A, B, desc):
    """Triangle count of a lower-triangular INT32 matrix A; B receives
    (A*A^T).*A.  Returns (ntris, tight_ms).  reference algorithm/tc.hpp:15-54"""
    ms = C.c_float(0)
    n = C.c_longlong(0)
    _check(_lib.load().gb200_tc(C.byref(n), A._h, B._h, desc._h, C.byref(ms)),
           "algorithm::tc")
    return n.value, ms.value
