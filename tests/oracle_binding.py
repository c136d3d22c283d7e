"""ctypes binding of the CPU checkers (test infrastructure only).

  oracle/libgboracle.so      C restatement (oracle/gb_oracle.c)
  oracle/_ref/libgbref.so    the reference's own CPU code (oracle/ref_wrapper.cpp),
                             present when it was built in a container that mounts
                             /root/reference

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libgboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libgbref.so")

_orc = None
_ref = None
FLT_MAX = np.finfo(np.float32).max


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libgboracle.so"])
        _orc = C.CDLL(ORACLE_SO)
        _orc.orc_identity.restype = C.c_float
        _orc.orc_add.restype = C.c_float
        _orc.orc_mul.restype = C.c_float
        _orc.orc_add.argtypes = [C.c_int, C.c_float, C.c_float]
        _orc.orc_mul.argtypes = [C.c_int, C.c_float, C.c_float]
        _orc.orc_tc.restype = C.c_longlong
        _orc.orc_build_csr.restype = C.c_longlong
        _orc.orc_build_csr.argtypes = [C.c_int, C.c_longlong, C.c_void_p,
                                       C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p]
        _orc.orc_tril.restype = C.c_longlong
        _orc.orc_rmat_edges.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong,
                                        C.c_longlong, C.c_void_p, C.c_void_p]
        _orc.orc_pr.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_float, C.c_float, C.c_int]
    return _orc


def ref():
    """The reference's own CPU implementation, or None when not built."""
    global _ref
    if _ref is None and os.path.exists(REF_SO):
        _ref = C.CDLL(REF_SO)
        _ref.ref_pr.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_float, C.c_float, C.c_int]
        _ref.ref_load_mtx.argtypes = [C.c_char_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p]
    return _ref


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- oracle (C restatement) ---------------------------------------------------

def identity(semiring):
    return float(lib().orc_identity(int(semiring)))


def bfs(rowptr, colind, src, stop=1 << 30):
    rowptr, colind = _i32(rowptr), _i32(colind)
    n = len(rowptr) - 1
    out = np.zeros(n, dtype=np.int32)
    lib().orc_bfs(n, _p(rowptr), _p(colind), _p(out), int(src), int(stop))
    return out


def sssp(rowptr, colind, val, src):
    rowptr, colind, val = _i32(rowptr), _i32(colind), _f32(val)
    n = len(rowptr) - 1
    out = np.zeros(n, dtype=np.float32)
    lib().orc_sssp(n, _p(rowptr), _p(colind), _p(val), _p(out), int(src))
    return out


def pr(rowptr, colind, alpha=0.85, eps=1e-8, max_niter=10):
    rowptr, colind = _i32(rowptr), _i32(colind)
    n = len(rowptr) - 1
    out = np.zeros(n, dtype=np.float32)
    lib().orc_pr(n, _p(rowptr), _p(colind), _p(out), alpha, eps, int(max_niter))
    return out


def tc(rowptr, colind):
    rowptr, colind = _i32(rowptr), _i32(colind)
    return int(lib().orc_tc(len(rowptr) - 1, _p(rowptr), _p(colind)))


def vxm(semiring, rowptr, colind, val, u, ncols=None, u_present=None,
        mask=None, scmp=False):
    """Returns (w, w_present) for w = u^T A over `semiring` (push formulation)."""
    rowptr, colind, val, u = _i32(rowptr), _i32(colind), _f32(val), _f32(u)
    nrows = len(rowptr) - 1
    ncols = nrows if ncols is None else ncols
    w = np.zeros(ncols, dtype=np.float32)
    wp = np.zeros(ncols, dtype=np.uint8)
    up = None if u_present is None else np.ascontiguousarray(u_present,
                                                             dtype=np.uint8)
    m = None if mask is None else _f32(mask)
    lib().orc_vxm(int(semiring), nrows, ncols, _p(rowptr), _p(colind), _p(val),
                  _p(u), _p(up) if up is not None else None,
                  _p(m) if m is not None else None, 1 if scmp else 0, _p(w),
                  _p(wp))
    return w, wp


def reduce_rows(rowptr, val):
    rowptr, val = _i32(rowptr), _f32(val)
    n = len(rowptr) - 1
    w = np.zeros(n, dtype=np.float32)
    lib().orc_reduce_rows(n, _p(rowptr), _p(val), _p(w))
    return w


def build_csr(n, src, dst, undirected=True):
    src, dst = _i32(src), _i32(dst)
    cap = len(src) * (2 if undirected else 1)
    rowptr = np.zeros(n + 1, dtype=np.int32)
    colind = np.zeros(max(cap, 1), dtype=np.int32)
    nnz = lib().orc_build_csr(n, len(src), _p(src), _p(dst),
                              1 if undirected else 0, _p(rowptr), _p(colind))
    return rowptr, colind[:nnz].copy()


def tril(rowptr, colind):
    rowptr, colind = _i32(rowptr).copy(), _i32(colind).copy()
    nnz = lib().orc_tril(len(rowptr) - 1, _p(rowptr), _p(colind))
    return rowptr, colind[:nnz].copy()


def rmat_edges(scale, edgefactor=16, seed=1):
    m = edgefactor << scale
    src = np.zeros(m, dtype=np.int32)
    dst = np.zeros(m, dtype=np.int32)
    lib().orc_rmat_edges(scale, m, seed, 0, _p(src), _p(dst))
    return src, dst


def rmat_csr(scale, edgefactor=16, seed=1):
    src, dst = rmat_edges(scale, edgefactor, seed)
    return build_csr(1 << scale, src, dst, True)


# ---- reference's own CPU code (when built) --------------------------------------

def ref_bfs(rowptr, colind, src, stop=1 << 30):
    rowptr, colind = _i32(rowptr), _i32(colind)
    n = len(rowptr) - 1
    out = np.zeros(n, dtype=np.int32)
    ref().ref_bfs(n, _p(rowptr), _p(colind), _p(out), int(src), int(stop))
    return out


def ref_sssp(rowptr, colind, val, src):
    rowptr, colind, val = _i32(rowptr), _i32(colind), _f32(val).copy()
    n = len(rowptr) - 1
    out = np.zeros(n, dtype=np.float32)
    ref().ref_sssp(n, _p(rowptr), _p(colind), _p(val), _p(out), int(src),
                   1 << 30)
    return out


def ref_pr(rowptr, colind, alpha=0.85, eps=1e-8, max_niter=10):
    rowptr, colind = _i32(rowptr), _i32(colind)
    n = len(rowptr) - 1
    val = np.ones(max(len(colind), 1), dtype=np.float32)
    out = np.zeros(n, dtype=np.float32)
    ref().ref_pr(n, _p(rowptr), _p(colind), _p(val), _p(out), alpha, eps,
                 int(max_niter))
    return out


def ref_tc(rowptr, colind):
    rowptr, colind = _i32(rowptr), _i32(colind)
    out = C.c_int(0)
    ref().ref_tc(len(rowptr) - 1, _p(rowptr), _p(colind), C.byref(out))
    return out.value


def ref_load_mtx(path, directed):
    n = C.c_int(0)
    nnz = ref().ref_load_mtx(path.encode(), int(directed), C.byref(n), None,
                             None, None)
    rowptr = np.zeros(n.value + 1, dtype=np.int32)
    colind = np.zeros(max(nnz, 1), dtype=np.int32)
    val = np.zeros(max(nnz, 1), dtype=np.float32)
    ref().ref_load_mtx(path.encode(), int(directed), C.byref(n), _p(rowptr),
                       _p(colind), _p(val))
    return rowptr, colind[:nnz].copy(), val[:nnz].copy()


def ref_uniform_weights(seed, lo, hi, n):
    out = np.zeros(n, dtype=np.float32)
    ref().ref_uniform_weights.argtypes = [C.c_int, C.c_int, C.c_int,
                                          C.c_longlong, C.c_void_p]
    ref().ref_uniform_weights(int(seed), int(lo), int(hi), int(n), _p(out))
    return out


def read_mtx_edges(path):
    """Minimal .mtx reader for tests (0-based src, dst, symmetric flag)."""
    with open(path) as f:
        header = f.readline().lower()
        symmetric = "symmetric" in header
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        n, _, m = [int(x) for x in line.split()[:3]]
        data = np.loadtxt(f, ndmin=2)
    src = data[:, 0].astype(np.int32) - 1
    dst = data[:, 1].astype(np.int32) - 1
    return n, src, dst, symmetric
