"""CPU oracle for the hot path (TEST INFRASTRUCTURE ONLY - see oracle/oracle.c).

numpy-facing wrappers over liboracle.so (plain C, built by oracle/Makefile).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Each wrapper cites the reference lines its C
function restates.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return so


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.oracle_sampled_nodes.restype = ctypes.c_int64
    return _LIB


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


_i64 = ctypes.c_int64


def build_adj(fullrowptr, rowptr, colidx, normfact, nrows):
    """create_coo_tensor (reference cuda_spmm.cu:787-827): -> (rows i64, cols i64, vals f32)."""
    fullrowptr, rowptr, normfact = _c(fullrowptr, np.int32), _c(rowptr, np.int32), _c(normfact, np.float32)
    nnz = int(rowptr[nrows])
    rows, cols, vals = np.empty(nnz, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float32)
    colidx = np.ascontiguousarray(colidx)
    if colidx.dtype == np.int16:
        fn = lib().oracle_build_adj
    elif colidx.dtype == np.int32:
        fn = lib().oracle_build_adj_i32
    else:
        raise TypeError("colidx must be int16 (reference sampler.py:136) or int32")
    fn(_p(fullrowptr), _p(rowptr), _p(colidx), _p(normfact), _i64(nrows), _p(rows), _p(cols), _p(vals))
    return rows, cols, vals


def _spmm(fn, rowptr, colidx, vals, M, X):
    rowptr, colidx, vals = _c(rowptr, np.int32), _c(colidx, np.int32), _c(vals, np.float32)
    X = _c(X, np.float32)
    D = X.shape[1]
    Y = np.empty((M, D), np.float32)
    fn(_p(rowptr), _p(colidx), _p(vals), _i64(M), _i64(D), _p(X), _i64(D), _p(Y), _i64(D))
    return Y


def spmm_f64acc(rowptr, colidx, vals, M, X):
    """Y = A.X, double accumulation: the arbiter for the 1e-5 tolerance."""
    return _spmm(lib().oracle_csr_spmm_f64acc, rowptr, colidx, vals, M, X)


def spmm_seqfma(rowptr, colidx, vals, M, X):
    """Y = A.X in the summation order of reference spmm_naive (cuda_spmm.cu:88-100)."""
    return _spmm(lib().oracle_csr_spmm_seqfma, rowptr, colidx, vals, M, X)


def spmm_chunk64(rowptr, colidx, vals, M, X):
    """Y = A.X in 64-nnz chunks like reference spmm_load_balance (cuda_spmm.cu:163-253)."""
    return _spmm(lib().oracle_csr_spmm_chunk64, rowptr, colidx, vals, M, X)


def _spmm_t(fn, rowptr, colidx, vals, M, K, G):
    rowptr, colidx, vals = _c(rowptr, np.int32), _c(colidx, np.int32), _c(vals, np.float32)
    G = _c(G, np.float32)
    D = G.shape[1]
    dX = np.empty((K, D), np.float32)
    fn(_p(rowptr), _p(colidx), _p(vals), _i64(M), _i64(K), _i64(D), _p(G), _i64(D), _p(dX), _i64(D))
    return dX


def spmm_t_f64acc(rowptr, colidx, vals, M, K, G):
    """dX = A^T.G (reference custom_sparse_ops.py:30-37), double accumulation."""
    return _spmm_t(lib().oracle_csr_spmm_t_f64acc, rowptr, colidx, vals, M, K, G)


def spmm_t_seqfma(rowptr, colidx, vals, M, K, G):
    """dX = A^T.G in the order of transpose().coalesce() + spmm_naive."""
    return _spmm_t(lib().oracle_csr_spmm_t_seqfma, rowptr, colidx, vals, M, K, G)


def csr_transpose(rowptr, colidx, M, K):
    """CSR of A^T with entries in ascending source row; -> (t_rowptr, t_colidx, perm)."""
    rowptr, colidx = _c(rowptr, np.int32), _c(colidx, np.int32)
    nnz = int(rowptr[M])
    t_rowptr, t_colidx, perm = np.empty(K + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.int32)
    lib().oracle_csr_transpose(_p(rowptr), _p(colidx), _i64(M), _i64(K), _p(t_rowptr), _p(t_colidx), _p(perm))
    return t_rowptr, t_colidx, perm


def coo_rows_to_rowptr(rows, M):
    rows = _c(rows, np.int64)
    rowptr = np.empty(M + 1, np.int32)
    lib().oracle_coo_rows_to_rowptr(_p(rows), _i64(rows.size), _i64(M), _p(rowptr))
    return rowptr


def sampled_nodes(after_nodes, previous_nodes):
    """reference sampler.py:143."""
    a, p = _c(after_nodes, np.int64), _c(previous_nodes, np.int64)
    out = np.empty(a.size, np.int64)
    n = lib().oracle_sampled_nodes(_p(a), _i64(a.size), _p(p), _i64(p.size), _p(out))
    return out[:n].copy()


def placement_remap(input_nodes, device_id_of_nodes, idx_of_nodes_on_device, devices):
    """reference sampler.py:150-158 -> (src_dev i32 [n0], slot i64 [n0])."""
    inp = _c(input_nodes, np.int64)
    did, idx = _c(device_id_of_nodes, np.int64), _c(idx_of_nodes_on_device, np.int64)
    devs = _c(devices, np.int64)
    src, slot = np.empty(inp.size, np.int32), np.empty(inp.size, np.int64)
    lib().oracle_placement_remap(_p(inp), _i64(inp.size), _p(did), _p(idx), _p(devs), _i64(devs.size), _p(src), _p(slot))
    return src, slot


def _bases(gpu_buffers, host_table):
    bufs = [_c(b, np.float32) for b in gpu_buffers] + [_c(host_table, np.float32)]
    ld = bufs[-1].shape[1]
    arr = (ctypes.c_void_p * len(bufs))(*[b.ctypes.data for b in bufs])
    return bufs, arr, ld


def gather_rows(gpu_buffers, host_table, src_dev, slot):
    """reference main.py:129-134 -> fp32 [n0, F]; bit-exact copy."""
    bufs, arr, ld = _bases(gpu_buffers, host_table)
    src, slot = _c(src_dev, np.int32), _c(slot, np.int64)
    out = np.zeros((src.size, ld), np.float32)
    lib().oracle_gather_rows(arr, _i64(len(bufs) - 1), _p(src), _p(slot), _i64(src.size), _i64(ld), _i64(ld), _p(out), _i64(ld))
    return out


def gather_spmm_f64acc(rowptr, colidx, vals, M, gpu_buffers, host_table, src_dev, slot):
    bufs, arr, ld = _bases(gpu_buffers, host_table)
    rowptr, colidx, vals = _c(rowptr, np.int32), _c(colidx, np.int32), _c(vals, np.float32)
    src, slot = _c(src_dev, np.int32), _c(slot, np.int64)
    Y = np.empty((M, ld), np.float32)
    lib().oracle_gather_spmm_f64acc(_p(rowptr), _p(colidx), _p(vals), _i64(M), _i64(ld), arr, _i64(len(bufs) - 1),
                                    _p(src), _p(slot), _i64(ld), _p(Y), _i64(ld))
    return Y


def rel_err(y, ref64):
    """Worst-row relative L2 error and max-abs/max error against an fp64-accumulated result."""
    y = np.asarray(y, np.float64)
    r = np.asarray(ref64, np.float64)
    num = np.sqrt(((y - r) ** 2).sum(axis=1))
    den = np.sqrt((r ** 2).sum(axis=1))
    row = np.where(den > 0, num / np.maximum(den, 1e-300), num)
    scale = np.abs(r).max() if r.size else 1.0
    return float(row.max() if row.size else 0.0), float(np.abs(y - r).max() / max(scale, 1e-300) if r.size else 0.0)
