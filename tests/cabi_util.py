"""Call the C ABI (include/gnn_b200.h) through ctypes on torch-owned device memory."""
import ctypes

import numpy as np
import torch

from gnn_b200 import _native


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def build_adj(fullrowptr, rowptr, colidx, normfact, M, K, want_indices=True, want_rows=False):
    lib = _native.cabi()
    nnz = int(colidx.numel())
    idx = torch.empty((2, nnz), dtype=torch.int64, device="cuda") if want_indices else None
    vals = torch.empty(nnz, dtype=torch.float32, device="cuda")
    col32 = torch.empty(nnz, dtype=torch.int32, device="cuda")
    row32 = torch.full((nnz,), -7, dtype=torch.int32, device="cuda") if want_rows else None
    rc = lib.gnn_build_adj(_ptr(fullrowptr), _ptr(rowptr), _ptr(colidx), colidx.element_size(), _ptr(normfact), M, K, nnz,
                           _ptr(idx), _ptr(vals), _ptr(col32), _ptr(row32), _stream())
    _native.check(rc, "gnn_build_adj")
    if want_rows:
        return idx, vals, col32, row32
    return idx, vals, col32


def csr_spmm(rowptr, colidx, vals, M, K, X, ldx=None, ldy=None):
    lib = _native.cabi()
    D = X.shape[1]
    ldx = ldx if ldx is not None else X.stride(0)
    nnz = int(vals.numel())
    ldy = ldy or D
    Ybuf = torch.full((M, ldy), float("nan"), dtype=torch.float32, device="cuda")
    wsb = lib.gnn_csr_spmm_workspace_bytes(M, nnz, D)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    rc = lib.gnn_csr_spmm_f32(_ptr(rowptr), _ptr(colidx), _ptr(vals), M, K, nnz, D, _ptr(X), ldx, _ptr(Ybuf), ldy,
                              _ptr(ws), wsb, _stream())
    _native.check(rc, "gnn_csr_spmm_f32")
    return Ybuf[:, :D]


def gather_spmm(rowptr, colidx, vals, M, K, D, xrows):
    lib = _native.cabi()
    nnz = int(vals.numel())
    Y = torch.full((M, D), float("nan"), dtype=torch.float32, device="cuda")
    wsb = lib.gnn_csr_spmm_workspace_bytes(M, nnz, D)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    rc = lib.gnn_gather_spmm_f32(_ptr(rowptr), _ptr(colidx), _ptr(vals), M, K, nnz, D, _ptr(xrows), _ptr(Y), D, _ptr(ws), wsb,
                                 _stream())
    _native.check(rc, "gnn_gather_spmm_f32")
    return Y


def csr_transpose(rowptr, colidx, vals, M, K, want_rows=False):
    lib = _native.cabi()
    nnz = int(vals.numel())
    t_rowptr = torch.empty(K + 1, dtype=torch.int32, device="cuda")
    t_colidx = torch.empty(nnz, dtype=torch.int32, device="cuda")
    t_vals = torch.empty(nnz, dtype=torch.float32, device="cuda")
    wsb = lib.gnn_csr_transpose_workspace_bytes(M, K, nnz)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    t_rowidx = torch.full((nnz,), -7, dtype=torch.int32, device="cuda") if want_rows else None
    rc = lib.gnn_csr_transpose(_ptr(rowptr), _ptr(colidx), _ptr(vals), M, K, nnz, _ptr(t_rowptr), _ptr(t_colidx), _ptr(t_vals),
                               _ptr(t_rowidx), _ptr(ws), wsb, _stream())
    _native.check(rc, "gnn_csr_transpose")
    if want_rows:
        return t_rowptr, t_colidx, t_vals, t_rowidx
    return t_rowptr, t_colidx, t_vals


def coo_to_csr(indices, M):
    lib = _native.cabi()
    nnz = indices.shape[1]
    rowptr = torch.empty(M + 1, dtype=torch.int32, device="cuda")
    col32 = torch.empty(nnz, dtype=torch.int32, device="cuda")
    _native.check(lib.gnn_coo_to_csr(_ptr(indices), M, nnz, _ptr(rowptr), _ptr(col32), _stream()), "gnn_coo_to_csr")
    return rowptr, col32


def placement_remap(input_nodes, dev_of, idx_of, devices, bases, ld_src, ld_host=None):
    lib = _native.cabi()
    n0 = int(input_nodes.numel())
    world = int(devices.numel())
    src = torch.empty(n0, dtype=torch.int32, device="cuda")
    slot = torch.empty(n0, dtype=torch.int64, device="cuda")
    xrows = torch.empty(n0, dtype=torch.int64, device="cuda")
    counts = torch.empty(world + 2, dtype=torch.int64, device="cuda")
    rc = lib.gnn_placement_remap(_ptr(input_nodes), n0, _ptr(dev_of), _ptr(idx_of), _ptr(devices), world, _ptr(bases), ld_src,
                                 ld_host if ld_host is not None else ld_src, _ptr(src), _ptr(slot), _ptr(xrows), _ptr(counts), _stream())
    _native.check(rc, "gnn_placement_remap")
    return src, slot, xrows, counts


def gather_rows(xrows, F, ld_out=None):
    lib = _native.cabi()
    n0 = int(xrows.numel())
    ld_out = ld_out or F
    out = torch.full((n0, ld_out), float("nan"), dtype=torch.float32, device="cuda")
    _native.check(lib.gnn_gather_rows_f32(_ptr(xrows), n0, F, _ptr(out), ld_out, _stream()), "gnn_gather_rows_f32")
    return out[:, :F]


def index_rows(X, idx):
    lib = _native.cabi()
    n, F = int(idx.numel()), X.shape[1]
    out = torch.empty((n, F), dtype=torch.float32, device="cuda")
    _native.check(lib.gnn_index_rows_f32(_ptr(X), X.stride(0), _ptr(idx), n, F, _ptr(out), F, _stream()), "gnn_index_rows_f32")
    return out


def csr_spmm_ex(rowptr, colidx, vals, M, K, X, counters, partials, zeroed, ldx=None, rowidx=None):
    """gnn_csr_spmm_f32_ex with caller-owned counter / partial regions (uint8 tensors)."""
    lib = _native.cabi()
    D = X.shape[1]
    ldx = ldx if ldx is not None else X.stride(0)
    nnz = int(vals.numel())
    Y = torch.full((M, D), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.gnn_csr_spmm_f32_ex(_ptr(rowptr), _ptr(rowidx), _ptr(colidx), _ptr(vals), M, K, nnz, D, _ptr(X), ldx, _ptr(Y), D,
                                 _ptr(counters), _ptr(partials), int(partials.numel()), 1 if zeroed else 0, _stream())
    _native.check(rc, "gnn_csr_spmm_f32_ex")
    return Y


def csr_spmm_t(rowptr, colidx, vals, M, K, G, lddx=None, rowidx=None):
    """gnn_csr_spmm_t_f32: dX = A^T.G from A's CSR (transpose-free)."""
    lib = _native.cabi()
    D = G.shape[1]
    nnz = int(vals.numel())
    lddx = lddx or D
    dX = torch.full((K, lddx), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.gnn_csr_spmm_t_f32(_ptr(rowptr), _ptr(rowidx), _ptr(colidx), _ptr(vals), M, K, nnz, D, _ptr(G), G.stride(0) if M > 1 else D,
                                _ptr(dX), lddx, _stream())
    _native.check(rc, "gnn_csr_spmm_t_f32")
    return dX


def set_transpose_budget(nbytes):
    return _native.cabi().gnn_set_transpose_budget(int(nbytes))
