"""B200-native replacement of the reference's ``custom_sparse_ops`` module.

Same public names and meaning (reference custom_sparse_ops.py:8-41):

    spmm_cpp                      the extension module (spmm_load_balance, spmm_naive, create_coo_tensor)
    spmm_forward_time / spmm_backward_time   floats read by main.py:196
    SparseDenseMM                 autograd.Function: forward(ctx, mat1, mat2), backward -> (None, grad_mat2)
    spmm = SparseDenseMM.apply    (models.py:18, :60)
    create_coo_tensor(fullrowptr, rowptr, colidx, normfact, nrows, ncols)   (sampler.py:139)

What changed underneath: the tensor returned by ``create_coo_tensor`` carries its
CSR (row pointer, int32 column ids, values) as a Python attribute, so ``spmm``
never rebuilds CSR from COO (the reference does on every call,
cuda_spmm.cu:620-667), and backward multiplies by a cached CSR of A^T built once
on the device instead of ``mat1.transpose(0,1).coalesce()`` per call
(custom_sparse_ops.py:34) - or, for sparse layers, needs no transposed index at all
(``BACKWARD``).  Sparse tensors from elsewhere are converted on first
use.  There is no CPU fallback: CPU operands raise.
"""
from __future__ import annotations

import torch

from . import _native

spmm_cpp = _native.extension()

spmm_forward_time = 0.0
spmm_backward_time = 0.0

# How ``SparseDenseMM.backward`` computes dX = A^T.G (reference custom_sparse_ops.py:30-37):
#   "index"    CSR of A^T built once per adjacency on the device, then the gather kernel (deterministic bits)
#   "scatter"  transpose-free: vector reductions into a zero-filled dX straight from A's CSR (sum order not fixed,
#              like the reference's atomicAdd kernel cuda_spmm.cu:205-209)
#   "auto"     scatter for short-row layers (top LADIES layer, sparse graphs: the index build costs more than the
#              product), index for the dense LADIES blocks (L2 reductions are ~2.3x slower than L2 reads there) -
#              measured A/B in profiles/r2_backward_ab.md.  An adjacency whose A^T index already exists (prebuilt by
#              the sampler / prefetcher, off the training stream) always uses it.
BACKWARD = "auto"
SCATTER_MEAN_ROW = 96          # "short-row": nnz < SCATTER_MEAN_ROW * nrows

_ATTR = "_gnn_b200_adj"


class Adjacency:
    """Device-resident CSR of one layer adjacency (+ lazily the CSR of its transpose)."""
    __slots__ = ("rowptr", "colidx", "vals", "rowidx", "nrows", "ncols", "_t")

    def __init__(self, rowptr, colidx, vals, nrows, ncols, rowidx=None):
        self.rowptr, self.colidx, self.vals = rowptr, colidx, vals
        self.rowidx = rowidx                 # int32 row id per entry (build_adj / csr_transpose emit it) or None
        self.nrows, self.ncols = int(nrows), int(ncols)
        self._t = None

    @property
    def nnz(self) -> int:
        return int(self.vals.numel())

    def matmul(self, dense: torch.Tensor, padded_rows: bool = False) -> torch.Tensor:
        """A . X   (forward, reference custom_sparse_ops.py:23).  ``padded_rows``: the result's rows start on 128-byte
        lines (a [M, D] view of a [M, ceil32(D)] buffer) for a consumer that reads them with 128-bit loads."""
        return spmm_cpp.csr_spmm(self.rowptr, self.colidx, self.vals, self.nrows, self.ncols, dense, self.rowidx, padded_rows)

    def transpose(self) -> "Adjacency":
        if self._t is None:
            t_rowptr, t_colidx, t_vals, t_rowidx = spmm_cpp.csr_transpose(self.rowptr, self.colidx, self.vals, self.nrows, self.ncols)
            self._t = Adjacency(t_rowptr, t_colidx, t_vals, self.ncols, self.nrows, t_rowidx)
        return self._t

    def device_tensors(self):
        """Every device tensor this adjacency (and its cached transpose) keeps alive - what a consumer on ANOTHER stream
        must ``record_stream`` (the caching allocator hands a freed block back to the producing stream at once)."""
        ts = [t for t in (self.rowptr, self.colidx, self.vals, self.rowidx) if t is not None]
        if self._t is not None:
            ts += self._t.device_tensors()
        return ts

    def short_rows(self) -> bool:
        return self.nnz < SCATTER_MEAN_ROW * max(self.nrows, 1)

    def matmul_t(self, dense: torch.Tensor, mode: str = None) -> torch.Tensor:
        """A^T . G   (backward, reference custom_sparse_ops.py:34); ``mode`` overrides the module-level BACKWARD."""
        mode = mode or BACKWARD
        if mode == "scatter" or (mode == "auto" and self._t is None and self.short_rows()):
            return spmm_cpp.csr_spmm_t(self.rowptr, self.colidx, self.vals, self.nrows, self.ncols, dense, self.rowidx)
        return self.transpose().matmul(dense)

    def gather_matmul(self, xrows: torch.Tensor, feat_dim: int) -> torch.Tensor:
        """A . gather(xrows) without materialising the gathered rows (main.py:129-134 + models.py:18)."""
        return spmm_cpp.gather_spmm(self.rowptr, self.colidx, self.vals, self.nrows, self.ncols, int(feat_dim), xrows, self.rowidx)


def adjacency_of(mat1: torch.Tensor) -> Adjacency:
    """CSR attached to a sparse COO tensor; built (and attached) on first use for foreign tensors."""
    adj = getattr(mat1, _ATTR, None)
    if adj is None:
        if not mat1.is_cuda:
            raise RuntimeError("sparseMat must be a CUDA tensor (no CPU path in gnn_b200)")
        if not mat1.is_sparse:
            raise RuntimeError("sparseMat must be a sparse COO tensor")
        if mat1.dtype != torch.float32:
            raise RuntimeError("sparseMat must be float32")
        rowptr, colidx = spmm_cpp.coo_to_csr(mat1)          # raises unless coalesced (spmm.cpp:12-13)
        adj = Adjacency(rowptr, colidx, mat1._values().contiguous(), mat1.shape[0], mat1.shape[1])
        try:
            setattr(mat1, _ATTR, adj)
        except Exception:
            pass
    return adj


def _check_dense(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    # the reference requires is_contiguous() (spmm.cpp:14-15); row-padded buffers (unit inner stride) are also
    # accepted because the gathered input-feature buffer keeps 16-byte-aligned rows
    if t.dim() != 2 or not (t.is_contiguous() or (t.stride(1) == 1 and t.stride(0) >= t.shape[1])):
        raise RuntimeError(f"{name} must be contiguous")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32")


class SparseDenseMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mat1, mat2):
        ctx.save_for_backward(mat1)
        adj = adjacency_of(mat1)
        ctx.adj = adj
        _check_dense(mat2, "denseMat")
        return adj.matmul(mat2)

    @staticmethod
    def backward(ctx, grad_output):
        adj = ctx.adj
        grad_mat2 = adj.matmul_t(grad_output.contiguous())
        return None, grad_mat2


spmm = SparseDenseMM.apply


def create_coo_tensor(fullrowptr, rowptr, colidx, normfact, nrows, ncols):
    """Drop-in for ``spmm_cpp.create_coo_tensor`` (reference spmm.cpp:44-50): returns the coalesced
    sparse COO tensor [nrows, ncols] (int64 indices, fp32 values) with its CSR attached."""
    coo, col32, row32 = spmm_cpp.build_adj(fullrowptr, rowptr, colidx, normfact, int(nrows), int(ncols))
    setattr(coo, _ATTR, Adjacency(rowptr, col32, coo._values(), nrows, ncols, row32))
    return coo
