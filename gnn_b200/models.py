"""Layer tail fused into one CUDA kernel per direction and the thin model shells that
carry it.

The reference's two layer types end with the same elementwise tail (reference models.py:21-25 and :61-64):

    out = F.elu(feat); mean = out.mean(1); var = out.var(1, unbiased=False) + 1e-9
    return (out - mean) * self.scale * torch.rsqrt(var) + self.offset

``elu_rownorm(feat, scale, offset)`` is that tail as ONE kernel forward and one backward (gnn_elu_rownorm_*_f32).
Two ways to use it:

* ``patch_reference_models(models)`` - for a checkout that has the reference: swaps the tail of the reference's OWN
  ``GraphSageConvolution`` / ``GraphConvolution`` classes in place; nothing of the reference is restated.
* ``build_model(...)`` - for a box without the reference (the GPU box, bench.py's training metric): a table-driven
  encoder whose parameter names equal the reference's (``encoder.gcs.<i>.linearW.weight`` ... ``linear.bias``), so a
  reference ``state_dict`` loads unchanged (tests/test_gpu_models.py checks outputs, loss and gradients against
  goldens of the unmodified reference modules).
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import custom_sparse_ops


class EluRowNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feat, scale, offset):
        ext = custom_sparse_ops.spmm_cpp
        feat = feat if feat.stride(-1) == 1 else feat.contiguous()
        y, mean, rstd = ext.elu_rownorm_fwd(feat, scale.contiguous(), offset.contiguous())
        ctx.save_for_backward(feat, scale, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        feat, scale, mean, rstd = ctx.saved_tensors
        ext = custom_sparse_ops.spmm_cpp
        dy = dy if dy.stride(-1) == 1 else dy.contiguous()
        dx, dscale, doffset = ext.elu_rownorm_bwd(dy, feat, scale.contiguous(), mean, rstd)
        return dx, dscale, doffset


elu_rownorm = EluRowNorm.apply


class TcLinear(torch.autograd.Function):
    """``x[rows] @ W.T + b`` on the tcgen05 tensor cores in 3xTF32 (csrc/linear_tc.cuh): fp32 in, fp32 out, agrees with
    an fp64 product to ~1e-6.  ``rows`` (int64 or None) is the ``x[sampled_nodes]`` gather of reference models.py:19,
    done by the operand loader.  Backward: dX = dY.W (same kernel on W^T), dW = dY^T.x[rows] (split over rows,
    fixed-order sum), db = column sums."""

    @staticmethod
    def forward(ctx, x, rows, W, b):
        ext = custom_sparse_ops.spmm_cpp
        x = x if x.stride(-1) == 1 else x.contiguous()
        need_dx = ctx.needs_input_grad[0]
        w_nk, w_kn = ext.linear_split_weights(W.detach(), need_dx)
        M = x.shape[0] if rows is None else rows.shape[0]
        out = torch.empty(M, W.shape[0], device=x.device, dtype=torch.float32)
        ext.linear_tf32x3(x, rows, w_nk, W.shape[1], b, out)
        ctx.save_for_backward(x, rows, w_kn)
        ctx.has_bias = b is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, rows, w_kn = ctx.saved_tensors
        ext = custom_sparse_ops.spmm_cpp
        dy = dy if dy.stride(-1) == 1 else dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dxr = torch.empty(dy.shape[0], x.shape[1], device=x.device, dtype=torch.float32)
            ext.linear_tf32x3(dy, None, w_kn, dy.shape[1], None, dxr)
            dx = dxr if rows is None else torch.zeros_like(x).index_add_(0, rows, dxr)
        dW = db = None
        if ctx.needs_input_grad[2] or (ctx.has_bias and ctx.needs_input_grad[3]):
            dW, db = ext.linear_wgrad_tf32x3(dy, x, rows, ctx.has_bias)        # db: column sums of dy from the same pass
        return dx, None, dW, (db if ctx.has_bias else None)


class SageLinears(torch.autograd.Function):
    """``cat[linearB(x[rows]), linearW(agg)]`` of reference models.py:19 without the index kernel and without the
    concat: both products write their column slice of one output buffer."""

    @staticmethod
    def forward(ctx, x, rows, agg, WB, bB, WW, bW):
        ext = custom_sparse_ops.spmm_cpp
        x = x if x.stride(-1) == 1 else x.contiguous()
        agg = agg if agg.stride(-1) == 1 else agg.contiguous()
        n, K = WB.shape
        sB, sBt = ext.linear_split_weights(WB.detach(), ctx.needs_input_grad[0])
        sW, sWt = ext.linear_split_weights(WW.detach(), ctx.needs_input_grad[2])
        pre = torch.empty(agg.shape[0], 2 * n, device=x.device, dtype=torch.float32)
        ext.linear_tf32x3(x, rows, sB, K, bB, pre[:, :n])
        ext.linear_tf32x3(agg, None, sW, K, bW, pre[:, n:])
        ctx.save_for_backward(x, rows, agg, sBt, sWt)
        return pre

    @staticmethod
    def backward(ctx, dpre):
        x, rows, agg, sBt, sWt = ctx.saved_tensors
        ext = custom_sparse_ops.spmm_cpp
        dpre = dpre if dpre.stride(-1) == 1 else dpre.contiguous()
        n = dpre.shape[1] // 2
        dB, dWv = dpre[:, :n], dpre[:, n:]
        dx = dagg = None
        if ctx.needs_input_grad[0]:
            dxr = torch.empty(dpre.shape[0], x.shape[1], device=x.device, dtype=torch.float32)
            ext.linear_tf32x3(dB, None, sBt, n, None, dxr)
            dx = torch.zeros_like(x).index_add_(0, rows, dxr)
        if ctx.needs_input_grad[2]:
            dagg = torch.empty(dpre.shape[0], agg.shape[1], device=x.device, dtype=torch.float32)
            ext.linear_tf32x3(dWv, None, sWt, n, None, dagg)
        dWB, dbB = ext.linear_wgrad_tf32x3(dB, x, rows, True)
        dWW, dbW = ext.linear_wgrad_tf32x3(dWv, agg, None, True)
        return dx, None, dagg, dWB, dbB, dWW, dbW


class SageLayer(torch.autograd.Function):
    """A whole GraphSAGE layer up to its pre-activation, reference models.py:18-19:
    ``agg = spmm(adj, x); pre = cat[linearB(x[rows]), linearW(agg)]``.
    Owning the SpMM as well lets the backward put both contributions to dX into ONE buffer: the SpMM backward writes
    dX = A^T.dagg, and the dX of linearB is added onto its rows ``rows`` by the GEMM's own epilogue (L2 reductions) -
    no zero-filled buffer, no index_add, no gradient-accumulation pass."""

    @staticmethod
    def forward(ctx, x, mat1, rows, WB, bB, WW, bW):
        ext = custom_sparse_ops.spmm_cpp
        adj = custom_sparse_ops.adjacency_of(mat1)
        custom_sparse_ops._check_dense(x, "denseMat")
        n, K = WB.shape
        agg = adj.matmul(x, padded_rows=True)
        need_dx = ctx.needs_input_grad[0]
        sB, sBt, sW, sWt = ext.linear_split_weights2(WB.detach(), WW.detach(), need_dx)
        pre = torch.empty(agg.shape[0], 2 * n, device=x.device, dtype=torch.float32)
        ext.linear_tf32x3(x, rows, sB, K, bB, pre[:, :n])
        ext.linear_tf32x3(agg, None, sW, K, bW, pre[:, n:])
        ctx.adj = adj
        ctx.save_for_backward(x, rows, agg, sBt, sWt)
        return pre

    @staticmethod
    def backward(ctx, dpre):
        x, rows, agg, sBt, sWt = ctx.saved_tensors
        ext = custom_sparse_ops.spmm_cpp
        dpre = dpre if dpre.stride(-1) == 1 else dpre.contiguous()
        n = dpre.shape[1] // 2
        dB, dWv = dpre[:, :n], dpre[:, n:]
        dx = None
        if ctx.needs_input_grad[0]:
            dagg = torch.empty(dpre.shape[0], x.shape[1], device=x.device, dtype=torch.float32)
            ext.linear_tf32x3(dWv, None, sWt, n, None, dagg)
            dx = ctx.adj.matmul_t(dagg)
            ext.linear_tf32x3(dB, None, sBt, n, None, dx, rows, True)      # dx[rows] += dB . WB
        dWB, dbB = ext.linear_wgrad_tf32x3(dB, x, rows, True)          # bias gradients: column sums of dpre from the same pass
        dWW, dbW = ext.linear_wgrad_tf32x3(dWv, agg, None, True)
        return dx, None, None, dWB, dbB, dWW, dbW


def _rows_tensor(rows, device):
    """The reference indexes with whatever the sampler returned (numpy int64 arrays, sampler.py:143); the kernels take an
    int64 tensor on the device."""
    if rows is None or (torch.is_tensor(rows) and rows.device == device and rows.dtype == torch.long):
        return rows
    if torch.is_tensor(rows):
        return rows.to(device=device, dtype=torch.long)
    import numpy as np
    return torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int64)).to(device)


def tc_linear(x, W, b, rows=None):
    return TcLinear.apply(x, _rows_tensor(rows, x.device), W, b)


def sage_layer(x, adj, rows, linearB, linearW):
    return SageLayer.apply(x, adj, _rows_tensor(rows, x.device), linearB.weight, linearB.bias, linearW.weight, linearW.bias)


def sage_linears(x, rows, agg, linearB, linearW):
    return SageLinears.apply(x, _rows_tensor(rows, x.device), agg, linearB.weight, linearB.bias, linearW.weight, linearW.bias)


def layer_tail(feat, scale, offset, fused: bool = True):
    """ELU + per-row standardisation + affine.  ``fused=False`` (or a CPU tensor) evaluates the reference's own torch
    expression - the baseline the fused kernel is measured against."""
    if fused and feat.is_cuda:
        return elu_rownorm(feat, scale, offset)
    out = F.elu(feat)
    mean = out.mean(dim=1, keepdim=True)
    var = out.var(dim=1, unbiased=False, keepdim=True) + 1e-9
    return (out - mean) * scale * torch.rsqrt(var) + offset


# --------------------------------------------------------------------------------------------------------------
# in-place patch of the reference's own classes
# --------------------------------------------------------------------------------------------------------------
def patch_reference_models(ref_models, tc: bool = False):
    """``import models; gnn_b200.models.patch_reference_models(models)``: the reference's layer classes keep their
    constructors, parameters, spmm call, concat and linears; only the tail after them becomes ``elu_rownorm``.
    ``tc=True`` additionally runs the linears (and the gather + concat around them) on the tensor cores."""
    def sage_forward(self, x, adj, sampled_nodes):
        if self.order > 0:
            if tc and x.is_cuda:
                pre = sage_layer(x, adj, sampled_nodes, self.linearB, self.linearW)
            else:
                agg = custom_sparse_ops.spmm(adj, x)
                pre = torch.cat([self.linearB(x[sampled_nodes]), self.linearW(agg)], 1)
        else:
            pre = tc_linear(x, self.linearW.weight, self.linearW.bias) if tc and x.is_cuda else self.linearW(x)
        return layer_tail(pre, self.scale, self.offset)

    def gcn_forward(self, x, adj):
        feat = custom_sparse_ops.spmm(adj, x) if self.order > 0 else x
        pre = tc_linear(feat, self.linear.weight, self.linear.bias) if tc and x.is_cuda else self.linear(feat)
        return layer_tail(pre, self.scale, self.offset)

    ref_models.GraphSageConvolution.forward = sage_forward
    ref_models.GraphConvolution.forward = gcn_forward
    return ref_models


# --------------------------------------------------------------------------------------------------------------
# stand-alone shells (parameter names = the reference's)
# --------------------------------------------------------------------------------------------------------------
class Conv(nn.Module):
    """One layer of either family.  ``sage``: linearW on the aggregate, linearB on the layer's own rows, concatenated
    (reference models.py:6-25); otherwise one ``linear`` on the aggregate (models.py:48-64)."""

    def __init__(self, sage: bool, n_in: int, n_out: int, order: int, fused: bool = True, spmm=None, tc: bool = False):
        super().__init__()
        self.sage, self.order, self.fused, self.tc = sage, order, fused, tc
        self._spmm = spmm
        width = n_out * ((1 + order) if sage else 1)
        if sage:
            self.linearW = nn.Linear(n_in, n_out)
            self.linearB = nn.Linear(n_in, n_out)
        else:
            self.linear = nn.Linear(n_in, n_out)
        self.offset = nn.Parameter(torch.zeros(width))
        self.scale = nn.Parameter(torch.ones(width))

    def forward(self, x, adj, own_rows):
        spmm = self._spmm or custom_sparse_ops.spmm
        if self.tc and x.is_cuda:           # dense linears on the tensor cores (3xTF32), gather and concat fused away
            if self.sage and self.order > 0:
                if self._spmm is None or self._spmm is custom_sparse_ops.spmm:
                    pre = sage_layer(x, adj, own_rows, self.linearB, self.linearW)
                else:                       # a caller-supplied spmm (instrumented, fused gather) keeps its own autograd node
                    pre = sage_linears(x, own_rows, spmm(adj, x), self.linearB, self.linearW)
            else:
                lin = self.linearW if self.sage else self.linear
                pre = tc_linear(spmm(adj, x) if (self.order > 0 and not self.sage) else x, lin.weight, lin.bias)
            return layer_tail(pre, self.scale, self.offset, self.fused)
        if self.sage:
            if self.order > 0:
                pre = torch.cat([self.linearB(x[own_rows]), self.linearW(spmm(adj, x))], 1)
            else:
                pre = self.linearW(x)
        else:
            pre = self.linear(spmm(adj, x) if self.order > 0 else x)
        return layer_tail(pre, self.scale, self.offset, self.fused)


class Encoder(nn.Module):
    """Stack of ``Conv`` layers under the attribute names the reference uses (``gcs``, ``dropout``, ``nhid``)."""

    def __init__(self, sage: bool, nfeat: int, nhid: int, orders, dropout: float, fused: bool = True, spmm=None,
                 tc: bool = False):
        super().__init__()
        widths = [nfeat] + [nhid * ((1 + o) if sage else 1) for o in orders]
        self.nhid = widths[-1]
        self.gcs = nn.ModuleList(Conv(sage, widths[i], nhid, orders[i], fused, spmm, tc) for i in range(len(orders)))
        self.dropout = nn.Dropout(dropout)

    def forward(self, x, adjs, sampled_nodes):
        for layer, adj, rows in zip(self.gcs, adjs, sampled_nodes):
            x = self.dropout(layer(x, adj, rows))
        return x


def GraphSage(nfeat, nhid, orders, dropout, fused=True, spmm=None):
    return Encoder(True, nfeat, nhid, orders, dropout, fused, spmm)


def GCN(nfeat, nhid, orders, dropout, fused=True, spmm=None):
    return Encoder(False, nfeat, nhid, orders, dropout, fused, spmm)


class GNN(nn.Module):
    """Encoder + L2-normalise + dropout + linear head (reference models.py:86-97)."""

    def __init__(self, encoder, num_classes, dropout, inp=None):
        super().__init__()
        self.encoder = encoder
        self.dropout = nn.Dropout(dropout)
        self.linear = nn.Linear(encoder.nhid, num_classes)

    def forward(self, feat, adjs, sampled_nodes):
        return self.linear(self.dropout(F.normalize(self.encoder(feat, adjs, sampled_nodes), p=2, dim=1)))


def build_model(kind: str, nfeat: int, nhid: int, orders, num_classes: int, dropout: float = 0.1, fused: bool = True, spmm=None,
                tc: bool = False):
    """``kind``: "graphsage" or "gcn" (reference main.py --model).  ``tc``: the layers' dense linears run on the tensor
    cores in 3xTF32 (TcLinear / SageLinears) instead of cuBLAS fp32."""
    enc = Encoder(kind == "graphsage", nfeat, nhid, list(orders), dropout, fused, spmm, tc)
    return GNN(enc, num_classes, dropout)
