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
def patch_reference_models(ref_models):
    """``import models; gnn_b200.models.patch_reference_models(models)``: the reference's layer classes keep their
    constructors, parameters, spmm call, concat and linears; only the tail after them becomes ``elu_rownorm``."""
    def sage_forward(self, x, adj, sampled_nodes):
        if self.order > 0:
            agg = custom_sparse_ops.spmm(adj, x)
            pre = torch.cat([self.linearB(x[sampled_nodes]), self.linearW(agg)], 1)
        else:
            pre = self.linearW(x)
        return layer_tail(pre, self.scale, self.offset)

    def gcn_forward(self, x, adj):
        pre = self.linear(custom_sparse_ops.spmm(adj, x) if self.order > 0 else x)
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

    def __init__(self, sage: bool, n_in: int, n_out: int, order: int, fused: bool = True, spmm=None):
        super().__init__()
        self.sage, self.order, self.fused = sage, order, fused
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

    def __init__(self, sage: bool, nfeat: int, nhid: int, orders, dropout: float, fused: bool = True, spmm=None):
        super().__init__()
        widths = [nfeat] + [nhid * ((1 + o) if sage else 1) for o in orders]
        self.nhid = widths[-1]
        self.gcs = nn.ModuleList(Conv(sage, widths[i], nhid, orders[i], fused, spmm) for i in range(len(orders)))
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


def build_model(kind: str, nfeat: int, nhid: int, orders, num_classes: int, dropout: float = 0.1, fused: bool = True, spmm=None):
    """``kind``: "graphsage" or "gcn" (reference main.py --model)."""
    enc = Encoder(kind == "graphsage", nfeat, nhid, list(orders), dropout, fused, spmm)
    return GNN(enc, num_classes, dropout)
