"""Fused layer epilogue (gnn_b200/models.py) vs the plain torch expression of the reference
(models.py:21-25) in fp32, forward and backward, tolerance 1e-5 relative; and the drop-in GraphSage/GNN against the
golden outputs of the unmodified reference models.py."""
import os

import numpy as np
import pytest
import torch

import oracle
from gnn_b200 import graphgen, sampler

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def models():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gnn_b200 import models as m
    return m


def _reference_tail(feat, scale, offset):
    out = torch.nn.functional.elu(feat)
    mean = out.mean(dim=1).view(out.shape[0], 1)
    var = out.var(dim=1, unbiased=False).view(out.shape[0], 1) + 1e-9
    return (out - mean) * scale * torch.rsqrt(var) + offset


@pytest.mark.parametrize("M,C", [(1, 7), (33, 32), (257, 100), (1000, 512), (513, 1024), (64, 1433), (40, 2048)])
def test_elu_rownorm_matches_torch(models, M, C):
    g = torch.Generator(device="cuda").manual_seed(M * 131 + C)
    feat = (torch.randn(M, C, device="cuda", generator=g) * 2).double()
    scale = (torch.rand(C, device="cuda", generator=g) + 0.5).double()
    offset = torch.randn(C, device="cuda", generator=g).double()
    dy = torch.randn(M, C, device="cuda", generator=g).double()
    f64 = [t.clone().requires_grad_(True) for t in (feat, scale, offset)]
    _reference_tail(*f64).backward(dy)                                      # fp64 arbiter
    f32 = [t.float().requires_grad_(True) for t in (feat, scale, offset)]
    y = models.elu_rownorm(*f32)
    y.backward(dy.float())
    t32 = [t.float().requires_grad_(True) for t in (feat, scale, offset)]
    yt = _reference_tail(*t32)
    yt.backward(dy.float())

    def rel(a, b):
        return ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()
    y64 = _reference_tail(*[t.detach() for t in f64])
    assert rel(y, y64) <= 1e-5
    for ours, theirs, ref in zip(f32, t32, f64):
        e_ours, e_torch = rel(ours.grad, ref.grad), rel(theirs.grad, ref.grad)
        assert e_ours <= max(1e-5, 3 * e_torch), (M, C, e_ours, e_torch)
    # reproducible: column sums are reduced in a fixed order
    f32b = [t.float().requires_grad_(True) for t in (feat, scale, offset)]
    models.elu_rownorm(*f32b).backward(dy.float())
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(f32, f32b))


def test_dropin_graphsage_matches_reference_golden(models, golden_dir):
    import custom_sparse_ops as cso
    z = np.load(os.path.join(golden_dir, "model_sage_tiny.npz"))
    shape = graphgen.SHAPES["tiny"]
    g = graphgen.generate(shape, seed=0)
    feats = graphgen.features(shape, seed=1)
    mb = sampler.ladies_sample(5, g.train_nodes[:24], [64] * 3, shape.num_nodes, g.indptr, g.indices, [1, 1, 1])
    adjs = [cso.create_coo_tensor(torch.from_numpy(l.fullrowptr).cuda(), torch.from_numpy(l.rowptr).cuda(), torch.from_numpy(l.colidx).cuda(),
                                  torch.from_numpy(l.normfact).cuda(), l.nrows, l.ncols) for l in mb.layers]
    enc = models.GraphSage(nfeat=shape.feat_dim, nhid=16, orders=[1, 1, 1], dropout=0.0)
    net = models.GNN(encoder=enc, num_classes=shape.num_classes, dropout=0.0, inp=shape.feat_dim)
    net.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_")})     # same parameter names
    net.cuda().train()
    out = net(torch.from_numpy(feats[mb.input_nodes]).cuda(), adjs, mb.sampled_nodes)
    assert np.allclose(out.detach().cpu().numpy(), z["out"], rtol=2e-5, atol=2e-6)
    labels = torch.nn.functional.one_hot(torch.from_numpy(graphgen.labels(shape, 3)[mb.batch_nodes]), shape.num_classes).float().cuda()
    w = torch.full((out.shape[0], 1), 1.0 / out.shape[0], device="cuda")
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, labels, weight=w, reduction="sum")
    assert abs(loss.item() - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))
    loss.backward()
    for name, p in net.named_parameters():
        ref = z["g_" + name]
        got = p.grad.cpu().numpy()
        assert np.allclose(got, ref, rtol=2e-4, atol=2e-6), name
