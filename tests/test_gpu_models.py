"""Fused layer epilogue (gnn_b200/models.py) vs the plain torch expression of the reference
(models.py:21-25) in fp32, forward and backward, tolerance 1e-5 relative; and the drop-in GraphSage/GNN against the
golden outputs of the unmodified reference models.py."""
import os

import numpy as np
import pytest
import torch

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


def test_training_loop_converges_on_tiny_graph(models):
    """End-to-end sanity in the reference's calling pattern (main.py:118-170): device LADIES sampler -> FeatureStore
    gather -> drop-in GraphSage/GNN on custom_sparse_ops.spmm -> BCE loss -> backward -> clip -> Adam; the loss of a
    fixed evaluation minibatch must fall."""
    import custom_sparse_ops as cso
    from gnn_b200 import gather, gpu_sampler, harness
    shape = graphgen.SHAPES["tiny"]
    g = graphgen.generate(shape, seed=0)
    feats = graphgen.features(shape, seed=1)
    # learnable labels: class = argmax of a fixed random projection of the node's own features
    rng = np.random.Generator(np.random.PCG64(0))
    labels = (feats @ rng.standard_normal((shape.feat_dim, shape.num_classes))).argmax(1)
    n = shape.num_nodes
    top = np.arange(0, n, 2)
    did = np.full(n, -1, dtype=np.int64)
    did[top] = 0
    idx = np.arange(n, dtype=np.int64)
    idx[top] = np.arange(top.size)
    dev = torch.device("cuda", 0)
    store = gather.FeatureStore(torch.from_numpy(feats), [top], did, idx, [0], 0, dev)
    dg = gpu_sampler.DeviceGraph(g.indptr, g.indices, dev)
    torch.manual_seed(0)
    enc = models.GraphSage(nfeat=shape.feat_dim, nhid=32, orders=[1, 1], dropout=0.0)
    net = models.GNN(encoder=enc, num_classes=shape.num_classes, dropout=0.0, inp=shape.feat_dim).to(dev)
    opt = torch.optim.Adam(net.parameters(), lr=0.01)

    def loss_of(mb):
        x0 = store.gather(torch.from_numpy(mb.input_nodes).to(dev))
        out = net(x0, mb.adjs, mb.sampled_nodes)
        y = torch.nn.functional.one_hot(torch.from_numpy(labels[mb.batch_nodes]), shape.num_classes).float().to(dev)
        return harness.bce_loss(out, y)

    eval_mb = gpu_sampler.ladies_sample_device(999, g.train_nodes[:64], [128] * 3, dg, [1, 1])
    net.eval()
    with torch.no_grad():
        before = loss_of(eval_mb).item()
    net.train()
    perm = np.random.Generator(np.random.PCG64(1)).permutation(g.train_nodes.size)
    for it in range(60):
        bn = g.train_nodes[perm[(it * 32) % (perm.size - 32):][:32]]
        mb = gpu_sampler.ladies_sample_device(1000 + it, bn, [128] * 3, dg, [1, 1])
        opt.zero_grad()
        loss = loss_of(mb)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 5)
        opt.step()
    net.eval()
    with torch.no_grad():
        after = loss_of(eval_mb).item()
    store.close()
    assert np.isfinite(after) and after < 0.8 * before, (before, after)
