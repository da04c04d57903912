"""GPU tests of the reference-facing Python API (custom_sparse_ops) - read like the tests the
reference never shipped: create_coo_tensor -> spmm -> autograd backward, on sampler output."""
import numpy as np
import pytest
import torch

import oracle
from gnn_b200 import graphgen, sampler

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def cso():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import custom_sparse_ops
    return custom_sparse_ops


@pytest.fixture(scope="module")
def mb():
    shape = graphgen.SHAPES["small"]
    g = graphgen.generate(shape, seed=0)
    return sampler.ladies_sample(99, g.train_nodes[:128], [1024] * 3, shape.num_nodes, g.indptr, g.indices, [1, 1, 1])


def _upload(layer, dev="cuda"):
    return (torch.from_numpy(layer.fullrowptr).to(dev), torch.from_numpy(layer.rowptr).to(dev),
            torch.from_numpy(layer.colidx).to(dev), torch.from_numpy(layer.normfact).to(dev), layer.nrows, layer.ncols)


def test_create_coo_tensor_contract(cso, mb):
    for layer in mb.layers:
        a = cso.create_coo_tensor(*_upload(layer))
        assert a.is_sparse and a.is_cuda and a.is_coalesced()
        assert tuple(a.shape) == (layer.nrows, layer.ncols)
        assert a._indices().dtype == torch.int64 and a._values().dtype == torch.float32
        rows, cols, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact, layer.nrows)
        assert np.array_equal(a._indices().cpu().numpy(), np.stack([rows, cols]))
        assert np.array_equal(a._values().cpu().numpy().view(np.uint32), vals.view(np.uint32))
        # the raw extension entry point keeps the reference signature and result
        b = cso.spmm_cpp.create_coo_tensor(*_upload(layer))
        assert torch.equal(b._indices(), a._indices()) and torch.equal(b._values(), a._values())
        # coalesce() of an already coalesced tensor is the identity, as the reference relies on
        assert torch.equal(a.coalesce()._values(), a._values())
        # dense view agrees with torch's own interpretation of the COO tensor
        if layer.nrows * layer.ncols < 4_000_000:
            dense = torch.zeros(layer.nrows, layer.ncols, device="cuda")
            dense[a._indices()[0], a._indices()[1]] = a._values()
            assert torch.equal(a.to_dense(), dense)


@pytest.mark.parametrize("D", [602, 128, 33])
def test_spmm_autograd_matches_oracle_and_torch_sparse(cso, mb, D):
    rng = np.random.Generator(np.random.PCG64(D))
    for layer in mb.layers:
        a = cso.create_coo_tensor(*_upload(layer))
        X = rng.standard_normal((layer.ncols, D)).astype(np.float32)
        G = rng.standard_normal((layer.nrows, D)).astype(np.float32)
        x = torch.from_numpy(X).cuda().requires_grad_(True)
        y = cso.spmm(a, x)
        y.backward(torch.from_numpy(G).cuda())
        _, _, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact, layer.nrows)
        yref = oracle.spmm_f64acc(layer.rowptr, layer.colidx32, vals, layer.nrows, X)
        gref = oracle.spmm_t_f64acc(layer.rowptr, layer.colidx32, vals, layer.nrows, layer.ncols, G)
        assert oracle.rel_err(y.detach().cpu().numpy(), yref)[0] <= TOL
        assert oracle.rel_err(x.grad.cpu().numpy(), gref)[0] <= TOL
        # torch.sparse on the same tensor (the reference's commented alternative, custom_sparse_ops.py:25,36)
        x2 = torch.from_numpy(X).cuda().requires_grad_(True)
        y2 = torch.sparse.mm(a, x2)
        y2.backward(torch.from_numpy(G).cuda())
        assert oracle.rel_err(y.detach().cpu().numpy(), y2.detach().double().cpu().numpy())[0] <= 1e-4
        assert oracle.rel_err(x.grad.cpu().numpy(), x2.grad.double().cpu().numpy())[0] <= 1e-4


def test_foreign_coo_and_reference_entry_points(cso, mb):
    layer = mb.layers[1]
    rows, cols, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact, layer.nrows)
    a = torch.sparse_coo_tensor(torch.from_numpy(np.stack([rows, cols])), torch.from_numpy(vals),
                                (layer.nrows, layer.ncols)).cuda().coalesce()
    X = torch.randn(layer.ncols, 96, device="cuda")
    yref = oracle.spmm_f64acc(layer.rowptr, layer.colidx32, vals, layer.nrows, X.cpu().numpy())
    for fn in (cso.spmm_cpp.spmm_load_balance, cso.spmm_cpp.spmm_naive, cso.spmm):
        assert oracle.rel_err(fn(a, X).cpu().numpy(), yref)[0] <= TOL
    # the reference's backward expression works unchanged on the extension (custom_sparse_ops.py:34)
    G = torch.randn(layer.nrows, 96, device="cuda")
    dx = cso.spmm_cpp.spmm_load_balance(a.transpose(0, 1).coalesce(), G.contiguous())
    gref = oracle.spmm_t_f64acc(layer.rowptr, layer.colidx32, vals, layer.nrows, layer.ncols, G.cpu().numpy())
    assert oracle.rel_err(dx.cpu().numpy(), gref)[0] <= TOL


def test_preconditions_raise(cso, mb):
    layer = mb.layers[2]
    a = cso.create_coo_tensor(*_upload(layer))
    X = torch.randn(layer.ncols, 64, device="cuda")
    with pytest.raises(RuntimeError):
        cso.spmm(a, X.t().contiguous().t())                       # not contiguous (spmm.cpp:14-15)
    with pytest.raises(RuntimeError):
        cso.spmm(a, X.cpu())                                      # not CUDA (spmm.cpp:10-11)
    unco = torch.sparse_coo_tensor(torch.tensor([[0, 0], [1, 1]]), torch.ones(2), (2, 2)).cuda()
    with pytest.raises(RuntimeError):
        cso.spmm_cpp.spmm_load_balance(unco, torch.ones(2, 2, device="cuda"))   # not coalesced (spmm.cpp:12-13)
    with pytest.raises(RuntimeError):
        cso.spmm(a, torch.randn(layer.ncols + 1, 8, device="cuda"))            # shape mismatch


def test_gcn_style_two_layer_grad_flow(cso, mb):
    """models.py:60-64 shape of use: spmm -> linear -> spmm, gradients reach the first weight."""
    l0, l1 = mb.layers[0], mb.layers[1]
    a0, a1 = cso.create_coo_tensor(*_upload(l0)), cso.create_coo_tensor(*_upload(l1))
    lin = torch.nn.Linear(100, 32).cuda()
    x = torch.randn(l0.ncols, 100, device="cuda")
    h = torch.nn.functional.elu(lin(cso.spmm(a0, x)))
    out = cso.spmm(a1, h)
    out.square().mean().backward()
    assert lin.weight.grad is not None and torch.isfinite(lin.weight.grad).all() and lin.weight.grad.abs().sum() > 0


def test_cuda_graph_capture_and_stream_semantics(cso, mb):
    """No host sync, no legacy-stream launch, no raw cudaMalloc on the path: forward + transpose + backward can be
    captured in a CUDA graph on a side stream and replayed (the reference issues 5 device syncs per call)."""
    layer = mb.layers[1]
    a = cso.create_coo_tensor(*_upload(layer))
    adj = cso.adjacency_of(a)
    x = torch.randn(layer.ncols, 128, device="cuda")
    go = torch.randn(layer.nrows, 128, device="cuda")
    y_ref, dx_ref = adj.matmul(x), adj.matmul_t(go, mode="index")
    adj._t = None
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        y = adj.matmul(x)
        dx = adj.matmul_t(go, mode="index")          # includes the A^T build
        adj._t = None
        dxs = adj.matmul_t(go, mode="scatter")       # transpose-free: zero fill + reductions, capturable too
    for _ in range(3):
        x.normal_()
        go.normal_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(y, adj.matmul(x)) and torch.equal(dx, adj.matmul_t(go, mode="index"))
        assert torch.allclose(dxs, dx, rtol=1e-4, atol=1e-5)
    assert y_ref.shape == y.shape and dx_ref.shape == dx.shape


def test_feature_store_prefetch_matches_gather(cso):
    from gnn_b200 import gather, graphgen
    shape = graphgen.SHAPES["tiny"]
    feats = graphgen.features(shape, seed=1)
    n = shape.num_nodes
    top = np.arange(0, n, 3)
    did = np.full(n, -1, dtype=np.int64)
    did[top] = 0
    idx = np.arange(n, dtype=np.int64)
    idx[top] = np.arange(top.size)
    store = gather.FeatureStore(torch.from_numpy(feats), [top], did, idx, [0], 0, torch.device("cuda", 0))
    nodes = torch.from_numpy(np.sort(np.random.Generator(np.random.PCG64(4)).choice(n, 200, replace=False))).cuda()
    side = torch.cuda.Stream()
    out, ev = store.prefetch(nodes, side)
    torch.cuda.current_stream().wait_event(ev)
    assert np.array_equal(out.cpu().numpy(), feats[nodes.cpu().numpy()])
    assert torch.equal(out, store.gather(nodes))
    masks = [(did[nodes.cpu().numpy()] == 0)]
    ref = store.gather_from_reference_tuple(masks, ~masks[0], [idx[nodes.cpu().numpy()[masks[0]]]], nodes.cpu().numpy()[~masks[0]], 200)
    assert torch.equal(ref, out)
    store.close()


def test_feature_store_fused_gather_spmm(cso):
    """Hybrid fused gather+SpMM: local rows read in place, remote (other shard / host) rows staged once."""
    from gnn_b200 import gather, graphgen, placement, sampler
    shape = graphgen.SHAPES["small"]
    g = graphgen.generate(shape, seed=0)
    feats = graphgen.features(shape, seed=1)
    world = 3
    pl = placement.create_placement(g.to_scipy(np.float64), g.train_nodes, int(0.1 * shape.num_nodes), list(range(world)), 3, alpha=0.0)
    for rank in (0, 2):
        store = gather.FeatureStore(torch.from_numpy(feats), pl.gpu_buffer_group, pl.device_id_of_nodes_group[rank],
                                    pl.idx_of_nodes_on_device_group[rank], list(range(world)), rank, torch.device("cuda", 0))
        mbx = sampler.ladies_sample(77 + rank, g.train_nodes[:128], [1024] * 3, shape.num_nodes, g.indptr, g.indices, [1, 1, 1])
        layer = mbx.layers[0]
        a = cso.create_coo_tensor(*_upload(layer))
        nodes = torch.from_numpy(mbx.input_nodes).cuda()
        y = store.gather_spmm(cso.adjacency_of(a), nodes)
        x = store.gather(nodes)
        assert np.array_equal(x.cpu().numpy(), feats[mbx.input_nodes])
        assert torch.equal(y, cso.adjacency_of(a).matmul(x)), "fused and staged paths must agree bit for bit"
        _, _, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact, layer.nrows)
        yref = oracle.spmm_f64acc(layer.rowptr, layer.colidx32, vals, layer.nrows, feats[mbx.input_nodes])
        assert oracle.rel_err(y.cpu().numpy(), yref)[0] <= TOL
        src = store.remap(nodes)[0].cpu().numpy()
        assert (src == rank).any() and (src != rank).any()
        store.close()


def test_device_prefetcher_matches_direct_path(cso):
    """pipeline.DevicePrefetcher (worker thread + side stream) hands over the same adjacencies and gathered rows as
    the direct calls, for several minibatches in flight."""
    from gnn_b200 import gather, graphgen, pipeline, sampler
    shape = graphgen.SHAPES["small"]
    g = graphgen.generate(shape, seed=0)
    feats = graphgen.features(shape, seed=1)
    n = shape.num_nodes
    top = np.arange(0, n, 4)
    did = np.full(n, -1, dtype=np.int64)
    did[top] = 0
    idx = np.arange(n, dtype=np.int64)
    idx[top] = np.arange(top.size)
    store = gather.FeatureStore(torch.from_numpy(feats), [top], did, idx, [0], 0, torch.device("cuda", 0))
    mbs = [sampler.ladies_sample(900 + i, g.train_nodes[i * 64:(i + 1) * 64], [512] * 3, n, g.indptr, g.indices, [1, 1, 1]) for i in range(5)]
    pre = pipeline.DevicePrefetcher(store, cso.create_coo_tensor, torch.device("cuda", 0), depth=2, prebuild_transpose=True)
    for mbx in mbs:
        pre.submit(pipeline.PinnedMinibatch(mbx))
    for mbx in mbs:
        adjs, x0, counts = pre.get()
        assert np.array_equal(x0.cpu().numpy(), feats[mbx.input_nodes])
        assert int(counts.sum().item()) == mbx.input_nodes.size
        for a, layer in zip(adjs, mbx.layers):
            rows, cols, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact, layer.nrows)
            assert np.array_equal(a._indices().cpu().numpy(), np.stack([rows, cols]))
            assert np.array_equal(a._values().cpu().numpy().view(np.uint32), vals.view(np.uint32))
        y = cso.spmm(adjs[0], x0)
        yref = oracle.spmm_f64acc(mbx.layers[0].rowptr, mbx.layers[0].colidx32, oracle.build_adj(
            mbx.layers[0].fullrowptr, mbx.layers[0].rowptr, mbx.layers[0].colidx, mbx.layers[0].normfact, mbx.layers[0].nrows)[2],
            mbx.layers[0].nrows, feats[mbx.input_nodes])
        assert oracle.rel_err(y.cpu().numpy(), yref)[0] <= TOL
        assert cso.adjacency_of(adjs[1])._t is not None and cso.adjacency_of(adjs[0])._t is None
    pre.close()
    store.close()
