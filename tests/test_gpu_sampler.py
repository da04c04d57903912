"""Device LADIES layer construction (gnn_b200/gpu_sampler.py) vs the host mirror, which is itself bit-identical to the
unmodified reference sampler (tests/test_sampler_golden.py): every hand-off array, remap and node list must be equal."""
import os

import numpy as np
import pytest
import torch

import oracle
from gnn_b200 import graphgen, sampler

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gs():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gnn_b200 import gpu_sampler
    return gpu_sampler


@pytest.mark.parametrize("one_call", [True, "device_compact", False],
                         ids=["one_call_layers", "one_call_device_compacted_support", "stepwise_layers"])
@pytest.mark.parametrize("shape_name,orders,samp,batch,seeds", [
    ("small", [1, 1, 1], 2048, 256, [4321, 7]),
    ("tiny", [1, 0, 1], 64, 16, [11]),
    ("cora", [1, 1], 512, 256, [1234, 1235]),
])
def test_device_sampler_bit_identical_to_host(gs, shape_name, orders, samp, batch, seeds, one_call, monkeypatch):
    if one_call == "device_compact":      # the path graphs of a million nodes and more take: gnn_support_compact on the device
        monkeypatch.setattr(gs, "DEVICE_COMPACT_MIN_NODES", 0)
    shape = graphgen.SHAPES[shape_name]
    g = graphgen.generate(shape, seed=0)
    dg = gs.DeviceGraph(g.indptr, g.indices, "cuda")
    rng = np.random.Generator(np.random.PCG64(5))
    for seed in seeds:
        bn = g.train_nodes[rng.permutation(g.train_nodes.size)[:batch]]
        ref = sampler.ladies_sample(seed, bn, [samp] * 5, shape.num_nodes, g.indptr, g.indices, orders)
        got = gs.ladies_sample_device(seed, bn, [samp] * 5, dg, orders, one_call_layers=bool(one_call))
        assert np.array_equal(got.input_nodes, ref.input_nodes)
        assert len(got.layers) == len(ref.layers)
        for li, (a, b) in enumerate(zip(got.layers, ref.layers)):
            if b is None:
                assert a is None and got.adjs[li] is None
                continue
            assert (a.nrows, a.ncols) == (b.nrows, b.ncols)
            assert np.array_equal(a.fullrowptr.cpu().numpy(), b.fullrowptr)
            assert np.array_equal(a.rowptr.cpu().numpy(), b.rowptr)
            assert a.colidx.dtype == torch.int16 and np.array_equal(a.colidx.cpu().numpy(), b.colidx)
            assert np.array_equal(a.normfact.cpu().numpy().view(np.uint32), b.normfact.view(np.uint32))
            assert np.array_equal(got.sampled_nodes[li], ref.sampled_nodes[li])
            rows, cols, vals = oracle.build_adj(b.fullrowptr, b.rowptr, b.colidx, b.normfact, b.nrows)
            adj = got.adjs[li]
            assert np.array_equal(adj._indices().cpu().numpy(), np.stack([rows, cols]))
            assert np.array_equal(adj._values().cpu().numpy().view(np.uint32), vals.view(np.uint32))
        # the membership bitmap is left clean for the next call
        assert int((dg.member_bits != 0).sum().item()) == 0


def test_device_sampler_matches_reference_golden(gs, golden_dir):
    """Directly against arrays captured from the unmodified reference sampler (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(golden_dir, "cora_gcn.npz"))
    shape = graphgen.SHAPES["cora"]
    g = graphgen.generate(shape, seed=0)
    dg = gs.DeviceGraph(g.indptr, g.indices, "cuda")
    for si, seed in enumerate(z["seeds"]):
        got = gs.ladies_sample_device(int(seed), z[f"s{si}_batch_nodes"], [int(z["samp_num"])] * 5, dg, [int(o) for o in z["orders"]])
        for li, a in enumerate(got.layers):
            lp = f"s{si}_l{li}_"
            for k in ["fullrowptr", "rowptr", "colidx", "normfact"]:
                assert np.array_equal(getattr(a, k).cpu().numpy(), z[lp + k]), k
            assert np.array_equal(got.sampled_nodes[li], z[lp + "sampled_nodes"])
        assert got.input_nodes.size == int(z[f"s{si}_n0"])


@pytest.mark.parametrize("tag", ["gcn", "sage"])
def test_device_sampler_locality_sampling_matches_reference(gs, golden_dir, tag):
    """--locality_sampling on the device sampler: scale factors 2, 1.5 and 16 (the reference truncates the scaled integer
    counts, sampler.py:119-121 on scipy's int64 column norm) against arrays captured from the unmodified reference."""
    z = np.load(os.path.join(golden_dir, "locality_small.npz"))
    shape = graphgen.SHAPES[str(z[tag + "_shape"])]
    g = graphgen.generate(shape, seed=0)
    dg = gs.DeviceGraph(g.indptr, g.indices, "cuda")
    orders = [int(o) for o in z[tag + "_orders"]]
    sets = [z[f"{tag}_set{i}"] for i in range(len(orders))]
    for ci in range(3):
        c = f"{tag}_c{ci}_"
        got = gs.ladies_sample_device(int(z[c + "seed"]), z[c + "batch_nodes"], [int(z[tag + "_samp_num"])] * 5, dg, orders,
                                      skewed_sampling_nodes=sets, scale_factor=float(z[c + "scale_factor"]))
        assert got.input_nodes.size == int(z[c + "n0"])
        for li, a in enumerate(got.layers):
            for k in ["fullrowptr", "rowptr", "colidx", "normfact"]:
                assert np.array_equal(getattr(a, k).cpu().numpy(), z[c + f"l{li}_{k}"]), (tag, ci, li, k)
            assert np.array_equal(got.sampled_nodes[li], z[c + f"l{li}_sampled_nodes"])
