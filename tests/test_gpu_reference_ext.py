"""Pin the oracle (and the CUDA path) against the reference's OWN CUDA extension, compiled
unmodified into oracle/_ref/ (oracle/build_ref.py) and run on the B200.  Skipped when the
compiled reference is not in the snapshot."""
import numpy as np
import pytest
import torch

import oracle
from gnn_b200 import graphgen, sampler

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle import build_ref
    mod = build_ref.load_ref()
    if mod is None:
        pytest.skip("oracle/_ref/spmm_ref.so not present")
    return mod


@pytest.fixture(scope="module")
def mb():
    shape = graphgen.SHAPES["small"]
    g = graphgen.generate(shape, seed=0)
    return sampler.ladies_sample(4321, g.train_nodes[:256], [2048] * 3, shape.num_nodes, g.indptr, g.indices, [1, 1, 1])


def _upload(layer):
    return (torch.from_numpy(layer.fullrowptr).cuda(), torch.from_numpy(layer.rowptr).cuda(),
            torch.from_numpy(layer.colidx).cuda(), torch.from_numpy(layer.normfact).cuda(), layer.nrows, layer.ncols)


def test_reference_create_coo_tensor_equals_oracle_and_ours(ref, mb):
    import custom_sparse_ops as cso
    for layer in mb.layers:
        r = ref.create_coo_tensor(*_upload(layer))
        rows, cols, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact, layer.nrows)
        assert np.array_equal(r._indices().cpu().numpy(), np.stack([rows, cols]))
        assert np.array_equal(r._values().cpu().numpy().view(np.uint32), vals.view(np.uint32))
        ours = cso.create_coo_tensor(*_upload(layer))
        assert torch.equal(ours._indices(), r._indices()) and torch.equal(ours._values(), r._values())


@pytest.mark.parametrize("D", [602, 64, 37])
def test_reference_spmm_naive_equals_seqfma_oracle(ref, mb, D):
    import custom_sparse_ops as cso
    rng = np.random.Generator(np.random.PCG64(D))
    for layer in mb.layers:
        a = ref.create_coo_tensor(*_upload(layer))
        X = rng.standard_normal((layer.ncols, D)).astype(np.float32)
        dX = torch.from_numpy(X).cuda()
        vals = a._values().cpu().numpy()
        y_naive = ref.spmm_naive(a, dX).cpu().numpy()
        y_seq = oracle.spmm_seqfma(layer.rowptr, layer.colidx32, vals, layer.nrows, X)
        assert np.array_equal(y_naive.view(np.uint32), y_seq.view(np.uint32)), "oracle seqfma order != reference spmm_naive"
        ref64 = oracle.spmm_f64acc(layer.rowptr, layer.colidx32, vals, layer.nrows, X)
        y_lb = ref.spmm_load_balance(a, dX).cpu().numpy()
        ours = cso.spmm(cso.create_coo_tensor(*_upload(layer)), dX).cpu().numpy()
        e_naive, e_lb, e_ours = (oracle.rel_err(y, ref64)[0] for y in (y_naive, y_lb, ours))
        assert e_ours <= 1e-5 and e_lb <= 1e-5 and e_naive <= 1e-5, (e_naive, e_lb, e_ours)
        # rows of <= 64 nonzeros: reference v2 is deterministic and equals v1 / the chunk64 oracle bit for bit
        short = np.diff(layer.rowptr) <= 64
        y_c64 = oracle.spmm_chunk64(layer.rowptr, layer.colidx32, vals, layer.nrows, X)
        assert np.array_equal(y_lb[short].view(np.uint32), y_c64[short].view(np.uint32))


def test_reference_backward_equals_oracle(ref, mb):
    import custom_sparse_ops as cso
    rng = np.random.Generator(np.random.PCG64(8))
    layer = mb.layers[1]
    a = ref.create_coo_tensor(*_upload(layer))
    G = rng.standard_normal((layer.nrows, 96)).astype(np.float32)
    dG = torch.from_numpy(G).cuda()
    vals = a._values().cpu().numpy()
    # custom_sparse_ops.py:34 with the deterministic kernel
    dx_ref = ref.spmm_naive(a.transpose(0, 1).coalesce(), dG.contiguous()).cpu().numpy()
    dx_seq = oracle.spmm_t_seqfma(layer.rowptr, layer.colidx32, vals, layer.nrows, layer.ncols, G)
    assert np.array_equal(dx_ref.view(np.uint32), dx_seq.view(np.uint32))
    ours_a = cso.create_coo_tensor(*_upload(layer))
    x = torch.zeros(layer.ncols, 96, device="cuda", requires_grad=True)
    cso.spmm(ours_a, x).backward(dG)
    ref64 = oracle.spmm_t_f64acc(layer.rowptr, layer.colidx32, vals, layer.nrows, layer.ncols, G)
    assert oracle.rel_err(x.grad.cpu().numpy(), ref64)[0] <= 1e-5
    assert oracle.rel_err(dx_ref, ref64)[0] <= 1e-5
