"""Full-size (BASELINE configs[1] shape) checks: one oracle comparison on a Reddit-sized LADIES-like block and
size-independent properties (linearity, adjoint identity <A x, g> = <x, A^T g>, row sums, reproducibility)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cso():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import custom_sparse_ops
    return custom_sparse_ops


@pytest.fixture(scope="module")
def big_block():
    """16 K x 23 K block, power-law row lengths (mean ~235, max ~3.5 K) and column popularity, like the sampled
    Reddit-shaped adjs[0]; built as the sampler would hand it over (fullrowptr, rowptr, int16 colidx, normfact)."""
    rng = np.random.Generator(np.random.PCG64(2024))
    M, K = 16157, 23207
    lens = np.minimum((rng.pareto(1.6, M) * 90 + 8).astype(np.int64), 3500)
    lens[rng.choice(M, 40, replace=False)] = 0
    colp = rng.pareto(1.2, K) + 0.05
    colp /= colp.sum()
    cols = []
    for n in lens:
        c = np.unique(rng.choice(K, int(n * 1.15) + 1, p=colp))[:n] if n else np.empty(0, np.int64)
        cols.append(c)
    lens = np.array([c.size for c in cols])
    rowptr = np.zeros(M + 1, np.int32)
    rowptr[1:] = np.cumsum(lens)
    colidx = np.concatenate(cols).astype(np.int16)
    full = np.zeros(M + 1, np.int32)
    full[1:] = np.cumsum(lens + rng.integers(1, 400, M))
    normfact = (1 / np.clip(8192 * colp, 1e-10, 1)).astype(np.float32)
    return M, K, full, rowptr, colidx, normfact


def _adj(cso, blk):
    M, K, full, rowptr, colidx, normfact = blk
    return cso.create_coo_tensor(torch.from_numpy(full).cuda(), torch.from_numpy(rowptr).cuda(), torch.from_numpy(colidx).cuda(),
                                 torch.from_numpy(normfact).cuda(), M, K)


def test_fullsize_forward_backward_vs_oracle(cso, big_block):
    M, K, full, rowptr, colidx, normfact = big_block
    a = _adj(cso, big_block)
    _, cols, vals = oracle.build_adj(full, rowptr, colidx, normfact, M)
    assert np.array_equal(a._values().cpu().numpy().view(np.uint32), vals.view(np.uint32))
    D = 602
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(K, D, device="cuda", generator=g).requires_grad_(True)
    go = torch.randn(M, D, device="cuda", generator=g)
    y = cso.spmm(a, x)
    y.backward(go)
    c32 = cols.astype(np.int32)
    yref = oracle.spmm_f64acc(rowptr, c32, vals, M, x.detach().cpu().numpy())
    assert oracle.rel_err(y.detach().cpu().numpy(), yref)[0] <= 1e-5
    gref = oracle.spmm_t_f64acc(rowptr, c32, vals, M, K, go.cpu().numpy())
    assert oracle.rel_err(x.grad.cpu().numpy(), gref)[0] <= 1e-5


def test_fullsize_properties(cso, big_block):
    M, K = big_block[0], big_block[1]
    a = _adj(cso, big_block)
    adj = cso.adjacency_of(a)
    g = torch.Generator(device="cuda").manual_seed(2)
    for D in (1024, 100):
        x1 = torch.randn(K, D, device="cuda", generator=g)
        x2 = torch.randn(K, D, device="cuda", generator=g)
        go = torch.randn(M, D, device="cuda", generator=g)
        y1, y2, y12 = adj.matmul(x1), adj.matmul(x2), adj.matmul(x1 + 2.0 * x2)
        assert torch.equal(adj.matmul(x1), y1), "forward not bit-reproducible"
        scale = y12.abs().max().item()
        assert (y12 - (y1 + 2.0 * y2)).abs().max().item() <= 2e-5 * scale            # linearity
        dx = adj.matmul_t(go, mode="index")
        assert torch.equal(adj.matmul_t(go, mode="index"), dx), "backward through the A^T index not bit-reproducible"
        dxs = adj.matmul_t(go, mode="scatter")                                       # transpose-free: same value to rounding
        assert (dxs - dx).abs().max().item() <= 2e-5 * dx.abs().max().item()
        lhs = (y1.double() * go.double()).sum().item()
        rhs = (x1.double() * dx.double()).sum().item()
        assert abs(lhs - rhs) <= 1e-6 * (y1.double().norm() * go.double().norm()).item()          # adjoint identity
        ones = torch.ones(K, D, device="cuda")
        rowsum = torch.zeros(M, device="cuda", dtype=torch.float64).index_add_(0, a._indices()[0], a._values().double())
        assert (adj.matmul(ones)[:, 0].double() - rowsum).abs().max().item() <= 1e-5 * rowsum.abs().max().item()
    # empty rows stay exactly zero at full size
    lens = np.diff(big_block[3])
    y = adj.matmul(torch.randn(K, 64, device="cuda", generator=g))
    assert torch.all(y[torch.from_numpy(lens == 0).cuda()] == 0)
