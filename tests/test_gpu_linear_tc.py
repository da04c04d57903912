"""Dense linears of a layer on the tcgen05 tensor cores (csrc/linear_tc.cuh, include/gnn_b200.h gnn_linear_*):
fp32 in / fp32 out, 3xTF32 inside.  Checked through the C ABI and through the autograd functions of gnn_b200/models.py
against an fp64 product of the same operands (reference models.py:18-19, :60 are `nn.Linear` on fp32 tensors).

Tolerance: relative L2 error <= 3e-6 per call and <= 1e-5 for the worst row (north_star: fp32 results within 1e-5); the
error that remains is the tensor core's truncating accumulation, ~2^-25 per 8-wide step on the main term."""
import ctypes
import os

import numpy as np
import pytest
import torch

from gnn_b200 import _native, graphgen, sampler

pytestmark = pytest.mark.gpu

TOL = 3e-6
ROW_TOL = 1e-5


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return _native.cabi()


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _rel(a, b):
    return ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()


def _row_rel(a, b):
    d = (a.double() - b).norm(dim=1)
    return (d / b.norm(dim=1).clamp_min(1e-30)).max().item()


def _split(lib, W, transposed):
    N, K = W.shape
    w_nk = torch.full((lib.gnn_linear_split_elems(N, K),), float("nan"), device="cuda")
    w_kn = torch.full((lib.gnn_linear_split_elems(K, N),), float("nan"), device="cuda") if transposed else None
    _native.check(lib.gnn_linear_split_weights_f32(_p(W), W.stride(0), N, K, _p(w_nk), _p(w_kn), _stream()), "split")
    return w_nk, w_kn


def _linear(lib, A, rows, M, K, w_split, N, bias, ldc=None, col_off=0):
    ldc = ldc or N
    buf = torch.full((M, ldc), float("nan"), device="cuda")
    out = buf[:, col_off:col_off + N]
    rc = lib.gnn_linear_tf32x3_f32(_p(A), A.stride(0), _p(rows), M, K, _p(w_split), N, _p(bias), _p(out), ldc, _stream())
    _native.check(rc, "gnn_linear_tf32x3_f32")
    return buf, out


@pytest.mark.parametrize("M,K,N,gather,pad,bias", [
    (1, 8, 8, False, 0, True), (128, 32, 256, False, 0, False), (129, 33, 257, True, 0, True), (300, 602, 512, True, 6, True),
    (1000, 602, 512, False, 0, True), (777, 512, 602, False, 0, False), (513, 1024, 512, True, 0, True), (64, 100, 47, False, 3, True),
    (2048, 1433, 16, False, 0, True), (4100, 128, 41, True, 0, True)])
def test_forward_matches_fp64(lib, M, K, N, gather, pad, bias):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + K * 3 + N)
    n_in = M + 37 if gather else M
    xfull = torch.randn(n_in, K + pad, generator=g).cuda()
    x = xfull[:, :K]
    W = (torch.randn(N, K, generator=g) * 0.1).cuda()
    b = torch.randn(N, generator=g).cuda() if bias else None
    rows = torch.randperm(n_in, generator=g)[:M].cuda() if gather else None
    w_nk, _ = _split(lib, W, False)
    buf, out = _linear(lib, x, rows, M, K, w_nk, N, b, ldc=N + 40, col_off=8)
    torch.cuda.synchronize()
    xa = x if rows is None else x[rows]
    ref = xa.double() @ W.double().t() + (b.double() if bias else 0)
    assert not torch.isnan(out).any()
    assert torch.isnan(buf[:, :8]).all() and torch.isnan(buf[:, 8 + N:]).all(), "columns outside the slice were written"
    assert _rel(out, ref) <= TOL
    assert _row_rel(out, ref) <= ROW_TOL


@pytest.mark.parametrize("K", [1433, 2500])
def test_long_reductions_are_chunked(lib, K):
    """K > 1024 runs as several launches whose results are added in fp32 (accumulation-bias bound, see the header)."""
    M, N = 700, 96
    x = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") * 0.1
    b = torch.randn(N, device="cuda")
    w_nk, _ = _split(lib, W, False)
    n0 = lib.gnn_launch_count()
    _, out = _linear(lib, x, None, M, K, w_nk, N, b)
    assert lib.gnn_launch_count() - n0 == (K + 1023) // 1024
    ref = x.double() @ W.double().t() + b.double()
    assert _rel(out, ref) <= TOL and _row_rel(out, ref) <= ROW_TOL


def test_accumulate_and_row_scatter(lib):
    """gnn_linear_tf32x3_f32_ex: C += A.W^T, and result rows added onto C[c_rows[m]] (the backward of x[sampled_nodes]),
    duplicates summed."""
    M, K, N, n_rows = 600, 512, 602, 900
    g = torch.Generator(device="cpu").manual_seed(5)
    A = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) * 0.1).cuda()
    w_nk, _ = _split(lib, W, False)
    base = torch.randn(n_rows, N, generator=g).cuda()
    prod = A.double() @ W.double().t()
    # plain accumulate
    C = base[:M].clone()
    _native.check(lib.gnn_linear_tf32x3_f32_ex(_p(A), K, None, M, K, _p(w_nk), N, None, _p(C), N, None, 1, _stream()), "ex")
    assert _rel(C, base[:M].double() + prod) <= TOL
    # scattered rows with duplicates
    rows = torch.randint(0, n_rows, (M,), generator=g).cuda()
    rows[:50] = rows[50:100]
    C = base.clone()
    _native.check(lib.gnn_linear_tf32x3_f32_ex(_p(A), K, None, M, K, _p(w_nk), N, None, _p(C), N, _p(rows), 1, _stream()), "ex")
    ref = base.double().index_add(0, rows, prod)
    assert _rel(C, ref) <= TOL
    untouched = torch.ones(n_rows, dtype=torch.bool, device="cuda")
    untouched[rows] = False
    assert torch.equal(C[untouched], base[untouched])
    # scatter without the accumulate flag is refused
    assert lib.gnn_linear_tf32x3_f32_ex(_p(A), K, None, M, K, _p(w_nk), N, None, _p(C), N, _p(rows), 0, _stream()) == -1


def test_split_planes_are_tf32_and_sum_back(lib):
    W = (torch.randn(70, 45, device="cuda") * 3)
    w_nk, w_kn = _split(lib, W, True)
    Kp, Np = 64, 96
    nk = w_nk.view(2, 70, Kp)
    kn = w_kn.view(2, 45, Np)
    assert (nk.view(torch.int32) & 0x1FFF).eq(0).all(), "planes must be exact TF32 values (13 low mantissa bits clear)"
    assert torch.equal(nk[:, :, 45:], torch.zeros_like(nk[:, :, 45:])) and torch.equal(kn[:, :, 70:], torch.zeros_like(kn[:, :, 70:]))
    back = nk[0, :, :45].double() + nk[1, :, :45].double()
    assert ((back - W.double()).abs() <= W.double().abs() * 2.0 ** -21).all()
    assert torch.equal(kn[:, :, :70], nk[:, :, :45].transpose(1, 2))


def test_two_weight_matrices_split_in_one_launch(lib):
    W0 = torch.randn(70, 45, device="cuda")
    W1 = torch.randn(33, 45, device="cuda") * 5
    a_nk, a_kn = _split(lib, W0, True)
    b_nk, b_kn = _split(lib, W1, True)
    outs = [torch.full_like(t, float("nan")) for t in (a_nk, a_kn, b_nk, b_kn)]
    n0 = lib.gnn_launch_count()
    _native.check(lib.gnn_linear_split_weights2_f32(_p(W0), 45, 70, 45, _p(outs[0]), _p(outs[1]), _p(W1), 45, 33, 45, _p(outs[2]), _p(outs[3]),
                                                    _stream()), "split2")
    assert lib.gnn_launch_count() - n0 == 1
    for got, ref in zip(outs, (a_nk, a_kn, b_nk, b_kn)):
        assert torch.equal(got, ref)


def test_dx_is_the_same_kernel_on_the_transposed_planes(lib):
    M, n_out, k_in = 900, 512, 602
    dY = torch.randn(M, 2 * n_out, device="cuda")[:, n_out:]          # column slice of a wider gradient
    W = torch.randn(n_out, k_in, device="cuda") * 0.1
    _, w_kn = _split(lib, W, True)
    _, dx = _linear(lib, dY, None, M, n_out, w_kn, k_in, None)
    ref = dY.double() @ W.double()
    assert _rel(dx, ref) <= TOL and _row_rel(dx, ref) <= ROW_TOL


@pytest.mark.parametrize("M,N,K,gather,off", [(1, 8, 8, False, 0), (32, 128, 256, False, 0), (333, 100, 47, False, 0),
                                              (5000, 512, 602, True, 512), (8689, 512, 1024, False, 512), (16157, 512, 602, True, 0),
                                              (40000, 41, 1024, False, 0)])
def test_wgrad_matches_fp64_and_is_reproducible(lib, M, N, K, gather, off):
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    n_in = M + 11 if gather else M
    X = torch.randn(n_in, K, generator=g).cuda()
    dY = torch.randn(M, N + off, generator=g).cuda()[:, off:]
    rows = torch.randperm(n_in, generator=g)[:M].cuda() if gather else None
    wsb = lib.gnn_linear_wgrad_workspace_bytes(M, N, K)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")

    db = torch.full((N + 3,), float("nan"), device="cuda")

    def run(with_bias=False):
        dW = torch.full((N, K + 5), float("nan"), device="cuda")
        rc = lib.gnn_linear_wgrad_tf32x3_f32(_p(dY), dY.stride(0), _p(X), X.stride(0), _p(rows), M, N, K, _p(dW), K + 5,
                                             _p(db) if with_bias else None, _p(ws), wsb, _stream())
        _native.check(rc, "gnn_linear_wgrad_tf32x3_f32")
        return dW
    dW = run(with_bias=True)
    # bias gradient = column sums of dY, from the same pass; plain fp32 sums (no TF32 involved)
    assert torch.isnan(db[N:]).all() and not torch.isnan(db[:N]).any()
    dbref = dY.double().sum(0)
    assert ((db[:N].double() - dbref).norm() / dbref.norm().clamp_min(1e-30)).item() <= 1e-6
    assert torch.isnan(dW[:, K:]).all() and not torch.isnan(dW[:, :K]).any()
    xa = X if rows is None else X[rows]
    ref = dY.double().t() @ xa.double()
    assert _rel(dW[:, :K], ref) <= TOL
    assert _row_rel(dW[:, :K], ref) <= ROW_TOL
    assert torch.equal(dW[:, :K], run()[:, :K]), "fixed-order split sum must be bit-reproducible"


def test_bad_arguments(lib):
    x = torch.zeros(4, 8, device="cuda")
    w = torch.zeros(lib.gnn_linear_split_elems(8, 8), device="cuda")
    out = torch.zeros(4, 8, device="cuda")
    assert lib.gnn_linear_tf32x3_f32(None, 8, None, 4, 8, _p(w), 8, None, _p(out), 8, _stream()) == -1
    assert lib.gnn_linear_tf32x3_f32(_p(x), 8, None, 4, 8, _p(w), 8, None, _p(out), 4, _stream()) == -1      # ldc < N
    assert lib.gnn_linear_tf32x3_f32(_p(x), 8, None, 0, 8, _p(w), 8, None, _p(out), 8, _stream()) == 0       # empty: no launch
    assert lib.gnn_linear_wgrad_tf32x3_f32(_p(x), 8, _p(x), 8, None, 4, 8, 8, _p(out), 8, None, None, 0, _stream()) == -2
    dW = torch.full((8, 8), float("nan"), device="cuda")
    db = torch.full((8,), float("nan"), device="cuda")
    assert lib.gnn_linear_wgrad_tf32x3_f32(_p(x), 8, _p(x), 8, None, 0, 8, 8, _p(dW), 8, _p(db), None, 0, _stream()) == 0
    torch.cuda.synchronize()
    assert torch.equal(dW, torch.zeros_like(dW)) and torch.equal(db, torch.zeros_like(db)), "an empty reduction is a zero gradient"


def test_sage_linears_autograd_matches_fp64():
    from gnn_b200 import models
    torch.manual_seed(0)
    M, n_in, K, n = 1000, 1500, 602, 512
    x = torch.randn(n_in, K, device="cuda", requires_grad=True)
    agg = torch.randn(M, K, device="cuda", requires_grad=True)
    rows = torch.randperm(n_in, device="cuda")[:M]
    WB = (torch.randn(n, K, device="cuda") * 0.05).requires_grad_(True)
    WW = (torch.randn(n, K, device="cuda") * 0.05).requires_grad_(True)
    bB = torch.randn(n, device="cuda", requires_grad=True)
    bW = torch.randn(n, device="cuda", requires_grad=True)
    ins = [x, agg, WB, bB, WW, bW]
    pre = models.SageLinears.apply(x, rows, agg, WB, bB, WW, bW)
    gout = torch.randn_like(pre)
    got = torch.autograd.grad(pre, ins, gout)
    x6, a6, WB6, bB6, WW6, bW6 = ins64 = [t.detach().double().requires_grad_(True) for t in ins]
    pre64 = torch.cat([x6[rows] @ WB6.t() + bB6, a6 @ WW6.t() + bW6], 1)
    ref = torch.autograd.grad(pre64, ins64, gout.double())
    assert _rel(pre, pre64.detach()) <= TOL
    for a, b in zip(got, ref):
        assert _rel(a, b) <= TOL
    # a layer whose input needs no gradient (layer 0 of the model): only parameter gradients come back
    x0 = x.detach()
    pre0 = models.SageLinears.apply(x0, rows, agg.detach(), WB, bB, WW, bW)
    g0 = torch.autograd.grad(pre0, [WB, bB, WW, bW], gout)
    for a, b in zip(g0, [ref[2], ref[3], ref[4], ref[5]]):
        assert _rel(a, b) <= TOL


def test_sage_layer_autograd_matches_fp64():
    """The whole layer (SpMM + both linears) as one autograd node: dX = A^T.dagg with linearB's dX added onto its rows."""
    from gnn_b200 import models
    torch.manual_seed(2)
    M, n_in, K, n = 300, 700, 100, 64
    dense = (torch.rand(M, n_in, device="cuda") < 0.05).float() * torch.rand(M, n_in, device="cuda")
    adj = dense.to_sparse().coalesce()
    x = torch.randn(n_in, K, device="cuda", requires_grad=True)
    rows = torch.randperm(n_in, device="cuda")[:M]
    WB = (torch.randn(n, K, device="cuda") * 0.1).requires_grad_(True)
    WW = (torch.randn(n, K, device="cuda") * 0.1).requires_grad_(True)
    bB = torch.randn(n, device="cuda", requires_grad=True)
    bW = torch.randn(n, device="cuda", requires_grad=True)
    ins = [x, WB, bB, WW, bW]
    pre = models.SageLayer.apply(x, adj, rows, WB, bB, WW, bW)
    gout = torch.randn_like(pre)
    got = torch.autograd.grad(pre, ins, gout)
    x6, WB6, bB6, WW6, bW6 = ins64 = [t.detach().double().requires_grad_(True) for t in ins]
    pre64 = torch.cat([x6[rows] @ WB6.t() + bB6, (dense.double() @ x6) @ WW6.t() + bW6], 1)
    ref = torch.autograd.grad(pre64, ins64, gout.double())
    assert _rel(pre, pre64.detach()) <= TOL
    for name, a, b in zip(["dx", "dWB", "dbB", "dWW", "dbW"], got, ref):
        assert _rel(a, b) <= TOL, name


def test_tc_linear_autograd_matches_fp64():
    from gnn_b200 import models
    torch.manual_seed(1)
    x = torch.randn(700, 100, device="cuda", requires_grad=True)
    W = (torch.randn(256, 100, device="cuda") * 0.1).requires_grad_(True)
    b = torch.randn(256, device="cuda", requires_grad=True)
    y = models.tc_linear(x, W, b)
    gout = torch.randn_like(y)
    got = torch.autograd.grad(y, [x, W, b], gout)
    x6, W6, b6 = ins64 = [t.detach().double().requires_grad_(True) for t in (x, W, b)]
    ref = torch.autograd.grad(x6 @ W6.t() + b6, ins64, gout.double())
    assert _rel(y, (x6 @ W6.t() + b6).detach()) <= TOL
    for a, r in zip(got, ref):
        assert _rel(a, r) <= TOL


@pytest.mark.parametrize("kind,golden", [("graphsage", "model_sage_tiny.npz")])
def test_dropin_model_with_tensor_core_linears_matches_reference_golden(kind, golden):
    """The same check as tests/test_gpu_models.py, with every layer's linears on the tensor cores."""
    import custom_sparse_ops as cso
    from gnn_b200 import models
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", golden))
    shape = graphgen.SHAPES["tiny"]
    g = graphgen.generate(shape, seed=0)
    feats = graphgen.features(shape, seed=1)
    mb = sampler.ladies_sample(5, g.train_nodes[:24], [64] * 3, shape.num_nodes, g.indptr, g.indices, [1, 1, 1])
    adjs = [cso.create_coo_tensor(torch.from_numpy(l.fullrowptr).cuda(), torch.from_numpy(l.rowptr).cuda(), torch.from_numpy(l.colidx).cuda(),
                                  torch.from_numpy(l.normfact).cuda(), l.nrows, l.ncols) for l in mb.layers]
    net = models.build_model(kind, shape.feat_dim, 16, [1, 1, 1], shape.num_classes, dropout=0.0, tc=True)
    net.load_state_dict({k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w_")})
    net.cuda().train()
    sn = [torch.from_numpy(np.ascontiguousarray(s, dtype=np.int64)).cuda() for s in mb.sampled_nodes]
    out = net(torch.from_numpy(feats[mb.input_nodes]).cuda(), adjs, sn)
    assert np.allclose(out.detach().cpu().numpy(), z["out"], rtol=2e-5, atol=2e-6)
    labels = torch.nn.functional.one_hot(torch.from_numpy(graphgen.labels(shape, 3)[mb.batch_nodes]), shape.num_classes).float().cuda()
    w = torch.full((out.shape[0], 1), 1.0 / out.shape[0], device="cuda")
    loss = torch.nn.functional.binary_cross_entropy_with_logits(out, labels, weight=w, reduction="sum")
    assert abs(loss.item() - float(z["loss"])) <= 2e-5 * abs(float(z["loss"]))
    loss.backward()
    for name, p in net.named_parameters():
        assert np.allclose(p.grad.cpu().numpy(), z["g_" + name], rtol=2e-4, atol=2e-6), name
