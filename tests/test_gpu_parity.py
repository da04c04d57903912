"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star): indices / remaps / gathers bit-exact; fp32 SpMM
outputs and gradients within 1e-5 relative of the fp64-accumulated product
(worst row, L2).  Differences below that are summation order: the kernel adds a
row's terms lane-strided and chunk-wise in a fixed order, the oracle sequentially.
"""
import hashlib
import os

import numpy as np
import pytest
import torch

import oracle
from gnn_b200 import graphgen, sampler

pytestmark = pytest.mark.gpu

TOL = 1e-5
WIDTHS = [1, 3, 16, 32, 64, 100, 128, 256, 512, 602, 1024, 1433]


@pytest.fixture(scope="module")
def cu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import cabi_util
    return cabi_util


def _golden_layers(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    out = []
    for si in range(len(z["seeds"])):
        for li in range(int(z[f"s{si}_nlayers"])):
            lp = f"s{si}_l{li}_"
            if lp + "none" in z.files:
                continue
            out.append({k: z[lp + k] for k in ["fullrowptr", "rowptr", "colidx", "normfact", "values", "x", "g",
                                               "y_torchsparse", "dx_torchsparse"]}
                       | {"nrows": int(z[lp + "nrows"]), "ncols": int(z[lp + "ncols"])})
    return out


@pytest.fixture(scope="module")
def small_mb():
    shape = graphgen.SHAPES["small"]
    g = graphgen.generate(shape, seed=0)
    mb = sampler.ladies_sample(4321, g.train_nodes[:256], [2048] * 3, shape.num_nodes, g.indptr, g.indices, [1, 1, 1])
    return shape, g, mb


def _layer_csr(layer):
    rows, cols, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx32, layer.normfact, layer.nrows)
    return layer.rowptr, layer.colidx32, vals


# --------------------------------------------------------------------------- build_adj
@pytest.mark.parametrize("name", ["cora_gcn", "tiny_sage3", "tiny_order0"])
def test_build_adj_bit_exact(cu, golden_dir, name):
    for L in _golden_layers(golden_dir, name):
        idx, vals, col32, row32 = cu.build_adj(cu.dev(L["fullrowptr"]), cu.dev(L["rowptr"]), cu.dev(L["colidx"]), cu.dev(L["normfact"]),
                                               L["nrows"], L["ncols"], want_rows=True)
        rows, cols, ovals = oracle.build_adj(L["fullrowptr"], L["rowptr"], L["colidx"], L["normfact"], L["nrows"])
        assert np.array_equal(idx[0].cpu().numpy(), rows)
        assert np.array_equal(row32.cpu().numpy(), rows.astype(np.int32))
        assert np.array_equal(idx[1].cpu().numpy(), cols)
        assert np.array_equal(col32.cpu().numpy(), cols.astype(np.int32))
        got = vals.cpu().numpy()
        assert np.array_equal(got.view(np.uint32), ovals.view(np.uint32)), "values not bit-exact vs oracle"
        assert np.array_equal(got.view(np.uint32), L["values"].view(np.uint32)), "values not bit-exact vs golden"


def test_build_adj_int32_and_empty(cu, small_mb):
    _, _, mb = small_mb
    layer = mb.layers[0]
    idx, vals, col32 = cu.build_adj(cu.dev(layer.fullrowptr), cu.dev(layer.rowptr), cu.dev(layer.colidx32), cu.dev(layer.normfact),
                                    layer.nrows, layer.ncols)
    rows, cols, ovals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx32, layer.normfact, layer.nrows)
    assert np.array_equal(idx.cpu().numpy(), np.stack([rows, cols]))
    assert np.array_equal(vals.cpu().numpy().view(np.uint32), ovals.view(np.uint32))
    # int16 and int32 column ids give the same adjacency
    idx16, vals16, _ = cu.build_adj(cu.dev(layer.fullrowptr), cu.dev(layer.rowptr), cu.dev(layer.colidx), cu.dev(layer.normfact),
                                    layer.nrows, layer.ncols)
    assert torch.equal(idx16, idx) and torch.equal(vals16, vals)
    # nnz == 0
    z = torch.zeros(5, dtype=torch.int32, device="cuda")
    i0, v0, c0 = cu.build_adj(z, z, torch.zeros(0, dtype=torch.int16, device="cuda"), torch.ones(3, device="cuda"), 4, 3)
    assert i0.shape == (2, 0) and v0.numel() == 0


# --------------------------------------------------------------------------- forward SpMM
@pytest.mark.parametrize("name", ["cora_gcn", "tiny_sage3"])
def test_spmm_golden(cu, golden_dir, name):
    for L in _golden_layers(golden_dir, name):
        rowptr, col32, vals = L["rowptr"], L["colidx"].astype(np.int32), L["values"]
        Y = cu.csr_spmm(cu.dev(rowptr), cu.dev(col32), cu.dev(vals), L["nrows"], L["ncols"], cu.dev(L["x"])).cpu().numpy()
        ref = oracle.spmm_f64acc(rowptr, col32, vals, L["nrows"], L["x"])
        err, _ = oracle.rel_err(Y, ref)
        assert err <= TOL, err
        # the reference's own CPU alternative (torch.sparse, custom_sparse_ops.py:25) is a tolerance-level oracle
        err_ts, _ = oracle.rel_err(Y, L["y_torchsparse"])
        assert err_ts <= 1e-4, err_ts


@pytest.mark.parametrize("D", WIDTHS)
def test_spmm_width_sweep(cu, small_mb, D):
    _, _, mb = small_mb
    rng = np.random.Generator(np.random.PCG64(D))
    for li, layer in enumerate(mb.layers):
        rowptr, col32, vals = _layer_csr(layer)
        X = rng.standard_normal((layer.ncols, D)).astype(np.float32)
        d_rowptr, d_col, d_vals, d_X = cu.dev(rowptr), cu.dev(col32), cu.dev(vals), cu.dev(X)
        Y = cu.csr_spmm(d_rowptr, d_col, d_vals, layer.nrows, layer.ncols, d_X)
        ref = oracle.spmm_f64acc(rowptr, col32, vals, layer.nrows, X)
        err, maxerr = oracle.rel_err(Y.cpu().numpy(), ref)
        # rows of 1-3 elements are pure cancellation (sum of signed terms near zero): judge them on the
        # normwise error max|y-ref|/max|ref| instead of the per-row relative L2
        assert (err if D >= 16 else maxerr) <= TOL, (li, D, err, maxerr)
        Y2 = cu.csr_spmm(d_rowptr, d_col, d_vals, layer.nrows, layer.ncols, d_X)
        assert torch.equal(Y, Y2), "not bit-reproducible"


def _random_csr(rng, M, K, row_lens):
    rowptr = np.zeros(M + 1, np.int32)
    rowptr[1:] = np.cumsum(row_lens)
    cols = np.concatenate([np.sort(rng.choice(K, n, replace=False)) for n in row_lens] + [np.empty(0, np.int64)]).astype(np.int32)
    vals = rng.standard_normal(cols.size).astype(np.float32)
    return rowptr, cols, vals


@pytest.mark.parametrize("D,pad", [(8, 0), (64, 0), (130, 2), (602, 2), (257, 7), (1024, 0)])
def test_spmm_ragged_empty_long_rows(cu, D, pad):
    rng = np.random.Generator(np.random.PCG64(7 + D))
    M, K = 300, 9000
    lens = rng.integers(0, 40, M)
    lens[[0, 1, 2]] = 0              # leading empty rows
    lens[[50, 51]] = 0               # empty rows in the middle
    lens[-3:] = 0                    # trailing empty rows
    lens[7] = 8000                   # one row spanning many chunks
    lens[120] = 3000
    lens[121] = 2049
    rowptr, cols, vals = _random_csr(rng, M, K, lens)
    ldx = D + pad
    Xfull = rng.standard_normal((K, ldx)).astype(np.float32)
    X = np.ascontiguousarray(Xfull[:, :D])
    dX = cu.dev(Xfull)[:, :D]
    Y = cu.csr_spmm(cu.dev(rowptr), cu.dev(cols), cu.dev(vals), M, K, dX, ldx=ldx, ldy=D + pad)
    ref = oracle.spmm_f64acc(rowptr, cols, vals, M, X)
    Yh = Y.cpu().numpy()
    assert not np.isnan(Yh).any(), "some output row was never written"
    assert np.all(Yh[lens == 0] == 0.0), "empty rows must be zero (reference cuda_spmm.cu:626)"
    err, maxerr = oracle.rel_err(Yh, ref)
    assert err <= TOL, (err, maxerr)


def test_spmm_all_empty_and_degenerate(cu):
    rowptr = torch.zeros(11, dtype=torch.int32, device="cuda")
    e_i = torch.zeros(0, dtype=torch.int32, device="cuda")
    e_f = torch.zeros(0, dtype=torch.float32, device="cuda")
    X = torch.ones(5, 12, device="cuda")
    Y = cu.csr_spmm(rowptr, e_i, e_f, 10, 5, X)
    assert Y.shape == (10, 12) and torch.all(Y == 0)
    # single entry, single column
    Y = cu.csr_spmm(cu.dev(np.array([0, 1], np.int32)), cu.dev(np.array([2], np.int32)), cu.dev(np.array([3.0], np.float32)), 1, 5,
                    torch.arange(5, dtype=torch.float32, device="cuda").reshape(5, 1).contiguous())
    assert Y.item() == 6.0


def test_spmm_nan_inf_not_leaked_from_other_rows(cu):
    """A 0-valued lane must never multiply a foreign row: X row 0 is inf/nan but unused."""
    rng = np.random.Generator(np.random.PCG64(3))
    M, K, D = 40, 64, 96
    lens = rng.integers(1, 37, M)
    rowptr = np.zeros(M + 1, np.int32)
    rowptr[1:] = np.cumsum(lens)
    cols = np.concatenate([np.sort(rng.choice(np.arange(1, K), n, replace=False)) for n in lens]).astype(np.int32)
    vals = rng.standard_normal(cols.size).astype(np.float32)
    X = rng.standard_normal((K, D)).astype(np.float32)
    X[0] = np.inf
    X[0, ::2] = np.nan
    Y = cu.csr_spmm(cu.dev(rowptr), cu.dev(cols), cu.dev(vals), M, K, cu.dev(X)).cpu().numpy()
    assert np.isfinite(Y).all()


# --------------------------------------------------------------------------- transpose + backward
def test_transpose_bit_exact_and_backward(cu, small_mb):
    _, _, mb = small_mb
    rng = np.random.Generator(np.random.PCG64(11))
    for layer in mb.layers:
        rowptr, col32, vals = _layer_csr(layer)
        M, K = layer.nrows, layer.ncols
        t_rowptr, t_col, t_vals = cu.csr_transpose(cu.dev(rowptr), cu.dev(col32), cu.dev(vals), M, K)
        o_rowptr, o_col, perm = oracle.csr_transpose(rowptr, col32, M, K)
        assert np.array_equal(t_rowptr.cpu().numpy(), o_rowptr)
        assert np.array_equal(t_col.cpu().numpy(), o_col)
        assert np.array_equal(t_vals.cpu().numpy().view(np.uint32), vals[perm].view(np.uint32))
        for D in (64, 602):
            G = rng.standard_normal((M, D)).astype(np.float32)
            dXg = cu.csr_spmm(t_rowptr, t_col, t_vals, K, M, cu.dev(G)).cpu().numpy()
            ref = oracle.spmm_t_f64acc(rowptr, col32, vals, M, K, G)
            err, _ = oracle.rel_err(dXg, ref)
            assert err <= TOL, err


def test_transpose_edge_cases(cu):
    rng = np.random.Generator(np.random.PCG64(5))
    for M, K in [(1, 1), (33, 70), (64, 31), (100, 3)]:
        lens = rng.integers(0, min(K, 20) + 1, M)
        rowptr, cols, vals = _random_csr(rng, M, K, lens)
        t_rowptr, t_col, t_vals = cu.csr_transpose(cu.dev(rowptr), cu.dev(cols), cu.dev(vals), M, K)
        o_rowptr, o_col, perm = oracle.csr_transpose(rowptr, cols, M, K)
        assert np.array_equal(t_rowptr.cpu().numpy(), o_rowptr)
        assert np.array_equal(t_col.cpu().numpy(), o_col)
        assert np.array_equal(t_vals.cpu().numpy(), vals[perm])


@pytest.mark.parametrize("name", ["cora_gcn", "tiny_sage3"])
def test_backward_golden(cu, golden_dir, name):
    for L in _golden_layers(golden_dir, name):
        rowptr, col32, vals = L["rowptr"], L["colidx"].astype(np.int32), L["values"]
        M, K = L["nrows"], L["ncols"]
        t = cu.csr_transpose(cu.dev(rowptr), cu.dev(col32), cu.dev(vals), M, K)
        dXg = cu.csr_spmm(*t, K, M, cu.dev(L["g"])).cpu().numpy()
        ref = oracle.spmm_t_f64acc(rowptr, col32, vals, M, K, L["g"])
        err, _ = oracle.rel_err(dXg, ref)
        assert err <= TOL, err
        err_ts, _ = oracle.rel_err(dXg, L["dx_torchsparse"])
        assert err_ts <= 1e-4, err_ts


def test_coo_to_csr(cu, small_mb):
    _, _, mb = small_mb
    layer = mb.layers[1]
    rows, cols, _ = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx32, layer.normfact, layer.nrows)
    rowptr, col32 = cu.coo_to_csr(cu.dev(np.stack([rows, cols])), layer.nrows)
    assert np.array_equal(rowptr.cpu().numpy(), layer.rowptr)
    assert np.array_equal(rowptr.cpu().numpy(), oracle.coo_rows_to_rowptr(rows, layer.nrows))
    assert np.array_equal(col32.cpu().numpy(), layer.colidx32)


# --------------------------------------------------------------------------- remap + gather
@pytest.mark.parametrize("name", ["cora_gcn", "tiny_sage3", "tiny_order0"])
def test_placement_remap_and_gather_bit_exact(cu, golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    shape = graphgen.SHAPES[str(z["shape"])]
    g = graphgen.generate(shape, seed=0)
    feats = graphgen.features(shape, seed=1)
    world = int(z["world"])
    F = shape.feat_dim
    ld = (F + 3) // 4 * 4                 # shards are padded to 16-byte rows
    def padded(a):
        out = np.zeros((a.shape[0], ld), np.float32)
        out[:, :F] = a
        return out
    bufs = [cu.dev(padded(feats[z["gpu_buffer_group"][i]])) for i in range(world)]
    host = cu.dev(padded(feats))          # stands in for the mapped host table
    bases = torch.tensor([b.data_ptr() for b in bufs] + [host.data_ptr()], dtype=torch.int64, device="cuda")
    devices = torch.arange(world, dtype=torch.int64, device="cuda")
    orders = [int(o) for o in z["orders"]]
    for si, seed in enumerate(z["seeds"]):
        pre = f"s{si}_"
        rank = int(z[pre + "rank"])
        mb = sampler.ladies_sample(int(seed), z[pre + "batch_nodes"], [int(z["samp_num"])] * 5, shape.num_nodes, g.indptr,
                                   g.indices, orders)
        did, idx = z["device_id_of_nodes_group"][rank], z["idx_of_nodes_on_device_group"][rank]
        src, slot, xrows, counts = cu.placement_remap(cu.dev(mb.input_nodes), cu.dev(did, torch.int64), cu.dev(idx, torch.int64),
                                                      devices, bases, ld)
        o_src, o_slot = oracle.placement_remap(mb.input_nodes, did, idx, list(range(world)))
        assert np.array_equal(src.cpu().numpy(), o_src)
        assert np.array_equal(slot.cpu().numpy(), o_slot)
        for i in range(world):                  # the reference's masks / slot lists (sampler.py:156-158)
            assert np.array_equal(src.cpu().numpy() == i, z[pre + f"mask_dev{i}"])
            assert np.array_equal(slot.cpu().numpy()[o_src == i], z[pre + f"idx_dev{i}"])
        assert np.array_equal(slot.cpu().numpy()[o_src == -1], z[pre + "idx_cpu"])
        c = counts.cpu().numpy()
        assert c[:world].tolist() == [int((o_src == i).sum()) for i in range(world)] and c[world] == int((o_src == -1).sum())
        out = cu.gather_rows(xrows, F).cpu().numpy()
        ref = oracle.gather_rows([feats[z["gpu_buffer_group"][i]] for i in range(world)], feats, o_src, o_slot)
        assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
        sha = np.frombuffer(hashlib.sha256(np.ascontiguousarray(out).tobytes()).digest(), dtype=np.uint8)
        assert np.array_equal(sha, z[pre + "input_feat_sha"]), "gathered rows differ from the reference gather (main.py:129-134)"
        # fused gather + SpMM on the deepest layer
        layer = mb.layers[0]
        rowptr, col32, vals = _layer_csr(layer)
        Y = cu.gather_spmm(cu.dev(rowptr), cu.dev(col32), cu.dev(vals), layer.nrows, layer.ncols, F, xrows).cpu().numpy()
        yref = oracle.gather_spmm_f64acc(rowptr, col32, vals, layer.nrows, [feats[z["gpu_buffer_group"][i]] for i in range(world)],
                                         feats, o_src, o_slot)
        err, _ = oracle.rel_err(Y, yref)
        assert err <= TOL, err


def test_index_rows(cu):
    rng = np.random.Generator(np.random.PCG64(1))
    for F in (5, 64, 602, 1024):
        X = rng.standard_normal((500, F)).astype(np.float32)
        idx = rng.integers(0, 500, 333)
        out = cu.index_rows(cu.dev(X), cu.dev(idx)).cpu().numpy()
        assert np.array_equal(out, X[idx])


# --------------------------------------------------------------------------- error paths and limits
def test_error_codes_and_workspace_contract(cu):
    import ctypes
    from gnn_b200 import _native
    lib = _native.cabi()
    rowptr = cu.dev(np.array([0, 2, 3], np.int32))
    col = cu.dev(np.array([0, 1, 1], np.int32))
    vals = cu.dev(np.array([1, 2, 3], np.float32))
    X = torch.ones(2, 8, device="cuda")
    Y = torch.zeros(2, 8, device="cuda")
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    need = lib.gnn_csr_spmm_workspace_bytes(2, 3, 8)
    ws = torch.empty(need, dtype=torch.uint8, device="cuda")
    assert lib.gnn_csr_spmm_f32(P(rowptr), P(col), P(vals), 2, 2, 3, 8, P(X), 8, P(Y), 8, None, 0, st) == -2          # no workspace
    assert lib.gnn_csr_spmm_f32(P(rowptr), P(col), P(vals), 2, 2, 3, 8, P(X), 8, P(Y), 8, P(ws), need - 1, st) == -2   # too small
    assert lib.gnn_csr_spmm_f32(P(rowptr), P(col), P(vals), 2, 2, 3, 8, P(X), 4, P(Y), 8, P(ws), need, st) == -1       # ldx < D
    assert lib.gnn_csr_spmm_f32(P(rowptr), P(col), P(vals), 2, 2, 3, 8, P(X), 8, P(Y), 8, P(ws), need, st) == 0
    torch.cuda.synchronize()
    assert Y.tolist() == [[3.0] * 8, [3.0] * 8]
    assert lib.gnn_csr_spmm_f32(P(rowptr), P(col), P(vals), 2, 2, 3, 1 << 25, P(X), 1 << 25, P(Y), 1 << 25, P(ws), need, st) == -3   # D limit
    assert lib.gnn_build_adj(P(rowptr), P(rowptr), P(col), 8, P(vals), 2, 2, 3, None, P(vals), None, None, st) == -1       # colidx width


def test_int32_columns_beyond_int16_range(cu):
    """K > 32767 needs int32 column ids (the reference's int16 hand-off silently wraps)."""
    rng = np.random.Generator(np.random.PCG64(9))
    M, K, D = 200, 70000, 40
    lens = rng.integers(1, 60, M)
    rowptr, cols, _ = _random_csr(rng, M, K, lens)
    full = np.zeros(M + 1, np.int32)
    full[1:] = np.cumsum(lens + 3)
    nf = rng.random(K).astype(np.float32) + 0.5
    idx, vals, col32 = cu.build_adj(cu.dev(full), cu.dev(rowptr), cu.dev(cols), cu.dev(nf), M, K)
    rows_o, cols_o, vals_o = oracle.build_adj(full, rowptr, cols, nf, M)
    assert np.array_equal(idx.cpu().numpy(), np.stack([rows_o, cols_o])) and cols_o.max() > 32767
    assert np.array_equal(vals.cpu().numpy().view(np.uint32), vals_o.view(np.uint32))
    X = rng.standard_normal((K, D)).astype(np.float32)
    Y = cu.csr_spmm(cu.dev(rowptr), col32, vals, M, K, cu.dev(X)).cpu().numpy()
    assert oracle.rel_err(Y, oracle.spmm_f64acc(rowptr, cols, vals_o, M, X))[0] <= TOL
    t = cu.csr_transpose(cu.dev(rowptr), col32, vals, M, K)
    o_rowptr, o_col, perm = oracle.csr_transpose(rowptr, cols, M, K)
    assert np.array_equal(t[0].cpu().numpy(), o_rowptr) and np.array_equal(t[1].cpu().numpy(), o_col)


def test_concurrent_threads_and_streams(cu, small_mb):
    """The reference is entered concurrently from trainer and sampler threads: four Python threads,
    each on its own stream, run create/spmm/transpose/spmm at once; results must equal the single-threaded ones."""
    import threading
    import custom_sparse_ops as cso
    _, _, mb = small_mb
    layer = mb.layers[0]
    args = [torch.from_numpy(a).cuda() for a in (layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact)]
    X = torch.randn(layer.ncols, 100, device="cuda")
    G = torch.randn(layer.nrows, 100, device="cuda")
    a0 = cso.create_coo_tensor(*args, layer.nrows, layer.ncols)
    y0 = cso.adjacency_of(a0).matmul(X)
    d0 = cso.adjacency_of(a0).matmul_t(G, mode="index")
    torch.cuda.synchronize()
    results, errors = {}, []

    def work(tid):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(5):
                    a = cso.create_coo_tensor(*args, layer.nrows, layer.ncols)
                    adj = cso.adjacency_of(a)
                    ds = adj.matmul_t(G, mode="scatter")          # transpose-free: same value up to summation order
                    y, d = adj.matmul(X), adj.matmul_t(G, mode="index")
                s.synchronize()
            results[tid] = (torch.equal(y, y0), torch.equal(d, d0), torch.equal(a._values(), a0._values()),
                            torch.allclose(ds, d0, rtol=1e-4, atol=1e-5))
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert all(all(v) for v in results.values()) and len(results) == 4


def test_spmm_and_transpose_randomised_shapes(cu):
    """Differential fuzz: random shapes, densities, skew, widths and leading dimensions against the oracle."""
    rng = np.random.Generator(np.random.PCG64(20261018))
    for trial in range(40):
        M = int(rng.integers(1, 3000))
        K = int(rng.integers(1, 6000))
        D = int(rng.choice([1, 2, 5, 16, 24, 33, 64, 96, 100, 128, 200, 256, 300, 512, 602, 640, 1000, 1024]))
        mean = float(rng.choice([0.5, 3, 20, 150]))
        lens = np.minimum(rng.poisson(mean, M) * (rng.random(M) < 0.9), K)
        if rng.random() < 0.3:
            lens[rng.integers(0, M)] = min(K, int(rng.integers(500, 5000)))       # a hub row
        rowptr, cols, vals = _random_csr(rng, M, K, lens)
        pad = int(rng.choice([0, 0, 1, 2, 4, 6, 30]))
        Xfull = rng.standard_normal((K, D + pad)).astype(np.float32)
        X = np.ascontiguousarray(Xfull[:, :D])
        d_rowptr, d_cols, d_vals = cu.dev(rowptr), cu.dev(cols), cu.dev(vals)
        Y = cu.csr_spmm(d_rowptr, d_cols, d_vals, M, K, cu.dev(Xfull)[:, :D], ldx=D + pad, ldy=D + int(rng.choice([0, 3]))).cpu().numpy()
        assert not np.isnan(Y).any(), (trial, M, K, D)
        ref = oracle.spmm_f64acc(rowptr, cols, vals, M, X)
        err, maxerr = oracle.rel_err(Y, ref)
        assert (err if D >= 16 else maxerr) <= TOL, (trial, M, K, D, pad, err, maxerr)
        if cols.size:
            t_rowptr, t_col, t_vals = cu.csr_transpose(d_rowptr, d_cols, d_vals, M, K)
            o_rowptr, o_col, perm = oracle.csr_transpose(rowptr, cols, M, K)
            assert np.array_equal(t_rowptr.cpu().numpy(), o_rowptr) and np.array_equal(t_col.cpu().numpy(), o_col)
            assert np.array_equal(t_vals.cpu().numpy(), vals[perm])


def test_corunner_reserve_changes_the_grid_not_the_result(cu):
    """gnn_set_corunner_ctas only changes how many chunks the planner cuts (include/gnn_b200.h): every setting stays
    within the fp32 bar of the oracle, repeats bit-identically, and the setter returns the previous value."""
    rng = np.random.Generator(np.random.PCG64(77))
    M, K, D = 6000, 9000, 256
    lens = np.minimum(rng.poisson(180, M), K)
    lens[17] = 5000                                                        # a hub row spanning many chunks
    rowptr, cols, vals = _random_csr(rng, M, K, lens)
    X = rng.standard_normal((K, D)).astype(np.float32)
    ref = oracle.spmm_f64acc(rowptr, cols, vals, M, X)
    d = [cu.dev(a) for a in (rowptr, cols, vals, X)]
    from gnn_b200 import _native
    lib = _native.cabi()
    assert lib.gnn_set_corunner_ctas(0) >= 0
    try:
        outs = {}
        for reserve in (0, 16, 64, 10 ** 6):                               # the last one leaves a single CTA slot
            prev = lib.gnn_set_corunner_ctas(reserve)
            assert prev >= 0
            y1 = cu.csr_spmm(d[0], d[1], d[2], M, K, d[3]).cpu().numpy()
            y2 = cu.csr_spmm(d[0], d[1], d[2], M, K, d[3]).cpu().numpy()
            assert np.array_equal(y1.view(np.uint32), y2.view(np.uint32)), reserve
            assert oracle.rel_err(y1, ref)[0] <= TOL, reserve
            outs[reserve] = y1
        assert lib.gnn_set_corunner_ctas(0) == 10 ** 6
    finally:
        lib.gnn_set_corunner_ctas(0)


def test_build_adj_hub_rows_and_chunk_boundaries(cu):
    """build_adj walks the entries in 256-entry chunks: rows far longer than a chunk, empty rows between them and a
    last partial chunk must still give the reference's (row, col, value) triples bit for bit."""
    rng = np.random.Generator(np.random.PCG64(5))
    M, K = 700, 30000
    lens = rng.integers(0, 4, M)
    lens[[3, 4, 350, 699]] = [4000, 257, 9000, 1]
    lens[5:40] = 0
    rowptr, cols, _ = _random_csr(rng, M, K, lens)
    full_lens = lens + rng.integers(1, 50, M)                                # full-graph degree >= sampled degree
    fullrowptr = np.zeros(M + 1, dtype=np.int32)
    np.cumsum(full_lens, out=fullrowptr[1:])
    normfact = (1.0 / rng.uniform(1e-3, 1.0, K)).astype(np.float32)
    for colt in (np.int16, np.int32):
        c = cols.astype(colt)
        idx, v, c32 = cu.build_adj(cu.dev(fullrowptr), cu.dev(rowptr), cu.dev(c), cu.dev(normfact), M, K)
        rows_o, cols_o, vals_o = oracle.build_adj(fullrowptr, rowptr, c, normfact, M)
        assert np.array_equal(idx.cpu().numpy(), np.stack([rows_o, cols_o]))
        assert np.array_equal(v.cpu().numpy().view(np.uint32), vals_o.view(np.uint32))
        assert np.array_equal(c32.cpu().numpy(), cols_o.astype(np.int32))
