"""CPU suite: the oracle against golden vectors.

(1) tests/golden/ref_gpu_*.npz - outputs of the reference's OWN CUDA extension (create_coo_tensor, spmm_naive,
    spmm_load_balance, and the transpose().coalesce() backward expression) run unmodified on a B200
    (tests/golden/make_golden_gpu.py).  oracle_build_adj and the sequential-FMA SpMM must match bit for bit.
(2) tests/golden/{cora_gcn,tiny_sage3}.npz - arrays captured from the unmodified reference Python on CPU plus the
    torch.sparse products (the reference's commented CPU alternative)."""
import os

import numpy as np
import pytest

import oracle

GPU_CASES = ["tiny", "cora", "small"]


@pytest.mark.parametrize("name", GPU_CASES)
def test_oracle_vs_reference_cuda_extension(golden_dir, name):
    z = np.load(os.path.join(golden_dir, f"ref_gpu_{name}.npz"))
    for li in range(int(z["nlayers"])):
        p = f"l{li}_"
        M, K = (int(v) for v in z[p + "shape"])
        rows, cols, vals = oracle.build_adj(z[p + "fullrowptr"], z[p + "rowptr"], z[p + "colidx"], z[p + "normfact"], M)
        assert np.array_equal(np.stack([rows, cols]), z[p + "indices"]), "indices differ from reference create_coo_tensor"
        assert np.array_equal(vals.view(np.uint32), z[p + "values"].view(np.uint32)), "values differ from reference create_coo_tensor"
        c32 = cols.astype(np.int32)
        rowptr = z[p + "rowptr"]
        X, G = z[p + "X"], z[p + "G"]
        # forward: reference spmm_naive == sequential FMA chain, bit for bit
        y_seq = oracle.spmm_seqfma(rowptr, c32, vals, M, X)
        assert np.array_equal(y_seq.view(np.uint32), z[p + "y_naive"].view(np.uint32))
        y64 = oracle.spmm_f64acc(rowptr, c32, vals, M, X)
        assert oracle.rel_err(z[p + "y_naive"], y64)[0] <= 1e-5 and oracle.rel_err(z[p + "y_load_balance"], y64)[0] <= 1e-5
        short = np.diff(rowptr) <= 64        # reference v2 is deterministic only for rows of one 64-nnz chunk
        y_c64 = oracle.spmm_chunk64(rowptr, c32, vals, M, X)
        assert np.array_equal(y_c64[short].view(np.uint32), z[p + "y_load_balance"][short].view(np.uint32))
        # backward expression of custom_sparse_ops.py:34 with the deterministic kernel
        dx_seq = oracle.spmm_t_seqfma(rowptr, c32, vals, M, K, G)
        assert np.array_equal(dx_seq.view(np.uint32), z[p + "dx_naive"].view(np.uint32))
        dx64 = oracle.spmm_t_f64acc(rowptr, c32, vals, M, K, G)
        assert oracle.rel_err(z[p + "dx_load_balance"], dx64)[0] <= 1e-5
        # transpose().coalesce() ordering == oracle CSR transpose
        t_rowptr, t_col, perm = oracle.csr_transpose(rowptr, c32, M, K)
        ti = z[p + "t_indices"]
        assert np.array_equal(ti[1], t_col) and np.array_equal(oracle.coo_rows_to_rowptr(ti[0], K), t_rowptr)
        assert np.array_equal(rows[perm], t_col) and np.array_equal(cols[perm], ti[0])


@pytest.mark.parametrize("name", ["cora_gcn", "tiny_sage3", "tiny_order0"])
def test_oracle_vs_reference_python_goldens(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    for si in range(len(z["seeds"])):
        for li in range(int(z[f"s{si}_nlayers"])):
            p = f"s{si}_l{li}_"
            if p + "none" in z.files:
                continue
            M = int(z[p + "nrows"])
            rows, cols, vals = oracle.build_adj(z[p + "fullrowptr"], z[p + "rowptr"], z[p + "colidx"], z[p + "normfact"], M)
            assert np.array_equal(vals.view(np.uint32), z[p + "values"].view(np.uint32))
            c32 = cols.astype(np.int32)
            y = oracle.spmm_f64acc(z[p + "rowptr"], c32, vals, M, z[p + "x"])
            dx = oracle.spmm_t_f64acc(z[p + "rowptr"], c32, vals, M, int(z[p + "ncols"]), z[p + "g"])
            assert oracle.rel_err(z[p + "y_torchsparse"], y)[0] <= 2e-5
            assert oracle.rel_err(z[p + "dx_torchsparse"], dx)[0] <= 2e-5


def test_oracle_self_consistency():
    rng = np.random.Generator(np.random.PCG64(3))
    M, K, D = 50, 70, 9
    lens = rng.integers(0, 30, M)
    rowptr = np.zeros(M + 1, np.int32)
    rowptr[1:] = np.cumsum(lens)
    cols = np.concatenate([np.sort(rng.choice(K, n, replace=False)) for n in lens]).astype(np.int32)
    vals = rng.standard_normal(cols.size).astype(np.float32)
    X = rng.standard_normal((K, D)).astype(np.float32)
    dense = np.zeros((M, K))
    dense[np.repeat(np.arange(M), lens), cols] = vals
    assert np.allclose(oracle.spmm_f64acc(rowptr, cols, vals, M, X), dense @ X, rtol=1e-6, atol=1e-6)
    G = rng.standard_normal((M, D)).astype(np.float32)
    assert np.allclose(oracle.spmm_t_f64acc(rowptr, cols, vals, M, K, G), dense.T @ G, rtol=1e-6, atol=1e-6)
    assert np.all(oracle.spmm_seqfma(rowptr, cols, vals, M, X)[lens == 0] == 0)
