"""Property tests of the CPU oracle itself (hypothesis): the checker has to be right on ragged, empty and degenerate
inputs before the GPU parity tests can lean on it.  Pure CPU, small sizes."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle


@st.composite
def csr(draw, max_m=24, max_k=40):
    M = draw(st.integers(0, max_m))
    K = draw(st.integers(1, max_k))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(0, K + 1, M) * (rng.random(M) < draw(st.sampled_from([0.0, 0.3, 0.9, 1.0])))
    lens = lens.astype(np.int64)
    rowptr = np.zeros(M + 1, np.int32)
    rowptr[1:] = np.cumsum(lens)
    cols = np.concatenate([np.sort(rng.choice(K, int(n), replace=False)) for n in lens] + [np.empty(0, np.int64)]).astype(np.int32)
    vals = rng.standard_normal(cols.size).astype(np.float32)
    return M, K, rowptr, cols, vals, lens, rng


def _dense(M, K, cols, vals, lens):
    a = np.zeros((M, K))
    a[np.repeat(np.arange(M), lens), cols] = vals
    return a


@settings(max_examples=60, deadline=None)
@given(csr(), st.integers(1, 19))
def test_spmm_variants_agree_with_dense_product(c, D):
    M, K, rowptr, cols, vals, lens, rng = c
    X = rng.standard_normal((K, D)).astype(np.float32)
    ref = _dense(M, K, cols, vals, lens) @ X.astype(np.float64)
    for fn in (oracle.spmm_f64acc, oracle.spmm_seqfma, oracle.spmm_chunk64):
        y = fn(rowptr, cols, vals, M, X)
        assert y.shape == (M, D) and y.dtype == np.float32
        assert np.allclose(y, ref, rtol=2e-5, atol=2e-5)
        assert np.all(y[lens == 0] == 0)                      # empty rows are zero rows (cuda_spmm.cu:626)


@settings(max_examples=60, deadline=None)
@given(csr())
def test_transpose_is_sorted_and_an_involution(c):
    M, K, rowptr, cols, vals, lens, rng = c
    t_rowptr, t_col, perm = oracle.csr_transpose(rowptr, cols, M, K)
    assert t_rowptr[0] == 0 and t_rowptr[-1] == cols.size and np.all(np.diff(t_rowptr) >= 0)
    t_lens = np.diff(t_rowptr).astype(np.int64)
    assert np.array_equal(_dense(K, M, t_col, vals[perm], t_lens), _dense(M, K, cols, vals, lens).T)
    for k in range(K):                                            # ascending source row inside every row of A^T
        seg = t_col[t_rowptr[k]:t_rowptr[k + 1]]
        assert np.all(np.diff(seg) > 0)
    if M > 0:
        b_rowptr, b_col, b_perm = oracle.csr_transpose(t_rowptr, t_col, K, M)
        assert np.array_equal(b_rowptr, rowptr) and np.array_equal(b_col, cols)
        assert np.array_equal(perm[b_perm], np.arange(cols.size))


@settings(max_examples=40, deadline=None)
@given(csr(), st.integers(1, 11))
def test_backward_equals_forward_on_the_transpose(c, D):
    M, K, rowptr, cols, vals, lens, rng = c
    G = rng.standard_normal((M, D)).astype(np.float32)
    t_rowptr, t_col, perm = oracle.csr_transpose(rowptr, cols, M, K)
    direct = oracle.spmm_t_f64acc(rowptr, cols, vals, M, K, G)
    via_t = oracle.spmm_f64acc(t_rowptr, t_col, vals[perm], K, G)
    assert np.array_equal(direct.view(np.uint32), via_t.view(np.uint32))      # same terms, same order, fp64 accumulate
    seq = oracle.spmm_t_seqfma(rowptr, cols, vals, M, K, G)
    assert np.array_equal(seq.view(np.uint32), oracle.spmm_seqfma(t_rowptr, t_col, vals[perm], K, G).view(np.uint32))


@settings(max_examples=60, deadline=None)
@given(csr(max_k=300))
def test_build_adj_formula_and_rowptr_round_trip(c):
    M, K, rowptr, cols, vals, lens, rng = c
    full_lens = lens + rng.integers(1, 9, M)
    fullrowptr = np.zeros(M + 1, np.int32)
    fullrowptr[1:] = np.cumsum(full_lens)
    normfact = (1.0 / rng.uniform(1e-6, 1.0, K)).astype(np.float32)
    for colt in (np.int16, np.int32):
        rows_o, cols_o, vals_o = oracle.build_adj(fullrowptr, rowptr, cols.astype(colt), normfact, M)
        rows_np = np.repeat(np.arange(M), lens)
        assert np.array_equal(rows_o, rows_np) and np.array_equal(cols_o, cols)
        want = ((1.0 / full_lens[rows_np].astype(np.float64)) * normfact[cols].astype(np.float64)).astype(np.float32)
        assert np.array_equal(vals_o.view(np.uint32), want.view(np.uint32))                 # cuda_spmm.cu:800
        assert np.array_equal(oracle.coo_rows_to_rowptr(rows_o, M), rowptr)                 # cuda_spmm.cu:255-265


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 4), st.integers(1, 9))
def test_remaps_and_gather_match_numpy_indexing(seed, world, F):
    rng = np.random.Generator(np.random.PCG64(seed))
    N = int(rng.integers(5, 200))
    device_id = rng.integers(-1, world, N).astype(np.int64)
    devices = np.arange(world, dtype=np.int64)
    idx_on_dev = np.arange(N, dtype=np.int64)
    buffers = []
    table = rng.standard_normal((N, F)).astype(np.float32)
    for d in range(world):
        held = np.flatnonzero(device_id == d)
        idx_on_dev[held] = rng.permutation(held.size)
        buf = np.zeros((max(held.size, 1), F), np.float32)
        buf[idx_on_dev[held]] = table[held]
        buffers.append(buf)
    after = np.unique(rng.integers(0, N, int(rng.integers(1, N + 1))))
    prev = rng.choice(after, int(rng.integers(0, after.size + 1)), replace=False)
    assert np.array_equal(oracle.sampled_nodes(after, prev), np.where(np.isin(after, prev))[0])      # sampler.py:143
    src, slot = oracle.placement_remap(after, device_id, idx_on_dev, devices)
    assert np.array_equal(src, device_id[after].astype(np.int32))                                    # sampler.py:152
    on_dev = src >= 0
    assert np.array_equal(slot[on_dev], idx_on_dev[after[on_dev]]) and np.array_equal(slot[~on_dev], after[~on_dev])
    got = oracle.gather_rows(buffers, table, src, slot)                                               # main.py:129-134
    assert np.array_equal(got.view(np.uint32), table[after].view(np.uint32))
