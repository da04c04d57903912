"""GPU parity of the round-2 kernels, through the C ABI, against the CPU oracle:

  * flat (nonzero-split) SpMM: short rows, long rows spanning many chunks, runs of empty rows longer than the
    row-pointer window, leading/trailing empty rows, every load width, padded leading dimensions;
  * self-cleaning arrival counters (GNN_SPMM_COUNTERS_ZEROED): one zeroed region reused by many calls;
  * transpose-free backward (gnn_csr_spmm_t_f32) within the 1e-5 bar of the fp64 oracle;
  * row-blocked CSR transpose (bitmap budget forced small) bit-identical to the single-block result and the oracle.
"""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def cu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import cabi_util
    return cabi_util


def random_csr(rng, M, K, row_lens):
    """CSR with the given row lengths, sorted unique columns per row, N(0,1) values."""
    row_lens = np.minimum(np.asarray(row_lens, dtype=np.int64), K)
    rowptr = np.zeros(M + 1, dtype=np.int32)
    np.cumsum(row_lens, out=rowptr[1:])
    cols = np.empty(int(rowptr[-1]), dtype=np.int32)
    for r in range(M):
        n = int(row_lens[r])
        if n:
            cols[rowptr[r]:rowptr[r + 1]] = np.sort(rng.choice(K, n, replace=False))
    vals = rng.standard_normal(cols.size).astype(np.float32)
    return rowptr, cols, vals


def shapes(rng):
    """(name, M, K, row_lens) cases; mean row length < 96 selects the flat kernel, above it the row-split kernel."""
    out = []
    out.append(("short", 3000, 5000, rng.integers(1, 12, 3000)))
    mixed = rng.integers(0, 6, 2500)
    mixed[[7, 900, 901, 2499]] = [1500, 700, 300, 2000]          # hub rows spanning dozens of 32-entry chunks
    out.append(("short+hubs", 2500, 4000, mixed))
    empties = np.zeros(4000, dtype=np.int64)
    empties[[300, 301, 900, 3500]] = [5, 40, 3, 9]               # runs of > 128 empty rows (window fallback), empty head/tail
    out.append(("mostly-empty", 4000, 300, empties))
    out.append(("one-row", 1, 5000, np.array([3000])))
    out.append(("top-layer", 512, 8000, rng.integers(20, 130, 512)))
    out.append(("dense-rows", 300, 6000, rng.integers(300, 1200, 300)))          # row-split kernel
    return out


@pytest.mark.parametrize("D", [1, 16, 37, 100, 128, 256, 602, 1024])
def test_spmm_flat_and_rowsplit_shapes(cu, D):
    from gnn_b200 import _native
    lib = _native.cabi()
    rng = np.random.Generator(np.random.PCG64(100 + D))
    for name, M, K, lens in shapes(rng):
        rowptr, cols, vals = random_csr(rng, M, K, lens)
        X = rng.standard_normal((K, D)).astype(np.float32)
        ref = oracle.spmm_f64acc(rowptr, cols, vals, M, X)
        d = [cu.dev(a) for a in (rowptr, cols, vals)]
        for ld in sorted({D, (D + 31) // 32 * 32, D + 1}):          # contiguous, 128-byte padded rows, odd leading dimension
            Xd = torch.zeros((K, ld), device="cuda")
            Xd[:, :D] = cu.dev(X)
            if ld > D:
                Xd[:, D:] = float("nan")                            # padding may be loaded but must never reach the result
            ldy = D if ld % 32 == 0 else D + 3                     # vector stores / scalar stores
            Y = cu.csr_spmm(d[0], d[1], d[2], M, K, Xd[:, :D], ldx=ld, ldy=ldy)
            got = Y.cpu().numpy()
            assert np.isfinite(got).all(), (name, D, ld, "a row was not written or padding leaked")
            err, maxerr = oracle.rel_err(got, ref)
            assert (err if D >= 16 else maxerr) <= TOL, (name, D, ld, err, maxerr)
            Y2 = cu.csr_spmm(d[0], d[1], d[2], M, K, Xd[:, :D], ldx=ld, ldy=ldy)
            assert torch.equal(Y, Y2), (name, D, ld, "not bit-reproducible")
        # with per-entry row ids (what create_coo_tensor attaches) the flat kernel skips the row search: same bits
        rowidx = cu.dev(np.repeat(np.arange(M, dtype=np.int32), np.diff(rowptr)))
        cb, pb = lib.gnn_csr_spmm_counter_bytes(M, len(vals), D), lib.gnn_csr_spmm_partial_bytes(M, len(vals), D)
        counters = torch.zeros(cb, dtype=torch.uint8, device="cuda")
        partials = torch.empty(pb, dtype=torch.uint8, device="cuda")
        Xc = cu.dev(X)
        Y3 = cu.csr_spmm_ex(d[0], d[1], d[2], M, K, Xc, counters, partials, zeroed=True, rowidx=rowidx)
        assert torch.equal(Y3, cu.csr_spmm(d[0], d[1], d[2], M, K, Xc)), (name, D, "row-id path differs from the search path")
        assert int(counters.count_nonzero()) == 0


def test_counters_self_cleaning(cu):
    """A zeroed counter region serves call after call without a memset (what spmm_ext.cpp does per stream)."""
    from gnn_b200 import _native
    lib = _native.cabi()
    rng = np.random.Generator(np.random.PCG64(5))
    cases = []
    for name, M, K, lens in shapes(rng):
        rowptr, cols, vals = random_csr(rng, M, K, lens)
        for D in (64, 256, 1024):
            cases.append((name, M, K, D, rowptr, cols, vals, rng.standard_normal((K, D)).astype(np.float32)))
    cb = max(lib.gnn_csr_spmm_counter_bytes(M, len(vals), D) for _, M, K, D, _, _, vals, _ in cases)
    pb = max(lib.gnn_csr_spmm_partial_bytes(M, len(vals), D) for _, M, K, D, _, _, vals, _ in cases)
    counters = torch.zeros(cb, dtype=torch.uint8, device="cuda")
    partials = torch.empty(pb, dtype=torch.uint8, device="cuda")
    for rep in range(2):
        for name, M, K, D, rowptr, cols, vals, X in cases:
            d = [cu.dev(a) for a in (rowptr, cols, vals)]
            Y = cu.csr_spmm_ex(d[0], d[1], d[2], M, K, cu.dev(X), counters, partials, zeroed=True)
            ref = oracle.spmm_f64acc(rowptr, cols, vals, M, X)
            err, _ = oracle.rel_err(Y.cpu().numpy(), ref)
            assert err <= TOL, (name, D, rep, err)
            assert int(counters.count_nonzero()) == 0, (name, D, rep, "counters must be back to zero when the call completes")
            # and the plain entry point (memset inside) gives the same bits
            assert torch.equal(Y, cu.csr_spmm(d[0], d[1], d[2], M, K, cu.dev(X))), (name, D)


@pytest.mark.parametrize("D", [4, 100, 128, 602, 1024])
def test_scatter_backward_matches_oracle(cu, D):
    rng = np.random.Generator(np.random.PCG64(300 + D))
    for name, M, K, lens in shapes(rng):
        rowptr, cols, vals = random_csr(rng, M, K, lens)
        G = rng.standard_normal((M, D)).astype(np.float32)
        ref = oracle.spmm_t_f64acc(rowptr, cols, vals, M, K, G)
        d = [cu.dev(a) for a in (rowptr, cols, vals)]
        rowidx = cu.dev(np.repeat(np.arange(M, dtype=np.int32), np.diff(rowptr)))
        for lddx in (D, D + 4, D + 1):                           # memset path, padded vector path, scalar-reduction path
            dX = cu.csr_spmm_t(d[0], d[1], d[2], M, K, cu.dev(G), lddx=lddx, rowidx=(rowidx if lddx != D + 4 else None))
            got = dX[:, :D].cpu().numpy()
            assert np.isfinite(got).all(), (name, D, lddx)
            err, maxerr = oracle.rel_err(got, ref)
            assert (err if D >= 16 else maxerr) <= TOL, (name, D, lddx, err, maxerr)
            if lddx > D:
                assert torch.isnan(dX[:, D:]).all(), "padding columns must stay untouched"
    # empty operands
    z = torch.zeros(4, dtype=torch.int32, device="cuda")
    out = cu.csr_spmm_t(z, torch.zeros(0, dtype=torch.int32, device="cuda"), torch.zeros(0, device="cuda"), 3, 5,
                        torch.ones((3, 8), device="cuda"))
    assert torch.count_nonzero(out) == 0


def test_transpose_row_blocked_equals_single_block(cu):
    rng = np.random.Generator(np.random.PCG64(9))
    for name, M, K, lens in shapes(rng):
        rowptr, cols, vals = random_csr(rng, M, K, lens)
        d = [cu.dev(a) for a in (rowptr, cols, vals)]
        t0 = cu.csr_transpose(d[0], d[1], d[2], M, K, want_rows=True)
        o_rowptr, o_col, perm = oracle.csr_transpose(rowptr, cols, M, K)
        assert np.array_equal(t0[3].cpu().numpy(), np.repeat(np.arange(K, dtype=np.int32), np.diff(o_rowptr))), name
        assert np.array_equal(t0[0].cpu().numpy(), o_rowptr) and np.array_equal(t0[1].cpu().numpy(), o_col), name
        assert np.array_equal(t0[2].cpu().numpy().view(np.uint32), vals[perm].view(np.uint32)), name
        for budget in (K * 8 * 1, K * 8 * 3, K * 8 * 17):         # 1, 3, 17 bitmap words per column -> many row blocks
            prev = cu.set_transpose_budget(budget)
            try:
                t1 = cu.csr_transpose(d[0], d[1], d[2], M, K, want_rows=True)
            finally:
                cu.set_transpose_budget(prev)
            for a, b in zip(t0, t1):
                assert torch.equal(a, b), (name, budget)
