"""The device sampler's array kernels called one by one through the C ABI (include/gnn_b200.h) against numpy: row slice +
column counts (reference sampler.py:113-117), support compaction, membership tables and the column slice
(sampler.py:133-136).  Shapes exercise what the entry-parallel decompositions must get right: empty rows, rows longer than a
tile (2,048 entries) and a chunk (1,024), repeated nodes, totals that are not multiples of the chunk."""
import ctypes

import numpy as np
import pytest
import torch

from gnn_b200 import _native

from .cabi_util import _ptr, _stream, dev

pytestmark = pytest.mark.gpu


def _graph(rng, n_nodes, long_rows, long_len, max_deg):
    deg = rng.integers(0, max_deg, n_nodes)
    deg[rng.choice(n_nodes, n_nodes // 10, replace=False)] = 0
    deg[rng.choice(n_nodes, long_rows, replace=False)] = long_len
    indptr = np.concatenate(([0], np.cumsum(deg))).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(n_nodes, d, replace=False)) for d in deg if d > 0]).astype(np.int32)
    return indptr, indices


@pytest.mark.parametrize("n_nodes,M,K,long_rows,long_len", [(20000, 700, 3000, 6, 5000), (3000, 64, 40, 2, 2500), (70000, 3000, 32768, 3, 9000)])
def test_slice_kernels_match_numpy(n_nodes, M, K, long_rows, long_len):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    lib = _native.cabi()
    rng = np.random.Generator(np.random.PCG64(n_nodes))
    indptr, indices = _graph(rng, n_nodes, long_rows, long_len, 40)
    deg = np.diff(indptr)
    nodes = rng.choice(n_nodes, M, replace=True).astype(np.int64)
    nodes[:long_rows] = np.argsort(-deg)[:long_rows]                      # the long rows are in, plus whatever repeats the draw gave
    indptr_d, indices_d, nodes_d = dev(indptr), dev(indices), dev(nodes)

    # U = lap_matrix[nodes, :] and its column counts
    lens = torch.empty(M, dtype=torch.int32, device="cuda")
    fullrowptr = torch.empty(M + 1, dtype=torch.int32, device="cuda")
    _native.check(lib.gnn_row_slice_count(_ptr(indptr_d), _ptr(nodes_d), M, _ptr(lens), _ptr(fullrowptr), _stream()), "row_slice_count")
    ref_full = np.concatenate(([0], np.cumsum(deg[nodes]))).astype(np.int32)
    assert np.array_equal(fullrowptr.cpu().numpy(), ref_full)
    total = int(ref_full[-1])
    assert total % 1024 != 0 and total > 2048
    ucols = torch.full((total,), -9, dtype=torch.int32, device="cuda")
    counts = torch.zeros(n_nodes, dtype=torch.int32, device="cuda")
    _native.check(lib.gnn_row_slice_fill(_ptr(indptr_d), _ptr(indices_d), _ptr(nodes_d), M, _ptr(fullrowptr), _ptr(ucols), _ptr(counts),
                                         _stream()), "row_slice_fill")
    ref_cols = np.concatenate([indices[indptr[v]:indptr[v + 1]] for v in nodes])
    assert np.array_equal(ucols.cpu().numpy(), ref_cols)
    ref_counts = np.bincount(ref_cols, minlength=n_nodes).astype(np.int32)
    assert np.array_equal(counts.cpu().numpy(), ref_counts)

    # support of the counts, compacted in id order
    chunks_n = int(lib.gnn_column_slice_chunks(n_nodes))
    scratch = torch.empty(2 * chunks_n + 2, dtype=torch.int32, device="cuda")
    nz_out = torch.full((n_nodes,), -1, dtype=torch.int64, device="cuda")
    cnt_out = torch.full((n_nodes,), -1, dtype=torch.int32, device="cuda")
    n_sup = torch.zeros(1, dtype=torch.int64, device="cuda")
    _native.check(lib.gnn_support_compact(_ptr(counts), n_nodes, _ptr(scratch), _ptr(nz_out), _ptr(cnt_out), _ptr(n_sup), _stream()),
                  "support_compact")
    ref_nz = np.flatnonzero(ref_counts)
    assert int(n_sup.item()) == ref_nz.size
    assert np.array_equal(nz_out.cpu().numpy()[:ref_nz.size], ref_nz)
    assert np.array_equal(cnt_out.cpu().numpy()[:ref_nz.size], ref_counts[ref_nz])
    assert int((nz_out[ref_nz.size:] != -1).sum().item()) == 0            # nothing written past the support

    # adj = U[:, after_nodes]
    after = np.sort(rng.choice(n_nodes, K, replace=False)).astype(np.int64)
    after_d = dev(after)
    words = (n_nodes + 31) // 32
    bits = torch.zeros(words, dtype=torch.int32, device="cuda")
    rank0 = torch.full((words,), -5, dtype=torch.int32, device="cuda")
    _native.check(lib.gnn_member_set(_ptr(bits), _ptr(rank0), _ptr(after_d), K, 1, _stream()), "member_set")
    chunks = int(lib.gnn_column_slice_chunks(total))
    assert chunks == (total + 1023) // 1024
    chunk_prefix = torch.empty(2 * chunks + 2, dtype=torch.int32, device="cuda")
    rowptr = torch.full((M + 1,), -3, dtype=torch.int32, device="cuda")
    _native.check(lib.gnn_column_slice_count(_ptr(ucols), total, _ptr(fullrowptr), M, _ptr(bits), _ptr(chunk_prefix), _ptr(rowptr),
                                             _stream()), "column_slice_count")
    lookup = np.full(n_nodes, -1, dtype=np.int64)
    lookup[after] = np.arange(K)
    local = lookup[ref_cols]
    keep = local >= 0
    kept_before = np.concatenate(([0], np.cumsum(keep)))
    ref_rowptr = kept_before[ref_full].astype(np.int32)
    assert np.array_equal(rowptr.cpu().numpy(), ref_rowptr)
    nnz = int(ref_rowptr[-1])
    assert nnz > 0
    for dtype, nbytes in ((torch.int16, 2), (torch.int32, 4)):
        colidx = torch.full((nnz,), -2, dtype=dtype, device="cuda")
        _native.check(lib.gnn_column_slice_fill(_ptr(ucols), total, _ptr(bits), _ptr(rank0), _ptr(chunk_prefix), _ptr(colidx), nbytes,
                                                _stream()), "column_slice_fill")
        assert np.array_equal(colidx.cpu().numpy().astype(np.int64), local[keep])
    _native.check(lib.gnn_member_set(_ptr(bits), _ptr(rank0), _ptr(after_d), K, 0, _stream()), "member_set (clear)")
    assert int((bits != 0).sum().item()) == 0
