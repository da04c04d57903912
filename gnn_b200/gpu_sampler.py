"""LADIES layer sampler with the array work on the device.

Same outputs as the reference ``ladies_sampler`` (sampler.py:90-160) bit for bit, because the only random step -
``np.random.choice(num_nodes, s_num, p=p, replace=False)`` (sampler.py:128) - runs on the host as numpy's own legacy
algorithm restated in C on the same MT19937 stream (gnn_legacy_choice_f64, checked against ``RandomState.choice``), on the
exact same probabilities (integer column counts come back from the device; ``p = pi / np.sum(pi)`` is the reference's
own expression).  What moves to the GPU are the passes that cost the reference ~1.9 s per Reddit-shaped minibatch:

    U = lap_matrix[previous_nodes, :]          (sampler.py:113)   gnn_row_slice_count / _fill
    pi = sp.linalg.norm(U, ord=0, axis=0)      (sampler.py:117)   column counts, fused into the fill pass
    adj = U[:, after_nodes]                    (sampler.py:133)   gnn_member_set + gnn_column_slice_count / _fill (stream compaction)
    create_coo_tensor(...)                     (sampler.py:139)   gnn_build_adj (unchanged)

Two D2H reads per layer synchronise the stream (the column counts - the whole array into pinned memory when the graph
has at most DENSE_COUNTS_MAX nodes, else a device-side compaction - and the kept count); everything else is
asynchronous: host arrays go up through pinned staging (a pageable source makes cudaMemcpyAsync synchronise the stream
before it copies), and the slice size is computed from the host copy of indptr.  The graph structure lives on the
device (int64 indptr, int32 indices).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native
from .sampler import sampled_nodes_remap, sorted_unique


@dataclasses.dataclass
class DeviceLayer:
    fullrowptr: torch.Tensor   # int32 [M+1]
    rowptr: torch.Tensor       # int32 [M+1]
    colidx: torch.Tensor       # int16 / int32 [nnz]
    normfact: torch.Tensor     # fp32 [K]
    nrows: int
    ncols: int


@dataclasses.dataclass
class DeviceMinibatch:
    layers: List[Optional[DeviceLayer]]
    adjs: List[Optional[torch.Tensor]]      # sparse COO tensors with their CSR attached (create_coo_tensor)
    sampled_nodes: List[np.ndarray]
    input_nodes: np.ndarray
    batch_nodes: np.ndarray


DENSE_COUNTS_MAX = 1 << 22      # up to this many nodes the whole count array crosses PCIe (16 MiB) into pinned memory
DEVICE_COMPACT_MIN_NODES = 1 << 20   # from this many nodes on the device compacts the support (gnn_support_compact): the
                                     # host pass over every counter costs more than the extra kernels (measured: 4.7 ms
                                     # per layer on a products-shaped graph, 0.1 ms on a Reddit-shaped one)


def h2d(arr: np.ndarray, device) -> torch.Tensor:
    """Host array -> device without synchronising the stream: through a pinned block of torch's caching host allocator
    (which keeps the block alive until the copy has run)."""
    t = torch.from_numpy(arr)
    if t.numel() == 0:
        return torch.empty(t.shape, dtype=t.dtype, device=device)
    return t.pin_memory().to(device, non_blocking=True)


class SamplerScratch:
    """Per-caller scratch tables (one per sampler thread): membership bitmap of the sampled columns (all zero between
    uses) with its word ranks (gnn_member_set), column counts and, for graphs of at most DENSE_COUNTS_MAX nodes, the pinned
    host mirror of the counts."""
    def __init__(self, num_nodes: int, device):
        words = (num_nodes + 31) // 32
        self.member_bits = torch.zeros(words, dtype=torch.int32, device=device)
        self.member_rank0 = torch.zeros(words, dtype=torch.int32, device=device)
        self.counts = torch.zeros(num_nodes, dtype=torch.int32, device=device)
        self.counts_host = None
        if num_nodes <= DENSE_COUNTS_MAX:
            self.counts_host = torch.zeros(num_nodes, dtype=torch.int32).pin_memory()
            self.counts_np = self.counts_host.numpy()


class DeviceGraph:
    """Structure of the row-normalised adjacency resident on one GPU."""
    def __init__(self, indptr: np.ndarray, indices: np.ndarray, device):
        self.device = torch.device(device)
        self.num_nodes = int(indptr.size - 1)
        self.indptr_host = np.asarray(indptr)      # the caller's array (not copied): slice sizes without a device read
        self.indptr_host_t = (torch.from_numpy(self.indptr_host) if self.indptr_host.dtype == np.int64 and
                              self.indptr_host.flags.c_contiguous else None)
        self.indptr = torch.from_numpy(np.ascontiguousarray(indptr, dtype=np.int64)).to(self.device)
        self.indices = torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int32)).to(self.device)
        self._default_scratch = None

    def scratch(self) -> SamplerScratch:
        return SamplerScratch(self.num_nodes, self.device)

    @property
    def member_bits(self):
        return self.default_scratch().member_bits

    def default_scratch(self) -> SamplerScratch:
        if self._default_scratch is None:
            self._default_scratch = self.scratch()
        return self._default_scratch


def legacy_choice_on_support(rs: np.random.RandomState, p_nz: np.ndarray, size: int) -> np.ndarray:
    """``rs.choice(N, size, p=p, replace=False)`` of numpy's legacy RandomState, evaluated on the SUPPORT of p only, by
    the native restatement ``gnn_legacy_choice_f64`` (include/gnn_b200.h): same MT19937 stream, same rounding, same
    result - about half the time of the numpy expressions below and outside the GIL, which is what the sampler threads
    of one rank contend for.  ``rs`` is advanced exactly as ``choice`` would advance it."""
    import ctypes
    lib = _native.cabi()
    kind, key, pos, has_gauss, gauss = rs.get_state()
    state = np.empty(625, dtype=np.uint32)
    state[:624] = key
    state[624] = pos
    p = np.ascontiguousarray(p_nz, dtype=np.float64)
    found = np.empty(int(size), dtype=np.int64)
    rc = lib.gnn_legacy_choice_f64(ctypes.c_void_p(state.ctypes.data), ctypes.c_void_p(p.ctypes.data), p.size, int(size),
                                   ctypes.c_void_p(found.ctypes.data))
    if rc != 0:
        raise ValueError("Fewer non-zero entries in p than size" if rc == -1 else _native.cabi().gnn_error_string(rc).decode())
    rs.set_state((kind, state[:624].copy(), int(state[624]), has_gauss, gauss))
    return found


def legacy_choice_on_support_numpy(rs: np.random.RandomState, p_nz: np.ndarray, size: int) -> np.ndarray:
    """The same draw in numpy expressions (the round-1/2 implementation, kept as the readable restatement and as a
    cross-check of the native one in the tests).

    ``p_nz`` holds the non-zero probabilities in index order; the result indexes into that support.  It is exactly
    what ``choice`` would return on the full-length p (mtrand.pyx, replace=False branch): zero entries neither change
    the running sums of ``np.cumsum`` nor can ``searchsorted(..., side='right')`` land on them, and the same uniform
    draws are consumed (``random_sample(size - n_uniq)`` per round).  Checked against ``choice`` in the tests."""
    p = np.array(p_nz, dtype=np.float64, copy=True)
    found = np.zeros(size, dtype=np.int64)
    n_uniq = 0
    while n_uniq < size:
        x = rs.random_sample(size - n_uniq)
        if n_uniq > 0:
            p[found[0:n_uniq]] = 0
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        # searchsorted with SORTED needles walks the cdf once instead of 8 K random binary searches over 200 K entries;
        # the positions found are the same, so the draw is unchanged
        xo = np.argsort(x, kind="stable")
        snew = cdf.searchsorted(x[xo], side="right")              # non-decreasing, like x[xo]
        new = np.empty(x.size, dtype=np.int64)
        new[xo] = snew
        # == `_, unique_indices = np.unique(new, return_index=True); unique_indices.sort()`: the FIRST draw of every
        # distinct value, in draw order = the smallest original index inside each run of equal values of the sorted array
        starts = np.flatnonzero(np.concatenate(([True], snew[1:] != snew[:-1])))
        unique_indices = np.sort(np.minimum.reduceat(xo, starts))
        new = new.take(unique_indices)
        found[n_uniq:n_uniq + new.size] = new
        n_uniq += new.size
    return found


def mt_state_of(rs: np.random.RandomState) -> np.ndarray:
    """uint32[625]: the 624 key words of a legacy RandomState followed by its position (what gnn_legacy_choice_f64 and
    gnn_ladies_layer_host advance in place)."""
    _, key, pos, _, _ = rs.get_state()
    state = np.empty(625, dtype=np.uint32)
    state[:624] = key
    state[624] = pos
    return state


_skew_cache = {}


def _sorted_skew_set(skewed_sampling_nodes, layer: int) -> np.ndarray:
    """Ascending distinct ids of one layer's locality set (np.isin of sampler.py:120 does not care about order); sorted
    once per set, not per minibatch."""
    arr = skewed_sampling_nodes[layer]
    key = (id(arr), len(arr))
    hit = _skew_cache.get(key)
    if hit is None or hit[0] is not arr:
        hit = (arr, np.unique(np.asarray(arr, dtype=np.int64)))
        if len(_skew_cache) > 64:
            _skew_cache.clear()
        _skew_cache[key] = hit
    return hit[1]


def host_layer_native(mt_state, nz, cnt, skew, scale_factor, previous_nodes, samp_num, counts_dense=None):
    """gnn_ladies_layer_host through ctypes -> (after_nodes int64, normfact float32, sampled positions int64, s_num).
    With ``counts_dense`` (the count of every node id): gnn_ladies_layer_host_ex, same outputs."""
    import ctypes
    lib = _native.cabi()
    nz = np.ascontiguousarray(nz, dtype=np.int64)
    cnt = np.ascontiguousarray(cnt, dtype=np.int32)
    prev = np.ascontiguousarray(previous_nodes, dtype=np.int64)
    s_num = min(int(nz.size), int(samp_num))
    cap = s_num + prev.size
    after = np.empty(cap, dtype=np.int64)
    normfact = np.empty(cap, dtype=np.float32)
    sampled = np.empty(prev.size, dtype=np.int64)
    n_sampled = ctypes.c_int64(0)
    vp = ctypes.c_void_p
    use_skew = skew is not None and scale_factor > 1
    if counts_dense is not None:
        dense = np.ascontiguousarray(counts_dense, dtype=np.int32)
        n_after = lib.gnn_ladies_layer_host_ex(vp(mt_state.ctypes.data), vp(nz.ctypes.data), vp(cnt.ctypes.data), nz.size,
                                               vp(dense.ctypes.data), dense.size,
                                               vp(skew.ctypes.data) if use_skew else None, skew.size if use_skew else 0,
                                               float(scale_factor), vp(prev.ctypes.data), prev.size, int(samp_num),
                                               vp(after.ctypes.data), vp(normfact.ctypes.data), vp(sampled.ctypes.data),
                                               ctypes.byref(n_sampled))
    else:
        n_after = lib.gnn_ladies_layer_host(vp(mt_state.ctypes.data), vp(nz.ctypes.data), vp(cnt.ctypes.data), nz.size,
                                            vp(skew.ctypes.data) if use_skew else None, skew.size if use_skew else 0,
                                            float(scale_factor), vp(prev.ctypes.data), prev.size, int(samp_num), vp(after.ctypes.data),
                                            vp(normfact.ctypes.data), vp(sampled.ctypes.data), ctypes.byref(n_sampled))
    if n_after < 0:
        _native.check(int(n_after), "gnn_ladies_layer_host")
    return after[:n_after], normfact[:n_after], sampled[:n_sampled.value], s_num


def host_layer_native_dense(mt_state, counts_dense, skew, scale_factor, previous_nodes, samp_num):
    """gnn_ladies_layer_host_dense: the same call on the whole count array (``counts_dense[v]`` for every node v)."""
    import ctypes
    lib = _native.cabi()
    cnt = np.ascontiguousarray(counts_dense, dtype=np.int32)
    prev = np.ascontiguousarray(previous_nodes, dtype=np.int64)
    cap = min(int(cnt.size), int(samp_num)) + prev.size
    after = np.empty(cap, dtype=np.int64)
    normfact = np.empty(cap, dtype=np.float32)
    sampled = np.empty(prev.size, dtype=np.int64)
    n_sampled, n_support = ctypes.c_int64(0), ctypes.c_int64(0)
    vp = ctypes.c_void_p
    use_skew = skew is not None and scale_factor > 1
    n_after = lib.gnn_ladies_layer_host_dense(vp(mt_state.ctypes.data), vp(cnt.ctypes.data), cnt.size,
                                              vp(skew.ctypes.data) if use_skew else None, skew.size if use_skew else 0,
                                              float(scale_factor), vp(prev.ctypes.data), prev.size, int(samp_num),
                                              vp(after.ctypes.data), vp(normfact.ctypes.data), vp(sampled.ctypes.data),
                                              ctypes.byref(n_sampled), ctypes.byref(n_support))
    if n_after < 0:
        _native.check(int(n_after), "gnn_ladies_layer_host_dense")
    return after[:n_after], normfact[:n_after], sampled[:n_sampled.value], min(int(n_support.value), int(samp_num))


def host_layer_numpy(rs, nz, cnt, skew, scale_factor, previous_nodes, samp_num):
    """The same host part in the reference's numpy expressions (sampler.py:117-143 restricted to the support of p):
    the readable restatement, and what the tests compare the native call against."""
    pi_nz = np.asarray(cnt).astype(np.int64)
    if scale_factor > 1 and skew is not None:
        # integer counts stay integer (the reference assigns the scaled values into scipy's int64 count array, which
        # truncates them), so the normaliser below is an exact integer sum on the support as on the full array
        sel = np.isin(nz, skew)
        pi_nz[sel] = pi_nz[sel] * scale_factor
    p_nz = pi_nz / np.sum(pi_nz)                                                            # :124 (same quotients)
    s_num = np.min([nz.size, samp_num])                                                     # :126 (count of p > 0)
    after_nodes = nz[legacy_choice_on_support_numpy(rs, p_nz, int(s_num))]                  # :128
    after_nodes = sorted_unique(np.concatenate((after_nodes, previous_nodes)))              # :131 (np.unique)
    pos = np.minimum(np.searchsorted(nz, after_nodes), nz.size - 1)
    p_after = np.where(nz[pos] == after_nodes, p_nz[pos], 0.0)                              # p[after_nodes]
    normfact = 1 / np.clip(s_num * p_after, 1e-10, 1).astype(np.float32)                    # :137
    return after_nodes, normfact, sampled_nodes_remap(after_nodes, previous_nodes), int(s_num)


def ladies_sample_device(seed: int, batch_nodes, samp_num_list: Sequence[int], graph: DeviceGraph, orders: Sequence[int],
                         create_coo_tensor=None, int16_ids: bool = True, skewed_sampling_nodes=None,
                         scale_factor: float = 1.0, scratch: Optional[SamplerScratch] = None,
                         prebuild_transpose: bool = True, one_call_layers: bool = True) -> DeviceMinibatch:
    """``one_call_layers``: a whole layer runs in one native call with the GIL released (``ladies_layer_device`` of the
    extension; graphs of at most DENSE_COUNTS_MAX nodes).  False keeps the step-by-step Python sequence below - the same
    kernels and the same host function, the readable order of operations, and what larger graphs use."""
    ext = _native.extension()
    if create_coo_tensor is None:
        from .custom_sparse_ops import create_coo_tensor
    dev, n = graph.device, graph.num_nodes
    scratch = scratch or graph.default_scratch()
    # == np.random.seed(seed) + the global legacy functions (sampler.py:96), as a private MT19937 state: thread-safe, and
    # the draws of consecutive layers continue one stream like the reference's
    mt_state = mt_state_of(np.random.RandomState(seed))
    mt_state_t = torch.from_numpy(mt_state.view(np.int32))           # the same 625 words, for the one-call layer
    previous_nodes = np.asarray(batch_nodes)
    batch = previous_nodes
    orders1 = list(orders)[::-1]
    layers: List[Optional[DeviceLayer]] = []
    adjs: List[Optional[torch.Tensor]] = []
    sampled: List[np.ndarray] = []
    for d in range(len(orders1)):
        if orders1[d] == 0:                                                # sampler.py:108-111
            layers.append(None)
            adjs.append(None)
            sampled.append([])
            continue
        prev_np = np.ascontiguousarray(previous_nodes, dtype=np.int64)
        skew = None
        if scale_factor > 1:                                                                # :119-121
            skew = _sorted_skew_set(skewed_sampling_nodes, len(orders1) - d - 1)
        if (one_call_layers and graph.indptr_host_t is not None and
                (scratch.counts_host is not None or n >= DEVICE_COMPACT_MIN_NODES)):
            fullrowptr, rowptr, colidx, nf_dev, after_t, sampled_t = ext.ladies_layer_device(
                graph.indptr, graph.indices, graph.indptr_host_t, scratch.member_bits, scratch.member_rank0, scratch.counts,
                scratch.counts_host, mt_state_t,
                torch.from_numpy(prev_np), torch.from_numpy(skew) if skew is not None else None, float(scale_factor),
                int(samp_num_list[d]), bool(int16_ids), int(DEVICE_COMPACT_MIN_NODES))
            after_nodes = after_t.numpy()
            layer = DeviceLayer(fullrowptr, rowptr, colidx, nf_dev, int(prev_np.size), int(after_nodes.size))
            layers.append(layer)
            adjs.append(create_coo_tensor(fullrowptr, rowptr, colidx, nf_dev, layer.nrows, layer.ncols))   # :139
            sampled.append(sampled_t.numpy())                                               # :143
            previous_nodes = after_nodes
            continue
        prev_dev = h2d(prev_np, dev)
        fullrowptr = ext.row_slice_count(graph.indptr, prev_dev)                           # :113-114
        total = int((graph.indptr_host[prev_np + 1] - graph.indptr_host[prev_np]).sum())    # == fullrowptr[-1], no device read
        scratch.counts.zero_()
        ucols = ext.row_slice_fill(graph.indptr, graph.indices, prev_dev, fullrowptr, total, scratch.counts)
        # :117-143 on the host in one native call (gnn_ladies_layer_host[_dense]): p, s_num, the legacy weighted draw, the
        # union with previous_nodes, normfact and the sampled_nodes remap - same bits as the numpy expressions of
        # host_layer_numpy() below, a fraction of the time and outside the GIL
        if scratch.counts_host is not None:
            # the whole count array in one transfer into pinned memory, compacted by the native call: one synchronisation
            scratch.counts_host.copy_(scratch.counts, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
            after_nodes, normfact, sampled_pos, s_num = host_layer_native_dense(mt_state, scratch.counts_np, skew, scale_factor,
                                                                                 prev_np, int(samp_num_list[d]))
        else:
            # only the columns that occur at all carry probability: compact them on the device and bring back
            # (index, count) pairs instead of an N-long array (N = 111 M on the papers100M shape)
            nz_dev = torch.nonzero(scratch.counts).flatten()
            nz = nz_dev.cpu().numpy()
            cnt = scratch.counts[nz_dev].cpu().numpy()                                      # :117 on the support (int32)
            after_nodes, normfact, sampled_pos, s_num = host_layer_native(mt_state, nz, cnt, skew, scale_factor, prev_np,
                                                                           int(samp_num_list[d]))
        after_dev = h2d(after_nodes, dev)
        ext.member_set(scratch.member_bits, scratch.member_rank0, after_dev, True)
        try:
            rowptr, chunk_prefix = ext.column_slice_count(ucols, fullrowptr, scratch.member_bits)  # :133,135
            nnz = int(rowptr[-1].item())
            use16 = int16_ids and after_nodes.size <= 32768
            colidx = ext.column_slice_fill(ucols, scratch.member_bits, scratch.member_rank0, chunk_prefix, nnz, use16)   # :136
        finally:
            ext.member_set(scratch.member_bits, scratch.member_rank0, after_dev, False)   # all zero again for the next minibatch, whatever happened
        nf_dev = h2d(normfact, dev)
        layer = DeviceLayer(fullrowptr, rowptr, colidx, nf_dev, int(previous_nodes.size), int(after_nodes.size))
        layers.append(layer)
        adjs.append(create_coo_tensor(fullrowptr, rowptr, colidx, nf_dev, layer.nrows, layer.ncols))   # :139
        sampled.append(sampled_pos)                                                         # :143
        previous_nodes = after_nodes
    layers.reverse()
    adjs.reverse()
    sampled.reverse()
    if prebuild_transpose:
        # every layer but the deepest gets a backward: build its A^T index here, on the sampler's
        # stream, so the training stream never pays for it
        from .custom_sparse_ops import adjacency_of
        for adj in adjs[1:]:
            if adj is not None:
                adjacency_of(adj).transpose()
    return DeviceMinibatch(layers, adjs, sampled, np.asarray(previous_nodes, dtype=np.int64), np.asarray(batch))


def record_stream(mb: DeviceMinibatch, stream) -> None:
    """Tell the caching allocator that the tensors of a minibatch built on a sampler stream are used on ``stream``."""
    from .custom_sparse_ops import adjacency_of
    for layer, adj in zip(mb.layers, mb.adjs):
        if layer is None:
            continue
        ts = [layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact, adj._indices(), adj._values()] + adjacency_of(adj).device_tensors()
        for t in ts:
            t.record_stream(stream)
