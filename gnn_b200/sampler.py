"""Host-side LADIES layer sampler and the per-minibatch remaps that feed the device path.

This mirrors reference sampler.py:90-160 (``ladies_sampler``).  The sampling
itself (sampler.py:113-131) is host numpy in the reference and stays host numpy
here - it is the *input generator* of the hot path, and it draws from a numpy
legacy ``RandomState`` seeded and called exactly like the reference's global
functions (``np.random.seed`` + ``np.random.choice(p=..., replace=False)``), so
sampled node sets are bit-identical to the reference by construction.  What is
restated differently is the data handling around it:

* row slice ``lap_matrix[previous_nodes, :]`` (sampler.py:113) and column slice
  ``U[:, after_nodes]`` (sampler.py:133) work on bare ``indptr``/``indices``
  arrays (a gather and a lookup-table filter) instead of scipy fancy indexing;
* the column-count ``sp.linalg.norm(U, ord=0, axis=0)`` (sampler.py:117) is a
  ``bincount``;
* the CSR hand-off (sampler.py:114,135-137) is returned as a :class:`LayerCSR`
  so callers decide when/where to upload it;
* the ``sampled_nodes`` remap (sampler.py:143) and the placement remap
  (sampler.py:150-158) are separate functions with a compact
  ``(src_dev, slot)`` form for the device gather (gnn_b200.gather).

Results are checked bit-for-bit against the unmodified reference sampler in
tests/test_sampler_golden.py (fixtures made by tests/golden/make_golden.py).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional, Sequence

import numpy as np


@dataclasses.dataclass
class LayerCSR:
    """One sampled layer adjacency as the reference hands it to create_coo_tensor
    (sampler.py:114,135-137 -> spmm.cpp:44).  All host numpy."""
    fullrowptr: np.ndarray   # int32 [M+1]  row pointer of the *unsliced* rows (full-graph degree)
    rowptr: np.ndarray       # int32 [M+1]
    colidx: np.ndarray       # int16 [nnz]  (reference sampler.py:136; wraps for K > 32767)
    normfact: np.ndarray     # fp32 [K]     1/clip(s_num * p[after], 1e-10, 1)
    nrows: int
    ncols: int
    colidx32: Optional[np.ndarray] = None  # int32 copy, exact for any K

    @property
    def nnz(self) -> int:
        return int(self.rowptr[-1])


@dataclasses.dataclass
class Minibatch:
    layers: List[Optional[LayerCSR]]     # index 0 = deepest layer (consumes input features)
    sampled_nodes: List[np.ndarray]      # per layer: positions of output nodes inside the input-node list
    input_nodes: np.ndarray              # int64 sorted unique global ids; row j of X <-> input_nodes[j]
    batch_nodes: np.ndarray              # the seed nodes (rows of the top layer)


def row_slice(indptr: np.ndarray, indices: np.ndarray, nodes: np.ndarray):
    """``lap_matrix[nodes, :]`` structure: returns (fullrowptr int64 [M+1], col ids of every entry)."""
    nodes = np.asarray(nodes, dtype=np.int64)
    starts = indptr[nodes]
    lens = indptr[nodes + 1] - starts
    fullrowptr = np.zeros(nodes.size + 1, dtype=np.int64)
    np.cumsum(lens, out=fullrowptr[1:])
    total = int(fullrowptr[-1])
    gather = np.repeat(starts - fullrowptr[:-1], lens) + np.arange(total, dtype=np.int64)
    return fullrowptr, indices[gather], lens


def column_slice(u_cols: np.ndarray, lens: np.ndarray, after_nodes: np.ndarray, num_nodes: int):
    """``U[:, after_nodes]`` structure for sorted unique ``after_nodes``: (rowptr int64, local col ids int32)."""
    k = after_nodes.size
    if num_nodes <= (1 << 27):
        lookup = np.full(num_nodes, -1, dtype=np.int32)
        lookup[after_nodes] = np.arange(k, dtype=np.int32)
        local = lookup[u_cols]
        keep = local >= 0
    else:  # huge graphs: no O(N) scratch per call
        pos = np.searchsorted(after_nodes, u_cols)
        np.minimum(pos, k - 1, out=pos)
        keep = after_nodes[pos] == u_cols
        local = pos.astype(np.int32)
    row_of_entry = np.repeat(np.arange(lens.size, dtype=np.int64), lens)
    counts = np.bincount(row_of_entry[keep], minlength=lens.size)
    rowptr = np.zeros(lens.size + 1, dtype=np.int64)
    np.cumsum(counts, out=rowptr[1:])
    return rowptr, local[keep]


def sorted_unique(a: np.ndarray) -> np.ndarray:
    """``np.unique(a)`` through sort + neighbour compare (numpy 2.3's hash-based unique is several times slower here)."""
    if a.size == 0:
        return a
    s = np.sort(a)
    keep = np.empty(s.size, dtype=bool)
    keep[0] = True
    np.not_equal(s[1:], s[:-1], out=keep[1:])
    return s[keep]


def sampled_nodes_remap(after_nodes: np.ndarray, previous_nodes: np.ndarray) -> np.ndarray:
    """reference sampler.py:143, ``np.where(np.in1d(after_nodes, previous_nodes))[0]``: positions inside the
    sorted-unique ``after_nodes`` of the nodes that also occur in ``previous_nodes`` (ascending)."""
    prev = sorted_unique(np.asarray(previous_nodes))
    pos = np.searchsorted(after_nodes, prev)
    ok = pos < after_nodes.size
    ok[ok] = after_nodes[pos[ok]] == prev[ok]
    return pos[ok]


def ladies_sample(seed: int, batch_nodes, samp_num_list: Sequence[int], num_nodes: int,
                  indptr: np.ndarray, indices: np.ndarray, orders: Sequence[int],
                  skewed_sampling_nodes=None, scale_factor: float = 1.0) -> Minibatch:
    """Host part of reference ``ladies_sampler`` (sampler.py:90-147), returning host arrays."""
    # a private legacy RandomState seeded like np.random.seed(seed) (sampler.py:96): the same MT19937 stream and
    # the same choice() algorithm as the global functions the reference calls, but safe under sampler threads
    rs = np.random.RandomState(seed)
    previous_nodes = np.asarray(batch_nodes)
    batch = previous_nodes
    orders1 = list(orders)[::-1]
    layers: List[Optional[LayerCSR]] = []
    sampled_nodes: List[np.ndarray] = []
    for d in range(len(orders1)):
        if orders1[d] == 0:                                # sampler.py:108-111
            layers.append(None)
            sampled_nodes.append([])
            continue
        fullrowptr, u_cols, lens = row_slice(indptr, indices, previous_nodes)      # :113-114
        pi = np.bincount(u_cols, minlength=num_nodes)                              # :117
        if scale_factor > 1:                                                       # :119-121
            # `pi` is an INTEGER array in the reference too (scipy's ord=0 column norm counts nonzeros as int64), so the
            # scaled counts are truncated toward zero by the assignment - 3 * 1.5 -> 4.  Kept: sampled sets stay identical.
            sel = skewed_sampling_nodes[len(orders1) - d - 1]
            pi[sel] = pi[sel] * scale_factor
        p = pi / np.sum(pi)                                                        # :124
        s_num = np.min([np.sum(p > 0), samp_num_list[d]])                          # :126
        after_nodes = rs.choice(num_nodes, s_num, p=p, replace=False)              # :128
        after_nodes = sorted_unique(np.concatenate((after_nodes, previous_nodes)))  # :131 (np.unique)
        rowptr, local_cols = column_slice(u_cols, lens, after_nodes, num_nodes)    # :133
        normfact = 1 / np.clip(s_num * p[after_nodes], 1e-10, 1).astype(np.float32)  # :137
        layers.append(LayerCSR(
            fullrowptr=fullrowptr.astype(np.int32), rowptr=rowptr.astype(np.int32),
            colidx=local_cols.astype(np.int16), normfact=normfact,
            nrows=int(lens.size), ncols=int(after_nodes.size),
            colidx32=local_cols.astype(np.int32, copy=False)))
        sampled_nodes.append(sampled_nodes_remap(after_nodes, previous_nodes))     # :143
        previous_nodes = after_nodes
    layers.reverse()                                                               # :147-148
    sampled_nodes.reverse()
    return Minibatch(layers, sampled_nodes, np.asarray(previous_nodes, dtype=np.int64), np.asarray(batch))


@dataclasses.dataclass
class PlacementRemap:
    """reference sampler.py:150-158 output plus the compact form the device gather consumes."""
    mask_on_devices: List[np.ndarray]   # world_size bool masks [n0]
    mask_on_cpu: np.ndarray             # bool [n0]
    idx_on_devices: List[np.ndarray]    # world_size int arrays: slot inside device i's buffer
    idx_on_cpu: np.ndarray              # global ids of uncached input nodes
    src_dev: np.ndarray                 # int32 [n0]: index into ``devices`` of the holder, -1 = host
    slot: np.ndarray                    # int64 [n0]: row inside the holder's buffer (global id for host)


def placement_remap(input_nodes: np.ndarray, device_id_of_nodes: np.ndarray,
                    idx_of_nodes_on_device: np.ndarray, devices: Sequence[int]) -> PlacementRemap:
    input_nodes_devices = device_id_of_nodes[input_nodes]                      # sampler.py:152
    mask_cpu = input_nodes_devices == -1                                       # :153
    idx_cpu = input_nodes[mask_cpu]                                            # :154
    masks, idxs = [], []
    src_dev = np.full(input_nodes.size, -1, dtype=np.int32)
    slot = np.asarray(input_nodes, dtype=np.int64).copy()
    for i in range(len(devices)):                                              # :156-158
        m = input_nodes_devices == devices[i]
        masks.append(m)
        idxs.append(idx_of_nodes_on_device[input_nodes[m]].copy())
        src_dev[m] = i
        slot[m] = idxs[-1]
    return PlacementRemap(masks, mask_cpu, idxs, idx_cpu, src_dev, slot)


def rank_batches(num_train: int, batch_size: int, rank: int, world_size: int, iter_num: int):
    """Minibatch scheduling of reference sampler.py:166-185 (global shuffle, contiguous chunk per rank).

    Returns a list of index arrays into ``train_nodes``; every rank gets the same
    number of batches only when the chunks are equal - callers that allreduce must
    equalise step counts."""
    import torch
    chunk = num_train // world_size + (1 if num_train % world_size else 0)
    start = rank * chunk
    end = min((rank + 1) * chunk, num_train)
    torch.manual_seed(iter_num)
    idxs = torch.randperm(num_train).numpy()
    nb = (end - start) // batch_size + (1 if (end - start) % batch_size else 0)
    return [idxs[start + j * batch_size: min(start + (j + 1) * batch_size, end)] for j in range(nb)]
