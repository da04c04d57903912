"""Placement tables consumed by the feature gather, and the locality-sampling node sets.

The reference computes both offline (``create_buffer``, preprocess.py:311-407, default branch; and
``get_skewed_sampled_nodes``, preprocess.py:414-423).  Their OUTPUT FORMAT is the input of the hot path
(sampler.py:150-158 reads ``device_id_of_nodes`` / ``idx_of_nodes_on_device``; sampler.py:119-121 reads the skew sets),
so a box without the reference needs a producer.  This one is organised differently from the reference's
per-candidate Python loop:

* access probability ``1^T L[train,:] L^(layers-1)`` by CSR row-vector products on the bare structure arrays;
* the alpha test ``prob[candidate] >= alpha * prob[replaced]`` does not depend on the running device loads, so the
  point where the reference's loop breaks is found up front with one vector comparison;
* the only sequential part - every ``world-1`` candidates the devices are re-ranked by accumulated probability - runs
  once per ROUND (``buffer`` rounds instead of ``buffer * (world-1)`` candidate steps), recording for each round the
  device order; the table updates (all ranks' views, the per-rank "my replaced copy now lives on the device that kept
  it" entries, the buffer slots) are then applied as whole-array assignments.

Tables are bit-identical to the reference's (tests/test_placement_golden.py, fixtures captured from the unmodified
reference for alpha = 0 and 0.5), including its tie-breaking (same ``np.argsort`` calls on the same values).
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence

import numpy as np


@dataclasses.dataclass
class Placement:
    device_id_of_nodes_group: List[np.ndarray]      # per rank: holder device id per node, -1 = host
    idx_of_nodes_on_device_group: List[np.ndarray]  # per rank (one shared array, like the reference): slot inside the holder's buffer
    gpu_buffer_group: List[np.ndarray]              # per device: node id of every slot
    sample_prob: np.ndarray
    accepted: int = 0                               # candidates placed before the alpha test failed


def _csr_parts(lap_matrix):
    """(indptr, indices, data) of a scipy CSR matrix or of a (indptr, indices, data) tuple."""
    if isinstance(lap_matrix, tuple):
        return lap_matrix
    return lap_matrix.indptr, lap_matrix.indices, lap_matrix.data


def _vec_times_csr(v: np.ndarray, indptr: np.ndarray, indices: np.ndarray, data, n_cols: int) -> np.ndarray:
    """Row vector times CSR matrix: out[c] = sum_r v[r] * A[r, c] (float64)."""
    w = np.repeat(np.asarray(v, dtype=np.float64), np.diff(indptr))
    if data is not None:
        w = w * data
    return np.bincount(indices, weights=w, minlength=n_cols)


def access_probability(lap_matrix, train_nodes, num_conv_layers: int) -> np.ndarray:
    """reference preprocess.py:343-345: ``ones(len(train)) * L[train, :]`` then ``* L`` per further layer.

    Same float64 additions in the same order as scipy's row-vector x CSR product (entries visited row by row, the
    rows of ``L[train, :]`` in ``train_nodes`` order), so the probabilities - and therefore the ranking, ties
    included - are bit-identical to the reference's."""
    indptr, indices, data = _csr_parts(lap_matrix)
    n = indptr.size - 1
    train = np.asarray(train_nodes, dtype=np.int64)
    starts, lens = indptr[train], indptr[train + 1] - indptr[train]
    offs = np.zeros(train.size + 1, dtype=np.int64)
    np.cumsum(lens, out=offs[1:])
    entry = np.repeat(starts - offs[:-1], lens) + np.arange(int(offs[-1]), dtype=np.int64)      # entries of L[train, :]
    prob = np.bincount(indices[entry], weights=None if data is None else data[entry].astype(np.float64), minlength=n).astype(np.float64)
    for _ in range(num_conv_layers - 1):
        prob = _vec_times_csr(prob, indptr, indices, data, n)
    return prob


def create_placement(lap_matrix, train_nodes, num_nodes_per_dev: int, devices: Sequence[int],
                     num_conv_layers: int, alpha: float = 0.0, sample_prob: np.ndarray | None = None) -> Placement:
    world = len(devices)
    devs = np.asarray(devices)
    indptr = _csr_parts(lap_matrix)[0]
    n = indptr.size - 1
    if sample_prob is None:
        sample_prob = access_probability(lap_matrix, train_nodes, num_conv_layers)
    ranked = np.argsort(-1 * sample_prob)[:num_nodes_per_dev * world]
    top, cand = ranked[:num_nodes_per_dev], ranked[num_nodes_per_dev:]

    # start: every GPU holds the same top set, every rank sees those nodes on itself
    views = []
    for d in range(world):
        view = np.full(n, -1, dtype=np.int64)
        view[top] = devs[d]
        views.append(view)
    slot_of = np.arange(n)
    slot_of[top] = np.arange(top.size)
    buffers = [top.copy() for _ in range(world)]

    accepted = 0
    if world > 1 and cand.size:
        per_round = world - 1
        new_slot = num_nodes_per_dev - 1 - np.arange(cand.size) // per_round     # slot each candidate would take
        replaced = ranked[new_slot]                                               # ... and the node it evicts there
        passes = sample_prob[cand] >= alpha * sample_prob[replaced]
        accepted = int(cand.size if passes.all() else np.argmin(passes))          # the reference breaks at the first failure
        rounds = -(-accepted // per_round)
        # the sequential part: device ranking at the start of each round
        load = np.zeros(world)
        holder = np.empty(accepted, dtype=np.int64)      # device index that takes candidate i
        keeper = np.empty(rounds, dtype=np.int64)        # device index that receives nothing in round r and keeps the evicted node
        cand_prob = sample_prob[cand[:accepted]]
        for r in range(rounds):
            order = np.argsort(load)
            lo, hi = r * per_round, min((r + 1) * per_round, accepted)
            holder[lo:hi] = order[:hi - lo]
            keeper[r] = order[-1]
            load[order[:hi - lo]] += cand_prob[lo:hi]
        acc_cand, acc_slot, acc_replaced = cand[:accepted], new_slot[:accepted], replaced[:accepted]
        for view in views:
            view[acc_cand] = devs[holder]
        slot_of[acc_cand] = acc_slot
        round_of = np.arange(accepted) // per_round
        for d in range(world):
            mine = holder == d
            views[d][acc_replaced[mine]] = devs[keeper[round_of[mine]]]
            buffers[d][acc_slot[mine]] = acc_cand[mine]
    return Placement(views, [slot_of] * world, buffers, sample_prob, accepted)


def locality_sampling_sets(indptr: np.ndarray, indices: np.ndarray, has_self_loops: bool,
                           gpu_buffer_group: Sequence[np.ndarray], num_layers: int, top: int = 8192) -> List[np.ndarray]:
    """Node sets whose sampling probability ``--locality_sampling`` scales up (reference preprocess.py:414-423 on
    ``adjacency + I``): layer 0 = every node cached on some GPU; layer i = the ``top`` nodes with the most i-hop
    paths from cached nodes.  Tie order follows the same ``np.argsort(-1 * v)`` call on the same float64 counts."""
    n = indptr.size - 1
    sets = [np.unique(np.concatenate([np.asarray(b) for b in gpu_buffer_group]))]
    v = np.zeros(n, dtype=np.float64)
    v[sets[0]] = 1
    for _ in range(1, num_layers):
        nxt = _vec_times_csr(v, indptr, indices, None, n)
        if not has_self_loops:
            nxt = nxt + v                                  # the "+ I" of preprocess.py call site main.py:257
        v = nxt
        sets.append(np.argsort(-1 * v)[:top])
    return sets


class ScaleFactorController:
    """The per-epoch adjustment of the locality-sampling ``scale_factor`` that the reference keeps - switched off, inside
    a string literal - at main.py:200-212: while the input-feature movement takes a fifth or more of the epoch the factor
    doubles (up to 16); once it falls under a tenth the factor settles halfway between the last two values; anything in
    between ends the search.  Same branches, same order, same constants; ``update`` is called once per epoch with the
    two times the reference accumulates (main.py:140-145 ``data_movement_time``, ``execution_time``) and returns the
    factor for the next epoch's ``prepare_data`` calls (main.py:117)."""

    def __init__(self, scale_factor: float = 1.0):
        self.scale_factor = float(scale_factor)
        self.factor_increase = True
        self.factor_before = float(scale_factor)
        self.factor_after = float(scale_factor)

    def update(self, data_movement_time: float, execution_time: float) -> float:
        if self.factor_increase:
            ratio = data_movement_time / execution_time
            if self.scale_factor >= 16:
                self.factor_increase = False
            elif ratio >= 0.2:
                self.factor_before = self.scale_factor
                self.scale_factor *= 2
            elif ratio < 0.1 and self.scale_factor != 1:
                self.factor_after = self.scale_factor
                self.scale_factor = (self.factor_before + self.factor_after) / 2
            else:
                self.factor_increase = False
        return self.scale_factor
