"""Placement tables consumed by the gather (producer side kept minimal).

The reference computes them offline in ``create_buffer`` (preprocess.py:311-407); that
code is out of scope as a subsystem, but its OUTPUT FORMAT is the input of the hot
path, so the default (non-PaGraph, non-naive) branch is restated here to be able to
produce tables on a box that does not have the reference: access probability
``1^T L[train,:] L^(layers-1)`` (preprocess.py:343-345), the top ``buffer*world`` nodes
cached, every GPU starting from the same top-``buffer`` set, then the greedy
alpha-swap (preprocess.py:355-383).  Checked against tables captured from the
reference in tests/test_placement_golden.py.
"""
from __future__ import annotations

import dataclasses
from typing import List, Sequence

import numpy as np


@dataclasses.dataclass
class Placement:
    device_id_of_nodes_group: List[np.ndarray]      # per rank: holder device id per node, -1 = host
    idx_of_nodes_on_device_group: List[np.ndarray]  # per rank (aliased): slot inside the holder's buffer
    gpu_buffer_group: List[np.ndarray]              # per device: node id of every slot
    sample_prob: np.ndarray


def access_probability(lap_matrix, train_nodes, num_conv_layers: int) -> np.ndarray:
    prob = np.ones(len(train_nodes)) * lap_matrix[train_nodes, :]      # preprocess.py:343
    for _ in range(num_conv_layers - 1):                               # :344-345
        prob = prob * lap_matrix
    return np.asarray(prob).ravel()


def create_placement(lap_matrix, train_nodes, num_nodes_per_dev: int, devices: Sequence[int],
                     num_conv_layers: int, alpha: float = 0.0) -> Placement:
    num_devs = len(devices)
    n = lap_matrix.shape[1]
    sample_prob = access_probability(lap_matrix, train_nodes, num_conv_layers)
    buffered_nodes = np.argsort(-1 * sample_prob)[:num_nodes_per_dev * num_devs]          # :346-347
    idx_of_nodes_on_device = np.arange(n)                                                  # :355
    gpu_buffer_group, device_id_of_nodes_group = [], []
    for i in range(num_devs):                                                              # :356-362
        device_id_of_nodes = np.array([-1] * n)
        gpu_buffer_group.append(buffered_nodes[:num_nodes_per_dev].copy())
        first = buffered_nodes[:num_nodes_per_dev]
        device_id_of_nodes[first] = devices[i]
        device_id_of_nodes_group.append(device_id_of_nodes.copy())
        idx_of_nodes_on_device[first] = np.arange(len(first))
    idx_group = [idx_of_nodes_on_device] * num_devs                                        # :364 (aliased on purpose)
    p_accum = np.array([0.0] * num_devs)
    device_order = np.argsort(p_accum)
    for i in range(len(buffered_nodes) - num_nodes_per_dev):                               # :367-383
        if i % (num_devs - 1) == 0:
            device_order = np.argsort(p_accum)
        candidate = buffered_nodes[num_nodes_per_dev + i]
        new_idx = num_nodes_per_dev - 1 - i // (num_devs - 1)
        replaced = buffered_nodes[new_idx]
        if sample_prob[candidate] >= alpha * sample_prob[replaced]:
            cur = device_order[i % (num_devs - 1)]
            p_accum[cur] += sample_prob[candidate]
            for j in range(num_devs):
                device_id_of_nodes_group[j][candidate] = devices[cur]
                idx_group[j][candidate] = new_idx
            device_id_of_nodes_group[cur][replaced] = devices[device_order[-1]]
            gpu_buffer_group[cur][new_idx] = candidate
        else:
            break
    return Placement(device_id_of_nodes_group, idx_group, gpu_buffer_group, sample_prob)
