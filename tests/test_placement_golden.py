"""Placement-table producer vs tables captured from reference preprocess.create_buffer."""
import os

import numpy as np
import pytest

from gnn_b200 import graphgen, placement

@pytest.mark.parametrize("name", ["cora_gcn", "tiny_sage3", "tiny_order0"])
def test_tables_match_reference(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    shape = graphgen.SHAPES[str(z["shape"])]
    g = graphgen.generate(shape, seed=0)
    world = int(z["world"])
    if world == 1:
        pytest.skip("reference create_buffer divides by (num_devs-1)")
    pl = placement.create_placement(g.to_scipy(np.float64), g.train_nodes, int(z["buffer_size"]), list(range(world)),
                                    int(np.sum(z["orders"])), alpha=float(z["alpha"]))
    for r in range(world):
        assert np.array_equal(pl.device_id_of_nodes_group[r], z["device_id_of_nodes_group"][r])
        assert np.array_equal(pl.idx_of_nodes_on_device_group[r], z["idx_of_nodes_on_device_group"][r])
        assert np.array_equal(pl.gpu_buffer_group[r], z["gpu_buffer_group"][r])
