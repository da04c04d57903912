"""Golden vectors for --locality_sampling (BASELINE configs[3]) from the UNMODIFIED reference Python, CPU only.

    python tests/golden/make_locality_golden.py        (build container only: reads /root/reference)

Executed from the reference, behind the import stubs of make_golden.py:
  * preprocess.create_buffer                  (preprocess.py:311-407)  placement tables, world 4, alpha 0
  * preprocess.get_skewed_sampled_nodes       (preprocess.py:414-423)  on adjacency + I, as main.py:257 calls it
  * sampler.ladies_sampler with scale_factor  (sampler.py:119-121)     2.0 and 1.5 (what the control loop in the
    string literal main.py:200-212 produces: doublings and midpoints)
Output: tests/golden/locality_small.npz
"""
import os
import sys
import tempfile
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

from gnn_b200 import graphgen  # noqa: E402


def main():
    import scipy.sparse as sp
    mg._install_stubs()
    sys.path.insert(0, mg.REF)
    import utils as ref_utils
    import preprocess as ref_pre
    import sampler as ref_sampler

    out = {}
    for tag, shape_name, orders, samp_num, batch, world in [("gcn", "tiny", [1, 1, 1], 96, 32, 4), ("sage", "small", [1, 1, 1], 1024, 128, 4)]:
        shape = graphgen.SHAPES[shape_name]
        g = graphgen.generate(shape, seed=0)
        n = shape.num_nodes
        ip, ix = g.indptr, g.indices
        rows = np.repeat(np.arange(n), np.diff(ip))
        offdiag = rows != ix
        adj = sp.csr_matrix((np.ones(int(offdiag.sum()), np.float32), (rows[offdiag], ix[offdiag])), shape=(n, n))
        lap = sp.csr_matrix(ref_utils.row_normalize(adj + sp.eye(n)) if shape.self_loops else ref_utils.row_normalize(adj))
        lap.sort_indices()
        assert np.array_equal(lap.indices, g.indices)
        feats = graphgen.features(shape, seed=1)
        labels = graphgen.labels(shape, seed=3)
        class_arr = sp.csr_matrix((np.ones(n, np.int32), (np.arange(n), labels)), shape=(n, shape.num_classes))
        devices = list(range(world))
        graph_data = (adj, class_arr, mg._CpuFeat(torch.from_numpy(feats)), shape.num_classes, g.train_nodes, g.valid_nodes, g.test_nodes)
        buffer_size = int(0.1 * n)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, "save"))
            os.chdir(tmp)
            try:
                did_group, idx_group, _, gpu_buffer_group, _ = ref_pre.create_buffer(lap, graph_data, buffer_size, devices,
                                                                                      "loc_" + tag, sum(orders), alpha=0.0)
            finally:
                os.chdir(cwd)
        sets = ref_pre.get_skewed_sampled_nodes(adj + sp.eye(n), gpu_buffer_group, orders)         # main.py:257
        p = tag + "_"
        out[p + "shape"] = shape_name
        out[p + "orders"] = np.array(orders)
        out[p + "samp_num"], out[p + "batch"], out[p + "world"], out[p + "buffer_size"] = samp_num, batch, world, buffer_size
        out[p + "gpu_buffer_group"] = np.stack([np.asarray(b) for b in gpu_buffer_group])
        for i, s in enumerate(sets):
            out[p + f"set{i}"] = np.asarray(s)
        rng = np.random.Generator(np.random.PCG64(5))
        for ci, (seed, sf) in enumerate([(21, 2.0), (22, 1.5), (23, 16.0)]):
            batch_nodes = g.train_nodes[rng.permutation(g.train_nodes.size)[:batch]]
            with mock.patch.object(torch.Tensor, "to", mg._cpu_to):
                res = ref_sampler.ladies_sampler(seed, batch_nodes, np.array([samp_num] * 5), n, lap, class_arr, orders,
                                                 did_group[0], idx_group[0], sets, sf, 0, devices)
            adjs, _, _, _, _, n0, _, sampled_nodes = res
            c = p + f"c{ci}_"
            out[c + "seed"], out[c + "scale_factor"], out[c + "batch_nodes"], out[c + "n0"] = seed, sf, batch_nodes, n0
            for li, a in enumerate(adjs):
                for k in ["fullrowptr", "rowptr", "colidx", "normfact", "nrows", "ncols"]:
                    out[c + f"l{li}_" + k] = getattr(a, k)
                out[c + f"l{li}_sampled_nodes"] = np.asarray(sampled_nodes[li])
    np.savez_compressed(os.path.join(mg.OUT, "locality_small.npz"), **out)
    print("wrote locality_small.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
