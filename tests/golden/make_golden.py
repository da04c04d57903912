"""Generate tests/golden/*.npz by running the UNMODIFIED reference Python on CPU.

Run in the build container only (reads /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

What is executed from the reference, unmodified, behind import stubs for the four
modules missing offline (matplotlib, matplotlib.pyplot, ogb.nodeproppred,
torch_geometric.utils - SURVEY.md section 8(c)):

  * utils.row_normalize                      (utils.py:56-64)
  * preprocess.create_buffer                 (preprocess.py:311-407), alpha = 0 and 0.5
  * sampler.ladies_sampler                   (sampler.py:90-160)
      - its `custom_sparse_ops.create_coo_tensor` calls are captured (the op itself is
        CUDA-only; its GPU outputs are pinned separately by make_golden_gpu.py)
      - integer device ids are kept (placement tables compare against them) and
        Tensor.to(<int>) is neutralised so uploads stay on the CPU
  * torch.sparse `mat1.mm(mat2)` / `mat1.transpose(0,1).mm(g)` - the reference's own
    commented CPU alternative (custom_sparse_ops.py:25,36) on the captured adjacencies.

main.py cannot be imported (argparse + globals at module level), so its gather block
(main.py:129-134) is restated below on CPU tensors, line for line.
"""
import os
import sys
import tempfile
import types
from unittest import mock

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, REPO)

from gnn_b200 import graphgen  # noqa: E402


class _Captured:
    """Stands in for the sparse tensor create_coo_tensor returns."""
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _install_stubs():
    for name in ["matplotlib", "matplotlib.pyplot", "ogb", "ogb.nodeproppred", "torch_geometric", "torch_geometric.utils"]:
        sys.modules[name] = types.ModuleType(name)
    sys.modules["ogb.nodeproppred"].PygNodePropPredDataset = object
    sys.modules["torch_geometric.utils"].to_undirected = None
    sys.modules["torch_geometric.utils"].dropout_adj = None
    cso = types.ModuleType("custom_sparse_ops")

    def create_coo_tensor(fullrowptr, rowptr, colidx, normfact, nrows, ncols):
        return _Captured(fullrowptr=fullrowptr.numpy().copy(), rowptr=rowptr.numpy().copy(),
                         colidx=colidx.numpy().copy(), normfact=normfact.numpy().copy(),
                         nrows=int(nrows), ncols=int(ncols))
    cso.create_coo_tensor = create_coo_tensor
    cso.spmm = None
    sys.modules["custom_sparse_ops"] = cso


class _CpuRows:
    def __init__(self, t):
        self.t = t

    def to(self, dev):
        return self.t


class _CpuFeat:
    """feat_data whose row-selection ignores `.to(int_device)` (preprocess.py:399)."""
    def __init__(self, t):
        self.t = t
        self.shape = t.shape

    def __getitem__(self, idx):
        return _CpuRows(self.t[idx])


_orig_to = torch.Tensor.to


def _cpu_to(self, *args, **kwargs):
    if args and isinstance(args[0], (int, np.integer)):
        return self
    return _orig_to(self, *args, **kwargs)


def main():
    import scipy.sparse as sp
    _install_stubs()
    sys.path.insert(0, REF)
    import utils as ref_utils          # noqa: F401
    import preprocess as ref_pre
    import sampler as ref_sampler

    cases = [
        # name, shape, model(+I?), orders, samp_num, batch, world, buffer_frac, alpha, seeds
        ("cora_gcn", "cora", [1, 1], 512, 256, 2, 0.1, 0.0, [1234, 1235]),
        ("tiny_sage3", "tiny", [1, 1, 1], 96, 32, 4, 0.1, 0.5, [7, 8]),
        ("tiny_order0", "tiny", [1, 0, 1], 64, 16, 1, 0.2, 0.0, [11]),
    ]
    for name, shape_name, orders, samp_num, batch, world, buf_frac, alpha, seeds in cases:
        shape = graphgen.SHAPES[shape_name]
        g = graphgen.generate(shape, seed=0)
        n = shape.num_nodes
        # adjacency WITHOUT self loops, then the reference's own normalisation (main.py:267-270)
        ip, ix = g.indptr, g.indices
        rows = np.repeat(np.arange(n), np.diff(ip))
        offdiag = rows != ix
        adj = sp.csr_matrix((np.ones(int(offdiag.sum()), np.float32), (rows[offdiag], ix[offdiag])), shape=(n, n))
        lap = ref_utils.row_normalize(adj + sp.eye(n)) if shape.self_loops else ref_utils.row_normalize(adj)
        lap = sp.csr_matrix(lap)
        lap.sort_indices()
        assert np.array_equal(lap.indptr, g.indptr) and np.array_equal(lap.indices, g.indices), "graphgen structure != reference lap_matrix"

        feats = graphgen.features(shape, seed=1)
        labels = graphgen.labels(shape, seed=3)
        class_arr = sp.csr_matrix((np.ones(n, np.int32), (np.arange(n), labels)), shape=(n, shape.num_classes))
        devices = list(range(world))
        graph_data = (adj, class_arr, _CpuFeat(torch.from_numpy(feats)), shape.num_classes, g.train_nodes, g.valid_nodes, g.test_nodes)
        buffer_size = int(buf_frac * n)
        cwd = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.makedirs(os.path.join(tmp, "save"))
            os.chdir(tmp)
            try:
                did_group, idx_group, gpu_buffers, gpu_buffer_group, _ = ref_pre.create_buffer(
                    lap, graph_data, buffer_size, devices, name, sum(orders), alpha=alpha)
            finally:
                os.chdir(cwd)
        out = {
            "shape": shape_name, "orders": np.array(orders), "samp_num": samp_num, "batch": batch, "world": world,
            "buffer_size": buffer_size, "alpha": alpha, "seeds": np.array(seeds),
            "device_id_of_nodes_group": np.stack([np.asarray(d) for d in did_group]),
            "idx_of_nodes_on_device_group": np.stack([np.asarray(d) for d in idx_group]),
            "gpu_buffer_group": np.stack([np.asarray(b) for b in gpu_buffer_group]),
        }
        for b, buf in enumerate(gpu_buffers):
            assert np.array_equal(buf.numpy(), feats[np.asarray(gpu_buffer_group[b])])

        rng = np.random.Generator(np.random.PCG64(99))
        for si, seed in enumerate(seeds):
            rank = si % world
            batch_nodes = g.train_nodes[rng.permutation(g.train_nodes.size)[:batch]]
            with mock.patch.object(torch.Tensor, "to", _cpu_to):
                res = ref_sampler.ladies_sampler(seed, batch_nodes, np.array([samp_num] * 5), n, lap, class_arr, orders,
                                                 did_group[rank], idx_group[rank], None, 1.0, rank, devices)
            adjs, masks_dev, mask_cpu, idx_dev, idx_cpu, n0, out_label, sampled_nodes = res
            pre = f"s{si}_"
            out[pre + "seed"] = seed
            out[pre + "rank"] = rank
            out[pre + "batch_nodes"] = batch_nodes
            out[pre + "n0"] = n0
            out[pre + "nlayers"] = len(adjs)
            out[pre + "labels_dense"] = out_label.numpy()
            out[pre + "mask_cpu"] = mask_cpu
            out[pre + "idx_cpu"] = idx_cpu
            for i in range(world):
                out[pre + f"mask_dev{i}"] = masks_dev[i]
                out[pre + f"idx_dev{i}"] = np.asarray(idx_dev[i])
            # ---- gather, main.py:129-134 restated on CPU tensors ----
            feat_data = torch.from_numpy(feats)
            input_feat = torch.zeros(n0, feats.shape[1])
            for i in range(world):
                input_feat[torch.from_numpy(masks_dev[i])] = gpu_buffers[i][torch.from_numpy(np.asarray(idx_dev[i]))].float()
            input_feat[torch.from_numpy(mask_cpu)] = feat_data[torch.from_numpy(idx_cpu)].float()
            out[pre + "input_feat_sha"] = np.frombuffer(
                __import__("hashlib").sha256(input_feat.numpy().tobytes()).digest(), dtype=np.uint8)
            if shape.feat_dim * n0 <= 200000:
                out[pre + "input_feat"] = input_feat.numpy()
            x = input_feat
            for li, a in enumerate(adjs):
                lp = pre + f"l{li}_"
                if a is None:
                    out[lp + "none"] = 1
                    continue
                for k in ["fullrowptr", "rowptr", "colidx", "normfact", "nrows", "ncols"]:
                    out[lp + k] = getattr(a, k)
                out[lp + "sampled_nodes"] = np.asarray(sampled_nodes[li])
                # values by an independent numpy statement of cuda_spmm.cu:800 (double math, one rounding)
                deg = np.diff(a.fullrowptr).astype(np.float64)
                rows_i = np.repeat(np.arange(a.nrows), np.diff(a.rowptr))
                vals = ((1.0 / deg[rows_i]) * a.normfact[a.colidx.astype(np.int64)].astype(np.float64)).astype(np.float32)
                out[lp + "values"] = vals
                idx = torch.from_numpy(np.vstack([rows_i, a.colidx.astype(np.int64)]))
                mat1 = torch.sparse_coo_tensor(idx, torch.from_numpy(vals), (a.nrows, a.ncols)).coalesce()
                d = min(x.shape[1], 48)              # keep fixtures small: first 48 feature columns
                xin = x[:, :d].contiguous() if x.shape[0] == a.ncols else torch.from_numpy(
                    graphgen.features(graphgen.GraphShape("x", a.ncols, 0, d, 1, 1), seed=5 + li))
                y = mat1.mm(xin)                                        # custom_sparse_ops.py:25
                gout = torch.from_numpy(graphgen.features(graphgen.GraphShape("g", a.nrows, 0, d, 1, 1), seed=17 + li))
                dx = mat1.transpose(0, 1).mm(gout)                      # custom_sparse_ops.py:36
                out[lp + "x"] = xin.numpy()
                out[lp + "y_torchsparse"] = y.numpy()
                out[lp + "g"] = gout.numpy()
                out[lp + "dx_torchsparse"] = dx.numpy()
                x = y[torch.from_numpy(np.asarray(sampled_nodes[li]))] if False else y
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
        print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(out.items())[:6]})


if __name__ == "__main__":
    main()
