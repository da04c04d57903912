"""Golden outputs of the UNMODIFIED reference models.py (GraphSage + GNN head, GCN) on CPU, for the harness model check.

Run in the build container only: python tests/golden/make_model_golden.py
The reference's `custom_sparse_ops.spmm` is stubbed by torch.sparse.mm (its own commented alternative,
custom_sparse_ops.py:25); dropout is 0 so the forward is deterministic."""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
for name in ["matplotlib", "matplotlib.pyplot"]:
    sys.modules[name] = types.ModuleType(name)
cso = types.ModuleType("custom_sparse_ops")
cso.spmm = lambda a, x: torch.sparse.mm(a, x)
sys.modules["custom_sparse_ops"] = cso
sys.path.insert(0, "/root/reference")
import models as ref_models  # noqa: E402
from gnn_b200 import graphgen, sampler  # noqa: E402
import oracle  # noqa: E402


def main():
    shape = graphgen.SHAPES["tiny"]
    g = graphgen.generate(shape, seed=0)
    feats = graphgen.features(shape, seed=1)
    mb = sampler.ladies_sample(5, g.train_nodes[:24], [64] * 3, shape.num_nodes, g.indptr, g.indices, [1, 1, 1])
    adjs = []
    for l in mb.layers:
        r, c, v = oracle.build_adj(l.fullrowptr, l.rowptr, l.colidx, l.normfact, l.nrows)
        adjs.append(torch.sparse_coo_tensor(torch.from_numpy(np.stack([r, c])), torch.from_numpy(v), (l.nrows, l.ncols)).coalesce())
    torch.manual_seed(0)
    enc = ref_models.GraphSage(nfeat=shape.feat_dim, nhid=16, orders=[1, 1, 1], dropout=0.0)
    net = ref_models.GNN(encoder=enc, num_classes=shape.num_classes, dropout=0.0, inp=shape.feat_dim)
    net.train()
    x = torch.from_numpy(feats[mb.input_nodes])
    out = net.forward(x, adjs, mb.sampled_nodes)
    labels = torch.nn.functional.one_hot(torch.from_numpy(graphgen.labels(shape, 3)[mb.batch_nodes]), shape.num_classes).float()
    import utils as ref_utils
    loss = ref_utils.loss(out, labels, True, torch.device("cpu"))
    loss.backward()
    state = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    grads = {k: p.grad.detach().numpy() for k, p in net.named_parameters()}
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "model_sage_tiny.npz"), out=out.detach().numpy(), loss=loss.item(),
                        **{"w_" + k: v for k, v in state.items()}, **{"g_" + k: v for k, v in grads.items()})
    print("wrote model_sage_tiny.npz", out.shape, float(loss), list(state)[:6])


if __name__ == "__main__":
    main()
