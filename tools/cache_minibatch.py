#!/usr/bin/env python
"""Sample ONE LADIES minibatch of a synthetic shape on the CPU and store its hand-off arrays in .cache/mb_<shape>_<i>.npz
(a few MB), so that GPU-side tools (tools/sweep_r2.py, tools/gather_roof.py) need not regenerate 100 M-edge graphs.

  python tools/cache_minibatch.py reddit products papers16
"""
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from gnn_b200 import graphgen, sampler  # noqa: E402


def main():
    for name in sys.argv[1:]:
        out = os.path.join(REPO, ".cache", f"mb_{name}_0.npz")
        if os.path.exists(out):
            print(f"{out} exists")
            continue
        shape = graphgen.SHAPES[name]
        t0 = time.time()
        g = graphgen.generate_cached(shape, seed=0, root=os.path.join(REPO, ".cache"))
        print(f"{name}: graph {g.num_nodes} nodes {g.nnz} nnz in {time.time() - t0:.0f}s", flush=True)
        samp, batch = (512, 256) if name == "cora" else (8192, 512)
        rng = np.random.Generator(np.random.PCG64(1000))
        nodes = g.train_nodes[rng.permutation(g.train_nodes.size)[:batch]]
        orders = [1, 1] if name == "cora" else [1, 1, 1]
        t0 = time.time()
        mb = sampler.ladies_sample(1234, nodes, [samp] * 5, g.num_nodes, g.indptr, g.indices, orders)
        arrs = {"input_nodes": mb.input_nodes}
        for li, l in enumerate(mb.layers):
            arrs.update({f"l{li}_fullrowptr": l.fullrowptr, f"l{li}_rowptr": l.rowptr, f"l{li}_colidx": l.colidx32,
                         f"l{li}_normfact": l.normfact, f"l{li}_shape": np.array([l.nrows, l.ncols], dtype=np.int64)})
        np.savez_compressed(out, **arrs)
        print(f"{name}: sampled in {time.time() - t0:.1f}s: " + "; ".join(f"{l.nrows}x{l.ncols} nnz {l.nnz}" for l in mb.layers), flush=True)


if __name__ == "__main__":
    main()
