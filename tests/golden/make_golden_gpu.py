"""Generate tests/golden/ref_gpu_*.npz by running the reference's OWN CUDA extension (compiled unmodified into
oracle/_ref/spmm_ref.so by oracle/build_ref.py) on a B200:

    gpurun -- 'python tests/golden/make_golden_gpu.py'      # writes gpurun_out/ref_gpu_*.npz; copy them to tests/golden/

Captured per case: the inputs handed to create_coo_tensor, its indices/values, and for a dense operand X / G the
outputs of spmm_naive (deterministic) and spmm_load_balance forward, and of the reference backward expression
spmm_*(mat1.transpose(0,1).coalesce(), G) (custom_sparse_ops.py:34).  The CPU suite then checks the oracle against
these files (tests/test_oracle_golden.py), so the pin does not depend on the compiled reference being present."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from gnn_b200 import graphgen, sampler  # noqa: E402
from oracle import build_ref  # noqa: E402


def main():
    ref = build_ref.load_ref()
    assert ref is not None, "oracle/_ref/spmm_ref.so missing: run oracle/build_ref.py in the build container first"
    out_dir = os.path.join(REPO, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    cases = [("tiny", [1, 1, 1], 96, 32, 7, 37), ("cora", [1, 1], 512, 256, 1234, 40), ("small", [1, 1], 1024, 128, 99, 24)]
    for shape_name, orders, samp, batch, seed, D in cases:
        shape = graphgen.SHAPES[shape_name]
        g = graphgen.generate(shape, seed=0)
        mb = sampler.ladies_sample(seed, g.train_nodes[:batch], [samp] * 5, shape.num_nodes, g.indptr, g.indices, orders)
        out = {"shape": shape_name, "orders": np.array(orders), "samp": samp, "batch": batch, "seed": seed, "nlayers": len(mb.layers)}
        rng = np.random.Generator(np.random.PCG64(seed))
        for li, l in enumerate(mb.layers):
            a = ref.create_coo_tensor(torch.from_numpy(l.fullrowptr).cuda(), torch.from_numpy(l.rowptr).cuda(),
                                      torch.from_numpy(l.colidx).cuda(), torch.from_numpy(l.normfact).cuda(), l.nrows, l.ncols)
            X = rng.standard_normal((l.ncols, D)).astype(np.float32)
            G = rng.standard_normal((l.nrows, D)).astype(np.float32)
            dX, dG = torch.from_numpy(X).cuda(), torch.from_numpy(G).cuda()
            at = a.transpose(0, 1).coalesce()
            pre = f"l{li}_"
            out.update({pre + "fullrowptr": l.fullrowptr, pre + "rowptr": l.rowptr, pre + "colidx": l.colidx, pre + "normfact": l.normfact,
                        pre + "shape": np.array([l.nrows, l.ncols]),
                        pre + "indices": a._indices().cpu().numpy(), pre + "values": a._values().cpu().numpy(),
                        pre + "X": X, pre + "G": G,
                        pre + "y_naive": ref.spmm_naive(a, dX).cpu().numpy(),
                        pre + "y_load_balance": ref.spmm_load_balance(a, dX).cpu().numpy(),
                        pre + "dx_naive": ref.spmm_naive(at, dG.contiguous()).cpu().numpy(),
                        pre + "dx_load_balance": ref.spmm_load_balance(at, dG.contiguous()).cpu().numpy(),
                        pre + "t_indices": at._indices().cpu().numpy()})
        path = os.path.join(out_dir, f"ref_gpu_{shape_name}.npz")
        np.savez_compressed(path, **out)
        print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
