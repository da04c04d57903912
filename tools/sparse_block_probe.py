#!/usr/bin/env python
"""One cold forward SpMM per width on the layer-0 block of a cached sparse-graph minibatch (papers16 / products), for
`ncu --set full -k regex:spmm_` captures of the short-row kernels.

  python tools/sparse_block_probe.py papers16 1024 256
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import custom_sparse_ops as cso  # noqa: E402

DEV = torch.device("cuda")
shape = sys.argv[1] if len(sys.argv) > 1 else "papers16"
widths = [int(w) for w in sys.argv[2:]] or [1024, 256]
z = np.load(os.path.join(REPO, ".cache", f"mb_{shape}_0.npz"))
flush = torch.empty(384 << 20, dtype=torch.uint8, device=DEV)
flush_src = torch.zeros(96 << 20, dtype=torch.int32, device=DEV)
for li in (0,):
    M, K = [int(v) for v in z[f"l{li}_shape"]]
    adj = cso.adjacency_of(cso.create_coo_tensor(torch.from_numpy(z[f"l{li}_fullrowptr"]).to(DEV), torch.from_numpy(z[f"l{li}_rowptr"]).to(DEV),
                                                 torch.from_numpy(z[f"l{li}_colidx"]).to(DEV), torch.from_numpy(z[f"l{li}_normfact"]).to(DEV), M, K))
    for D in widths:
        X = torch.randn(K, D, device=DEV)
        for rep in range(3):
            flush.zero_()
            flush_src.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            Y = adj.matmul(X)
            e1.record()
            torch.cuda.synchronize()
        print(f"{shape} L{li} {M}x{K} nnz {adj.nnz} D {D}: {e0.elapsed_time(e1) * 1e3:.1f} us", flush=True)
