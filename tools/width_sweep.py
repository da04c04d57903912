#!/usr/bin/env python
"""SpMM width sweep (BASELINE.json configs[4]: D in 16..1024) on LADIES blocks of a synthetic graph.

  python tools/width_sweep.py --workload papers16 --orders 1,1,1 > profiles/r1_width_sweep_papers16.md

For every sampled layer block and every width: forward A.X and backward A^T.G (A^T index cached, so the
backward number is the product alone; the one-off A^T build is listed separately), CUDA events, L2 flushed
before every launch, median of 5; algorithmic bytes = SURVEY.md 8(d) formula; roof = MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import custom_sparse_ops as cso  # noqa: E402
from gnn_b200 import graphgen, sampler  # noqa: E402


_FLUSH_SRC = None


def timed(fn, flush, reps=5):
    global _FLUSH_SRC
    if _FLUSH_SRC is None:
        _FLUSH_SRC = torch.zeros(96 << 20, dtype=torch.int32, device=flush.device)
    ts = []
    for r in range(reps + 2):
        flush.zero_()
        _FLUSH_SRC.sum()          # leave the cache full of clean lines (no write-back inside the timed region)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="papers16")
    ap.add_argument("--orders", default="1,1,1")
    ap.add_argument("--samp", type=int, default=8192)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--widths", default="16,32,64,128,256,512,1024")
    args = ap.parse_args()
    orders = [int(x) for x in args.orders.split(",")]
    shape = graphgen.SHAPES[args.workload]
    g = graphgen.generate_cached(shape, seed=0)
    rng = np.random.Generator(np.random.PCG64(1000))
    mb = sampler.ladies_sample(1234, g.train_nodes[rng.permutation(g.train_nodes.size)[:args.batch]], [args.samp] * 5, g.num_nodes,
                               g.indptr, g.indices, orders)
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    dev = torch.device("cuda")
    flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev)
    print(f"# SpMM width sweep on {shape.name}-shaped LADIES blocks (samp_num {args.samp}, batch {args.batch}, orders {orders})\n")
    print(f"Graph: {g.num_nodes} nodes, {g.nnz} directed nnz, max degree {int(g.degrees().max())}; HBM roof {hbm} GB/s "
          f"({'measured' if peaks else 'fallback'}). Times are CUDA-event medians with L2 flushed before each launch.\n")
    print("| block (MxK, nnz, mean/max row) | D | fwd us | fwd GB/s | fwd % HBM | bwd us | bwd GB/s | bwd % HBM | A^T build us |")
    print("|---|---|---|---|---|---|---|---|---|")
    for li, layer in enumerate(mb.layers):
        if layer is None:
            continue
        a = cso.create_coo_tensor(torch.from_numpy(layer.fullrowptr).to(dev), torch.from_numpy(layer.rowptr).to(dev),
                                  torch.from_numpy(layer.colidx32).to(dev), torch.from_numpy(layer.normfact).to(dev),
                                  layer.nrows, layer.ncols)
        adj = cso.adjacency_of(a)
        rl = np.diff(layer.rowptr)

        def build():
            adj._t = None
            adj.transpose()
        t_build = timed(build, flush)
        for D in [int(x) for x in args.widths.split(",")]:
            x = torch.randn(layer.ncols, D, device=dev)
            go = torch.randn(layer.nrows, D, device=dev)
            B = bench.algorithmic_bytes(layer.nnz, layer.nrows, layer.ncols, D)
            tf = timed(lambda: adj.matmul(x), flush)
            tb = timed(lambda: adj.matmul_t(go), flush)
            print(f"| L{li} {layer.nrows}x{layer.ncols}, {layer.nnz}, {rl.mean():.1f}/{rl.max()} | {D} | {tf * 1e6:.1f} | {B / tf / 1e9:.0f} | "
                  f"{100 * B / tf / 1e9 / hbm:.1f} | {tb * 1e6:.1f} | {B / tb / 1e9:.0f} | {100 * B / tb / 1e9 / hbm:.1f} | {t_build * 1e6:.1f} |")


if __name__ == "__main__":
    main()
