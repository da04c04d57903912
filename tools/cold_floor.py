#!/usr/bin/env python
"""What does the timing protocol itself cost?  Under the bench's protocol (L2 flushed, CUDA events around ONE launch)
time (a) an empty kernel-sized op, (b) a plain copy moving B bytes (B/2 read + B/2 written) for the algorithmic byte
counts of the small LADIES blocks, (c) a random row gather (gnn_gather_rows_f32) of the same rows an SpMM would read
once.  These are the floors any SpMM on those blocks sits on in this measurement: HBM-roof microseconds are not
reachable when a cold 2 MB copy already takes ~10 us.

  python tools/cold_floor.py > gpurun_out/cold_floor.md
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import custom_sparse_ops as cso  # noqa: E402

DEV = torch.device("cuda")
FLUSH = torch.empty(384 << 20, dtype=torch.uint8, device=DEV)
FLUSH_SRC = torch.zeros(96 << 20, dtype=torch.int32, device=DEV)
ext = cso.spmm_cpp


def timed(fn, reps=7):
    ts = []
    for r in range(reps + 2):
        FLUSH.zero_()
        FLUSH_SRC.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


one = torch.zeros(32, device=DEV)
print("# Cold-launch floors under the bench timing protocol (L2 flushed, CUDA events around one launch), B200\n")
print(f"empty-ish kernel (32-float fill): {timed(lambda: one.fill_(1.0)):.1f} us\n")
print("| bytes moved (MB) | HBM roof us @6553.6 GB/s | cold copy us | copy % of roof | cold random row gather us (rows x 4 KB) |")
print("|---|---|---|---|---|")
for mb in (1, 2, 4, 8, 16, 38, 64, 100, 168, 256):
    n = mb * (1 << 20) // 8            # floats per half
    src = torch.randn(n, device=DEV)
    dst = torch.empty(n, device=DEV)
    t = timed(lambda: dst.copy_(src))
    rows = max(1, mb * (1 << 20) // 2 // 4096)
    table = torch.randn(max(rows * 2, 1024), 1024, device=DEV)
    idx = torch.randint(0, table.shape[0], (rows,), device=DEV)
    ptrs = table.data_ptr() + idx * 4096
    tg = timed(lambda: ext.gather_rows(ptrs, 1024, 1024))
    roof = mb * (1 << 20) / 6553.6e9 * 1e6
    print(f"| {mb} | {roof:.1f} | {t:.1f} | {100 * roof / t:.0f} | {tg:.1f} ({rows} rows) |")
