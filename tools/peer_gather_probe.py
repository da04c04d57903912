#!/usr/bin/env python
"""Peer-gather over NVLink, ONE process driving two GPUs (so that ncu may profile it: `ncu --metrics
nvlrx__bytes.sum,nvltx__bytes.sum,...`): a feature shard lives on cuda:1, the gather kernels run on cuda:0 and read it
through peer addresses - the same loads the multi-process FeatureStore issues on CUDA-IPC mappings.

  python tools/peer_gather_probe.py [rows] [F]        prints rows, bytes, us, GB/s for the plain gather and for the fused
                                                      gather+SpMM reading peer rows directly (papers-shaped block)
"""
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import custom_sparse_ops as cso  # noqa: E402
from gnn_b200 import gather as gmod  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
F = int(sys.argv[2]) if len(sys.argv) > 2 else 602
ext = cso.spmm_cpp
d0, d1 = torch.device("cuda", 0), torch.device("cuda", 1)
ld = gmod.padded_ld(F)
shard = torch.randn(262144, ld, device=d1)
torch.zeros(1, device=d1).to(d0)
torch.cuda.set_device(0)
import ctypes
rt = ctypes.CDLL("libcudart.so.12")                    # already loaded by torch; kernels on cuda:0 may then dereference cuda:1 memory
rc = rt.cudaDeviceEnablePeerAccess(ctypes.c_int(1), ctypes.c_uint(0))
assert rc in (0, 704), f"cudaDeviceEnablePeerAccess -> {rc}"        # 704 = already enabled
rt.cudaGetLastError()
gen = torch.Generator(device=d0).manual_seed(1)
slots = torch.randint(0, shard.shape[0], (rows,), device=d0, generator=gen)
ptrs = shard.data_ptr() + slots * (ld * 4)
src = torch.zeros(rows, dtype=torch.int32, device=d0)
out = torch.empty(rows, ld, device=d0)
res = {}


def timed(fn, reps=5):
    ts = []
    for r in range(reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


t = timed(lambda: ext.gather_rows_src(ptrs, src, 0, F, out))
ok = bool(torch.equal(out[:, :F].cpu(), shard[slots.to(d1)][:, :F].cpu()))
res["gather"] = {"rows": rows, "bytes": rows * F * 4, "us": round(t, 1), "GBps": round(rows * F * 4 / t / 1e3, 1), "bit_exact": ok,
                 "frac_of_measured_peer_copy_770": round(rows * F * 4 / t / 1e3 / 770, 3)}

# fused gather+SpMM on a papers-shaped block (nnz/K ~ 5): every nonzero reads its X row straight from the peer shard,
# against staging the rows first (gather, then SpMM on the local copy)
z = os.path.join(REPO, ".cache", "mb_papers16_0.npz")
if os.path.exists(z) and F == 128:
    z = np.load(z)
    M, K = [int(v) for v in z["l0_shape"]]
    a = cso.create_coo_tensor(torch.from_numpy(z["l0_fullrowptr"]).to(d0), torch.from_numpy(z["l0_rowptr"]).to(d0),
                              torch.from_numpy(z["l0_colidx"]).to(d0), torch.from_numpy(z["l0_normfact"]).to(d0), M, K)
    adj = cso.adjacency_of(a)
    sl = torch.randint(0, shard.shape[0], (K,), device=d0, generator=gen)
    xr = shard.data_ptr() + sl * (ld * 4)
    s0 = torch.zeros(K, dtype=torch.int32, device=d0)
    stage = torch.empty(K, ld, device=d0)

    def staged():
        ext.gather_rows_src(xr, s0, 0, F, stage)
        return adj.matmul(stage[:, :F])
    y_direct = adj.gather_matmul(xr, F)
    y_staged = staged()
    res["fused_direct_vs_staged"] = {"block": f"{M}x{K} nnz {adj.nnz} D {F}", "direct_us": round(timed(lambda: adj.gather_matmul(xr, F)), 1),
                                     "staged_us": round(timed(staged), 1), "same_bits": bool(torch.equal(y_direct, y_staged)),
                                     "peer_bytes_direct": adj.nnz * F * 4, "peer_bytes_staged": K * F * 4}
print(json.dumps(res))
