"""Speed of light of the L2->SM row gather the SpMM performs (experiment build)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, '.')
lib = ctypes.CDLL('tools/_build/libgnn_b200_tune.so')
z = np.load('.cache/mb_reddit_0.npz')
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
sink = torch.zeros(4, device='cuda')
flush = torch.empty(384 << 20, dtype=torch.uint8, device='cuda')
for li, D in [(1, 1024), (0, 1024)]:
    M, K = [int(v) for v in z[f'l{li}_shape']]
    col = torch.from_numpy(z[f'l{li}_colidx'].astype(np.int32)).cuda()
    nnz = col.numel()
    X = torch.randn(K, D, device='cuda')
    print(f"layer{li} K={K} nnz={nnz} D={D}: X is {K*D*4/1e6:.0f} MB")
    for nv, u in [(8, 1), (8, 2), (4, 2), (4, 4), (2, 4), (2, 8), (1, 8), (1, 16)]:
        for warps_per_sm in (16, 32, 64):
            warps = 148 * warps_per_sm
            slabs = D // (128 * nv)
            per_warp = max(32, (nnz * slabs // warps) // 32 * 32)
            bytes_ = warps * per_warp * nv * 512
            ts = []
            for r in range(4):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = lib.gnn_debug_gather_roof(P(X), D, K, P(col), nnz, nv, u, warps, per_warp, P(sink), st)
                e1.record(); torch.cuda.synchronize()
                assert rc == 0
                if r: ts.append(e0.elapsed_time(e1))
            t = float(np.median(ts)) * 1e-3
            print(f"  nv={nv} u={u} warps/SM={warps_per_sm}: {bytes_/1e9:.2f} GB in {t*1e6:.0f} us = {bytes_/t/1e12:.2f} TB/s")
