"""Hub-cached SpMM prototype (experiment build): how much does serving the most frequent columns' X rows from shared
memory buy over the L2-bound production kernel?  Checks the result against the production kernel."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import oracle
lib = ctypes.CDLL('tools/_build/libgnn_b200_tune.so')
vp, i64, sz, ci = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t, ctypes.c_int
lib.gnn_csr_spmm_workspace_bytes.restype = sz
lib.gnn_csr_spmm_workspace_bytes.argtypes = [i64, i64, i64]
lib.gnn_csr_spmm_f32.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, sz, vp]
lib.gnn_debug_spmm_hub.argtypes = [vp, vp, vp, ci, ci, ci, vp, ci, vp, ci, vp, vp, ci, ci, ci, ci, vp]
z = np.load('.cache/mb_reddit_0.npz')
flush = torch.empty(384 << 20, dtype=torch.uint8, device='cuda'); fsrc = torch.zeros(96 << 20, dtype=torch.int32, device='cuda')
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def timed(fn, reps=5):
    ts = []
    for r in range(reps + 2):
        flush.zero_(); fsrc.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3

for li, D in [(0, 602), (1, 1024)]:
    M, K = [int(v) for v in z[f'l{li}_shape']]
    rows, cols, vals = oracle.build_adj(z[f'l{li}_fullrowptr'], z[f'l{li}_rowptr'], z[f'l{li}_colidx'], z[f'l{li}_normfact'], M)
    rowptr = torch.from_numpy(z[f'l{li}_rowptr']).cuda(); col = torch.from_numpy(cols.astype(np.int32)).cuda(); v = torch.from_numpy(vals).cuda()
    nnz = len(vals); ld = (D + 31) // 32 * 32
    X = torch.randn(K, ld, device='cuda'); Y = torch.empty(M, D, device='cuda'); Y2 = torch.zeros(M, D, device='cuda')
    wsb = lib.gnn_csr_spmm_workspace_bytes(M, nnz, D); ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    t_prod = timed(lambda: lib.gnn_csr_spmm_f32(P(rowptr), P(col), P(v), M, K, nnz, D, P(X), ld, P(Y), D, P(ws), wsb, st))
    cnt = np.bincount(cols, minlength=K)
    order = np.argsort(-cnt, kind='stable')
    print(f"layer{li} {M}x{K} nnz={nnz} D={D}: production kernel {t_prod:.1f} us")
    for H in (0, 512, 1024):
        hubcols = order[:max(H, 1)].astype(np.int32)
        slot = np.full(K, -1, np.int16)
        if H: slot[hubcols] = np.arange(H, dtype=np.int16)
        cover = cnt[hubcols].sum() / nnz if H else 0.0
        d_slot = torch.from_numpy(slot).cuda(); d_hub = torch.from_numpy(hubcols).cuda()
        for ranges in (8, 16, 31):
            for unr in (2, 4, 8):
                fn = lambda: lib.gnn_debug_spmm_hub(P(rowptr), P(col), P(v), M, nnz, D, P(X), ld, P(Y2), D, P(d_slot), P(d_hub), H, K, ranges, unr, st)
                rc = fn(); torch.cuda.synchronize()
                assert rc == 0, rc
                ok = torch.allclose(Y2, Y, rtol=1e-4, atol=1e-4)
                t = timed(fn, reps=3)
                print(f"  H={H} (covers {cover*100:.0f}% of nnz) ranges={ranges} unr={unr}: {t:.1f} us  ok={ok}")
