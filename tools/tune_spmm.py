"""Sweep (NV, U, MINB, C) of the row-split SpMM on the cached Reddit-shaped blocks (experiment build, -DGNN_TUNE)."""
import ctypes, os, sys, itertools, json
import numpy as np, torch
sys.path.insert(0, '.')
import oracle
lib = ctypes.CDLL('tools/_build/libgnn_b200_tune.so')
vp, i64, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
lib.gnn_csr_spmm_workspace_bytes.restype = sz
lib.gnn_csr_spmm_workspace_bytes.argtypes = [i64, i64, i64]
lib.gnn_csr_spmm_f32.argtypes = [vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, sz, vp]
z = np.load('.cache/mb_reddit_0.npz')
flush = torch.empty(384 << 20, dtype=torch.uint8, device='cuda')
def P(t): return ctypes.c_void_p(t.data_ptr())
blocks = []
for li, D in [(0, 602), (1, 1024), (2, 1024)]:
    M, K = [int(v) for v in z[f'l{li}_shape']]
    rows, cols, vals = oracle.build_adj(z[f'l{li}_fullrowptr'], z[f'l{li}_rowptr'], z[f'l{li}_colidx'], z[f'l{li}_normfact'], M)
    ld = (D + 31) // 32 * 32
    X = torch.randn(K, ld, device='cuda')
    blocks.append(dict(li=li, M=M, K=K, D=D, ld=ld, nnz=len(vals), rowptr=torch.from_numpy(z[f'l{li}_rowptr']).cuda(),
                       col=torch.from_numpy(cols.astype(np.int32)).cuda(), vals=torch.from_numpy(vals).cuda(), X=X,
                       Y=torch.empty(M, D, device='cuda')))
def run(b, reps=5):
    wsb = lib.gnn_csr_spmm_workspace_bytes(b['M'], b['nnz'], b['D'])
    ws = torch.empty(wsb, dtype=torch.uint8, device='cuda')
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ts = []
    for r in range(reps + 2):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.gnn_csr_spmm_f32(P(b['rowptr']), P(b['col']), P(b['vals']), b['M'], b['K'], b['nnz'], b['D'], P(b['X']), b['ld'],
                                  P(b['Y']), b['D'], P(ws), wsb, st)
        e1.record()
        assert rc == 0, rc
        torch.cuda.synchronize()
        if r >= 2: ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
ref = {}
for b in blocks:
    for k in ['GNN_TUNE_U', 'GNN_TUNE_MINB', 'GNN_TUNE_NV', 'GNN_TUNE_C']: os.environ.pop(k, None)
    t = run(b); ref[b['li']] = b['Y'].clone()
    print(f"layer{b['li']} D={b['D']} default: {t*1e3:.1f} us", flush=True)
res = []
for b in blocks:
    cs = [0] if b['li'] < 2 else [0, 64]            # 0 = leave the chunk to the wave-fitting planner
    for nv, u, minb, c in itertools.product([1, 2, 3, 4, 5], [1, 2, 4, 8], [2, 3, 4], cs):
        if nv * u > 16 or nv * u < 4: continue
        os.environ.update(GNN_TUNE_NV=str(nv), GNN_TUNE_U=str(u), GNN_TUNE_MINB=str(minb))
        if c: os.environ['GNN_TUNE_C'] = str(c)
        else: os.environ.pop('GNN_TUNE_C', None)
        t = run(b, reps=3)
        ok = torch.allclose(b['Y'], ref[b['li']], rtol=1e-4, atol=1e-5)
        res.append((b['li'], nv, u, minb, c, t, ok))
for li in (0, 1, 2):
    rs = sorted([r for r in res if r[0] == li], key=lambda r: r[5])
    print(f"--- layer {li}: best 14 of {len(rs)}; all ok = {all(r[6] for r in rs)}")
    for r in rs[:14]: print(f"  nv={r[1]} u={r[2]} minb={r[3]} C={r[4]}: {r[5]*1e3:.1f} us ok={r[6]}")
    for nv in [1, 2, 3, 4, 5, 6, 8]:
        rr = [r for r in rs if r[1] == nv]
        if rr: print(f"  best nv={nv}: u={rr[0][2]} minb={rr[0][3]} C={rr[0][4]}: {rr[0][5]*1e3:.1f} us")
