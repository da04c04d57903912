#!/usr/bin/env python
"""Round-2 kernel A/B on cached LADIES minibatches (.cache/mb_<shape>_0.npz, tools/cache_minibatch.py), experiment build
(-DGNN_TUNE: kernel choice and plan parameters from GNN_TUNE_* environment variables).

For every layer block and width: forward A.X with the row-split kernel and with flat-kernel variants; the A^T index
build; backward A^T.G through the index (row-split / flat) and transpose-free (scatter reductions).  CUDA events, L2
flushed before every launch, median of 5.  Output: a markdown table per shape.

  sh tools/build_tune.sh && python tools/sweep_r2.py reddit papers16 products cora > gpurun_out/sweep_r2.md
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import custom_sparse_ops as cso  # noqa: E402  (production build: adjacency construction only)

lib = ctypes.CDLL(os.path.join(REPO, "tools/_build/libgnn_b200_tune.so"))
vp, i64, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
lib.gnn_csr_spmm_counter_bytes.restype = sz
lib.gnn_csr_spmm_counter_bytes.argtypes = [i64, i64, i64]
lib.gnn_csr_spmm_partial_bytes.restype = sz
lib.gnn_csr_spmm_partial_bytes.argtypes = [i64, i64, i64]
lib.gnn_csr_spmm_f32_ex.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp, vp, sz, ctypes.c_uint, vp]
lib.gnn_csr_spmm_t_f32.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, vp, i64, vp, i64, vp]
lib.gnn_csr_transpose_workspace_bytes.restype = sz
lib.gnn_csr_transpose_workspace_bytes.argtypes = [i64, i64, i64]
lib.gnn_csr_transpose.argtypes = [vp, vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, sz, vp]

DEV = torch.device("cuda")
FLUSH = torch.empty(384 << 20, dtype=torch.uint8, device=DEV)
FLUSH_SRC = torch.zeros(96 << 20, dtype=torch.int32, device=DEV)
TUNE_KEYS = ["GNN_TUNE_FLAT", "GNN_TUNE_FC", "GNN_TUNE_FNV", "GNN_TUNE_FU", "GNN_TUNE_SC", "GNN_TUNE_SNV", "GNN_TUNE_C",
             "GNN_TUNE_NV", "GNN_TUNE_U", "GNN_TUNE_MINB"]


def P(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


USE_ROWIDS = os.environ.get("SWEEP_ROWIDS", "1") == "1"


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, reps=5):
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
    return float(np.median(ts)) * 1e3           # us


def setenv(**kw):
    for k in TUNE_KEYS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        os.environ["GNN_TUNE_" + k] = str(v)


class Csr:
    def __init__(self, rowptr, col, vals, M, K, rowidx=None):
        self.rowptr, self.col, self.vals, self.M, self.K = rowptr, col, vals, M, K
        self.rowidx = rowidx if USE_ROWIDS else None
        self.nnz = int(vals.numel())


def spmm(a: Csr, X, Y, zeroed_ws=None):
    D = X.shape[1]
    cb, pb = lib.gnn_csr_spmm_counter_bytes(a.M, a.nnz, D), lib.gnn_csr_spmm_partial_bytes(a.M, a.nnz, D)
    if zeroed_ws is None and os.environ.get("SWEEP_MEMSET", "0") != "1":
        zeroed_ws = torch.zeros(cb, dtype=torch.uint8, device=DEV)       # what spmm_ext.cpp keeps per stream: no memset launch
    if zeroed_ws is None:
        counters = torch.empty(cb, dtype=torch.uint8, device=DEV)
        flags = 0
    else:
        counters, flags = zeroed_ws, 1
        assert counters.numel() >= cb
    partials = torch.empty(pb, dtype=torch.uint8, device=DEV)

    def run():
        rc = lib.gnn_csr_spmm_f32_ex(P(a.rowptr), P(a.rowidx), P(a.col), P(a.vals), a.M, a.K, a.nnz, D, P(X), X.stride(0), P(Y), D,
                                     P(counters), P(partials), pb, flags, stream())
        assert rc == 0, rc
    return run


def transpose(a: Csr):
    wsb = lib.gnn_csr_transpose_workspace_bytes(a.M, a.K, a.nnz)
    ws = torch.empty(wsb, dtype=torch.uint8, device=DEV)
    t_rowptr = torch.empty(a.K + 1, dtype=torch.int32, device=DEV)
    t_col = torch.empty(a.nnz, dtype=torch.int32, device=DEV)
    t_vals = torch.empty(a.nnz, dtype=torch.float32, device=DEV)
    t_rowidx = torch.empty(a.nnz, dtype=torch.int32, device=DEV)

    def run():
        rc = lib.gnn_csr_transpose(P(a.rowptr), P(a.col), P(a.vals), a.M, a.K, a.nnz, P(t_rowptr), P(t_col), P(t_vals), P(t_rowidx),
                                   P(ws), wsb, stream())
        assert rc == 0, rc
    return run, Csr(t_rowptr, t_col, t_vals, a.K, a.M, t_rowidx)


def scatter(a: Csr, G, dX):
    D = G.shape[1]

    def run():
        rc = lib.gnn_csr_spmm_t_f32(P(a.rowptr), P(a.rowidx), P(a.col), P(a.vals), a.M, a.K, a.nnz, D, P(G), G.stride(0), P(dX), D, stream())
        assert rc == 0, rc
    return run


FLAT_VARIANTS = [("flat C32 nv1 u16", dict(FLAT=1, FC=32, FNV=1, FU=16)), ("flat C64 nv1 u16", dict(FLAT=1, FC=64, FNV=1, FU=16)),
                 ("flat C128 nv1 u16", dict(FLAT=1, FC=128, FNV=1, FU=16)), ("flat C32 nv1 u8", dict(FLAT=1, FC=32, FNV=1, FU=8)),
                 ("flat C64 nv1 u8", dict(FLAT=1, FC=64, FNV=1, FU=8)), ("flat C128 nv1 u8", dict(FLAT=1, FC=128, FNV=1, FU=8)),
                 ("flat C32 nv2 u8", dict(FLAT=1, FC=32, FNV=2, FU=8)), ("flat C64 nv2 u8", dict(FLAT=1, FC=64, FNV=2, FU=8)),
                 ("flat C128 nv2 u8", dict(FLAT=1, FC=128, FNV=2, FU=8)), ("flat C32 nv4 u4", dict(FLAT=1, FC=32, FNV=4, FU=4)),
                 ("flat C64 nv4 u4", dict(FLAT=1, FC=64, FNV=4, FU=4)), ("flat C128 nv4 u4", dict(FLAT=1, FC=128, FNV=4, FU=4))]
SCATTER_VARIANTS = [("scatter C32 nv1", dict(SC=32, SNV=1)), ("scatter C64 nv1", dict(SC=64, SNV=1)),
                    ("scatter C128 nv1", dict(SC=128, SNV=1)), ("scatter C64 nv2", dict(SC=64, SNV=2)),
                    ("scatter C128 nv2", dict(SC=128, SNV=2))]


def best_of(a, X, Y, ref, variants, skip_wide_for_narrow=True):
    """-> (rowsplit_us, best flat name, best flat us, default-plan us)"""
    res = {}
    D = X.shape[1]
    for name, env in [("rowsplit", dict(FLAT=0))] + variants + [("default", {})]:
        if env.get("FNV", 1) > 1 and D < 128 * env["FNV"]:
            continue
        if a.nnz > (4 << 20) and env.get("FLAT") == 1 and env.get("FC", 128) < 128:
            continue
        setenv(**env)
        run = spmm(a, X, Y)
        Y.fill_(float("nan"))
        try:
            run()
        except AssertionError as exc:
            print(f"<!-- {name} on {a.M}x{a.K} nnz {a.nnz} D {D}: rc {exc} -->", flush=True)
            res[name] = (float("inf"), False)
            continue
        torch.cuda.synchronize()
        ok = bool(torch.allclose(Y, ref, rtol=2e-4, atol=2e-5)) if ref is not None else True
        if ref is None:
            ref = Y.clone()
        res[name] = (timed(run), ok)
    setenv()
    return res, ref


def main():
    shapes = sys.argv[1:] or ["reddit", "papers16", "products", "cora"]
    widths = {"reddit": {0: [602], 1: [1024], 2: [1024]}, "products": {0: [100], 1: [512], 2: [512]},
              "cora": {0: [1433], 1: [512]},
              "papers16": {li: [16, 64, 128, 256, 512, 1024] for li in range(3)}}
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(REPO, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    for shape in shapes:
        path = os.path.join(REPO, ".cache", f"mb_{shape}_0.npz")
        if not os.path.exists(path):
            print(f"\n## {shape}: {path} missing (tools/cache_minibatch.py)\n")
            continue
        z = np.load(path)
        print(f"\n## {shape}-shaped LADIES minibatch (HBM roof {hbm} GB/s measured)\n")
        print("| block | D | op | row-split us | best flat/scatter | us | default plan us | HBM roof us | default % of HBM roof | all variants (us) |")
        print("|---|---|---|---|---|---|---|---|---|---|")
        li = 0
        while f"l{li}_shape" in z.files:
            M, K = [int(v) for v in z[f"l{li}_shape"]]
            cols = z[f"l{li}_colidx"]
            adj = cso.adjacency_of(cso.create_coo_tensor(torch.from_numpy(z[f"l{li}_fullrowptr"]).to(DEV), torch.from_numpy(z[f"l{li}_rowptr"]).to(DEV),
                                                         torch.from_numpy(cols).to(DEV), torch.from_numpy(z[f"l{li}_normfact"]).to(DEV), M, K))
            a = Csr(adj.rowptr, adj.colidx, adj.vals, M, K, adj.rowidx)
            rl = np.diff(z[f"l{li}_rowptr"])
            tag = f"L{li} {M}x{K} nnz {a.nnz} row {rl.mean():.1f}/{rl.max()}"
            setenv()
            tr_run, at = transpose(a)
            tr_run()
            torch.cuda.synchronize()
            t_build = timed(tr_run)
            for D in widths.get(shape, {}).get(li, [128]):
                ld = (D + 31) // 32 * 32
                X = torch.randn(K, ld, device=DEV)[:, :D]
                G = torch.randn(M, D, device=DEV)
                Y = torch.empty(M, D, device=DEV)
                dX = torch.empty(K, D, device=DEV)
                B = 8 * a.nnz + 4 * (M + 1) + 4 * K * D + 4 * M * D
                roof = B / (hbm * 1e9) * 1e6
                res, _ = best_of(a, X, Y, None, FLAT_VARIANTS)
                flat = {k: v for k, v in res.items() if k.startswith("flat")}
                bname = min(flat, key=lambda k: flat[k][0]) if flat else "-"
                allv = " ".join(f"{k.replace('flat ', '')}={v[0]:.1f}{'' if v[1] else '(BAD)'}" for k, v in res.items())
                print(f"| {tag} | {D} | fwd | {res['rowsplit'][0]:.1f} | {bname} | {flat[bname][0] if flat else 0:.1f} | {res['default'][0]:.1f} | "
                      f"{roof:.1f} | {100 * roof / res['default'][0]:.1f} | {allv} |", flush=True)
                # zeroed-counter path (no memset launch) with the default plan
                cb = lib.gnn_csr_spmm_counter_bytes(M, a.nnz, D)
                zc = torch.zeros(cb, dtype=torch.uint8, device=DEV)
                t_nomemset = timed(spmm(a, X, Y, zeroed_ws=zc))
                assert int(zc.count_nonzero()) == 0
                # backward through the index
                resb, refb = best_of(at, G, dX, None, FLAT_VARIANTS)
                flatb = {k: v for k, v in resb.items() if k.startswith("flat")}
                bb = min(flatb, key=lambda k: flatb[k][0]) if flatb else "-"
                allb = " ".join(f"{k.replace('flat ', '')}={v[0]:.1f}{'' if v[1] else '(BAD)'}" for k, v in resb.items())
                print(f"| {tag} | {D} | bwd via A^T (build {t_build:.1f} us extra) | {resb['rowsplit'][0]:.1f} | {bb} | {flatb[bb][0] if flatb else 0:.1f} | "
                      f"{resb['default'][0]:.1f} | {roof:.1f} | {100 * roof / resb['default'][0]:.1f} | {allb}; fwd default without memset {t_nomemset:.1f} |", flush=True)
                # transpose-free
                sres = {}
                for name, env in SCATTER_VARIANTS + [("scatter default", {})]:
                    if env.get("SNV", 1) > 1 and D < 256:
                        continue
                    setenv(**env)
                    run = scatter(a, G, dX)
                    run()
                    torch.cuda.synchronize()
                    ok = bool(torch.allclose(dX, refb, rtol=2e-4, atol=2e-5))
                    sres[name] = (timed(run), ok)
                setenv()
                sb = min(sres, key=lambda k: sres[k][0])
                alls = " ".join(f"{k.replace('scatter ', '')}={v[0]:.1f}{'' if v[1] else '(BAD)'}" for k, v in sres.items())
                print(f"| {tag} | {D} | bwd transpose-free (incl. zero fill) | - | {sb} | {sres[sb][0]:.1f} | {sres['scatter default'][0]:.1f} | {roof:.1f} | "
                      f"{100 * roof / sres['scatter default'][0]:.1f} | {alls}; index path total {t_build + resb['default'][0]:.1f} |", flush=True)
            li += 1


if __name__ == "__main__":
    main()
