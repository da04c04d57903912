"""Bring-up and timing probe of the tcgen05 3xTF32 linears (csrc/linear_tc.cuh) against fp64 and against cuBLAS fp32.

    python tools/linear_tc_probe.py nt|tn|time|model
Each group runs in its own process (a trapped kernel kills the CUDA context)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gnn_b200 import custom_sparse_ops as cso  # noqa: E402

ext = cso.spmm_cpp
dev = torch.device("cuda")


def rel(a, b):
    return ((a.double() - b).norm() / b.norm().clamp_min(1e-30)).item()


def errmap(got, ref, rb=32, cb=32, maxr=8, maxc=8):
    d = (got.double() - ref).abs()
    scale = ref.abs().mean().item() + 1e-30
    rows = []
    for i in range(min(maxr, (d.shape[0] + rb - 1) // rb)):
        rows.append(" ".join(f"{d[i*rb:(i+1)*rb, j*cb:(j+1)*cb].mean().item()/scale:8.1e}"
                             for j in range(min(maxc, (d.shape[1] + cb - 1) // cb))))
    return "\n      ".join(rows)


def run_nt(M, K, N, gather=False, ld_pad=0, bias=True, col_slice=False, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_in = M + 37 if gather else M
    xfull = torch.randn(n_in, K + ld_pad, generator=g).to(dev)
    x = xfull[:, :K]
    W = (torch.randn(N, K, generator=g) * 0.1).to(dev)
    b = torch.randn(N, generator=g).to(dev) if bias else None
    rows = torch.randperm(n_in, generator=g)[:M].to(dev) if gather else None
    w_nk, _ = ext.linear_split_weights(W, False)
    buf = torch.full((M, N + (64 if col_slice else 0)), float("nan"), device=dev)
    out = buf[:, 32:32 + N] if col_slice else buf
    ext.linear_tf32x3(x, rows, w_nk, K, b, out)
    torch.cuda.synchronize()
    xa = x if rows is None else x[rows]
    ref = xa.double() @ W.double().t() + (b.double() if bias else 0)
    e = rel(out, ref)
    e32 = rel(xa @ W.t() + (b if bias else 0), ref)
    ok = e <= 3e-6 and not torch.isnan(out).any().item()
    if col_slice:
        ok = ok and torch.isnan(buf[:, :32]).all().item() and torch.isnan(buf[:, 32 + N:]).all().item()
    print(f"NT M={M} K={K} N={N} gather={gather} pad={ld_pad} slice={col_slice}: rel {e:.2e} (cuBLAS fp32 {e32:.2e}) {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        print("      " + errmap(out, ref), flush=True)
    return ok


def run_tn(M, N, K, gather=False, seed=0, lddy_extra=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    n_in = M + 11 if gather else M
    X = torch.randn(n_in, K, generator=g).to(dev)
    dyfull = torch.randn(M, N + lddy_extra, generator=g).to(dev)
    dY = dyfull[:, lddy_extra:]
    rows = torch.randperm(n_in, generator=g)[:M].to(dev) if gather else None
    dW, _ = ext.linear_wgrad_tf32x3(dY, X, rows)
    torch.cuda.synchronize()
    xa = X if rows is None else X[rows]
    ref = dY.double().t() @ xa.double()
    e = rel(dW, ref)
    e32 = rel(dY.t() @ xa, ref)
    ok = e <= 3e-6
    print(f"TN M={M} N={N} K={K} gather={gather} lddy+{lddy_extra}: rel {e:.2e} (cuBLAS fp32 {e32:.2e}) {'OK' if ok else 'FAIL'}", flush=True)
    if not ok:
        print("      " + errmap(dW, ref), flush=True)
    return ok


def tn_decode():
    """One-hot operands: where does the hardware think element (m, n) of dY and (m, k) of X live?"""
    M, N, K = 8, 128, 256
    for (m0, n0) in [(0, 0), (0, 1), (0, 4), (0, 32), (1, 0), (7, 5), (3, 100)]:
        dY = torch.zeros(M, N, device=dev); dY[m0, n0] = 1.0
        X = torch.zeros(M, K, device=dev); X[m0, :] = torch.arange(1, K + 1, device=dev).float()
        dW, _ = ext.linear_wgrad_tf32x3(dY, X, None)
        torch.cuda.synchronize()
        nz = dW.nonzero()
        head = [(int(a), int(b), float(dW[a, b])) for a, b in nz[:6].tolist()]
        rows = sorted(set(nz[:, 0].tolist()))[:8]
        print(f"dY one-hot (m={m0}, n={n0}): {len(nz)} nonzeros, max {dW.abs().max().item():.3g}, rows {rows}, head {head}", flush=True)
    for (m0, k0) in [(0, 0), (0, 1), (0, 4), (0, 32), (1, 0), (7, 5), (3, 200)]:
        X = torch.zeros(M, K, device=dev); X[m0, k0] = 1.0
        dY = torch.zeros(M, N, device=dev); dY[m0, :] = torch.arange(1, N + 1, device=dev).float()
        dW, _ = ext.linear_wgrad_tf32x3(dY, X, None)
        torch.cuda.synchronize()
        nz = dW.nonzero()
        head = [(int(a), int(b), float(dW[a, b])) for a, b in nz[:6].tolist()]
        cols = sorted(set(nz[:, 1].tolist()))[:8]
        print(f"X one-hot (m={m0}, k={k0}): {len(nz)} nonzeros, max {dW.abs().max().item():.3g}, cols {cols}, head {head}", flush=True)
    dY = torch.ones(M, N, device=dev); X = torch.ones(M, K, device=dev)
    dW, _ = ext.linear_wgrad_tf32x3(dY, X, None)
    print("all-ones: min", dW.min().item(), "max", dW.max().item(), "expected", M, flush=True)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def group_time():
    torch.backends.cuda.matmul.allow_tf32 = False
    print("| op | shape | ours us | TF/s (fp32-equivalent) | cuBLAS fp32 us |\n|---|---|---:|---:|---:|")
    for (M, K, N) in [(16157, 602, 512), (8689, 1024, 512), (512, 1024, 512), (8689, 512, 1024), (16157, 512, 602)]:
        x = torch.randn(M, K, device=dev)
        W = torch.randn(N, K, device=dev) * 0.1
        b = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev)
        w_nk, _ = ext.linear_split_weights(W, False)
        t = timed(lambda: ext.linear_tf32x3(x, None, w_nk, K, b, out))
        t0 = timed(lambda: torch.addmm(b, x, W.t()))
        print(f"| NT | {M}x{K}x{N} | {t:.1f} | {2*M*K*N/t/1e6:.0f} | {t0:.1f} |", flush=True)
        if K == 602:
            xp = torch.randn(M, 608, device=dev)[:, :602]
            t = timed(lambda: ext.linear_tf32x3(xp, None, w_nk, K, b, out))
            print(f"| NT (ld 608) | {M}x{K}x{N} | {t:.1f} | {2*M*K*N/t/1e6:.0f} | |", flush=True)
        ts = timed(lambda: ext.linear_split_weights(W, True))
        print(f"| split W | {N}x{K} | {ts:.1f} | | |", flush=True)
    for (M, N, K) in [(16157, 512, 602), (8689, 512, 1024), (512, 512, 1024)]:
        X = torch.randn(M, (K + 31) // 32 * 32, device=dev)[:, :K]          # rows on 128-byte lines, as the layer feeds them
        dY = torch.randn(M, 2 * N, device=dev)[:, N:]
        t = timed(lambda: ext.linear_wgrad_tf32x3(dY, X, None))
        t0 = timed(lambda: torch.mm(dY.t(), X))
        print(f"| TN | M{M} N{N} K{K} | {t:.1f} | {2*M*K*N/t/1e6:.0f} | {t0:.1f} |", flush=True)
        if K % 4:
            Xu = torch.randn(M, K, device=dev)
            t = timed(lambda: ext.linear_wgrad_tf32x3(dY, Xu, None))
            print(f"| TN (rows {K} floats apart: scalar loads) | M{M} N{N} K{K} | {t:.1f} | {2*M*K*N/t/1e6:.0f} | |", flush=True)
        rows = torch.randperm(M, device=dev)
        t = timed(lambda: ext.linear_wgrad_tf32x3(dY, X, rows))
        print(f"| TN (gathered rows) | M{M} N{N} K{K} | {t:.1f} | {2*M*K*N/t/1e6:.0f} | |", flush=True)


def group_model():
    from gnn_b200 import models
    torch.manual_seed(0)
    M, n_in, K, n = 1000, 1500, 602, 512
    x = torch.randn(n_in, K, device=dev, requires_grad=True)
    agg = torch.randn(M, K, device=dev, requires_grad=True)
    rows = torch.randperm(n_in, device=dev)[:M]
    WB = (torch.randn(n, K, device=dev) * 0.05).requires_grad_(True)
    WW = (torch.randn(n, K, device=dev) * 0.05).requires_grad_(True)
    bB = torch.randn(n, device=dev, requires_grad=True)
    bW = torch.randn(n, device=dev, requires_grad=True)
    ins = [x, agg, WB, bB, WW, bW]
    pre = models.SageLinears.apply(x, rows, agg, WB, bB, WW, bW)
    gout = torch.randn_like(pre)
    got = torch.autograd.grad(pre, ins, gout)
    ins64 = [t.detach().double().requires_grad_(True) for t in ins]
    x6, a6, WB6, bB6, WW6, bW6 = ins64
    pre64 = torch.cat([x6[rows] @ WB6.t() + bB6, a6 @ WW6.t() + bW6], 1)
    ref = torch.autograd.grad(pre64, ins64, gout.double())
    print(f"SageLinears fwd rel {rel(pre, pre64.detach()):.2e}")
    ok = rel(pre, pre64.detach()) <= 3e-6
    for name, a, b in zip(["dx", "dagg", "dWB", "dbB", "dWW", "dbW"], got, ref):
        e = rel(a, b)
        ok = ok and e <= 3e-6
        print(f"  {name}: rel {e:.2e}")
    print("MODEL", "OK" if ok else "FAIL")
    return ok


if __name__ == "__main__":
    grp = sys.argv[1]
    ok = True
    if grp == "nt":
        for args in [dict(M=128, K=32, N=256, bias=False), dict(M=128, K=64, N=256), dict(M=256, K=128, N=512),
                     dict(M=300, K=602, N=512, gather=True), dict(M=1000, K=602, N=512, ld_pad=6, col_slice=True),
                     dict(M=777, K=512, N=602, bias=False), dict(M=129, K=1024, N=512, gather=True, col_slice=True),
                     dict(M=64, K=100, N=47), dict(M=16157, K=602, N=512, gather=True, ld_pad=6)]:
            ok = run_nt(**args) and ok
    elif grp == "tn":
        for args in [dict(M=32, N=128, K=256), dict(M=128, N=128, K=256), dict(M=1024, N=512, K=602),
                     dict(M=5000, N=512, K=602, gather=True, lddy_extra=512), dict(M=8689, N=512, K=1024, lddy_extra=512),
                     dict(M=333, N=100, K=47), dict(M=16157, N=512, K=602, gather=True)]:
            ok = run_tn(**args) and ok
    elif grp == "ncu":      # one launch of each production shape, for `ncu --set full -k regex:linear_tc`
        for (M, K, N) in [(16157, 602, 512), (8689, 1024, 512), (8689, 512, 1024)]:
            x = torch.randn(M, 608 if K == 602 else K, device=dev)[:, :K]
            W = torch.randn(N, K, device=dev) * 0.1
            out = torch.empty(M, N, device=dev)
            w_nk, _ = ext.linear_split_weights(W, False)
            ext.linear_tf32x3(x, None, w_nk, K, None, out)
        for (M, N, K) in [(16157, 512, 602), (8689, 512, 1024)]:
            X = torch.randn(M, (K + 31) // 32 * 32, device=dev)[:, :K]
            dY = torch.randn(M, 2 * N, device=dev)[:, N:]
            ext.linear_wgrad_tf32x3(dY, X, None)
        torch.cuda.synchronize()
    elif grp == "tndbg":
        tn_decode()
    elif grp == "time":
        group_time()
    elif grp == "model":
        ok = group_model()
    print("GROUP", grp, "OK" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)
