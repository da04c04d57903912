"""How fast are the dense linears' GEMM shapes on this GPU in fp32 (SIMT), TF32 (tensor cores) and as 3 TF32 GEMMs?"""
import torch, numpy as np
dev = torch.device("cuda")
def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print("torch", torch.__version__, "fp32_precision attr:", getattr(torch.backends.cuda.matmul, "fp32_precision", None))
for (m, k, n) in [(16157, 602, 512), (8689, 1024, 512), (512, 1024, 512)]:
    a = torch.randn(m, k, device=dev); b = torch.randn(k, n, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    t32 = timed(lambda: torch.mm(a, b)); ref = torch.mm(a.double(), b.double())
    e32 = ((torch.mm(a, b).double() - ref).norm() / ref.norm()).item()
    torch.backends.cuda.matmul.allow_tf32 = True
    ttf = timed(lambda: torch.mm(a, b))
    etf = ((torch.mm(a, b).double() - ref).norm() / ref.norm()).item()
    ahi = (a.view(torch.int32) & -8192).view(torch.float32); alo = a - ahi
    bhi = (b.view(torch.int32) & -8192).view(torch.float32); blo = b - bhi
    def three():
        o = torch.mm(alo, bhi); o.addmm_(ahi, blo); o.addmm_(ahi, bhi); return o
    t3 = timed(three)
    e3 = ((three().double() - ref).norm() / ref.norm()).item()
    torch.backends.cuda.matmul.allow_tf32 = False
    fl = 2 * m * k * n
    print(f"{m}x{k}x{n}: fp32 {t32:.0f} us ({fl/t32/1e6:.0f} TF/s, err {e32:.1e}) | tf32 {ttf:.0f} us ({fl/ttf/1e6:.0f} TF/s, err {etf:.1e}) | 3xtf32 (pre-split) {t3:.0f} us (err {e3:.1e})")
