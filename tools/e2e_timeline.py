"""torch.profiler (CUPTI) timeline of bench.py's e2e leg -> gpurun_out/e2e_trace.json; summarise with tools/e2e_timeline_summary.py"""
import sys, os; sys.path.insert(0, '.')
import torch, bench, custom_sparse_ops as cso
from gnn_b200 import gather as gmod
from torch.profiler import profile, ProfilerActivity
class A: pass
args = A(); args.workload='reddit'; args.minibatches=3; args.buffer_size=0.1; args.steps=8; args.warmup=3
log = lambda m: None
device = torch.device('cuda', 0)
shape, g, mbs, samp, batch = bench.build_workload(args, 0, 1, log)
store = bench.build_store(args, gmod, shape, g, device, 0, 1, log)
nl = len(mbs[0].layers)
widths = bench.layer_widths(shape.feat_dim, nl, gcn=shape.self_loops)
step_bytes = [sum(bench.algorithmic_bytes(l.nnz, l.nrows, l.ncols, D) * (1 if li == 0 else 2) for li, (l, D) in enumerate(zip(mb.layers, widths))) for mb in mbs]
r = bench.run_e2e(args, cso, store, mbs, widths, step_bytes, device, 0, 1, log)
print("plain:", r["value"], r["ms_per_step"], flush=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    r = bench.run_e2e(args, cso, store, mbs, widths, step_bytes, device, 0, 1, log)
print("profiled:", r["value"], r["ms_per_step"], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/e2e_trace.json")
