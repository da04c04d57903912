import sys, os, time, json
sys.path.insert(0, '.')
import numpy as np, torch
import bench, custom_sparse_ops as cso
from gnn_b200 import gather as gmod, harness, graphgen
class A: pass
args = A(); args.workload='reddit'; args.minibatches=3; args.buffer_size=0.1; args.steps=10; args.warmup=3
log = lambda m: print(m, file=sys.stderr)
device = torch.device('cuda', 0)
shape, g, mbs, samp, batch = bench.build_workload(args, 0, 1, log)
store = bench.build_store(args, gmod, shape, g, device, 0, 1, log)
# reuse harness but with profiler around steps
from torch.profiler import profile, ProfilerActivity
res = harness.bench_train(args, cso, store, shape, g, mbs, bench.ORDERS, bench.NHID, device, 0, 1, log)
print(res)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    res = harness.bench_train(args, cso, store, shape, g, mbs, bench.ORDERS, bench.NHID, device, 0, 1, log)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=60))
