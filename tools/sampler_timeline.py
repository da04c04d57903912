"""torch.profiler (CUPTI) view of the device LADIES sampler alone (Reddit-shaped minibatches, idle GPU): device time per
minibatch by kernel -> printed table.  Also the job the live-training sampler threads run (sampler + input gather).

  python tools/sampler_timeline.py
"""
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import bench
import custom_sparse_ops as cso
from gnn_b200 import gather as gmod, gpu_sampler
from torch.autograd import DeviceType
from torch.profiler import profile, ProfilerActivity


class A:
    pass


args = A(); args.workload = 'reddit'; args.minibatches = 3; args.buffer_size = 0.1; args.steps = 10; args.warmup = 3
log = lambda m: None
device = torch.device('cuda', 0)
torch.cuda.set_device(device)
shape, g, mbs, samp, batch = bench.build_workload(args, 0, 1, log)
store = bench.build_store(args, gmod, shape, g, device, 0, 1, log)
dg = gpu_sampler.DeviceGraph(g.indptr, g.indices, device)
rng = np.random.Generator(np.random.PCG64(0))
bns = [g.train_nodes[rng.permutation(g.train_nodes.size)[:batch]] for _ in range(16)]


def job(i):
    mb = gpu_sampler.ladies_sample_device(2000 + i, bns[i], [samp] * 5, dg, bench.ORDERS, create_coo_tensor=cso.create_coo_tensor)
    x0 = store.gather(gpu_sampler.h2d(mb.input_nodes, device))
    return mb, x0


for i in range(4):
    job(i)
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(4, 12):
    job(i)
torch.cuda.synchronize()
print(f"sampler + gather job, idle GPU: {(time.perf_counter() - t) / 8 * 1e3:.2f} ms per minibatch wall")
n = 8
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(n):
        job(i)
    torch.cuda.synchronize()
rows = sorted([(k.self_device_time_total, k.count, k.key) for k in prof.key_averages()
               if k.device_type == DeviceType.CUDA and k.self_device_time_total > 0], reverse=True)
tot = sum(t for t, _, _ in rows)
print(f"\ndevice time per minibatch: {tot / n / 1e3:.3f} ms; launches + copies per minibatch: {sum(c for _, c, _ in rows) / n:.0f}\n")
print("| us/minibatch | launches/minibatch | kernel |\n|---:|---:|---|")
for t, c, name in rows[:30]:
    print(f"| {t / n:.1f} | {c / n:.1f} | `{name.replace('(anonymous namespace)::', '')[:110]}` |")
