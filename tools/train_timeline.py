"""torch.profiler (CUPTI) view of the training step of bench.py (pre-sampled minibatches): GPU busy time per step and the
kernel classes it is made of -> printed table (gpurun_out/train_timeline.md when redirected).

  python tools/train_timeline.py [fused] [flat] [tc]
"""
import collections
import sys, os; sys.path.insert(0, '.')
import torch, bench, custom_sparse_ops as cso
from gnn_b200 import gather as gmod, harness
from torch.profiler import profile, ProfilerActivity
class A: pass
args = A(); args.workload = 'reddit'; args.minibatches = 3; args.buffer_size = 0.1; args.steps = 10; args.warmup = 3
fused, flat, tc = 'fused' in sys.argv, 'flat' in sys.argv, 'tc' in sys.argv
log = lambda m: None
device = torch.device('cuda', 0)
shape, g, mbs, samp, batch = bench.build_workload(args, 0, 1, log)
store = bench.build_store(args, gmod, shape, g, device, 0, 1, log)
r = harness.bench_train(args, cso, store, shape, g, mbs, bench.ORDERS, bench.NHID, device, 0, 1, log, fused=fused, flat_grads=flat, tc=tc)
print(f"plain (fused={fused}, flat={flat}): {r['minibatches_per_s']} minibatches/s, {r['ms_per_step_device']} ms device, {r['ms_per_step_wall']} ms wall", flush=True)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    r = harness.bench_train(args, cso, store, shape, g, mbs, bench.ORDERS, bench.NHID, device, 0, 1, log, fused=fused, flat_grads=flat, tc=tc)
print(f"profiled: {r['ms_per_step_device']} ms device", flush=True)
ev = [e for e in prof.profiler.kineto_results.events()] if False else None
ka = prof.key_averages()
steps = 6 + r['steps'] * (2 if r.get('timed_region_repeated_after_allocator_growth') else 1)   # warm-up + timed steps of bench_train
from torch.autograd import DeviceType
rows = sorted([(k.self_device_time_total, k.count, k.key) for k in ka
               if k.device_type == DeviceType.CUDA and k.self_device_time_total > 0], reverse=True)      # kernels and memcpys only
tot = sum(t for t, _, _ in rows)
print(f"\nGPU busy per step: {tot / steps / 1e3:.3f} ms over {steps} steps; kernels per step: {sum(c for _, c, _ in rows) / steps:.0f}\n")
print("| us/step | launches/step | kernel |\n|---:|---:|---|")
for t, c, name in rows[:45]:
    print(f"| {t / steps:.1f} | {c / steps:.1f} | `{name.replace('(anonymous namespace)::', '')[:110]}` |")
