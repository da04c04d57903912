"""Live-sampler training rate under a restricted core count (taskset), spin vs blocking synchronisation, N sampler threads.

  taskset -c 0-3 python tools/live_probe.py <threads> <blocking 0|1>
"""
import os
import sys
sys.path.insert(0, '.')
import torch
import bench
import custom_sparse_ops as cso
from gnn_b200 import gather as gmod, harness

threads, blocking = int(sys.argv[1]), int(sys.argv[2])


class A:
    pass


args = A(); args.workload = 'reddit'; args.minibatches = 3; args.buffer_size = 0.1; args.steps = 24; args.warmup = 3
log = lambda m: None
device = torch.device('cuda', 0)
torch.cuda.set_device(device)
if blocking:
    cso.spmm_cpp.set_blocking_sync(True)
shape, g, mbs, samp, batch = bench.build_workload(args, 0, 1, log)
store = bench.build_store(args, gmod, shape, g, device, 0, 1, log)
r = harness.bench_train_live(args, cso, store, shape, g, bench.ORDERS, bench.NHID, samp, batch, device, 0, 1, log, pool_num=threads,
                             fused=True, flat_grads=True, tc=True)
print(f"cores {len(os.sched_getaffinity(0))} threads {threads} blocking {blocking}: {r['minibatches_per_s']} minibatches/s, "
      f"{r['ms_per_step_wall']} ms/step", flush=True)
