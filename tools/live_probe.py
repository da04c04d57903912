"""Live-sampler training rate under a restricted core count, N sampler threads, sampler-stream priority.

  python tools/live_probe.py "<cores>:<threads>:<priority>[:<sync 0 spin|1 blocking|2 yield>[:<split gather 0|1>]]" ...     e.g. 4:4:0 4:4:-1 16:8:0

One process, one graph: the configurations run back to back (the host-core restriction is sched_setaffinity on the
whole process, set before the sampler pool of that configuration starts).  Also prints the sampler's own time per
minibatch on an idle GPU (single thread, no training beside it).
"""
import os
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import bench
import custom_sparse_ops as cso
from gnn_b200 import gather as gmod, gpu_sampler, harness


class A:
    pass


args = A(); args.workload = 'reddit'; args.minibatches = 3; args.buffer_size = 0.1; args.steps = 24; args.warmup = 3
log = lambda m: None
device = torch.device('cuda', 0)
torch.cuda.set_device(device)
all_cores = sorted(os.sched_getaffinity(0))
shape, g, mbs, samp, batch = bench.build_workload(args, 0, 1, log)
store = bench.build_store(args, gmod, shape, g, device, 0, 1, log)

dg = gpu_sampler.DeviceGraph(g.indptr, g.indices, device)
rng = np.random.Generator(np.random.PCG64(0))
bns = [g.train_nodes[rng.permutation(g.train_nodes.size)[:batch]] for _ in range(16)]
ts = []
for i, bn in enumerate(bns):
    torch.cuda.synchronize(); t = time.perf_counter()
    gpu_sampler.ladies_sample_device(2000 + i, bn, [samp] * 5, dg, bench.ORDERS, create_coo_tensor=cso.create_coo_tensor)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t)
print(f"sampler alone, idle GPU: median {np.median(ts[4:]) * 1e3:.2f} ms per minibatch (first calls {[round(x * 1e3, 1) for x in ts[:4]]})", flush=True)
if os.environ.get("LIVE_PROBE_PROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for i in range(8):
        gpu_sampler.ladies_sample_device(3000 + i, bns[i], [samp] * 5, dg, bench.ORDERS, create_coo_tensor=cso.create_coo_tensor)
    torch.cuda.synchronize()
    pr.disable()
    print("cProfile of 8 minibatches (sampler alone):")
    pstats.Stats(pr).sort_stats('tottime').print_stats(22)
del dg

for spec in sys.argv[1:]:
    f = spec.split(':')
    cores, threads, prio = int(f[0]), int(f[1]), int(f[2])
    blocking = int(f[3]) if len(f) > 3 else 0
    co_split = bool(int(f[4])) if len(f) > 4 else True
    for tid in os.listdir('/proc/self/task'):          # every existing thread (autograd engine, CUDA workers), new ones inherit
        try:
            os.sched_setaffinity(int(tid), all_cores[:cores])
        except OSError:
            pass
    cso.spmm_cpp.set_blocking_sync(blocking)      # 0 spin, 1 blocking, 2 yield
    r = harness.bench_train_live(args, cso, store, shape, g, bench.ORDERS, bench.NHID, samp, batch, device, 0, 1, log, pool_num=threads,
                                 fused=True, flat_grads=True, tc=True, sampler_stream_priority=prio, co_split=co_split)
    print(f"cores {len(os.sched_getaffinity(0))} threads {threads} priority {prio} blocking {blocking} split-gather {int(co_split)}: {r['minibatches_per_s']} minibatches/s, "
          f"{r['ms_per_step_wall']} ms/step over {r['steps']} steps; sampler job {r['sampler_job_ms']} ms, trainer waits {r['trainer_wait_ms_per_step']} ms/step, cudaMallocs {r['cuda_mallocs_in_timed_region']}, repeated {r['timed_region_repeated_after_allocator_growth']}", flush=True)
