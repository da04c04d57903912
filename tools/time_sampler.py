import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from gnn_b200 import graphgen, sampler, gpu_sampler
g = graphgen.generate_cached('reddit')
dg = gpu_sampler.DeviceGraph(g.indptr, g.indices, 'cuda')
rng = np.random.Generator(np.random.PCG64(0))
bns = [g.train_nodes[rng.permutation(g.train_nodes.size)[:512]] for _ in range(12)]
ts = []
for i, bn in enumerate(bns):
    torch.cuda.synchronize(); t = time.perf_counter()
    mb = gpu_sampler.ladies_sample_device(2000+i, bn, [8192]*5, dg, [1,1,1]); torch.cuda.synchronize()
    ts.append(time.perf_counter()-t)
print("per-call ms:", [round(x*1e3,1) for x in ts])
print("mem reserved MB", torch.cuda.memory_reserved()/1e6, "allocated", torch.cuda.memory_allocated()/1e6)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(3):
    mb = gpu_sampler.ladies_sample_device(3000+i, bns[i], [8192]*5, dg, [1,1,1]); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats('tottime').print_stats(12)
