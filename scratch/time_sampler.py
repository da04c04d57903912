import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from gnn_b200 import graphgen, sampler, gpu_sampler
g = graphgen.generate_cached('reddit')
dg = gpu_sampler.DeviceGraph(g.indptr, g.indices, 'cuda')
rng = np.random.Generator(np.random.PCG64(0))
bns = [g.train_nodes[rng.permutation(g.train_nodes.size)[:512]] for _ in range(6)]
t = time.perf_counter(); ref = sampler.ladies_sample(1234, bns[0], [8192]*5, g.num_nodes, g.indptr, g.indices, [1,1,1]); th = time.perf_counter()-t
got = gpu_sampler.ladies_sample_device(1234, bns[0], [8192]*5, dg, [1,1,1]); torch.cuda.synchronize()
ok = all(np.array_equal(a.colidx.cpu().numpy(), b.colidx) and np.array_equal(a.rowptr.cpu().numpy(), b.rowptr) for a, b in zip(got.layers, ref.layers))
ts = []
for i, bn in enumerate(bns[1:]):
    torch.cuda.synchronize(); t = time.perf_counter()
    mb = gpu_sampler.ladies_sample_device(2000+i, bn, [8192]*5, dg, [1,1,1]); torch.cuda.synchronize()
    ts.append(time.perf_counter()-t)
print(f"host mirror {th*1e3:.0f} ms; device sampler {np.median(ts)*1e3:.1f} ms (min {min(ts)*1e3:.1f}); identical={ok}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
mb = gpu_sampler.ladies_sample_device(3000, bns[0], [8192]*5, dg, [1,1,1]); torch.cuda.synchronize()
pr.disable(); pstats.Stats(pr).sort_stats('cumulative').print_stats(14)
