"""Device LADIES sampler alone (idle GPU, one thread): wall ms per minibatch on a generated graph shape.

  python tools/sampler_alone.py [reddit|products|papers16] [layers]
"""
import sys
import time
sys.path.insert(0, '.')
import numpy as np
import torch
import custom_sparse_ops as cso
from gnn_b200 import graphgen, gpu_sampler

name = sys.argv[1] if len(sys.argv) > 1 else 'reddit'
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = graphgen.generate_cached(name)
dg = gpu_sampler.DeviceGraph(g.indptr, g.indices, 'cuda')
rng = np.random.Generator(np.random.PCG64(0))
bns = [g.train_nodes[rng.permutation(g.train_nodes.size)[:512]] for _ in range(14)]
ts = []
for i, bn in enumerate(bns):
    torch.cuda.synchronize(); t = time.perf_counter()
    mb = gpu_sampler.ladies_sample_device(2000 + i, bn, [8192] * 5, dg, [1] * layers, create_coo_tensor=cso.create_coo_tensor)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t)
print(f"{name}: {dg.num_nodes} nodes, {layers} layers, input rows {mb.input_nodes.size}: median {np.median(ts[4:]) * 1e3:.2f} ms per minibatch "
      f"(first calls {[round(x * 1e3, 1) for x in ts[:4]]}); device compaction {'on' if dg.num_nodes >= gpu_sampler.DEVICE_COMPACT_MIN_NODES else 'off'}")
