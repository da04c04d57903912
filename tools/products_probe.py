"""BASELINE configs[3] alone (products-shaped 3-layer GCN, locality sampling scale factor 2, live device sampler): the
`other_workloads.products_gcn_locality` object of bench.py without the rest of the run.

  python tools/products_probe.py
"""
import argparse
import json
import sys
sys.path.insert(0, '.')
import torch
import bench
import custom_sparse_ops as cso
from gnn_b200 import gather as gmod

args = argparse.Namespace(steps=20, warmup=3, buffer_size=0.1, minibatches=3, verbose=False)
device = torch.device('cuda', 0)
torch.cuda.set_device(device)
flush_buf = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=device)
out = bench.run_products_locality(args, cso, gmod, device, 0, 1, flush_buf, lambda m: None)
for k, v in out.items():
    if k.startswith('train'):
        print(k, json.dumps({kk: v.get(kk) for kk in ('minibatches_per_s', 'ms_per_step_wall', 'sampler_threads', 'sampler_job_ms',
                                                      'trainer_wait_ms_per_step', 'steps', 'error')}))
print('parity', {k: v.get('parity_ok') for k, v in out.get('sampling', {}).items()})
