"""How much does work on a side stream slow the training stream?  

Main stream: the five SpMMs of one Reddit-shaped minibatch / 100 tiny kernels / 100 memsets.
Side stream: nothing / a looping host-row gather (zero-copy PCIe reads) / device-row gathers / H2D DMA copies.
CORUN=<n> sets gnn_set_corunner_ctas(n) first.   usage: CORUN=16 python tools/corunner_contention.py
"""
import sys, os; sys.path.insert(0, '.')
import torch, bench, custom_sparse_ops as cso
from gnn_b200 import gather as gmod
class A: pass
args = A(); args.workload='reddit'; args.minibatches=3; args.buffer_size=0.1; args.steps=12; args.warmup=3
log = lambda m: None
device = torch.device('cuda', 0)
shape, g, mbs, samp, batch = bench.build_workload(args, 0, 1, log)
store = bench.build_store(args, gmod, shape, g, device, 0, 1, log)
nl = len(mbs[0].layers)
widths = bench.layer_widths(shape.feat_dim, nl, gcn=shape.self_loops)
mb = mbs[0]
adjs, xs, gs = [], [], []
for li, (layer, D) in enumerate(zip(mb.layers, widths)):
    a = cso.create_coo_tensor(torch.from_numpy(layer.fullrowptr).to(device), torch.from_numpy(layer.rowptr).to(device),
                              torch.from_numpy(layer.colidx).to(device), torch.from_numpy(layer.normfact).to(device), layer.nrows, layer.ncols)
    adjs.append(cso.adjacency_of(a))
    xs.append(torch.randn(layer.ncols, gmod.padded_ld(D), device=device)[:, :D])
    gs.append(torch.randn(layer.nrows, D, device=device) if li > 0 else None)
small = torch.zeros(1024, device=device)
print('corunner prev', store.ext.set_corunner_ctas(int(os.environ.get('CORUN', '0'))), 'now', os.environ.get('CORUN', '0'), flush=True)
def main_work(kind):
    if kind == "spmm":
        for li in range(nl): adjs[li].matmul(xs[li])
        for li in range(1, nl):
            adjs[li]._t = None
            adjs[li].matmul_t(gs[li])
    elif kind == "tiny":
        for _ in range(100): small.add_(1.0)
    elif kind == "memset":
        for _ in range(100): small.zero_()
nodes = torch.from_numpy(mb.input_nodes).to(device)
src_dev, slot, xrows, c = store.remap(nodes)
out = torch.empty((nodes.numel(), store.ld), device=device)
side = torch.cuda.Stream()
pin = torch.empty(64 << 20, dtype=torch.uint8).pin_memory(); dbuf = torch.empty(64 << 20, dtype=torch.uint8, device=device)
def side_work(kind, n):
    with torch.cuda.stream(side):
        for _ in range(n):
            if kind == "hostgather": store.ext.gather_rows_src(xrows, src_dev, -1, store.feat_dim, out)
            elif kind == "devgather": store.ext.gather_rows_src(xrows, src_dev, 0, store.feat_dim, out)
            elif kind == "h2d": dbuf.copy_(pin, non_blocking=True)
for mk in ("spmm", "tiny", "memset"):
    for sk, n in (("none", 0), ("hostgather", 6), ("devgather", 400), ("h2d", 4)):
        ts = []
        for rep in range(4):
            torch.cuda.synchronize()
            side_work(sk, n)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); main_work(mk); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"main={mk:7s} side={sk:10s}: " + " ".join(f"{t:.3f}" for t in ts), flush=True)
