#!/usr/bin/env python
"""bench.py - LADIES-layer SpMM (fwd A.X + bwd A^T.G) on synthetic Reddit-shaped minibatches.

Contract (driver): python bench.py --gpus N --steps K --warmup W [--impl reference]
prints ONE JSON line on rank 0.

  step      one minibatch of BASELINE.json configs[1] (GraphSAGE, LADIES samp_num 8192, batch 512,
            Reddit-shaped synthetic graph): forward SpMM of its 3 layer blocks (D = 602, 1024, 1024)
            and backward SpMM (CSR-of-A^T build + product) of layers 1 and 2 - the deepest block has
            no backward.
  value     whole-job algorithmic GB/s of those SpMMs (SURVEY.md 8(d) byte formula), inputs resident
            in HBM, CUDA-event time of the steps, max over ranks; L2 flushed between steps.
  e2e       same metric through the public API with HOST inputs: pinned sampler CSR arrays -> H2D ->
            create_coo_tensor, placement remap + feature gather (local shard / peer shards over NVLink /
            mapped pinned host), spmm forward + autograd backward, D2H of the loss.  The hand-off runs in
            pipeline.DevicePrefetcher (worker thread, two side streams); the pipeline is EMPTY when the clock
            starts, so every copy/build/gather of the K timed steps is inside the timed region.
  warm_l2   companion of `value` without the L2 flush.
  roofline  dominant kernel (the forward/backward row-split SpMM launch with the largest share).
  cpu_baseline / --impl reference
            the reference's CPU path, torch.sparse COO mm (reference custom_sparse_ops.py:25,36), on a
            bounded row-sample of the same blocks, best of a thread sweep.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

ORDERS = [1, 1, 1]
NHID = 512
# Measured on this pool's B200 (profiles/gather_roof_r1.txt): a kernel that only performs the SpMM's row gather
# (same column stream, float4 loads, no FMA/stores) tops out at ~20 TB/s of L2->SM traffic.
L2_GATHER_ROOF_BPS = 20.0e12


# ----------------------------------------------------------------------------- workload
def algorithmic_bytes(nnz, M, K, D):
    """each operand once, fp32 values, int32 indices (same formula both directions)."""
    return 8 * nnz + 4 * (M + 1) + 4 * K * D + 4 * M * D


def layer_widths(feat_dim, nlayers, gcn=False):
    """SpMM operand widths. GraphSAGE aggregates before the linear and concatenates: nfeat, then 2*nhid
    (reference models.py:18-19,34-36); GCN: nfeat, then nhid (models.py:60-61,73-76)."""
    return [feat_dim] + [(NHID if gcn else 2 * NHID)] * (nlayers - 1)


def build_workload(args, rank, world, log):
    from gnn_b200 import graphgen, sampler
    shape = graphgen.SHAPES[args.workload]
    t0 = time.time()
    cache_root = os.path.join(REPO, ".cache") if os.path.isdir(os.path.join(REPO, ".cache")) else None
    g = graphgen.generate_cached(shape, seed=0, root=cache_root)
    log(f"graph {shape.name}: {g.num_nodes} nodes, {g.nnz} directed nnz, max deg {int(g.degrees().max())} ({time.time() - t0:.1f}s)")
    samp, batch = (8192, 512) if args.workload != "small" else (2048, 256)
    if args.workload == "cora":
        samp, batch = 512, 256
    # rank r trains on its own chunk of one shuffled permutation (reference sampler.py:170-189)
    rng = np.random.Generator(np.random.PCG64(1000 + rank))
    chunk = (g.train_nodes.size + world - 1) // world
    own = g.train_nodes[rank * chunk: min((rank + 1) * chunk, g.train_nodes.size)]
    mbs = []
    t0 = time.time()
    for i in range(args.minibatches):
        batch_nodes = own[rng.permutation(own.size)[:batch]]
        mbs.append(sampler.ladies_sample(1234 + 100 * rank + i, batch_nodes, [samp] * 5, g.num_nodes, g.indptr, g.indices, ORDERS))
    log(f"sampled {len(mbs)} minibatches ({time.time() - t0:.1f}s): " + "; ".join(
        f"{l.nrows}x{l.ncols} nnz {l.nnz}" for l in mbs[0].layers))
    return shape, g, mbs, samp, batch


def block_stats(mb, widths):
    out = []
    for li, (layer, D) in enumerate(zip(mb.layers, widths)):
        rl = np.diff(layer.rowptr)
        out.append({"layer": li, "M": layer.nrows, "K": layer.ncols, "nnz": layer.nnz, "D": D,
                    "row_nnz_mean": round(float(rl.mean()), 1), "row_nnz_max": int(rl.max()),
                    "density": round(layer.nnz / (layer.nrows * layer.ncols), 5)})
    return out


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            mx = max(mx, cmax)
            if t_begin - 0.05 <= ts <= t_end + 0.15:
                sm.append(clk)
                for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(r[1].split(",")[1]) for r in self.rows[-3:] if len(r[1].split(",")) > 2] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU reference path
def cpu_reference_sample(mb, widths, target_s, log, steps=1, warmup=0):
    """torch.sparse COO `mat1.mm(mat2)` and `mat1.transpose(0,1).mm(g)` (reference custom_sparse_ops.py:25,36)
    on the first rows of every layer block of one minibatch; best thread count of a sweep."""
    import torch
    import oracle
    blocks = []
    total_nnz = sum(l.nnz for l in mb.layers)
    # bounded sample: keep a fraction of the rows of each block so the whole pass is ~target_s at ~0.25 GFMA/s/thread
    work = sum(l.nnz * D * (1 if i == 0 else 2) for i, (l, D) in enumerate(zip(mb.layers, widths)))
    frac = min(1.0, target_s * 0.5e9 / max(work, 1))
    nbytes = 0
    for li, (layer, D) in enumerate(zip(mb.layers, widths)):
        rows = max(1, int(round(layer.nrows * frac)))
        rp = layer.rowptr[:rows + 1]
        nnz = int(rp[-1])
        r_, c_, v_ = oracle.build_adj(layer.fullrowptr[:rows + 1], rp, layer.colidx32[:nnz], layer.normfact, rows)
        a = torch.sparse_coo_tensor(torch.from_numpy(np.stack([r_, c_])), torch.from_numpy(v_), (rows, layer.ncols)).coalesce()
        x = torch.randn(layer.ncols, D)
        gq = torch.randn(rows, D)
        blocks.append((li, a, x, gq))
        nbytes += algorithmic_bytes(nnz, rows, layer.ncols, D) * (1 if li == 0 else 2)
    ncpu = len(os.sched_getaffinity(0))
    sweep = sorted({t for t in (1, 2, 4, 8, 16, 32, ncpu) if t <= ncpu})

    def one_pass(split=None):
        t0 = time.perf_counter()
        for li, a, x, gq in blocks:
            t1 = time.perf_counter()
            a.mm(x)                                          # custom_sparse_ops.py:25
            t2 = time.perf_counter()
            if li > 0:
                a.transpose(0, 1).mm(gq)                     # custom_sparse_ops.py:36
            if split is not None:
                split[0] += t2 - t1
                split[1] += time.perf_counter() - t2
        return time.perf_counter() - t0

    default_threads = torch.get_num_threads()
    sweep_ms = {}
    best_t, best_threads = None, 1
    for th in sweep:
        torch.set_num_threads(th)
        one_pass()
        t = min(one_pass() for _ in range(2))
        sweep_ms[str(th)] = round(t * 1e3, 1)
        if best_t is None or t < best_t:
            best_t, best_threads = t, th
    torch.set_num_threads(best_threads)
    for _ in range(max(warmup, 2)):
        one_pass()
    split = [0.0, 0.0]
    nrep = max(steps, 5) if steps <= 1 else steps
    times = [one_pass(split) for _ in range(nrep)]
    t = float(np.median(times))
    # the reference's gather block (main.py:129-134) on CPU tensors: boolean-mask scatter of fancy-indexed rows
    n0, F = mb.layers[0].ncols, widths[0]
    table = torch.randn(max(4 * n0, 1024), F)
    ids = torch.randint(0, table.shape[0], (n0,))
    mask = torch.ones(n0, dtype=torch.bool)
    tg = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = torch.empty(n0, F)
        out[mask] = table[ids].float()
        tg.append(time.perf_counter() - t0)
    sample = (f"first {frac * 100:.1f}% of the rows of each of the {len(blocks)} layer blocks of one minibatch "
              f"(torch.sparse COO mm fwd + transpose().mm bwd, reference custom_sparse_ops.py:25,36), "
              f"best of threads {sweep} on {ncpu} usable cores, median of {nrep} passes after 2 warm-ups")
    log(f"cpu reference: {nbytes / t / 1e9:.3f} GB/s at {best_threads} threads, {t * 1e3:.1f} ms per sampled pass, frac {frac:.4f}")
    return {"value": nbytes / t / 1e9, "unit": "GB/s", "cores": best_threads, "kind": "reference", "sample": sample,
            "ms_per_sample": t * 1e3, "fwd_ms": round(split[0] / nrep * 1e3, 1), "bwd_ms": round(split[1] / nrep * 1e3, 1),
            "thread_sweep_ms": sweep_ms, "host_cpus": os.cpu_count(), "usable_cores": ncpu, "torch_default_threads": default_threads,
            "gather_cpu_ms": round(float(np.median(tg)) * 1e3, 2), "total_nnz_full": total_nnz}, times


# ----------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="reddit", choices=["reddit", "products", "small", "cora"])
    ap.add_argument("--minibatches", type=int, default=3, help="distinct pre-sampled minibatches rotated through")
    ap.add_argument("--buffer-size", type=float, default=0.1, help="fraction of nodes cached per GPU (reference --buffer_size)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--ref-gpu", action="store_true", help="also time the reference's CUDA kernels from oracle/_ref")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: libraries that print there (NCCL's version banner) go to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def log(msg):
        if args.verbose or os.environ.get("BENCH_VERBOSE"):
            print(f"[bench r{rank}] {msg}", file=sys.stderr, flush=True)

    if args.impl == "reference":
        if rank != 0:
            return 0
        return run_reference(args, log)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    import custom_sparse_ops as cso
    from gnn_b200 import gather as gmod
    ext = cso.spmm_cpp

    def barrier():
        if world > 1:
            dist.barrier()

    # rank 0 generates the graph first so the others hit the cache
    if world > 1 and rank != 0:
        barrier()
    shape, g, mbs, samp, batch = build_workload(args, rank, world, log)
    if world > 1 and rank == 0:
        barrier()
    nl = len(mbs[0].layers)
    widths = layer_widths(shape.feat_dim, nl, gcn=shape.self_loops)

    # ---- device-resident operands per minibatch
    gen = torch.Generator(device=device)
    gen.manual_seed(7 + rank)
    dev_mbs = []
    for mb in mbs:
        adjs, xs, gs = [], [], []
        for li, (layer, D) in enumerate(zip(mb.layers, widths)):
            a = cso.create_coo_tensor(torch.from_numpy(layer.fullrowptr).to(device), torch.from_numpy(layer.rowptr).to(device),
                                      torch.from_numpy(layer.colidx).to(device), torch.from_numpy(layer.normfact).to(device),
                                      layer.nrows, layer.ncols)
            adjs.append(cso.adjacency_of(a))
            # rows padded to 16 bytes, exactly like the buffer the product's gather hands to the first SpMM
            ld = gmod.padded_ld(D)
            xs.append(torch.randn(layer.ncols, ld, device=device, generator=gen)[:, :D])
            gs.append(torch.randn(layer.nrows, D, device=device, generator=gen) if li > 0 else None)
        dev_mbs.append((adjs, xs, gs))
    step_bytes = [sum(algorithmic_bytes(l.nnz, l.nrows, l.ncols, D) * (1 if li == 0 else 2)
                      for li, (l, D) in enumerate(zip(mb.layers, widths))) for mb in mbs]

    flush_buf = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=device)   # 3x L2
    flush_src = torch.zeros(96 * 1024 * 1024, dtype=torch.int32, device=device)   # 384 MiB, read-only

    def flush_l2():
        # write 384 MiB, then read 384 MiB: the cache ends up full of CLEAN foreign lines, so the timed kernels
        # start cold without also paying for the write-back of the flush's own dirty lines
        flush_buf.zero_()
        flush_src.sum()

    op_names = [f"fwd{li}" for li in range(nl)] + [f"bwd{li}" for li in range(1, nl)]

    def run_step(i, events=None):
        adjs, xs, gs = dev_mbs[i % len(dev_mbs)]
        k = 0
        for li in range(nl):
            if events is not None:
                events[k][0].record()
            adjs[li].matmul(xs[li])
            if events is not None:
                events[k][1].record()
            k += 1
        for li in range(1, nl):
            adjs[li]._t = None                      # a fresh adjacency every minibatch: the A^T build is part of backward
            if events is not None:
                events[k][0].record()
            adjs[li].matmul_t(gs[li])
            if events is not None:
                events[k][1].record()
            k += 1

    # the clock sampler (nvidia-smi -lms) starts BEFORE the warm-up: its start-up (NVML init, first query) stalls
    # launches for milliseconds, which must not land inside the timed region
    clocks = ClockSampler(local_rank)
    clocks.start()

    # ---- warm-up
    for i in range(args.warmup):
        flush_l2()
        run_step(i)
    torch.cuda.synchronize()
    barrier()

    # ---- timed region: K steps, CUDA events per step and per op on the launching (current) stream
    import gc
    gc.collect()
    gc.disable()            # a generational collection pause between two event records would be billed to that op
    t_wait = time.time()
    while len(clocks.rows) < 2 and time.time() - t_wait < 5.0:
        time.sleep(0.05)
    launches0 = ext.launch_count()
    step_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    op_ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in op_names] for _ in range(args.steps)]
    torch.cuda.synchronize()
    barrier()
    t_begin = time.time()
    wall0 = time.perf_counter()
    for s in range(args.steps):
        flush_l2()
        step_ev[s][0].record()
        run_step(s, op_ev[s])
        step_ev[s][1].record()
    torch.cuda.synchronize()
    barrier()
    wall = time.perf_counter() - wall0
    t_end = time.time()
    gc.enable()
    launches = ext.launch_count() - launches0
    clk = clocks.stop(t_begin, t_end)

    step_ms = np.array([a.elapsed_time(b) for a, b in step_ev])
    if os.environ.get("BENCH_DUMP_OPS"):
        print(np.array2string(np.array([[a.elapsed_time(b) for a, b in row] for row in op_ev]), precision=3, max_line_width=200),
              file=sys.stderr)
    op_ms = np.array([[a.elapsed_time(b) for a, b in row] for row in op_ev])          # [steps, ops]
    total_ms = float(step_ms.sum())
    total_bytes = float(sum(step_bytes[s % len(mbs)] for s in range(args.steps)))
    if world > 1:
        t = torch.tensor([total_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms_max = float(t.item())
        b = torch.tensor([total_bytes], device=device, dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        total_bytes_all = float(b.item())
    else:
        total_ms_max, total_bytes_all = total_ms, total_bytes
    value = total_bytes_all / (total_ms_max * 1e-3) / 1e9

    # ---- companion number: the same K steps WITHOUT the flush, one minibatch repeated - in training X
    # was just produced and A just built, so the operands of a step are partly L2-resident.  Not the headline.
    wev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for _ in range(2):
        run_step(0)
    wev[0].record()
    for s in range(args.steps):
        run_step(0)
    wev[1].record()
    torch.cuda.synchronize()
    warm_ms = wev[0].elapsed_time(wev[1]) / args.steps
    warm_l2 = {"value": round(step_bytes[0] / (warm_ms * 1e-3) / 1e9, 2), "unit": "GB/s per GPU", "ms_per_step": round(warm_ms, 4),
               "note": "no L2 flush, one minibatch repeated back to back (rank 0's figure)"}

    # ---- per-op table + roofline of the dominant launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    sm_clk_ghz = (clk["sm_mhz"] if clk and clk["sm_mhz"] > 0 else 1965.0) / 1e3
    ops = []
    for k, name in enumerate(op_names):
        li = int(name[3:])
        bytes_k = np.array([algorithmic_bytes(mbs[s % len(mbs)].layers[li].nnz, mbs[s % len(mbs)].layers[li].nrows,
                                              mbs[s % len(mbs)].layers[li].ncols, widths[li]) for s in range(args.steps)], dtype=np.float64)
        nnzD = np.array([mbs[s % len(mbs)].layers[li].nnz * widths[li] for s in range(args.steps)], dtype=np.float64)
        ms = op_ms[:, k]
        t_hbm = bytes_k.mean() / (hbm_peak * 1e9)
        t_l2 = 4 * nnzD.mean() / L2_GATHER_ROOF_BPS                            # gather bytes / measured pure-gather roof
        t_fma = nnzD.mean() / (148 * 128 * sm_clk_ghz * 1e9)
        ops.append({"op": name, "ms": round(float(ms.mean()), 4), "ms_median": round(float(np.median(ms)), 4),
                    "ms_max": round(float(ms.max()), 4), "share": round(float(ms.sum() / op_ms.sum()), 4),
                    "algorithmic_GBps": round(float(bytes_k.sum() / (ms.sum() * 1e-3) / 1e9), 1),
                    "bytes": int(bytes_k.mean()), "gather_GB": round(float(4 * nnzD.mean() / 1e9), 3),
                    "gflop": round(float(2 * nnzD.mean() / 1e9), 3),
                    "t_bound_us": {"hbm": round(t_hbm * 1e6, 1), "l2_gather": round(t_l2 * 1e6, 1), "fma": round(t_fma * 1e6, 1)},
                    "frac_of_t_bound": round(float(max(t_hbm, t_l2, t_fma) / (ms.mean() * 1e-3)), 3)})
    dom = max(ops, key=lambda o: o["share"])
    roofline = {"bound": "hbm", "kernel": "spmm_rowsplit_kernel", "op": dom["op"], "achieved": dom["algorithmic_GBps"],
                "peak": hbm_peak, "unit": "GB/s", "frac": round(dom["algorithmic_GBps"] / hbm_peak, 4), "traffic": None,
                "peak_source": peak_src, "binding_roof": max(dom["t_bound_us"], key=dom["t_bound_us"].get),
                "frac_of_binding_roof": dom["frac_of_t_bound"],
                "binding_roof_note": "l2_gather = 4*nnz*D bytes requested from L2 / 20 TB/s, the measured speed of light of a "
                                     "pure row-gather kernel on this GPU (profiles/gather_roof_r1.txt)"}
    prof = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(dom["op"])
        except (OSError, ValueError):
            pass

    # ---- end-to-end through the public API with host inputs
    e2e, train, store = None, None, None
    if not (args.no_e2e and args.no_train):
        store = build_store(args, gmod, shape, g, device, rank, world, log)
    if not args.no_e2e:
        e2e = run_e2e(args, cso, store, mbs, widths, step_bytes, device, rank, world, log)

    # ---- training minibatches/s (full step incl. NCCL allreduce), secondary metric of BASELINE.json
    if not args.no_train and shape.self_loops:
        train = {"skipped": "the training harness restates the GraphSAGE model only; GCN shapes report the SpMM path"}
    elif not args.no_train:
        from gnn_b200 import harness
        # reference-shaped model first (the reference's own models.py would run exactly these torch ops), then the
        # same model with the fused ELU+row-norm epilogue of gnn_b200/models.py
        train = harness.bench_train(args, cso, store, shape, g, mbs, ORDERS, NHID, device, rank, world, log)
        # sampler threads per GPU: the reference's default --pool_num is 4 (main.py:77); with the host cores to spare
        # (>= 2 per thread and rank) the second live number uses 8, since 4 threads x 26 ms per minibatch is sampler-bound
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
        pool_wide = 8 if cores // max(world, 1) >= 16 else 4
        for key, fn, kw in [("fused_epilogue_model", harness.bench_train, dict(fused=True)),
                            ("live_sampler", harness.bench_train_live, dict(fused=False, pool_num=4)),
                            ("live_sampler_fused_epilogue", harness.bench_train_live, dict(fused=True, pool_num=pool_wide))]:
            try:
                if fn is harness.bench_train:
                    train[key] = fn(args, cso, store, shape, g, mbs, ORDERS, NHID, device, rank, world, log, **kw)
                else:
                    train[key] = fn(args, cso, store, shape, g, ORDERS, NHID, samp, batch, device, rank, world, log, **kw)
            except Exception as exc:                   # secondary numbers must not take the headline down
                train[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if store is not None:
        store.close()

    ref_gpu = None
    if args.ref_gpu and rank == 0:
        ref_gpu = run_ref_gpu(mbs, widths, device, log)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_reference_sample(mbs[0], widths, args.cpu_seconds, log)

    if rank == 0:
        line = {
            "metric": "LADIES-layer SpMM HBM GB/s (fwd+bwd)", "value": round(value, 2), "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms_max / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{shape.name}-shaped {'GCN' if shape.self_loops else 'GraphSAGE'} LADIES samp_num {samp} batch {batch}"
                                   + (" (BASELINE configs[1])" if shape.name == "reddit" else ""),
                       "graph": {"nodes": g.num_nodes, "directed_nnz": g.nnz, "feat_dim": shape.feat_dim, "alpha": shape.alpha,
                                 "max_degree": int(g.degrees().max())},
                       "blocks": block_stats(mbs[0], widths), "minibatches_rotated": len(mbs),
                       "l2": "flushed between steps (384 MiB write + 384 MiB read) and inputs rotate over >L2 working sets",
                       "sharding": "each rank its own minibatches, no data-path collective in `value`",
                       "bwd_includes": "CSR-of-A^T build (gnn_csr_transpose) every step"},
            "wall_ms_per_step_incl_flush": round(wall / args.steps * 1e3, 4),
            "gpu_launches": int(launches), "clocks": clk, "ops": ops, "warm_l2": warm_l2, "roofline": roofline,
            "e2e": e2e, "train": train, "cpu_baseline": cpu_baseline,
        }
        if ref_gpu is not None:
            line["reference_cuda_kernels_same_gpu"] = ref_gpu
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def build_store(args, gmod, shape, g, device, rank, world, log):
    """Placement tables (reference create_buffer semantics, alpha = 0) + the placement-partitioned feature store."""
    import torch
    import torch.distributed as dist
    from gnn_b200 import graphgen, placement
    t0 = time.time()
    devices = list(range(world))
    buffer_rows = int(args.buffer_size * g.num_nodes)
    if world > 1:
        pl = placement.create_placement(g.to_scipy(np.float64), g.train_nodes, buffer_rows, devices, sum(ORDERS), alpha=0.0)
        did, idx, bufs = pl.device_id_of_nodes_group[rank], pl.idx_of_nodes_on_device_group[rank], pl.gpu_buffer_group
    else:
        prob = placement.access_probability(g.to_scipy(np.float64), g.train_nodes, sum(ORDERS))
        top = np.argsort(-1 * prob)[:buffer_rows]
        did = np.full(g.num_nodes, -1, dtype=np.int64)
        did[top] = 0
        idx = np.arange(g.num_nodes, dtype=np.int64)
        idx[top] = np.arange(top.size)
        bufs = [top]
    log(f"placement: {buffer_rows} rows/GPU ({time.time() - t0:.1f}s)")
    t0 = time.time()
    feats = torch.from_numpy(graphgen.features(shape, seed=1))
    store = gmod.FeatureStore(feats, bufs, did, idx, devices, rank, device, group=(dist.group.WORLD if world > 1 else None))
    log(f"feature store: host table {tuple(feats.shape)} pinned+mapped, shard {len(bufs[rank])} rows ({time.time() - t0:.1f}s)")
    return store


def run_e2e(args, cso, store, mbs, widths, step_bytes, device, rank, world, log):
    """Public-API step with host inputs (see module docstring)."""
    import torch
    import torch.distributed as dist
    nl = len(widths)
    from gnn_b200 import pipeline
    host_mbs = [pipeline.PinnedMinibatch(mb) for mb in mbs]
    gen = torch.Generator(device=device)
    gen.manual_seed(11 + rank)
    acts = [[torch.randn(mb.layers[li].ncols, widths[li], device=device, generator=gen).requires_grad_(True) for li in range(1, nl)]
            for mb in mbs]
    # the reference uploads CSR pieces and builds adjacencies in sampler threads (sampler.py:135-139); same shape here:
    # one worker thread + side stream prepares minibatch i+1 (H2D, create_coo_tensor, remap, gather) while i computes
    pre = pipeline.DevicePrefetcher(store, cso.create_coo_tensor, device, depth=2,
                                    prebuild_transpose=os.environ.get("BENCH_E2E_PREBUILD", "1") == "1")
    src_counts = []

    loss_host = torch.empty(1, dtype=torch.float32).pin_memory()
    losses, waits, syncs = [], [], []
    defer = os.environ.get("BENCH_E2E_DEFER", "1") == "1"
    if os.environ.get("BENCH_SWITCH"):
        sys.setswitchinterval(float(os.environ["BENCH_SWITCH"]))

    def step(i, prev_ready):
        """Launch minibatch i, then read the loss of minibatch i-1 (one D2H read per step, one step late, so the
        GPU is never idle while the host launches the next step)."""
        t_a = time.perf_counter()
        adjs, x0, counts = pre.get()
        waits.append((time.perf_counter() - t_a) * 1e3)
        loss = cso.spmm(adjs[0], x0).sum()
        for li in range(1, nl):
            h = acts[i % len(mbs)][li - 1]
            h.grad = None
            loss = loss + cso.spmm(adjs[li], h).sum()
        loss.backward()
        if not defer:
            losses.append(float(loss.item()))
            return None, counts
        if prev_ready is not None:
            t_a = time.perf_counter()
            prev_ready.synchronize()
            syncs.append((time.perf_counter() - t_a) * 1e3)
            losses.append(float(loss_host[0]))
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
        ready = torch.cuda.Event()
        ready.record()
        return ready, counts

    nwarm = max(args.warmup, 3)
    for i in range(nwarm):
        pre.submit(host_mbs[i % len(mbs)])
    ready = None
    for i in range(nwarm):
        ready, _ = step(i, ready)
    if ready is not None:
        ready.synchronize()
    torch.cuda.synchronize()                  # the pipeline is EMPTY here: nothing of the timed steps has been copied yet
    losses.clear()
    waits.clear()
    syncs.clear()
    import gc
    gc.collect()
    gc.disable()            # a full collection of a torch process takes 5-12 ms: two of them landed in 20 timed steps
    if world > 1:
        dist.barrier()
    ready = None
    mallocs0 = torch.cuda.memory_stats(device).get("num_device_alloc", 0)
    seg0 = {k: torch.cuda.memory_stats(device).get(f"segment.{k}.allocated", 0) for k in ("small_pool", "large_pool")}
    t0 = time.perf_counter()
    for s in range(args.steps):               # the sampler side hands over K host minibatches; every H2D copy, build and
        pre.submit(host_mbs[(nwarm + s) % len(mbs)])     # gather of the K timed steps happens after t0 (pipeline fill included)
    marks = []
    for s in range(args.steps):
        ready, counts = step(nwarm + s, ready)
        src_counts.append(counts)
        marks.append(time.perf_counter())
    if ready is not None:
        ready.synchronize()                   # the last step's loss is read inside the timed region too
        losses.append(float(loss_host[0]))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    gc.enable()
    mallocs = torch.cuda.memory_stats(device).get("num_device_alloc", 0) - mallocs0
    log("e2e cudaMalloc in the timed region: %d (%s)" % (mallocs, ", ".join(
        f"{k} +{torch.cuda.memory_stats(device).get(f'segment.{k}.allocated', 0) - v}" for k, v in seg0.items())))
    pre.close()
    log("e2e host-side step intervals (ms): " + " ".join(f"{(b - a) * 1e3:.2f}" for a, b in zip([t0] + marks, marks))
        + f" | drain {(t0 + dt - marks[-1]) * 1e3:.2f}")
    log("e2e wait for the prefetcher (ms): " + " ".join(f"{w:.2f}" for w in waits))
    log("e2e wait for the previous loss (ms): " + " ".join(f"{w:.2f}" for w in syncs))
    assert len(losses) >= args.steps and all(np.isfinite(losses)), "every step's loss must have been read on the host"
    counts = torch.stack(src_counts).double().mean(0).cpu().numpy()      # rows per source per step
    bytes_all = float(sum(step_bytes[(nwarm + s) % len(mbs)] for s in range(args.steps)))
    if world > 1:
        t = torch.tensor([dt], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        b = torch.tensor([bytes_all], device=device, dtype=torch.float64)
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        bytes_all = float(b.item())
    F4 = store.ld * 4
    csr_bytes = float(np.mean([pm.h2d_bytes() for pm in host_mbs]))
    host_rows = float(counts[world])
    peer_rows = float(sum(counts[i] for i in range(world) if i != rank))
    out = {"value": round(bytes_all / dt / 1e9, 2), "unit": "GB/s", "ms_per_step": round(dt / args.steps * 1e3, 4),
           "h2d_bytes_per_step": int(csr_bytes + host_rows * F4), "d2h_bytes_per_step": 4,
           "h2d_detail": {"sampler_csr_bytes": int(csr_bytes), "host_feature_rows": int(host_rows),
                          "host_feature_bytes_zero_copy": int(host_rows * F4)},
           "gather_rows_per_step": {"local": int(counts[rank]), "peer": int(peer_rows), "host": int(host_rows)},
           "peer_bytes_per_step": int(peer_rows * F4), "cuda_mallocs_in_timed_region": int(mallocs),
           "api": "pipeline.DevicePrefetcher (H2D of pinned sampler arrays + custom_sparse_ops.create_coo_tensor + A^T index + "
                  "FeatureStore remap/gather on a worker thread and two side streams) + custom_sparse_ops.spmm (autograd) + loss read",
           "pipelining": "the pipeline is empty when the clock starts (the K host minibatches are handed over at t0, so the fill is timed); inputs of minibatch i+1 are copied/built/gathered while minibatch i computes and the loss of minibatch i is read (pinned D2H) after i+1 is launched; every copy and read of the K steps is inside the timed region"}
    # gather alone, for the NVLink / PCIe roofs
    nodes = host_mbs[0].input_nodes.to(device)
    src_dev, slot, xrows, c = store.remap(nodes)
    outbuf = torch.empty((nodes.numel(), store.ld), device=device)
    c = c.cpu().numpy()
    for name, sid, rows in [("local", rank, c[rank]), ("host", -1, c[world])] + (
            [("peer", None, sum(c[i] for i in range(world) if i != rank))] if world > 1 else []):
        if rows == 0:
            continue
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(4):
            if rep == 1:
                ev[0].record()
            if sid is None:
                for i in range(world):
                    if i != rank:
                        store.ext.gather_rows_src(xrows, src_dev, i, store.feat_dim, outbuf)
            else:
                store.ext.gather_rows_src(xrows, src_dev, sid, store.feat_dim, outbuf)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        out.setdefault("gather_GBps", {})[name] = round(float(rows) * store.feat_dim * 4 / (ms * 1e-3) / 1e9, 1)
    # NVLink roof of the peer-gather path: 64 Ki random rows of the next rank's shard (158 MB at F=602), one launch
    if world > 1:
        peer = (rank + 1) % world
        shard = store.shards[peer]
        gen = torch.Generator(device=device)
        gen.manual_seed(3 + rank)
        nrows = 65536
        slots = torch.randint(0, shard.shape[0], (nrows,), device=device, generator=gen)
        ptrs = shard.data_ptr() + slots * (store.ld * 4)
        big = torch.empty((nrows, store.ld), device=device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        for rep in range(4):
            if rep == 1:
                ev[0].record()
            store.ext.gather_rows_src(ptrs, torch.zeros(nrows, dtype=torch.int32, device=device), 0, store.feat_dim, big)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 3
        ok = bool(torch.equal(big[:, :store.feat_dim], shard[slots][:, :store.feat_dim]))
        gbs = nrows * store.feat_dim * 4 / (ms * 1e-3) / 1e9
        out["peer_gather_roof"] = {"GBps": round(gbs, 1), "rows": nrows, "bytes": int(nrows * store.feat_dim * 4),
                                   "frac_of_measured_peer_copy_770": round(gbs / 770.0, 3), "frac_of_nominal_900": round(gbs / 900.0, 3),
                                   "bit_exact": ok, "note": "one-sided reads of the next rank's shard over NVLink, one kernel"}
    return out


def run_ref_gpu(mbs, widths, device, log):
    """The reference's own CUDA kernels (oracle/_ref, built unmodified) on the same blocks and GPU: the kernel to beat."""
    import torch
    from oracle import build_ref
    mod = build_ref.load_ref()
    if mod is None:
        return {"unavailable": "oracle/_ref/spmm_ref.so not in this snapshot"}
    mb = mbs[0]
    res = {}
    tot_ms, tot_bytes = 0.0, 0.0
    for li, (layer, D) in enumerate(zip(mb.layers, widths)):
        a = mod.create_coo_tensor(torch.from_numpy(layer.fullrowptr).to(device), torch.from_numpy(layer.rowptr).to(device),
                                  torch.from_numpy(layer.colidx).to(device), torch.from_numpy(layer.normfact).to(device),
                                  layer.nrows, layer.ncols)
        x = torch.randn(layer.ncols, D, device=device)
        gq = torch.randn(layer.nrows, D, device=device)
        for name, fn in [("fwd", lambda: mod.spmm_load_balance(a, x))] + (
                [("bwd", lambda: mod.spmm_load_balance(a.transpose(0, 1).coalesce(), gq.contiguous()))] if li > 0 else []):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / 3 * 1e3
            res[f"{name}{li}_ms"] = round(ms, 3)
            tot_ms += ms
            tot_bytes += algorithmic_bytes(layer.nnz, layer.nrows, layer.ncols, D)
    res["algorithmic_GBps"] = round(tot_bytes / (tot_ms * 1e-3) / 1e9, 1)
    res["note"] = "spmm_load_balance (+ transpose().coalesce() in bwd), wall clock around device syncs, warm L2"
    return res


def run_reference(args, log):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only)."""
    from gnn_b200 import graphgen  # noqa: F401
    shape, g, mbs, samp, batch = build_workload(args, 0, 1, log)
    widths = layer_widths(shape.feat_dim, len(mbs[0].layers), gcn=shape.self_loops)
    cpu, times = cpu_reference_sample(mbs[0], widths, args.cpu_seconds, log, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "LADIES-layer SpMM HBM GB/s (fwd+bwd)", "value": round(cpu["value"], 4), "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(float(np.mean(times)) * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{shape.name}-shaped GraphSAGE LADIES samp_num {samp} batch {batch} (BASELINE configs[1])",
                       "blocks": block_stats(mbs[0], widths)},
            "cpu_baseline": cpu, "gpu_launches": 0,
            "e2e": {"value": round(cpu["value"], 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    sys.exit(main())
