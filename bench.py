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
# The L2->SM row-gather roof is MEASURED in every run (measure_gather_roof: the production library's probe kernel on the
# dominant block's own column stream); this constant is only the fallback when the probe cannot run.
L2_GATHER_ROOF_FALLBACK_BPS = 18.0e12


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


def make_config(shape, g, mbs, widths, samp, batch):
    """`config` of the JSON line - identical keys in the product arm and in the reference arm."""
    return {"workload": f"{shape.name}-shaped {'GCN' if shape.self_loops else 'GraphSAGE'} LADIES samp_num {samp} batch {batch}"
                        + (" (BASELINE configs[1])" if shape.name == "reddit" else ""),
            "graph": {"nodes": g.num_nodes, "directed_nnz": g.nnz, "feat_dim": shape.feat_dim, "alpha": shape.alpha,
                      "max_degree": int(g.degrees().max())},
            "blocks": block_stats(mbs[0], widths), "minibatches_rotated": len(mbs),
            "l2": "flushed between steps (384 MiB write + 384 MiB read) and inputs rotate over >L2 working sets",
            "sharding": "each rank its own minibatches, no data-path collective in `value`",
            "bwd_includes": "everything a backward needs from a fresh adjacency: the CSR-of-A^T build (dense blocks) or the "
                            "zero fill of the transpose-free path (short-row blocks)"}


# ----------------------------------------------------------------------------- parity gate
GATE_TOL = 1e-5
GATE_ROWS = 1024


def _row_sample(rowptr, cols, vals, rows):
    """Sub-CSR of the given rows (host numpy)."""
    lens = (rowptr[rows + 1] - rowptr[rows]).astype(np.int64)
    sub = np.zeros(rows.size + 1, dtype=np.int32)
    np.cumsum(lens, out=sub[1:])
    take = np.repeat(rowptr[rows].astype(np.int64) - sub[:-1], lens) + np.arange(int(sub[-1]), dtype=np.int64)
    return sub, cols[take], vals[take]


def parity_gate(mb, widths, dev_mb, store, device, log):
    """Before anything is timed: the operands of the TIMED minibatch 0 through the product path vs the CPU oracle.

    * adjacency values of every layer bit-equal to oracle.build_adj (reference cuda_spmm.cu:787-803);
    * every timed op (fwd of each layer, bwd of layers >= 1) on GATE_ROWS evenly spaced output rows vs the oracle's
      fp64-accumulated product (reference custom_sparse_ops.py:16-37), worst-row relative L2 <= 1e-5;
    * the gathered input features of the minibatch bit-equal to feats[input_nodes] (reference main.py:129-134).
    The oracle is the checker here, never the thing measured."""
    import torch
    import oracle
    adjs, xs, gs = dev_mb
    checks, ok = [], True
    host = []
    for li, (layer, D) in enumerate(zip(mb.layers, widths)):
        _, cols, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx32, layer.normfact, layer.nrows)
        cols = cols.astype(np.int32)
        same = bool(np.array_equal(adjs[li].vals.cpu().numpy().view(np.uint32), vals.view(np.uint32)))
        checks.append({"op": f"adj{li}", "values_bit_exact": same})
        ok &= same
        host.append((cols, vals))
    for li, (layer, D) in enumerate(zip(mb.layers, widths)):
        cols, vals = host[li]
        X = xs[li].cpu().numpy()
        rows = np.unique(np.linspace(0, layer.nrows - 1, min(GATE_ROWS, layer.nrows)).astype(np.int64))
        sub = _row_sample(layer.rowptr, cols, vals, rows)
        ref = oracle.spmm_f64acc(sub[0], sub[1], sub[2], rows.size, np.ascontiguousarray(X))
        got = adjs[li].matmul(xs[li])[torch.from_numpy(rows).to(device)].cpu().numpy()
        err = oracle.rel_err(got, ref)[0]
        checks.append({"op": f"fwd{li}", "rows_checked": int(rows.size), "worst_row_rel_err": float(f"{err:.3e}")})
        ok &= err <= GATE_TOL
        if li == 0:
            continue
        t_rowptr, t_col, perm = oracle.csr_transpose(layer.rowptr, cols, layer.nrows, layer.ncols)
        crow = np.unique(np.linspace(0, layer.ncols - 1, min(GATE_ROWS, layer.ncols)).astype(np.int64))
        subt = _row_sample(t_rowptr, t_col, vals[perm], crow)
        G = gs[li].cpu().numpy()
        reft = oracle.spmm_f64acc(subt[0], subt[1], subt[2], crow.size, np.ascontiguousarray(G))
        adjs[li]._t = None
        gott = adjs[li].matmul_t(gs[li])[torch.from_numpy(crow).to(device)].cpu().numpy()
        adjs[li]._t = None
        errt = oracle.rel_err(gott, reft)[0]
        checks.append({"op": f"bwd{li}", "rows_checked": int(crow.size), "worst_row_rel_err": float(f"{errt:.3e}")})
        ok &= errt <= GATE_TOL
    if store is not None:
        nodes = torch.from_numpy(mb.input_nodes).to(device)
        x0 = store.gather(nodes).cpu().numpy()
        want = store.host_rows(mb.input_nodes)
        same = bool(np.array_equal(x0.view(np.uint32), want.view(np.uint32)))
        _, _, _, counts = store.remap(nodes)
        c = counts.cpu().numpy()
        checks.append({"op": "gather", "rows": int(nodes.numel()), "bit_exact": same,
                       "rows_by_source": {"gpu_shards": [int(v) for v in c[:store.world]], "host": int(c[store.world])}})
        ok &= same
    log("parity gate: " + json.dumps(checks))
    return {"passed": bool(ok), "tolerance": GATE_TOL, "checks": checks,
            "checker": "oracle/ (CPU restatement of the reference, fp64 accumulation), on the operands of timed minibatch 0"}


def measure_gather_roof(ext, X, colidx, flush_l2):
    """L2->SM row-gather speed of light on THIS GPU, on the dominant block's own column stream (the production
    library's probe kernel: loads only, no FMA, no stores).  Best of a few layouts; TB/s."""
    import torch
    best, detail = 0.0, {}
    Xp = X if (X.stride(0) % 4 == 0 and X.data_ptr() % 16 == 0) else X.contiguous()
    for nv, wps in [(4, 32), (4, 16), (2, 32), (1, 32), (1, 64)]:
        if Xp.shape[1] < 128 * nv:
            continue
        ts = []
        for rep in range(4):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nbytes = ext.probe_row_gather(Xp, colidx, nv, wps)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                ts.append(e0.elapsed_time(e1))
        tbps = nbytes / (float(np.median(ts)) * 1e-3) / 1e12
        detail[f"nv{nv}_w{wps}"] = round(tbps, 2)
        best = max(best, tbps)
    return best, detail


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
    ap.add_argument("--ref-gpu", action="store_true", help="(default on) time the reference's CUDA kernels from oracle/_ref")
    ap.add_argument("--no-ref-gpu", action="store_true")
    ap.add_argument("--skip-gate", action="store_true", help="profiler captures only: do not run the parity gate (the line says so)")
    ap.add_argument("--no-other-workloads", action="store_true", help="skip BASELINE configs[3]/[4] (products GCN + locality, papers sweep)")
    ap.add_argument("--papers-scale", type=int, default=16, help="papers100M-shaped graph at 1/scale of the nodes and edges")
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

    # ---- placement-partitioned feature store (needed by the gate's gather check, e2e and train)
    store = None
    if not (args.no_e2e and args.no_train):
        store = build_store(args, gmod, shape, g, device, rank, world, log)

    # ---- parity gate on the operands of timed minibatch 0, on every rank, BEFORE anything is timed
    gate = ({"passed": True, "skipped": "--skip-gate (profiler capture); not a reportable run"} if args.skip_gate
            else parity_gate(mbs[0], widths, dev_mbs[0], store, device, log))
    gate_ok = torch.tensor([1.0 if gate["passed"] else 0.0], device=device)
    if world > 1:
        dist.all_reduce(gate_ok, op=dist.ReduceOp.MIN)
    if gate_ok.item() < 1.0:
        if rank == 0 or not gate["passed"]:
            print(f"[bench r{rank}] PARITY GATE FAILED - no number is reported: {json.dumps(gate)}", file=sys.stderr, flush=True)
        if world > 1:
            dist.destroy_process_group()
        return 3
    gate["ranks_passed"] = world

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
    # L2->SM gather speed of light, measured now, on the largest forward block of timed minibatch 0
    big = max(range(nl), key=lambda li: mbs[0].layers[li].nnz * widths[li])
    try:
        gather_roof_tbps, gather_roof_detail = measure_gather_roof(ext, dev_mbs[0][1][big], dev_mbs[0][0][big].colidx, flush_l2)
        gather_roof_src = "measured in this run (gnn_probe_row_gather_f32 on the block's own column stream, cold L2)"
    except Exception as exc:                             # the headline must not die with the probe
        gather_roof_tbps, gather_roof_detail = L2_GATHER_ROOF_FALLBACK_BPS / 1e12, {"error": str(exc)[:200]}
        gather_roof_src = "fallback constant"
    ops = []
    for k, name in enumerate(op_names):
        li = int(name[3:])
        bytes_k = np.array([algorithmic_bytes(mbs[s % len(mbs)].layers[li].nnz, mbs[s % len(mbs)].layers[li].nrows,
                                              mbs[s % len(mbs)].layers[li].ncols, widths[li]) for s in range(args.steps)], dtype=np.float64)
        nnzD = np.array([mbs[s % len(mbs)].layers[li].nnz * widths[li] for s in range(args.steps)], dtype=np.float64)
        ms = op_ms[:, k]
        t_hbm = bytes_k.mean() / (hbm_peak * 1e9)
        t_l2 = 4 * nnzD.mean() / (gather_roof_tbps * 1e12)                     # gather bytes / measured pure-gather roof
        t_fma = nnzD.mean() / (148 * 128 * sm_clk_ghz * 1e9)
        ops.append({"op": name, "ms": round(float(ms.mean()), 4), "ms_median": round(float(np.median(ms)), 4),
                    "ms_max": round(float(ms.max()), 4), "share": round(float(ms.sum() / op_ms.sum()), 4),
                    "algorithmic_GBps": round(float(bytes_k.sum() / (ms.sum() * 1e-3) / 1e9), 1),
                    "frac_of_hbm_roof": round(float(bytes_k.sum() / (ms.sum() * 1e-3) / 1e9 / hbm_peak), 4),
                    "bytes": int(bytes_k.mean()), "gather_GB": round(float(4 * nnzD.mean() / 1e9), 3),
                    "gflop": round(float(2 * nnzD.mean() / 1e9), 3),
                    "t_bound_us": {"hbm": round(t_hbm * 1e6, 1), "l2_gather": round(t_l2 * 1e6, 1), "fma": round(t_fma * 1e6, 1)},
                    "frac_of_t_bound": round(float(max(t_hbm, t_l2, t_fma) / (ms.mean() * 1e-3)), 3)})
    dom = max(ops, key=lambda o: o["share"])
    roofline = {"bound": "hbm", "kernel": "spmm_rowsplit_kernel", "op": dom["op"], "achieved": dom["algorithmic_GBps"],
                "peak": hbm_peak, "unit": "GB/s", "frac": round(dom["algorithmic_GBps"] / hbm_peak, 4), "traffic": None,
                "peak_source": peak_src, "binding_roof": max(dom["t_bound_us"], key=dom["t_bound_us"].get),
                "frac_of_binding_roof": dom["frac_of_t_bound"],
                "l2_gather_roof_measured_TBps": round(gather_roof_tbps, 2), "l2_gather_roof_layouts_TBps": gather_roof_detail,
                "l2_gather_roof_source": gather_roof_src,
                "binding_roof_note": "a row-wise fp32 SpMM on a dense LADIES block re-reads every X row nnz/K (~100-160) times "
                                     "from L2; l2_gather = 4*nnz*D bytes / the gather-only speed of light measured above "
                                     "(SURVEY.md 8(d): HBM is not the binding roof of these blocks; it is for the short-row ones)"}
    prof = os.path.join(REPO, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(dom["op"])
        except (OSError, ValueError):
            pass

    # ---- end-to-end through the public API with host inputs
    e2e, train = None, None
    if not args.no_e2e:
        e2e = run_e2e(args, cso, store, mbs, widths, step_bytes, device, rank, world, log)

    # ---- training minibatches/s (full step incl. NCCL allreduce), secondary metric of BASELINE.json
    if not args.no_train:
        from gnn_b200 import harness
        # reference-shaped model first (the reference's own models.py would run exactly these torch ops), then the
        # same model with the fused ELU+row-norm epilogue of gnn_b200/models.py and the
        # flat-gradient clip + exchange
        train = harness.bench_train(args, cso, store, shape, g, mbs, ORDERS, NHID, device, rank, world, log)
        # sampler threads per GPU: the reference's default --pool_num is 4 (main.py:77); with the host cores to spare
        # (>= 2 per thread and rank) the second live number uses 8, since 4 threads x 26 ms per minibatch is sampler-bound
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
        pool_wide = 8 if cores // max(world, 1) >= 16 else 4
        for key, fn, kw in [("fused_epilogue_model", harness.bench_train, dict(fused=True)),
                            ("fused_epilogue_flat_gradients", harness.bench_train, dict(fused=True, flat_grads=True)),
                            # + the layers' dense linears (and the gather / concat around them) on tcgen05 in 3xTF32
                            ("tensor_core_linears", harness.bench_train, dict(fused=True, flat_grads=True, tc=True)),
                            ("live_sampler", harness.bench_train_live, dict(fused=False, pool_num=4)),
                            ("live_sampler_fused_epilogue", harness.bench_train_live, dict(fused=True, pool_num=pool_wide, flat_grads=True)),
                            ("live_sampler_tensor_core_linears", harness.bench_train_live,
                             dict(fused=True, pool_num=pool_wide, flat_grads=True, tc=True))]:
            try:
                if fn is harness.bench_train:
                    train[key] = fn(args, cso, store, shape, g, mbs, ORDERS, NHID, device, rank, world, log, **kw)
                else:
                    train[key] = fn(args, cso, store, shape, g, ORDERS, NHID, samp, batch, device, rank, world, log, **kw)
            except Exception as exc:                   # secondary numbers must not take the headline down
                train[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
        # one-glance summary: the reference-shaped torch model on the new ops (top level of `train`) against the fastest
        # variant of the optional fused model pieces, pre-sampled and with the sampler in the loop
        pre = {k: v["minibatches_per_s"] for k, v in train.items() if isinstance(v, dict) and "minibatches_per_s" in v and not k.startswith("live")}
        live = {k: v["minibatches_per_s"] for k, v in train.items() if isinstance(v, dict) and "minibatches_per_s" in v and k.startswith("live")}
        train["summary"] = {"reference_shaped_model": train.get("minibatches_per_s"),
                            "best_pre_sampled": max(pre.items(), key=lambda kv: kv[1]) if pre else None,
                            "best_live_sampler": max(live.items(), key=lambda kv: kv[1]) if live else None, "unit": "minibatches/s (all ranks)"}
    if store is not None:
        store.close()
        store = None

    # ---- the kernel to beat: the reference's own CUDA extension (compiled unmodified into oracle/_ref) on the same
    # blocks, same GPU, same clock (CUDA events, cold L2)
    ref_gpu = None
    if not args.no_ref_gpu and rank == 0:
        try:
            ref_gpu = run_ref_gpu(mbs, widths, device, flush_l2, step_bytes[0], total_ms_max / args.steps, log)
        except Exception as exc:
            ref_gpu = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    # ---- BASELINE configs[3] and [4]: bounded runs of the other named workloads, every rank its own minibatches
    other = None
    if not args.no_other_workloads and args.workload == "reddit":
        del dev_mbs, flush_src
        torch.cuda.empty_cache()
        other = {}
        for key, fn in [("products_gcn_locality", run_products_locality), ("papers_width_sweep", run_papers_sweep)]:
            try:
                other[key] = fn(args, cso, gmod, device, rank, world, flush_buf, log)
            except Exception as exc:
                import traceback
                log(traceback.format_exc())
                other[key] = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_baseline, _ = cpu_reference_sample(mbs[0], widths, args.cpu_seconds, log)

    if rank == 0:
        line = {
            "metric": "LADIES-layer SpMM HBM GB/s (fwd+bwd)", "value": round(value, 2), "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total_ms_max / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(shape, g, mbs, widths, samp, batch),
            "parity_gate": gate,
            "wall_ms_per_step_incl_flush": round(wall / args.steps * 1e3, 4),
            "gpu_launches": int(launches), "clocks": clk, "ops": ops, "warm_l2": warm_l2, "roofline": roofline,
            "e2e": e2e, "train": train, "cpu_baseline": cpu_baseline,
            "reference_cuda_kernels_same_gpu": ref_gpu, "other_workloads": other,
        }
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
                store.ext.gather_rows_src(xrows, src_dev, -100000 - rank, store.feat_dim, outbuf)     # GNN_SRC_PEERS(rank): one launch
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


def run_ref_gpu(mbs, widths, device, flush_l2, step_bytes, our_ms_per_step, log):
    """The reference's own CUDA kernels (oracle/_ref, built unmodified by oracle/build_ref.py) on the blocks of timed
    minibatch 0, on this GPU, clocked like `value`: CUDA events around each op, L2 flushed before each op, median of 3.
    Forward = spmm_load_balance (spmm.cpp:23-27); backward = mat1.transpose(0,1).coalesce() + spmm_load_balance, exactly
    what custom_sparse_ops.py:30-37 executes.  Used as a yardstick only."""
    import torch
    from oracle import build_ref
    mod = build_ref.load_ref()
    if mod is None:
        return {"unavailable": "oracle/_ref/spmm_ref.so not in this snapshot"}
    mb = mbs[0]
    res = {}
    tot_ms = 0.0
    for li, (layer, D) in enumerate(zip(mb.layers, widths)):
        a = mod.create_coo_tensor(torch.from_numpy(layer.fullrowptr).to(device), torch.from_numpy(layer.rowptr).to(device),
                                  torch.from_numpy(layer.colidx).to(device), torch.from_numpy(layer.normfact).to(device),
                                  layer.nrows, layer.ncols)
        x = torch.randn(layer.ncols, D, device=device)
        gq = torch.randn(layer.nrows, D, device=device)
        for name, fn in [("fwd", lambda: mod.spmm_load_balance(a, x))] + (
                [("bwd", lambda: mod.spmm_load_balance(a.transpose(0, 1).coalesce(), gq.contiguous()))] if li > 0 else []):
            fn()
            ts = []
            for _ in range(3):
                flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = float(np.median(ts))
            res[f"{name}{li}_ms"] = round(ms, 3)
            tot_ms += ms
    res["ms_per_step"] = round(tot_ms, 3)
    res["algorithmic_GBps"] = round(step_bytes / (tot_ms * 1e-3) / 1e9, 1)
    res["ours_speedup"] = round(tot_ms / our_ms_per_step, 2)
    res["note"] = ("reference spmm_load_balance (+ transpose().coalesce() in bwd) compiled unmodified for sm_100a; CUDA events, "
                   "cold L2 per op, median of 3, minibatch 0 (rank 0)")
    return res


# ----------------------------------------------------------------------------- BASELINE configs[3] / [4]
def _timed_us(fn, flush_buf, reps=3):
    import torch
    ts = []
    for r in range(reps + 1):
        flush_buf.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if r:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


def _graph_rank0_first(name, rank, world, log):
    """Rank 0 generates (and caches) the graph, the others load the cache."""
    import torch.distributed as dist
    from gnn_b200 import graphgen
    shape = graphgen.SHAPES[name]
    cache_root = os.path.join(REPO, ".cache") if os.path.isdir(os.path.join(REPO, ".cache")) else None
    if world > 1 and rank != 0:
        dist.barrier()
    t0 = time.time()
    g = graphgen.generate_cached(shape, seed=0, root=cache_root)
    log(f"graph {shape.name}: {g.num_nodes} nodes, {g.nnz} directed nnz ({time.time() - t0:.1f}s)")
    if world > 1 and rank == 0:
        dist.barrier()
    return shape, g


def _spmm_block_sweep(cso, mb, widths_per_layer, device, flush_buf, check_rows=512):
    """fwd + bwd of every layer block at the given widths: CUDA-event us (cold L2), algorithmic bytes, oracle parity on a
    row sample.  Returns (rows, total_bytes, total_us, worst_err)."""
    import torch
    import oracle
    rows_out, tot_b, tot_us, worst = [], 0.0, 0.0, 0.0
    for li, layer in enumerate(mb.layers):
        a = cso.create_coo_tensor(torch.from_numpy(layer.fullrowptr).to(device), torch.from_numpy(layer.rowptr).to(device),
                                  torch.from_numpy(layer.colidx32 if layer.ncols > 32767 else layer.colidx).to(device),
                                  torch.from_numpy(layer.normfact).to(device), layer.nrows, layer.ncols)
        adj = cso.adjacency_of(a)
        _, cols, vals = oracle.build_adj(layer.fullrowptr, layer.rowptr, layer.colidx32, layer.normfact, layer.nrows)
        cols = cols.astype(np.int32)
        rsel = np.unique(np.linspace(0, layer.nrows - 1, min(check_rows, layer.nrows)).astype(np.int64))
        sub = _row_sample(layer.rowptr, cols, vals, rsel)
        for D in widths_per_layer[li]:
            from gnn_b200 import gather as gmod
            x = torch.randn(layer.ncols, gmod.padded_ld(D), device=device)[:, :D]
            go = torch.randn(layer.nrows, D, device=device)
            B = algorithmic_bytes(layer.nnz, layer.nrows, layer.ncols, D)
            y = adj.matmul(x)
            err = oracle.rel_err(y[torch.from_numpy(rsel).to(device)].cpu().numpy(),
                                 oracle.spmm_f64acc(sub[0], sub[1], sub[2], rsel.size, np.ascontiguousarray(x.cpu().numpy())))[0]
            worst = max(worst, err)
            t_f = _timed_us(lambda: adj.matmul(x), flush_buf)

            def bwd():
                adj._t = None                     # a fresh adjacency per minibatch: whatever the backward needs is inside
                adj.matmul_t(go)
            t_b = _timed_us(bwd, flush_buf) if li > 0 else None
            rows_out.append({"layer": li, "M": layer.nrows, "K": layer.ncols, "nnz": layer.nnz, "D": D, "fwd_us": round(t_f, 1),
                             "bwd_us": (round(t_b, 1) if t_b is not None else None),
                             "fwd_GBps": round(B / t_f / 1e3, 1), "bwd_GBps": (round(B / t_b / 1e3, 1) if t_b else None),
                             "fwd_rel_err": float(f"{err:.2e}")})
            tot_b += B * (2 if li > 0 else 1)
            tot_us += t_f + (t_b or 0.0)
    return rows_out, tot_b, tot_us, worst


def _allreduce_sum_max(vals_sum, vals_max, device, world):
    import torch
    import torch.distributed as dist
    if world <= 1:
        return vals_sum, vals_max
    a = torch.tensor(vals_sum, device=device, dtype=torch.float64)
    b = torch.tensor(vals_max, device=device, dtype=torch.float64)
    dist.all_reduce(a, op=dist.ReduceOp.SUM)
    dist.all_reduce(b, op=dist.ReduceOp.MAX)
    return a.tolist(), b.tolist()


def run_products_locality(args, cso, gmod, device, rank, world, flush_buf, log):
    """BASELINE configs[3]: 3-layer GCN with --locality_sampling on the ogbn-products-shaped graph, features placed by
    the placement model (buffer 0.1 per GPU, alpha 0), every rank its own minibatches.  Reports the SpMM path (fwd+bwd
    GB/s with oracle parity), where the input rows come from with and without locality sampling (reference
    sampler.py:119-121 + preprocess.py:414-423), and GCN training minibatches/s with the device sampler in the loop."""
    import torch
    import torch.distributed as dist
    from gnn_b200 import graphgen, harness, placement, sampler
    shape, g = _graph_rank0_first("products", rank, world, log)
    orders, nhid, samp, batch, scale_factor = [1, 1, 1], 512, 8192, 512, 2.0
    devices = list(range(world))
    buffer_rows = int(args.buffer_size * g.num_nodes)
    t0 = time.time()
    lap = (g.indptr, g.indices, np.repeat(1.0 / np.maximum(g.degrees(), 1), g.degrees()))
    prob = placement.access_probability(lap, g.train_nodes, sum(orders))
    if world > 1:
        pl = placement.create_placement(lap, g.train_nodes, buffer_rows, devices, sum(orders), alpha=0.0, sample_prob=prob)
        did, idx, bufs = pl.device_id_of_nodes_group[rank], pl.idx_of_nodes_on_device_group[rank], pl.gpu_buffer_group
    else:
        top = np.argsort(-1 * prob)[:buffer_rows]
        did = np.full(g.num_nodes, -1, dtype=np.int64)
        did[top] = 0
        idx = np.arange(g.num_nodes, dtype=np.int64)
        idx[top] = np.arange(top.size)
        bufs = [top]
    skew = placement.locality_sampling_sets(g.indptr, g.indices, shape.self_loops, bufs, len(orders))
    log(f"products: placement + locality sets ({time.time() - t0:.1f}s)")
    feats = torch.from_numpy(graphgen.features(shape, seed=1))
    store = gmod.FeatureStore(feats, bufs, did, idx, devices, rank, device, group=(dist.group.WORLD if world > 1 else None))
    rng = np.random.Generator(np.random.PCG64(4000 + rank))
    chunk = (g.train_nodes.size + world - 1) // world
    own = g.train_nodes[rank * chunk: min((rank + 1) * chunk, g.train_nodes.size)]
    out = {"config": f"products-shaped 3-layer GCN (nhid {nhid}), LADIES samp_num {samp} batch {batch}, buffer_size {args.buffer_size}, "
                     f"alpha 0, locality_sampling scale_factor {scale_factor} (BASELINE configs[3])",
           "graph": {"nodes": g.num_nodes, "directed_nnz": g.nnz, "feat_dim": shape.feat_dim}}
    widths = [[shape.feat_dim], [nhid], [nhid]]
    src = {}
    for tag, sf in [("uniform", 1.0), ("locality", scale_factor)]:
        mbs = [sampler.ladies_sample(7000 + 10 * rank + i, own[rng.permutation(own.size)[:batch]], [samp] * 5, g.num_nodes, g.indptr,
                                     g.indices, orders, skewed_sampling_nodes=skew, scale_factor=sf) for i in range(2)]
        counts = np.zeros(world + 2)
        for mb in mbs:
            _, _, _, c = store.remap(torch.from_numpy(mb.input_nodes).to(device))
            counts += c.cpu().numpy()
        counts /= len(mbs)
        local, host = counts[rank], counts[world]
        peer = counts[:world].sum() - local
        rows, tot_b, tot_us, worst = _spmm_block_sweep(cso, mbs[0], widths, device, flush_buf)
        x0 = store.gather(torch.from_numpy(mbs[0].input_nodes).to(device)).cpu().numpy()
        gather_ok = bool(np.array_equal(x0.view(np.uint32), store.host_rows(mbs[0].input_nodes).view(np.uint32)))
        (b_all, us_sum), (us_max, err_max) = _allreduce_sum_max([tot_b, tot_us], [tot_us, worst], device, world)
        src[tag] = {"input_rows_per_minibatch": {"local": int(local), "peer": int(peer), "host": int(host)},
                    "spmm_fwd_bwd_GBps_all_ranks": round(b_all / us_max / 1e3, 1), "spmm_us_per_minibatch_max_rank": round(us_max, 1),
                    "blocks": rows if rank == 0 else None, "worst_row_rel_err": float(f"{err_max:.2e}"),
                    "parity_ok": bool(err_max <= GATE_TOL and gather_ok), "gather_bit_exact": gather_ok}
    out["sampling"] = src
    # GCN training with the device sampler in the loop, locality sampling on
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 4)
    pool = 8 if cores // max(world, 1) >= 16 else 4      # like the Reddit-shaped live legs: the reference's default --pool_num is 4
    try:
        targs = argparse.Namespace(steps=min(args.steps, 12))
        out["train_gcn_live_locality"] = harness.bench_train_live(targs, cso, store, shape, g, orders, nhid, samp, batch, device, rank, world, log,
                                                                   pool_num=pool, fused=True, flat_grads=True, skewed_sampling_nodes=skew,
                                                                   scale_factor=scale_factor)
    except Exception as exc:
        out["train_gcn_live_locality"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    try:                  # the same with the GCN layers' linears on the tensor cores (3xTF32)
        targs = argparse.Namespace(steps=min(args.steps, 12))
        out["train_gcn_live_locality_tensor_core_linears"] = harness.bench_train_live(
            targs, cso, store, shape, g, orders, nhid, samp, batch, device, rank, world, log, pool_num=pool, fused=True, flat_grads=True,
            skewed_sampling_nodes=skew, scale_factor=scale_factor, tc=True)
    except Exception as exc:
        out["train_gcn_live_locality_tensor_core_linears"] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    store.close()
    return out


def run_papers_sweep(args, cso, gmod, device, rank, world, flush_buf, log):
    """BASELINE configs[4]: ogbn-papers100M-shaped graph (at 1/--papers-scale of the nodes and edges - stated in the
    output; degree statistics kept), features in pinned host memory only (no GPU cache: every input row crosses PCIe),
    SpMM width sweep 16..1024 on each rank's own LADIES blocks, oracle parity at every width."""
    import torch
    import torch.distributed as dist
    from gnn_b200 import graphgen, sampler
    name = "papers16" if args.papers_scale == 16 else None
    if name is None:
        base = graphgen.SHAPES["papers16"]
        f = 16.0 / args.papers_scale
        graphgen.SHAPES[f"papers{args.papers_scale}"] = graphgen.GraphShape(f"papers{args.papers_scale}", int(base.num_nodes * f),
                                                                             int(base.num_undirected_edges * f), base.feat_dim,
                                                                             base.num_classes, base.max_degree, base.alpha, base.self_loops)
        name = f"papers{args.papers_scale}"
    shape, g = _graph_rank0_first(name, rank, world, log)
    orders, samp, batch = [1, 1, 1], 8192, 512
    rng = np.random.Generator(np.random.PCG64(9000 + rank))
    chunk = (g.train_nodes.size + world - 1) // world
    own = g.train_nodes[rank * chunk: min((rank + 1) * chunk, g.train_nodes.size)]
    mb = sampler.ladies_sample(8000 + rank, own[rng.permutation(own.size)[:batch]], [samp] * 5, g.num_nodes, g.indptr, g.indices, orders)
    sweep = [16, 32, 64, 128, 256, 512, 1024]
    rows, tot_b, tot_us, worst = _spmm_block_sweep(cso, mb, [sweep] * len(mb.layers), device, flush_buf)
    # host-resident features: the whole table pinned + mapped, no GPU shard; gather of one minibatch's input rows
    t0 = time.time()
    feats = torch.from_numpy(graphgen.features(shape, seed=1))
    did = np.full(g.num_nodes, -1, dtype=np.int64)
    store = gmod.FeatureStore(feats, [np.empty(0, dtype=np.int64) for _ in range(world)], did, np.arange(g.num_nodes, dtype=np.int64),
                              list(range(world)), rank, device, group=(dist.group.WORLD if world > 1 else None))
    log(f"papers: host table {tuple(feats.shape)} pinned ({time.time() - t0:.1f}s)")
    nodes = torch.from_numpy(mb.input_nodes).to(device)
    x0 = store.gather(nodes)
    gather_ok = bool(np.array_equal(x0.cpu().numpy().view(np.uint32), store.host_rows(mb.input_nodes).view(np.uint32)))
    src_dev, _, xrows, _ = store.remap(nodes)
    buf = torch.empty((nodes.numel(), store.ld), device=device)
    t_g = _timed_us(lambda: store.ext.gather_rows_src(xrows, src_dev, -1, store.feat_dim, buf), flush_buf)
    gbytes = nodes.numel() * store.feat_dim * 4
    store.close()
    per_width = []
    for D in sweep:
        rs = [r for r in rows if r["D"] == D]
        b = sum(algorithmic_bytes(r["nnz"], r["M"], r["K"], D) * (2 if r["bwd_us"] is not None else 1) for r in rs)
        us = sum(r["fwd_us"] + (r["bwd_us"] or 0.0) for r in rs)
        (b_all,), (us_max,) = _allreduce_sum_max([b], [us], device, world)
        per_width.append({"D": D, "fwd_bwd_GBps_all_ranks": round(b_all / us_max / 1e3, 1), "us_per_minibatch_max_rank": round(us_max, 1)})
    (_,), (err_max, tg_max) = _allreduce_sum_max([0.0], [worst, t_g], device, world)
    return {"config": f"papers100M-shaped graph at 1/{args.papers_scale} scale ({g.num_nodes} nodes, {g.nnz} directed nnz, mean degree "
                      f"{g.nnz / g.num_nodes:.1f}), 128-d features in pinned host memory, LADIES samp_num {samp} batch {batch}, width sweep "
                      "(BASELINE configs[4])",
            "width_sweep": per_width, "blocks_rank0": rows if rank == 0 else None,
            "worst_row_rel_err": float(f"{err_max:.2e}"), "parity_ok": bool(err_max <= GATE_TOL and gather_ok),
            "host_gather": {"rows": int(nodes.numel()), "bytes": int(gbytes), "us_max_rank": round(tg_max, 1),
                            "GBps_per_gpu": round(gbytes / tg_max / 1e3, 1), "bit_exact": gather_ok,
                            "note": "every input row of the minibatch read zero-copy from the pinned host table over PCIe"}}


def run_reference(args, log):
    """--impl reference: the reference's CPU path on the host cores (rank 0 only)."""
    from gnn_b200 import graphgen  # noqa: F401
    shape, g, mbs, samp, batch = build_workload(args, 0, 1, log)
    widths = layer_widths(shape.feat_dim, len(mbs[0].layers), gcn=shape.self_loops)
    cpu, times = cpu_reference_sample(mbs[0], widths, args.cpu_seconds, log, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "LADIES-layer SpMM HBM GB/s (fwd+bwd)", "value": round(cpu["value"], 4), "unit": "GB/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(float(np.mean(times)) * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(shape, g, mbs, widths, samp, batch),
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
