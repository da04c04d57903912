"""Training-step harness for the secondary BASELINE metric (training minibatches/s at 1/2/4/8 B200).

NOT part of the product path: the reference's models.py / main.py are dense torch code that must
keep working unchanged on top of ``custom_sparse_ops``.  The GPU
box does not have /root/reference, so the few dense pieces a training step needs are restated
here with the same math, only to drive the hot path in its real calling pattern:

  * the model shells of gnn_b200/models.py (parameter names = the reference's, pinned to goldens of the unmodified
    reference modules): GraphSAGE / GCN layers around ``spmm(adj, x)``, head = L2-normalise, dropout, linear
  * loss: BCE-with-logits weighted 1/batch, summed          (utils.py:129-140, sigmoid_loss default)
  * step: gather -> forward -> loss -> backward -> clip_grad_norm_(5) -> gradient exchange -> Adam
    (main.py:129-170); the exchange is ONE NCCL allreduce(SUM) of the flattened gradient - the
    reference sums, it does not average (main.py:159) - replacing threads + barrier + peer copies.
"""
from __future__ import annotations

import gc
import os
import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def bce_loss(preds, labels):
    w = torch.full((preds.shape[0], 1), 1.0 / preds.shape[0], device=preds.device)
    return F.binary_cross_entropy_with_logits(preds, labels, weight=w, reduction="sum")


def exchange_gradients(params, world):
    """One allreduce(SUM) of the flattened gradient (reference main.py:149-168 / utils.py:152-160)."""
    if world <= 1:
        return 0
    import torch.distributed as dist
    grads = [p.grad for p in params if p.grad is not None]
    flat = torch.cat([gr.reshape(-1) for gr in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    off = 0
    for gr in grads:
        n = gr.numel()
        gr.copy_(flat[off:off + n].view_as(gr))
        off += n
    return flat.numel() * 4


class FlatGradients:
    """All parameter gradients as views of ONE buffer.

    The reference clips every replica's gradient to norm 5 (main.py:146), then sums the replicas (main.py:149-168:
    threads + barrier + per-parameter peer copies).  With the gradients living in one flat buffer from the start,
    zeroing is one memset, the clip is one norm + one in-place scale over the buffer (instead of a per-parameter norm,
    a stack and a per-parameter multiply), and the exchange is one NCCL allreduce(SUM) straight on the buffer with no
    flatten / copy-back.  ``buckets`` > 1 splits the allreduce so that the optimizer can start on the first bucket while
    the next one is still in flight (not used by default: 11 MB over NVSwitch is latency-bound)."""

    def __init__(self, params, world: int, max_norm: float = 5.0, comm_stream=None):
        self.params = [p for p in params if p.requires_grad]
        self.world, self.max_norm = world, max_norm
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n
        self.bytes = total * 4

    def zero(self):
        self.flat.zero_()

    def clip_and_exchange(self):
        """clip_grad_norm_(params, max_norm) followed by the gradient sum over ranks; returns bytes exchanged."""
        norm = torch.linalg.vector_norm(self.flat)
        self.flat.mul_(torch.clamp(self.max_norm / (norm + 1e-6), max=1.0))      # same coefficient as clip_grad_norm_
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            return self.bytes
        return 0


def make_model(cso, shape, orders, nhid, device, fused, tc=False):
    """Replica of the reference's model for this shape: GraphSAGE (main.py default) or GCN for the +I shapes."""
    from . import models
    torch.manual_seed(1234)                      # same initial replica on every rank (reference main.py:91-97 builds one per thread)
    kind = "gcn" if shape.self_loops else "graphsage"
    return models.build_model(kind, shape.feat_dim, nhid, orders, shape.num_classes, dropout=0.1, fused=fused,
                              spmm=cso.spmm, tc=tc).to(device), kind


def bench_train(args, cso, store, shape, g, mbs, orders, nhid, device, rank, world, log, fused=False, flat_grads=False, tc=False):
    """Full training steps over the rotated pre-sampled minibatches (sampling excluded, as stated in the line)."""
    import torch.distributed as dist
    from . import graphgen
    model, kind = make_model(cso, shape, orders, nhid, device, fused, tc)
    params = [p for p in model.parameters() if p.requires_grad]
    flat = FlatGradients(params, world) if flat_grads else None
    opt = torch.optim.Adam(params, lr=0.01, fused=bool(flat_grads))
    labels_all = graphgen.labels(shape, seed=3)
    prepared = []
    for mb in mbs:
        adjs = [cso.create_coo_tensor(torch.from_numpy(l.fullrowptr).to(device), torch.from_numpy(l.rowptr).to(device),
                                      torch.from_numpy(l.colidx).to(device), torch.from_numpy(l.normfact).to(device),
                                      l.nrows, l.ncols) for l in mb.layers]
        sn = [torch.from_numpy(np.ascontiguousarray(s, dtype=np.int64)).to(device) for s in mb.sampled_nodes]
        y = F.one_hot(torch.from_numpy(labels_all[mb.batch_nodes]), shape.num_classes).float().to(device)
        prepared.append((adjs, sn, y, torch.from_numpy(mb.input_nodes).to(device)))

    model.train()
    comm_bytes = 0
    # input features of minibatch i+1 are gathered on a side stream while minibatch i computes (the reference
    # prefetches whole minibatches in sampler threads; its gather itself is synchronous, main.py:129-137)
    side = store.side_stream()
    pending = {}

    def step(i):
        nonlocal comm_bytes
        adjs, sn, y, nodes = prepared[i % len(prepared)]
        for a in adjs:
            cso.adjacency_of(a)._t = None        # a fresh adjacency every minibatch: backward rebuilds its A^T index
        if flat is not None:
            flat.zero()
        else:
            opt.zero_grad(set_to_none=False)
        if i in pending:
            x0, ev = pending.pop(i)
            torch.cuda.current_stream().wait_event(ev)
            x0.record_stream(torch.cuda.current_stream())
        else:
            x0 = store.gather(nodes)
        side.wait_stream(torch.cuda.current_stream())
        pending[i + 1] = store.prefetch(prepared[(i + 1) % len(prepared)][3], side)
        out = model(x0, adjs, sn)
        loss = bce_loss(out, y)
        loss.backward()
        if flat is not None:
            comm_bytes = flat.clip_and_exchange()
        else:
            torch.nn.utils.clip_grad_norm_(params, 5)
            comm_bytes = exchange_gradients(params, world)
        opt.step()
        return loss

    steps = max(2, min(args.steps, 20))
    co_token = store.begin_co_running()          # the prefetched host-row gather runs beside the step's SpMMs
    for i in range(6):               # allocator pools, cuBLAS workspaces and autotuned plans settle (3 left 11 ms steps in a cold process)
        step(i)
    pending.clear()
    torch.cuda.synchronize()

    torch.cuda.synchronize()
    gc.collect()
    gc.disable()              # a full collection of a torch process takes 5-12 ms; training scripts freeze/disable it too
    if world > 1:
        dist.barrier()
    # A cudaMalloc inside the timed steps (the caching allocator still growing: each one stalls every stream for tens of
    # ms) means the loop has not reached its steady state; every rank then times the same number of steps once more, and
    # the line says so.
    repeats = 0
    step_base = 0
    while True:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        mallocs0 = torch.cuda.memory_stats(device).get("num_device_alloc", 0)
        t0 = time.perf_counter()
        ev0.record()
        for s in range(steps):
            loss = step(step_base + s)
        ev1.record()
        last = float(loss.item())
        torch.cuda.synchronize()
        mallocs = int(torch.cuda.memory_stats(device).get("num_device_alloc", 0) - mallocs0)
        again = torch.tensor([1 if (mallocs > 0 and repeats == 0) else 0], device=device)
        if world > 1:
            dist.all_reduce(again, op=dist.ReduceOp.MAX)
            dist.barrier()
        wall = time.perf_counter() - t0
        if int(again.item()) == 0:
            break
        repeats += 1
        step_base += steps
    pending.clear()
    gc.enable()
    ms = ev0.elapsed_time(ev1)
    store.end_co_running(co_token)
    if world > 1:
        t = torch.tensor([ms, wall * 1e3], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, wall = float(t[0].item()), float(t[1].item()) / 1e3
    nparams = sum(p.numel() for p in params)
    return {"minibatches_per_s": round(world * steps / max(ms * 1e-3, wall), 2), "unit": "minibatches/s", "steps": steps,
            "ms_per_step_device": round(ms / steps, 3), "ms_per_step_wall": round(wall / steps * 1e3, 3),
            "allreduce_bytes_per_step": int(comm_bytes), "parameters": int(nparams), "final_loss": round(last, 4),
            "fused_epilogue": bool(fused), "flat_gradients": bool(flat_grads), "tensor_core_linears": bool(tc), "model": kind,
            "cuda_mallocs_in_timed_region": mallocs, "timed_region_repeated_after_allocator_growth": bool(repeats),
            "note": f"gather (next minibatch prefetched on a side stream) + {kind} fwd + BCE loss + bwd + clip + NCCL "
                    "allreduce(sum) + Adam on pre-sampled minibatches (host LADIES sampling and adjacency upload excluded)"}


_SAMPLER_STREAMS = {}


def _sampler_stream(device, priority, index, pipeline):
    """The ``index``-th sampler stream of a device, created once per process with its allocator pool sized up front
    (pipeline.reserve_stream_pool) and reused by every later run: torch hands out streams from a pool of 32 per priority
    and wraps around, and every fresh stream would reserve its gigabyte again."""
    key = (str(device), int(priority), int(index))
    st = _SAMPLER_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=device, priority=priority)
        pipeline.reserve_stream_pool(st, 1 << 30)     # no cudaMalloc in the loop (pipeline.py)
        _SAMPLER_STREAMS[key] = st
    return st


def bench_train_live(args, cso, store, shape, g, orders, nhid, samp, batch, device, rank, world, log, pool_num=4, fused=False,
                     prebuild_transpose=True, flat_grads=False, skewed_sampling_nodes=None, scale_factor=1.0, tc=False,
                     sampler_stream_priority=0, co_split=True):
    """Training with the sampler IN the loop (BASELINE's second minibatches/s number): ``pool_num`` sampler threads
    (reference main.py:77 uses a ThreadPoolExecutor of --pool_num=4 per GPU) run the device LADIES sampler
    (gnn_b200/gpu_sampler.py: native legacy draw on the host, array passes on the GPU) and the feature gather on their own
    CUDA streams, a bounded queue feeds the training stream."""
    import collections
    import threading
    from concurrent.futures import ThreadPoolExecutor
    import torch.distributed as dist
    from . import gpu_sampler, graphgen, pipeline
    model, kind = make_model(cso, shape, orders, nhid, device, fused, tc)
    params = [p for p in model.parameters() if p.requires_grad]
    flat = FlatGradients(params, world) if flat_grads else None
    opt = torch.optim.Adam(params, lr=0.01, fused=bool(flat_grads))
    labels_all = graphgen.labels(shape, seed=3)
    dg = gpu_sampler.DeviceGraph(g.indptr, g.indices, device)
    tls = threading.local()
    main_stream = torch.cuda.current_stream(device)
    # sustained rate: the timed region is several times the queue depth, so that neither a full queue at its start nor an
    # empty one can carry the figure
    depth = 2 * pool_num
    steps = max(args.steps, 6 * depth)
    # every sampler thread owns a stream, and the caching allocator keeps one block pool per stream: until each pool
    # has seen the largest adjacency it will hold, jobs hit cudaMalloc (device-synchronising).  Warm up long enough
    # for that (untimed, like any production run's first steps); measured: 4 steps left runs at 19-52 ms/step, then
    # the same code settles at 7.2 ms/step.
    warm = 8 * pool_num
    total = 2 * steps + warm                          # the timed region may run twice (see below)
    rng = np.random.Generator(np.random.PCG64(77 + rank))
    chunk = (g.train_nodes.size + world - 1) // world
    own = g.train_nodes[rank * chunk: min((rank + 1) * chunk, g.train_nodes.size)]
    batches = [own[rng.permutation(own.size)[:batch]] for _ in range(total)]

    import itertools
    thread_ids = itertools.count()                    # (next() on a count is atomic under the GIL)
    job_s = [0.0, 0]                                  # seconds spent inside sampler jobs, jobs finished (timed region only)
    wait_s = [0.0]                                    # seconds the training thread waited for a minibatch

    def job(i):
        t_job = time.perf_counter()
        torch.cuda.set_device(device)
        if not hasattr(tls, "stream"):
            tls.stream = _sampler_stream(device, sampler_stream_priority, next(thread_ids), pipeline)
            tls.scratch = dg.scratch()
        with torch.cuda.stream(tls.stream):
            mb = gpu_sampler.ladies_sample_device(5000 + 1000 * rank + i, batches[i], [samp] * 5, dg, orders,
                                                  create_coo_tensor=cso.create_coo_tensor, scratch=tls.scratch,
                                                  prebuild_transpose=prebuild_transpose,
                                                  skewed_sampling_nodes=skewed_sampling_nodes, scale_factor=scale_factor)
            # pinned staging for every upload: a pageable source makes the copy synchronise the stream first
            nodes = gpu_sampler.h2d(mb.input_nodes, device)
            x0 = store.gather_co_running(nodes) if co_split else store.gather(nodes)
            sn = [gpu_sampler.h2d(np.ascontiguousarray(s_, dtype=np.int64), device) for s_ in mb.sampled_nodes]
            y = F.one_hot(torch.from_numpy(labels_all[mb.batch_nodes]), shape.num_classes).float().pin_memory().to(device, non_blocking=True)
        tls.stream.synchronize()
        gpu_sampler.record_stream(mb, main_stream)
        for t in [x0, y] + sn:
            t.record_stream(main_stream)
        job_s[0] += time.perf_counter() - t_job       # (unsynchronised adds of a diagnostic: good to a few percent)
        job_s[1] += 1
        return mb, x0, sn, y

    # GNN_B200_SYNC_MODE: how host threads wait in a stream synchronise (0 spin = CUDA default, 1 blocking, 2 spin + yield);
    # an experiment switch - see profiles/README.md for what was measured
    sync_mode = os.environ.get("GNN_B200_SYNC_MODE")
    if sync_mode is not None:
        cso.spmm_cpp.set_blocking_sync(int(sync_mode))
    model.train()
    pool = ThreadPoolExecutor(max_workers=pool_num)
    pending = collections.deque()
    nxt = 0

    def refill():
        nonlocal nxt
        while len(pending) < depth and nxt < total:
            pending.append(pool.submit(job, nxt))
            nxt += 1

    def step():
        refill()
        t_wait = time.perf_counter()
        mb, x0, sn, y = pending.popleft().result()
        wait_s[0] += time.perf_counter() - t_wait
        refill()
        if flat is not None:
            flat.zero()
        else:
            opt.zero_grad(set_to_none=False)
        out = model(x0, mb.adjs, sn)
        loss = bce_loss(out, y)
        loss.backward()
        if flat is not None:
            flat.clip_and_exchange()
        else:
            torch.nn.utils.clip_grad_norm_(params, 5)
            exchange_gradients(params, world)
        opt.step()
        return loss

    co_token = store.begin_co_running()          # sampler threads gather host rows beside the step's SpMMs
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()
    if world > 1:
        dist.barrier()
    # like bench_train: a cudaMalloc inside the timed steps (caching allocator still growing - every call stalls all
    # streams) means the loop is not in its steady state yet; all ranks then time the same number of steps once more
    repeats = 0
    while True:
        job_s[0], job_s[1], wait_s[0] = 0.0, 0, 0.0
        mallocs0 = torch.cuda.memory_stats(device).get("num_device_alloc", 0)
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = step()
        last = float(loss.item())
        torch.cuda.synchronize()
        mallocs = int(torch.cuda.memory_stats(device).get("num_device_alloc", 0) - mallocs0)
        again = torch.tensor([1 if (mallocs > 0 and repeats == 0) else 0], device=device)
        if world > 1:
            dist.all_reduce(again, op=dist.ReduceOp.MAX)
            dist.barrier()
        wall = time.perf_counter() - t0
        if int(again.item()) == 0:
            break
        repeats += 1
    gc.enable()
    pool.shutdown(wait=True)
    store.end_co_running(co_token)
    if world > 1:
        t = torch.tensor([wall], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall = float(t.item())
    return {"minibatches_per_s": round(world * steps / wall, 2), "unit": "minibatches/s", "steps": steps,
            "ms_per_step_wall": round(wall / steps * 1e3, 3), "sampler_threads": pool_num, "warmup_steps": warm, "final_loss": round(last, 4),
            "queue_depth": depth, "sampler_job_ms": round(job_s[0] / max(job_s[1], 1) * 1e3, 2),
            "trainer_wait_ms_per_step": round(wait_s[0] / steps * 1e3, 3),
            "cuda_mallocs_in_timed_region": mallocs, "timed_region_repeated_after_allocator_growth": bool(repeats),
            "fused_epilogue": bool(fused), "flat_gradients": bool(flat_grads), "tensor_core_linears": bool(tc), "model": kind,
            "scale_factor": float(scale_factor),
            "note": "live LADIES sampling: legacy weighted draw on the host (native, same MT19937 stream) + device array passes "
                    "(bit-identical sampled sets), gather in the sampler threads, then the same training step; wall clock "
                    "incl. sampling, timed region = 6x the queue depth"}
