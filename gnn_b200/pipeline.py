"""Host -> device minibatch pipeline (the reference's sampler threads, reference sampler.py:135-139 + main.py:118-134).

In the reference every sampler job uploads its CSR pieces (``torch.from_numpy(...).to(device)``) and calls
``create_coo_tensor`` inside a pool thread, so uploads and adjacency construction overlap with training; only the
feature gather is synchronous in the training loop.  ``DevicePrefetcher`` keeps that shape with one worker thread and
one side stream per instance: for each host minibatch it copies the pinned CSR arrays, builds the adjacencies, runs
the placement remap + feature gather, and hands the training stream an event to wait on.  PCIe traffic of minibatch
i+1 (CSR arrays and zero-copy host feature rows) therefore overlaps with the SpMMs of minibatch i.
"""
from __future__ import annotations

import queue
import threading
from typing import Optional

import numpy as np
import torch


class PinnedMinibatch:
    """A sampler.Minibatch with its hand-off arrays in pinned host memory (what a sampler thread would own)."""
    def __init__(self, mb):
        self.mb = mb
        self.layers = []
        for layer in mb.layers:
            if layer is None:
                self.layers.append(None)
                continue
            self.layers.append(tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                                     for a in (layer.fullrowptr, layer.rowptr, layer.colidx, layer.normfact)))
        self.input_nodes = torch.from_numpy(np.ascontiguousarray(mb.input_nodes, dtype=np.int64)).pin_memory()

    def h2d_bytes(self) -> int:
        n = self.input_nodes.numel() * 8
        for t4 in self.layers:
            if t4 is not None:
                n += sum(t.numel() * t.element_size() for t in t4)
        return n


def reserve_stream_pool(stream: "torch.cuda.Stream", nbytes: int, small_blocks: int = 64) -> None:
    """Put one ``nbytes`` block into the caching allocator's pool of ``stream``.

    PyTorch keeps one block pool per stream, and a pool only grows through ``cudaMalloc`` - 1.7-2.9 ms per call for the
    60-70 MB pieces of a Reddit-shaped minibatch (torch.profiler trace of bench.py's e2e leg), paid by the first
    dozen minibatches of every side stream until the pool covers the pipeline depth.  With 180 GB of HBM the pool is
    simply sized up front: the block is allocated and freed here, stays cached for this stream, and later requests
    split it.  Call once per side stream, before the loop.  Size it generously (DevicePrefetcher: 8 GiB for ~0.2 GB
    minibatches): blocks handed to another stream come back only after that stream's recorded events completed, so
    several minibatches beyond the queue depth are outstanding, and ONE late ``cudaMalloc`` on a side stream stalls
    the launches of every other thread for its duration (87 ms observed with a 2 GiB pool)."""
    if nbytes <= 0:
        return
    free, _ = torch.cuda.mem_get_info(stream.device)
    nbytes = min(int(nbytes), free // 4)            # never more than a quarter of what is free right now
    with torch.cuda.stream(stream):
        block = torch.empty(int(nbytes), dtype=torch.uint8, device=stream.device)
        # requests up to 1 MiB (index arrays, counters) are served from a separate pool of 2 MiB segments
        small = [torch.empty((1 << 20) - 512, dtype=torch.uint8, device=stream.device) for _ in range(small_blocks)]
        del block, small


class DevicePrefetcher:
    def __init__(self, store, create_coo_tensor, device, depth: int = 2, prebuild_transpose: bool = False,
                 reserve_bytes: int = 8 << 30):
        self.store, self.create = store, create_coo_tensor
        self.device = torch.device(device)
        self.depth = depth
        self.prebuild_transpose = prebuild_transpose
        # two side streams: the feature gather (PCIe-bound on its host rows, ~1 ms for a Reddit-shaped minibatch) and
        # the CSR uploads + adjacency builds are independent, so they run side by side instead of queueing behind each
        # other (e2e step 2.13 -> 2.02 ms when this was the only change; DESIGN.md 8a has the rest)
        self.stream = torch.cuda.Stream(device=self.device)
        self.gather_stream = torch.cuda.Stream(device=self.device)
        reserve_stream_pool(self.stream, reserve_bytes)
        reserve_stream_pool(self.gather_stream, reserve_bytes // 4)
        # the host-row gather of minibatch i+1 is resident for most of step i: tell the SpMM planner to leave its
        # CTA slots out of the one-wave fit (include/gnn_b200.h, gnn_set_corunner_ctas)
        self._prev_corunner = store.begin_co_running()
        self._in: "queue.Queue" = queue.Queue()
        self._out: "queue.Queue" = queue.Queue(maxsize=depth)
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def submit(self, pinned: Optional[PinnedMinibatch]):
        self._in.put(pinned)

    def _run(self):
        torch.cuda.set_device(self.device)
        while True:
            item = self._in.get()
            if item is None:
                self._out.put(None)
                return
            try:
                self._out.put(self._build(item))
            except Exception as exc:  # noqa: BLE001  surfaced in get()
                self._out.put(exc)

    def _build(self, pm: PinnedMinibatch):
        dev = self.device
        with torch.cuda.stream(self.gather_stream):
            nodes = pm.input_nodes.to(dev, non_blocking=True)
            src_dev, _, xrows, counts = self.store.remap(nodes)
            buf = torch.empty((nodes.numel(), self.store.ld), dtype=torch.float32, device=dev)
            self.store.ext.gather_rows_src(xrows, src_dev, -100, self.store.feat_dim, buf)   # HBM / NVLink rows
            self.store.ext.gather_rows_src(xrows, src_dev, -1, self.store.feat_dim, buf)     # host rows (PCIe)
            gathered = torch.cuda.Event()
            gathered.record(self.gather_stream)
        with torch.cuda.stream(self.stream):
            adjs = []
            for layer, t4 in zip(pm.mb.layers, pm.layers):
                if t4 is None:
                    adjs.append(None)
                    continue
                frp, rp, ci, nf = (t.to(dev, non_blocking=True) for t in t4)
                adjs.append(self.create(frp, rp, ci, nf, layer.nrows, layer.ncols))
            if self.prebuild_transpose:
                from .custom_sparse_ops import adjacency_of
                for a in adjs[1:]:
                    if a is not None:
                        adjacency_of(a).transpose()
            self.stream.wait_event(gathered)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        return adjs, buf[:, :self.store.feat_dim], counts, ev

    def get(self):
        """Next device-ready minibatch; makes the CURRENT stream wait for its uploads/gather."""
        item = self._out.get()
        if isinstance(item, Exception):
            raise item
        if item is None:
            return None
        adjs, x0, counts, ev = item
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        from .custom_sparse_ops import adjacency_of
        tensors = [x0, counts]
        for a in adjs:
            if a is not None:
                tensors += [a._indices(), a._values()] + adjacency_of(a).device_tensors()
        for t in tensors:
            t.record_stream(cur)
        return adjs, x0, counts

    def close(self):
        self._in.put(None)
        self._thread.join(timeout=10)
        self.store.end_co_running(self._prev_corunner)
        self._prev_corunner = None
