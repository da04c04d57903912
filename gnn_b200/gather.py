"""Placement-partitioned feature buffers and the per-minibatch input gather.

Host-side mirror of the reference's gather block (main.py:129-134, copies at
:185-190 and :228-233) and of the buffers ``create_buffer`` uploads
(preprocess.py:397-399):

    input_feat_data[mask_i]   = gpu_buffers[i][slots_i].to(device)     for every GPU i
    input_feat_data[mask_cpu] = feat_data[ids_cpu].to(device)

Here the per-GPU buffers are *shards* allocated through the C ABI so that every
rank (one process per GPU) can map its peers' shards over NVLink (CUDA IPC), the
host table is pinned and mapped for zero-copy reads, the placement tables live
on the device, and one remap kernel + one gather kernel on the consuming GPU
replace the mask/scatter sequence.  Row j of the result belongs to
``input_nodes[j]`` exactly as in the reference.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _native


def padded_ld(feat_dim: int) -> int:
    """Rows of shards / gathered buffers start on 128-byte lines (ld = multiple of 32 floats): float4 loads for any F,
    and no X row of the first SpMM straddles a cache line it does not own (measured -10 % on the F=602 block)."""
    return (int(feat_dim) + 31) // 32 * 32


class FeatureStore:
    def __init__(self, feat_data: torch.Tensor, gpu_buffer_nodes: Sequence[np.ndarray], device_id_of_nodes: np.ndarray,
                 idx_of_nodes_on_device: np.ndarray, devices: Sequence[int], rank: int, device: torch.device,
                 group=None, map_host: bool = True, pad_host: bool = True):
        """feat_data: CPU float32 [N, F] (the reference's ``feat_data``); gpu_buffer_nodes[i]: node id of every
        slot of GPU i's buffer (``gpu_buffer_group``); device_id_of_nodes / idx_of_nodes_on_device: THIS rank's
        view of the placement tables; devices: device ids as they appear in the tables; ``group``: a
        torch.distributed process group when there is one process per GPU (None = single process).
        ``pad_host=False`` registers ``feat_data`` as it is (leading dimension F) instead of copying it into a table
        with 128-byte-aligned rows: no second copy of a papers100M-sized table in host RAM, at the price of narrower
        loads on the (PCIe-bound) host rows.  ``map_host=False``: no host table at all - a minibatch that needs an
        uncached node then raises instead of reading through a NULL base."""
        self.ext = _native.extension()
        self.rank, self.world = int(rank), len(devices)
        self.device = torch.device(device)
        self.feat_dim = int(feat_data.shape[1])
        self.ld = padded_ld(self.feat_dim)
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()

        # ---- host table: padded, pinned, mapped (zero-copy reads of uncached rows, main.py:134)
        self.host = None
        self._feat = feat_data                  # the caller's table (not copied): reference rows for checks
        self.ld_host = self.ld
        host_alias = 0
        if map_host:
            if (feat_data.shape[1] == self.ld or not pad_host) and feat_data.is_contiguous():
                self.host = feat_data
                self.ld_host = int(feat_data.shape[1])
            else:
                self.host = torch.zeros((feat_data.shape[0], self.ld), dtype=torch.float32)
                self.host[:, :self.feat_dim] = feat_data
            host_alias = self.ext.host_register(self.host)

        # ---- shards
        self.shards: List[Optional[torch.Tensor]] = [None] * self.world
        self._handles = None
        multi_process = group is not None and self.world > 1
        if multi_process:
            import torch.distributed as dist
            nodes = np.asarray(gpu_buffer_nodes[self.rank])
            shard, handle = self.ext.shard_alloc(max(len(nodes), 1), self.ld, dev_index)      # an empty buffer still maps
            self._fill(shard, feat_data, nodes)
            self.shards[self.rank] = shard
            handles = [None] * self.world
            dist.all_gather_object(handles, (bytes(handle), int(len(nodes))), group=group)
            torch.cuda.synchronize(self.device)
            dist.barrier(group=group)
            for i, (h, rows) in enumerate(handles):
                if i != self.rank:
                    self.shards[i] = self.ext.shard_open(h, max(rows, 1), self.ld, dev_index)
        else:
            # single process: every buffer lives on this device (world == 1, or a test emulating more ranks)
            for i in range(self.world):
                nodes = np.asarray(gpu_buffer_nodes[i])
                shard, _ = self.ext.shard_alloc(max(len(nodes), 1), self.ld, dev_index)
                self._fill(shard, feat_data, nodes)
                self.shards[i] = shard

        ptrs = [int(s.data_ptr()) for s in self.shards] + [int(host_alias)]
        self.bases = torch.tensor(ptrs, dtype=torch.int64, device=self.device)
        self.devices = torch.tensor([int(d) for d in devices], dtype=torch.int64, device=self.device)
        self.device_id_of_nodes = torch.from_numpy(np.ascontiguousarray(device_id_of_nodes, dtype=np.int64)).to(self.device)
        self.idx_of_nodes_on_device = torch.from_numpy(np.ascontiguousarray(idx_of_nodes_on_device, dtype=np.int64)).to(self.device)

    def _fill(self, shard: torch.Tensor, feat_data: torch.Tensor, nodes: np.ndarray):
        if len(nodes) == 0:
            return
        shard.zero_()
        rows = feat_data[torch.from_numpy(np.ascontiguousarray(nodes, dtype=np.int64))]
        shard[:len(nodes), :self.feat_dim].copy_(rows.to(self.device, non_blocking=False))

    # ------------------------------------------------------------------
    def remap(self, input_nodes: torch.Tensor):
        """Device placement remap (reference sampler.py:150-158): -> (src_dev i32, slot i64, xrows i64, counts i64)."""
        out = self.ext.placement_remap(input_nodes, self.device_id_of_nodes, self.idx_of_nodes_on_device, self.devices,
                                       self.bases, self.ld, self.ld_host)
        return out

    def host_rows(self, nodes) -> np.ndarray:
        """``feat_data[nodes]`` from the caller's own table (what the reference's gather must reproduce bit for bit)."""
        return self._feat[torch.as_tensor(np.asarray(nodes), dtype=torch.int64)].numpy()

    def check_all_resident(self, counts: torch.Tensor) -> None:
        """With ``map_host=False`` a node that no GPU caches has nowhere to come from: fail loudly."""
        if self.host is None and int(counts[self.world].item()) > 0:
            raise RuntimeError(f"{int(counts[self.world].item())} input nodes are not cached on any GPU and the store has no host table")

    def gather(self, input_nodes: torch.Tensor) -> torch.Tensor:
        """input_feat_data of main.py:129-134 as a [n0, F] view of a 16-byte-row-aligned buffer."""
        _, _, xrows, counts = self.remap(input_nodes)
        if self.host is None:
            self.check_all_resident(counts)
        return self.ext.gather_rows(xrows, self.feat_dim, self.ld)

    def gather_co_running(self, input_nodes: torch.Tensor) -> torch.Tensor:
        """:meth:`gather` for callers that run NEXT TO the training step's kernels (sampler threads): rows in HBM / on
        peers by a full-size launch, host rows by the small PCIe-bound grid that :meth:`begin_co_running` reserved -
        the same two launches :meth:`prefetch` issues, on the current stream.  Same bytes as :meth:`gather`."""
        if self.host is None:
            return self.gather(input_nodes)
        src_dev, _, xrows, _ = self.remap(input_nodes)
        buf = torch.empty((input_nodes.numel(), self.ld), dtype=torch.float32, device=self.device)
        self.ext.gather_rows_src(xrows, src_dev, -100, self.feat_dim, buf)      # GNN_SRC_DEVICES: HBM / NVLink rows
        self.ext.gather_rows_src(xrows, src_dev, -1, self.feat_dim, buf)        # host rows: small grid, PCIe-bound
        return buf[:, :self.feat_dim]

    def begin_co_running(self):
        """Declare that host-row gathers of this store will run on side streams NEXT TO SpMMs (prefetch of the next
        minibatch): the SpMM planner then keeps the gather's CTA slots out of its one-wave grid
        (include/gnn_b200.h, gnn_set_corunner_ctas).  Returns a token for :meth:`end_co_running`."""
        if self.host is None:
            return None
        return self.ext.set_corunner_ctas(self.ext.host_gather_ctas())

    def end_co_running(self, token) -> None:
        if token is not None:
            self.ext.set_corunner_ctas(token)

    def side_stream(self) -> "torch.cuda.Stream":
        """The store's own prefetch stream, created once, with its allocator pool sized up front: a fresh stream
        starts with an empty pool, and the ``cudaMalloc`` calls that fill it were the 20-160 ms stalls seen in
        training loops that made a new side stream per run (pipeline.reserve_stream_pool)."""
        if getattr(self, "_side", None) is None:
            from .pipeline import reserve_stream_pool
            self._side = torch.cuda.Stream(device=self.device)
            reserve_stream_pool(self._side, 8 * 4 * self.ld * 32768)      # eight Reddit-sized input blocks
        return self._side

    def prefetch(self, input_nodes: torch.Tensor, stream: "Optional[torch.cuda.Stream]" = None):
        """Issue remap + gather for a FUTURE minibatch on ``stream`` (default: :meth:`side_stream`; host rows cross
        PCIe while the current step computes).  Returns (buffer view, event); wait on the event before the first use."""
        stream = stream or self.side_stream()
        with torch.cuda.stream(stream):
            src_dev, _, xrows, _ = self.remap(input_nodes)
            buf = torch.empty((input_nodes.numel(), self.ld), dtype=torch.float32, device=self.device)
            self.ext.gather_rows_src(xrows, src_dev, -100, self.feat_dim, buf)      # GNN_SRC_DEVICES: HBM / NVLink rows
            self.ext.gather_rows_src(xrows, src_dev, -1, self.feat_dim, buf)        # host rows: small grid, PCIe-bound
            ev = torch.cuda.Event()
            ev.record(stream)
        return buf[:, :self.feat_dim], ev

    def gather_spmm(self, adjacency, input_nodes: torch.Tensor) -> torch.Tensor:
        """Fused gather + first-layer SpMM (main.py:129-134 + models.py:18): rows in the LOCAL shard are read by the
        SpMM straight from the shard through the pointer table; rows held by peers or the host are first staged
        once each (never per nonzero - a LADIES block re-reads every X row ~100x) and the table is pointed at
        the staging rows.  ``adjacency`` is a custom_sparse_ops.Adjacency whose columns are ``input_nodes``."""
        src_dev, _, xrows, _ = self.remap(input_nodes)
        n0 = input_nodes.numel()
        stage = torch.empty((n0, self.ld), dtype=torch.float32, device=self.device)
        self.ext.gather_rows_src(xrows, src_dev, -200 - self.rank, self.feat_dim, stage)       # GNN_SRC_NOT(rank)
        staged = stage.data_ptr() + torch.arange(n0, device=self.device, dtype=torch.int64) * (self.ld * 4)
        table = torch.where(src_dev == self.rank, xrows, staged)
        out = adjacency.gather_matmul(table, self.feat_dim)
        stage.record_stream(torch.cuda.current_stream())
        return out

    def gather_from_reference_tuple(self, masks_on_devices, mask_on_cpu, idx_on_devices, idx_on_cpu, num_input_nodes):
        """Same gather driven by the reference sampler's own outputs (sampler.py:160 tuple), for main.py drop-in use:
        the masks / slot lists are folded into a pointer table on the host, then one gather kernel runs."""
        slot = np.zeros(num_input_nodes, dtype=np.int64)
        src = np.full(num_input_nodes, self.world, dtype=np.int64)      # index into bases; world = host
        for i in range(self.world):
            m = np.asarray(masks_on_devices[i])
            slot[m] = np.asarray(idx_on_devices[i])
            src[m] = i
        mc = np.asarray(mask_on_cpu)
        slot[mc] = np.asarray(idx_on_cpu)
        bases = self.bases.cpu().numpy()
        xrows = torch.from_numpy(bases[src] + slot * (self.ld * 4)).to(self.device)
        return self.ext.gather_rows(xrows, self.feat_dim, self.ld)

    def close(self):
        if self.host is not None:
            try:
                self.ext.host_unregister(self.host)
            except RuntimeError:
                pass
            self.host = None
        self.shards = []
