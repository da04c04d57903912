#!/usr/bin/env python
"""NVLink data counters around the in-step peer gather (torchrun, one rank per GPU): every rank pulls its peers' rows of
real Reddit-shaped minibatches in ONE launch (GNN_SRC_PEERS), R times; rank 0 reads the NVLink RX/TX data counters of
its GPU (NVML field values, KiB) before and after and compares them with the bytes the kernels requested.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/nvlink_counters.py > gpurun_out/nvlink_n8.json
"""
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
import custom_sparse_ops as cso  # noqa: E402
from gnn_b200 import gather as gmod  # noqa: E402


def nvml_counters(index):
    """-> (rx_KiB, tx_KiB, how) summed over the links of GPU `index`, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ids = [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF), (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF)]
        vals = pynvml.nvmlDeviceGetFieldValues(h, ids)
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal))
        return out[0], out[1], "NVML field values NVLINK_THROUGHPUT_DATA_RX/TX (all links, KiB)"
    except Exception as exc:  # noqa: BLE001
        return None


def smi_counters(index):
    try:
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        rx = sum(int(l.split(":")[-1].split()[0]) for l in txt.splitlines() if "Data Rx" in l)
        tx = sum(int(l.split(":")[-1].split()[0]) for l in txt.splitlines() if "Data Tx" in l)
        return (rx, tx, "nvidia-smi nvlink -gt d (sum over links, KiB)") if (rx or tx) else None
    except Exception:  # noqa: BLE001
        return None


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)

    class A:
        workload, minibatches, buffer_size = "reddit", 3, 0.1
    log = lambda m: None
    if rank != 0:
        dist.barrier()
    shape, g, mbs, samp, batch = bench.build_workload(A, rank, world, log)
    if rank == 0:
        dist.barrier()
    store = bench.build_store(A, gmod, shape, g, device, rank, world, log)
    reps = 200
    prepared = []
    for mb in mbs:
        nodes = torch.from_numpy(mb.input_nodes).to(device)
        src_dev, slot, xrows, counts = store.remap(nodes)
        prepared.append((xrows, src_dev, torch.empty((nodes.numel(), store.ld), device=device), counts.cpu().numpy()))
    peer_rows = float(np.mean([sum(c[i] for i in range(world) if i != rank) for _, _, _, c in prepared]))
    for xrows, src_dev, out, _ in prepared:
        store.ext.gather_rows_src(xrows, src_dev, -100000 - rank, store.feat_dim, out)
    torch.cuda.synchronize()
    dist.barrier()
    read = nvml_counters if nvml_counters(local) else smi_counters
    c0 = read(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        xrows, src_dev, out, _ = prepared[r % len(prepared)]
        store.ext.gather_rows_src(xrows, src_dev, -100000 - rank, store.feat_dim, out)        # GNN_SRC_PEERS(rank), one launch
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    dist.barrier()
    time.sleep(0.5)
    c1 = read(local) if rank == 0 else None
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    b = torch.tensor([peer_rows * store.feat_dim * 4 * reps], device=device, dtype=torch.float64)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    if rank == 0:
        mine = peer_rows * store.feat_dim * 4 * reps
        res = {"n_gpus": world, "launches_per_rank": reps, "peer_rows_per_launch_rank0": round(peer_rows),
               "requested_bytes_rank0": int(mine), "us_per_launch_max_rank": round(float(t.item()) / reps * 1e3, 1),
               "peer_gather_GBps_per_gpu": round(mine / (float(t.item()) * 1e-3) / 1e9, 1),
               "peer_gather_GBps_all_ranks": round(float(b.item()) / (float(t.item()) * 1e-3) / 1e9, 1),
               "frac_of_measured_peer_copy_770_per_gpu": round(mine / (float(t.item()) * 1e-3) / 1e9 / 770, 3)}
        if c0 and c1:
            res["nvlink_counters_gpu0"] = {"how": c0[2], "rx_bytes_delta": (c1[0] - c0[0]) * 1024, "tx_bytes_delta": (c1[1] - c0[1]) * 1024,
                                           "rx_over_requested": round((c1[0] - c0[0]) * 1024 / mine, 3),
                                           "note": "RX of GPU 0 = rows it pulled from its peers (+ request/response protocol overhead); "
                                                   "TX = rows its peers pulled from its shard; both include the barrier/allreduce traffic of this tool (KBs)"}
        else:
            res["nvlink_counters_gpu0"] = {"unavailable": "neither NVML field values nor nvidia-smi nvlink -gt d answered on this box"}
        print(json.dumps(res))
    store.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
