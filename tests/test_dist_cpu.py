"""world_size-2 gloo tests (CPU) of the host-side multi-rank logic: minibatch scheduling
(reference sampler.py:166-185) and the gradient exchange (reference main.py:149-168: SUM, not mean)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnn_b200 import harness, sampler


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    params = list(model.parameters())
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    params[1].grad = None                                   # parameters without a gradient are skipped (main.py:156)
    nbytes = harness.exchange_gradients(params, world)
    expect = sum(r + 1 for r in range(world))
    ok = all(torch.allclose(p.grad, torch.full_like(p, float(expect) * (i + 1))) for i, p in enumerate(params) if p.grad is not None)
    ok = ok and params[1].grad is None and nbytes == 4 * sum(p.numel() for i, p in enumerate(params) if i != 1)
    # FlatGradients: clip every replica to norm 5 (main.py:146), THEN sum the replicas (main.py:149-168) - on one flat
    # buffer the parameters' .grad are views of; must equal clip_grad_norm_ + exchange_gradients on a twin model
    torch.manual_seed(1)
    twin_a = torch.nn.Sequential(torch.nn.Linear(7, 9), torch.nn.ELU(), torch.nn.Linear(9, 4))
    twin_b = torch.nn.Sequential(torch.nn.Linear(7, 9), torch.nn.ELU(), torch.nn.Linear(9, 4))
    twin_b.load_state_dict(twin_a.state_dict())
    pa, pb = list(twin_a.parameters()), list(twin_b.parameters())
    flat = harness.FlatGradients(pb, world, max_norm=5.0)
    for step in range(2):
        x = torch.randn(6, 7, generator=torch.Generator().manual_seed(10 * rank + step)) * (40.0 if step == 0 else 0.01)
        for q in pa:
            q.grad = None
        twin_a(x).square().sum().backward()
        torch.nn.utils.clip_grad_norm_(pa, 5)
        harness.exchange_gradients(pa, world)
        flat.zero()
        twin_b(x).square().sum().backward()
        nb = flat.clip_and_exchange()
        ok = ok and nb == 4 * sum(q.numel() for q in pb)
        ok = ok and all(q.grad.data_ptr() == flat.flat[o:].data_ptr() for q, o in zip(pb, np.cumsum([0] + [q.numel() for q in pb[:-1]])))
        ok = ok and all(torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-7) for a, b in zip(pa, pb))
    batches = sampler.rank_batches(1001, 64, rank, world, iter_num=3)
    np.save(os.path.join(out_dir, f"b{rank}.npy"), np.concatenate(batches))
    np.save(os.path.join(out_dir, f"n{rank}.npy"), np.array([len(batches), int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_schedule_and_gradient_sum(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    idx = [np.load(tmp_path / f"b{r}.npy") for r in range(world)]
    meta = [np.load(tmp_path / f"n{r}.npy") for r in range(world)]
    assert all(m[1] == 1 for m in meta), "allreduce(SUM) of the flattened gradient gave the wrong result"
    # one global permutation, contiguous chunk per rank, chunks disjoint and complete (sampler.py:170-189)
    allidx = np.concatenate(idx)
    assert np.array_equal(np.sort(allidx), np.arange(1001))
    torch.manual_seed(3)
    perm = torch.randperm(1001).numpy()
    assert np.array_equal(idx[0], perm[:501]) and np.array_equal(idx[1], perm[501:])
    assert meta[0][0] == 8 and meta[1][0] == 8       # ceil(501/64), ceil(500/64)


def test_single_rank_exchange_is_noop():
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert harness.exchange_gradients([p], 1) == 0 and torch.all(p.grad == 2.0)
