"""World-size-2 gloo test (CPU) of the multi-GPU host logic: the batch partition is a disjoint cover in rank
order and gathered results reassemble the unsharded answer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vcr_net_b200.shard import allreduce_gradients, gather_results, shard_batch, shard_bounds


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    src = torch.arange(n_items * 3 * 5, dtype=torch.float32).reshape(n_items, 3, 5)
    (mine,) = shard_batch([src], world, rank)
    # stand-in for the per-rank registration: any per-pair function commutes with the sharding
    local = mine.sum(dim=2) * 2.0 + 1.0
    full = gather_results(local, n_items, dst=0)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover():
    for n in (1, 2, 7, 16, 24, 255):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather(tmp_path):
    n_items, world = 7, 2
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(world, _free_port(), n_items, out), nprocs=world, join=True)
    src = torch.arange(n_items * 3 * 5, dtype=torch.float32).reshape(n_items, 3, 5)
    want = (src.sum(dim=2) * 2.0 + 1.0).numpy()
    assert np.array_equal(np.load(out), want)


def _grad_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))     # same init on every rank
    x = torch.arange(8 * 5, dtype=torch.float32).reshape(8, 5) / 10.0
    (mine,) = shard_batch([x], world, rank)
    net(mine).pow(2).sum().backward()                                           # per-rank gradient of its shard
    n = allreduce_gradients(net, average=False)
    if rank == 0:
        np.save(out_path, torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy())
        assert n == sum(p.numel() for p in net.parameters())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce(tmp_path):
    """Sum of per-shard gradients over 2 ranks == gradient of the unsharded batch (training all-reduce)."""
    out = str(tmp_path / "grad.npy")
    mp.spawn(_grad_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 4), torch.nn.Linear(4, 2))
    x = torch.arange(8 * 5, dtype=torch.float32).reshape(8, 5) / 10.0
    net(x).pow(2).sum().backward()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()]).numpy()
    assert np.allclose(np.load(out), want, rtol=1e-5, atol=1e-6)
