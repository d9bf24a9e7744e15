"""Batch sharding across GPUs: registration pairs are independent, so rank r of W takes
a contiguous slice of the batch, runs the whole path on its own device / stream, and results are concatenated
in rank order.  No collective sits on the data path; `gather_results` exists for callers that want every rank's
poses on rank 0 (evaluation bookkeeping), over whatever backend the process group uses (NCCL on GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int):
    """Contiguous, balanced partition: the first n_items % world_size ranks take one extra item."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, world_size: int, rank: int):
    n = tensors[0].shape[0]
    lo, hi = shard_bounds(n, world_size, rank)
    return [t[lo:hi] for t in tensors]


def gather_results(local: torch.Tensor, n_items: int, dst: int = 0):
    """Concatenate per-rank result slices (shard_bounds order) on rank ``dst``; other ranks get None."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    W, r = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_items, W, i) for i in range(W)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(W)] if r == dst else None
    dist.gather(pad, bufs, dst=dst)
    if r != dst:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def allreduce_gradients(module: torch.nn.Module, average: bool = True):
    """Data-parallel LPD pre-training (BASELINE config 3 at N > 1): ONE all-reduce of the flattened fp32 gradient
    (5.6 M parameters for the full VCRNet, 0.4 M for the LPDNet embedding) per step -- the only collective in the
    framework (NCCL over NVLink on GPUs; gloo in the CPU tests).  Replaces nn.DataParallel's per-step replicate +
    reduce in the reference's train loop (util/initPara.py:260, model/lpdnet_model.py:232-276).
    Parameters without a gradient contribute zeros so every rank reduces the same layout."""
    params = [p for p in module.parameters() if p.requires_grad]
    if not params or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for p in params:
        n = p.numel()
        g = flat[off:off + n].view_as(p)
        if p.grad is None:
            p.grad = g.clone()
        else:
            p.grad.copy_(g)
        off += n
    return off
