"""Multi-GPU plumbing for the NSA hot path.

The path shards by batch (x KV group) with NO exchange step: scores, Eq.10 sums, top-n, all three branch attentions,
the gate and their gradients are independent across (b, g).  So the only distributed logic is
  * which sequences a rank owns (`shard_batch`),
  * device-side timing reduced as the max over ranks (`max_over_ranks`),
  * and, for the DDP training config only, the reference's bf16-compressed gradient allreduce
    (scripts/train_showcase.py:654-665) -- torch DDP + NCCL over NVLink, registered by `register_bf16_compress`.
One process per GPU; rendezvous through the usual RANK / WORLD_SIZE / MASTER_* variables.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of `global_batch` sequences: returns (first, count) for `rank`.
    Ranks differ by at most one sequence; no sequence is split (selection is group-consistent, so a sequence's
    (b, g) rows may live on one GPU without any collective)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(global_batch, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (elapsed milliseconds measured with CUDA events) over the job."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def register_bf16_compress(ddp_model) -> None:
    """The reference's DDP setting (train_showcase.py:654-665): gradients are cast to bf16 for the bucketed allreduce and
    cast back.  78.3 M parameters -> 156.6 MB per step over NCCL / NVLink."""
    from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
    ddp_model.register_comm_hook(state=None, hook=default_hooks.bf16_compress_hook)


def broadcast_parameters(module: torch.nn.Module, src: int = 0) -> None:
    """What DDP's constructor does: every rank starts from rank `src`'s parameters and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)


def allreduce_grads_bf16(params, *, compress_dtype: torch.dtype = torch.bfloat16) -> int:
    """The reference's DDP gradient exchange with `bf16_compress_hook` (scripts/train_showcase.py:654-665) as ONE flat collective
    that a CUDA graph can capture (DDP's reducer cannot be captured): gradients are divided by the world size, cast to bf16,
    summed over the ranks in one all_reduce, cast back and written into `p.grad` -- the hook's arithmetic (divide, compress,
    allreduce, decompress) on one bucket holding every gradient.  Returns the bytes each rank contributes (2 per element)."""
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return 0
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    grads = [p.grad for p in ps]
    flat = torch.cat([g.reshape(-1) for g in grads])
    if world > 1:
        packed = torch.empty(flat.shape, dtype=compress_dtype, device=flat.device)
        torch.div(flat, world, out=packed)  # divide, then compress: the hook's order (one pass)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        views, off = [], 0
        for g in grads:
            n = g.numel()
            views.append(packed[off:off + n].view_as(g))
            off += n
        torch._foreach_copy_(grads, views)  # decompress into p.grad: one multi-tensor kernel instead of one copy per parameter
    return int(flat.numel()) * 2
