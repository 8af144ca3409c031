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


class OverlappedGradExchange:
    """The same exchange (divide by world, round to bf16, all_reduce, copy back -- bf16_compress_hook, train_showcase.py:654-665)
    issued PER BUCKET while the backward pass is still running: a post-accumulate hook on every parameter counts a bucket's
    gradients in; when the last one lands the bucket is flattened, compressed and handed to NCCL as an ASYNC all_reduce, which runs
    on the process group's own stream next to the backward kernels of the earlier layers.  `finish()` (after backward) waits for
    the collectives and writes the reduced values into p.grad.  This is what DDP's reducer does with its buckets (the reference
    relies on it, train_showcase.py:604-665), without the parts of the reducer that a CUDA graph cannot capture: everything here
    is stream-ordered, so a captured training step replays the overlap.  One flat exchange after backward cost ~1.7 ms of a 26 ms
    step on 8 GPUs; the collective itself is ~0.3 ms and hides under the backward of the next layer."""

    def __init__(self, buckets, *, compress_dtype: torch.dtype = torch.bfloat16):
        self.buckets = [[p for p in b if p.requires_grad] for b in buckets]
        self.buckets = [b for b in self.buckets if b]
        self.compress_dtype = compress_dtype
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self._left = [len(b) for b in self.buckets]
        self._pending = []
        self._handles = []
        self.enabled = True
        if self.world > 1:
            for bi, b in enumerate(self.buckets):
                for p in b:
                    self._handles.append(p.register_post_accumulate_grad_hook(lambda _p, bi=bi: self._ready(bi)))

    def _ready(self, bi: int) -> None:
        if not self.enabled:
            return
        self._left[bi] -= 1
        if self._left[bi] == 0:
            self._launch(bi)

    def _launch(self, bi: int) -> None:
        grads = [p.grad for p in self.buckets[bi]]
        flat = torch.cat([g.reshape(-1) for g in grads])
        packed = torch.empty(flat.shape, dtype=self.compress_dtype, device=flat.device)
        torch.div(flat, self.world, out=packed)  # divide, then compress: the hook's order
        work = dist.all_reduce(packed, op=dist.ReduceOp.SUM, async_op=True)
        self._pending.append((work, packed, grads))

    def finish(self) -> int:
        """Wait for every bucket (buckets whose hooks did not fire -- parameters without a gradient this step -- are exchanged now)
        and decompress into p.grad.  Returns the bytes this rank contributed."""
        if self.world == 1:
            return 0
        for bi, left in enumerate(self._left):
            if left != 0 and any(p.grad is not None for p in self.buckets[bi]):
                ps = [p for p in self.buckets[bi] if p.grad is not None]
                flat = torch.cat([p.grad.reshape(-1) for p in ps])
                packed = torch.empty(flat.shape, dtype=self.compress_dtype, device=flat.device)
                torch.div(flat, self.world, out=packed)
                self._pending.append((dist.all_reduce(packed, op=dist.ReduceOp.SUM, async_op=True), packed, [p.grad for p in ps]))
        nbytes = 0
        for work, packed, grads in self._pending:
            work.wait()
            views, off = [], 0
            for g in grads:
                n = g.numel()
                views.append(packed[off:off + n].view_as(g))
                off += n
            torch._foreach_copy_(grads, views)
            nbytes += packed.numel() * packed.element_size()
        self._pending = []
        self._left = [len(b) for b in self.buckets]
        return nbytes

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def check_bf16_exchange(device, n: int = 1 << 16, seed: int = 0) -> dict:
    """Numerics of the compressed exchange on the live process group: every rank contributes g_r (fp32); the result must be the sum
    over ranks of bf16(g_r / world) -- divide, THEN round -- accumulated by the collective in bf16 or better.  Returns the max
    deviation from the fp32 sum of the compressed contributions in units of the bf16 spacing at the result (world = 2: one bf16
    addition, so at most half a spacing; longer rings may round at every hop).  For power-of-two worlds dividing is exact, so the
    order of divide and round cannot be told apart here; tests/test_dist_cpu.py pins it with a world of three."""
    world, rank = dist.get_world_size(), dist.get_rank()
    g = torch.Generator(device="cpu").manual_seed(seed + rank)
    grad = (torch.randn(n, generator=g) * (1.0 + rank)).to(device)
    p = torch.nn.Parameter(torch.zeros(n, device=device))
    p.grad = grad.clone()
    allreduce_grads_bf16([p])
    mine = (grad / world).to(torch.bfloat16)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    want = torch.stack([t.float() for t in parts]).sum(0)
    got = p.grad.float()
    # every hop of the reduction rounds the running sum once: |error| <= (world - 1) * 2^-8 * sum_r |part_r|
    bound = (world - 1) * 2.0 ** -8 * torch.stack([t.float().abs() for t in parts]).sum(0)
    worst = float(((got - want).abs() / bound.clamp_min(1e-30)).max())
    return {"world": world, "elements": n, "max_error_over_bound": worst, "exact": bool(torch.equal(got, want.to(torch.bfloat16).float())),
            "ok": worst <= 1.0, "bound": "(world-1) * 2^-8 * sum_r |bf16(g_r/world)| per element (one bf16 rounding per hop)"}
