#!/usr/bin/env python3
"""Config C5: m7c TinyLM training step with the NSA fwd/bwd kernels, DDP over NCCL with the reference's
bf16-compressed gradient allreduce (scripts/train_showcase.py:654-665).  Synthetic byte tokens, random init.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_ddp_bench.py [--layers 12]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from nsa_vibe_b200 import _lib
from nsa_vibe_b200 import dist as nd
from nsa_vibe_b200.model.tiny_lm import TinyLM


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--S", type=int, default=2048)
    ap.add_argument("--B", type=int, default=2, help="sequences per GPU")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--graph", action="store_true", help="capture forward + backward + optimizer step in one CUDA graph and replay it")
    ap.add_argument("--flat", action="store_true", help="with --graph: one flat all_reduce after backward instead of per-layer buckets overlapped with it")
    ap.add_argument("--profile", action="store_true", help="cProfile of the host side of the timed steps (rank 0)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1337 + rank)
    os.environ.setdefault("NSA_PREFILL_BATCHED", "1")
    model = TinyLM(256, 768, a.layers, 12, 2, 64, 64, 32, 16, 64, 16, 512).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    flat_exchange = a.graph and world > 1
    if flat_exchange:
        # DDP's reducer cannot be captured in a CUDA graph; the same exchange (divide, bf16, allreduce, cast back) as one flat NCCL
        # all_reduce after the backward can (nd.allreduce_grads_bf16).  156.6 MB over NVLink: ~0.5 ms of a ~27 ms step.
        nd.broadcast_parameters(model, 0)
        exch = None if a.flat else nd.OverlappedGradExchange(model.grad_buckets())
    elif world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
        nd.register_bf16_compress(model)
    # NSA_OPT_FUSED=1: the reference's opt-in fused AdamW (scripts/train_showcase.py:747-758)
    fused = os.getenv("NSA_OPT_FUSED", "0").lower() in ("1", "true", "yes")
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, capturable=a.graph, fused=fused)
    ids = torch.randint(0, 256, (a.B, a.S + 1), device=dev)
    lib = _lib.load()

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(ids[:, :-1])
        loss = F.cross_entropy(logits.float().reshape(-1, 256), ids[:, 1:].reshape(-1))
        loss.backward()
        if flat_exchange:
            exch.finish() if exch is not None else nd.allreduce_grads_bf16(model.parameters())
        opt.step()
        return loss

    mode = "eager"
    if a.graph:
        # whole-step capture: the step is host-bound in eager mode (~170 NSA launches + ~2000 torch ops per step); the C ABI
        # allocates nothing and never synchronises, so its launches are captured like any other kernel
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(3, a.warmup)):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        holder = {}
        with torch.cuda.graph(graph):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                logits = model(ids[:, :-1])
            holder["loss"] = F.cross_entropy(logits.float().reshape(-1, 256), ids[:, 1:].reshape(-1))
            holder["loss"].backward()
            if flat_exchange:
                exch.finish() if exch is not None else nd.allreduce_grads_bf16(model.parameters())
            opt.step()

        def step():  # noqa: F811
            graph.replay()
            return holder["loss"]

        mode = "cuda-graph replay of forward+backward+AdamW"
    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = lib.nsa_kernel_launches()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = None
    if a.profile and rank == 0:
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    s.record()
    for _ in range(a.steps):
        loss = step()
    e.record()
    if prof is not None:
        prof.disable()
        import pstats
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(45)
    torch.cuda.synchronize()
    ms = nd.max_over_ranks(s.elapsed_time(e) / a.steps, device=dev)
    if rank == 0:
        print(json.dumps({"config": "C5 m7c TinyLM DDP training step", "layers": a.layers, "params": n_params, "S": a.S,
                          "batch_per_gpu": a.B, "n_gpus": world, "mode": mode, "adamw": "fused" if fused else "foreach", "ms_per_step": ms, "tokens_per_s": world * a.B * a.S / (ms * 1e-3),
                          "loss": float(loss.detach()), "nsa_kernel_launches_per_step": (lib.nsa_kernel_launches() - n0) / a.steps,
                          "grad_allreduce": (("one flat bf16 NCCL all_reduce captured in the graph" if a.flat else "per-layer bf16 buckets, async NCCL all_reduce overlapped with backward, captured in the graph") if flat_exchange else "DDP bf16_compress_hook over NCCL") if world > 1 else "none (single GPU)",
                          "allreduce_bytes_per_step": 2 * n_params if world > 1 else 0}))
    if flat_exchange:
        # a process group whose collective sits in a live CUDA graph does not tear down cleanly here (destroy_process_group hung
        # on 2 B200s until the launcher's timeout): flush and leave without the NCCL destructor
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
