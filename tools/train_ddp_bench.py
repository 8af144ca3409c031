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
from nsa_vibe_b200.model.llama_block_nsa import LlamaBlockNSA, RMSNorm


class TinyLM(nn.Module):  # scripts/train_showcase.py:30-117 (embedding -> blocks -> norm -> lm_head)
    def __init__(self, vocab, dim, n_layers, heads, groups, dk, dv, l, d, l_sel, n_sel, w):
        super().__init__()
        self.embed = nn.Embedding(vocab, dim)
        self.blocks = nn.ModuleList([LlamaBlockNSA(dim, heads, groups, dk, dv, l, d, l_sel, n_sel, w) for _ in range(n_layers)])
        self.norm_f = RMSNorm(dim)
        self.lm_head = nn.Linear(dim, vocab, bias=False)

    def forward(self, ids):
        x = self.embed(ids)
        for b in self.blocks:
            x = b(x)
        return self.lm_head(self.norm_f(x))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--S", type=int, default=2048)
    ap.add_argument("--B", type=int, default=2, help="sequences per GPU")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
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
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local], gradient_as_bucket_view=True)
        nd.register_bf16_compress(model)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    ids = torch.randint(0, 256, (a.B, a.S + 1), device=dev)
    lib = _lib.load()

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(ids[:, :-1])
        loss = F.cross_entropy(logits.float().reshape(-1, 256), ids[:, 1:].reshape(-1))
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = lib.nsa_kernel_launches()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(a.steps):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = nd.max_over_ranks(s.elapsed_time(e) / a.steps, device=dev)
    if rank == 0:
        print(json.dumps({"config": "C5 m7c TinyLM DDP training step", "layers": a.layers, "params": n_params, "S": a.S,
                          "batch_per_gpu": a.B, "n_gpus": world, "ms_per_step": ms, "tokens_per_s": world * a.B * a.S / (ms * 1e-3),
                          "loss": float(loss), "nsa_kernel_launches_per_step": (lib.nsa_kernel_launches() - n0) / a.steps,
                          "grad_allreduce": "DDP bf16_compress_hook over NCCL" if world > 1 else "none (single GPU)",
                          "allreduce_bytes_per_step": 2 * n_params if world > 1 else 0}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
