"""Module-level decode (NSAAttention.forward(prefill=False), m7c dims, bf16) after a prefill of S tokens: time per step and the
top GPU kernels.    python tools/prof_module_decode.py [S] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NSA_PREFILL_BATCHED", "1")
import torch
from torch.profiler import ProfilerActivity, profile

from nsa_vibe_b200.cache.kv_cache import create_empty_kv
from nsa_vibe_b200.core.block_index import build_block_meta
from nsa_vibe_b200.core.nsa_attention import NSAAttention

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
dev = torch.device("cuda", 0)
attn = NSAAttention(dim=768, n_heads=12, n_kv_groups=2, d_k=64, d_v=64).to(dev).bfloat16()
meta = build_block_meta(S + 256, 32, 16, 64, 16, 512)
kv = create_empty_kv(B, 2, 64, 64, meta, device=dev, dtype=torch.bfloat16)
with torch.no_grad():
    for b0 in range(0, 1):
        attn(torch.randn(B, S - 40, 768, device=dev).bfloat16(), kv, prefill=True)
    x1 = torch.randn(B, 1, 768, device=dev).bfloat16()
    for _ in range(8):
        attn(x1, kv, prefill=False)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    n = 24
    for _ in range(n):
        attn(x1, kv, prefill=False)
    e.record()
    torch.cuda.synchronize()
    us = s.elapsed_time(e) / n * 1e3
    print(f"NSAAttention decode S~{S} B={B}: {us:.1f} us per step, {us / B:.3f} us per token")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        attn(x1, kv, prefill=False)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
