"""Host-side cost of one module-level decode step (cProfile over 300 steps at B=1, where the GPU is never the bound).
    python tools/hostprof_module_decode.py [S]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NSA_PREFILL_BATCHED", "1")
import torch

from nsa_vibe_b200.cache.kv_cache import create_empty_kv
from nsa_vibe_b200.core.block_index import build_block_meta
from nsa_vibe_b200.core.nsa_attention import NSAAttention

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda", 0)
attn = NSAAttention(dim=768, n_heads=12, n_kv_groups=2, d_k=64, d_v=64).to(dev).bfloat16()
kv = create_empty_kv(1, 2, 64, 64, build_block_meta(S + 512, 32, 16, 64, 16, 512), device=dev, dtype=torch.bfloat16)
with torch.no_grad():
    attn(torch.randn(1, S - 400, 768, device=dev).bfloat16(), kv, prefill=True)
    kv.reserve(S + 512)
    x1 = torch.randn(1, 1, 768, device=dev).bfloat16()
    for _ in range(20):
        attn(x1, kv, prefill=False)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        attn(x1, kv, prefill=False)
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(40)
