"""Module-level prefill (NSAAttention.forward, m7c dims, bf16): time per call and the top GPU kernels.
    python tools/prof_module.py [S] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("NSA_PREFILL_BATCHED", "1")
import torch
from torch.profiler import ProfilerActivity, profile

from nsa_vibe_b200.cache.kv_cache import create_empty_kv
from nsa_vibe_b200.core.block_index import build_block_meta
from nsa_vibe_b200.core.nsa_attention import NSAAttention

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
attn = NSAAttention(dim=768, n_heads=12, n_kv_groups=2, d_k=64, d_v=64).to(dev).bfloat16()
x = torch.randn(B, S, 768, device=dev).bfloat16()
meta = build_block_meta(S, 32, 16, 64, 16, 512)


def call():
    kv = create_empty_kv(B, 2, 64, 64, meta, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        return attn(x, kv, prefill=True)[0]


for _ in range(3):
    call()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    call()
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print(f"NSAAttention prefill S={S} B={B}: {ms:.3f} ms per call, {B * S / ms * 1e3 / 1e6:.2f} M tok/s")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    call()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))
