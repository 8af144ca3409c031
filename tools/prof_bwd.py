"""Run the fused backward a few times (for ncu captures): python tools/prof_bwd.py [S] [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nsa_vibe_b200 import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
G, h, D, l, d, ls, n, w = 2, 6, 64, 32, 16, 64, 16, 512
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device=dev).bfloat16().requires_grad_(True)
S_cmp = (S - l) // d + 1
leaves = [r(B, S, G, h, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S_cmp, D), r(B, G, S_cmp, D)]
gate = tuple(t.requires_grad_(True) for t in (torch.randn(32, 64, device=dev) * 0.1, torch.zeros(32, device=dev),
                                              torch.randn(3, 32, device=dev) * 0.1, torch.zeros(3, device=dev)))
cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
O, _, _ = ops.prefill_core(*leaves, gate, cfg, sel_mode=0)
dO = torch.randn_like(O)
for _ in range(3):
    torch.autograd.grad(O, leaves + list(gate), dO, retain_graph=True)
torch.cuda.synchronize()
print("ok")
