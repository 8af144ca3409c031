"""Run the split scorer's two kernels a few times at S=64k (for ncu captures): python tools/prof_fused.py [S]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nsa_vibe_b200 import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
G, h, D, l, d, ls, n, w = 2, 6, 64, 32, 16, 64, 16, 512
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device="cuda").bfloat16()
S_cmp = (S - l) // d + 1
Q, Kc, Vc = r(1, S, G, h, D), r(1, G, S_cmp, D), r(1, G, S_cmp, D)
cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
with torch.no_grad():
    for _ in range(3):
        st = ops.score_stats(Q, Kc, cfg)
        ops.score_cmp(Q, Kc, Vc, cfg, st)
torch.cuda.synchronize()
print("ok")
