import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nsa_vibe_b200 import ops
G, h, D, l, d, ls, n, w = 2, 6, 64, 32, 16, 64, 16, 512
S, B = 65536, 1
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device=dev).bfloat16()
S_cmp = (S - l) // d + 1
Q, Kw, Vw, Kc, Vc = r(B, S, G, h, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S_cmp, D), r(B, G, S_cmp, D)
cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
br = ops.BR_WIN if (len(sys.argv) > 1 and sys.argv[1] == "win") else ops.BR_CMP
K, V = (Kw, Vw) if br == ops.BR_WIN else (Kc, Vc)
with torch.no_grad():
    for _ in range(4):
        ops.branch_attention(br, Q, K, V, cfg)
torch.cuda.synchronize()
