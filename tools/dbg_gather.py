import sys
sys.path.insert(0, '/root/repo')
import torch
from nsa_vibe_b200 import ops
G, h, D, l, d, ls, n, w = 2, 6, 64, 32, 16, 64, 16, 512
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device=dev).bfloat16()
Sd, Bd = 4096, 512
cap = Sd + 64
Ks2, Vs2, Kw2, Vw2 = r(Bd, G, cap, D), r(Bd, G, cap, D), r(Bd, G, cap, D), r(Bd, G, cap, D)
Sc = (Sd - l) // d + 1
Kc2, Vc2 = r(Bd, G, Sc + 8, D), r(Bd, G, Sc + 8, D)
q = r(Bd, 1, G, h, D)
gate = (torch.randn(32, 64, device=dev) * 0.1, torch.zeros(32, device=dev), torch.randn(3, 32, device=dev) * 0.1, torch.zeros(3, device=dev))
c2 = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
out = torch.empty((Bd, 1, G, h, D), dtype=torch.bfloat16, device=dev)
rg = torch.empty((Bd, G, n, 2), dtype=torch.int32, device=dev)
for _ in range(6):
    ops.decode_core(q, Ks2, Vs2, Kw2, Vw2, Kc2, Vc2, gate, c2, t=Sd - 1, S_sel_kv=Sd, S_win_kv=Sd, win_off=0, S_cmp=Sc, ranges_out=rg, out=out)
torch.cuda.synchronize()
