"""Run one hot-path kernel a few times (for ncu captures): python tools/prof_one.py {score|score_select|cmp|win|sel|sel2} [S]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nsa_vibe_b200 import ops

what = sys.argv[1] if len(sys.argv) > 1 else "score"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
G, h, D, l, d, ls, n, w = 2, 6, 64, 32, 16, 64, 16, 512
B = 1
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g, device=dev).bfloat16()
S_cmp = (S - l) // d + 1
Q, Ks, Vs, Kw, Vw, Kc, Vc = r(B, S, G, h, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S_cmp, D), r(B, G, S_cmp, D)
cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
with torch.no_grad():
    ranges = ops.score_select(Q, Kc, cfg, mode=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        if what == "score":
            ops.score_pgrp(Q, Kc, cfg)
        elif what == "score_select":  # the scorer as the hot path runs it: pass 2 stops at the causal limit
            ops.score_select(Q, Kc, cfg, mode=0)
        elif what == "cmp":
            ops.branch_attention(ops.BR_CMP, Q, Kc, Vc, cfg)
        elif what == "win":
            ops.branch_attention(ops.BR_WIN, Q, Kw, Vw, cfg)
        elif what == "sel2":
            ops.sel_attention_blockmajor(Q, Ks, Vs, cfg, ranges, ranges_trusted=True)
        else:
            ops.branch_attention(ops.BR_SEL, Q, Ks, Vs, cfg, ranges, ranges_trusted=True)
    e1.record()
torch.cuda.synchronize()
print(f"ok {what} S={S}: {e0.elapsed_time(e1) / reps:.3f} ms per call (first call included)")
