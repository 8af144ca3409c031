import os, subprocess, sys
CASES = [(1, 2048, 6), (2, 4173, 6)]
STOPS = ["0"]
CHILD = r'''
import sys
sys.path.insert(0, "%s")
import torch
from nsa_vibe_b200 import ops
B, S, h = %d, %d, %d
G, l, d, ls, n, w = 2, 32, 16, 64, 16, 512
gen = torch.Generator().manual_seed(3)
Q = torch.randn(B, S, G, h, 64, generator=gen).bfloat16()
Kc = torch.randn(B, G, (S - l) // d + 1, 64, generator=gen).bfloat16()
cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
got = ops.score_pgrp(Q.cuda(), Kc.cuda(), cfg)
torch.cuda.synchronize()
print("ok", float(got.sum()))
'''
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for c, stop in [(c, st) for c in CASES for st in STOPS]:
    env = dict(os.environ, NSA_B200_SCORE_MT="4", NSA_B200_SCORE_STOP=stop)
    r = subprocess.run([sys.executable, "-c", CHILD % ((root,) + c)], env=env, capture_output=True, text=True, timeout=120)
    tail = (r.stdout.strip().splitlines() or [""])[-1] if r.returncode == 0 else (r.stderr.strip().splitlines() or [""])[-1][:100]
    print(c, "stop", stop, "rc", r.returncode, tail, flush=True)
