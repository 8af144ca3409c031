"""Run the standalone selection kernel a few times (for ncu captures): python tools/prof_select.py [S]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nsa_vibe_b200 import ops

S = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
G, ls, n = 2, 64, 16
pg = torch.rand(1, S, G, (S + ls - 1) // ls, device="cuda")
for _ in range(3):
    ops.select_ranges_prefill(pg, ls, n, S)
torch.cuda.synchronize()
print("ok")
