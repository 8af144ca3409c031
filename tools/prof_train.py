"""GPU-time breakdown of one C5 training step (torch.profiler, CUDA activities): python tools/prof_train.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F
from torch.profiler import ProfilerActivity, profile

from train_ddp_bench import TinyLM

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
os.environ.setdefault("NSA_PREFILL_BATCHED", "1")
model = TinyLM(256, 768, 12, 12, 2, 64, 64, 32, 16, 64, 16, 512).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
ids = torch.randint(0, 256, (B, 2049), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        logits = model(ids[:, :-1])
    loss = F.cross_entropy(logits.float().reshape(-1, 256), ids[:, 1:].reshape(-1))
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=110))
