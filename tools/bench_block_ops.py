"""RMSNorm(+residual) kernel vs the reference's ATen chain at the C5 training shape (16384 rows x 768, fp32 stream, bf16 out).
    python tools/bench_block_ops.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nsa_vibe_b200 import ops
from nsa_vibe_b200.model.llama_block_nsa import rmsnorm_torch

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


rows, dim = 8 * 2048, 768
# 12 layers' worth of distinct rows so that the working set (12 x 50 MB) does not sit in the 126 MB L2
xs = [torch.randn(rows, dim, device="cuda", requires_grad=True) for _ in range(12)]
rs = [torch.randn(rows, dim, device="cuda").bfloat16() for _ in range(12)]
w = torch.ones(dim, device="cuda", requires_grad=True)
dy = torch.randn(rows, dim, device="cuda").bfloat16()
ds = torch.randn(rows, dim, device="cuda")
out = {}
i = [0]


def nxt():
    i[0] = (i[0] + 1) % 12
    return i[0]


with torch.no_grad():
    out["fwd_kernel_us"] = timed(lambda: ops.rmsnorm(xs[nxt()], w, 1e-6, out_dtype=torch.bfloat16))
    out["fwd_res_kernel_us"] = timed(lambda: ops.rmsnorm(xs[nxt()], w, 1e-6, residual=rs[i[0]], out_dtype=torch.bfloat16))
    out["fwd_torch_us"] = timed(lambda: rmsnorm_torch(xs[nxt()], w, 1e-6).bfloat16())
    out["fwd_res_torch_us"] = timed(lambda: rmsnorm_torch(xs[nxt()] + rs[i[0]], w, 1e-6).bfloat16())


def fb_kernel():
    k = nxt()
    s, y = ops.rmsnorm(xs[k], w, 1e-6, residual=rs[k].requires_grad_(True), out_dtype=torch.bfloat16)
    torch.autograd.backward([s, y], [ds, dy])


def fb_torch():
    k = nxt()
    s = xs[k] + rs[k].requires_grad_(True)
    y = rmsnorm_torch(s, w, 1e-6).bfloat16()
    torch.autograd.backward([s, y], [ds, dy])


out["fwd_bwd_res_kernel_us"] = timed(fb_kernel, 30)
out["fwd_bwd_res_torch_us"] = timed(fb_torch, 30)
hbm = PEAK.get("hbm_gbs", 6549.1)
b_fwd = rows * dim * (4 + 2)
out["fwd_kernel_GBps"] = b_fwd / out["fwd_kernel_us"] / 1e3
out["fwd_kernel_frac_of_hbm"] = out["fwd_kernel_GBps"] / float(hbm)
out["algorithmic_bytes_fwd"] = b_fwd
out["hbm_peak_GBps"] = float(hbm)
print(json.dumps(out))
