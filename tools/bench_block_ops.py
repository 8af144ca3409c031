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


# backward kernels timed through the C ABI directly (eager autograd around them is host-bound at this size)
import ctypes as C

from nsa_vibe_b200 import _lib

lib = _lib.load()
F32, BF16 = _lib.NSA_F32, _lib.NSA_BF16
rstd = torch.rand(rows, device="cuda") + 0.5
dxs = [torch.empty(rows, dim, device="cuda") for _ in range(12)]
dw = torch.empty(dim, device="cuda")
part = torch.empty(int(lib.nsa_rmsnorm_partials(rows)), dim, device="cuda")
P = lambda t: C.c_void_p(t.data_ptr())


def bwd_kernel():
    k = nxt()
    rc = lib.nsa_rmsnorm_bwd(P(dy), P(xs[k]), P(w), P(rstd), P(ds), P(dxs[k]), P(dw), P(part), rows, dim, F32, F32, BF16, ops._stream())
    assert rc == 0


def bwd_torch():
    k = nxt()
    y = rmsnorm_torch(xs[k], w, 1e-6)
    torch.autograd.grad([y], [xs[k], w], [dy.float()])


ys = [torch.empty(rows, dim, device="cuda", dtype=torch.bfloat16) for _ in range(12)]
ss = [torch.empty(rows, dim, device="cuda") for _ in range(12)]


def fwd_kernel():
    k = nxt()
    assert lib.nsa_rmsnorm_fwd(P(xs[k]), None, P(w), None, P(ys[k]), P(rstd), rows, dim, C.c_float(1e-6), F32, 0, F32, BF16, ops._stream()) == 0


def fwd_res_kernel():
    k = nxt()
    assert lib.nsa_rmsnorm_fwd(P(xs[k]), P(rs[k]), P(w), P(ss[k]), P(ys[k]), P(rstd), rows, dim, C.c_float(1e-6), F32, BF16, F32, BF16,
                               ops._stream()) == 0


out["fwd_kernel_direct_us"] = timed(fwd_kernel)
out["fwd_res_kernel_direct_us"] = timed(fwd_res_kernel)
out["bwd_kernel_us (dx + ds, dw)"] = timed(bwd_kernel)
out["fwd_bwd_torch_us (eager autograd, fp32 out, no residual)"] = timed(bwd_torch, 30)
b_bwd = rows * dim * (4 + 2 + 4 + 4)
out["bwd_kernel_GBps"] = b_bwd / out["bwd_kernel_us (dx + ds, dw)"] / 1e3
out["algorithmic_bytes_bwd"] = b_bwd
hbm = PEAK.get("hbm_gbs", 6549.1)
b_fwd = rows * dim * (4 + 2)
out["fwd_kernel_GBps"] = b_fwd / out["fwd_kernel_direct_us"] / 1e3
out["fwd_res_kernel_GBps"] = rows * dim * (4 + 2 + 4 + 2) / out["fwd_res_kernel_direct_us"] / 1e3
out["fwd_kernel_frac_of_hbm"] = out["fwd_kernel_GBps"] / float(hbm)
out["algorithmic_bytes_fwd"] = b_fwd
out["bwd_kernel_frac_of_hbm"] = out["bwd_kernel_GBps"] / float(hbm)
out["hbm_peak_GBps"] = float(hbm)
print(json.dumps(out))
