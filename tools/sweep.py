#!/usr/bin/env python3
"""Configs C2 / C3 of SURVEY 8d on one GPU: m7c prefill at S=2048 (B=1, 32) and the decode sweep S=512..4096
(B=1 and B=592), one JSON line each."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nsa_vibe_b200 import ops

G, h, D, l, d, ls, n, w = 2, 6, 64, 32, 16, 64, 16, 512
dev = "cuda"
cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
gen = torch.Generator(device=dev).manual_seed(0)
r = lambda *s: torch.randn(*s, generator=gen, device=dev).bfloat16()
gate = (torch.randn(32, 64, device=dev) * 0.1, torch.zeros(32, device=dev), torch.randn(3, 32, device=dev) * 0.1, torch.zeros(3, device=dev))


def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


with torch.no_grad():
    for B, S in ((1, 2048), (32, 2048), (8, 8192), (1, 16384)):
        S_cmp = (S - l) // d + 1
        Q, Ks, Vs, Kw, Vw, Kc, Vc = r(B, S, G, h, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S_cmp, D), r(B, G, S_cmp, D)

        def step():
            rg = ops.score_select(Q, Kc, cfg, mode=0)
            return ops.prefill_core(Q, Ks, Vs, Kw, Vw, Kc, Vc, gate, cfg, sel_mode=0, ranges=rg)[0]
        ms = timeit(step)
        print(json.dumps({"config": "C2 m7c NSA core prefill (one layer), bf16", "B": B, "S": S, "ms": ms, "tok_per_s": B * S / (ms * 1e-3)}), flush=True)
    gc = ops._gate_struct(gate, torch.device(dev))
    for Bd in (1, 592):
        for Sd in (512, 1024, 2048, 4096):
            cap = Sd + 64
            Ks2, Vs2, Kw2, Vw2 = r(Bd, G, cap, D), r(Bd, G, cap, D), r(Bd, G, cap, D), r(Bd, G, cap, D)
            Sc = (Sd - l) // d + 1
            Kc2, Vc2 = r(Bd, G, Sc + 8, D), r(Bd, G, Sc + 8, D)
            q = r(Bd, 1, G, h, D)
            out = torch.empty((Bd, 1, G, h, D), dtype=torch.bfloat16, device=dev)
            rg = torch.empty((Bd, G, n, 2), dtype=torch.int32, device=dev)
            f = lambda: ops.decode_core(q, Ks2, Vs2, Kw2, Vw2, Kc2, Vc2, gate, cfg, t=Sd - 1, S_sel_kv=Sd, S_win_kv=Sd, win_off=0, S_cmp=Sc,
                                        ranges_out=rg, out=out, gate_cache=gc)
            ms = timeit(f, n=20)
            reads = Sc + n * ls + min(w, Sd)
            byts = reads * G * 2 * D * 2
            print(json.dumps({"config": "C3 decode step (fused kernel), bf16", "B": Bd, "S": Sd, "us_per_step": ms * 1e3, "us_per_token": ms * 1e3 / Bd,
                              "reads_per_token": reads, "algorithmic_GBps": Bd * byts / (ms * 1e-3) / 1e9, "frac_of_hbm_6549": Bd * byts / (ms * 1e-3) / 1e9 / 6549.1}), flush=True)

# ---- the same two configs through the module API (NSAAttention.forward: projections, cache append, hot path, output projection) ----
os.environ.setdefault("NSA_PREFILL_BATCHED", "1")
from nsa_vibe_b200.cache.kv_cache import create_empty_kv  # noqa: E402
from nsa_vibe_b200.core.block_index import build_block_meta  # noqa: E402
from nsa_vibe_b200.core.nsa_attention import NSAAttention  # noqa: E402

torch.manual_seed(0)
attn = NSAAttention(dim=768, n_heads=12, n_kv_groups=G, d_k=D, d_v=D, l=l, d=d, l_sel=ls, n_sel=n, w=w).to(dev).bfloat16()
mk = lambda B_, S_: create_empty_kv(B_, G, D, D, build_block_meta(S_ + 64, l, d, ls, n, w), device=torch.device(dev), dtype=torch.bfloat16)
with torch.no_grad():
    for B in (1, 32):  # C2: m7c prefill S=2048 (bench/bench_prefill.py:76-85)
        x = torch.randn(B, 2048, 768, device=dev).bfloat16()
        ms = timeit(lambda: attn(x, mk(B, 2048), prefill=True), n=10)
        print(json.dumps({"config": "C2 NSAAttention.forward(prefill=True), m7c dims, bf16 (module API, one layer)", "B": B, "S": 2048, "ms": ms,
                          "tok_per_s": B * 2048 / (ms * 1e-3)}), flush=True)
    for Bd in (1, 64, 592):  # C3: bench/bench_decode.py:113-136 -- prefill a random context, then timed single-token steps
        for Sd in (512, 1024, 2048, 4096):
            kv = mk(Bd, Sd)
            attn(torch.randn(Bd, Sd - 48, 768, device=dev).bfloat16(), kv, prefill=True)
            kv.reserve(Sd + 64)
            x1 = torch.randn(Bd, 1, 768, device=dev).bfloat16()
            us = timeit(lambda: attn(x1, kv, prefill=False), n=32, warm=8) * 1e3
            print(json.dumps({"config": "C3 NSAAttention.forward(prefill=False), m7c dims, bf16 (module API)", "B": Bd, "S": int(kv.K_sel.shape[2]),
                              "us_per_step": us, "us_per_token": us / Bd}), flush=True)
            del kv
