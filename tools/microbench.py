#!/usr/bin/env python3
"""Per-kernel timings of the hot path at the bench shapes (CUDA events, warm, inputs > L2)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from nsa_vibe_b200 import ops


def timeit(fn, n=5, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--S", type=int, default=65536)
    ap.add_argument("--B", type=int, default=1)
    ap.add_argument("--what", default="sel,cmp,win,score,fwd")
    ap.add_argument("--decode-B", type=int, default=512)
    a = ap.parse_args()
    G, h, D, l, d, ls, n, w = 2, 6, 64, 32, 16, 64, 16, 512
    S, B = a.S, a.B
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g, device=dev).bfloat16()
    S_cmp = (S - l) // d + 1
    Q, Ks, Vs, Kw, Vw, Kc, Vc = r(B, S, G, h, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S, D), r(B, G, S_cmp, D), r(B, G, S_cmp, D)
    gate = (torch.randn(32, 64, device=dev) * 0.1, torch.zeros(32, device=dev), torch.randn(3, 32, device=dev) * 0.1, torch.zeros(3, device=dev))
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
    what = a.what.split(",")
    with torch.no_grad():
        t0 = time.time()
        ranges = ops.score_select(Q, Kc, cfg, mode=0)
        torch.cuda.synchronize()
        print(f"S={S} B={B} first score_select {1e3 * (time.time() - t0):.1f} ms")
        if "score" in what:
            print(f"score_select          : {timeit(lambda: ops.score_select(Q, Kc, cfg, mode=0), n=3, warm=1):9.3f} ms")
        for name, br, K, V in (("sel", 1, Ks, Vs), ("cmp", 0, Kc, Vc), ("win", 2, Kw, Vw)):
            if name in what:
                for impl, nm in ((ops.IMPL_AUTO, "auto"), (ops.IMPL_SIMT, "simt")):
                    c2 = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=impl)
                    ms = timeit(lambda: ops.branch_attention(br, Q, K, V, c2, ranges if br == 1 else None), n=3, warm=1)
                    print(f"branch {name} ({nm:4s})     : {ms:9.3f} ms")
        if "sel2" in what:
            ms = timeit(lambda: ops.sel_attention_blockmajor(Q, Ks, Vs, cfg, ranges), n=3, warm=1)
            print(f"branch sel block-major: {ms:9.3f} ms")
        if "decode" in what:
            Sd, Bd = 4096, a.decode_B
            cap = Sd + 64
            Ks2, Vs2, Kw2, Vw2 = r(Bd, G, cap, D), r(Bd, G, cap, D), r(Bd, G, cap, D), r(Bd, G, cap, D)
            Sc = (Sd - l) // d + 1
            Kc2, Vc2 = r(Bd, G, Sc + 8, D), r(Bd, G, Sc + 8, D)
            q = r(Bd, 1, G, h, D)
            gc = ops._gate_struct(gate, torch.device(dev))
            out = torch.empty((Bd, 1, G, h, D), dtype=torch.bfloat16, device=dev)
            rg = torch.empty((Bd, G, n, 2), dtype=torch.int32, device=dev)
            for impl, nm in ((ops.IMPL_AUTO, "auto"), (ops.IMPL_SIMT, "simt")):
                c2 = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=impl)
                f = lambda: ops.decode_core(q, Ks2, Vs2, Kw2, Vw2, Kc2, Vc2, gate, c2, t=Sd - 1, S_sel_kv=Sd, S_win_kv=Sd, win_off=0,
                                            S_cmp=Sc, ranges_out=rg, out=out, gate_cache=gc)
                ms = timeit(f, n=20, warm=3)
                print(f"decode step B={Bd} S={Sd} ({nm:4s}): {1e3 * ms:9.1f} us  -> {Bd * 916992 / (ms * 1e-3) / 1e9:7.1f} GB/s algorithmic")
            c2 = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
            dm_q = q
            ms = timeit(lambda: ops.score_select(q, Kc2, c2, mode=1, t0=Sd - 1, S_total=Sd, S_cmp=Sc), n=20, warm=3)
            print(f"decode score_select only: {1e3 * ms:9.1f} us")
            ms = timeit(lambda: ops.gate_forward(q, gate, c2), n=20, warm=3)
            print(f"decode gate only        : {1e3 * ms:9.1f} us")
        if "fwd" in what:
            ms = timeit(lambda: ops.prefill_core(Q, Ks, Vs, Kw, Vw, Kc, Vc, gate, cfg, sel_mode=0, ranges=ranges), n=3, warm=1)
            print(f"prefill_fwd (given ranges): {ms:9.3f} ms")
    if "bwd" in what:  # backward of the fused hot path alone (graph retained), tensor-core vs SIMT kernels
        impls = [(ops.IMPL_AUTO, "auto")] + ([(ops.IMPL_SIMT, "simt")] if S <= 8192 else [])
        for impl, nm in impls:
            c2 = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=impl)
            leaves = [t.clone().requires_grad_(True) for t in (Q, Ks, Vs, Kw, Vw, Kc, Vc)]
            O, _, _ = ops.prefill_core(*leaves, gate, c2, sel_mode=0, ranges=ranges)
            dO = torch.randn_like(O)
            ms = timeit(lambda: torch.autograd.grad(O, leaves, dO, retain_graph=True), n=3, warm=1)
            print(f"prefill_bwd ({nm:4s}) incl. torch zero-fill/casts: {ms:9.3f} ms")
            del O, leaves


if __name__ == "__main__":
    main()
