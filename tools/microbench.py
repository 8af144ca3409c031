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
        if "fwd" in what:
            ms = timeit(lambda: ops.prefill_core(Q, Ks, Vs, Kw, Vw, Kc, Vc, gate, cfg, sel_mode=0, ranges=ranges), n=3, warm=1)
            print(f"prefill_fwd (given ranges): {ms:9.3f} ms")


if __name__ == "__main__":
    main()
