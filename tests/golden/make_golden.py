#!/usr/bin/env python3
"""Generate golden vectors by running the REAL reference (seconds-0/nsa-vibe) on CPU.

Run in the build container only (the reference is mounted read-only at /root/reference and
does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes small ``.npz`` fixtures next to this file.  Nothing here is product code; the
fixtures pin ``oracle/nsa_oracle.py`` (tests/test_oracle_golden.py) and, through the oracle
and directly, the CUDA kernels (tests/test_*_gpu.py).  Every array name says which reference
function produced it.
"""
import os
import sys

REF = os.environ.get("NSA_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
os.environ.setdefault("NSA_DEBUG_LOG", "0")

import numpy as np  # noqa: E402
import torch  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(4)


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"wrote {name}.npz ({len(out)} arrays)")


def gen_meta():
    from nsa.core.block_index import build_block_meta

    arrs = {}
    cfgs = [(200, 32, 16, 64), (2048, 32, 16, 64), (128, 16, 8, 32), (96, 32, 16, 64), (31, 32, 16, 64),
            (64, 4, 2, 4), (70, 8, 8, 16), (4096, 32, 16, 64), (50, 16, 8, 32)]
    for i, (S, l, d, ls) in enumerate(cfgs):
        m = build_block_meta(S, l, d, ls, 16, 512)
        arrs[f"cfg{i}"] = np.array([S, l, d, ls])
        arrs[f"rows{i}"] = m.M_csl_coo_indices[0]
        arrs[f"cols{i}"] = m.M_csl_coo_indices[1]
        arrs[f"vals{i}"] = m.M_csl_coo_values
        arrs[f"ncmp{i}"] = np.array([m.cmp_starts.numel(), m.sel_starts.numel()])
    arrs["n"] = np.array(len(cfgs))
    save("meta", **arrs)


def gen_rope_phi():
    from nsa.core.compress_pool import avg_pool_phi_rope_kv
    from nsa.core.rope import apply_rope

    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 2, 70, 16, generator=g)
    pos = torch.arange(5, 75)
    y = apply_rope(x, pos)
    y8 = apply_rope(x, pos, scale=8.0)
    K = torch.randn(2, 2, 70, 16, generator=g)
    V = torch.randn(2, 2, 70, 8, generator=g)
    Kc, Vc = avg_pool_phi_rope_kv(K, V, 16, 8)
    save("rope_phi", x=x, pos=pos, rope=y, rope_scale8=y8, K_raw=K, V_raw=V, K_cmp=Kc, V_cmp=Vc,
         ld=np.array([16, 8]))


def gen_scores():
    from nsa.core.block_index import build_block_meta
    from nsa.core.selection_scorer import compute_pcmp_all, map_pcmp_to_pslc_batched

    g = torch.Generator().manual_seed(3)
    B, S, G, h, Dk = 2, 200, 2, 3, 16
    l, d, ls = 32, 16, 64
    meta = build_block_meta(S, l, d, ls, 16, 512)
    S_cmp = meta.cmp_starts.numel()
    Q = torch.randn(B, S, G, h, Dk, generator=g)
    Kc = torch.randn(B, G, S_cmp, Dk, generator=g)
    p = compute_pcmp_all(Q, Kc, 1.0 / Dk ** 0.5)
    pslc = map_pcmp_to_pslc_batched(p, meta)
    pgrp = pslc.sum(dim=3)
    # second geometry (showcase block sizes), fewer cmp rows than the meta covers (decode-like)
    meta2 = build_block_meta(128, 16, 8, 32, 8, 64)
    Q2 = torch.randn(1, 4, 2, 4, 8, generator=g)
    Kc2 = torch.randn(1, 2, 9, 8, generator=g)
    p2 = compute_pcmp_all(Q2, Kc2, 1.0 / 8 ** 0.5)
    pslc2 = map_pcmp_to_pslc_batched(p2, meta2)
    save("scores", Q=Q, K_cmp=Kc, p_cmp=p, p_slc=pslc, p_grp=pgrp, cfg=np.array([S, l, d, ls]),
         Q2=Q2, K_cmp2=Kc2, p_slc2=pslc2, p_grp2=pslc2.sum(dim=3), cfg2=np.array([128, 16, 8, 32]))


def _rand_pgrp(g, shape, kind):
    if kind == "rand":
        return torch.rand(shape, generator=g)
    if kind == "softmaxlike":  # many exact zeros, like uncovered blocks
        x = torch.rand(shape, generator=g)
        return torch.where(x > 0.6, x, torch.zeros_like(x))
    if kind == "small":  # tiny values where the 1e-8 bias is effective
        return torch.rand(shape, generator=g) * 1e-6
    raise ValueError(kind)


def gen_select():
    from nsa.core.block_index import build_block_meta
    from nsa.core.selection_scorer import select_topn_ranges, select_topn_ranges_batched

    g = torch.Generator().manual_seed(5)
    arrs = {}
    # ---- decode mode -----------------------------------------------------------------
    dec = []
    i = 0
    for (S_ctx, ls, n, kind) in [(700, 64, 16, "rand"), (2048, 64, 16, "softmaxlike"), (300, 32, 8, "rand"),
                                 (130, 64, 16, "rand"), (64, 64, 16, "rand"), (1500, 64, 5, "small"),
                                 (520, 64, 3, "rand"), (4096, 64, 16, "rand"), (260, 32, 2, "rand")]:
        meta = build_block_meta(S_ctx, ls // 2, ls // 4, ls, n, 512)
        S_sel = meta.sel_starts.numel()
        for t in sorted({0, 1, ls - 2, ls - 1, ls, 2 * ls - 1, 2 * ls, S_ctx // 2, S_ctx - 2, S_ctx - 1}):
            if t < 0 or t >= S_ctx:
                continue
            p = _rand_pgrp(g, (2, 2, S_sel), kind)
            r = select_topn_ranges(p, meta, n, t, True, 2)
            arrs[f"dec_p{i}"] = p
            arrs[f"dec_r{i}"] = r
            arrs[f"dec_c{i}"] = np.array([ls, n, t])
            i += 1
    arrs["dec_n"] = np.array(i)
    # ---- prefill (batched) mode ------------------------------------------------------
    i = 0
    for (S, ls, n, kind) in [(200, 64, 16, "rand"), (700, 64, 16, "rand"), (128, 32, 8, "rand"), (40, 64, 16, "rand"),
                             (100, 64, 16, "rand"), (300, 32, 8, "softmaxlike"), (1200, 64, 16, "softmaxlike"),
                             (512, 64, 4, "small"), (256, 64, 3, "rand"), (192, 64, 2, "rand"), (500, 32, 16, "rand"),
                             (96, 32, 1, "rand"), (640, 64, 10, "rand")]:
        meta = build_block_meta(S, ls // 2, ls // 4, ls, n, 512)
        S_sel = meta.sel_starts.numel()
        p = _rand_pgrp(g, (1, S, 2, S_sel), kind)
        r = select_topn_ranges_batched(p, meta, n, S, True, 2)
        arrs[f"pre_p{i}"] = p
        arrs[f"pre_r{i}"] = r
        arrs[f"pre_c{i}"] = np.array([ls, n, S])
        i += 1
    arrs["pre_n"] = np.array(i)
    save("select", **arrs)


def gen_attention():
    from nsa.core.attention_kernels import (
        grouped_selection_attention_masked,
        sliding_window_attention,
    )
    from nsa.core.block_index import build_block_meta
    from nsa.core.selection_scorer import select_topn_ranges_batched
    from nsa.kernels.flash_wrappers import attention_bgh

    g = torch.Generator().manual_seed(7)
    B, S, G, h, Dk, Dv = 2, 160, 2, 3, 16, 16
    l, d, ls, n, w = 16, 8, 32, 4, 48
    Q = torch.randn(B, S, G, h, Dk, generator=g)
    K = torch.randn(B, G, S, Dk, generator=g)
    V = torch.randn(B, G, S, Dv, generator=g)
    meta = build_block_meta(S, l, d, ls, n, w)
    p = torch.rand(B, S, G, meta.sel_starts.numel(), generator=g)
    ranges = select_topn_ranges_batched(p, meta, n, S, True, 2)
    O_sel = grouped_selection_attention_masked(Q, K, V, ranges)
    O_win = sliding_window_attention(Q, K, V, w)
    # cmp: true softmax through reference primitives: attention_bgh(causal=False) on the first
    # num_cmp(t) compressed tokens (lengths as attention_kernels.py:121)
    S_cmp = meta.cmp_starts.numel()
    Kc = torch.randn(B, G, S_cmp, Dk, generator=g)
    Vc = torch.randn(B, G, S_cmp, Dv, generator=g)
    O_cmp = torch.zeros(B, S, G, h, Dv)
    for t in range(S):
        L = 0 if t + 1 < l else min((t + 1 - l) // d + 1, S_cmp)
        if L > 0:
            O_cmp[:, t] = attention_bgh(Q[:, t].contiguous(), Kc[:, :, :L].contiguous(), Vc[:, :, :L].contiguous(),
                                        causal=False)
    save("attention", Q=Q, K=K, V=V, ranges=ranges, O_sel=O_sel, O_win=O_win, K_cmp=Kc, V_cmp=Vc, O_cmp=O_cmp,
         cfg=np.array([l, d, ls, n, w]))


def gen_gate():
    from nsa.core.nsa_attention import GateMLP

    torch.manual_seed(13)
    gm = GateMLP(16)
    q = torch.randn(50, 16)
    p = gm(q, tau=1.0)
    p_tau = gm(q, tau=0.5)
    with torch.no_grad():
        gm.fc2.bias.copy_(torch.tensor([-1000.0, 1000.0, -1000.0]))
    p_hard = gm(q, tau=1.0)
    save("gate", q=q, fc1_w=gm.fc1.weight, fc1_b=gm.fc1.bias, fc2_w=gm.fc2.weight,
         fc2_b_soft=np.zeros(3, dtype=np.float32), fc2_b_hard=gm.fc2.bias, p=p, p_tau=p_tau, p_hard=p_hard)


def _empty_kv(B, G, dk, dv, meta):
    from nsa.cache.kv_cache import NSA_KV

    z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt)
    return NSA_KV(K_sel=z(B, G, 0, dk), V_sel=z(B, G, 0, dv), K_win=z(B, G, 0, dk), V_win=z(B, G, 0, dv),
                  K_cmp_raw_seq=z(B, G, 0, dk), V_cmp_raw_seq=z(B, G, 0, dv), K_cmp=z(B, G, 0, dk), V_cmp=z(B, G, 0, dv),
                  win_ptr=z(B, G, dt=torch.int32), cmp_emit_next=z(B, G, dt=torch.int32), meta=meta,
                  reads_pred=z(0, dt=torch.int64), reads_act_total=z(0, dt=torch.int64), reads_act_sel=z(0, dt=torch.int64),
                  reads_act_cmp=z(0, dt=torch.int64), reads_act_win=z(0, dt=torch.int64))


def gen_module():
    """Whole-module vectors.  The reference's sel (NSA_FORCE_SEL_MASK=1) and win (batched) routes are
    true softmax; its cmp route is degenerate, so 'intended' vectors swap in a true-softmax cmp
    built from the reference's attention_bgh(causal=False) (SURVEY 8c last row)."""
    os.environ["NSA_PREFILL_BATCHED"] = "1"
    os.environ["NSA_FORCE_SEL_MASK"] = "1"
    import nsa.core.attention_kernels as ak
    import nsa.core.nsa_attention as na
    from nsa.core.block_index import build_block_meta
    from nsa.kernels.flash_wrappers import attention_bgh

    def cmp_true(Q, K_cmp, V_cmp, l, d):
        B, S, G, h, _ = Q.shape
        S_cmp = K_cmp.shape[2]
        out = torch.zeros(B, S, G, h, V_cmp.shape[-1], dtype=V_cmp.dtype)
        for t in range(S):
            L = 0 if t + 1 < l else min((t + 1 - l) // d + 1, S_cmp)
            if L > 0:
                out[:, t] = attention_bgh(Q[:, t].contiguous(), K_cmp[:, :, :L].contiguous(),
                                          V_cmp[:, :, :L].contiguous(), causal=False)
        return out

    torch.manual_seed(1337)
    dim, H, G, dk, dv, l, d, ls, n, w = 64, 4, 2, 16, 16, 16, 8, 32, 4, 40
    B, S = 2, 150
    m = na.NSAAttention(dim, H, G, dk, dv, l=l, d=d, l_sel=ls, n_sel=n, w=w)
    x = torch.randn(B, S, dim)
    meta = build_block_meta(S, l, d, ls, n, w)
    arrs = {"x": x, "cfg": np.array([dim, H, G, dk, dv, l, d, ls, n, w])}
    for k, v in m.state_dict().items():
        arrs["sd__" + k] = v
    # literal reference output (cmp degenerate)
    with torch.no_grad():
        out_lit, kv = m(x, _empty_kv(B, G, dk, dv, meta), prefill=True)
    arrs["out_literal"] = out_lit
    arrs["kv_K_sel"], arrs["kv_V_sel"] = kv.K_sel, kv.V_sel
    arrs["kv_K_win"], arrs["kv_V_win"] = kv.K_win, kv.V_win
    arrs["kv_K_cmp"], arrs["kv_V_cmp"] = kv.K_cmp, kv.V_cmp
    # intended: swap cmp
    orig = ak.batched_causal_attention_compressed_masked
    ak.batched_causal_attention_compressed_masked = cmp_true
    try:
        xg = x.clone().requires_grad_(True)
        out_int, _ = m(xg, _empty_kv(B, G, dk, dv, meta), prefill=True)
        arrs["out_intended"] = out_int
        gsum = torch.randn(out_int.shape, generator=torch.Generator().manual_seed(2))
        (out_int * gsum).sum().backward()
        arrs["grad_out"] = gsum
        arrs["grad_x"] = xg.grad
        for k, p in m.named_parameters():
            if p.grad is not None:
                arrs["grad__" + k] = p.grad
        # per-branch outputs with forced gates (module-level plumbing pins)
        for fb in ("cmp", "sel", "win"):
            m.gate._force_branch = fb
            with torch.no_grad():
                o, _ = m(x, _empty_kv(B, G, dk, dv, meta), prefill=True)
            arrs["out_force_" + fb] = o
        m.gate._force_branch = None
    finally:
        ak.batched_causal_attention_compressed_masked = orig

    # ---- decode: every token goes through a decode step from an empty cache (what NSA_PREFILL_TILE does,
    # nsa_attention.py:1507-1519), so the emission schedule counts absolute tokens.  NOTE: the reference's
    # batched/sequential prefill never fills K_cmp_raw_seq, so "prefill then decode" restarts the emission
    # count at zero; that quirk is documented in DESIGN.md and not reproduced.
    # Intended semantics = attention_bgh(causal=False) + masked selection.
    S0, T = 70, 12
    xs = torch.randn(B, S0 + T, dim, generator=torch.Generator().manual_seed(4))
    orig_bgh = na.attention_bgh
    na.attention_bgh = lambda Q, K, V, causal=True: orig_bgh(Q, K, V, causal=False)
    try:
        with torch.no_grad():
            kv = _empty_kv(B, G, dk, dv, build_block_meta(S0 + T, l, d, ls, n, w))
            outs = []
            for i in range(S0 + T):
                o, kv = m(xs[:, i:i + 1], kv, prefill=False)
                outs.append(o)
        arrs["dec_x"] = xs
        arrs["dec_S0T"] = np.array([S0, T])
        arrs["dec_out_steps"] = torch.cat(outs, dim=1)
        arrs["dec_K_cmp_final"] = kv.K_cmp
        arrs["dec_V_cmp_final"] = kv.V_cmp
        arrs["dec_K_win_final"] = kv.K_win
        arrs["dec_reads_total"] = kv.reads_act_total
    finally:
        na.attention_bgh = orig_bgh
    save("module", **arrs)


def gen_free_functions():
    """The reference's stage-by-stage free functions: selection with other force_init / force_local settings, p_cmp,
    Eq.9 for several geometries, ids -> ranges, the one-query attention primitive."""
    from nsa.core.block_index import build_block_meta
    from nsa.core.selection_scorer import (compute_pcmp_all, convert_indices_to_ranges_batched, convert_indices_to_ranges_batched_v2,
                                           map_pcmp_to_pslc, map_pcmp_to_pslc_batched, select_topn_ranges, select_topn_ranges_batched)
    from nsa.kernels.flash_wrappers import attention_bgh

    g = torch.Generator().manual_seed(23)
    arrs = {}
    flags = [(0, 0), (1, 0), (0, 1), (0, 2), (1, 1), (0, 3), (1, 2)]
    i = 0
    for (S_ctx, ls, n) in [(700, 64, 16), (300, 32, 8), (130, 64, 5), (64, 64, 16), (260, 32, 2)]:
        meta = build_block_meta(S_ctx, ls // 2, ls // 4, ls, n, 512)
        S_sel = meta.sel_starts.numel()
        for (fi, fl) in flags:
            for t in sorted({0, ls - 1, ls, 2 * ls, S_ctx // 2, S_ctx - 1}):
                if t >= S_ctx:
                    continue
                p = torch.rand((1, 2, S_sel), generator=g)
                arrs[f"dec_p{i}"], arrs[f"dec_r{i}"] = p, select_topn_ranges(p, meta, n, t, bool(fi), fl)
                arrs[f"dec_c{i}"] = np.array([ls, n, t, fi, fl])
                i += 1
    arrs["dec_n"] = np.array(i)
    i = 0
    for (S, ls, n) in [(200, 64, 16), (40, 64, 16), (100, 64, 16), (128, 32, 8), (300, 32, 3), (192, 64, 2), (96, 32, 1), (640, 64, 10)]:
        meta = build_block_meta(S, ls // 2, ls // 4, ls, n, 512)
        S_sel = meta.sel_starts.numel()
        for (fi, fl) in flags:
            p = torch.rand((1, S, 2, S_sel), generator=g)
            arrs[f"pre_p{i}"], arrs[f"pre_r{i}"] = p, select_topn_ranges_batched(p, meta, n, S, bool(fi), fl)
            arrs[f"pre_c{i}"] = np.array([ls, n, S, fi, fl])
            i += 1
    arrs["pre_n"] = np.array(i)
    # compute_pcmp_all (selection_scorer.py:42-61)
    Q = torch.randn(2, 20, 2, 3, 16, generator=g)
    Kc = torch.randn(2, 2, 9, 16, generator=g)
    arrs["pcmp_Q"], arrs["pcmp_K"], arrs["pcmp_p"] = Q, Kc, compute_pcmp_all(Q, Kc, 1.0 / 4.0)
    # Eq.9, several geometries (selection_scorer.py:64-116)
    for k, (S, l, d, ls) in enumerate([(512, 32, 16, 64), (256, 16, 8, 32), (64, 4, 2, 4), (96, 8, 8, 16), (200, 32, 16, 64), (300, 64, 16, 64)]):
        meta = build_block_meta(S, l, d, ls, 8, 64)
        S_cmp = meta.cmp_starts.numel()
        pc = torch.rand(1, 7, 2, 3, S_cmp, generator=g)
        arrs[f"map_c{k}"] = np.array([S, l, d, ls])
        arrs[f"map_p{k}"], arrs[f"map_o{k}"] = pc, map_pcmp_to_pslc_batched(pc, meta)
        short = torch.rand(2, 2, 3, max(S_cmp - 3, 1), generator=g)  # fewer compressed rows than the map covers (early decode)
        arrs[f"map_ps{k}"], arrs[f"map_os{k}"] = short, map_pcmp_to_pslc(short, meta)
    arrs["map_n"] = np.array(6)
    # ids -> ranges (selection_scorer.py:380-605)
    for k, (B, S, G, K, ls, hi) in enumerate([(2, 8, 2, 12, 4, 16), (1, 40, 2, 16, 64, 12), (1, 1, 1, 5, 32, 4), (2, 16, 1, 32, 64, 40)]):
        meta = build_block_meta(max(hi * ls, S), ls // 2, ls // 4, ls, 8, 64)
        idx = torch.randint(-1, hi, (B, S, G, K), generator=g)
        idx = torch.sort(torch.where(idx < 0, torch.full_like(idx, 10 ** 6), idx), dim=-1).values
        idx = torch.where(idx == 10 ** 6, torch.full_like(idx, -1), idx)
        arrs[f"i2r_c{k}"] = np.array([ls, meta.sel_starts.numel()])
        arrs[f"i2r_i{k}"] = idx
        arrs[f"i2r_v1_{k}"] = convert_indices_to_ranges_batched(idx, meta, S)
        arrs[f"i2r_v2_{k}"] = convert_indices_to_ranges_batched_v2(idx, meta, S)
    arrs["i2r_n"] = np.array(4)
    # attention_bgh, all keys (flash_wrappers.py:191-282; causal=True with one query row is the degenerate first-key route)
    q = torch.randn(2, 2, 3, 16, generator=g)
    K = torch.randn(2, 2, 37, 16, generator=g)
    V = torch.randn(2, 2, 37, 16, generator=g)
    arrs["bgh_q"], arrs["bgh_K"], arrs["bgh_V"], arrs["bgh_O"] = q, K, V, attention_bgh(q, K, V, causal=False)
    save("free_functions", **arrs)


def gen_phi_mlp():
    """phi="mlp" (learnable depthwise Conv1d, nsa_attention.py:275-291, :1741-1777).  The reference's constructor clears
    phi_v_conv unconditionally (the `self.phi_v_conv = None` of the else-branch is dedented one level, :289-291), so the module as
    shipped asserts on first use; the generating script repairs exactly that attribute and then calls the reference's own
    _phi_apply_seq / _phi_apply_last and a whole prefill + decode loop."""
    import torch.nn as nn
    from nsa.cache.kv_cache import NSA_KV  # noqa: F401
    from nsa.core.block_index import build_block_meta
    from nsa.core.nsa_attention import NSAAttention

    torch.manual_seed(41)
    dim, H, G, dk, dv, l, d, ls, n, w = 64, 4, 2, 16, 16, 8, 4, 16, 4, 32
    m = NSAAttention(dim=dim, n_heads=H, n_kv_groups=G, d_k=dk, d_v=dv, l=l, d=d, l_sel=ls, n_sel=n, w=w, phi="mlp")
    assert m.phi_k_conv is not None and m.phi_v_conv is None  # the bug
    m.phi_v_conv = nn.Conv1d(dv, dv, kernel_size=l, stride=d, groups=dv, bias=False)
    with torch.no_grad():
        m.phi_k_conv.weight.copy_(torch.randn_like(m.phi_k_conv.weight) * 0.3)
        m.phi_v_conv.weight.copy_(torch.randn_like(m.phi_v_conv.weight) * 0.3)
    arrs = {"cfg": np.array([dim, H, G, dk, dv, l, d, ls, n, w])}
    for k, v in m.state_dict().items():
        arrs["sd__" + k] = v
    K_raw, V_raw = torch.randn(2, G, 37, dk), torch.randn(2, G, 37, dv)
    with torch.no_grad():
        Kc, Vc = m._phi_apply_seq(K_raw, V_raw, torch.arange(37))
        Kl, Vl = m._phi_apply_last(K_raw[:, :, 20:28], V_raw[:, :, 20:28], torch.arange(20, 28))
    arrs.update(K_raw=K_raw, V_raw=V_raw, K_cmp=Kc, V_cmp=Vc, K_last=Kl, V_last=Vl)
    # whole module: prefill S0 tokens, then decode (emission through _phi_apply_last); only the compressed stream is compared
    # (attention outputs on the reference's default routes are the degenerate first-key ones)
    x = torch.randn(1, 30, dim)
    kv = _empty_kv(1, G, dk, dv, build_block_meta(64, l, d, ls, n, w))
    with torch.no_grad():
        _, kv = m(x[:, :18], kv, prefill=True)
        arrs["K_cmp_after_prefill"] = kv.K_cmp.clone()
        # the reference's prefill does not record the raw stream (its decode would restart the emission count): seed it
        kv.K_cmp_raw_seq = m._shape_kv(m.W_K_cmp(x[:, :18]), 1, 18)
        kv.V_cmp_raw_seq = m._shape_kv(m.W_V_cmp(x[:, :18]), 1, 18)
        for i in range(18, 30):
            _, kv = m(x[:, i:i + 1], kv, prefill=False)
    arrs.update(x=x, K_cmp_final=kv.K_cmp, V_cmp_final=kv.V_cmp)
    save("phi_mlp", **arrs)


if __name__ == "__main__":
    if len(sys.argv) > 1:  # python tests/golden/make_golden.py free_functions [...]: regenerate only the named fixtures
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_free_functions()
    gen_phi_mlp()
    gen_meta()
    gen_rope_phi()
    gen_scores()
    gen_select()
    gen_attention()
    gen_gate()
    gen_module()
