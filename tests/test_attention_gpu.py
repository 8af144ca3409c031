"""Scoring, branch attention, gate and fused hot-path parity against the oracle and the reference-generated
golden vectors.  Tolerances (the reference's own vocabulary in its GPU tests): fp32 max-abs <= 5e-5;
bf16 inputs vs the fp32 oracle max-abs <= 2e-2 and MAE <= 1e-3 per branch and for the gated output; gradients
relative error <= 5e-3 (fp32) / 3e-2 (bf16)."""

import pytest
import torch

from conftest import T, load_golden
from oracle import nsa_oracle as O

pytestmark = pytest.mark.gpu

FP32_ATOL = 5e-5
BF16_MAXABS, BF16_MAE = 2e-2, 1e-3


def _ops():
    from nsa_vibe_b200 import ops
    return ops


def _rel(a, b):
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def test_scores_golden_fp32():
    ops = _ops()
    g = load_golden("scores")
    S, l, d, ls = [int(v) for v in g["cfg"]]
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=16, w=512)
    got = ops.score_pgrp(T(g["Q"]).cuda(), T(g["K_cmp"]).cuda(), cfg).cpu()
    assert torch.allclose(got, T(g["p_grp"]), atol=2e-6), (got - T(g["p_grp"])).abs().max()
    # early-decode geometry: fewer compressed rows than the Eq.9 map covers
    S2, l2, d2, ls2 = [int(v) for v in g["cfg2"]]
    cfg2 = ops.NSAConfig(l=l2, d=d2, l_sel=ls2, n_sel=8, w=64)
    got2 = ops.score_pgrp(T(g["Q2"]).cuda(), T(g["K_cmp2"]).cuda(), cfg2, S_sel=4).cpu()
    assert torch.allclose(got2, T(g["p_grp2"]), atol=2e-6)


@pytest.mark.parametrize("norm", ["full_row", "causal"])
def test_scores_vs_oracle_m7c_dims(norm):
    ops = _ops()
    B, S, G, h, Dk, l, d, ls = 1, 512, 2, 6, 64, 32, 16, 64
    gen = torch.Generator().manual_seed(1)
    Q = torch.randn(B, S, G, h, Dk, generator=gen)
    Kc = torch.randn(B, G, O.num_cmp_blocks(S, l, d), Dk, generator=gen)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, norm_mode=ops.NORM_CAUSAL if norm == "causal" else ops.NORM_FULL_ROW)
    want = O.prefill_scores(Q, Kc, l, d, ls, 16, 512, norm)
    got = ops.score_pgrp(Q.cuda(), Kc.cuda(), cfg).cpu()
    assert torch.allclose(got, want, atol=5e-6), (got - want).abs().max()
    # fused scoring+selection agrees with selecting on the oracle's scores except at fp32 near-ties
    r = ops.score_select(Q.cuda(), Kc.cuda(), cfg, mode=0).cpu()
    ok, bad = O.ranges_equivalent(r, O.select_ranges_prefill(want, ls, 16, S))
    assert bad <= 2, f"{bad} of {B * S * G} rows differ"
    r_same = ops.select_ranges_prefill(got.cuda(), ls, 16, S).cpu()
    assert torch.equal(r, r_same)  # fused path == standalone kernels on the same scores


def test_branches_golden_fp32():
    ops = _ops()
    g = load_golden("attention")
    l, d, ls, n, w = [int(v) for v in g["cfg"]]
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
    Q, K, V = T(g["Q"]).cuda(), T(g["K"]).cuda(), T(g["V"]).cuda()
    o = ops.branch_attention(ops.BR_SEL, Q, K, V, cfg, T(g["ranges"]).cuda()).cpu()
    assert torch.allclose(o, T(g["O_sel"]), atol=FP32_ATOL), (o - T(g["O_sel"])).abs().max()
    o = ops.branch_attention(ops.BR_WIN, Q, K, V, cfg).cpu()
    assert torch.allclose(o, T(g["O_win"]), atol=FP32_ATOL), (o - T(g["O_win"])).abs().max()
    o = ops.branch_attention(ops.BR_CMP, Q, T(g["K_cmp"]).cuda(), T(g["V_cmp"]).cuda(), cfg).cpu()
    assert torch.allclose(o, T(g["O_cmp"]), atol=FP32_ATOL), (o - T(g["O_cmp"])).abs().max()


def test_sel_empty_rows_zero_no_nan():
    ops = _ops()
    Q = torch.randn(1, 3, 1, 2, 16).cuda()
    K = torch.randn(1, 1, 10, 16).cuda()
    V = torch.randn(1, 1, 10, 16).cuda()
    r = torch.zeros(1, 3, 1, 2, 2, dtype=torch.int32)
    r[0, 1, 0, 0] = torch.tensor([2, 5])
    r[0, 2, 0, 1] = torch.tensor([7, 3])  # end < start: skipped like the reference's consumers do
    o, lse = ops.branch_attention(ops.BR_SEL, Q, K, V, ops.NSAConfig(), r.cuda(), return_lse=True)
    assert torch.isfinite(o).all() and torch.all(o[0, 0] == 0) and torch.all(o[0, 2] == 0)
    assert torch.isinf(lse[0, 0]).all() and torch.isfinite(lse[0, 1]).all()


def _rand_case(B, S, G, h, Dk, Dv, l, d, ls, n, w, seed, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen)
    Q, K_sel, V_sel, K_win, V_win = r(B, S, G, h, Dk), r(B, G, S, Dk), r(B, G, S, Dv), r(B, G, S, Dk), r(B, G, S, Dv)
    S_cmp = O.num_cmp_blocks(S, l, d)
    K_cmp, V_cmp = r(B, G, S_cmp, Dk), r(B, G, S_cmp, Dv)
    hid = max(1, Dk // 2)
    gate = (r(hid, Dk) * 0.3, r(hid) * 0.1, r(3, hid) * 0.5, r(3) * 0.1)
    ts = [t.to(dtype).float() for t in (Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp)]  # values representable in dtype
    return ts, gate


@pytest.mark.parametrize("dims", [
    dict(B=2, S=200, G=2, h=4, Dk=16, Dv=16, l=16, d=8, ls=32, n=8, w=64),    # train_showcase shapes
    dict(B=1, S=300, G=2, h=6, Dk=64, Dv=64, l=32, d=16, ls=64, n=16, w=128),  # m7c head dims
    dict(B=1, S=130, G=1, h=3, Dk=32, Dv=16, l=8, d=4, ls=16, n=5, w=20),     # ragged / Dk != Dv
])
@pytest.mark.parametrize("sel_mode", [0, 1])
def test_prefill_core_fp32_vs_oracle(dims, sel_mode):
    ops = _ops()
    D = dims
    ts, gate = _rand_case(D["B"], D["S"], D["G"], D["h"], D["Dk"], D["Dv"], D["l"], D["d"], D["ls"], D["n"], D["w"], seed=3)
    cfg = ops.NSAConfig(l=D["l"], d=D["d"], l_sel=D["ls"], n_sel=D["n"], w=D["w"])
    Oc, ranges, gates = ops.prefill_core(*[t.cuda() for t in ts], tuple(t.cuda() for t in gate), cfg, sel_mode=sel_mode)
    # oracle on the ranges the kernel chose (scores differ only at fp32 rounding; index parity is tested separately)
    want = O.prefill_core(*ts, gate, l=D["l"], d=D["d"], l_sel=D["ls"], n_sel=D["n"], w=D["w"], ranges=ranges.cpu())
    assert torch.allclose(gates.cpu(), want["gates"], atol=1e-5)
    assert torch.allclose(Oc.cpu(), want["O"], atol=FP32_ATOL), (Oc.cpu() - want["O"]).abs().max()
    # and the ranges themselves against the oracle's rule on the oracle's scores
    pg = O.prefill_scores(ts[0], ts[5], D["l"], D["d"], D["ls"], D["n"], D["w"])
    if sel_mode == 0:
        wr = O.select_ranges_prefill(pg, D["ls"], D["n"], D["S"])
    else:
        wr = torch.stack([O.select_ranges_decode(pg[:, t], D["ls"], D["n"], t) for t in range(D["S"])], dim=1)
    ok, bad = O.ranges_equivalent(ranges.cpu(), wr)
    assert bad <= 1, f"{bad} rows differ"


def test_prefill_core_bf16_vs_fp32_oracle():
    ops = _ops()
    D = dict(B=2, S=384, G=2, h=6, Dk=64, Dv=64, l=32, d=16, ls=64, n=16, w=128)
    ts, gate = _rand_case(D["B"], D["S"], D["G"], D["h"], D["Dk"], D["Dv"], D["l"], D["d"], D["ls"], D["n"], D["w"], seed=5,
                          dtype=torch.bfloat16)
    cfg = ops.NSAConfig(l=D["l"], d=D["d"], l_sel=D["ls"], n_sel=D["n"], w=D["w"])
    Oc, ranges, gates = ops.prefill_core(*[t.cuda().bfloat16() for t in ts], tuple(t.cuda() for t in gate), cfg, sel_mode=0)
    want = O.prefill_core(*ts, gate, l=D["l"], d=D["d"], l_sel=D["ls"], n_sel=D["n"], w=D["w"], ranges=ranges.cpu())
    err = (Oc.float().cpu() - want["O"]).abs()
    assert err.max() <= BF16_MAXABS and err.mean() <= BF16_MAE, (err.max(), err.mean())
    for br, name in ((ops.BR_CMP, "O_cmp"), (ops.BR_SEL, "O_sel"), (ops.BR_WIN, "O_win")):
        K, V = {0: (ts[5], ts[6]), 1: (ts[1], ts[2]), 2: (ts[3], ts[4])}[br]
        ob = ops.branch_attention(br, ts[0].cuda().bfloat16(), K.cuda().bfloat16(), V.cuda().bfloat16(), cfg,
                                  ranges if br == 1 else None)
        err = (ob.float().cpu() - want[name]).abs()
        assert err.max() <= BF16_MAXABS and err.mean() <= BF16_MAE, (name, err.max(), err.mean())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 5e-3), (torch.bfloat16, 3e-2)])
def test_prefill_core_backward_vs_oracle_autograd(dtype, tol):
    ops = _ops()
    D = dict(B=1, S=160, G=2, h=3, Dk=32, Dv=32, l=16, d=8, ls=32, n=6, w=48)
    ts, gate = _rand_case(D["B"], D["S"], D["G"], D["h"], D["Dk"], D["Dv"], D["l"], D["d"], D["ls"], D["n"], D["w"], seed=9,
                          dtype=dtype)
    cfg = ops.NSAConfig(l=D["l"], d=D["d"], l_sel=D["ls"], n_sel=D["n"], w=D["w"])
    dev = [t.cuda().to(dtype).requires_grad_(True) for t in ts]
    gdev = tuple(t.cuda().requires_grad_(True) for t in gate)
    Oc, ranges, _ = ops.prefill_core(*dev, gdev, cfg, sel_mode=0)
    dO = torch.randn(Oc.shape, generator=torch.Generator().manual_seed(1))
    (Oc.float() * dO.cuda()).sum().backward()
    cpu = [t.clone().requires_grad_(True) for t in ts]
    gcpu = tuple(t.clone().requires_grad_(True) for t in gate)
    want = O.prefill_core(*cpu, gcpu, l=D["l"], d=D["d"], l_sel=D["ls"], n_sel=D["n"], w=D["w"], ranges=ranges.cpu())
    (want["O"] * dO).sum().backward()
    names = ["Q", "K_sel", "V_sel", "K_win", "V_win", "K_cmp", "V_cmp"]
    for nme, a, b in zip(names, dev, cpu):
        assert _rel(a.grad.float().cpu(), b.grad) <= tol, (nme, _rel(a.grad.float().cpu(), b.grad))
    for nme, a, b in zip(["fc1_w", "fc1_b", "fc2_w", "fc2_b"], gdev, gcpu):
        assert _rel(a.grad.float().cpu(), b.grad) <= max(tol, 1e-2), (nme, _rel(a.grad.float().cpu(), b.grad))


def test_branch_backward_matches_reference_formula_edges():
    """Empty / adjacent / long ranges (the reference's test_selection_backward_edges.py:27-91 cases) against
    autograd through the oracle."""
    ops = _ops()
    gen = torch.Generator().manual_seed(2)
    B, S, G, h, D, Skv = 1, 4, 1, 2, 16, 300
    Q = torch.randn(B, S, G, h, D, generator=gen)
    K = torch.randn(B, G, Skv, D, generator=gen)
    V = torch.randn(B, G, Skv, D, generator=gen)
    r = torch.zeros(B, S, G, 3, 2, dtype=torch.int32)
    r[0, 1, 0, 0] = torch.tensor([0, 8]); r[0, 1, 0, 1] = torch.tensor([8, 16])      # adjacent
    r[0, 2, 0, 0] = torch.tensor([5, 6])                                               # single key
    r[0, 3, 0, 0] = torch.tensor([0, 128]); r[0, 3, 0, 2] = torch.tensor([140, 300])  # long L, gap
    q, k, v = (t.cuda().requires_grad_(True) for t in (Q, K, V))
    o = ops.branch_attention(ops.BR_SEL, q, k, v, ops.NSAConfig(), r.cuda())
    dO = torch.randn(o.shape, generator=gen)
    (o * dO.cuda()).sum().backward()
    qc, kc, vc = (t.clone().requires_grad_(True) for t in (Q, K, V))
    oc, _ = O.sel_attention(qc, kc, vc, r)
    (oc * dO).sum().backward()
    assert torch.allclose(o.detach().cpu(), oc.detach(), atol=FP32_ATOL)
    for a, b in ((q, qc), (k, kc), (v, vc)):
        assert torch.allclose(a.grad.cpu(), b.grad, atol=1e-4), (a.grad.cpu() - b.grad).abs().max()
    assert torch.all(q.grad[0, 0] == 0)  # empty row: no gradient


def test_gate_kernel_golden():
    ops = _ops()
    g = load_golden("gate")
    q = T(g["q"]).cuda()[None, :, None, None, :]  # [1,50,1,1,16]: one head, so q_gp == q
    args = (T(g["fc1_w"]).cuda(), T(g["fc1_b"]).cuda(), T(g["fc2_w"]).cuda())
    p = ops.gate_forward(q, (*args, T(g["fc2_b_soft"]).cuda()), ops.NSAConfig()).cpu()[0, :, 0]
    assert torch.allclose(p, T(g["p"]), atol=1e-6)
    p = ops.gate_forward(q, (*args, T(g["fc2_b_soft"]).cuda()), ops.NSAConfig(gate_tau=0.5)).cpu()[0, :, 0]
    assert torch.allclose(p, T(g["p_tau"]), atol=1e-6)
    p = ops.gate_forward(q, (*args, T(g["fc2_b_hard"]).cuda()), ops.NSAConfig()).cpu()[0, :, 0]
    assert torch.equal(p, T(g["p_hard"]))
    p = ops.gate_forward(q, None, ops.NSAConfig(gate_mode=ops.GATE_UNIFORM)).cpu()
    assert torch.allclose(p, torch.full_like(p, 1 / 3))


def test_decode_core_vs_oracle():
    ops = _ops()
    B, G, h, Dk, Dv, l, d, ls, n, w = 2, 2, 4, 32, 32, 16, 8, 32, 6, 40
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
    gen = torch.Generator().manual_seed(11)
    r = lambda *s: torch.randn(*s, generator=gen)
    cap = 256
    for t in (0, 5, 15, 16, 31, 32, 47, 100, 200):
        n_tok = t + 1
        S_cmp = O.num_cmp_blocks(n_tok, l, d)
        q = r(B, G, h, Dk)
        K_sel, V_sel, K_win, V_win = r(B, G, cap, Dk), r(B, G, cap, Dv), r(B, G, cap, Dk), r(B, G, cap, Dv)
        K_cmp, V_cmp = r(B, G, 64, Dk), r(B, G, 64, Dv)
        gate = (r(16, Dk) * 0.3, r(16) * 0.1, r(3, 16) * 0.5, r(3) * 0.1)
        lo = max(0, n_tok - w)
        want = O.decode_core(q, K_sel[:, :, :n_tok], V_sel[:, :, :n_tok], K_win[:, :, lo:n_tok], V_win[:, :, lo:n_tok],
                             K_cmp[:, :, :S_cmp], V_cmp[:, :, :S_cmp], gate, l=l, d=d, l_sel=ls, n_sel=n, w=w)
        rg = torch.empty(B, G, n, 2, dtype=torch.int32, device="cuda")
        got = ops.decode_core(q[:, None].cuda(), K_sel.cuda(), V_sel.cuda(), K_win.cuda(), V_win.cuda(), K_cmp.cuda(),
                              V_cmp.cuda(), tuple(x.cuda() for x in gate), cfg, t=t, S_sel_kv=n_tok, S_win_kv=n_tok,
                              win_off=0, S_cmp=S_cmp, ranges_out=rg)
        assert torch.equal(rg.cpu(), want["ranges"]), f"t={t}"
        assert torch.allclose(got[:, 0].cpu(), want["O"], atol=FP32_ATOL), (t, (got[:, 0].cpu() - want["O"]).abs().max())


def test_many_short_ranges_beyond_16_blocks_are_served_not_truncated():
    """the tcgen05 selected-branch kernels hold <= 16 blocks of 64 keys per row.  Caller-supplied ranges that break
    that (here 20 disjoint 1-token ranges) must not be silently truncated: ops checks them (nsa_ranges_max_blocks) and runs
    the SIMT kernel, forward and backward."""
    ops = _ops()
    B, S, G, h, D, S_kv, K = 1, 8, 2, 6, 64, 256, 20
    gen = torch.Generator().manual_seed(5)
    r16 = lambda *s: torch.randn(*s, generator=gen).bfloat16().float()
    Q, Kk, V = r16(B, S, G, h, D), r16(B, G, S_kv, D), r16(B, G, S_kv, D)
    ranges = torch.zeros(B, S, G, K, 2, dtype=torch.int32)
    for s in range(S):
        for g in range(G):
            starts = torch.arange(K) * 11 + s + g          # disjoint, unaligned 1-token ranges
            ranges[0, s, g, :, 0] = starts
            ranges[0, s, g, :, 1] = starts + 1
    cfg = ops.NSAConfig()
    assert ops.ranges_max_blocks(ranges.cuda(), S_kv) == K
    want = O.sel_attention(Q, Kk, V, ranges)[0]
    dev = [t.cuda().bfloat16().requires_grad_(True) for t in (Q, Kk, V)]
    got = ops.branch_attention(ops.BR_SEL, *dev, cfg, ranges.cuda())
    assert (got.float().cpu() - want).abs().max() <= BF16_MAXABS
    dO = r16(*got.shape)
    (got.float() * dO.cuda()).sum().backward()
    cpu = [t.clone().requires_grad_(True) for t in (Q, Kk, V)]
    (O.sel_attention(*cpu, ranges)[0] * dO).sum().backward()
    for a, b in zip(dev, cpu):
        assert _rel(a.grad.float().cpu(), b.grad) <= 3e-2
    # ranges within the invariant stay on the tensor-core path and agree too
    assert ops.ranges_max_blocks(ranges[..., :16, :].contiguous().cuda(), S_kv) == 16
    with pytest.raises(RuntimeError, match="more than 16 blocks"):
        ops.sel_attention_blockmajor(dev[0].detach(), dev[1].detach(), dev[2].detach(), cfg, ranges.cuda())
