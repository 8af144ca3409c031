"""tcgen05 / TMA kernels against the SIMT kernels and the fp32 oracle (bf16 tolerance: max-abs 2e-2, MAE 1e-3)."""
import pytest
import torch

from oracle import nsa_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from nsa_vibe_b200 import ops
    return ops


def _case(B, S, G, h, l, d, ls, n, w, seed, dtype):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen).to(dtype).float()
    S_cmp = O.num_cmp_blocks(S, l, d)
    return [r(B, S, G, h, 64), r(B, G, S, 64), r(B, G, S, 64), r(B, G, S, 64), r(B, G, S, 64), r(B, G, S_cmp, 64), r(B, G, S_cmp, 64)]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sel_mode,S,h", [(0, 700, 6), (1, 700, 6), (0, 1500, 4), (1, 333, 8), (0, 64, 1)])
def test_sel_branch_tc_vs_simt_vs_oracle(dtype, sel_mode, S, h):
    ops = _ops()
    B, G, l, d, ls, n, w = 2, 2, 32, 16, 64, 16, 512
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=S + h, dtype=dtype)
    cfg_tc = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    cfg_simt = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_SIMT)
    Q, K, V = (t.cuda().to(dtype) for t in ts[:3])
    ranges = ops.score_select(Q, ts[5].cuda().to(dtype), cfg_simt, mode=sel_mode)
    o_tc, lse_tc = ops.branch_attention(ops.BR_SEL, Q, K, V, cfg_tc, ranges, return_lse=True)
    o_si, lse_si = ops.branch_attention(ops.BR_SEL, Q, K, V, cfg_simt, ranges, return_lse=True)
    want, lse_w = O.sel_attention(ts[0], ts[1], ts[2], ranges.cpu())
    err = (o_tc.float().cpu() - want).abs()
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (err.max(), err.mean())
    assert (o_tc.float() - o_si.float()).abs().max() <= 2e-2
    fin = torch.isfinite(lse_w)
    assert torch.equal(torch.isfinite(lse_tc.cpu()), fin)
    assert (lse_tc.cpu()[fin] - lse_w[fin]).abs().max() <= 2e-2
    empty = ~fin.any(dim=-1)
    if empty.any():
        assert torch.all(o_tc.float().cpu()[empty] == 0)


def test_prefill_core_tc_composition_bf16():
    """nsa_prefill_fwd with the selected branch on tensor cores (auto dispatch) == oracle."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 1, 900, 2, 6, 32, 16, 64, 16, 512
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=4, dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(1)
    gate = (torch.randn(32, 64, generator=gen) * 0.3, torch.randn(32, generator=gen) * 0.1, torch.randn(3, 32, generator=gen) * 0.5,
            torch.zeros(3))
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
    Oc, ranges, gates = ops.prefill_core(*[t.cuda().bfloat16() for t in ts], tuple(x.cuda() for x in gate), cfg, sel_mode=0)
    want = O.prefill_core(*ts, gate, l=l, d=d, l_sel=ls, n_sel=n, w=w, ranges=ranges.cpu())
    err = (Oc.float().cpu() - want["O"]).abs()
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (err.max(), err.mean())
    assert torch.allclose(gates.cpu(), want["gates"], atol=1e-3)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("norm", ["full_row", "causal"])
@pytest.mark.parametrize("S,h,t0", [(700, 6, 0), (2064, 6, 0), (333, 8, 0), (96, 1, 0), (1200, 4, 0), (48, 16, 0)])
def test_score_tc_vs_oracle(dtype, norm, S, h, t0):
    """tcgen05 scorer (Q.K_cmp^T -> softmax -> Eq.9 -> Eq.10) against the fp32 oracle on the same 16-bit inputs.
    Tolerance: p_grp entries are probabilities summed over h heads; max-abs 2e-5 * h (fp32 softmax of exact products)."""
    ops = _ops()
    B, G, l, d, ls, n, w = 2, 2, 32, 16, 64, 16, 512
    gen = torch.Generator().manual_seed(S * 7 + h)
    Q = torch.randn(B, S, G, h, 64, generator=gen).to(dtype).float()
    Kc = torch.randn(B, G, O.num_cmp_blocks(S, l, d), 64, generator=gen).to(dtype).float()
    nm = ops.NORM_CAUSAL if norm == "causal" else ops.NORM_FULL_ROW
    cfg_tc = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC, norm_mode=nm)
    cfg_si = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_SIMT, norm_mode=nm)
    want = O.prefill_scores(Q, Kc, l, d, ls, n, w, norm)
    got = ops.score_pgrp(Q.cuda().to(dtype), Kc.cuda().to(dtype), cfg_tc).cpu()
    simt = ops.score_pgrp(Q.cuda().to(dtype), Kc.cuda().to(dtype), cfg_si).cpu()
    assert got.shape == want.shape
    assert torch.isfinite(got).all()
    assert (got - want).abs().max() <= 2e-5 * h, (got - want).abs().max()
    assert (got - simt).abs().max() <= 2e-5 * h
    # rows sum to (number of heads with any key) like the reference's p_grp
    assert torch.allclose(got.sum(-1), want.sum(-1), atol=1e-4 * h)
    # fused scoring + selection == standalone selection on the same scores; near-ties only vs the oracle's scores
    r = ops.score_select(Q.cuda().to(dtype), Kc.cuda().to(dtype), cfg_tc, mode=0).cpu()
    assert torch.equal(r, ops.select_ranges_prefill(got.cuda(), ls, n, S).cpu())
    _assert_only_near_ties(r, O.select_ranges_prefill(want, ls, n, S), want, got, ls, n)


def _assert_only_near_ties(r_gpu, r_oracle, p_oracle, p_gpu, l_sel, n_sel, t0=0, max_frac=0.005):
    """The fused scoring + selection may differ from the selection on the oracle's fp32 scores ONLY where the scores tie within
    the scorer's own error: every differing row is checked (O.selection_difference_is_near_tie with delta = that row's
    max |p_grp_gpu - p_grp_oracle| over the columns the row can select), and such rows stay below max_frac of all rows."""
    B, S, G = r_gpu.shape[:3]
    bad = 0
    for b in range(B):
        for s in range(S):
            for g in range(G):
                if O.nonempty_ranges(r_gpu[b, s, g].tolist()) == O.nonempty_ranges(r_oracle[b, s, g].tolist()):
                    continue
                bad += 1
                t = t0 + s
                nv = min((t + 1) // l_sel, p_oracle.shape[-1])
                delta = float((p_gpu[b, s, g, :nv] - p_oracle[b, s, g, :nv]).abs().max()) if nv else 0.0
                assert O.selection_difference_is_near_tie(p_oracle[b, s, g], r_gpu[b, s, g].tolist(), r_oracle[b, s, g].tolist(), l_sel,
                                                          n_sel, t, delta), (
                    f"row (b={b}, t={t}, g={g}) selects different blocks and the scores do not tie within 2*{delta:.2e}")
    assert bad <= max(2, int(B * S * G * max_frac)), f"{bad} of {B * S * G} rows differ"


def test_score_tc_chunked_t0():
    """Rows t0..t0+S-1 of a longer sequence (chunked prefill): same scores as the corresponding rows of the full run."""
    ops = _ops()
    B, G, h, l, d, ls, n, w = 1, 2, 6, 32, 16, 64, 16, 512
    S_full, t0, S = 1024, 640, 200
    gen = torch.Generator().manual_seed(11)
    Q = torch.randn(B, S_full, G, h, 64, generator=gen).bfloat16()
    Kc = torch.randn(B, G, O.num_cmp_blocks(S_full, l, d), 64, generator=gen).bfloat16()
    for nm in (ops.NORM_FULL_ROW, ops.NORM_CAUSAL):
        cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC, norm_mode=nm)
        full = ops.score_pgrp(Q.cuda(), Kc.cuda(), cfg)
        part = ops.score_pgrp(Q[:, t0:t0 + S].contiguous().cuda(), Kc.cuda(), cfg, t0=t0, S_sel=full.shape[-1])
        assert torch.equal(part, full[:, t0:t0 + S])


def _check_branch(o, lse, want, lse_w):
    err = (o.float().cpu() - want).abs()
    assert torch.isfinite(o.float()).all()
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (err.max(), err.mean())
    fin = torch.isfinite(lse_w)
    assert torch.equal(torch.isfinite(lse.cpu()), fin)
    if fin.any():
        assert (lse.cpu()[fin] - lse_w[fin]).abs().max() <= 2e-2
    empty = ~fin
    if empty.any():
        assert torch.all(o.float().cpu()[empty] == 0)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("S,h,w", [(700, 6, 512), (1500, 4, 512), (333, 8, 64), (40, 1, 512), (2100, 6, 300), (130, 16, 1),
                                   (3300, 6, 512), (3201, 4, 100)])  # the last two: >= 4*TOK*148 rows -> 4 M-tiles x 64-key tiles
def test_dense_branches_tc_vs_oracle(dtype, S, h, w):
    """tcgen05 dense ranged attention (cmp and win branches) against the fp32 oracle on the same 16-bit inputs.
    Tolerance: max-abs 2e-2, MAE 1e-3 (bf16 P and O rounding)."""
    ops = _ops()
    B, G, l, d, ls, n = 2, 2, 32, 16, 64, 16
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=S + h + w, dtype=dtype)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    Q = ts[0].cuda().to(dtype)
    o, lse = ops.branch_attention(ops.BR_WIN, Q, ts[3].cuda().to(dtype), ts[4].cuda().to(dtype), cfg, return_lse=True)
    _check_branch(o, lse, *O.win_attention(ts[0], ts[3], ts[4], w))
    o, lse = ops.branch_attention(ops.BR_CMP, Q, ts[5].cuda().to(dtype), ts[6].cuda().to(dtype), cfg, return_lse=True)
    _check_branch(o, lse, *O.cmp_attention(ts[0], ts[5], ts[6], l, d))


def test_dense_branches_tc_chunked_rows():
    """Query rows t0..t0+S-1 against full caches (chunked prefill): equals the same rows of the full run."""
    ops = _ops()
    B, G, h, l, d, ls, n, w = 1, 2, 6, 32, 16, 64, 16, 512
    S_full, t0, S = 1400, 777, 300
    ts = _case(B, S_full, G, h, l, d, ls, n, w, seed=5, dtype=torch.bfloat16)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    dev = [t.cuda().bfloat16() for t in ts]
    Qp = dev[0][:, t0:t0 + S].contiguous()
    for br, K, V in ((ops.BR_WIN, dev[3], dev[4]), (ops.BR_CMP, dev[5], dev[6])):
        full, lse_f = ops.branch_attention(br, dev[0], K, V, cfg, return_lse=True)
        part, lse_p = ops.branch_attention(br, Qp, K, V, cfg, t0=t0, return_lse=True)
        assert (part.float() - full[:, t0:t0 + S].float()).abs().max() <= 1e-2
        assert (lse_p - lse_f[:, t0:t0 + S]).abs().max() <= 1e-3


def test_prefill_core_all_tc_bf16_m7c():
    """nsa_prefill_fwd with scoring and all three branches on tensor cores == oracle (given the same ranges)."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 2, 2048, 2, 6, 32, 16, 64, 16, 512
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=9, dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(2)
    gate = (torch.randn(32, 64, generator=gen) * 0.3, torch.randn(32, generator=gen) * 0.1, torch.randn(3, 32, generator=gen) * 0.5,
            torch.zeros(3))
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    Oc, ranges, gates = ops.prefill_core(*[t.cuda().bfloat16() for t in ts], tuple(x.cuda() for x in gate), cfg, sel_mode=0)
    want = O.prefill_core(*ts, gate, l=l, d=d, l_sel=ls, n_sel=n, w=w, ranges=ranges.cpu())
    err = (Oc.float().cpu() - want["O"]).abs()
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (err.max(), err.mean())
    pg = O.prefill_scores(ts[0], ts[5], l, d, ls, n, w)
    pg_gpu = ops.score_pgrp(ts[0].cuda().bfloat16(), ts[5].cuda().bfloat16(), cfg).cpu()
    _assert_only_near_ties(ranges.cpu(), O.select_ranges_prefill(pg, ls, n, S), pg, pg_gpu, ls, n)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("t,h,w,win_off", [(4095, 6, 512, 0), (1000, 6, 512, 0), (700, 4, 64, 200), (63, 8, 512, 0), (40, 2, 512, 0),
                                            (2500, 6, 300, 0), (129, 1, 20, 100)])
def test_decode_step_tc_vs_oracle(dtype, t, h, w, win_off):
    """Fused decode step on the tcgen05 gather kernel (scoring + decode-rule selection + gate + cmp/sel/win attention +
    combine in one launch) against the fp32 oracle on the same 16-bit inputs.  Ranges must be equivalent except at fp32
    near-ties of p_grp; rows with the same ranges must agree within max-abs 2e-2 / MAE 1e-3."""
    ops = _ops()
    B, G, l, d, ls, n = 5, 2, 32, 16, 64, 16
    n_tok = t + 1
    cap = n_tok + 37
    gen = torch.Generator().manual_seed(t + h)
    r = lambda *s: torch.randn(*s, generator=gen).to(dtype).float()
    S_cmp = O.num_cmp_blocks(n_tok, l, d)
    q = r(B, G, h, 64)
    K_sel, V_sel = r(B, G, cap, 64), r(B, G, cap, 64)
    n_win = n_tok - win_off                       # the window cache holds tokens win_off .. t
    K_win, V_win = r(B, G, n_win + 5, 64), r(B, G, n_win + 5, 64)
    K_cmp, V_cmp = r(B, G, S_cmp + 3, 64), r(B, G, S_cmp + 3, 64)
    gate = (torch.randn(32, 64, generator=gen) * 0.3, torch.randn(32, generator=gen) * 0.1, torch.randn(3, 32, generator=gen) * 0.5,
            torch.randn(3, generator=gen) * 0.1)
    lo = max(0, n_tok - w) - win_off
    assert lo >= 0
    want = O.decode_core(q, K_sel[:, :, :n_tok], V_sel[:, :, :n_tok], K_win[:, :, lo:n_win], V_win[:, :, lo:n_win],
                         K_cmp[:, :, :S_cmp], V_cmp[:, :, :S_cmp], gate, l=l, d=d, l_sel=ls, n_sel=n, w=w)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    rg = torch.full((B, G, n, 2), -7, dtype=torch.int32, device="cuda")
    cu = lambda x: x.cuda().to(dtype)
    got = ops.decode_core(cu(q[:, None]), cu(K_sel), cu(V_sel), cu(K_win), cu(V_win), cu(K_cmp), cu(V_cmp),
                          tuple(x.cuda() for x in gate), cfg, t=t, S_sel_kv=n_tok, S_win_kv=n_win, win_off=win_off, S_cmp=S_cmp,
                          ranges_out=rg)
    assert torch.isfinite(got.float()).all()
    rg = rg.cpu()
    same = torch.tensor([[O.nonempty_ranges(rg[b, g].tolist()) == O.nonempty_ranges(want["ranges"][b, g].tolist())
                          for g in range(G)] for b in range(B)])
    assert (~same).sum() <= 1, f"{(~same).sum()} of {B * G} rows picked different blocks"
    err = (got[:, 0].float().cpu() - want["O"]).abs()[same]
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (err.max(), err.mean())
    # causality and clamping of what was selected
    assert (rg[..., 1] <= n_tok).all() and (rg[..., 0] >= 0).all()


@pytest.mark.parametrize("norm", ["full_row", "causal"])
def test_score_tc_four_mtile_path(norm):
    """Enough rows (B*G*S >= 4*21*148) that the scorer runs 4 M-tiles per CTA with single-buffered accumulators."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 2, 4096 + 77, 2, 6, 32, 16, 64, 16, 512
    gen = torch.Generator().manual_seed(3)
    Q = torch.randn(B, S, G, h, 64, generator=gen).bfloat16().float()
    Kc = torch.randn(B, G, O.num_cmp_blocks(S, l, d), 64, generator=gen).bfloat16().float()
    nm = ops.NORM_CAUSAL if norm == "causal" else ops.NORM_FULL_ROW
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC, norm_mode=nm)
    want = O.prefill_scores(Q, Kc, l, d, ls, n, w, norm)
    got = ops.score_pgrp(Q.cuda().bfloat16(), Kc.cuda().bfloat16(), cfg).cpu()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max() <= 2e-5 * h, (got - want).abs().max()
    r = ops.score_select(Q.cuda().bfloat16(), Kc.cuda().bfloat16(), cfg, mode=0).cpu()
    _assert_only_near_ties(r, O.select_ranges_prefill(want, ls, n, S), want, got, ls, n)


@pytest.mark.parametrize("norm", ["full_row", "causal"])
@pytest.mark.parametrize("S", [1500, 4096 + 77])
def test_score_tc_logits_growing_along_the_sequence(norm, S):
    """Compressed keys that grow 40x / 400x along the sequence: later logits sit hundreds of log2 units above the first tiles', the
    online maximum of pass 1 moves many times and nearly all probability mass is on a few keys.  Both scorer shapes (2 and 4
    M-tiles per CTA) stay finite and match the fp32 oracle; the tolerance is the fp32 spacing of logits of that size (s*c ~ 500:
    one ulp of the exponent argument is 3e-5, i.e. ~2e-5 relative on probabilities up to 2)."""
    ops = _ops()
    B, G, h, l, d, ls, n, w = 2, 2, 6, 32, 16, 64, 16, 512
    gen = torch.Generator().manual_seed(S)
    Q = torch.randn(B, S, G, h, 64, generator=gen).bfloat16().float()
    Kc = torch.randn(B, G, O.num_cmp_blocks(S, l, d), 64, generator=gen)
    Kc[:, :, 40:] *= 40.0
    Kc[:, :, 70:] *= 10.0
    Kc = Kc.bfloat16().float()
    nm = ops.NORM_CAUSAL if norm == "causal" else ops.NORM_FULL_ROW
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC, norm_mode=nm)
    want = O.prefill_scores(Q, Kc, l, d, ls, n, w, norm)
    got = ops.score_pgrp(Q.cuda().bfloat16(), Kc.cuda().bfloat16(), cfg).cpu()
    assert torch.isfinite(got).all()
    assert (got - want).abs().max() <= 1e-4 * h, (got - want).abs().max()


def test_dense_tc_reference_jump():
    """Later key tiles hold logits far above the first tile's maximum: the lazy softmax reference must move (rows redo the
    tile and rescale the accumulated O in TMEM) and still match the oracle."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 1, 1100, 2, 6, 32, 16, 64, 16, 512
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=21, dtype=torch.bfloat16)
    # window keys: tokens >= 400 are 30x larger; compressed keys: entries >= 20 are 30x larger
    ts[3][:, :, 400:] *= 30.0
    ts[5][:, :, 20:] *= 30.0
    ts = [t.bfloat16().float() for t in ts]
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    Q = ts[0].cuda().bfloat16()
    o, lse = ops.branch_attention(ops.BR_WIN, Q, ts[3].cuda().bfloat16(), ts[4].cuda().bfloat16(), cfg, return_lse=True)
    want, lse_w = O.win_attention(ts[0], ts[3], ts[4], w)
    assert torch.isfinite(o.float()).all()
    assert (o.float().cpu() - want).abs().max() <= 3e-2, (o.float().cpu() - want).abs().max()
    assert (lse.cpu() - lse_w).abs().max() <= 0.25  # logits are O(1e3) here; bf16-rounded inputs, fp32 accumulate
    o, lse = ops.branch_attention(ops.BR_CMP, Q, ts[5].cuda().bfloat16(), ts[6].cuda().bfloat16(), cfg, return_lse=True)
    want, lse_w = O.cmp_attention(ts[0], ts[5], ts[6], l, d)
    assert torch.isfinite(o.float()).all()
    assert (o.float().cpu() - want).abs().max() <= 3e-2, (o.float().cpu() - want).abs().max()


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("sel_mode,S,h,B", [(0, 700, 6, 2), (1, 700, 6, 1), (0, 1500, 4, 2), (1, 333, 8, 1), (0, 2100, 6, 1),
                                             (0, 200, 8, 3), (1, 130, 1, 2)])
def test_sel_blockmajor_vs_gather_vs_oracle(dtype, sel_mode, S, h, B):
    """KV-block-major selected branch (index build -> tcgen05 per-block attention -> merge of partials) against the fp32
    oracle and the query-major gather kernel.  Tolerance: max-abs 2e-2 / MAE 1e-3 (partials are stored in 16 bits)."""
    ops = _ops()
    G, l, d, ls, n, w = 2, 32, 16, 64, 16, 512
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=3 * S + h, dtype=dtype)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    Q, K, V = (t.cuda().to(dtype) for t in ts[:3])
    ranges = ops.score_select(Q, ts[5].cuda().to(dtype), cfg, mode=sel_mode)
    o_bm, lse_bm = ops.sel_attention_blockmajor(Q, K, V, cfg, ranges, return_lse=True)
    cfg_auto = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)  # gather kernel where it exists (h <= 8), else SIMT
    o_g, lse_g = ops.branch_attention(ops.BR_SEL, Q, K, V, cfg_auto, ranges, return_lse=True)
    want, lse_w = O.sel_attention(ts[0], ts[1], ts[2], ranges.cpu())
    assert torch.isfinite(o_bm.float()).all()
    err = (o_bm.float().cpu() - want).abs()
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (err.max(), err.mean())
    assert (o_bm.float() - o_g.float()).abs().max() <= 2e-2
    fin = torch.isfinite(lse_w)
    assert torch.equal(torch.isfinite(lse_bm.cpu()), fin)
    assert (lse_bm.cpu()[fin] - lse_w[fin]).abs().max() <= 2e-2
    empty = ~fin.any(dim=-1)
    if empty.any():
        assert torch.all(o_bm.float().cpu()[empty] == 0)


def test_prefill_core_long_uses_blockmajor_sel():
    """B*S*G >= 16384 and S >= 4096: nsa_prefill_fwd routes the selected branch through the block-major kernels."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 2, 4096, 2, 6, 32, 16, 64, 16, 512
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=13, dtype=torch.bfloat16)
    gen = torch.Generator().manual_seed(5)
    gate = (torch.randn(32, 64, generator=gen) * 0.3, torch.randn(32, generator=gen) * 0.1, torch.randn(3, 32, generator=gen) * 0.5,
            torch.zeros(3))
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
    dev = [t.cuda().bfloat16() for t in ts]
    Oc, ranges, gates = ops.prefill_core(*dev, tuple(x.cuda() for x in gate), cfg, sel_mode=0)
    # the selected branch alone, both kernels, same ranges
    o_bm = ops.sel_attention_blockmajor(dev[0], dev[1], dev[2], cfg, ranges)
    o_g = ops.branch_attention(ops.BR_SEL, dev[0], dev[1], dev[2], cfg, ranges)
    assert (o_bm.float() - o_g.float()).abs().max() <= 2e-2
    want = O.prefill_core(*ts, gate, l=l, d=d, l_sel=ls, n_sel=n, w=w, ranges=ranges.cpu())
    err = (Oc.float().cpu() - want["O"]).abs()
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (err.max(), err.mean())


@pytest.mark.parametrize("S,h,B", [(300, 6, 2), (2100, 6, 1), (5000, 4, 1), (1100, 8, 2), (64, 1, 1)])
def test_score_select_causal_pass2_is_bit_exact(S, h, B):
    """nsa_score_select stops the scorer's second pass at each CTA's causal limit (blocks a row may never select are not
    scored); the ranges must equal those selected from the complete p_grp -- bit-exact, both rules, full rows and chunks."""
    ops = _ops()
    G, l, d, ls, n, w = 2, 32, 16, 64, 16, 512
    ts = _case(B, S, G, h, l, d, ls, n, w, seed=S + 3 * h, dtype=torch.bfloat16)
    Q, Kc = ts[0].cuda().bfloat16(), ts[5].cuda().bfloat16()
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    pg = ops.score_pgrp(Q, Kc, cfg)
    fused0 = ops.score_select(Q, Kc, cfg, mode=0)
    assert torch.equal(fused0, ops.select_ranges_prefill(pg, ls, n, S))
    fused1 = ops.score_select(Q, Kc, cfg, mode=1)
    for t in sorted({0, 1, ls - 1, ls, S // 3, S - 2, S - 1} & set(range(S))):
        assert torch.equal(fused1[:, t], ops.select_ranges_decode(pg[:, t].contiguous(), ls, n, t)), t
    if S >= 2000:  # chunk of rows (the sequence so far ends with the chunk; still more than n_sel blocks) against the same cache
        t0, Sc = S // 2 + 5, S // 4
        part = ops.score_select(Q[:, t0:t0 + Sc].contiguous(), Kc, cfg, mode=0, t0=t0)
        assert torch.equal(part, fused0[:, t0:t0 + Sc])


def test_needle_at_64k_full_size_is_selected_and_retrieved():
    """BASELINE config 4 at full size (S=65536, m7c head dims, bf16): the needle check of bench/needle_64k_smoke.py:39-66 run through
    the whole path instead of with a hand-made range -- a needle (32 identical key rows = one compressed window, with a distinctive
    value) is planted at S/2, every later query matches it; scoring + selection must pick its selection block for every later row
    and the selected-branch output must be the needle value (cosine > 0.999).  Rows before
    the needle must never reach it (causality).  Size-independent property, no oracle call."""
    ops = _ops()
    B, S, G, h, D, l, d, ls, n, w = 1, 65536, 2, 6, 64, 32, 16, 64, 16, 512
    pos = S // 2
    gen = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=gen, device="cuda")
    K, V, Q = r(B, G, S, D), r(B, G, S, D), r(B, S, G, h, D)
    k_star, v_star = r(B, G, D), r(B, G, D)
    K[:, :, pos:pos + l] = k_star[:, :, None]          # one whole compressed window [pos, pos+l) (pos is a multiple of d)
    V[:, :, pos:pos + l] = v_star[:, :, None]
    Q[:, pos + l:] = 4.0 * k_star[:, None, :, None, :]   # every later query is aligned with the needle
    # phi without RoPE for this synthetic check: K_cmp[i] = mean of rows [i*d, i*d + l)
    K_cmp = K.unfold(2, l, d).mean(dim=-1).contiguous()
    assert K_cmp.shape[2] == (S - l) // d + 1
    Qb, Kb, Vb, Kcb = Q.bfloat16(), K.bfloat16(), V.bfloat16(), K_cmp.bfloat16()
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w)
    ranges = ops.score_select(Qb, Kcb, cfg, mode=0)                  # [1,S,G,K,2]
    s_, e_ = ranges[..., 0].long(), ranges[..., 1].long()
    t = torch.arange(S, device="cuda").view(1, S, 1, 1)
    covers = ((s_ <= pos) & (e_ >= pos + l) & (e_ > s_)).any(dim=-1)  # some range holds the whole needle window
    later = slice(pos + ls, S)                                        # rows whose block pos//ls is complete
    assert bool(covers[:, later].all()), f"{int((~covers[:, later]).sum())} later rows missed the needle block"
    assert not bool(((e_ > pos) & (t < pos)).any()), "a row before the needle reached it"
    rows = torch.tensor([pos + ls, pos + 4096, 50000, S - 1], device="cuda")
    out = ops.sel_attention_blockmajor(Qb, Kb, Vb, cfg, ranges)      # the kernel nsa_prefill_fwd uses at this length
    got = out[0, rows].float()                                       # [4,G,h,D]
    cos = torch.nn.functional.cosine_similarity(got, v_star[0][None, :, None, :].expand_as(got), dim=-1)
    assert float(cos.min()) > 0.999, float(cos.min())
