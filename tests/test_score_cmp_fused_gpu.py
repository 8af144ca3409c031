"""Pass 2 of the scorer fused with the compressed branch (tc_score_cmp.cu, used by nsa_prefill_full_fwd for long 16-bit prefill):
  * the ranges are BIT-EQUAL to the stand-alone scorer + selection (same exponentials, same Eq.9 / Eq.10 summation order);
  * O_cmp agrees with the stand-alone dense kernel and with the oracle (max-abs <= 2e-2, MAE <= 1e-3 for 16-bit inputs);
  * the saved lse / branch outputs drive the same backward (gradients vs the separately computed path, rel-err <= 3e-2);
  * rows whose causal logits sit far below the full-row reference (future compressed keys dominate) take their own reference."""
import pytest
import torch

from oracle import nsa_oracle as O

pytestmark = pytest.mark.gpu
L, D, LS, N, W = 32, 16, 64, 16, 512


def _case(B, S, G, h, seed, dtype):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen).to(dtype).float()
    S_cmp = O.num_cmp_blocks(S, L, D)
    ts = [r(B, S, G, h, 64), r(B, G, S, 64), r(B, G, S, 64), r(B, G, S, 64), r(B, G, S, 64), r(B, G, S_cmp, 64), r(B, G, S_cmp, 64)]
    gate = (torch.randn(32, 64, generator=gen) * 0.3, torch.randn(32, generator=gen) * 0.1, torch.randn(3, 32, generator=gen) * 0.5,
            torch.zeros(3))
    return ts, gate


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("norm", ["full_row", "causal"])
@pytest.mark.parametrize("B,S,h,sel_mode", [(2, 3200, 6, 0), (1, 6300, 6, 1), (5, 1100, 8, 0), (6, 1700, 4, 0)])
def test_fused_pass2_cmp_equals_separate_kernels(dtype, norm, B, S, h, sel_mode):
    from nsa_vibe_b200 import ops
    G = 2
    assert B * G * S >= 4 * (128 // h) * 148, "the fused kernel serves launches that fill the machine"
    ts, gate = _case(B, S, G, h, seed=S + h, dtype=dtype)
    nm = ops.NORM_CAUSAL if norm == "causal" else ops.NORM_FULL_ROW
    dev = [t.cuda().to(dtype) for t in ts]
    gd = tuple(x.cuda() for x in gate)
    # ranges: bit-equal to the stand-alone scorer + selection
    cfg = ops.NSAConfig(l=L, d=D, l_sel=LS, n_sel=N, w=W, norm_mode=nm)
    with torch.no_grad():
        O_f, r_f, g_f = ops.prefill_core(*dev, gd, cfg, sel_mode=sel_mode)
        r_s = ops.score_select(dev[0], dev[5], cfg, mode=sel_mode)
        assert torch.equal(r_f, r_s)
        O_s, _, g_s = ops.prefill_core(*dev, gd, cfg, sel_mode=sel_mode, ranges=r_s, ranges_trusted=True)
    assert torch.equal(g_f, g_s)
    dd = (O_f.float() - O_s.float()).abs()  # two 16-bit results: one rounding step of an O(1) value apart at most, equal on average
    assert dd.max() <= 4e-2 and dd.mean() <= 5e-4, (float(dd.max()), float(dd.mean()))
    # the compressed branch alone (forced gate): fused vs stand-alone dense kernel vs oracle
    cfg_c = ops.NSAConfig(l=L, d=D, l_sel=LS, n_sel=N, w=W, norm_mode=nm, gate_mode=ops.GATE_CMP)
    with torch.no_grad():
        Oc_f = ops.prefill_core(*dev, None, cfg_c, sel_mode=sel_mode)[0]
        Oc_d = ops.branch_attention(ops.BR_CMP, dev[0], dev[5], dev[6], cfg)
    want, lse_w = O.cmp_attention(ts[0], ts[5], ts[6], L, D)
    for got in (Oc_f, Oc_d):
        err = (got.float().cpu() - want).abs()
        assert torch.isfinite(got.float()).all()
        assert err.max() <= 2e-2 and err.mean() <= 1e-3, (float(err.max()), float(err.mean()))
    assert torch.all(Oc_f[:, :L - 1] == 0)  # rows before the first compressed token attend nothing


def test_fused_pass2_cmp_saves_what_backward_needs():
    from nsa_vibe_b200 import ops
    B, S, G, h, dtype = 2, 3200, 2, 6, torch.bfloat16
    ts, gate = _case(B, S, G, h, seed=77, dtype=dtype)
    cfg = ops.NSAConfig(l=L, d=D, l_sel=LS, n_sel=N, w=W)
    gen = torch.Generator().manual_seed(5)
    dO = torch.randn(B, S, G, h, 64, generator=gen).to(dtype).cuda()

    def run(with_ranges):
        leaves = [t.cuda().to(dtype).requires_grad_(True) for t in ts]
        gl = tuple(x.cuda().requires_grad_(True) for x in gate)
        rg = None
        if with_ranges:
            with torch.no_grad():
                rg = ops.score_select(leaves[0], leaves[5], cfg, mode=0)
        Oc, ranges, _ = ops.prefill_core(*leaves, gl, cfg, sel_mode=0, ranges=rg, ranges_trusted=True)
        grads = torch.autograd.grad(Oc, leaves + list(gl), dO)
        return Oc, ranges, grads
    O_f, r_f, g_f = run(False)   # nsa_prefill_full_fwd: lse / O_cmp saved by the fused kernel
    O_s, r_s, g_s = run(True)    # nsa_prefill_fwd: the dense compressed kernel
    assert torch.equal(r_f, r_s)
    assert (O_f.float() - O_s.float()).abs().max() <= 4e-2
    for a, b in zip(g_f, g_s):
        assert torch.isfinite(a.float()).all()
        assert float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)) <= 3e-2


def test_rows_dominated_by_future_keys_take_their_own_reference():
    """Full-row normaliser: compressed keys that lie in a row's future enter its p_cmp normaliser.  When they dwarf
    every causal logit, exp2(s*c - offs) underflows for all causal keys; the branch then uses the causal maximum as reference."""
    from nsa_vibe_b200 import ops
    B, S, G, h, dtype = 1, 6300, 2, 6, torch.bfloat16
    ts, gate = _case(B, S, G, h, seed=3, dtype=dtype)
    ts[5][:, :, 60:] *= 60.0   # compressed keys 60.. are huge: rows t < ~1000 only see small ones causally
    ts[5] = ts[5].to(dtype).float()  # what the kernels read
    cfg = ops.NSAConfig(l=L, d=D, l_sel=LS, n_sel=N, w=W)
    cfg_c = ops.NSAConfig(l=L, d=D, l_sel=LS, n_sel=N, w=W, gate_mode=ops.GATE_CMP)
    dev = [t.cuda().to(dtype) for t in ts]
    with torch.no_grad():
        got = ops.prefill_core(*dev, None, cfg_c, sel_mode=0)[0]
        st = ops.score_stats(dev[0], dev[5], cfg)
        dense = ops.branch_attention(ops.BR_CMP, dev[0], dev[5], dev[6], cfg)
    want, _ = O.cmp_attention(ts[0], ts[5], ts[6], L, D)
    assert torch.isfinite(got.float()).all()
    # rows whose causal keys are all small (num_cmp(t) <= 60) but whose full-row reference is set by the huge future keys
    hi = torch.tensor([O.num_cmp_at(t, L, D, ts[5].shape[2]) for t in range(S)])
    at_stake = (hi > 0) & (hi <= 60)
    gap = (st[..., 0] - st[..., 1]).cpu()[0][at_stake]
    assert at_stake.sum() > 800 and (gap > 100).float().mean() > 0.9, "the test data must exercise the own-reference path"
    err = (got.float().cpu() - want).abs()[0][at_stake]
    assert err.max() <= 2e-2 and err.mean() <= 1e-3, (float(err.max()), float(err.mean()))
    assert got[0][at_stake].float().abs().mean() > 0.01   # they did not collapse to zero
    # every other row: the same function as the stand-alone dense kernel (with logits this peaked a 16-bit P is far from the fp32
    # oracle for both, so the two kernels are compared with each other)
    dd = (got.float() - dense.float()).abs()
    assert dd.max() <= 4e-2 and dd.mean() <= 5e-4, (float(dd.max()), float(dd.mean()))
