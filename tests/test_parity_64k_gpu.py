"""Full-size parity (BASELINE.json config 4: S = 65536, m7c head dims, bf16): the GPU step's p_grp, selection, per-branch
outputs and gated output against the oracle on the same inputs, at sampled rows {31, 64, 4095, 32768, 65535} (8 rows from
each position where they fit).  Tolerances as everywhere: p_grp max-abs <= 2e-5 * h; selection identical except at fp32
near-ties of the scores (each differing row is checked); bf16 outputs vs the fp32 oracle max-abs <= 2e-2, MAE <= 1e-3."""
import pytest
import torch

from oracle import nsa_oracle as O

pytestmark = pytest.mark.gpu

S, G, H, D, L, DD, LS, N, W = 65536, 2, 6, 64, 32, 16, 64, 16, 512
POSITIONS = (31, 64, 4095, 32768, 65535)
ROWS = 8


@pytest.fixture(scope="module")
def step():
    from nsa_vibe_b200 import ops
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(2024)
    r = lambda *s: torch.randn(*s, generator=g, device=dev).bfloat16()
    S_cmp = (S - L) // DD + 1
    t = dict(Q=r(1, S, G, H, D), K_sel=r(1, G, S, D), V_sel=r(1, G, S, D), K_win=r(1, G, S, D), V_win=r(1, G, S, D),
             K_cmp=r(1, G, S_cmp, D), V_cmp=r(1, G, S_cmp, D))
    gate = (torch.randn(32, D, generator=g, device=dev) * 0.3, torch.randn(32, generator=g, device=dev) * 0.1,
            torch.randn(3, 32, generator=g, device=dev) * 0.5, torch.zeros(3, device=dev))
    cfg = ops.NSAConfig(l=L, d=DD, l_sel=LS, n_sel=N, w=W)
    with torch.no_grad():
        p_grp = ops.score_pgrp(t["Q"], t["K_cmp"], cfg)
        ranges = ops.score_select(t["Q"], t["K_cmp"], cfg, mode=0)
        Oc, ranges2, gates = ops.prefill_core(t["Q"], t["K_sel"], t["V_sel"], t["K_win"], t["V_win"], t["K_cmp"], t["V_cmp"], gate, cfg,
                                              sel_mode=0)
        assert torch.equal(ranges, ranges2)
        branches = dict(O_cmp=ops.branch_attention(ops.BR_CMP, t["Q"], t["K_cmp"], t["V_cmp"], cfg),
                        O_win=ops.branch_attention(ops.BR_WIN, t["Q"], t["K_win"], t["V_win"], cfg),
                        O_sel=ops.sel_attention_blockmajor(t["Q"], t["K_sel"], t["V_sel"], cfg, ranges, ranges_trusted=True))
    cpu = {k: v.float().cpu() for k, v in t.items()}
    return dict(cpu=cpu, gate=tuple(x.float().cpu() for x in gate), p_grp=p_grp, ranges=ranges, O=Oc, gates=gates, **branches)


@pytest.mark.parametrize("pos", POSITIONS)
def test_step_at_64k_matches_oracle_on_sampled_rows(step, pos):
    t0 = min(pos, S - ROWS) if pos + ROWS > S else pos
    c = step["cpu"]
    sl = slice(t0, t0 + ROWS)
    kw = dict(l=L, d=DD, l_sel=LS, n_sel=N, w=W, t0=t0, S_total=S)
    # scores and selection
    want_p = O.prefill_scores(c["Q"][:, sl], c["K_cmp"], L, DD, LS, N, W, "full_row", t0, S)
    got_p = step["p_grp"][:, sl].cpu()
    assert (got_p - want_p).abs().max() <= 2e-5 * H, (got_p - want_p).abs().max()
    want_r = O.select_ranges_prefill(want_p, LS, N, S, t0)
    got_r = step["ranges"][:, sl].cpu()
    same = torch.ones(ROWS, G, dtype=torch.bool)
    for s in range(ROWS):
        for g in range(G):
            if O.nonempty_ranges(got_r[0, s, g].tolist()) != O.nonempty_ranges(want_r[0, s, g].tolist()):
                same[s, g] = False
                t = t0 + s
                nv = min((t + 1) // LS, want_p.shape[-1])
                delta = float((got_p[0, s, g, :nv] - want_p[0, s, g, :nv]).abs().max()) if nv else 0.0
                assert O.selection_difference_is_near_tie(want_p[0, s, g], got_r[0, s, g].tolist(), want_r[0, s, g].tolist(), LS, N, t,
                                                          delta), f"row t={t}, g={g}: different blocks without a score tie"
    assert (~same).sum() <= 1
    assert bool((got_r[..., 1] <= torch.arange(t0, t0 + ROWS).view(1, -1, 1, 1) + 1).all())
    # branches and gated output, given the GPU's ranges
    want = O.prefill_core(c["Q"][:, sl], c["K_sel"], c["V_sel"], c["K_win"], c["V_win"], c["K_cmp"], c["V_cmp"], step["gate"],
                          ranges=got_r, **kw)
    assert torch.allclose(step["gates"][:, sl].cpu(), want["gates"], atol=1e-5)
    for k in ("O_cmp", "O_win", "O_sel", "O"):
        err = (step[k][:, sl].float().cpu() - want[k]).abs()
        assert torch.isfinite(step[k][:, sl].float()).all()
        assert err.max() <= 2e-2 and err.mean() <= 1e-3, (k, pos, float(err.max()), float(err.mean()))
