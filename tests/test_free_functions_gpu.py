"""The reference's stage-by-stage free functions and named wrappers (compute_pcmp_all, map_pcmp_to_pslc(_batched),
select_topn_ranges(_batched) with any force_init / force_local, convert_indices_to_ranges_batched(_v2), attention_bgh,
grouped_selection_attention*, sliding_window_attention, batched_causal_attention_compressed,
selection_attention_backward_reference) against vectors the REAL reference produced (tests/golden/free_functions.npz, written
by tests/golden/make_golden.py) and against the oracle.  Integer outputs bit-exact / range-equivalent, fp32 <= 5e-5."""
import pytest
import torch

from conftest import T, load_golden
from oracle import nsa_oracle as O
from test_oracle_golden import decode_case_is_well_defined

pytestmark = pytest.mark.gpu


def _meta(S, l, d, ls, n=8, w=64):
    from nsa_vibe_b200.core.block_index import build_block_meta
    return build_block_meta(S, l, d, ls, n, w)


def test_select_with_force_flags_matches_reference():
    from nsa_vibe_b200.core import selection_scorer as ss
    g = load_golden("free_functions")
    checked = 0
    for i in range(int(g["dec_n"])):
        ls, ns, t, fi, fl = [int(v) for v in g[f"dec_c{i}"]]
        p = T(g[f"dec_p{i}"])
        if not decode_case_is_well_defined(p.shape[-1], ls, ns, t, fi, fl):
            continue
        meta = _meta(p.shape[-1] * ls, ls // 2, ls // 4, ls, ns, 512)
        mine = ss.select_topn_ranges(p.cuda(), meta, ns, t, force_init=bool(fi), force_local=fl).cpu()
        ok, bad = O.ranges_equivalent(mine, T(g[f"dec_r{i}"]))
        assert ok, f"decode case {i} (l_sel={ls}, n={ns}, t={t}, force=({fi},{fl})): {bad} rows differ"
        checked += 1
    assert checked > 100
    for i in range(int(g["pre_n"])):
        ls, ns, S, fi, fl = [int(v) for v in g[f"pre_c{i}"]]
        p = T(g[f"pre_p{i}"])
        meta = _meta(S, ls // 2, ls // 4, ls, ns, 512)
        mine = ss.select_topn_ranges_batched(p.cuda(), meta, ns, S, force_init=bool(fi), force_local=fl).cpu()
        ref = T(g[f"pre_r{i}"])
        assert mine.shape == ref.shape, f"prefill case {i} force=({fi},{fl}): K {mine.shape} vs {ref.shape}"
        assert torch.equal(mine, ref), f"prefill case {i} (l_sel={ls}, n={ns}, S={S}, force=({fi},{fl}))"


def test_tie_break_is_dtype_independent_and_prefers_lower_index():
    """nsa/tests/test_selection_tiebreak.py:17-58: equal scores everywhere, no forced blocks."""
    from nsa_vibe_b200.core import selection_scorer as ss
    meta = _meta(64, 4, 2, 4, 8, 8)
    S_sel = meta.sel_starts.numel()
    outs = [ss.select_topn_ranges(torch.ones(1, 1, S_sel, dtype=dt).cuda(), meta, 3, 63, force_init=False, force_local=0).cpu()
            for dt in (torch.float32, torch.float16, torch.bfloat16)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert O.nonempty_ranges(outs[0][0, 0].tolist()) == [(0, 12)]  # blocks 0, 1, 2: the lower indices win the tie
    outs = [ss.select_topn_ranges_batched(torch.ones(1, 3, 1, S_sel, dtype=dt).cuda(), meta, 3, 3, force_init=False, force_local=0).cpu()
            for dt in (torch.float32, torch.float16, torch.bfloat16)]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert ss.validate_selection_determinism(torch.randn(1, 1, S_sel).cuda(), meta, n_top=5, t_token=63, num_trials=3) is True


def test_pcmp_map_and_group_reduce_match_reference():
    from nsa_vibe_b200.core import selection_scorer as ss
    g = load_golden("free_functions")
    p = ss.compute_pcmp_all(T(g["pcmp_Q"]).cuda(), T(g["pcmp_K"]).cuda(), 0.25).cpu()
    assert torch.allclose(p, T(g["pcmp_p"]), atol=2e-6), (p - T(g["pcmp_p"])).abs().max()
    with pytest.raises(RuntimeError):
        ss.compute_pcmp_all(T(g["pcmp_Q"]).cuda(), T(g["pcmp_K"]).cuda(), 0.5)  # only 1/sqrt(Dk) is implemented
    for k in range(int(g["map_n"])):
        S, l, d, ls = [int(v) for v in g[f"map_c{k}"]]
        meta = _meta(S, l, d, ls)
        got = ss.map_pcmp_to_pslc_batched(T(g[f"map_p{k}"]).cuda(), meta).cpu()
        assert got.shape == g[f"map_o{k}"].shape and torch.allclose(got, T(g[f"map_o{k}"]), atol=1e-6)
        got = ss.map_pcmp_to_pslc(T(g[f"map_ps{k}"]).cuda(), meta).cpu()
        assert torch.allclose(got, T(g[f"map_os{k}"]), atol=1e-6)
        # Eq.10 + the fused scorer's identity: group_reduce(map(p)) == h * map of one head when heads agree (test_group_consistency.py)
        base = torch.rand(2, 3, 1, T(g[f"map_ps{k}"]).shape[-1]).cuda()
        rep = ss.group_reduce_pslc(ss.map_pcmp_to_pslc(base.repeat(1, 1, 4, 1), meta))
        assert torch.allclose(rep, ss.map_pcmp_to_pslc(base, meta).squeeze(2) * 4, atol=1e-5)


def test_scoring_stages_compose_to_the_fused_scorer():
    """compute_pcmp_all -> map_pcmp_to_pslc_batched -> sum over heads == compute_pgrp_all (the fused kernel), fp32."""
    from nsa_vibe_b200.core import selection_scorer as ss
    S, l, d, ls = 300, 32, 16, 64
    meta = _meta(S, l, d, ls, 16, 512)
    Q = torch.randn(2, S, 2, 4, 32).cuda()
    Kc = torch.randn(2, 2, meta.cmp_starts.numel(), 32).cuda()
    staged = ss.map_pcmp_to_pslc_batched(ss.compute_pcmp_all(Q, Kc, 32 ** -0.5), meta).sum(dim=3)
    fused = ss.compute_pgrp_all(Q, Kc, meta)
    assert torch.allclose(staged, fused, atol=5e-6), (staged - fused).abs().max()


def test_indices_to_ranges_match_reference():
    from nsa_vibe_b200.core import selection_scorer as ss
    g = load_golden("free_functions")
    for k in range(int(g["i2r_n"])):
        ls, S_sel = [int(v) for v in g[f"i2r_c{k}"]]
        idx = T(g[f"i2r_i{k}"])
        meta = _meta(S_sel * ls, ls // 2, ls // 4, ls)
        v2 = ss.convert_indices_to_ranges_batched_v2(idx.cuda(), meta, idx.shape[1]).cpu()
        v1 = ss.convert_indices_to_ranges_batched(idx.cuda(), meta, idx.shape[1]).cpu()
        assert v2.shape == g[f"i2r_v2_{k}"].shape and v1.shape == g[f"i2r_v1_{k}"].shape
        assert torch.equal(v1, T(g[f"i2r_v1_{k}"]).to(torch.int32))      # the loop version: bit-exact
        ok, bad = O.ranges_equivalent(v2, T(g[f"i2r_v2_{k}"]))           # v2: the reference's own equivalence criterion
        assert ok, bad
        tpos = torch.arange(idx.shape[1]).view(1, -1, 1, 1)
        assert bool((v2[..., 1] <= tpos + 1).all())                      # causality (test_selection_v2_equiv.py:114-129)
    empty = ss.convert_indices_to_ranges_batched_v2(torch.zeros(1, 4, 1, 0, dtype=torch.int32).cuda(), _meta(32, 4, 2, 4), 4)
    assert empty.shape == (1, 4, 1, 0, 2)


def test_named_attention_wrappers_match_reference_and_oracle():
    from nsa_vibe_b200 import kernels as K
    from nsa_vibe_b200.core import attention_kernels as ak
    g = load_golden("free_functions")
    o = K.attention_bgh(T(g["bgh_q"]).cuda(), T(g["bgh_K"]).cuda(), T(g["bgh_V"]).cuda(), causal=False).cpu()
    assert torch.allclose(o, T(g["bgh_O"]), atol=5e-5), (o - T(g["bgh_O"])).abs().max()
    a = load_golden("attention")
    l, d, ls, n, w = [int(v) for v in a["cfg"]]
    Q, Kk, V = T(a["Q"]).cuda(), T(a["K"]).cuda(), T(a["V"]).cuda()
    for fn in (ak.grouped_selection_attention, ak.grouped_selection_attention_masked, ak.selection_attention_varlen_all,
               K.selection_attention_cuda, K.selection_attention_triton):
        assert torch.allclose(fn(Q, Kk, V, T(a["ranges"]).cuda()).cpu(), T(a["O_sel"]), atol=5e-5)
    assert torch.allclose(ak.sliding_window_attention(Q, Kk, V, w).cpu(), T(a["O_win"]), atol=5e-5)
    assert torch.allclose(ak.batched_causal_attention_compressed(Q, T(a["K_cmp"]).cuda(), T(a["V_cmp"]).cuda(), l, d).cpu(),
                          T(a["O_cmp"]), atol=5e-5)
    assert ak.sliding_window_attention(Q, Kk[:, :, :0], V[:, :, :0], w).abs().sum() == 0
    # analytical backward under the intended semantics == autograd through the oracle's masked attention
    dO = torch.randn_like(T(a["O_sel"]))
    dQ, dK, dV = K.selection_attention_backward_reference(Q, Kk, V, T(a["ranges"]).cuda(), dO.cuda())
    cpu = [T(a[k]).clone().requires_grad_(True) for k in ("Q", "K", "V")]
    (O.sel_attention(*cpu, T(a["ranges"]))[0] * dO).sum().backward()
    for mine, ref in zip((dQ, dK, dV), cpu):
        assert float((mine.cpu() - ref.grad).norm() / ref.grad.norm()) <= 5e-3
