"""Pins oracle/nsa_oracle.py against vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest
import torch

from conftest import T, load_golden
from oracle import nsa_oracle as O


def test_meta_eq9_weights():
    g = load_golden("meta")
    for i in range(int(g["n"])):
        S, l, d, ls = [int(v) for v in g[f"cfg{i}"]]
        m = O.build_meta(S, l, d, ls, 16, 512)
        assert [m.S_cmp, m.S_sel] == g[f"ncmp{i}"].tolist()
        assert np.array_equal(m.coo_rows, g[f"rows{i}"])
        assert np.array_equal(m.coo_cols, g[f"cols{i}"])
        assert np.array_equal(m.coo_vals, g[f"vals{i}"])  # bit-exact fp32 weights


def test_meta_divisibility_guard():
    with pytest.raises(ValueError):
        O.build_meta(128, 32, 12, 64, 16, 512)


def test_rope_and_phi():
    g = load_golden("rope_phi")
    x, pos = T(g["x"]), T(g["pos"])
    assert torch.allclose(O.rope(x, pos), T(g["rope"]), atol=1e-6)
    assert torch.allclose(O.rope(x, pos, scale=8.0), T(g["rope_scale8"]), atol=1e-6)
    l, d = [int(v) for v in g["ld"]]
    Kc, Vc = O.phi_avg_pool(T(g["K_raw"]), T(g["V_raw"]), l, d)
    assert torch.allclose(Kc, T(g["K_cmp"]), atol=1e-6)
    assert torch.allclose(Vc, T(g["V_cmp"]), atol=1e-6)


def test_scores_pcmp_pslc_pgrp():
    g = load_golden("scores")
    S, l, d, ls = [int(v) for v in g["cfg"]]
    Q, Kc = T(g["Q"]), T(g["K_cmp"])
    p = O.pcmp_all(Q, Kc, 1.0 / math.sqrt(Q.shape[-1]))
    assert torch.allclose(p, T(g["p_cmp"]), atol=1e-6)
    meta = O.build_meta(S, l, d, ls, 16, 512)
    pslc = O.pslc_from_pcmp(T(g["p_cmp"]), meta)
    assert torch.allclose(pslc, T(g["p_slc"]), atol=1e-7)
    assert torch.allclose(O.pgrp_from_pslc(pslc), T(g["p_grp"]), atol=1e-6)
    # fewer compressed rows than the map covers (early decode)
    S2, l2, d2, ls2 = [int(v) for v in g["cfg2"]]
    meta2 = O.build_meta(S2, l2, d2, ls2, 8, 64)
    p2 = O.pcmp_all(T(g["Q2"]), T(g["K_cmp2"]), 1.0 / math.sqrt(8))
    pslc2 = O.pslc_from_pcmp(p2, meta2)
    assert torch.allclose(pslc2, T(g["p_slc2"]), atol=1e-6)


def test_select_decode_matches_reference():
    g = load_golden("select")
    n = int(g["dec_n"])
    assert n > 50
    for i in range(n):
        ls, ns, t = [int(v) for v in g[f"dec_c{i}"]]
        mine = O.select_ranges_decode(T(g[f"dec_p{i}"]), ls, ns, t)
        ok, bad = O.ranges_equivalent(mine, T(g[f"dec_r{i}"]))
        assert ok, f"decode case {i} (l_sel={ls}, n={ns}, t={t}): {bad} rows differ"
        assert int(mine[..., 1].max()) <= t + 1


def test_select_prefill_matches_reference_bit_exact():
    g = load_golden("select")
    n = int(g["pre_n"])
    for i in range(n):
        ls, ns, S = [int(v) for v in g[f"pre_c{i}"]]
        mine = O.select_ranges_prefill(T(g[f"pre_p{i}"]), ls, ns, S)
        ref = T(g[f"pre_r{i}"])
        assert mine.shape == ref.shape, f"prefill case {i}: K {mine.shape} vs {ref.shape}"
        assert torch.equal(mine, ref), f"prefill case {i} (l_sel={ls}, n={ns}, S={S})"


def test_branch_attention():
    g = load_golden("attention")
    l, d, ls, n, w = [int(v) for v in g["cfg"]]
    Q, K, V = T(g["Q"]), T(g["K"]), T(g["V"])
    Osel, _ = O.sel_attention(Q, K, V, T(g["ranges"]))
    assert torch.allclose(Osel, T(g["O_sel"]), atol=2e-6)
    Owin, _ = O.win_attention(Q, K, V, w)
    assert torch.allclose(Owin, T(g["O_win"]), atol=2e-6)
    Ocmp, _ = O.cmp_attention(Q, T(g["K_cmp"]), T(g["V_cmp"]), l, d)
    assert torch.allclose(Ocmp, T(g["O_cmp"]), atol=2e-6)


def test_sel_attention_empty_rows_are_zero():
    Q = torch.randn(1, 3, 1, 2, 8)
    K = torch.randn(1, 1, 10, 8)
    V = torch.randn(1, 1, 10, 8)
    r = torch.zeros(1, 3, 1, 2, 2, dtype=torch.int32)
    r[0, 1, 0, 0] = torch.tensor([2, 5])
    o, lse = O.sel_attention(Q, K, V, r)
    assert torch.all(o[0, 0] == 0) and torch.all(o[0, 2] == 0) and torch.isfinite(o).all()
    assert torch.isinf(lse[0, 0]).all()


def test_gate_mlp():
    g = load_golden("gate")
    q = T(g["q"])
    args = (T(g["fc1_w"]), T(g["fc1_b"]), T(g["fc2_w"]))
    assert torch.allclose(O.gate_mlp(q, *args, T(g["fc2_b_soft"])), T(g["p"]), atol=1e-6)
    assert torch.allclose(O.gate_mlp(q, *args, T(g["fc2_b_soft"]), tau=0.5), T(g["p_tau"]), atol=1e-6)
    assert torch.equal(O.gate_mlp(q, *args, T(g["fc2_b_hard"])), T(g["p_hard"]))
    assert torch.allclose(O.gate_mlp(q, *args, T(g["fc2_b_soft"]), mode=O.GATE_UNIFORM), torch.full((50, 3), 1 / 3))


def _module_inputs(g):
    dim, H, G, dk, dv, l, d, ls, n, w = [int(v) for v in g["cfg"]]
    sd = {k[4:]: T(v) for k, v in g.items() if k.startswith("sd__")}
    return (dim, H, G, dk, dv, l, d, ls, n, w), sd


def _project(x, sd, cfg, t0=0):
    dim, H, G, dk, dv, l, d, ls, n, w = cfg
    B, S, _ = x.shape
    pos = torch.arange(t0, t0 + S)
    # the reference rotates Q as ONE vector of width n_heads*d_k (nsa_attention.py:1002-1009), K per head-dim
    Q = O.rope(x @ sd["W_Q.weight"].T, pos).reshape(B, S, G, H // G, dk)
    kv = lambda name: (x @ sd[name].T).view(B, S, G, -1).permute(0, 2, 1, 3)
    K_sel, V_sel = O.rope(kv("W_K_sel.weight"), pos), kv("W_V_sel.weight")
    K_win, V_win = O.rope(kv("W_K_win.weight"), pos), kv("W_V_win.weight")
    return Q, K_sel, V_sel, K_win, V_win, kv("W_K_cmp.weight"), kv("W_V_cmp.weight")


def test_whole_module_prefill_intended_semantics():
    """Oracle hot path + plain projections reproduces the reference module (batched prefill,
    masked selection, true-softmax cmp swapped in)."""
    g = load_golden("module")
    cfg, sd = _module_inputs(g)
    dim, H, G, dk, dv, l, d, ls, n, w = cfg
    x = T(g["x"])
    Q, K_sel, V_sel, K_win, V_win, Kc_raw, Vc_raw = _project(x, sd, cfg)
    K_cmp, V_cmp = O.phi_avg_pool(Kc_raw, Vc_raw, l, d)
    assert torch.allclose(K_cmp, T(g["kv_K_cmp"]), atol=1e-5)
    assert torch.allclose(K_sel, T(g["kv_K_sel"]), atol=1e-5)
    assert torch.allclose(K_win[:, :, -w:], T(g["kv_K_win"]), atol=1e-5)
    gp = (sd["gate.fc1.weight"], sd["gate.fc1.bias"], sd["gate.fc2.weight"], sd["gate.fc2.bias"])
    r = O.prefill_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gp, l=l, d=d, l_sel=ls, n_sel=n, w=w)
    B, S = x.shape[:2]
    out = r["O"].reshape(B, S, H * dv) @ sd["out.weight"].T
    assert torch.allclose(out, T(g["out_intended"]), atol=2e-5), (out - T(g["out_intended"])).abs().max()
    for mode, name in ((O.GATE_CMP, "cmp"), (O.GATE_SEL, "sel"), (O.GATE_WIN, "win")):
        r = O.prefill_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gp, l=l, d=d, l_sel=ls, n_sel=n, w=w, gate_mode=mode)
        out = r["O"].reshape(B, S, H * dv) @ sd["out.weight"].T
        assert torch.allclose(out, T(g["out_force_" + name]), atol=2e-5), name


def test_whole_module_decode_intended_semantics():
    g = load_golden("module")
    cfg, sd = _module_inputs(g)
    dim, H, G, dk, dv, l, d, ls, n, w = cfg
    xs = T(g["dec_x"])
    S0, Tn = [int(v) for v in g["dec_S0T"]]
    gp = (sd["gate.fc1.weight"], sd["gate.fc1.bias"], sd["gate.fc2.weight"], sd["gate.fc2.bias"])
    Q, K_sel, V_sel, K_win, V_win, Kc_raw, Vc_raw = _project(xs, sd, cfg)
    B = xs.shape[0]
    outs = []
    for t in range(S0 + Tn):
        n_raw = t + 1
        ncmp = O.num_cmp_blocks(n_raw, l, d)
        # decode emission == pooling of the prefix (nsa/tests/test_decode_step.py:226-278)
        K_cmp, V_cmp = O.phi_avg_pool(Kc_raw[:, :, :n_raw], Vc_raw[:, :, :n_raw], l, d)
        assert K_cmp.shape[2] == ncmp
        assert O.decode_emits(n_raw, l, d) == (ncmp > O.num_cmp_blocks(n_raw - 1, l, d))
        lo = max(0, n_raw - w)
        r = O.decode_core(Q[:, t], K_sel[:, :, :n_raw], V_sel[:, :, :n_raw], K_win[:, :, lo:n_raw], V_win[:, :, lo:n_raw],
                          K_cmp, V_cmp, gp, l=l, d=d, l_sel=ls, n_sel=n, w=w)
        outs.append(r["O"].reshape(B, 1, H * dv) @ sd["out.weight"].T)
    out = torch.cat(outs, dim=1)
    assert torch.allclose(out, T(g["dec_out_steps"]), atol=3e-5), (out - T(g["dec_out_steps"])).abs().max()
    assert torch.allclose(K_cmp, T(g["dec_K_cmp_final"]), atol=1e-5)
    reads = [O.expected_reads(t + 1, l, d, n, ls, w)[0] for t in range(S0 + Tn)]
    assert reads == g["dec_reads_total"].tolist()


def test_oracle_autograd_matches_reference_grads():
    """Gradients of the oracle hot path (torch autograd through the restatement) against the
    reference module's own backward on the same inputs."""
    g = load_golden("module")
    cfg, sd = _module_inputs(g)
    dim, H, G, dk, dv, l, d, ls, n, w = cfg
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x = T(g["x"]).clone().requires_grad_(True)
    Q, K_sel, V_sel, K_win, V_win, Kc_raw, Vc_raw = _project(x, sd, cfg)
    K_cmp, V_cmp = O.phi_avg_pool(Kc_raw, Vc_raw, l, d)
    gp = (sd["gate.fc1.weight"], sd["gate.fc1.bias"], sd["gate.fc2.weight"], sd["gate.fc2.bias"])
    r = O.prefill_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gp, l=l, d=d, l_sel=ls, n_sel=n, w=w)
    B, S = x.shape[:2]
    out = r["O"].reshape(B, S, H * dv) @ sd["out.weight"].T
    (out * T(g["grad_out"])).sum().backward()
    assert torch.allclose(x.grad, T(g["grad_x"]), atol=1e-4), (x.grad - T(g["grad_x"])).abs().max()
    for k, v in sd.items():
        ref = T(g["grad__" + k])
        assert torch.allclose(v.grad, ref, atol=2e-4, rtol=1e-3), (k, (v.grad - ref).abs().max())


def decode_case_is_well_defined(S_sel, l_sel, n_sel, t, fi, fl):
    """select_topn_ranges keeps picks that landed on -inf (future / incomplete) blocks; which -inf entry torch.topk returns is
    unspecified, and without a forced local block such a pick can merge into a visible [cb*l_sel, t+1) range.  The rule is
    defined when the current block is forced anyway (force_local >= 1) or enough finite candidates exist."""
    if fl >= 1:
        return True
    nvalid = min((t + 1) // l_sel, S_sel)
    forced_valid = 1 if (fi and nvalid > 0) else 0
    return nvalid - forced_valid >= min(max(n_sel - fi - fl, 0), S_sel)


def test_free_function_goldens_select_with_force_flags():
    """select_topn_ranges(_batched) with every (force_init, force_local) the reference's callers and tests use."""
    g = load_golden("free_functions")
    checked = 0
    for i in range(int(g["dec_n"])):
        ls, ns, t, fi, fl = [int(v) for v in g[f"dec_c{i}"]]
        if not decode_case_is_well_defined(g[f"dec_p{i}"].shape[-1], ls, ns, t, fi, fl):
            continue
        mine = O.select_ranges_decode(T(g[f"dec_p{i}"]), ls, ns, t, bool(fi), fl)
        ok, bad = O.ranges_equivalent(mine, T(g[f"dec_r{i}"]))
        assert ok, f"decode case {i} (l_sel={ls}, n={ns}, t={t}, force=({fi},{fl})): {bad} rows differ"
        checked += 1
    assert checked > 100
    for i in range(int(g["pre_n"])):
        ls, ns, S, fi, fl = [int(v) for v in g[f"pre_c{i}"]]
        mine = O.select_ranges_prefill(T(g[f"pre_p{i}"]), ls, ns, S, 0, bool(fi), fl)
        ref = T(g[f"pre_r{i}"])
        assert mine.shape == ref.shape, f"prefill case {i} force=({fi},{fl}): K {mine.shape} vs {ref.shape}"
        assert torch.equal(mine, ref), f"prefill case {i} (l_sel={ls}, n={ns}, S={S}, force=({fi},{fl}))"


def test_free_function_goldens_pcmp_map_and_ranges():
    g = load_golden("free_functions")
    p = O.pcmp_all(T(g["pcmp_Q"]), T(g["pcmp_K"]), 0.25)
    assert torch.allclose(p, T(g["pcmp_p"]), atol=1e-6)
    for k in range(int(g["map_n"])):
        S, l, d, ls = [int(v) for v in g[f"map_c{k}"]]
        meta = O.build_meta(S, l, d, ls, 8, 64)
        assert torch.allclose(O.pslc_from_pcmp(T(g[f"map_p{k}"]), meta), T(g[f"map_o{k}"]), atol=1e-6)
        short = T(g[f"map_ps{k}"])
        assert torch.allclose(O.pslc_from_pcmp(short[:, None], meta)[:, 0], T(g[f"map_os{k}"]), atol=1e-6)
    for k in range(int(g["i2r_n"])):
        ls, S_sel = [int(v) for v in g[f"i2r_c{k}"]]
        mine = O.indices_to_ranges(T(g[f"i2r_i{k}"]), S_sel, ls)
        # v2 leaves clamped-away runs as rows with end <= start; the criterion of the reference's own equivalence test
        # (test_selection_v2_equiv.py:80-111) compares the non-empty ranges in order
        for ref in (T(g[f"i2r_v2_{k}"]), T(g[f"i2r_v1_{k}"])):
            ok, bad = O.ranges_equivalent(mine, ref)
            assert ok, bad
    q, K, V = T(g["bgh_q"]), T(g["bgh_K"]), T(g["bgh_V"])
    o, _ = O._masked_attention(q[:, None], K, V, torch.ones(q.shape[0], 1, q.shape[1], K.shape[2], dtype=torch.bool))
    assert torch.allclose(o[:, 0], T(g["bgh_O"]), atol=2e-6)


def test_phi_mlp_depthwise_conv():
    """phi="mlp" against the reference's own _phi_apply_seq / _phi_apply_last (its constructor bug repaired in make_golden.py)."""
    g = load_golden("phi_mlp")
    dim, H, G, dk, dv, l, d, ls, n, w = [int(v) for v in g["cfg"]]
    wk, wv = T(g["sd__phi_k_conv.weight"]), T(g["sd__phi_v_conv.weight"])
    Kc, Vc = O.phi_conv(T(g["K_raw"]), T(g["V_raw"]), wk, wv, l, d)
    assert torch.allclose(Kc, T(g["K_cmp"]), atol=2e-6) and torch.allclose(Vc, T(g["V_cmp"]), atol=2e-6)
    Kl, Vl = O.phi_conv(T(g["K_raw"])[:, :, 20:28], T(g["V_raw"])[:, :, 20:28], wk, wv, l, d, pos=torch.arange(20, 28))
    assert torch.allclose(Kl, T(g["K_last"]), atol=2e-6) and torch.allclose(Vl, T(g["V_last"]), atol=2e-6)
    # initialised to 1/l it is the average pool
    Ka, Va = O.phi_avg_pool(T(g["K_raw"]), T(g["V_raw"]), l, d)
    Kc1, Vc1 = O.phi_conv(T(g["K_raw"]), T(g["V_raw"]), torch.full_like(wk, 1.0 / l), torch.full_like(wv, 1.0 / l), l, d)
    assert torch.allclose(Kc1, Ka, atol=1e-6) and torch.allclose(Vc1, Va, atol=1e-6)
