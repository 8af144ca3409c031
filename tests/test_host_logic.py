"""Host-side mirror of the reference interface: block metadata, cache container, RoPE / phi producers and the
module's argument checking.  CPU only (no kernel is launched)."""
import numpy as np
import pytest
import torch

from conftest import T, load_golden


def test_block_meta_matches_reference():
    from nsa_vibe_b200 import build_block_meta
    g = load_golden("meta")
    for i in range(int(g["n"])):
        S, l, d, ls = [int(v) for v in g[f"cfg{i}"]]
        m = build_block_meta(S, l, d, ls, 16, 512)
        assert [m.cmp_starts.numel(), m.sel_starts.numel()] == g[f"ncmp{i}"].tolist()
        assert np.array_equal(m.M_csl_coo_indices[0].numpy(), g[f"rows{i}"])
        assert np.array_equal(m.M_csl_coo_indices[1].numpy(), g[f"cols{i}"])
        assert np.array_equal(m.M_csl_coo_values.numpy(), g[f"vals{i}"])
        assert m.M_csl_indptr.dtype == torch.int32 and m.M_csl_indptr[-1] == len(g[f"rows{i}"])
    with pytest.raises(ValueError):
        build_block_meta(128, 32, 12, 64, 16, 512)  # nsa/tests/test_block_math.py:43-47
    build_block_meta(66047, 32, 16, 64, 16, 512)  # 2.85 s in the reference; closed form here


def test_rope_and_phi_match_reference():
    from nsa_vibe_b200.core.compress_pool import avg_pool_phi_rope_kv
    from nsa_vibe_b200.core.rope import apply_rope
    g = load_golden("rope_phi")
    x, pos = T(g["x"]), T(g["pos"])
    assert torch.allclose(apply_rope(x, pos), T(g["rope"]), atol=1e-6)
    assert torch.allclose(apply_rope(x, pos, scale=8.0), T(g["rope_scale8"]), atol=1e-6)
    assert apply_rope(x.bfloat16(), pos).dtype == torch.bfloat16  # nsa/tests/test_rope_dtype.py
    l, d = [int(v) for v in g["ld"]]
    Kc, Vc = avg_pool_phi_rope_kv(T(g["K_raw"]), T(g["V_raw"]), l, d)
    assert torch.allclose(Kc, T(g["K_cmp"]), atol=1e-6) and torch.allclose(Vc, T(g["V_cmp"]), atol=1e-6)
    Kc0, _ = avg_pool_phi_rope_kv(T(g["K_raw"])[:, :, :5], T(g["V_raw"])[:, :, :5], l, d)
    assert Kc0.shape[2] == 0


def test_kv_cache_semantics():
    from nsa_vibe_b200 import build_block_meta, create_empty_kv
    B, G, D, w = 2, 2, 8, 5
    kv = create_empty_kv(B, G, D, D, build_block_meta(64, 16, 8, 32, 4, w), device="cpu")
    ref_k = torch.zeros(B, G, 0, D)
    for step in range(12):
        k = torch.randn(B, G, 1, D)
        kv.update_selection_raw(k, k * 2)
        kv.update_window(k, k * 2, w)
        kv.append_cmp_raw(k, k)
        ref_k = torch.cat([ref_k, k], dim=2)
        assert torch.equal(kv.K_sel, ref_k) and torch.equal(kv.V_sel, ref_k * 2)
        assert torch.equal(kv.K_win, ref_k[:, :, -w:]) and kv.K_win.shape[2] == min(w, step + 1)  # kv_cache.py:32-38
        assert kv.K_cmp_raw_seq.shape[2] == step + 1
    assert kv.length("K_win") == 12 and kv.slab("K_win").shape[2] >= 12
    ptr = kv.slab("K_sel").data_ptr()
    kv.reserve(100)
    assert kv.slab("K_sel").shape[2] >= 100 and torch.equal(kv.K_sel, ref_k)
    p2 = kv.slab("K_sel").data_ptr()
    kv.update_selection_raw(torch.randn(B, G, 3, D), torch.randn(B, G, 3, D))
    assert kv.slab("K_sel").data_ptr() == p2 and kv.K_sel.shape[2] == 15  # no reallocation after reserve
    # a field assigned directly (reference style) is adopted on the next append
    kv.K_sel = ref_k.clone()
    kv.V_sel = ref_k.clone()
    kv.update_selection_raw(torch.ones(B, G, 1, D), torch.ones(B, G, 1, D))
    assert kv.K_sel.shape[2] == 13 and torch.equal(kv.K_sel[:, :, :12], ref_k)
    kv.append_reads_pred(7)
    kv.append_reads_actual(7, 4, 1, 2)
    assert kv.reads_pred.tolist() == [7] and kv.reads_act_win.tolist() == [2]


def test_module_contract_on_cpu():
    from nsa_vibe_b200 import NSAAttention, build_block_meta, create_empty_kv
    with pytest.raises(ValueError):
        NSAAttention(64, 4, 2, 16, 16, l=32, d=12)
    with pytest.raises(ValueError):
        NSAAttention(64, 4, 2, 16, 16, phi="conv2d")
    m = NSAAttention(64, 4, 2, 16, 16, l=8, d=4, l_sel=16, phi="mlp")  # learnable phi: depthwise Conv1d initialised to the mean
    assert m.phi_k_conv.weight.shape == (16, 1, 8) and m.phi_v_conv is not None
    assert torch.all(m.phi_k_conv.weight == 1.0 / 8)
    m = NSAAttention(64, 4, 2, 16, 16, l=16, d=8, l_sel=32, n_sel=4, w=40)
    assert sorted(m.state_dict()) == sorted([
        "W_Q.weight", "W_K_sel.weight", "W_V_sel.weight", "W_K_win.weight", "W_V_win.weight", "W_K_cmp.weight",
        "W_V_cmp.weight", "out.weight", "gate.fc1.weight", "gate.fc1.bias", "gate.fc2.weight", "gate.fc2.bias"])
    assert m.h_per_group == 2 and m.gate.fc1.out_features == 8 and torch.all(m.gate.fc2.bias == 0)
    kv = create_empty_kv(1, 2, 16, 16, build_block_meta(64, 16, 8, 32, 4, 40), device="cpu")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 8, 64), kv, prefill=True)  # the product path never computes on the CPU
    assert m.get_gate_stats() is None and all(v == 0 for v in m.get_fallback_counters().values())


def test_gate_mlp_torch_forward_matches_reference():
    from nsa_vibe_b200 import GateMLP
    g = load_golden("gate")
    gm = GateMLP(16)
    with torch.no_grad():
        gm.fc1.weight.copy_(T(g["fc1_w"])); gm.fc1.bias.copy_(T(g["fc1_b"])); gm.fc2.weight.copy_(T(g["fc2_w"]))
        assert torch.allclose(gm(T(g["q"])), T(g["p"]), atol=1e-6)
        gm.fc2.bias.copy_(T(g["fc2_b_hard"]))
        assert torch.equal(gm(T(g["q"])), T(g["p_hard"]))
