"""RoPE + re-layout and phi average-pool kernels (producers of the hot path's inputs) against the torch restatements of
nsa/core/rope.py:16-51 and nsa/core/compress_pool.py:9-38 (nsa_vibe_b200/core/rope.py, compress_pool.py), forward and backward.
Tolerance: fp32 max-abs 2e-6 (same operation order, libdevice sin/cos/pow), 16-bit one rounding step (2^-8 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mods():
    from nsa_vibe_b200 import ops
    from nsa_vibe_b200.core.compress_pool import avg_pool_phi_rope_kv
    from nsa_vibe_b200.core.rope import apply_rope
    return ops, apply_rope, avg_pool_phi_rope_kv


def _tol(dtype):
    return dict(atol=2e-6, rtol=1e-6) if dtype == torch.float32 else dict(atol=2e-2, rtol=2e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,S,V,D,t0,scale", [(2, 37, 2, 64, 0, 1.0), (1, 300, 12, 64, 0, 1.0), (1, 5, 2, 16, 1234, 2.0), (1, 1, 2, 64, 4095, 1.0)])
def test_rope_shape_matches_torch_chain(dtype, B, S, V, D, t0, scale):
    ops, apply_rope, _ = _mods()
    g = torch.Generator(device="cuda").manual_seed(S + V)
    x = torch.randn(B, S, V * D, generator=g, device="cuda").to(dtype)
    pos = torch.arange(t0, t0 + S, device="cuda")
    for rope, cache in (("vector", True), ("token", False), ("none", True)):
        xa = x.clone().requires_grad_(True)
        xb = x.clone().requires_grad_(True)
        got = ops.rope_shape(xa, V, D, rope=rope, to_cache_layout=cache, t0=t0, scale=scale)
        if rope == "vector":    # reference order for K: reshape to [B,G,S,D], then rotate each D-vector
            want = apply_rope(xb.view(B, S, V, D).permute(0, 2, 1, 3).contiguous(), pos, scale=scale)
        elif rope == "token":   # reference order for Q: rotate the token's V*D values as one vector, then view
            want = apply_rope(xb, pos, scale=scale).view(B, S, V, D)
        else:
            want = xb.view(B, S, V, D).permute(0, 2, 1, 3).contiguous()
        assert got.shape == want.shape
        assert torch.allclose(got.float(), want.float(), **_tol(dtype)), (rope, (got.float() - want.float()).abs().max())
        dy = torch.randn(want.shape, generator=g, device="cuda").to(dtype)
        got.backward(dy)
        want.backward(dy)
        assert torch.allclose(xa.grad.float(), xb.grad.float(), **_tol(dtype)), (rope, "grad")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,G,S,D,l,d,t0", [(2, 2, 200, 64, 32, 16, 0), (1, 2, 33, 16, 16, 8, 0), (1, 1, 32, 64, 32, 16, 992), (1, 2, 20, 64, 32, 16, 0)])
def test_phi_avgpool_matches_torch_chain(dtype, B, G, S, D, l, d, t0):
    ops, _, avg_pool_phi_rope_kv = _mods()
    g = torch.Generator(device="cuda").manual_seed(S + l)
    K = torch.randn(B, G, S, D, generator=g, device="cuda").to(dtype)
    V = torch.randn(B, G, S, D, generator=g, device="cuda").to(dtype)
    ka, va, kb, vb = (t.clone().requires_grad_(True) for t in (K, V, K, V))
    Kc, Vc = ops.phi_avgpool(ka, va, l, d, t0=t0)
    Kw, Vw = avg_pool_phi_rope_kv(kb, vb, l, d, pos=torch.arange(t0, t0 + S, device="cuda"))
    assert Kc.shape == Kw.shape and Vc.shape == Vw.shape
    if Kc.numel() == 0:
        return
    assert torch.allclose(Kc.float(), Kw.float(), **_tol(dtype)), (Kc.float() - Kw.float()).abs().max()
    assert torch.allclose(Vc.float(), Vw.float(), **_tol(dtype))
    dk = torch.randn(Kw.shape, generator=g, device="cuda").to(dtype)
    dv = torch.randn(Vw.shape, generator=g, device="cuda").to(dtype)
    (Kc.float() * dk.float()).sum().backward()
    (Vc.float() * dv.float()).sum().backward()
    (Kw.float() * dk.float()).sum().backward()
    (Vw.float() * dv.float()).sum().backward()
    assert torch.allclose(ka.grad.float(), kb.grad.float(), **_tol(dtype)), (ka.grad.float() - kb.grad.float()).abs().max()
    assert torch.allclose(va.grad.float(), vb.grad.float(), **_tol(dtype))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,H,G,Dk,Dv,t,scale", [(3, 12, 2, 64, 64, 0, 1.0), (2, 4, 2, 16, 32, 777, 1.0), (5, 12, 2, 64, 64, 4095, 2.0)])
def test_decode_produce_matches_the_seven_chains(dtype, B, H, G, Dk, Dv, t, scale):
    """One launch = rope(Q as one vector), rope(K_sel), rope(K_win) per Dk-vector, six rows scattered into their slabs
    (nsa_attention.py:545-586), and the step's read counters (kv_cache.py:51-65)."""
    ops, _, _ = _mods()
    g = torch.Generator(device="cuda").manual_seed(B + t)
    widths = [H * Dk] + [G * Dk, G * Dv] * 3
    y = torch.randn(B, sum(widths), generator=g, device="cuda").to(dtype)
    parts = torch.split(y, widths, dim=-1)
    caps = [9, 11, 7, 8, 13, 10]
    rows = [4, 10, 0, 3, 12, 5]
    slabs = [torch.full((B, G, caps[i], Dv if i & 1 else Dk), 7.0, device="cuda", dtype=dtype) for i in range(6)]
    q = torch.empty(B, H * Dk, device="cuda", dtype=dtype)
    ctr = torch.zeros(5, 16, dtype=torch.int64, device="cuda")
    ops.decode_produce(y, q, slabs, rows, H=H, G=G, Dk=Dk, Dv=Dv, t=t, scale=scale, counters=ctr, counters_idx=3,
                       counter_vals=(11, 12, 13, 14, 15))
    want_q = ops.rope_shape(parts[0].contiguous().view(B, 1, -1), H, Dk, rope="token", t0=t, scale=scale).reshape(B, H * Dk)
    assert torch.equal(q, want_q)
    for i in range(6):
        D = Dv if i & 1 else Dk
        rope = "vector" if i in (0, 2) else "none"
        want = ops.rope_shape(parts[1 + i].contiguous().view(B, 1, -1), G, D, rope=rope, to_cache_layout=True, t0=t, scale=scale)
        assert torch.equal(slabs[i][:, :, rows[i]], want[:, :, 0]), i
        mask = torch.ones(caps[i], dtype=torch.bool, device="cuda")
        mask[rows[i]] = False
        assert bool((slabs[i][:, :, mask] == 7.0).all()), f"slab {i}: rows other than {rows[i]} were touched"
    assert ctr[:, 3].tolist() == [11, 12, 13, 14, 15] and int(ctr.sum()) == 65


def test_kv_inplace_append_keeps_reference_views():
    """NSA_KV.token_append_slots / commit_token_append / counter_slot keep the reference's public tensors (kv_cache.py:8-65)."""
    from nsa_vibe_b200.cache.kv_cache import create_empty_kv
    from nsa_vibe_b200.core.block_index import build_block_meta
    kv = create_empty_kv(2, 2, 16, 16, build_block_meta(64, 8, 4, 16, 2, 4), device=torch.device("cuda"), dtype=torch.float32)
    like = torch.empty(1, device="cuda", dtype=torch.bfloat16)
    for step in range(70):
        slabs, rows = kv.token_append_slots(like)
        assert rows == [step] * 6
        for s in slabs:
            s[:, :, step] = step
        c, idx = kv.counter_slot()
        c[:, idx] = step
        kv.commit_token_append(4)
        kv.commit_counters()
    assert kv.K_sel.shape == (2, 2, 70, 16) and kv.K_sel.dtype == torch.bfloat16 and kv.K_win.shape == (2, 2, 4, 16)
    assert kv.K_win[0, 0, :, 0].tolist() == [66, 67, 68, 69] and kv.K_sel[1, 1, :, 3].tolist() == list(range(70))
    assert kv.reads_pred.tolist() == list(range(70)) and kv.reads_act_win.shape == (70,)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,S,H,G,Dk,Dv,t0,scale", [(2, 37, 12, 2, 64, 64, 0, 1.0), (1, 300, 4, 2, 16, 32, 50, 2.0), (3, 1, 12, 2, 64, 64, 9, 1.0)])
def test_project_split_matches_the_seven_chains_forward_and_backward(dtype, B, S, H, G, Dk, Dv, t0, scale):
    """ops.project_split (one launch per direction) == rope_shape per tensor on column slices of the fused projection output,
    values and gradients bit for bit (same arithmetic, nsa_attention.py:998-1016)."""
    ops, _, _ = _mods()
    g = torch.Generator(device="cuda").manual_seed(S + H)
    widths = [H * Dk] + [G * Dk, G * Dv] * 3
    y = torch.randn(B, S, sum(widths), generator=g, device="cuda").to(dtype)
    ya = y.clone().requires_grad_(True)
    yb = y.clone().requires_grad_(True)
    got = ops.project_split(ya, H=H, G=G, Dk=Dk, Dv=Dv, t0=t0, scale=scale)
    parts = torch.split(yb, widths, dim=-1)
    want = [ops.rope_shape(parts[0].contiguous(), H, Dk, rope="token", t0=t0, scale=scale).view(B, S, G, H // G, Dk)]
    for i in range(6):
        D = Dv if i & 1 else Dk
        want.append(ops.rope_shape(parts[1 + i].contiguous(), G, D, rope="vector" if i in (0, 2) else "none", to_cache_layout=True,
                                   t0=t0, scale=scale))
    loss_a = loss_b = 0.0
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)
        dy = torch.randn(a.shape, generator=g, device="cuda").to(dtype)
        loss_a = loss_a + (a.float() * dy.float()).sum()
        loss_b = loss_b + (b.float() * dy.float()).sum()
    loss_a.backward()
    loss_b.backward()
    assert torch.equal(ya.grad, yb.grad)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("t0,scale", [(0, 1.0), (1000, 2.0)])
def test_project_split_with_rotation_tables_is_bit_identical(dtype, t0, scale, monkeypatch):
    """Rows of 256 tokens and more take sin / cos from a cached table (nsa_rope_table) instead of sincosf per element: the table is
    the producers' own arithmetic, so values and gradients equal the per-tensor rope_shape chains bit for bit, and the table itself
    equals the torch expression of rope.py:14-33 rounded to the dtype."""
    ops, _, _ = _mods()
    B, S, H, G, Dk, Dv = 2, 700, 12, 2, 64, 64
    assert S >= ops._ROPE_TABLE_MIN_ROWS
    ops._rope_table_cache.clear()
    g = torch.Generator(device="cuda").manual_seed(5)
    widths = [H * Dk] + [G * Dk, G * Dv] * 3
    y = torch.randn(B, S, sum(widths), generator=g, device="cuda").to(dtype)
    ya = y.clone().requires_grad_(True)
    yb = y.clone().requires_grad_(True)
    got = ops.project_split(ya, H=H, G=G, Dk=Dk, Dv=Dv, t0=t0, scale=scale)
    assert len(ops._rope_table_cache) == 1, "the table path was not taken"
    parts = torch.split(yb, widths, dim=-1)
    want = [ops.rope_shape(parts[0].contiguous(), H, Dk, rope="token", t0=t0, scale=scale).view(B, S, G, H // G, Dk)]
    for i in range(6):
        D = Dv if i & 1 else Dk
        want.append(ops.rope_shape(parts[1 + i].contiguous(), G, D, rope="vector" if i in (0, 2) else "none", to_cache_layout=True,
                                   t0=t0, scale=scale))
    loss_a = loss_b = 0.0
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)
        dy = torch.randn(a.shape, generator=g, device="cuda").to(dtype)
        loss_a = loss_a + (a.float() * dy.float()).sum()
        loss_b = loss_b + (b.float() * dy.float()).sum()
    loss_a.backward()
    loss_b.backward()
    assert torch.equal(ya.grad, yb.grad)
    # the table against the reference's expression (fp32 angles; the device sin / cos may differ from torch's in the last bit)
    tq, tk = ops.rope_tables(y.device, dtype, H * Dk, Dk, t0, S, 10000.0, scale)
    pos = torch.arange(t0, t0 + S, device="cuda", dtype=torch.float32) / scale
    for tab, dim in ((tq, H * Dk), (tk, Dk)):
        inv = 10000.0 ** (-2.0 * torch.arange(dim // 2, device="cuda", dtype=torch.float32) / dim)
        ang = pos[:, None] * inv[None, :]
        tol = 2e-6 if dtype == torch.float32 else (8e-3 if dtype == torch.bfloat16 else 1e-3)
        assert (tab[..., 0].float() - ang.sin()).abs().max() <= tol + 2e-3 * (t0 > 0)
        assert (tab[..., 1].float() - ang.cos()).abs().max() <= tol + 2e-3 * (t0 > 0)


def test_phi_mlp_conv_matches_reference_and_autograd():
    """Learnable phi (depthwise Conv1d, phi="mlp"): kernels vs the reference's own outputs (tests/golden/phi_mlp.npz), gradients
    of inputs and taps vs autograd through the oracle, and the module (prefill, eager decode and CUDA-graph decode emission)."""
    from conftest import T, load_golden
    from nsa_vibe_b200 import ops
    from oracle import nsa_oracle as O
    g = load_golden("phi_mlp")
    dim, H, G, dk, dv, l, d, ls, n, w = [int(v) for v in g["cfg"]]
    wk, wv = T(g["sd__phi_k_conv.weight"]), T(g["sd__phi_v_conv.weight"])
    K_raw, V_raw = T(g["K_raw"]), T(g["V_raw"])
    Kc, Vc = ops.phi_conv(K_raw.cuda(), V_raw.cuda(), wk.cuda(), wv.cuda(), l, d)
    assert torch.allclose(Kc.cpu(), T(g["K_cmp"]), atol=5e-6) and torch.allclose(Vc.cpu(), T(g["V_cmp"]), atol=5e-6)
    Kl, Vl = ops.phi_conv(K_raw[:, :, 20:28].cuda().contiguous(), V_raw[:, :, 20:28].cuda().contiguous(), wk.cuda(), wv.cuda(), l, d, t0=20)
    assert torch.allclose(Kl.cpu(), T(g["K_last"]), atol=5e-6) and torch.allclose(Vl.cpu(), T(g["V_last"]), atol=5e-6)
    # gradients: dx (rotation transposed), dw (reduction over batch and positions)
    dev = [t.clone().cuda().requires_grad_(True) for t in (K_raw, V_raw, wk, wv)]
    cpu = [t.clone().requires_grad_(True) for t in (K_raw, V_raw, wk, wv)]
    gk, gv = torch.randn_like(T(g["K_cmp"])), torch.randn_like(T(g["V_cmp"]))
    a, b = ops.phi_conv(*dev, l, d)
    ((a * gk.cuda()).sum() + (b * gv.cuda()).sum()).backward()
    a2, b2 = O.phi_conv(*cpu, l, d)
    ((a2 * gk).sum() + (b2 * gv).sum()).backward()
    for x, y in zip(dev, cpu):
        assert x.grad.shape == y.grad.shape
        assert float((x.grad.cpu() - y.grad).norm() / y.grad.norm()) <= 1e-5
    # module: state-dict compatible parameters, prefill + decode emission equals the reference's compressed stream
    from nsa_vibe_b200.cache.kv_cache import create_empty_kv
    from nsa_vibe_b200.core.block_index import build_block_meta
    from nsa_vibe_b200.core.nsa_attention import NSAAttention
    m = NSAAttention(dim=dim, n_heads=H, n_kv_groups=G, d_k=dk, d_v=dv, l=l, d=d, l_sel=ls, n_sel=n, w=w, phi="mlp").cuda()
    m.load_state_dict({k[4:]: T(v).cuda() for k, v in g.items() if k.startswith("sd__")})
    x = T(g["x"]).cuda()
    kv = create_empty_kv(1, G, dk, dv, build_block_meta(64, l, d, ls, n, w), device="cuda")
    with torch.no_grad():
        _, kv = m(x[:, :18], kv, prefill=True)
        assert torch.allclose(kv.K_cmp.cpu(), T(g["K_cmp_after_prefill"]), atol=5e-6)
        for i in range(18, 30):
            _, kv = m(x[:, i:i + 1], kv, prefill=False)
    assert torch.allclose(kv.K_cmp.cpu(), T(g["K_cmp_final"]), atol=5e-6) and torch.allclose(kv.V_cmp.cpu(), T(g["V_cmp_final"]), atol=5e-6)
    # bf16, m7c head dims: the CUDA-graph decode step emits through the same taps as the eager step
    torch.manual_seed(0)
    mg = NSAAttention(dim=256, n_heads=12, n_kv_groups=2, d_k=64, d_v=64, l=32, d=16, l_sel=64, n_sel=16, w=128, phi="mlp").cuda().bfloat16()
    with torch.no_grad():
        mg.phi_k_conv.weight.add_(torch.randn_like(mg.phi_k_conv.weight) * 0.02)
        mg.phi_v_conv.weight.add_(torch.randn_like(mg.phi_v_conv.weight) * 0.02)
    xs = torch.randn(2, 80, 256, device="cuda").bfloat16()
    kv = create_empty_kv(2, 2, 64, 64, build_block_meta(64, 32, 16, 64, 16, 128), device="cuda", dtype=torch.bfloat16)
    with torch.no_grad():
        for i in range(80):
            _, kv = mg(xs[:, i:i + 1], kv, prefill=False)
        assert getattr(kv, "_decode_graph")[2] is not None
        want_k, want_v = ops.phi_conv(kv.K_cmp_raw_seq, kv.V_cmp_raw_seq, mg.phi_k_conv.weight, mg.phi_v_conv.weight, 32, 16)
    assert kv.K_cmp.shape[2] == 4 and torch.equal(kv.K_cmp, want_k) and torch.equal(kv.V_cmp, want_v)
