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
