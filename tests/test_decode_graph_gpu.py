"""The decode step as ONE replayed CUDA graph driven by a device-side step record (ops.DecodeGraphStep: nsa_decode_produce /
nsa_decode_emit / nsa_decode_fwd_stepped / nsa_decode_advance) against the eager step of the same module (NSA_DECODE_GRAPH=0):
same outputs, same caches, same emission schedule, same read counters, token after token (bench/bench_decode.py:123-136 loop)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _module(env, dtype):
    from nsa_vibe_b200.core.nsa_attention import NSAAttention
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        torch.manual_seed(3)
        m = NSAAttention(dim=256, n_heads=12, n_kv_groups=2, d_k=64, d_v=64, l=32, d=16, l_sel=64, n_sel=16, w=128).cuda().to(dtype)
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    return m


def _kv(B, dtype):
    from nsa_vibe_b200.cache.kv_cache import create_empty_kv
    from nsa_vibe_b200.core.block_index import build_block_meta
    return create_empty_kv(B, 2, 64, 64, build_block_meta(64, 32, 16, 64, 16, 128), device="cuda", dtype=dtype)


def _same_step(og, oe, kv_eager, i):
    """Bit-equal whenever both paths run the fused tcgen05 decode kernel.  Before the first compressed token exists the eager step
    runs the SIMT kernels (nsa_decode_fwd needs S_cmp >= 1 for the fused kernel; the stepped entry point always uses it): same
    function, different summation order."""
    if kv_eager.K_cmp.shape[2] >= 1:
        assert torch.equal(og, oe), f"step {i}: max diff {(og.float() - oe.float()).abs().max()}"
    else:
        assert (og.float() - oe.float()).abs().max() <= 2e-2, i


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("S0,steps", [(0, 70), (200, 90), (37, 40)])
def test_graph_decode_equals_eager_decode(dtype, S0, steps):
    from nsa_vibe_b200 import ops
    B = 3
    mg, me = _module({"NSA_DECODE_GRAPH": "1"}, dtype), _module({"NSA_DECODE_GRAPH": "0"}, dtype)
    me.load_state_dict(mg.state_dict())
    assert mg._env_cache["decode_graph"] and not me._env_cache["decode_graph"]
    g = torch.Generator(device="cuda").manual_seed(S0 + steps)
    xs = torch.randn(B, S0 + steps, 256, generator=g, device="cuda").to(dtype)
    kvg, kve = _kv(B, dtype), _kv(B, dtype)
    with torch.no_grad():
        if S0:
            og, kvg = mg(xs[:, :S0], kvg, prefill=True)
            oe, kve = me(xs[:, :S0], kve, prefill=True)
            assert torch.equal(og, oe)
        n0 = ops.launch_count
        for i in range(S0, S0 + steps):
            og, kvg = mg(xs[:, i:i + 1], kvg, prefill=False)
            oe, kve = me(xs[:, i:i + 1], kve, prefill=False)
            assert og.shape == (B, 1, 256)
            _same_step(og, oe, kve, i)
            assert torch.equal(mg._last_ranges, me._last_ranges), f"step {i}: selected ranges differ"
    assert getattr(kvg, "_decode_graph")[2] is not None, "the graph path was not taken"
    assert getattr(kve, "_decode_graph", None) is None
    for f in ("K_sel", "V_sel", "K_win", "V_win", "K_cmp_raw_seq", "V_cmp_raw_seq", "K_cmp", "V_cmp", "reads_pred", "reads_act_total",
              "reads_act_sel", "reads_act_cmp", "reads_act_win"):
        a, b = getattr(kvg, f), getattr(kve, f)
        assert a.shape == b.shape, (f, a.shape, b.shape)
        assert torch.equal(a, b), f
    n = S0 + steps
    assert kvg.K_sel.shape[2] == n and kvg.K_win.shape[2] == min(128, n) and kvg.K_cmp.shape[2] == (0 if n < 32 else (n - 32) // 16 + 1)
    assert kvg.reads_pred.numel() == steps


def test_graph_decode_survives_cache_growth_and_weight_updates():
    """Slabs that fill up are reallocated (the graph is re-captured on the new pointers); a changed weight re-captures too."""
    dtype, B = torch.bfloat16, 2
    mg, me = _module({"NSA_DECODE_GRAPH": "1"}, dtype), _module({"NSA_DECODE_GRAPH": "0"}, dtype)
    me.load_state_dict(mg.state_dict())
    g = torch.Generator(device="cuda").manual_seed(1)
    xs = torch.randn(B, 700, 256, generator=g, device="cuda").to(dtype)
    kvg, kve = _kv(B, dtype), _kv(B, dtype)
    with torch.no_grad():
        for i in range(700):
            if i == 350:
                for m in (mg, me):
                    m.gate.fc2.bias.data.add_(torch.tensor([0.5, -0.25, 0.0], device="cuda", dtype=dtype))
                    m.W_Q.weight.data.mul_(1.01)
            og, kvg = mg(xs[:, i:i + 1], kvg, prefill=False)
            oe, kve = me(xs[:, i:i + 1], kve, prefill=False)
            _same_step(og, oe, kve, i)
    assert torch.equal(kvg.K_cmp, kve.K_cmp) and torch.equal(kvg.K_sel, kve.K_sel) and kvg.K_sel.shape[2] == 700


def test_fp32_module_keeps_the_eager_step():
    m = _module({"NSA_DECODE_GRAPH": "1"}, torch.float32)
    kv = _kv(1, torch.float32)
    with torch.no_grad():
        for i in range(3):
            o, kv = m(torch.randn(1, 1, 256, device="cuda"), kv, prefill=False)
    assert getattr(kv, "_decode_graph", None) is None and torch.isfinite(o).all()
