"""Tensor-core backward (tc_bwd.cu) against autograd through the fp32 oracle and against the SIMT backward.

Tolerance: gradients of 16-bit inputs, relative error ||g - g_ref|| / ||g_ref|| <= 3e-2 against the fp32 oracle
(P and dS are rounded to 16 bits before the gradient contractions, like the forward's P), <= 2e-2 against the SIMT kernel fed
the same saved O / lse."""
import pytest
import torch

from oracle import nsa_oracle as O

pytestmark = pytest.mark.gpu

TOL_ORACLE, TOL_SIMT = 3e-2, 2e-2


def _ops():
    from nsa_vibe_b200 import ops
    return ops


def _rel(a, b):
    # relative to the reference norm; a reference that is exactly zero (w = 1: dS = P o (dP - D) = 0) is compared in
    # absolute terms instead (rms error <= tol * 1e-4)
    return float((a - b).norm() / b.norm().clamp_min(1e-4 * b.numel() ** 0.5))


def _case(B, S, G, h, l, d, seed, dtype, S_kv=None):
    gen = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=gen).to(dtype).float()
    S_kv = S if S_kv is None else S_kv
    S_cmp = O.num_cmp_blocks(S_kv, l, d)
    return [r(B, S, G, h, 64), r(B, G, S_kv, 64), r(B, G, S_kv, 64), r(B, G, S_kv, 64), r(B, G, S_kv, 64),
            r(B, G, S_cmp, 64), r(B, G, S_cmp, 64)]


def _grads_dev(ops, br, Q, K, V, cfg, ranges, dO, dtype, t0=0):
    q, k, v = (t.cuda().to(dtype).requires_grad_(True) for t in (Q, K, V))
    o = ops.branch_attention(br, q, k, v, cfg, ranges, t0=t0)
    (o.float() * dO.cuda()).sum().backward()
    return [t.grad.float().cpu() for t in (q, k, v)]


def _grads_oracle(br, Q, K, V, ranges, dO, l, d, w, t0=0):
    q, k, v = (t.clone().requires_grad_(True) for t in (Q, K, V))
    if br == 1:
        o, _ = O.sel_attention(q, k, v, ranges)
    elif br == 2:
        o, _ = O.win_attention(q, k, v, w, t0=t0)
    else:
        o, _ = O.cmp_attention(q, k, v, l, d, t0=t0)
    (o * dO).sum().backward()
    return [q.grad, k.grad, v.grad]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("S,h,w,B", [(700, 6, 512, 2), (1500, 4, 128, 1), (333, 8, 64, 2), (2100, 6, 300, 1), (260, 1, 512, 1),
                                     (513, 16, 1, 1)])
def test_bwd_tc_branches_vs_oracle_and_simt(dtype, S, h, w, B):
    ops = _ops()
    G, l, d, ls, n = 2, 32, 16, 64, 16
    ts = _case(B, S, G, h, l, d, seed=S + h + w, dtype=dtype)
    # the forward gather kernel serves h <= 8: beyond that AUTO = SIMT forward + tensor-core backward
    cfg_tc = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC if h <= 8 else ops.IMPL_AUTO)
    cfg_si = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_SIMT)
    dO = torch.randn(B, S, G, h, 64, generator=torch.Generator().manual_seed(7)).to(dtype).float()
    ranges = ops.score_select(ts[0].cuda().to(dtype), ts[5].cuda().to(dtype), cfg_si, mode=0)
    for br, K, V in ((ops.BR_CMP, ts[5], ts[6]), (ops.BR_SEL, ts[1], ts[2]), (ops.BR_WIN, ts[3], ts[4])):
        rg = ranges if br == ops.BR_SEL else None
        g_tc = _grads_dev(ops, br, ts[0], K, V, cfg_tc, rg, dO, dtype)
        g_si = _grads_dev(ops, br, ts[0], K, V, cfg_si, rg, dO, dtype)
        g_or = _grads_oracle(br, ts[0], K, V, None if rg is None else rg.cpu(), dO, l, d, w)
        for nme, a, b, c in zip(("dQ", "dK", "dV"), g_tc, g_si, g_or):
            assert torch.isfinite(a).all(), (br, nme)
            assert _rel(a, c) <= TOL_ORACLE, (br, nme, "oracle", _rel(a, c))
            assert _rel(a, b) <= TOL_SIMT, (br, nme, "simt", _rel(a, b))


def test_bwd_tc_chunked_rows_t0():
    """Query rows t0..t0+S-1 against longer caches (chunked prefill): same key ranges as the oracle's t0 variant."""
    ops = _ops()
    B, G, h, l, d, ls, n, w = 1, 2, 6, 32, 16, 64, 16, 256
    S_full, t0, S = 1400, 777, 300
    dtype = torch.bfloat16
    ts = _case(B, S, G, h, l, d, seed=11, dtype=dtype, S_kv=S_full)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC)
    dO = torch.randn(B, S, G, h, 64, generator=torch.Generator().manual_seed(3)).to(dtype).float()
    for br, K, V in ((ops.BR_CMP, ts[5], ts[6]), (ops.BR_WIN, ts[3], ts[4])):
        g_tc = _grads_dev(ops, br, ts[0], K, V, cfg, None, dO, dtype, t0=t0)
        g_or = _grads_oracle(br, ts[0], K, V, None, dO, l, d, w, t0=t0)
        for nme, a, c in zip(("dQ", "dK", "dV"), g_tc, g_or):
            assert _rel(a, c) <= TOL_ORACLE, (br, nme, _rel(a, c))


def test_bwd_tc_sel_empty_and_clamped_rows():
    """Hand-made ranges: empty rows (no gradient), a partially valid last block, adjacent and repeated-block-free ranges."""
    ops = _ops()
    B, S, G, h = 1, 256, 1, 6
    dtype = torch.bfloat16
    gen = torch.Generator().manual_seed(5)
    Q, K, V = (torch.randn(*s, generator=gen).to(dtype).float() for s in ((B, S, G, h, 64), (B, G, S, 64), (B, G, S, 64)))
    ranges = torch.zeros(B, S, G, 4, 2, dtype=torch.int32)
    for t in range(S):
        if t % 7 == 0:
            continue  # empty row
        ranges[0, t, 0, 0] = torch.tensor([0, min(64, t + 1)])
        if t >= 128:
            ranges[0, t, 0, 1] = torch.tensor([64, 128]) if t % 2 else torch.tensor([128, t + 1])
    cfg = ops.NSAConfig(impl=ops.IMPL_TC)
    dO = torch.randn(B, S, G, h, 64, generator=gen).to(dtype).float()
    g_tc = _grads_dev(ops, ops.BR_SEL, Q, K, V, cfg, ranges.cuda(), dO, dtype)
    g_or = _grads_oracle(ops.BR_SEL, Q, K, V, ranges, dO, 32, 16, 512)
    for nme, a, c in zip(("dQ", "dK", "dV"), g_tc, g_or):
        assert torch.isfinite(a).all()
        assert _rel(a, c) <= TOL_ORACLE, (nme, _rel(a, c))
    assert torch.all(g_tc[0][0, 0] == 0) and torch.all(g_tc[0][0, 7] == 0)  # empty rows: no dQ


@pytest.mark.parametrize("gate_mode", ["mlp", "uniform"])
def test_prefill_core_backward_tc_m7c_vs_oracle(gate_mode):
    """Fused path at m7c head dims, bf16: all gradients (inputs and gate MLP) against autograd through the oracle."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 2, 640, 2, 6, 32, 16, 64, 16, 128
    dtype = torch.bfloat16
    ts = _case(B, S, G, h, l, d, seed=21, dtype=dtype)
    gen = torch.Generator().manual_seed(2)
    gate = (torch.randn(32, 64, generator=gen) * 0.3, torch.randn(32, generator=gen) * 0.1,
            torch.randn(3, 32, generator=gen) * 0.5, torch.randn(3, generator=gen) * 0.1)
    gm = ops.GATE_MLP if gate_mode == "mlp" else ops.GATE_UNIFORM
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC, gate_mode=gm)
    dev = [t.cuda().to(dtype).requires_grad_(True) for t in ts]
    gdev = tuple(t.cuda().requires_grad_(True) for t in gate)
    Oc, ranges, _ = ops.prefill_core(*dev, gdev if gate_mode == "mlp" else None, cfg, sel_mode=0)
    dO = torch.randn(Oc.shape, generator=gen).to(dtype).float()
    (Oc.float() * dO.cuda()).sum().backward()
    cpu = [t.clone().requires_grad_(True) for t in ts]
    gcpu = tuple(t.clone().requires_grad_(True) for t in gate)
    want = O.prefill_core(*cpu, gcpu, l=l, d=d, l_sel=ls, n_sel=n, w=w, ranges=ranges.cpu(), gate_mode=gm)
    (want["O"] * dO).sum().backward()
    for nme, a, b in zip(["Q", "K_sel", "V_sel", "K_win", "V_win", "K_cmp", "V_cmp"], dev, cpu):
        assert _rel(a.grad.float().cpu(), b.grad) <= TOL_ORACLE, (nme, _rel(a.grad.float().cpu(), b.grad))
    if gate_mode == "mlp":
        for nme, a, b in zip(["fc1_w", "fc1_b", "fc2_w", "fc2_b"], gdev, gcpu):
            assert _rel(a.grad.float().cpu(), b.grad) <= TOL_ORACLE, (nme, _rel(a.grad.float().cpu(), b.grad))


def test_bwd_tc_linearity_large():
    """Size-independent property at a training-size shape (S=2048, m7c dims): the backward is linear in dO, so
    grad(dO1 + dO2) = grad(dO1) + grad(dO2) up to the 16-bit rounding of dS, and grad(0) = 0 exactly."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 2, 2048, 2, 6, 32, 16, 64, 16, 512
    dtype = torch.bfloat16
    ts = _case(B, S, G, h, l, d, seed=33, dtype=dtype)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, impl=ops.IMPL_TC, gate_mode=ops.GATE_UNIFORM)
    gen = torch.Generator().manual_seed(4)
    dO1 = torch.randn(B, S, G, h, 64, generator=gen).to(dtype)
    dO2 = torch.randn(B, S, G, h, 64, generator=gen).to(dtype)

    def run(dO):
        dev = [t.cuda().to(dtype).requires_grad_(True) for t in ts]
        Oc, _, _ = ops.prefill_core(*dev, None, cfg, sel_mode=0)
        Oc.backward(dO.cuda())
        return [t.grad.float() for t in dev]

    g1, g2, g12, g0 = run(dO1), run(dO2), run((dO1.float() + dO2.float()).to(dtype)), run(torch.zeros_like(dO1))
    for a, b, c, z in zip(g1, g2, g12, g0):
        assert torch.all(z == 0)
        assert _rel(a + b, c) <= 3e-2, _rel(a + b, c)


def test_bwd_tc_long_sequence_vs_simt_same_saved_tensors(monkeypatch):
    """S = 16384 (256 selection blocks, 1023 compressed keys, chunked M-tile walks): the tensor-core backward against the SIMT
    backward fed the SAME saved O / lse / ranges (NSA_B200_IMPL=1 switches the kernel family at call time), so only the
    backward kernels differ.  Tolerance: rel-err 2e-2 per gradient (16-bit P and dS in the tensor-core path)."""
    ops = _ops()
    B, S, G, h, l, d, ls, n, w = 1, 16384, 2, 6, 32, 16, 64, 16, 512
    dtype = torch.bfloat16
    ts = _case(B, S, G, h, l, d, seed=44, dtype=dtype)
    cfg = ops.NSAConfig(l=l, d=d, l_sel=ls, n_sel=n, w=w, gate_mode=ops.GATE_UNIFORM)
    dev = [t.cuda().to(dtype).requires_grad_(True) for t in ts]
    Oc, _, _ = ops.prefill_core(*dev, None, cfg, sel_mode=0)
    dO = torch.randn(Oc.shape, generator=torch.Generator().manual_seed(6)).to(dtype).cuda()
    g_tc = torch.autograd.grad(Oc, dev, dO, retain_graph=True)
    monkeypatch.setenv("NSA_B200_IMPL", "1")
    g_si = torch.autograd.grad(Oc, dev, dO, retain_graph=True)
    monkeypatch.delenv("NSA_B200_IMPL")
    for nme, a, b in zip(["Q", "K_sel", "V_sel", "K_win", "V_win", "K_cmp", "V_cmp"], g_tc, g_si):
        assert torch.isfinite(a.float()).all(), nme
        assert _rel(a.float(), b.float()) <= TOL_SIMT, (nme, _rel(a.float(), b.float()))
