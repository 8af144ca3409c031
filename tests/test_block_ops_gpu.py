"""RMSNorm (+ residual add, + output cast) kernel, forward and backward, against the reference's formula
(nsa/model/llama_block_nsa.py:13-22, restated as model.llama_block_nsa.rmsnorm_torch) evaluated in fp32 with autograd.
Tolerances: fp32 in/out max-abs 2e-5 (the only difference is the summation order of mean(x^2) and of the dw column sums);
16-bit outputs one rounding step (2^-8 relative) around the fp32 result."""
import pytest
import torch

pytestmark = pytest.mark.gpu

_EPS = 1e-6


def _ref(x, w, r=None):
    from nsa_vibe_b200.model.llama_block_nsa import rmsnorm_torch
    s = x if r is None else x + r
    return s, rmsnorm_torch(s, w, _EPS)


def _close(a, b, dt):
    a, b = a.float(), b.float()
    tol = 2e-5 if dt == torch.float32 else 1.6e-2
    assert torch.allclose(a, b, atol=tol, rtol=tol), (dt, (a - b).abs().max())


@pytest.mark.parametrize("shape", [(2, 37, 768), (1, 1, 64), (5, 132), (3000, 256)])
@pytest.mark.parametrize("xdt,wdt,odt", [(torch.float32, torch.float32, torch.float32), (torch.float32, torch.float32, torch.bfloat16),
                                         (torch.bfloat16, torch.bfloat16, torch.bfloat16), (torch.float16, torch.float32, torch.float16)])
@pytest.mark.parametrize("with_res", [False, True])
def test_rmsnorm_forward_backward(shape, xdt, wdt, odt, with_res):
    from nsa_vibe_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(shape[0] + shape[-1])
    x32 = torch.randn(*shape, generator=g, device="cuda") * 1.7
    w32 = torch.rand(shape[-1], generator=g, device="cuda") + 0.5
    r32 = torch.randn(*shape, generator=g, device="cuda") if with_res else None
    rdt = torch.bfloat16 if (with_res and xdt == torch.float32 and odt == torch.bfloat16) else xdt  # residual of another dtype
    x = x32.to(xdt).requires_grad_(True)
    w = w32.to(wdt).requires_grad_(True)
    r = None if r32 is None else r32.to(rdt).requires_grad_(True)
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().float().requires_grad_(True)
    rr = None if r is None else r.detach().float().requires_grad_(True)
    got = ops.rmsnorm(x, w, _EPS, residual=r, out_dtype=odt)
    s_ref, y_ref = _ref(xr, wr, rr)
    if with_res:
        s, y = got
        assert s.dtype == xdt and y.dtype == odt
        _close(s, s_ref, xdt)
    else:
        y = got
        assert y.dtype == odt
    # a 16-bit running sum is rounded before the norm sees it (as in torch): compare against the norm of the rounded sum
    if with_res and xdt != torch.float32:
        _, y_ref2 = _ref(s.detach().float(), wr.detach())
        _close(y, y_ref2, odt)
    else:
        _close(y, y_ref, odt)
    dy = torch.randn(*shape, generator=g, device="cuda")
    dsum = torch.randn(*shape, generator=g, device="cuda") if with_res else None
    loss = (y.float() * dy).sum() + ((s.float() * dsum).sum() if with_res else 0.0)
    loss_ref = (y_ref * dy).sum() + ((s_ref * dsum).sum() if with_res else 0.0)
    loss.backward()
    loss_ref.backward()
    gt = 2e-4 if (xdt == torch.float32 and odt == torch.float32) else 6e-2
    scale = max(1.0, float(xr.grad.abs().max()))
    assert (x.grad.float() - xr.grad).abs().max() <= gt * scale, (x.grad.float() - xr.grad).abs().max()
    wscale = max(1.0, float(wr.grad.abs().max()))
    assert (w.grad.float() - wr.grad).abs().max() <= gt * wscale, ((w.grad.float() - wr.grad).abs().max(), wscale)
    if with_res:
        assert (r.grad.float() - rr.grad).abs().max() <= gt * scale


def test_rmsnorm_module_autocast_emits_bf16_equal_to_cast_of_fp32():
    """Under autocast the norm emits bf16 directly: bit-equal to casting the fp32 output (what the consuming nn.Linear would do)."""
    from nsa_vibe_b200.model.llama_block_nsa import RMSNorm
    torch.manual_seed(3)
    n = RMSNorm(768).cuda()
    x = torch.randn(4, 50, 768, device="cuda")
    y32 = n(x)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = n(x)
    assert y32.dtype == torch.float32 and y16.dtype == torch.bfloat16
    assert torch.equal(y16, y32.to(torch.bfloat16))


def test_rmsnorm_rejects_cpu_and_bad_dims():
    from nsa_vibe_b200 import ops
    with pytest.raises(RuntimeError):
        ops.rmsnorm(torch.randn(2, 8), torch.ones(8))
    with pytest.raises(RuntimeError):
        ops.rmsnorm(torch.randn(2, 6, device="cuda"), torch.ones(6, device="cuda"))


def test_block_stack_with_deferred_residual_matches_plain_stack_and_torch_norm():
    """Two LlamaBlockNSA blocks (llama_block_nsa.py:33-106): (a) the stack with every residual add folded into the next norm's
    kernel, (b) the plain stack, (c) the plain stack with the reference's ATen RMSNorm formula -- same outputs and gradients
    (fp32: 2e-4 max-abs on values, 2e-3 relative to the largest gradient)."""
    import os
    from nsa_vibe_b200.model import llama_block_nsa as M
    os.environ["NSA_PREFILL_BATCHED"] = "1"
    try:
        torch.manual_seed(11)
        blocks = [M.LlamaBlockNSA(64, 4, 2, 16, 16, l=16, d=8, l_sel=32, n_sel=4, w=40).cuda() for _ in range(2)]
    finally:
        os.environ.pop("NSA_PREFILL_BATCHED", None)
    x0 = torch.randn(2, 96, 64, device="cuda")

    def run(mode):
        for b in blocks:
            b.zero_grad(set_to_none=True)
        x = x0.clone().requires_grad_(True)
        if mode == "deferred":
            h, delta = x, None
            for b in blocks:
                h, delta = b(h, delta, defer_residual=True)
            y = h + delta
        else:
            y = x
            for b in blocks:
                y = b(y)
        y.square().sum().backward()
        return y.detach(), x.grad.detach(), [p.grad.detach().clone() for b in blocks for p in b.parameters()]

    ya, ga, pa = run("deferred")
    yb, gb, pb = run("plain")
    orig = M.RMSNorm.forward
    M.RMSNorm.forward = lambda self, x, residual=None: (
        M.rmsnorm_torch(x, self.weight, self.eps) if residual is None else (x + residual, M.rmsnorm_torch(x + residual, self.weight, self.eps)))
    try:
        yc, gc, pc = run("plain")
    finally:
        M.RMSNorm.forward = orig
    for y, g_, p_ in ((yb, gb, pb), (yc, gc, pc)):
        assert (ya - y).abs().max() <= 2e-4, (ya - y).abs().max()
        assert (ga - g_).abs().max() <= 2e-3 * max(1.0, float(g_.abs().max()))
        for u, v in zip(pa, p_):
            assert (u - v).abs().max() <= 2e-3 * max(1.0, float(v.abs().max())), (u.shape, (u - v).abs().max())


@pytest.mark.parametrize("n_rows", [1, 37, 5000, 200003])
def test_device_stats_match_the_reference_formulas(n_rows):
    """ops.stats against the reference's ATen formulas (nsa_attention.py:127-165 gate health, :455-507 selection statistics):
    integers and extrema exactly, means to 1e-6."""
    from nsa_vibe_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(n_rows)
    logits = torch.randn(n_rows, 3, generator=g, device="cuda") * 4.0
    logits[::7] = torch.tensor([60.0, 0.0, -3.0], device="cuda")  # collapsed rows
    gates = torch.softmax(logits, dim=-1)
    K = 5
    starts = torch.randint(0, 500, (n_rows, K), generator=g, device="cuda", dtype=torch.int32)
    lens = torch.randint(-3, 64, (n_rows, K), generator=g, device="cuda", dtype=torch.int32)  # negative: empty / padded ranges
    ranges = torch.stack([starts, starts + lens], dim=-1)
    got = ops.stats(gates=gates.view(-1, 1, 3) if n_rows > 1 else gates, ranges=ranges.view(n_rows, 1, K, 2))
    ent = -(gates * (gates + 1e-8).log()).sum(dim=-1)
    mx = gates.max(dim=-1)[0]
    assert got["total_gates"] == n_rows and got["rows"] == n_rows
    assert abs(got["entropy_mean"] - ent.double().mean().item()) <= 1e-6 and abs(got["max_gate_mean"] - mx.double().mean().item()) <= 1e-6
    assert abs(got["entropy_min"] - ent.min().item()) <= 1e-7 and got["max_gate_max"] == mx.max().item()
    assert abs(got["collapse_fraction"] - ((ent < 0.1) & (mx > 0.95)).double().mean().item()) <= 1e-12
    for a, b in zip(got["branch_shares"], gates.double().mean(dim=0).tolist()):
        assert abs(a - b) <= 1e-6
    L = (ranges[..., 1] - ranges[..., 0]).clamp_min(0).sum(dim=-1).to(torch.int64)
    assert got["k_max"] == int(L.max()) and abs(got["k_mean"] - L.double().mean().item()) <= 1e-9
    assert abs(got["pct_at_max"] - (L == L.max()).double().mean().item()) <= 1e-12
    assert ops.stats(ranges=torch.zeros(0, 2, 4, 2, dtype=torch.int32, device="cuda")) == {"k_mean": 0.0, "k_max": 0, "rows": 0, "pct_at_max": 0.0}
