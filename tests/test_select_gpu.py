"""Selection kernel parity (bit-exact indices).  Calls go through the C ABI (nsa_select_ranges_*)."""
import numpy as np
import pytest
import torch

from conftest import T, load_golden
from oracle import nsa_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from nsa_vibe_b200 import ops
    return ops


def test_prefill_golden_bit_exact():
    ops = _ops()
    g = load_golden("select")
    for i in range(int(g["pre_n"])):
        ls, ns, S = [int(v) for v in g[f"pre_c{i}"]]
        ref = T(g[f"pre_r{i}"])
        got = ops.select_ranges_prefill(T(g[f"pre_p{i}"]).cuda(), ls, ns, S).cpu()
        assert got.shape == ref.shape, (i, got.shape, ref.shape)
        assert torch.equal(got, ref), f"prefill case {i} (l_sel={ls}, n={ns}, S={S}): {(got != ref).any(-1).any(-1).sum()} rows differ"


def test_decode_golden_equivalent():
    ops = _ops()
    g = load_golden("select")
    for i in range(int(g["dec_n"])):
        ls, ns, t = [int(v) for v in g[f"dec_c{i}"]]
        got = ops.select_ranges_decode(T(g[f"dec_p{i}"]).cuda(), ls, ns, t).cpu()
        ok, bad = O.ranges_equivalent(got, T(g[f"dec_r{i}"]))
        assert ok, f"decode case {i} (l_sel={ls}, n={ns}, t={t}): {bad} rows differ"
        assert torch.equal(got, O.select_ranges_decode(T(g[f"dec_p{i}"]), ls, ns, t))  # same padding as the oracle


def _oracle_prefill_row(p_row, t, ls, n, S):
    """Oracle ranges of ONE row t (p_row [B,G,S_sel]): the rule is row-local, so feed the row at index t of an
    otherwise empty tensor."""
    B, G, S_sel = p_row.shape
    pr = torch.zeros(B, t + 1, G, S_sel)
    pr[:, t] = p_row
    return O.select_ranges_prefill(pr[:, t:t + 1] if t == 0 else pr, ls, n, S)[:, t]


@pytest.mark.parametrize("S,ls,n", [(2048, 64, 16), (1000, 32, 8), (130, 64, 16), (4096, 64, 16), (333, 16, 5)])
def test_prefill_vs_oracle_random(S, ls, n):
    ops = _ops()
    gen = torch.Generator().manual_seed(S + n)
    S_sel = O.num_sel_blocks(S, ls)
    p = torch.rand(2, S, 2, S_sel, generator=gen)
    got = ops.select_ranges_prefill(p.cuda(), ls, n, S).cpu()
    want = O.select_ranges_prefill(p, ls, n, S)
    assert torch.equal(got, want), f"{(got != want).any(-1).any(-1).sum()} rows differ"


def test_ties_prefer_lower_index():
    ops = _ops()
    S, ls, n = 1024, 64, 8
    S_sel = S // ls
    p = torch.ones(1, S, 1, S_sel)
    got = ops.select_ranges_prefill(p.cuda(), ls, n, S).cpu()
    want = O.select_ranges_prefill(p, ls, n, S)
    assert torch.equal(got, want)
    t = S - 1  # forced {0,14,15(valid at the last token)}, picks 1..5 -> [0,384) and [896,1024)
    assert O.nonempty_ranges(got[0, t, 0].tolist()) == [(0, 6 * ls), (14 * ls, 16 * ls)]


def test_64k_properties_and_sample():
    ops = _ops()
    S, ls, n = 65536, 64, 16
    S_sel = S // ls
    gen = torch.Generator(device="cuda").manual_seed(0)
    p = torch.rand(1, S, 2, S_sel, generator=gen, device="cuda")
    r = ops.select_ranges_prefill(p, ls, n, S)
    assert r.shape == (1, S, 2, 16, 2)
    s, e = r[..., 0].long(), r[..., 1].long()
    t = torch.arange(S, device="cuda").view(1, S, 1, 1)
    assert bool((e <= t + 1).all()) and bool((s >= 0).all())                       # causality
    ne = e > s
    assert bool((((s % ls) == 0) | ~ne).all()) and bool((((e % ls) == 0) | ~ne).all())  # whole blocks in prefill mode
    nxt_s = torch.where(ne[..., 1:], s[..., 1:], torch.full_like(s[..., 1:], 1 << 40))
    assert bool(((e[..., :-1] < nxt_s) | ~ne[..., :-1]).all())                      # sorted, merged (gap between runs)
    tot = (e - s).clamp_min(0).sum(-1)
    # once >= n complete blocks exist: {0, cb-1} + 13 picks, plus cb itself only on the last token of a block (F2)
    full = ((t[..., 0] + 1) % ls == 0).long()
    assert bool((tot == (n - 1 + full) * ls)[:, 17 * ls:].all())
    assert bool((s[:, ls - 1:, :, 0] == 0).all())                                   # block 0 forced once complete
    rows = [0, 63, 64, 127, 128, 1000, 4095, 4096, 30000, 65535]
    pc = p.cpu()
    for tt in rows:
        want = O.select_ranges_prefill(pc[:, :tt + 1], ls, n, S)[:, tt] if tt < 200 else _row_oracle_big(pc, tt, ls, n, S)
        assert torch.equal(r[:, tt].cpu(), want), f"t={tt}"


def _row_oracle_big(pc, tt, ls, n, S):
    # evaluate the oracle for one row without looping over all earlier rows
    pr = pc[:, tt:tt + 1]
    p = pr[0, 0].numpy()
    res = []
    for g in range(p.shape[0]):
        nvalid = min((tt + 1) // ls, p.shape[1])
        cb = tt // ls
        forced = sorted({0, cb, max(cb - 1, 0)})
        masked = p[g].copy()
        masked[nvalid:] = -np.inf
        for j in forced:
            masked[j] = -np.inf
        comp = O._composite(masked)
        picks = [int(j) for j in O._rank_desc_lower_index(comp, n - 3)]
        ids = sorted(j for j in list(forced) + picks if j < nvalid)
        runs = []
        for j in ids:
            if runs and j - runs[-1][1] in (0, 1):
                runs[-1][1] = j
            else:
                runs.append([j, j])
        row = [[a * ls, min((b + 1) * ls, tt + 1)] for a, b in runs] + [[0, 0]] * (16 - len(runs))
        res.append(row)
    return torch.tensor(res, dtype=torch.int32)[None]


def test_decode_long_context():
    ops = _ops()
    ls, n = 64, 16
    for t in (4095, 65534, 70000):
        S_sel = O.num_sel_blocks(t + 1, ls)
        p = torch.rand(3, 2, S_sel, generator=torch.Generator().manual_seed(t))
        got = ops.select_ranges_decode(p.cuda(), ls, n, t).cpu()
        assert torch.equal(got, O.select_ranges_decode(p, ls, n, t))


@pytest.mark.parametrize("pattern", ["quantised", "ones", "one_lane", "peaked", "zeros", "ramp"])
def test_threshold_path_structured_rows(pattern):
    """128 < S_sel <= 1024 takes the threshold form of the picks (select.cuh: select_threshold_1024) and falls back to the rounds
    when more than 32 candidates reach the threshold: rows built to sit on both sides of that switch, and on exact ties."""
    ops = _ops()
    S, ls, n = 32768, 64, 16
    S_sel = S // ls
    gen = torch.Generator().manual_seed(11)
    rows = sorted(set([ls * 16 - 1, ls * 16, ls * 17 + 5, 2047, 2048, 8191, 8200, 20000, S - 1] +
                      [int(v) for v in torch.randint(ls * 13, S, (40,), generator=gen)]))
    if pattern == "quantised":
        p = torch.randint(0, 4, (1, S, 2, S_sel), generator=gen).float() / 4
    elif pattern == "ones":
        p = torch.ones(1, S, 2, S_sel)
    elif pattern == "one_lane":  # every large value in the blocks j = 5 mod 32: few lanes hold the top of the row
        p = torch.rand(1, S, 2, S_sel, generator=gen) * 0.5
        p[..., 5::32] += 1.0
    elif pattern == "peaked":
        p = torch.zeros(1, S, 2, S_sel)
        idx = torch.randint(0, S_sel, (1, S, 2, 6), generator=gen)
        p.scatter_(-1, idx, torch.rand(1, S, 2, 6, generator=gen) + 0.5)
    elif pattern == "zeros":
        p = torch.zeros(1, S, 2, S_sel)
    else:
        p = torch.arange(S_sel).float().expand(1, S, 2, S_sel).contiguous() / S_sel
    got = ops.select_ranges_prefill(p.cuda(), ls, n, S).cpu()
    for tt in rows:
        want = _row_oracle_big(p, tt, ls, n, S)
        assert torch.equal(got[:, tt], want), f"{pattern}: t={tt}\n{got[0, tt, 0].tolist()}\n{want[0, 0].tolist()}"
