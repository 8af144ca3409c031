"""CPU oracle for the NSA hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``nsa_vibe_b200``) never imports it and raises if the CUDA library is missing.

It is a restatement (numpy / torch-CPU, no reference imports) of the algorithm of
seconds-0/nsa-vibe for the path BASELINE.json names.  Every function cites the reference
file:line it follows (paths relative to the reference checkout).  Where the reference's
default route is degenerate (``is_causal=True`` with one query row attends to key 0 only,
SURVEY.md section 0 F1) the oracle implements the *intended* semantics (softmax over all
allowed keys) and offers ``literal_first_key`` so a parity report can show both.

Parity pinning: ``tests/golden/make_golden.py`` imports the real reference in the build
container and writes fixtures (``tests/golden/*.npz``); ``tests/test_oracle_golden.py``
checks every function below against them.  Status per stage:
  * Eq.9 weights, p_cmp -> p_slc -> p_grp, top-n + ranges (both modes), sel attention,
    sliding attention, gate MLP, rope, phi avg-pool, cache emission: PINNED on reference
    outputs generated here.
  * compressed-branch true softmax and the gated three-branch output: the reference has
    no non-degenerate implementation, so these are pinned only through the reference's
    own mask/length helpers + a reference SDPA call in the golden script ("restatement
    pinned on reference primitives"); see DESIGN.md.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

NEG_INF = float("-inf")


# ----------------------------------------------------------------------------------------
# Block geometry and the Eq.9 map                      (nsa/core/block_index.py:25-99)
# ----------------------------------------------------------------------------------------
@dataclass
class OracleMeta:
    l: int
    d: int
    l_sel: int
    n_sel: int
    w: int
    S_cmp: int
    S_sel: int
    coo_rows: np.ndarray  # [nnz] int64, ascending compressed index
    coo_cols: np.ndarray  # [nnz] int64 selection block
    coo_vals: np.ndarray  # [nnz] float32 = overlap / total overlap


def num_cmp_blocks(seq_len: int, l: int, d: int) -> int:
    """block_index.py:31  --  0 if seq_len < l else (seq_len - l)//d + 1."""
    return 0 if seq_len < l else (seq_len - l) // d + 1


def num_sel_blocks(seq_len: int, l_sel: int) -> int:
    """block_index.py:34  --  ceil(seq_len / l_sel)."""
    return 0 if seq_len <= 0 else (seq_len + l_sel - 1) // l_sel


def build_meta(seq_len: int, l: int, d: int, l_sel: int, n_sel: int, w: int) -> OracleMeta:
    """Closed-form restatement of build_M_csl_csr (block_index.py:43-71) + COO
    (block_index.py:74-99): compressed block i covers tokens [i*d, i*d+l); selection block j
    covers [j*l_sel, (j+1)*l_sel) for j < S_sel; weight = overlap / sum of overlaps over
    the selection blocks that exist (so a compressed block hanging past the last selection
    block is renormalised), rows ascending in i, columns ascending in j inside a row."""
    if l % d != 0 or l_sel % d != 0:
        raise ValueError("Require d|l and d|l_sel in M0")  # block_index.py:75-77
    S_cmp = num_cmp_blocks(seq_len, l, d)
    S_sel = num_sel_blocks(seq_len, l_sel)
    rows: List[int] = []
    cols: List[int] = []
    vals: List[float] = []
    for i in range(S_cmp):
        a0, a1 = i * d, i * d + l
        j_lo = a0 // l_sel
        j_hi = min((a1 - 1) // l_sel, S_sel - 1)
        ovs = []
        for j in range(j_lo, j_hi + 1):
            ov = max(0, min(a1, (j + 1) * l_sel) - max(a0, j * l_sel))
            if ov > 0:
                ovs.append((j, ov))
        tot = sum(o for _, o in ovs)
        for j, ov in ovs:
            rows.append(i)
            cols.append(j)
            vals.append(ov / tot)  # python float -> float32 on store (torch.tensor(.., float32))
    return OracleMeta(
        l, d, l_sel, n_sel, w, S_cmp, S_sel,
        np.asarray(rows, dtype=np.int64),
        np.asarray(cols, dtype=np.int64),
        np.asarray(vals, dtype=np.float32),
    )


# ----------------------------------------------------------------------------------------
# RoPE and phi                       (nsa/core/rope.py:16-51, nsa/core/compress_pool.py:9-38)
# ----------------------------------------------------------------------------------------
def rope(x: torch.Tensor, pos: torch.Tensor, base: float = 10000.0, scale: float = 1.0) -> torch.Tensor:
    """Interleaved-pair rotary embedding, fp32 angles, sin/cos cast to x.dtype (rope.py:30-51)."""
    D = x.shape[-1]
    half = D // 2
    inv_freq = base ** (-2.0 * torch.arange(half, dtype=torch.float32) / D)
    if scale <= 0:
        scale = 1.0
    ang = (pos.to(torch.float32) / float(scale))[..., None] * inv_freq  # [S, D/2]
    sin, cos = torch.sin(ang).to(x.dtype), torch.cos(ang).to(x.dtype)
    xe, xo = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = xe * cos - xo * sin
    out[..., 1::2] = xe * sin + xo * cos
    return out


def phi_avg_pool(K_raw: torch.Tensor, V_raw: torch.Tensor, l: int, d: int,
                 pos: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """K_cmp[i] = mean_{s in [i*d, i*d+l)} RoPE(K_raw[s]); V_cmp[i] = mean V_raw[s]
    (compress_pool.py:9-38; V is not rotated; NSA_ROPE_SCALE is ignored there)."""
    B, G, S, Dk = K_raw.shape
    if pos is None:
        pos = torch.arange(S)
    Kr = rope(K_raw, pos)
    n = num_cmp_blocks(S, l, d)
    if n == 0:
        return K_raw.new_zeros(B, G, 0, Dk), V_raw.new_zeros(B, G, 0, V_raw.shape[-1])
    Kc = torch.stack([Kr[:, :, i * d:i * d + l].mean(dim=2) for i in range(n)], dim=2)
    Vc = torch.stack([V_raw[:, :, i * d:i * d + l].mean(dim=2) for i in range(n)], dim=2)
    return Kc, Vc


def phi_conv(K_raw: torch.Tensor, V_raw: torch.Tensor, w_k: torch.Tensor, w_v: torch.Tensor, l: int, d: int,
             pos: Optional[torch.Tensor] = None, rope_scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Learnable phi (phi="mlp"): depthwise Conv1d over time, kernel l, stride d, no bias, on RoPE(K_raw) and V_raw
    (_phi_apply_seq, nsa_attention.py:1741-1758); w_* [D, l] (Conv1d.weight [D,1,l] squeezed):
    K_cmp[i, e] = sum_r w_k[e, r] * RoPE(K_raw)[i*d + r, e]."""
    B, G, S, Dk = K_raw.shape
    if pos is None:
        pos = torch.arange(S)
    Kr = rope(K_raw, pos, scale=rope_scale)
    n = num_cmp_blocks(S, l, d)
    if n == 0:
        return K_raw.new_zeros(B, G, 0, Dk), V_raw.new_zeros(B, G, 0, V_raw.shape[-1])
    wk, wv = w_k.reshape(Dk, l), w_v.reshape(V_raw.shape[-1], l)
    Kc = torch.stack([(Kr[:, :, i * d:i * d + l] * wk.t()).sum(dim=2) for i in range(n)], dim=2)
    Vc = torch.stack([(V_raw[:, :, i * d:i * d + l] * wv.t()).sum(dim=2) for i in range(n)], dim=2)
    return Kc, Vc


def num_cmp_at(t: int, l: int, d: int, S_cmp: int) -> int:
    """packing.py:15-23 / attention_kernels.py:121  --  compressed tokens visible at row t."""
    return 0 if t + 1 < l else min((t + 1 - l) // d + 1, S_cmp)


# ----------------------------------------------------------------------------------------
# Scores: p_cmp, Eq.9, Eq.10                        (nsa/core/selection_scorer.py:42-121)
# ----------------------------------------------------------------------------------------
def pcmp_all(Q: torch.Tensor, K_cmp: torch.Tensor, scale: float, norm: str = "full_row",
             l: int = 0, d: int = 0, t0: int = 0) -> torch.Tensor:
    """Q [B,S,G,h,Dk], K_cmp [B,G,S_cmp,Dk] -> p [B,S,G,h,S_cmp].
    norm="full_row": softmax over every compressed key for every row, the reference's prefill
    behaviour (selection_scorer.py:58-61; non-causal normaliser).
    norm="causal": softmax over the first num_cmp(t) keys only (what decode sees)."""
    logits = torch.einsum("bsghd,bgcd->bsghc", Q.float(), K_cmp.float()) * scale
    if norm == "causal":
        S, S_cmp = Q.shape[1], K_cmp.shape[2]
        nc = torch.tensor([num_cmp_at(t0 + t, l, d, S_cmp) for t in range(S)])
        dis = torch.arange(S_cmp)[None, :] >= nc[:, None]
        logits = logits.masked_fill(dis[None, :, None, None, :], NEG_INF)
        p = torch.softmax(logits, dim=-1)
        return torch.nan_to_num(p, nan=0.0)
    return torch.softmax(logits, dim=-1)


def pslc_from_pcmp(p: torch.Tensor, meta: OracleMeta) -> torch.Tensor:
    """Eq.9: p_slc[..., c] += p[..., r] * w for COO entries with r < S_cmp present, accumulated
    in ascending r (CPU scatter_add order; selection_scorer.py:89-116)."""
    S_cmp = p.shape[-1]
    out = torch.zeros(*p.shape[:-1], meta.S_sel, dtype=p.dtype)
    for r, c, w in zip(meta.coo_rows, meta.coo_cols, meta.coo_vals):
        if r < S_cmp:
            out[..., int(c)] += p[..., int(r)] * torch.tensor(w, dtype=p.dtype)
    return out


def pgrp_from_pslc(p_slc: torch.Tensor) -> torch.Tensor:
    """Eq.10: sum over the h heads of the group (nsa_attention.py:1091, :670)."""
    return p_slc.sum(dim=-2)


# ----------------------------------------------------------------------------------------
# Top-n selection and ranges                       (nsa/core/selection_scorer.py:124-605)
# ----------------------------------------------------------------------------------------
def _rank_desc_lower_index(composite: np.ndarray, k: int) -> np.ndarray:
    """Top-k by descending fp32 composite, exact ties -> lower index (documented intent,
    selection_scorer.py:179-181)."""
    order = np.argsort(-composite.astype(np.float32), kind="stable")
    return order[:k]


def _composite(masked: np.ndarray) -> np.ndarray:
    """fp32(score) - fp32(fp32(j) * fp32(1e-8))   (selection_scorer.py:182-184, :312-318)."""
    S_sel = masked.shape[-1]
    bias = (np.arange(S_sel, dtype=np.float32) * np.float32(1e-8)).astype(np.float32)
    return (masked.astype(np.float32) - bias).astype(np.float32)


def select_ranges_decode(p_grp: torch.Tensor, l_sel: int, n_sel: int, t: int, force_init: bool = True,
                         force_local: int = 2) -> torch.Tensor:
    """select_topn_ranges (selection_scorer.py:124-249): p_grp [B,G,S_sel] -> [B,G,n_sel,2] int32.
    Forced {0, cb, max(cb-1,0)} always included; other picks from valid blocks
    ((j+1)*l_sel <= t+1) by composite; sort, dedup, merge adjacent, clamp end to t+1, pad [0,0].
    Picks that fall on -inf (future) blocks are dropped here: the reference keeps them as
    rows with end <= start which every consumer skips."""
    p = p_grp.detach().float().numpy()
    B, G, S_sel = p.shape
    out = np.zeros((B, G, n_sel, 2), dtype=np.int32)
    cb = max(t // l_sel, 0)
    forced = ([0] if force_init else []) + [max(cb - i, 0) for i in range(force_local)]  # selection_scorer.py:159-170
    k_rest = max(n_sel - len(forced), 0)
    for b in range(B):
        for g in range(G):
            masked = p[b, g].copy()
            for j in range(S_sel):
                if (j + 1) * l_sel > t + 1:
                    masked[j] = NEG_INF
            for j in forced:
                if j < S_sel:
                    masked[j] = NEG_INF
            picks: List[int] = []
            if k_rest > 0:
                comp = _composite(masked)
                for j in _rank_desc_lower_index(comp, min(k_rest, S_sel)):
                    if comp[j] > NEG_INF:
                        picks.append(int(j))
            blocks = sorted(set([j for j in forced if j < S_sel] + picks))
            merged: List[List[int]] = []
            for j in blocks:
                s0 = j * l_sel
                if merged and merged[-1][1] == s0:
                    merged[-1][1] = s0 + l_sel
                else:
                    merged.append([s0, s0 + l_sel])
            for i, (s0, e0) in enumerate(merged[:n_sel]):
                out[b, g, i, 0] = s0
                out[b, g, i, 1] = min(e0, t + 1)
    return torch.from_numpy(out)


def prefill_forced_cols(S: int, l_sel: int) -> int:
    """Number of forced columns after the reference's column-wise unique_consecutive
    (selection_scorer.py:299-300): 1 if every row sits in block 0, 2 if at most two blocks."""
    if S <= l_sel:
        return 1
    if S <= 2 * l_sel:
        return 2
    return 3


def prefill_forced_table(S: int, l_sel: int, force_init: bool = True, force_local: int = 2) -> np.ndarray:
    """Forced block ids of every row t < S as the reference builds them (selection_scorer.py:283-300): [0 if force_init] +
    [max(cb - k, 0) for k < force_local], sorted per row, then COLUMNS equal to their left neighbour on every row are dropped
    (unique_consecutive(dim=-1) acts on whole columns).  Returns [S, nf] int."""
    rows = []
    for t in range(S):
        cb = t // l_sel
        rows.append(sorted(([0] if force_init else []) + [max(cb - k, 0) for k in range(force_local)]))
    arr = np.array(rows, dtype=np.int64).reshape(S, -1)
    keep = [c for c in range(arr.shape[1]) if c == 0 or not np.array_equal(arr[:, c], arr[:, c - 1])]
    return arr[:, keep]


def prefill_range_cols(S: int, l_sel: int, n_sel: int, force_init: bool = True, force_local: int = 2) -> int:
    """Width K of the batched ranges tensor [B,S,G,K,2]."""
    S_sel = num_sel_blocks(S, l_sel)
    if n_sel >= S_sel:
        return S_sel
    nf = prefill_forced_table(S, l_sel, force_init, force_local).shape[1]
    k_rest = max(0, n_sel - nf)
    return nf + min(k_rest, S_sel) if k_rest > 0 else min(nf, n_sel)


def select_ranges_prefill(p_grp: torch.Tensor, l_sel: int, n_sel: int, S: int, t0: int = 0, force_init: bool = True,
                          force_local: int = 2) -> torch.Tensor:
    """select_topn_ranges_batched + convert_indices_to_ranges_batched_v2
    (selection_scorer.py:255-362, :434-605): p_grp [B,S,G,S_sel] -> [B,S,G,K,2] int32.
    Every entry (forced included) must be a *complete* block at row t; if n_sel >= S_sel the
    row is all valid blocks; runs of equal/+1 ids merge; end clamped to t+1; left aligned."""
    p = p_grp.detach().float().numpy()
    B, S_q, G, S_sel = p.shape
    table = prefill_forced_table(S, l_sel, force_init, force_local)
    nf = table.shape[1]
    K = prefill_range_cols(S, l_sel, n_sel, force_init, force_local)
    k_rest = max(0, n_sel - nf)
    out = np.zeros((B, S_q, G, K, 2), dtype=np.int32)
    for b in range(B):
        for tq in range(S_q):
            t = t0 + tq  # absolute position of this row
            nvalid = min((t + 1) // l_sel, S_sel)
            forced = [int(v) for v in table[t]] if t < S else []
            for g in range(G):
                if n_sel >= S_sel:
                    ids = list(range(nvalid))
                else:
                    masked = p[b, tq, g].copy()
                    masked[nvalid:] = NEG_INF
                    for j in forced:
                        if j < S_sel:
                            masked[j] = NEG_INF
                    picks: List[int] = []
                    if k_rest > 0:
                        comp = _composite(masked)
                        picks = [int(j) for j in _rank_desc_lower_index(comp, min(k_rest, S_sel))]
                        sel = list(forced) + picks
                    else:
                        sel = list(forced)[:n_sel]
                    ids = sorted(j for j in sel if j < nvalid)  # invalid -> -1 -> dropped
                runs: List[List[int]] = []
                for j in ids:
                    if runs and j - runs[-1][1] in (0, 1):
                        runs[-1][1] = j
                    else:
                        runs.append([j, j])
                for i, (j0, j1) in enumerate(runs):
                    out[b, tq, g, i, 0] = j0 * l_sel
                    out[b, tq, g, i, 1] = min((j1 + 1) * l_sel, t + 1)
    return torch.from_numpy(out)


def indices_to_ranges(indices: torch.Tensor, S_sel: int, l_sel: int, t0: int = 0) -> torch.Tensor:
    """convert_indices_to_ranges_batched (selection_scorer.py:380-431; v2 :434-605 gives the same runs, padded to K columns):
    ids [B,S,G,K] ascending, negative = padding -> [B,S,G,K,2] int32."""
    idx = indices.numpy()
    B, S, G, K = idx.shape
    out = np.zeros((B, S, G, K, 2), dtype=np.int32)
    for b in range(B):
        for s in range(S):
            for g in range(G):
                spans: List[List[int]] = []
                prev = None
                for bid in idx[b, s, g].tolist():
                    if bid < 0 or bid >= S_sel or bid == prev:
                        continue
                    prev = bid
                    s0 = bid * l_sel
                    e0 = min(s0 + l_sel, t0 + s + 1)
                    if e0 <= s0:
                        continue
                    if spans and spans[-1][1] == s0:
                        spans[-1][1] = e0
                    else:
                        spans.append([s0, e0])
                for i, (s0, e0) in enumerate(spans):
                    out[b, s, g, i] = (s0, e0)
    return torch.from_numpy(out)


def nonempty_ranges(row: Sequence[Sequence[int]]) -> List[Tuple[int, int]]:
    """The reference's own equivalence criterion: ordered list of ranges with end > start
    (nsa/tests/test_selection_v2_equiv.py:80-111)."""
    return [(int(s), int(e)) for s, e in row if int(e) > int(s)]


def ranges_equivalent(a: torch.Tensor, b: torch.Tensor) -> Tuple[bool, int]:
    """Compare two ranges tensors [..., K, 2] (K may differ) row by row with `nonempty_ranges`.
    Returns (all_equal, number_of_differing_rows)."""
    A = a.reshape(-1, a.shape[-2], 2).tolist()
    Bm = b.reshape(-1, b.shape[-2], 2).tolist()
    assert len(A) == len(Bm)
    bad = sum(1 for ra, rb in zip(A, Bm) if nonempty_ranges(ra) != nonempty_ranges(rb))
    return bad == 0, bad


def blocks_of_ranges(row: Sequence[Sequence[int]], l_sel: int) -> set:
    """Selection-block ids covered by a row of [start, end) ranges (the last block of a range may be clamped)."""
    out = set()
    for s0, e0 in nonempty_ranges(row):
        out.update(range(s0 // l_sel, (e0 + l_sel - 1) // l_sel))
    return out


def selection_difference_is_near_tie(p_row: torch.Tensor, row_a, row_b, l_sel: int, n_sel: int, t: int, delta: float,
                                     mode: int = 0) -> bool:
    """True when two selections of one row differ only by candidates whose fp32 composite lies within 2*delta of the cut (the
    n-th best candidate of `p_row`): a scorer whose p_grp is within `delta` of p_row can order such candidates either way,
    and no others.  mode 0 = prefill rule, 1 = decode rule (forced blocks {0, cb-1, cb} are never candidates)."""
    a, b = blocks_of_ranges(row_a, l_sel), blocks_of_ranges(row_b, l_sel)
    diff = a ^ b
    if not diff:
        return True
    p = p_row.detach().float().numpy()
    S_sel = p.shape[0]
    nvalid = min((t + 1) // l_sel, S_sel)
    cb = t // l_sel
    forced = {0, cb, max(cb - 1, 0)}
    masked = p.copy()
    masked[nvalid:] = NEG_INF
    for j in forced:
        if j < S_sel:
            masked[j] = NEG_INF
    comp = _composite(masked)
    k = min(max(n_sel - 3, 0), S_sel)
    order = np.sort(comp)[::-1]
    if k == 0 or k > order.shape[0] or not np.isfinite(order[k - 1]):
        return False
    cut = float(order[k - 1])
    return all(j < S_sel and np.isfinite(comp[j]) and abs(float(comp[j]) - cut) <= 2.0 * delta + 1e-12 for j in diff)


# ----------------------------------------------------------------------------------------
# Branch attention (intended semantics: softmax over every allowed key)
# ----------------------------------------------------------------------------------------
def _masked_attention(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor,
                      allowed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Q [B,S,G,h,Dk], K [B,G,Skv,Dk], V [B,G,Skv,Dv], allowed [B,S,G,Skv] bool.
    Returns (O [B,S,G,h,Dv], LSE [B,S,G,h]); rows with no allowed key give O = 0, LSE = -inf
    (attention_kernels.py:737-771 zeroes empty rows the same way)."""
    Dk = Q.shape[-1]
    dt = torch.float64 if Q.dtype == torch.float64 else torch.float32
    logits = torch.einsum("bsghd,bgkd->bsghk", Q.to(dt), K.to(dt)) / math.sqrt(Dk)
    logits = logits.masked_fill(~allowed[:, :, :, None, :], NEG_INF)
    lse = torch.logsumexp(logits, dim=-1)
    p = torch.exp(logits - torch.where(torch.isinf(lse), torch.zeros_like(lse), lse)[..., None])
    p = torch.where(allowed[:, :, :, None, :], p, torch.zeros_like(p))
    O = torch.einsum("bsghk,bgkd->bsghd", p, V.to(dt))
    return O, lse


def allowed_from_ranges(ranges: torch.Tensor, S_kv: int) -> torch.Tensor:
    """ranges [B,S,G,n,2] -> allowed [B,S,G,S_kv]: union of [s,e) with e > s, clamped to the
    cache (attention_kernels.py:724-735)."""
    B, S, G, n, _ = ranges.shape
    pos = torch.arange(S_kv)
    s = ranges[..., 0].clamp(0, S_kv).long()[..., None]
    e = ranges[..., 1].clamp(0, S_kv).long()[..., None]
    return ((pos >= s) & (pos < e)).any(dim=3)


def sel_attention(Q, K, V, ranges):
    """grouped_selection_attention_masked (attention_kernels.py:705-772)."""
    return _masked_attention(Q, K, V, allowed_from_ranges(ranges, K.shape[2]))


def win_attention(Q, K, V, w: int, t0: int = 0):
    """sliding_window_attention (attention_kernels.py:146-178): row t sees keys
    [t-w+1 .. t]; `t0` = absolute position of row 0 when K holds the full history."""
    B, S, G = Q.shape[:3]
    S_kv = K.shape[2]
    if w <= 0 or S_kv == 0:
        dv = V.shape[-1]
        return Q.new_zeros(B, S, G, Q.shape[3], dv, dtype=torch.float32), Q.new_full((B, S, G, Q.shape[3]), NEG_INF, dtype=torch.float32)
    row = torch.arange(S)[:, None] + t0
    col = torch.arange(S_kv)[None, :]
    allowed = (col <= row) & (col >= row - (w - 1))
    return _masked_attention(Q, K, V, allowed[None, :, None, :].expand(B, S, G, S_kv))


def cmp_attention(Q, K_cmp, V_cmp, l: int, d: int, t0: int = 0):
    """Restatement (the reference's routes are degenerate): softmax over the first
    num_cmp(t) compressed keys, zero when there are none.  Mask = the one the reference builds
    and discards at attention_kernels.py:117-126; lengths = packing.py:15-23."""
    B, S, G = Q.shape[:3]
    S_cmp = K_cmp.shape[2]
    nc = torch.tensor([num_cmp_at(t0 + t, l, d, S_cmp) for t in range(S)])
    allowed = torch.arange(S_cmp)[None, :] < nc[:, None]
    return _masked_attention(Q, K_cmp, V_cmp, allowed[None, :, None, :].expand(B, S, G, S_cmp))


def literal_first_key(V: torch.Tensor, first_idx: torch.Tensor, has_any: torch.Tensor, h: int) -> torch.Tensor:
    """Literal-mode shim: what the reference's default routes return (V at the first allowed
    key).  V [B,G,Skv,Dv], first_idx/has_any [B,S,G] -> [B,S,G,h,Dv]."""
    B, S, G = first_idx.shape
    idx = first_idx.permute(0, 2, 1)[..., None].expand(B, G, S, V.shape[-1]).long()
    v = torch.gather(V, 2, idx).permute(0, 2, 1, 3)  # [B,S,G,Dv]
    v = torch.where(has_any[..., None], v, torch.zeros_like(v))
    return v[:, :, :, None, :].expand(B, S, G, h, V.shape[-1])


# ----------------------------------------------------------------------------------------
# Gate MLP and combine                     (nsa/core/nsa_attention.py:32-82, :1356-1398)
# ----------------------------------------------------------------------------------------
GATE_MLP, GATE_UNIFORM, GATE_CMP, GATE_SEL, GATE_WIN = 0, 1, 2, 3, 4


def gate_mlp(q_gp: torch.Tensor, fc1_w, fc1_b, fc2_w, fc2_b, tau: float = 1.0, mode: int = GATE_MLP) -> torch.Tensor:
    """softmax(fc2(silu(fc1(q_gp))) / max(tau,1e-6)); rows whose top-2 logit gap exceeds 50
    become hard one-hot (nsa_attention.py:70-82); NSA_FORCE_UNIFORM_GATE / NSA_FORCE_BRANCH
    overrides (:51-68)."""
    shape = (*q_gp.shape[:-1], 3)
    if mode == GATE_UNIFORM:
        return torch.full(shape, 1.0 / 3.0, dtype=q_gp.dtype)
    if mode in (GATE_CMP, GATE_SEL, GATE_WIN):
        one = torch.zeros(shape, dtype=q_gp.dtype)
        one[..., mode - GATE_CMP] = 1.0
        return one
    x = torch.nn.functional.silu(torch.nn.functional.linear(q_gp, fc1_w, fc1_b))
    g = torch.nn.functional.linear(x, fc2_w, fc2_b) / max(tau, 1e-6)
    p = torch.softmax(g, dim=-1)
    with torch.no_grad():
        top2 = torch.topk(g, k=2, dim=-1).values
        peaked = (top2[..., 0] - top2[..., 1]) > 50.0
    hard = torch.nn.functional.one_hot(torch.argmax(g, dim=-1), 3).to(p.dtype)
    return torch.where(peaked[..., None], hard, p)


def combine(gates: torch.Tensor, O_cmp, O_sel, O_win) -> torch.Tensor:
    """O = g_cmp*O_cmp + g_sel*O_sel + g_win*O_win; gate order [cmp, sel, win]
    (nsa_attention.py:1393-1396)."""
    g = gates[..., None, :]  # [B,S,G,1,3]
    return g[..., 0:1] * O_cmp + g[..., 1:2] * O_sel + g[..., 2:3] * O_win


# ----------------------------------------------------------------------------------------
# End-to-end hot path
# ----------------------------------------------------------------------------------------
def prefill_scores(Q, K_cmp, l, d, l_sel, n_sel, w, norm="full_row", t0=0, S_total=None):
    """nsa_attention.py:1073-1091: p_cmp -> p_slc -> p_grp for every row (rows t0 .. t0+S-1 of S_total)."""
    S = Q.shape[1]
    meta = build_meta(S_total if S_total is not None else t0 + S, l, d, l_sel, n_sel, w)
    scale = 1.0 / math.sqrt(Q.shape[-1])
    if K_cmp.shape[2] == 0:
        return torch.zeros(Q.shape[0], S, Q.shape[2], meta.S_sel)
    p = pcmp_all(Q, K_cmp, scale, norm, l, d, t0)
    return pgrp_from_pslc(pslc_from_pcmp(p, meta))


def prefill_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate_params, *, l, d, l_sel, n_sel, w,
                 tau=1.0, gate_mode=GATE_MLP, norm="full_row", ranges=None, t0=0, S_total=None):
    """Hot path of _forward_prefill_batched (nsa_attention.py:1066-1398) between "Q/K/V
    projected + RoPE'd" and "O handed to self.out", intended semantics.  Returns dict."""
    S = Q.shape[1]
    S_total = t0 + S if S_total is None else S_total
    if ranges is None:
        p_grp = prefill_scores(Q, K_cmp, l, d, l_sel, n_sel, w, norm, t0, S_total)
        ranges = select_ranges_prefill(p_grp, l_sel, n_sel, S_total, t0)
    else:
        p_grp = None
    O_cmp, lse_cmp = cmp_attention(Q, K_cmp, V_cmp, l, d, t0)
    O_sel, lse_sel = sel_attention(Q, K_sel, V_sel, ranges)
    O_win, lse_win = win_attention(Q, K_win, V_win, w, t0)
    q_gp = Q.float().mean(dim=3) if Q.dtype != torch.float64 else Q.mean(dim=3)
    gates = gate_mlp(q_gp, *gate_params, tau=tau, mode=gate_mode)
    O = combine(gates.to(O_cmp.dtype), O_cmp, O_sel, O_win)
    return dict(O=O, O_cmp=O_cmp, O_sel=O_sel, O_win=O_win, gates=gates, ranges=ranges,
                p_grp=p_grp, lse=torch.stack([lse_cmp, lse_sel, lse_win], dim=-1))


def decode_core(q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate_params, *, l, d, l_sel, n_sel, w,
                tau=1.0, gate_mode=GATE_MLP):
    """Hot path of one decode step (nsa_attention.py:648-971) after the caches were appended:
    q [B,G,h,Dk]; K_sel holds t+1 tokens; K_win the last min(w,t+1); K_cmp the emitted blocks.
    Scores use only emitted blocks (causal by construction); selection = select_topn_ranges."""
    B, G, h, Dk = q.shape
    t = K_sel.shape[2] - 1
    Q = q[:, None]  # [B,1,G,h,Dk]
    S_sel = num_sel_blocks(max(t + 1, l_sel), l_sel)
    meta = build_meta(max(t + 1, l_sel), l, d, l_sel, n_sel, w)
    if K_cmp.shape[2] > 0:
        p = pcmp_all(Q, K_cmp, 1.0 / math.sqrt(Dk))
        p_grp = pgrp_from_pslc(pslc_from_pcmp(p, meta))[:, 0]
    else:
        p_grp = torch.zeros(B, G, S_sel)
    ranges = select_ranges_decode(p_grp, l_sel, n_sel, t)
    allowed = torch.ones(B, 1, G, K_cmp.shape[2], dtype=torch.bool)
    O_cmp, lse_cmp = _masked_attention(Q, K_cmp, V_cmp, allowed) if K_cmp.shape[2] > 0 else (
        torch.zeros(B, 1, G, h, V_cmp.shape[-1]), torch.full((B, 1, G, h), NEG_INF))
    O_sel, lse_sel = sel_attention(Q, K_sel, V_sel, ranges[:, None])
    aw = torch.ones(B, 1, G, K_win.shape[2], dtype=torch.bool)
    O_win, lse_win = _masked_attention(Q, K_win, V_win, aw)
    gates = gate_mlp(Q.float().mean(dim=3), *gate_params, tau=tau, mode=gate_mode)
    O = combine(gates, O_cmp, O_sel, O_win)
    return dict(O=O[:, 0], O_cmp=O_cmp[:, 0], O_sel=O_sel[:, 0], O_win=O_win[:, 0], gates=gates[:, 0],
                ranges=ranges, p_grp=p_grp)


def decode_emits(S_raw: int, l: int, d: int) -> bool:
    """A compressed token is emitted after appending raw token number S_raw (1-based count)
    iff S_raw >= l and (S_raw - l) % d == 0  (nsa_attention.py:587-588)."""
    return S_raw >= l and (S_raw - l) % d == 0


def expected_reads(S: int, l: int, d: int, n_sel: int, l_sel: int, w: int) -> Tuple[int, int, int, int]:
    """Decode read counters (nsa_attention.py:634-638; bench_decode.py:36-38)."""
    nc = num_cmp_blocks(S, l, d)
    return nc + n_sel * l_sel + min(w, S), n_sel * l_sel, nc, min(w, S)
