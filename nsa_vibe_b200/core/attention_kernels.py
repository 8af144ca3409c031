"""Branch attention entry points with the reference's names (nsa/core/attention_kernels.py), all routed to the
one CUDA path with true-softmax semantics and an analytical backward."""
from __future__ import annotations

import torch

from .. import ops


def _cfg(l=32, d=16, l_sel=64, n_sel=16, w=512) -> ops.NSAConfig:
    return ops.NSAConfig(l=l, d=d, l_sel=l_sel, n_sel=n_sel, w=w)


def grouped_selection_attention(Q, K, V, ranges):
    """Q [B,S,G,h,Dk], K/V [B,G,S_kv,D*], ranges [B,S,G,n,2] -> [B,S,G,h,Dv]; softmax over every token in the
    ranges, empty rows -> 0 (the semantics of grouped_selection_attention_masked, attention_kernels.py:705-772)."""
    return ops.branch_attention(ops.BR_SEL, Q, K, V, _cfg(), ranges)


grouped_selection_attention_masked = grouped_selection_attention
grouped_selection_attention_packed = grouped_selection_attention
selection_attention_varlen_all = grouped_selection_attention


def sliding_window_attention(Q, K, V, w: int):
    """attention_kernels.py:146-178: row t attends to keys [t-w+1 .. t]."""
    B, S, G, h, _ = Q.shape
    if w <= 0 or K.shape[2] == 0 or S == 0:
        return torch.zeros((B, S, G, h, V.shape[-1]), dtype=V.dtype, device=V.device)
    return ops.branch_attention(ops.BR_WIN, Q, K, V, _cfg(w=w))


def batched_causal_attention_compressed(Q, K_cmp, V_cmp, l: int, d: int):
    """attention_kernels.py:106-143 with the intended semantics: softmax over the first num_cmp(t) compressed
    tokens, zero where there are none."""
    B, S, G, h, _ = Q.shape
    if K_cmp.shape[2] == 0:
        return torch.zeros((B, S, G, h, V_cmp.shape[-1]), dtype=V_cmp.dtype, device=V_cmp.device)
    return ops.branch_attention(ops.BR_CMP, Q, K_cmp, V_cmp, _cfg(l=l, d=d, l_sel=d))
