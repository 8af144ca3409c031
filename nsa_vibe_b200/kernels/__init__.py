"""Names the reference exposes under nsa/kernels (flash_wrappers.attention_bgh, cuda_sel_kernel.selection_attention_cuda,
triton_sel_kernel.selection_attention_triton), all backed by the sm_100a kernels."""
from __future__ import annotations

import torch

from .. import ops
from ..core.attention_kernels import grouped_selection_attention as selection_attention_cuda  # noqa: F401
from ..core.attention_kernels import grouped_selection_attention as selection_attention_triton  # noqa: F401


def attention_bgh(Q: torch.Tensor, K: torch.Tensor, V: torch.Tensor, causal: bool = True) -> torch.Tensor:
    """One-query GQA attention (nsa/kernels/flash_wrappers.py:191-282): Q [B,G,h,Dk], K/V [B,G,S,D*] -> [B,G,h,Dv],
    softmax over all S keys.  (`causal` is accepted and ignored: with one query row the reference's is_causal=True
    degenerates to key 0; the intended semantics is every key <= t, i.e. all of K.)"""
    S = K.shape[2]
    cfg = ops.NSAConfig(w=max(S, 1))
    O = ops.branch_attention(ops.BR_WIN, Q[:, None], K, V, cfg, t0=S - 1)
    return O[:, 0]
