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


def selection_attention_backward_reference(Q, K, V, ranges, dO):
    """_selection_attention_backward (nsa/kernels/triton_sel_kernel/__init__.py:163-231) under the intended semantics (softmax over
    every selected key; the reference keeps key 0 only, :217-219): Q [B,S,G,h,Dk], K/V [B,G,S_kv,D*], ranges [B,S,G,n,2], dO
    [B,S,G,h,Dv] -> (dQ, dK, dV) through the analytical backward kernels."""
    Qr, Kr, Vr = (t.detach().clone().requires_grad_(True) for t in (Q, K, V))
    O = ops.branch_attention(ops.BR_SEL, Qr, Kr, Vr, ops.NSAConfig(), ranges)
    return torch.autograd.grad(O, (Qr, Kr, Vr), dO.to(O.dtype))
