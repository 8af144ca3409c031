"""Rotary embedding as the reference applies it (nsa/core/rope.py:16-51): interleaved pairs, fp32 angles,
sin/cos cast to the input dtype.  Producer of the hot path's inputs, kept in torch."""
from __future__ import annotations

import torch


def build_inv_freq(dim: int, base: float = 10000.0, device=None) -> torch.Tensor:
    assert dim % 2 == 0, "RoPE requires even dimension"
    idx = torch.arange(dim // 2, device=device, dtype=torch.float32)
    return base ** (-2 * idx / dim)


def apply_rope(x: torch.Tensor, pos: torch.Tensor, base: float = 10000.0, *, scale: float = 1.0) -> torch.Tensor:
    D = x.shape[-1]
    assert D % 2 == 0, "RoPE requires even dimension"
    inv_freq = build_inv_freq(D, base=base, device=x.device)
    while pos.dim() < x.dim() - 1:
        pos = pos.unsqueeze(0)
    if scale <= 0:
        scale = 1.0
    ang = (pos.to(torch.float32) / float(scale)).unsqueeze(-1) * inv_freq
    sin, cos = torch.sin(ang).to(x.dtype), torch.cos(ang).to(x.dtype)
    x2 = x.reshape(*x.shape[:-1], D // 2, 2)
    x0, x1 = x2[..., 0], x2[..., 1]
    return torch.stack((x0 * cos - x1 * sin, x0 * sin + x1 * cos), dim=-1).reshape(x.shape)
