"""Torch restatement of the rotary embedding the reference applies (nsa/core/rope.py:16-51: interleaved pairs, angles in fp32,
sin / cos rounded to the input dtype, every product rounded to that dtype).  The product path rotates inside the CUDA producers
(`ops.project_split`, `ops.rope_shape`, `ops.phi_avgpool`); this module is the comparator those kernels are tested against and
keeps the reference's two public names."""
from __future__ import annotations

import torch


def build_inv_freq(dim: int, base: float = 10000.0, device=None) -> torch.Tensor:
    """base^(-2i/dim) for the dim/2 rotation pairs, fp32."""
    if dim % 2:
        raise AssertionError("RoPE requires even dimension")
    pair = torch.arange(dim // 2, dtype=torch.float32, device=device)
    return base ** (-2 * pair / dim)


def apply_rope(x: torch.Tensor, pos: torch.Tensor, base: float = 10000.0, *, scale: float = 1.0) -> torch.Tensor:
    """x [..., S, D] (D even), pos [S] or [..., S] -> x rotated pair-wise: (x_2i, x_2i+1) by the angle (pos / scale) * base^(-2i/D)."""
    width = x.shape[-1]
    freq = build_inv_freq(width, base=base, device=x.device)
    div = float(scale) if scale > 0 else 1.0
    p = pos.to(torch.float32)
    p = p.reshape((1,) * (x.dim() - 1 - p.dim()) + tuple(p.shape)) if p.dim() < x.dim() - 1 else p
    theta = (p / div)[..., None] * freq                      # fp32 [..., S, D/2]
    s, c = theta.sin().to(x.dtype), theta.cos().to(x.dtype)
    even, odd = x[..., 0::2], x[..., 1::2]
    out = torch.empty_like(x)
    out[..., 0::2] = even * c - odd * s
    out[..., 1::2] = even * s + odd * c
    return out
