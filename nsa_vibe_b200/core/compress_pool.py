"""Torch restatement of phi = avg-pool(kernel l, stride d) over RoPE(K) and raw V (nsa/core/compress_pool.py:9-38).  The product path
pools inside `ops.phi_avgpool` (one CUDA pass per tensor, forward and backward); this function is the comparator that kernel is tested
against and keeps the reference's public name."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F

from .rope import apply_rope


def avg_pool_phi_rope_kv(K_raw: torch.Tensor, V_raw: torch.Tensor, l: int, d: int,
                         pos: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    B, G, S, Dk = K_raw.shape
    Dv = V_raw.shape[-1]
    if pos is None:
        pos = torch.arange(S, device=K_raw.device)
    K_rope = apply_rope(K_raw, pos)
    if S < l:
        return (K_raw.new_zeros((B, G, 0, Dk)), V_raw.new_zeros((B, G, 0, Dv)))
    Kp = F.avg_pool1d(K_rope.reshape(B * G, S, Dk).transpose(1, 2), kernel_size=l, stride=d)
    Vp = F.avg_pool1d(V_raw.reshape(B * G, S, Dv).transpose(1, 2), kernel_size=l, stride=d)
    S_cmp = Kp.shape[-1]
    return (Kp.transpose(1, 2).reshape(B, G, S_cmp, Dk).contiguous(), Vp.transpose(1, 2).reshape(B, G, S_cmp, Dv).contiguous())
