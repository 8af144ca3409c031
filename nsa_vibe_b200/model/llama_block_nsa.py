"""LlamaBlockNSA (nsa/model/llama_block_nsa.py:33-106): RMSNorm -> NSAAttention -> residual -> RMSNorm -> SiLU MLP.
The caller of the hot path; plain torch around the B200 NSAAttention."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..cache.kv_cache import create_empty_kv
from ..core.block_index import build_block_meta
from ..core.nsa_attention import NSAAttention


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        rms = x.pow(2).mean(dim=-1, keepdim=True).add(self.eps).rsqrt()
        return (x * rms) * self.weight


class MLP(nn.Module):
    def __init__(self, dim: int, hidden_mult: int = 4) -> None:
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden_mult * dim, bias=False)
        self.fc2 = nn.Linear(hidden_mult * dim, dim, bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc2(F.silu(self.fc1(x)))


class LlamaBlockNSA(nn.Module):
    def __init__(self, dim: int, n_heads: int, n_kv_groups: int, d_k: int, d_v: int, l: int = 32, d: int = 16,
                 l_sel: int = 64, n_sel: int = 16, w: int = 512) -> None:
        super().__init__()
        self.norm1 = RMSNorm(dim)
        self.attn = NSAAttention(dim=dim, n_heads=n_heads, n_kv_groups=n_kv_groups, d_k=d_k, d_v=d_v, l=l, d=d,
                                 l_sel=l_sel, n_sel=n_sel, w=w)
        self.norm2 = RMSNorm(dim)
        self.mlp = MLP(dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, S, _ = x.shape
        xn = self.norm1(x)
        a = self.attn
        meta = build_block_meta(S, a.l, a.d, a.l_sel, a.n_sel, a.w)
        kv = create_empty_kv(B, a.n_kv_groups, a.d_k, a.d_v, meta, device=x.device, dtype=xn.dtype)
        out, _ = a(xn, kv, prefill=True)
        x = x + out
        return x + self.mlp(self.norm2(x))
