"""LlamaBlockNSA (nsa/model/llama_block_nsa.py:33-106): RMSNorm -> NSAAttention -> residual -> RMSNorm -> SiLU MLP.
The caller of the hot path: nn.Linear GEMMs around the B200 NSAAttention, RMSNorm (+ residual add) as one kernel per direction."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..cache.kv_cache import create_empty_kv
from ..core.block_index import build_block_meta
from ..core.nsa_attention import NSAAttention


def rmsnorm_torch(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """The reference's formula (llama_block_nsa.py:19-22), kept as the comparator of the kernel in tests."""
    rms = x.pow(2).mean(dim=-1, keepdim=True).add(eps).rsqrt()
    return (x * rms) * weight


def _norm_out_dtype(x: torch.Tensor) -> torch.dtype:
    # under autocast every consumer of a norm output is an nn.Linear that would cast it: emit that dtype directly
    if torch.is_autocast_enabled("cuda") and x.dtype == torch.float32:
        return torch.get_autocast_dtype("cuda")
    return x.dtype


class RMSNorm(nn.Module):
    """llama_block_nsa.py:13-22; one kernel per direction (ops.rmsnorm).  forward(x, residual=r) returns (x + r, norm(x + r))."""

    def __init__(self, dim: int, eps: float = 1e-6) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x: torch.Tensor, residual: torch.Tensor = None):
        return ops.rmsnorm(x, self.weight, self.eps, residual=residual, out_dtype=_norm_out_dtype(x))


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

    def forward(self, x: torch.Tensor, delta: torch.Tensor = None, defer_residual: bool = False):
        """x -> x + attn(norm1(x)) -> (+ mlp(norm2(.))).  `delta`: a pending residual of the previous block (x stands for
        x + delta; the add happens inside norm1's kernel).  defer_residual=True returns (x, mlp_out) and leaves the last add to
        the next norm -- a stack of blocks then never runs a stand-alone residual add (llama_block_nsa.py:102-106)."""
        B, S, _ = x.shape
        if delta is not None:
            x, xn = self.norm1(x, residual=delta)
        else:
            xn = self.norm1(x)
        a = self.attn
        meta = build_block_meta(S, a.l, a.d, a.l_sel, a.n_sel, a.w)
        kv = create_empty_kv(B, a.n_kv_groups, a.d_k, a.d_v, meta, device=x.device, dtype=xn.dtype)
        out, _ = a(xn, kv, prefill=True)
        x, xn2 = self.norm2(x, residual=out)  # x = x + out and norm2(x) in one pass
        m = self.mlp(xn2)
        return (x, m) if defer_residual else x + m
