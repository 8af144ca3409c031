"""TinyLM (scripts/train_showcase.py:30-117 of the reference): embedding -> LlamaBlockNSA stack -> RMSNorm -> lm_head.  The caller
of the hot path in the DDP training config; each block's last residual add runs inside the next norm's kernel."""
from __future__ import annotations

import torch
import torch.nn as nn

from .llama_block_nsa import LlamaBlockNSA, RMSNorm


class TinyLM(nn.Module):
    def __init__(self, vocab: int, dim: int, n_layers: int, heads: int, groups: int, dk: int, dv: int, l: int, d: int, l_sel: int,
                 n_sel: int, w: int):
        super().__init__()
        self.embed = nn.Embedding(vocab, dim)
        self.blocks = nn.ModuleList([LlamaBlockNSA(dim, heads, groups, dk, dv, l, d, l_sel, n_sel, w) for _ in range(n_layers)])
        self.norm_f = RMSNorm(dim)
        self.lm_head = nn.Linear(dim, vocab, bias=False)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        x, delta = self.embed(ids), None
        for b in self.blocks:
            x, delta = b(x, delta, defer_residual=True)
        _, xn = self.norm_f(x, residual=delta)
        return self.lm_head(xn)

    def grad_buckets(self):
        """Parameters grouped in the order their gradients become final during backward: head, then the blocks last to first, then
        the embedding -- one bucket per block (6.49 M parameters = 13 MB of bf16 at m7c dims)."""
        buckets = [list(self.lm_head.parameters()) + list(self.norm_f.parameters())]
        buckets += [list(b.parameters()) for b in reversed(self.blocks)]
        buckets.append(list(self.embed.parameters()))
        return [b for b in buckets if b]


def m7c_tiny_lm(n_layers: int = 12) -> TinyLM:
    """configs/m7c_125m_80g.yaml shapes with the byte vocabulary the synthetic-data runs use (78.3 M parameters at 12 layers)."""
    return TinyLM(256, 768, n_layers, 12, 2, 64, 64, 32, 16, 64, 16, 512)
