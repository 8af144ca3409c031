"""Length helpers (nsa/core/packing.py:6-23)."""
import torch


def compute_sliding_lengths(S: int, w: int, device) -> torch.Tensor:
    return (torch.arange(S, device=device) + 1).clamp_max(w)


def compute_compressed_lengths(S: int, l: int, d: int, S_cmp: int, device) -> torch.Tensor:
    t = torch.arange(S, device=device)
    return torch.where(t + 1 < l, 0, ((t + 1 - l) // d) + 1).clamp(min=0, max=S_cmp)
