"""Free functions of the reference's nsa/core/selection_scorer.py, same names and argument meaning, backed by
the CUDA kernels.  `meta` arguments are accepted for signature compatibility; the kernels evaluate the Eq.9
stencil from (l, d, l_sel) directly and only read the block sizes from it.
"""
from __future__ import annotations

import math

import torch

from .. import ops
from .block_index import BlockMeta


def _cfg(meta: BlockMeta, n_sel=None, norm=ops.NORM_FULL_ROW) -> ops.NSAConfig:
    return ops.NSAConfig(l=meta.l, d=meta.d, l_sel=meta.l_sel, n_sel=meta.n_sel if n_sel is None else n_sel, w=meta.w,
                         norm_mode=norm)


def compute_pgrp_all(Q_all: torch.Tensor, K_cmp: torch.Tensor, meta: BlockMeta, scale: float | None = None) -> torch.Tensor:
    """Fused compute_pcmp_all -> map_pcmp_to_pslc_batched -> sum over heads (selection_scorer.py:42-61, :89-116,
    nsa_attention.py:1091): Q [B,S,G,h,Dk], K_cmp [B,G,S_cmp,Dk] -> p_grp [B,S,G,S_sel] fp32.  The [.., S_cmp]
    probabilities the reference materialises (12.9 GB per sequence at 64k) never exist here."""
    if scale is not None and abs(scale - 1.0 / math.sqrt(Q_all.shape[-1])) > 1e-12:
        raise RuntimeError("the CUDA scorer uses scale = 1/sqrt(Dk)")
    return ops.score_pgrp(Q_all, K_cmp, _cfg(meta), S_sel=int(meta.sel_starts.numel()))


def group_reduce_pslc(p_slc: torch.Tensor) -> torch.Tensor:
    """Eq.10 (selection_scorer.py:119-121); trivial, kept for API compatibility."""
    return p_slc.sum(dim=2)


def select_topn_ranges(p_grp: torch.Tensor, meta: BlockMeta, n_top: int, t_token: int, force_init: bool = True,
                       force_local: int = 2, _skip_validation: bool = False) -> torch.Tensor:
    """selection_scorer.py:124-249 -> [B,G,n_top,2] int32."""
    if not force_init or force_local != 2:
        raise RuntimeError("the CUDA selector implements force_init=True, force_local=2 (the only call sites)")
    return ops.select_ranges_decode(p_grp, meta.l_sel, n_top, int(t_token))


def select_topn_ranges_batched(p_grp_all: torch.Tensor, meta: BlockMeta, n_top: int, S: int, force_init: bool = True,
                               force_local: int = 2) -> torch.Tensor:
    """selection_scorer.py:255-362 (with ranges v2, :434-605) -> [B,S,G,K,2] int32."""
    if not force_init or force_local != 2:
        raise RuntimeError("the CUDA selector implements force_init=True, force_local=2 (the only call sites)")
    return ops.select_ranges_prefill(p_grp_all, meta.l_sel, n_top, S_total=S)
