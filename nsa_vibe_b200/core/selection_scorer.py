"""Free functions of the reference's nsa/core/selection_scorer.py, same names and argument meaning, backed by
the CUDA kernels.  `meta` arguments are accepted for signature compatibility; the kernels evaluate the Eq.9
stencil from (l, d, l_sel) directly and only read the block sizes from it.

The hot path (NSAAttention.forward) does not call these one by one -- it runs scoring, Eq.9, Eq.10 and the selection
fused (compute_pgrp_all / ops.score_select) and never materialises p_cmp or p_slc.  The stage-by-stage functions exist
for the callers and tests of the reference that import them.
"""
from __future__ import annotations

import math
import os

import torch

from .. import ops
from .block_index import BlockMeta


def _cfg(meta: BlockMeta, n_sel=None, norm=ops.NORM_FULL_ROW) -> ops.NSAConfig:
    return ops.NSAConfig(l=meta.l, d=meta.d, l_sel=meta.l_sel, n_sel=meta.n_sel if n_sel is None else n_sel, w=meta.w,
                         norm_mode=norm)


def compute_pcmp_all(Q_all: torch.Tensor, K_cmp: torch.Tensor, scale: float) -> torch.Tensor:
    """selection_scorer.py:42-61: Q [B,S,G,h,Dk], K_cmp [B,G,S_cmp,Dk] -> p_cmp [B,S,G,h,S_cmp], softmax over ALL S_cmp keys, in
    the dtype of Q.  Stand-alone stage (S_cmp up to ~12k); the hot path uses compute_pgrp_all."""
    if abs(scale - 1.0 / math.sqrt(Q_all.shape[-1])) > 1e-9 * max(1.0, abs(scale)):
        raise RuntimeError("the CUDA scorer uses scale = 1/sqrt(Dk)")
    return ops.pcmp_all(Q_all, K_cmp).to(Q_all.dtype)


def compute_pgrp_all(Q_all: torch.Tensor, K_cmp: torch.Tensor, meta: BlockMeta, scale: float | None = None) -> torch.Tensor:
    """Fused compute_pcmp_all -> map_pcmp_to_pslc_batched -> sum over heads (selection_scorer.py:42-61, :89-116,
    nsa_attention.py:1091): Q [B,S,G,h,Dk], K_cmp [B,G,S_cmp,Dk] -> p_grp [B,S,G,S_sel] fp32.  The [.., S_cmp]
    probabilities the reference materialises (12.9 GB per sequence at 64k) never exist here."""
    if scale is not None and abs(scale - 1.0 / math.sqrt(Q_all.shape[-1])) > 1e-12:
        raise RuntimeError("the CUDA scorer uses scale = 1/sqrt(Dk)")
    return ops.score_pgrp(Q_all, K_cmp, _cfg(meta), S_sel=int(meta.sel_starts.numel()))


def map_pcmp_to_pslc(p_cmp: torch.Tensor, meta: BlockMeta) -> torch.Tensor:
    """selection_scorer.py:64-86: p_cmp [B,G,h,S_cmp] -> p_slc [B,G,h,S_sel] (Eq.9; rows beyond meta's S_cmp are ignored)."""
    S_cmp = min(int(p_cmp.shape[-1]), int(meta.cmp_starts.numel()))
    out = ops.map_pcmp_to_pslc(p_cmp[..., :S_cmp], int(meta.sel_starts.numel()), meta.l, meta.d, meta.l_sel)
    return out.to(p_cmp.dtype)


def map_pcmp_to_pslc_batched(p_cmp_all: torch.Tensor, meta: BlockMeta) -> torch.Tensor:
    """selection_scorer.py:89-116: p_cmp_all [B,S,G,h,S_cmp] -> [B,S,G,h,S_sel]."""
    return map_pcmp_to_pslc(p_cmp_all, meta)


def group_reduce_pslc(p_slc: torch.Tensor) -> torch.Tensor:
    """Eq.10 (selection_scorer.py:119-121); trivial, kept for API compatibility."""
    return p_slc.sum(dim=2)


def select_topn_ranges(p_grp: torch.Tensor, meta: BlockMeta, n_top: int, t_token: int, force_init: bool = True,
                       force_local: int = 2, _skip_validation: bool = False) -> torch.Tensor:
    """selection_scorer.py:124-249 -> [B,G,n_top,2] int32.  Ranking happens in fp32 whatever the dtype of p_grp (:182-184)."""
    out = ops.select_ranges_decode(p_grp, meta.l_sel, n_top, int(t_token), force_init=force_init, force_local=force_local)
    if not _skip_validation and os.getenv("NSA_VALIDATE_SELECTION_DETERMINISM", "0").lower() in ("1", "true", "yes"):
        validate_selection_determinism(p_grp, meta, n_top, t_token)
    return out


def select_topn_ranges_batched(p_grp_all: torch.Tensor, meta: BlockMeta, n_top: int, S: int, force_init: bool = True,
                               force_local: int = 2) -> torch.Tensor:
    """selection_scorer.py:255-362 (with ranges v2, :434-605) -> [B,S,G,K,2] int32."""
    return ops.select_ranges_prefill(p_grp_all, meta.l_sel, n_top, S_total=S, force_init=force_init, force_local=force_local)


def convert_indices_to_ranges_batched_v2(indices: torch.Tensor, meta: BlockMeta, S: int) -> torch.Tensor:
    """selection_scorer.py:434-605: indices [B,S,G,K] (ascending, -1 padded) -> ranges [B,S,G,K,2] int32, [0,0] padded."""
    return ops.indices_to_ranges(indices, int(meta.sel_starts.numel()), meta.l_sel)


def convert_indices_to_ranges_batched(indices: torch.Tensor, meta: BlockMeta, S: int) -> torch.Tensor:
    """selection_scorer.py:380-431: as v2, but trimmed to the widest row ([B,S,G,max_ranges,2]) like the reference's loop.  The
    trim reads one integer back from the device."""
    r = convert_indices_to_ranges_batched_v2(indices, meta, S)
    if r.numel() == 0:
        return r[..., :0, :] if r.shape[3] else r
    n = int(((r[..., 1] > r[..., 0]).sum(dim=-1)).max().item())
    return r[..., :n, :].contiguous()


def convert_indices_to_ranges_batched_dispatch(indices: torch.Tensor, meta: BlockMeta, S: int) -> torch.Tensor:
    """selection_scorer.py:364-377."""
    if os.getenv("NSA_SEL_RANGES_V2", "1").lower() in ("1", "true", "yes"):
        return convert_indices_to_ranges_batched_v2(indices, meta, S)
    return convert_indices_to_ranges_batched(indices, meta, S)


def validate_selection_determinism(p_grp: torch.Tensor, meta: BlockMeta, n_top: int, t_token: int, num_trials: int = 3) -> bool:
    """M8 helper of the reference: the same inputs select the same ranges on every trial."""
    first = select_topn_ranges(p_grp, meta, n_top, t_token, _skip_validation=True)
    for _ in range(max(0, num_trials - 1)):
        if not torch.equal(first, select_topn_ranges(p_grp, meta, n_top, t_token, _skip_validation=True)):
            return False
    return True
