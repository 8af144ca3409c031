"""Tensor-level wrappers over the C ABI (include/nsa_b200.h).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every op below launches hand-written
sm_100a kernels through libnsa_b200.so on torch's current stream.  CPU tensors are rejected -- there is no
fallback (BASELINE.json north_star).
"""
from __future__ import annotations

import collections
import ctypes as C
import math
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (GATE_CMP, GATE_MLP, GATE_SEL, GATE_UNIFORM, GATE_WIN, IMPL_AUTO, IMPL_SIMT, IMPL_TC,  # noqa: F401
                   NORM_CAUSAL, NORM_FULL_ROW)

_DTYPES = {torch.float32: _lib.NSA_F32, torch.bfloat16: _lib.NSA_BF16, torch.float16: _lib.NSA_F16}

# how many kernels this process launched through the C ABI (bench.py reports it as gpu_launches)
launch_count = 0


@dataclass
class NSAConfig:
    """Static NSA geometry + switches for one module (nsa_attention.py:188-206 ctor arguments)."""
    l: int = 32
    d: int = 16
    l_sel: int = 64
    n_sel: int = 16
    w: int = 512
    gate_tau: float = 1.0
    gate_mode: int = GATE_MLP
    norm_mode: int = NORM_FULL_ROW
    impl: int = IMPL_AUTO


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    # raw handle of torch's current stream on the current device (torch.cuda.current_stream() costs ~10 us of Python per call)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("nsa_vibe_b200 ops need CUDA tensors: the NSA hot path has no CPU fallback")


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _gate_struct(gate, dev):
    """gate = (fc1_w, fc1_b, fc2_w, fc2_b) tensors (any float dtype) or None -> (GateParams, keepalive, hidden)."""
    gp = _lib.GateParams()
    if gate is None:
        return gp, (), 0
    keep = tuple(None if t is None else _c(t.detach().to(device=dev, dtype=torch.float32)) for t in gate)
    gp.fc1_w, gp.fc1_b, gp.fc2_w, gp.fc2_b = (None if t is None else t.data_ptr() for t in keep)
    return gp, keep, int(keep[0].shape[0])


def make_dims(Q: torch.Tensor, cfg: NSAConfig, *, t0: int = 0, K_sel=None, K_win=None, K_cmp=None, V=None,
              S_sel_kv=None, S_win_kv=None, S_cmp=None, win_off: int = 0, n_ranges: int = 0, gate_hidden: int = 0,
              Dv: Optional[int] = None) -> _lib.Dims:
    B, S, G, h, Dk = Q.shape
    dm = _lib.Dims()
    dm.B, dm.S, dm.G, dm.h, dm.Dk = B, S, G, h, Dk
    dm.Dv = int(Dv if Dv is not None else (V.shape[-1] if V is not None else Dk))
    dm.l, dm.d, dm.l_sel, dm.n_sel, dm.w = cfg.l, cfg.d, cfg.l_sel, cfg.n_sel, cfg.w
    dm.t0 = int(t0)
    dm.cap_sel = int(K_sel.shape[2]) if K_sel is not None else 0
    dm.S_sel_kv = dm.cap_sel if S_sel_kv is None else int(S_sel_kv)
    dm.cap_win = int(K_win.shape[2]) if K_win is not None else 0
    dm.S_win_kv = dm.cap_win if S_win_kv is None else int(S_win_kv)
    dm.win_off = int(win_off)
    dm.cap_cmp = int(K_cmp.shape[2]) if K_cmp is not None else 0
    dm.S_cmp = dm.cap_cmp if S_cmp is None else int(S_cmp)
    dm.n_ranges = int(n_ranges)
    dm.dtype = _DTYPES[Q.dtype]
    dm.gate_mode, dm.gate_hidden, dm.norm_mode = int(cfg.gate_mode), int(gate_hidden), int(cfg.norm_mode)
    dm.impl = int(cfg.impl) if int(cfg.impl) == IMPL_SIMT else int(os.environ.get("NSA_B200_IMPL", cfg.impl))
    dm.gate_tau = float(cfg.gate_tau)
    dm.scale = 1.0 / math.sqrt(Dk)
    return dm


def _call(name: str, *args):
    global launch_count
    lib = _lib.load()
    rc = getattr(lib, name)(*args)
    _lib.check(rc, name)
    launch_count += 1


def _workspace(dm, which: int, dev) -> Optional[torch.Tensor]:
    """Scratch the kernels of one call need (nsa_workspace_bytes); None when the call needs none."""
    n = int(_lib.load().nsa_workspace_bytes(C.byref(dm), which))
    return torch.empty(n, dtype=torch.uint8, device=dev) if n else None


def prefill_range_cols(S_total: int, l_sel: int, n_sel: int, force_init: bool = True, force_local: int = 2) -> int:
    return int(_lib.load().nsa_prefill_range_cols_ex(S_total, l_sel, n_sel, int(bool(force_init)), int(force_local)))


def num_sel_blocks(seq_len: int, l_sel: int) -> int:
    return 0 if seq_len <= 0 else (seq_len + l_sel - 1) // l_sel


# ----------------------------------------------------------------------------------------------------
# (2) selection
# ----------------------------------------------------------------------------------------------------
def select_ranges_prefill(p_grp: torch.Tensor, l_sel: int, n_sel: int, S_total: Optional[int] = None, t0: int = 0, *,
                          force_init: bool = True, force_local: int = 2) -> torch.Tensor:
    """p_grp [B,S,G,S_sel] fp32 -> ranges [B,S,G,K,2] int32; bit-exact vs select_topn_ranges_batched
    (nsa/core/selection_scorer.py:255-362)."""
    _require_cuda(p_grp)
    B, S, G, S_sel = p_grp.shape
    S_total = S if S_total is None else S_total
    p = _c(p_grp.detach().float())
    K = prefill_range_cols(S_total, l_sel, n_sel, force_init, force_local)
    out = torch.empty((B, S, G, K, 2), dtype=torch.int32, device=p.device)
    if out.numel():
        _call("nsa_select_ranges_prefill", _ptr(p), B, S, G, S_sel, l_sel, n_sel, S_total, t0, K, int(bool(force_init)), int(force_local),
              _ptr(out), _stream())
    return out


def select_ranges_decode(p_grp: torch.Tensor, l_sel: int, n_sel: int, t: int, *, force_init: bool = True,
                         force_local: int = 2) -> torch.Tensor:
    """p_grp [B,G,S_sel] fp32 -> ranges [B,G,n_sel,2] int32 (select_topn_ranges, selection_scorer.py:124-249)."""
    _require_cuda(p_grp)
    B, G, S_sel = p_grp.shape
    p = _c(p_grp.detach().float())
    out = torch.empty((B, G, n_sel, 2), dtype=torch.int32, device=p.device)
    if out.numel():
        _call("nsa_select_ranges_decode", _ptr(p), B, G, S_sel, l_sel, n_sel, int(t), int(bool(force_init)), int(force_local),
              _ptr(out), _stream())
    return out


def pcmp_all(Q: torch.Tensor, K_cmp: torch.Tensor) -> torch.Tensor:
    """compute_pcmp_all (selection_scorer.py:42-61) as a stand-alone stage: softmax over ALL S_cmp keys of Q.K_cmp^T/sqrt(Dk);
    Q [B,S,G,h,Dk], K_cmp [B,G,S_cmp,Dk] -> p_cmp [B,S,G,h,S_cmp] fp32.  For callers of the reference's free functions: the hot
    path never materialises this tensor, and this stage serves S_cmp up to ~12k."""
    _require_cuda(Q, K_cmp)
    Q, K_cmp = _c(Q.detach()), _c(K_cmp.detach())
    B, S, G, h, _ = Q.shape
    dm = make_dims(Q, NSAConfig(), K_cmp=K_cmp)
    out = torch.empty((B, S, G, h, K_cmp.shape[2]), dtype=torch.float32, device=Q.device)
    if out.numel():
        _call("nsa_pcmp_all", C.byref(dm), _ptr(Q), _ptr(K_cmp), _ptr(out), _stream())
    return out


def map_pcmp_to_pslc(p_cmp: torch.Tensor, S_sel: int, l: int, d: int, l_sel: int) -> torch.Tensor:
    """Eq.9 as a stand-alone stage (map_pcmp_to_pslc(_batched), selection_scorer.py:64-116): p_cmp [..., S_cmp] -> [..., S_sel] fp32."""
    _require_cuda(p_cmp)
    p = _c(p_cmp.detach().float())
    S_cmp = int(p.shape[-1])
    n_rows = p.numel() // S_cmp if S_cmp else 0
    out = torch.zeros((*p.shape[:-1], int(S_sel)), dtype=torch.float32, device=p.device)
    if out.numel() and S_cmp:
        _call("nsa_map_pcmp_to_pslc", _ptr(p), n_rows, S_cmp, int(S_sel), int(l), int(d), int(l_sel), _ptr(out), _stream())
    return out


def indices_to_ranges(indices: torch.Tensor, S_sel: int, l_sel: int, t0: int = 0) -> torch.Tensor:
    """convert_indices_to_ranges_batched(_v2) (selection_scorer.py:380-605): block ids [B,S,G,K] (ascending per row, negative =
    padding) -> [B,S,G,K,2] int32 token ranges, duplicates dropped, adjacent blocks merged, ends clamped to t + 1, [0,0] padded."""
    _require_cuda(indices)
    B, S, G, K = indices.shape
    idx = _c(indices.detach().to(torch.int32))
    out = torch.zeros((B, S, G, K, 2), dtype=torch.int32, device=idx.device)
    if out.numel():
        _call("nsa_indices_to_ranges", _ptr(idx), B, S, G, K, int(S_sel), int(l_sel), int(t0), _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------------------------------
# (1) scoring
# ----------------------------------------------------------------------------------------------------
def score_pgrp(Q: torch.Tensor, K_cmp: torch.Tensor, cfg: NSAConfig, *, t0: int = 0, S_sel: Optional[int] = None) -> torch.Tensor:
    """Q [B,S,G,h,Dk], K_cmp [B,G,S_cmp,Dk] -> p_grp [B,S,G,S_sel] fp32 (Eq.8-10)."""
    _require_cuda(Q, K_cmp)
    Q, K_cmp = _c(Q.detach()), _c(K_cmp.detach())
    B, S, G, h, Dk = Q.shape
    if S_sel is None:
        S_sel = num_sel_blocks(t0 + S, cfg.l_sel)
    dm = make_dims(Q, cfg, t0=t0, K_cmp=K_cmp)
    out = torch.empty((B, S, G, S_sel), dtype=torch.float32, device=Q.device)
    if out.numel():
        _call("nsa_score", C.byref(dm), _ptr(Q), _ptr(K_cmp), S_sel, _ptr(out), _stream())
    return out


def score_select(Q: torch.Tensor, K_cmp: torch.Tensor, cfg: NSAConfig, *, mode: int = 0, t0: int = 0,
                 S_total: Optional[int] = None, S_sel: Optional[int] = None, S_cmp: Optional[int] = None) -> torch.Tensor:
    """Fused scoring + selection -> ranges [B,S,G,K,2] int32.  mode 0: batched-prefill rule, 1: decode rule."""
    _require_cuda(Q, K_cmp)
    Q, K_cmp = _c(Q.detach()), _c(K_cmp.detach())
    B, S, G, h, Dk = Q.shape
    S_total = t0 + S if S_total is None else S_total
    if S_sel is None:
        S_sel = num_sel_blocks(max(S_total, cfg.l_sel), cfg.l_sel)
    K = prefill_range_cols(S_total, cfg.l_sel, cfg.n_sel) if mode == 0 else cfg.n_sel
    dm = make_dims(Q, cfg, t0=t0, K_cmp=K_cmp, n_ranges=K, S_cmp=S_cmp)
    out = torch.zeros((B, S, G, K, 2), dtype=torch.int32, device=Q.device)
    ws_bytes = int(_lib.load().nsa_workspace_bytes(C.byref(dm), _lib.WS_SCORE_SELECT))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=Q.device)
    if out.numel():
        _call("nsa_score_select", C.byref(dm), _ptr(Q), _ptr(K_cmp), S_sel, S_total, mode, _ptr(out), _ptr(ws), _stream())
    return out


def score_stats(Q: torch.Tensor, K_cmp: torch.Tensor, cfg: NSAConfig, *, t0: int = 0) -> torch.Tensor:
    """Pass 1 of the split scorer (nsa_score_stats): [B,S,G,h,2] fp32 row statistics.  Long 16-bit prefill only."""
    _require_cuda(Q, K_cmp)
    Q, K_cmp = _c(Q.detach()), _c(K_cmp.detach())
    B, S, G, h, _ = Q.shape
    dm = make_dims(Q, cfg, t0=t0, K_cmp=K_cmp)
    out = torch.empty((B, S, G, h, 2), dtype=torch.float32, device=Q.device)
    _call("nsa_score_stats", C.byref(dm), _ptr(Q), _ptr(K_cmp), _ptr(out), _stream())
    return out


def score_cmp(Q: torch.Tensor, K_cmp: torch.Tensor, V_cmp: torch.Tensor, cfg: NSAConfig, stats: torch.Tensor, *, t0: int = 0,
              S_sel: Optional[int] = None):
    """Pass 2 of the scorer fused with the compressed branch (nsa_score_cmp) -> (p_grp staging [B,S,G,S_sel] fp32, O_cmp, lse_cmp)."""
    _require_cuda(Q, K_cmp, V_cmp, stats)
    Q, K_cmp, V_cmp = _c(Q.detach()), _c(K_cmp.detach()), _c(V_cmp.detach())
    B, S, G, h, _ = Q.shape
    if S_sel is None:
        S_sel = num_sel_blocks(max(t0 + S, cfg.l_sel), cfg.l_sel)
    dm = make_dims(Q, cfg, t0=t0, K_cmp=K_cmp, V=V_cmp)
    pg = torch.empty((B, S, G, S_sel), dtype=torch.float32, device=Q.device)
    O = torch.empty((B, S, G, h, V_cmp.shape[-1]), dtype=Q.dtype, device=Q.device)
    lse = torch.empty((B, S, G, h), dtype=torch.float32, device=Q.device)
    _call("nsa_score_cmp", C.byref(dm), _ptr(Q), _ptr(K_cmp), _ptr(V_cmp), int(S_sel), _ptr(_c(stats)), _ptr(pg), _ptr(O), _ptr(lse), _stream())
    return pg, O, lse


# ----------------------------------------------------------------------------------------------------
# (3)(4) single-branch attention with autograd
# ----------------------------------------------------------------------------------------------------
BR_CMP, BR_SEL, BR_WIN = 0, 1, 2


def ranges_max_blocks(ranges: torch.Tensor, S_kv: int) -> int:
    """Largest number of 64-key blocks any row of `ranges` [...,K,2] is cut into (nsa_ranges_max_blocks).  Synchronises:
    one int32 comes back to the host."""
    _require_cuda(ranges)
    K = int(ranges.shape[-2])
    r = _c(ranges.detach().to(torch.int32))
    n_rows = r.numel() // max(2 * K, 1)
    out = torch.empty(1, dtype=torch.int32, device=r.device)
    _call("nsa_ranges_max_blocks", _ptr(r), n_rows, K, int(S_kv), _ptr(out), _stream())
    return int(out.item())


def _cfg_for_ranges(cfg: "NSAConfig", ranges: Optional[torch.Tensor], S_kv: int, trusted: bool) -> "NSAConfig":
    """Caller-supplied ranges may break the invariant the tcgen05 selected-branch kernels rely on (<= 16 blocks of 64 keys per
    row, include/nsa_b200.h): such calls run on the SIMT kernels, which take any ranges like the reference's
    grouped_selection_attention (attention_kernels.py:181-226).  Ranges the library selected itself are `trusted`."""
    if ranges is None or trusted or cfg.impl == IMPL_SIMT or ranges.numel() == 0:
        return cfg
    if ranges_max_blocks(ranges, S_kv) <= _lib.MAX_SEL_BLOCKS:
        return cfg
    if cfg.impl == IMPL_TC or int(os.environ.get("NSA_B200_IMPL", 0)) == IMPL_TC:
        raise RuntimeError("ranges cut into more than 16 blocks of 64 keys per row: the tcgen05 selected-branch kernels cannot "
                           "serve them (impl=TC was forced)")
    from dataclasses import replace
    return replace(cfg, impl=IMPL_SIMT)


def _branch_dims(branch, Q, K, V, cfg, ranges, t0, win_off):
    kw = dict(t0=t0, V=V, n_ranges=0 if ranges is None else ranges.shape[3])
    if branch == BR_CMP:
        return make_dims(Q, cfg, K_cmp=K, **kw)
    if branch == BR_SEL:
        return make_dims(Q, cfg, K_sel=K, **kw)
    return make_dims(Q, cfg, K_win=K, win_off=win_off, **kw)


class _BranchAttn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, Q, K, V, ranges, branch, cfg, t0, win_off):
        Qc, Kc, Vc = _c(Q), _c(K), _c(V)
        rg = None if ranges is None else _c(ranges.to(torch.int32))
        dm = _branch_dims(branch, Qc, Kc, Vc, cfg, rg, t0, win_off)
        B, S, G, h, _ = Qc.shape
        O = torch.empty((B, S, G, h, Vc.shape[-1]), dtype=Qc.dtype, device=Qc.device)
        lse = torch.empty((B, S, G, h), dtype=torch.float32, device=Qc.device)
        if O.numel():
            _call("nsa_branch_attn_fwd", C.byref(dm), branch, _ptr(Qc), _ptr(Kc), _ptr(Vc), _ptr(rg), _ptr(O), _ptr(lse), _stream())
        ctx.save_for_backward(Qc, Kc, Vc, rg if rg is not None else torch.empty(0, device=Qc.device), O, lse)
        ctx.meta = (branch, cfg, t0, win_off, rg is not None)
        ctx.mark_non_differentiable(lse)
        return O, lse

    @staticmethod
    def backward(ctx, dO, _dlse):
        Q, K, V, rg, O, lse = ctx.saved_tensors
        branch, cfg, t0, win_off, has_r = ctx.meta
        rg = rg if has_r else None
        dm = _branch_dims(branch, Q, K, V, cfg, rg, t0, win_off)
        ws = _workspace(dm, _lib.WS_BWD, Q.device)
        # one zero-filled fp32 buffer (segments padded to 64 elements), one cast back: see _PrefillCore.backward
        srcs = (Q, K, V)
        sizes = [(t.numel() + 63) // 64 * 64 for t in srcs]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=Q.device)
        offs = [0, sizes[0], sizes[0] + sizes[1]]
        dQ, dK, dV = (flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, srcs))
        dOc = _c(dO.to(Q.dtype))
        if O.numel():
            _call("nsa_branch_attn_bwd", C.byref(dm), branch, _ptr(Q), _ptr(K), _ptr(V), _ptr(rg), _ptr(O), _ptr(lse),
                  _ptr(dOc), _ptr(dQ), _ptr(dK), _ptr(dV), _ptr(ws), _stream())
        if K.dtype == Q.dtype and V.dtype == Q.dtype:
            low = flat.to(Q.dtype)
            return (*(low[o:o + t.numel()].view(t.shape) for o, t in zip(offs, srcs)), None, None, None, None, None)
        return dQ.to(Q.dtype), dK.to(K.dtype), dV.to(V.dtype), None, None, None, None, None


def branch_attention(branch: int, Q, K, V, cfg: NSAConfig, ranges=None, *, t0: int = 0, win_off: int = 0,
                     return_lse: bool = False, ranges_trusted: bool = False):
    """True-softmax attention of one NSA branch (0 cmp, 1 sel, 2 win) with analytical backward.  ranges_trusted: the ranges
    come from this library's selection (<= 16 blocks of 64 keys per row); otherwise they are checked first (one host sync)."""
    _require_cuda(Q, K, V, ranges)
    if branch == BR_SEL:
        cfg = _cfg_for_ranges(cfg, ranges, K.shape[2], ranges_trusted)
    O, lse = _BranchAttn.apply(Q, K, V, ranges, branch, cfg, t0, win_off)
    return (O, lse) if return_lse else O


def sel_attention_blockmajor(Q, K, V, cfg: NSAConfig, ranges, *, t0: int = 0, return_lse: bool = False,
                             ranges_trusted: bool = False):
    """Selected-branch attention, KV-block-major (forward only; nsa_sel_attn_fwd_blockmajor).  Same result as
    branch_attention(BR_SEL, ...)."""
    _require_cuda(Q, K, V, ranges)
    if not ranges_trusted and ranges.numel() and ranges_max_blocks(ranges, K.shape[2]) > _lib.MAX_SEL_BLOCKS:
        raise RuntimeError("sel_attention_blockmajor: a row's ranges cut into more than 16 blocks of 64 keys; use "
                           "branch_attention(BR_SEL, ...), which serves such ranges on the SIMT kernel")
    Qc, Kc, Vc = _c(Q.detach()), _c(K.detach()), _c(V.detach())
    rg = _c(ranges.to(torch.int32))
    dm = _branch_dims(BR_SEL, Qc, Kc, Vc, cfg, rg, t0, 0)
    B, S, G, h, _ = Qc.shape
    O = torch.empty((B, S, G, h, Vc.shape[-1]), dtype=Qc.dtype, device=Qc.device)
    lse = torch.empty((B, S, G, h), dtype=torch.float32, device=Qc.device)
    ws_bytes = int(_lib.load().nsa_workspace_bytes(C.byref(dm), _lib.WS_SEL_BLOCKMAJOR))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=Qc.device)
    if O.numel():
        _call("nsa_sel_attn_fwd_blockmajor", C.byref(dm), _ptr(Qc), _ptr(Kc), _ptr(Vc), _ptr(rg), _ptr(O), _ptr(lse), _ptr(ws), _stream())
    return (O, lse) if return_lse else O


# ----------------------------------------------------------------------------------------------------
# (4) gate
# ----------------------------------------------------------------------------------------------------
def gate_forward(Q: torch.Tensor, gate, cfg: NSAConfig) -> torch.Tensor:
    """GateMLP on q_gp = mean_h(Q): Q [B,S,G,h,Dk] -> gates [B,S,G,3] fp32 (nsa_attention.py:32-82)."""
    _require_cuda(Q)
    Qc = _c(Q.detach())
    gp, keep, hid = _gate_struct(gate, Qc.device)
    dm = make_dims(Qc, cfg, gate_hidden=hid)
    B, S, G = Qc.shape[:3]
    out = torch.empty((B, S, G, 3), dtype=torch.float32, device=Qc.device)
    if out.numel():
        _call("nsa_gate_fwd", C.byref(dm), _ptr(Qc), C.byref(gp), _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------------------------------
# fused hot path: prefill
# ----------------------------------------------------------------------------------------------------
class _PrefillCore(torch.autograd.Function):
    """The drop-in for nsa_attention.py:1066-1398 (scores -> ranges -> three branches -> gated combine)."""

    @staticmethod
    def forward(ctx, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, fc1_w, fc1_b, fc2_w, fc2_b, cfg, sel_mode, t0,
                stopgrad_gates, ranges_in, geom):
        Q, K_sel, V_sel, K_win, V_win = _c(Q), _c(K_sel), _c(V_sel), _c(K_win), _c(V_win)
        K_cmp, V_cmp = _c(K_cmp), _c(V_cmp)
        B, S, G, h, Dk = Q.shape
        Dv = V_sel.shape[-1]
        dev = Q.device
        geom = dict(geom or {})
        gate = (fc1_w, fc1_b, fc2_w, fc2_b) if fc1_w is not None else None
        gp, keep, hid = _gate_struct(gate, dev)
        # needs_input_grad ignores torch.no_grad(): geom carries the caller's grad mode, so that inference neither allocates the
        # per-branch outputs / lse (309 MB per 64k sequence) nor gives up the fused merge + blend
        grad_mode = bool(geom.pop("__grad", True))
        need_grad = grad_mode and any(ctx.needs_input_grad[:11])
        O = torch.empty((B, S, G, h, Dv), dtype=Q.dtype, device=dev)
        gates = torch.empty((B, S, G, 3), dtype=torch.float32, device=dev)
        lse = torch.empty((3, B, S, G, h), dtype=torch.float32, device=dev) if need_grad else None
        O_br = torch.empty((3, B, S, G, h, Dv), dtype=Q.dtype, device=dev) if need_grad else None
        if ranges_in is None:
            # scoring + selection + the three branches + gated combine in ONE C-ABI call (nsa_prefill_full_fwd): for long 16-bit
            # prefill the scorer's second pass and the compressed branch share one kernel
            S_total = t0 + S
            S_sel = num_sel_blocks(max(S_total, cfg.l_sel), cfg.l_sel)
            K = prefill_range_cols(S_total, cfg.l_sel, cfg.n_sel) if sel_mode == 0 else cfg.n_sel
            dm = make_dims(Q, cfg, t0=t0, K_sel=K_sel, K_win=K_win, K_cmp=K_cmp, V=V_sel, n_ranges=K, gate_hidden=hid, **geom)
            ranges = torch.zeros((B, S, G, K, 2), dtype=torch.int32, device=dev)
            ws_bytes = int(_lib.load().nsa_workspace_bytes(C.byref(dm), _lib.WS_PREFILL_FULL))
            ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev)
            if O.numel():
                _call("nsa_prefill_full_fwd", C.byref(dm), _ptr(Q), _ptr(K_sel), _ptr(V_sel), _ptr(K_win), _ptr(V_win), _ptr(K_cmp),
                      _ptr(V_cmp), C.byref(gp), S_sel, S_total, int(sel_mode), _ptr(ranges), _ptr(O), _ptr(lse), _ptr(gates), _ptr(O_br),
                      _ptr(ws), _stream())
        else:
            ranges = _c(ranges_in.to(torch.int32))
            dm = make_dims(Q, cfg, t0=t0, K_sel=K_sel, K_win=K_win, K_cmp=K_cmp, V=V_sel, n_ranges=ranges.shape[3],
                           gate_hidden=hid, **geom)
            ws_bytes = int(_lib.load().nsa_workspace_bytes(C.byref(dm), _lib.WS_PREFILL))
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
            if O.numel():
                _call("nsa_prefill_fwd", C.byref(dm), _ptr(Q), _ptr(K_sel), _ptr(V_sel), _ptr(K_win), _ptr(V_win), _ptr(K_cmp),
                      _ptr(V_cmp), _ptr(ranges), C.byref(gp), _ptr(O), _ptr(lse), _ptr(gates), _ptr(O_br), _ptr(ws), _stream())
        if need_grad:
            ctx.save_for_backward(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, ranges, lse, gates, O_br, *[k for k in keep if k is not None])
            ctx.geom = geom
            ctx.meta = (cfg, t0, hid, stopgrad_gates, gate is not None, [k is not None for k in keep],
                        [None if t is None else t.dtype for t in (fc1_w, fc1_b, fc2_w, fc2_b)])
        ctx.mark_non_differentiable(ranges, gates)
        return O, ranges, gates

    @staticmethod
    def backward(ctx, dO, _dr, _dg):
        Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, ranges, lse, gates, O_br, *gk = ctx.saved_tensors
        cfg, t0, hid, stopgrad, has_gate, gmask, gdt = ctx.meta
        dev = Q.device
        dm = make_dims(Q, cfg, t0=t0, K_sel=K_sel, K_win=K_win, K_cmp=K_cmp, V=V_sel, n_ranges=ranges.shape[3], gate_hidden=hid,
                       **ctx.geom)
        f32 = dict(dtype=torch.float32, device=dev)
        # ONE zero-filled fp32 buffer for the eight accumulators (segments padded to 64 elements: the kernels reduce into them with
        # 16-byte bulk operations) and, below, ONE cast back to the tensors' dtype: 2 launches instead of 15 per layer and step
        srcs = (Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp)
        sizes = [(t.numel() + 63) // 64 * 64 for t in srcs]
        flat = torch.zeros(sum(sizes) + gates.numel(), **f32)
        offs = [0]
        for n in sizes:
            offs.append(offs[-1] + n)
        views = [flat[o:o + t.numel()].view(t.shape) for o, t in zip(offs, srcs)]
        dQ, grads = views[0], views[1:]
        dgates = flat[offs[-1]:].view(gates.shape)
        dOc = _c(dO.to(Q.dtype))
        ws = _workspace(dm, _lib.WS_BWD, dev)
        if Q.numel():
            _call("nsa_prefill_bwd", C.byref(dm), _ptr(Q), _ptr(K_sel), _ptr(V_sel), _ptr(K_win), _ptr(V_win), _ptr(K_cmp),
                  _ptr(V_cmp), _ptr(ranges), _ptr(O_br), _ptr(lse), _ptr(gates), _ptr(dOc), _ptr(dQ),
                  *[_ptr(g) for g in grads], _ptr(dgates), _ptr(ws), _stream())
        dparams = [None, None, None, None]
        if has_gate and cfg.gate_mode == GATE_MLP and not stopgrad:
            it = iter(gk)
            full = [next(it) if m else None for m in gmask]
            gp = _lib.GateParams()
            gp.fc1_w, gp.fc1_b, gp.fc2_w, gp.fc2_b = (None if t is None else t.data_ptr() for t in full)
            H, Dk = full[0].shape
            d1w, d1b = torch.zeros((H, Dk), **f32), torch.zeros((H,), **f32)
            d2w, d2b = torch.zeros((3, H), **f32), torch.zeros((3,), **f32)
            if Q.numel():
                _call("nsa_gate_bwd", C.byref(dm), _ptr(Q), C.byref(gp), _ptr(dgates), _ptr(dQ), _ptr(d1w), _ptr(d1b),
                      _ptr(d2w), _ptr(d2b), _stream())
            dparams = [d1w, d1b if gmask[1] else None, d2w, d2b if gmask[3] else None]
            dparams = [None if g is None else g.to(dt) for g, dt in zip(dparams, gdt)]
        if all(t.dtype == Q.dtype for t in srcs):
            low = flat[:offs[-1]].to(Q.dtype)
            outs = [low[o:o + t.numel()].view(t.shape) for o, t in zip(offs, srcs)]
        else:
            outs = [v.to(t.dtype) for v, t in zip(views, srcs)]
        return (*outs, *dparams, None, None, None, None, None, None)


def prefill_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate, cfg: NSAConfig, *, sel_mode: int = 0, t0: int = 0,
                 stopgrad_gates: bool = False, ranges: Optional[torch.Tensor] = None, S_sel_kv: Optional[int] = None,
                 S_win_kv: Optional[int] = None, S_cmp: Optional[int] = None, win_off: int = 0, ranges_trusted: bool = False):
    """NSA hot path for S query rows.  gate = (fc1_w, fc1_b, fc2_w, fc2_b) or None for forced/uniform gates.
    Returns (O [B,S,G,h,Dv], ranges [B,S,G,K,2] int32, gates [B,S,G,3] fp32).  ranges: use these instead of scoring and
    selecting; unless ranges_trusted (they came from score_select / select_ranges_*), they are checked against the 16-block
    invariant of the tcgen05 kernels first (one host sync) and served by the SIMT kernels when they break it."""
    _require_cuda(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp)
    cfg = _cfg_for_ranges(cfg, ranges, S_sel_kv if S_sel_kv is not None else K_sel.shape[2], ranges_trusted)
    g = gate if gate is not None else (None, None, None, None)
    if gate is not None:
        # the kernels read fp32 gate weights; keep autograd attached to the caller's parameters
        g = tuple(None if t is None else t for t in gate)
    geom = {k: v for k, v in dict(S_sel_kv=S_sel_kv, S_win_kv=S_win_kv, S_cmp=S_cmp).items() if v is not None}
    if win_off:
        geom["win_off"] = win_off
    geom["__grad"] = torch.is_grad_enabled()
    return _PrefillCore.apply(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, g[0], g[1], g[2], g[3], cfg, sel_mode, t0,
                              stopgrad_gates, ranges, geom)


# ----------------------------------------------------------------------------------------------------
# producers of the hot path's inputs: RoPE + re-layout, phi average pool
# ----------------------------------------------------------------------------------------------------
class _RopeShape(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, V, D, rot_dim, dst_layout, t0, base, scale):
        xc = _c(x)
        B, S = xc.shape[0], xc.shape[1]
        y = torch.empty((B, S, V, D) if dst_layout == 0 else (B, V, S, D), dtype=xc.dtype, device=xc.device)
        if y.numel():
            _call("nsa_rope_shape", _ptr(xc), _ptr(y), B, S, V, D, 0, dst_layout, rot_dim, int(t0), float(base), float(scale), 0,
                  _DTYPES[xc.dtype], _stream())
        ctx.meta = (B, S, V, D, rot_dim, dst_layout, int(t0), float(base), float(scale), x.shape)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, S, V, D, rot_dim, dst_layout, t0, base, scale, xshape = ctx.meta
        dyc = _c(dy)
        dx = torch.empty((B, S, V, D), dtype=dyc.dtype, device=dyc.device)
        if dx.numel():
            _call("nsa_rope_shape", _ptr(dyc), _ptr(dx), B, S, V, D, dst_layout, 0, rot_dim, t0, base, scale, 1,
                  _DTYPES[dyc.dtype], _stream())
        return dx.view(xshape), None, None, None, None, None, None, None


def rope_shape(x: torch.Tensor, V: int, D: int, *, rope: str = "none", to_cache_layout: bool = False, t0: int = 0,
               base: float = 10000.0, scale: float = 1.0) -> torch.Tensor:
    """x [B,S,V*D] (a projection output) -> [B,S,V,D] or, with to_cache_layout, [B,V,S,D] in ONE pass, with RoPE
    (nsa/core/rope.py:16-51) applied per D-vector (rope="vector"), across the V*D values of a token as one vector
    (rope="token": what the reference does to Q, nsa_attention.py:1002-1009) or not at all (rope="none": V tensors)."""
    _require_cuda(x)
    rot = {"none": 0, "vector": D, "token": V * D}[rope]
    if not (scale > 0):
        scale = 1.0
    return _RopeShape.apply(x, V, D, rot, 1 if to_cache_layout else 0, t0, base, scale)


class _PhiAvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, l, d, rope, t0):
        xc = _c(x)
        B, G, S, D = xc.shape
        S_cmp = (S - l) // d + 1
        y = torch.empty((B, G, S_cmp, D), dtype=xc.dtype, device=xc.device)
        if y.numel():
            _call("nsa_phi_avgpool", _ptr(xc), _ptr(y), B * G, S, D, l, d, int(rope), int(t0), 10000.0, 1.0, 0, _DTYPES[xc.dtype],
                  _stream())
        ctx.meta = (B, G, S, D, l, d, int(rope), int(t0))
        return y

    @staticmethod
    def backward(ctx, dy):
        B, G, S, D, l, d, rope, t0 = ctx.meta
        dyc = _c(dy)
        dx = torch.empty((B, G, S, D), dtype=dyc.dtype, device=dyc.device)
        if dx.numel():
            _call("nsa_phi_avgpool", _ptr(dyc), _ptr(dx), B * G, S, D, l, d, rope, t0, 10000.0, 1.0, 1, _DTYPES[dyc.dtype], _stream())
        return dx, None, None, None, None


def phi_avgpool(K_raw: torch.Tensor, V_raw: torch.Tensor, l: int, d: int, *, t0: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
    """avg_pool_phi_rope_kv (nsa/core/compress_pool.py:9-38) for rows at positions t0, t0+1, ...: K_cmp = avgpool_{l,d}(RoPE(K_raw)),
    V_cmp = avgpool_{l,d}(V_raw); [B,G,S,D] -> [B,G,(S-l)//d+1,D].  S < l gives empty outputs."""
    _require_cuda(K_raw, V_raw)
    B, G, S, Dk = K_raw.shape
    if S < l:
        return K_raw.new_zeros((B, G, 0, Dk)), V_raw.new_zeros((B, G, 0, V_raw.shape[-1]))
    return _PhiAvgPool.apply(K_raw, l, d, 1, t0), _PhiAvgPool.apply(V_raw, l, d, 0, t0)


class _PhiConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, l, d, rope, t0, scale):
        xc, wc = _c(x), _c(w.detach().float())
        B, G, S, D = xc.shape
        S_cmp = (S - l) // d + 1
        y = torch.empty((B, G, S_cmp, D), dtype=xc.dtype, device=xc.device)
        if y.numel():
            _call("nsa_phi_conv", _ptr(xc), _ptr(wc), _ptr(y), None, B * G, S, D, l, d, int(rope), int(t0), 10000.0, float(scale), 0,
                  _DTYPES[xc.dtype], _stream())
        ctx.save_for_backward(xc, wc)
        ctx.meta = (B, G, S, D, l, d, int(rope), int(t0), float(scale), w.dtype, tuple(w.shape))
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, wc = ctx.saved_tensors
        B, G, S, D, l, d, rope, t0, scale, wdt, wshape = ctx.meta
        dyc = _c(dy.to(xc.dtype))
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(xc)
            _call("nsa_phi_conv", _ptr(dyc), _ptr(wc), _ptr(dx), None, B * G, S, D, l, d, rope, t0, 10000.0, scale, 1, _DTYPES[xc.dtype], _stream())
        if ctx.needs_input_grad[1]:
            dwf = torch.empty((D, l), dtype=torch.float32, device=xc.device)
            _call("nsa_phi_conv", _ptr(xc), None, _ptr(dwf), _ptr(dyc), B * G, S, D, l, d, rope, t0, 10000.0, scale, 2, _DTYPES[xc.dtype], _stream())
            dw = dwf.to(wdt).view(wshape)
        return dx, dw, None, None, None, None, None


def phi_conv(K_raw: torch.Tensor, V_raw: torch.Tensor, w_k: torch.Tensor, w_v: torch.Tensor, l: int, d: int, *, t0: int = 0,
             rope_scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """Learnable phi (phi = "mlp": depthwise Conv1d over time, kernel l, stride d, no bias; nsa_attention.py:275-291, :1741-1777) for
    rows at positions t0, t0+1, ...: K_cmp = conv_k(RoPE(K_raw)), V_cmp = conv_v(V_raw); w_* are the Conv1d weights ([D,1,l] or [D,l]).
    [B,G,S,D] -> [B,G,(S-l)//d+1,D]; gradients flow to the inputs and to the taps.  S < l gives empty outputs."""
    _require_cuda(K_raw, V_raw, w_k, w_v)
    B, G, S, Dk = K_raw.shape
    if S < l:
        return K_raw.new_zeros((B, G, 0, Dk)), V_raw.new_zeros((B, G, 0, V_raw.shape[-1]))
    sc = rope_scale if rope_scale > 0 else 1.0
    return (_PhiConv.apply(K_raw, w_k.reshape(Dk, l), l, d, 1, t0, sc),
            _PhiConv.apply(V_raw, w_v.reshape(V_raw.shape[-1], l), l, d, 0, t0, 1.0))


_ROPE_TABLE_MIN_ROWS = int(os.environ.get("NSA_B200_ROPE_TABLE_MIN_ROWS", "256"))  # 0 or a huge value turns the tables off / on for all
_rope_table_cache: "collections.OrderedDict" = collections.OrderedDict()


def rope_tables(device, dtype, q_dim: int, k_dim: int, t0: int, rows: int, base: float, scale: float):
    """(sin, cos) tables of the producers' rotations for positions [t0, t0 + rows): Q is rotated as one q_dim-wide vector, the K
    streams per k_dim-vector (nsa_attention.py:1002-1009) -> ([rows, q_dim/2, 2], [rows, k_dim/2, 2]) in `dtype`, built by
    nsa_rope_table with the producers' own arithmetic.  Position-only data, so the few most recent geometries are kept (a 64k table
    is 108 MB in bf16); every layer and every step of a model shares them."""
    key = (str(device), dtype, q_dim, k_dim, t0, rows, float(base), float(scale))
    hit = _rope_table_cache.get(key)
    if hit is not None:
        if hit[2] is not None and not torch.cuda.is_current_stream_capturing():  # (event calls are not allowed while capturing)
            hit[2].synchronize()  # built a moment ago, possibly on another stream: wait once, then the entry is plain data
            hit = _rope_table_cache[key] = (hit[0], hit[1], None)
        return hit[0], hit[1]
    tq = torch.empty((rows, q_dim // 2, 2), dtype=dtype, device=device)
    tk = torch.empty((rows, k_dim // 2, 2), dtype=dtype, device=device)
    for tab, dim in ((tq, q_dim), (tk, k_dim)):
        _call("nsa_rope_table", rows, dim // 2, dim, int(t0), float(base), float(scale), _DTYPES[dtype], _ptr(tab), _stream())
    if torch.cuda.is_current_stream_capturing():
        return tq, tk  # built inside the capture (its memory belongs to the graph's pool): not shared
    ev = torch.cuda.Event()
    ev.record()
    _rope_table_cache[key] = (tq, tk, ev)
    # captured graphs hold raw pointers into these tables, so entries leave only under real memory pressure (2 GB of tables)
    while sum(t[0].numel() * t[0].element_size() + t[1].numel() * t[1].element_size() for t in _rope_table_cache.values()) > (2 << 30) \
            and len(_rope_table_cache) > 1:
        _rope_table_cache.popitem(last=False)
    return tq, tk


def decode_produce(y: torch.Tensor, q_out: torch.Tensor, slabs, rows, *, H: int, G: int, Dk: int, Dv: int, t: int,
                   base: float = 10000.0, scale: float = 1.0, counters: Optional[torch.Tensor] = None, counters_idx: int = 0,
                   counter_vals=(0, 0, 0, 0, 0), inverse: bool = False) -> None:
    """Fused projection output y [B,(S,) H*Dk + G*(3*Dk+3*Dv)] = (Q | K_sel | V_sel | K_win | V_win | K_raw | V_raw) -> RoPE'd Q
    into q_out [B,(S,) H*Dk] and the six streams into rows rows[i].. of slabs[i] ([B,G,cap,D], contiguous), in ONE launch: the
    reference's seven rope/view/permute/cat chains (nsa_attention.py:545-586, :998-1016, kv_cache.py:28-49).  inverse=True is the
    backward direction (reads q_out / slabs, writes y).  counters ([5,cap] int64, optional) receives the step's read counters
    (kv_cache.py:51-65) at column counters_idx."""
    _require_cuda(y, q_out, *slabs)
    B = y.shape[0]
    S = y.shape[1] if y.dim() == 3 else 1
    if not y.is_contiguous() or y.shape[-1] != H * Dk + G * (3 * Dk + 3 * Dv) or not q_out.is_contiguous() or q_out.dtype != y.dtype:
        raise RuntimeError("decode_produce: y must be contiguous [B,(S,) H*Dk + G*(3*Dk+3*Dv)] and q_out of the same dtype")
    if q_out.numel() != B * S * H * Dk:
        raise RuntimeError("decode_produce: q_out must hold B*S*H*Dk elements")
    a = _lib.DecodeProduce()
    a.y, a.q_out = y.data_ptr(), q_out.data_ptr()
    for i, (sl, r) in enumerate(zip(slabs, rows)):
        D = Dv if (i & 1) else Dk
        if not sl.is_contiguous() or sl.dtype != y.dtype or sl.shape[0] != B or sl.shape[1] != G or sl.shape[3] != D:
            raise RuntimeError(f"decode_produce: slab {i} must be contiguous [B,G,cap,{D}] of dtype {y.dtype}")
        a.slab[i], a.cap[i], a.row[i] = sl.data_ptr(), int(sl.shape[2]), int(r)
    if counters is not None:
        a.counters, a.counters_cap, a.counters_idx = counters.data_ptr(), int(counters.shape[1]), int(counters_idx)
        for i, v in enumerate(counter_vals):
            a.counter_val[i] = int(v)
    a.B, a.H, a.G, a.Dk, a.Dv, a.t = B, H, G, Dk, Dv, int(t)
    a.base, a.scale, a.dtype = float(base), float(scale if scale > 0 else 1.0), _DTYPES[y.dtype]
    a.S, a.inverse = int(S), int(bool(inverse))
    if S >= _ROPE_TABLE_MIN_ROWS:  # long rows: the rotation's sin / cos come from a cached table instead of sincosf per element
        tq, tk = rope_tables(y.device, y.dtype, H * Dk, Dk, int(t), int(S), a.base, a.scale)
        a.rope_q, a.rope_k, a.rope_t0, a.rope_rows = tq.data_ptr(), tk.data_ptr(), int(t), int(S)
    if B * S:
        _call("nsa_decode_produce", C.byref(a), _stream())


class _ProjectSplit(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, H, G, Dk, Dv, t0, base, scale):
        yc = _c(y)
        B, S, _ = yc.shape
        q = torch.empty((B, S, G, H // G, Dk), dtype=yc.dtype, device=yc.device)
        outs = [torch.empty((B, G, S, Dv if (i & 1) else Dk), dtype=yc.dtype, device=yc.device) for i in range(6)]
        decode_produce(yc, q, outs, [0] * 6, H=H, G=G, Dk=Dk, Dv=Dv, t=t0, base=base, scale=scale)
        ctx.meta = (H, G, Dk, Dv, int(t0), float(base), float(scale), tuple(yc.shape))
        return (q, *outs)

    @staticmethod
    def backward(ctx, dq, *douts):
        H, G, Dk, Dv, t0, base, scale, yshape = ctx.meta
        dy = torch.empty(yshape, dtype=dq.dtype, device=dq.device)
        decode_produce(dy, _c(dq), [_c(g) for g in douts], [0] * 6, H=H, G=G, Dk=Dk, Dv=Dv, t=t0, base=base, scale=scale, inverse=True)
        return dy, None, None, None, None, None, None, None


def project_split(y: torch.Tensor, *, H: int, G: int, Dk: int, Dv: int, t0: int = 0, base: float = 10000.0, scale: float = 1.0):
    """y [B,S,H*Dk + G*(3Dk+3Dv)] (ONE GEMM over the seven stacked projection weights) -> (Q [B,S,G,h,Dk] RoPE'd as the reference
    does (one H*Dk-wide vector per token, nsa_attention.py:1002-1009), K_sel, V_sel, K_win, V_win, K_raw, V_raw [B,G,S,D] with
    K_sel / K_win RoPE'd per Dk-vector) in one launch; the backward is one launch too."""
    _require_cuda(y)
    if y.dim() != 3 or y.shape[-1] != H * Dk + G * (3 * Dk + 3 * Dv) or y.dtype not in _DTYPES:
        raise RuntimeError("project_split: y must be [B,S,H*Dk + G*(3*Dk+3*Dv)] in fp32/bf16/fp16")
    return _ProjectSplit.apply(y, H, G, Dk, Dv, t0, base, scale if scale > 0 else 1.0)


# ----------------------------------------------------------------------------------------------------
# fused hot path: decode
# ----------------------------------------------------------------------------------------------------
def decode_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate, cfg: NSAConfig, *, t: int, S_sel_kv: int,
                S_win_kv: int, win_off: int, S_cmp: int, ranges_out: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, gate_cache=None) -> torch.Tensor:
    """One decode step.  Q [B,1,G,h,Dk]; caches may be over-allocated ([B,G,cap,D], rows present given by S_*).
    Returns O [B,1,G,h,Dv]."""
    _require_cuda(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp)
    Q = _c(Q)
    B, S, G, h, Dk = Q.shape
    if S != 1:
        raise AssertionError(f"Decode mode requires S=1 (single token), got S={S}.")  # nsa_attention.py:532-535
    Dv = V_sel.shape[-1]
    if gate_cache is not None:
        gp, keep, hid = gate_cache
    else:
        gp, keep, hid = _gate_struct(gate, Q.device)
    dm = make_dims(Q, cfg, t0=t, K_sel=K_sel, K_win=K_win, K_cmp=K_cmp, Dv=Dv, S_sel_kv=S_sel_kv, S_win_kv=S_win_kv,
                   S_cmp=S_cmp, win_off=win_off, n_ranges=cfg.n_sel, gate_hidden=hid)
    O = out if out is not None else torch.empty((B, 1, G, h, Dv), dtype=Q.dtype, device=Q.device)
    if ranges_out is None:
        ranges_out = torch.empty((B, G, cfg.n_sel, 2), dtype=torch.int32, device=Q.device)
    ws_bytes = int(_lib.load().nsa_workspace_bytes(C.byref(dm), _lib.WS_DECODE))
    ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=Q.device)
    for name, t_ in (("K_sel", K_sel), ("V_sel", V_sel), ("K_win", K_win), ("V_win", V_win), ("K_cmp", K_cmp), ("V_cmp", V_cmp)):
        if not t_.is_contiguous():
            raise RuntimeError(f"decode_core: {name} must be contiguous [B,G,cap,D]")
    _call("nsa_decode_fwd", C.byref(dm), _ptr(Q), _ptr(K_sel), _ptr(V_sel), _ptr(K_win), _ptr(V_win), _ptr(K_cmp),
          _ptr(V_cmp), C.byref(gp), _ptr(O), _ptr(ranges_out), _ptr(ws), _stream())
    return O


class DecodeStepPlan:
    """Prebuilt argument blocks of one module-level decode step (nsa_decode_produce + nsa_decode_fwd) for a fixed batch and set of
    cache slabs: a step then only updates the scalars that move (position, rows, counts) -- the eager Python step was the bound of
    the module-level decode (197 us per step against 135 us of GPU time at S=4096, B=592).  Same two C-ABI calls as
    decode_produce / decode_core; Q, O and the workspace are persistent buffers of the plan (consumed on the same stream)."""

    def __init__(self, y: torch.Tensor, slabs, cmp_slabs, counters, *, H: int, G: int, Dk: int, Dv: int, cfg: NSAConfig, gate,
                 rope_scale: float):
        _require_cuda(y, *slabs, *cmp_slabs)
        B, dev, dt = y.shape[0], y.device, y.dtype
        h = H // G
        for i, sl in enumerate(tuple(slabs) + tuple(cmp_slabs)):
            D = Dv if (i & 1) else Dk
            if not sl.is_contiguous() or sl.dtype != dt or sl.shape[0] != B or sl.shape[1] != G or sl.shape[3] != D:
                raise RuntimeError(f"DecodeStepPlan: cache slab {i} must be contiguous [B,G,cap,{D}] of dtype {dt}")
        self.slabs, self.cmp_slabs, self.counters = tuple(slabs), tuple(cmp_slabs), counters
        self.B, self.H, self.G, self.h, self.Dk, self.Dv, self.N = B, H, G, h, Dk, Dv, H * Dk + G * (3 * Dk + 3 * Dv)
        self.dtype, self.device = dt, dev
        self.Q = torch.empty((B, 1, G, h, Dk), dtype=dt, device=dev)
        self.O = torch.empty((B, 1, G, h, Dv), dtype=dt, device=dev)
        self.gate_keep = _gate_struct(gate, dev)
        gp, _, hid = self.gate_keep
        self.gp_ref = C.byref(gp)
        a = _lib.DecodeProduce()
        a.q_out = self.Q.data_ptr()
        for i, sl in enumerate(self.slabs):
            a.slab[i], a.cap[i] = sl.data_ptr(), int(sl.shape[2])
        if counters is not None:
            a.counters, a.counters_cap = counters.data_ptr(), int(counters.shape[1])
        a.B, a.H, a.G, a.Dk, a.Dv, a.S, a.inverse = B, H, G, Dk, Dv, 1, 0
        a.base, a.scale, a.dtype = 10000.0, float(rope_scale if rope_scale > 0 else 1.0), _DTYPES[dt]
        self.a, self.a_ref = a, C.byref(a)
        dm = make_dims(self.Q, cfg, K_sel=self.slabs[0], K_win=self.slabs[2], K_cmp=cmp_slabs[0], Dv=Dv, n_ranges=cfg.n_sel,
                       gate_hidden=hid)
        self.dm, self.dm_ref = dm, C.byref(dm)
        lib = _lib.load()
        self.ws = torch.empty(max(int(lib.nsa_workspace_bytes(self.dm_ref, _lib.WS_DECODE)), 1), dtype=torch.uint8, device=dev)
        self.fn_produce, self.fn_decode = lib.nsa_decode_produce, lib.nsa_decode_fwd
        P = lambda t: C.c_void_p(t.data_ptr())
        self.dec_args = (P(self.Q), P(self.slabs[0]), P(self.slabs[1]), P(self.slabs[2]), P(self.slabs[3]), P(cmp_slabs[0]),
                         P(cmp_slabs[1]), self.gp_ref, P(self.O))
        self.ws_ptr = P(self.ws)

    def produce(self, y: torch.Tensor, t: int, rows, counters_idx: int, counter_vals) -> None:
        global launch_count
        if y.shape != (self.B, self.N) or y.dtype != self.dtype or not y.is_contiguous():
            raise RuntimeError("DecodeStepPlan.produce: y must be contiguous [B, H*Dk + G*(3*Dk+3*Dv)] of the plan's dtype")
        a = self.a
        a.y, a.t = y.data_ptr(), t
        for i in range(6):
            a.row[i] = rows[i]
        if self.counters is not None:
            a.counters_idx = counters_idx
            for i in range(5):
                a.counter_val[i] = counter_vals[i]
        _lib.check(self.fn_produce(self.a_ref, _stream()), "nsa_decode_produce")
        launch_count += 1

    def attend(self, t: int, S_win_kv: int, S_cmp: int, ranges_out: torch.Tensor) -> torch.Tensor:
        global launch_count
        dm = self.dm
        dm.t0, dm.S_sel_kv, dm.S_win_kv, dm.win_off, dm.S_cmp = t, t + 1, S_win_kv, (t + 1) - S_win_kv, S_cmp
        _lib.check(self.fn_decode(self.dm_ref, *self.dec_args, C.c_void_p(ranges_out.data_ptr()), self.ws_ptr, _stream()),
                   "nsa_decode_fwd")
        launch_count += 1
        return self.O


class DecodeGraphStep:
    """One module-level decode step -- projection GEMM, nsa_decode_produce, nsa_decode_emit, nsa_decode_fwd_stepped, output GEMM,
    nsa_decode_advance -- captured ONCE in a CUDA graph and replayed per token with no host-side argument patching: position and
    row counts live in a device record (nsa_decode_state_t) that the last kernel of the step advances.  The reference's loop is
    bench/bench_decode.py:123-136 (an eager Python step: ~150 us of host time against ~125 us of GPU time at S=4096, B=592).
    The host only copies the token in, replays, and mirrors the bookkeeping (cache lengths) in Python while the GPU runs."""

    KERNELS_PER_REPLAY = 4  # produce, emit, fused decode, advance (+ two cuBLAS GEMMs)

    def __init__(self, x_like: torch.Tensor, W_cat: torch.Tensor, W_out: torch.Tensor, slabs, cmp_slabs, counters, *, H: int, G: int,
                 Dk: int, Dv: int, cfg: NSAConfig, gate, rope_scale: float, phi_w=None):
        _require_cuda(x_like, W_cat, W_out, *slabs, *cmp_slabs)
        B, dim = x_like.shape[0], x_like.shape[-1]
        dev, dt = x_like.device, x_like.dtype
        h = H // G
        self.B, self.dim, self.H, self.G, self.h, self.Dk, self.Dv, self.cfg = B, dim, H, G, h, Dk, Dv, cfg
        self.N = H * Dk + G * (3 * Dk + 3 * Dv)
        self.W_cat, self.W_out = W_cat, W_out
        self.slabs, self.cmp_slabs, self.counters = tuple(slabs), tuple(cmp_slabs), counters
        self.x = torch.zeros((B, dim), dtype=dt, device=dev)
        self.Q = torch.empty((B, 1, G, h, Dk), dtype=dt, device=dev)
        self.O = torch.empty((B, 1, G, h, Dv), dtype=dt, device=dev)
        self.ranges = torch.zeros((B, G, cfg.n_sel, 2), dtype=torch.int32, device=dev)
        self.state = torch.zeros(8, dtype=torch.int32, device=dev)          # nsa_decode_state_t
        self.state_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self.expected = None                                                 # host mirror of the device record (tuple of 5 ints)
        self.out = None                                                      # [B,1,dim], allocated inside the captured region
        self.gate_keep = _gate_struct(gate, dev)
        gp, _, hid = self.gate_keep
        a = _lib.DecodeProduce()
        a.q_out = self.Q.data_ptr()
        for i, sl in enumerate(self.slabs):
            a.slab[i], a.cap[i] = sl.data_ptr(), int(sl.shape[2])
        if counters is not None:
            a.counters, a.counters_cap = counters.data_ptr(), int(counters.shape[1])
        a.B, a.H, a.G, a.Dk, a.Dv, a.S, a.inverse = B, H, G, Dk, Dv, 1, 0
        a.base, a.scale, a.dtype = 10000.0, float(rope_scale if rope_scale > 0 else 1.0), _DTYPES[dt]
        a.state = self.state.data_ptr()
        a.l, a.d, a.l_sel, a.n_sel, a.w = cfg.l, cfg.d, cfg.l_sel, cfg.n_sel, cfg.w
        self.a = a
        e = _lib.DecodeEmit()
        e.state = self.state.data_ptr()
        e.K_raw, e.V_raw = self.slabs[4].data_ptr(), self.slabs[5].data_ptr()
        e.K_cmp, e.V_cmp = cmp_slabs[0].data_ptr(), cmp_slabs[1].data_ptr()
        e.BG, e.cap_raw, e.cap_cmp, e.Dk, e.Dv, e.l, e.d = B * G, int(self.slabs[4].shape[2]), int(cmp_slabs[0].shape[2]), Dk, Dv, cfg.l, cfg.d
        e.base, e.scale, e.dtype = 10000.0, 1.0, _DTYPES[dt]  # phi pools RoPE'd keys at scale 1 (ops.phi_avgpool)
        if phi_w is not None:  # learnable phi: fp32 copies of the Conv1d taps [D, l] (the graph is re-captured when they change)
            self.phi_w = tuple(_c(t.detach().float().reshape(t.shape[0], -1)) for t in phi_w)
            e.w_k, e.w_v = self.phi_w[0].data_ptr(), self.phi_w[1].data_ptr()
            e.scale = float(rope_scale if rope_scale > 0 else 1.0)
        self.e = e
        self.dm = make_dims(self.Q, cfg, K_sel=self.slabs[0], K_win=self.slabs[2], K_cmp=cmp_slabs[0], Dv=Dv, n_ranges=cfg.n_sel,
                            gate_hidden=hid)
        self.gp = gp
        self.graph = None

    @staticmethod
    def supported(x_like, slabs, cmp_slabs, cfg: NSAConfig, *, H: int, G: int, Dk: int, Dv: int, gate_hidden: int) -> bool:
        if x_like.dtype not in (torch.bfloat16, torch.float16):
            return False
        q = torch.empty((x_like.shape[0], 1, G, H // G, Dk), dtype=x_like.dtype, device="meta")
        dm = make_dims(q, cfg, K_sel=slabs[0], K_win=slabs[2], K_cmp=cmp_slabs[0], Dv=Dv, n_ranges=cfg.n_sel, gate_hidden=gate_hidden)
        return bool(_lib.load().nsa_decode_stepped_supported(C.byref(dm)))

    def set_state(self, t: int, row_win: int, row_raw: int, S_cmp: int, ctr_idx: int) -> None:
        h = self.state_host
        h[0], h[1], h[2], h[3], h[4] = t, row_win, row_raw, S_cmp, ctr_idx
        self.state.copy_(h, non_blocking=True)
        self.expected = (t, row_win, row_raw, S_cmp, ctr_idx)

    def _body(self) -> None:
        import torch.nn.functional as F
        y = F.linear(self.x, self.W_cat)
        self.a.y = y.data_ptr()
        P = lambda t_: C.c_void_p(t_.data_ptr())
        _call("nsa_decode_produce", C.byref(self.a), _stream())
        _call("nsa_decode_emit", C.byref(self.e), _stream())
        _call("nsa_decode_fwd_stepped", C.byref(self.dm), P(self.Q), P(self.slabs[0]), P(self.slabs[1]), P(self.slabs[2]), P(self.slabs[3]),
              P(self.cmp_slabs[0]), P(self.cmp_slabs[1]), C.byref(self.gp), P(self.O), P(self.ranges), P(self.state), _stream())
        self.out = F.linear(self.O.view(self.B, 1, self.H * self.Dv), self.W_out)
        _call("nsa_decode_advance", P(self.state), self.cfg.l, self.cfg.d, _stream())
        self._keep = y

    def capture(self) -> None:
        """Warm up once eagerly (library handles, function attributes), then capture.  The caller sets the state before and after:
        the warm-up advances the record and writes this step's rows, which the first replay rewrites identically."""
        cur = torch.cuda.current_stream(self.x.device)
        side = torch.cuda.Stream(device=self.x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._body()
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def replay(self) -> None:
        global launch_count
        self.graph.replay()
        launch_count += self.KERNELS_PER_REPLAY


# ----------------------------------------------------------------------------------------------------
# caller-side row kernels of the block around the hot path
# ----------------------------------------------------------------------------------------------------
class _RMSNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, eps, out_dtype):
        xc = _c(x)
        rc = None if residual is None else _c(residual)
        dim = xc.shape[-1]
        rows = xc.numel() // dim if dim else 0
        y = torch.empty(xc.shape, dtype=out_dtype, device=xc.device)
        s = torch.empty_like(xc) if rc is not None else xc
        rstd = torch.empty(rows, dtype=torch.float32, device=xc.device)
        wc = _c(weight)
        if rows:
            _call("nsa_rmsnorm_fwd", _ptr(xc), _ptr(rc), _ptr(wc), _ptr(s) if rc is not None else None, _ptr(y), _ptr(rstd), rows, dim,
                  float(eps), _DTYPES[xc.dtype], _DTYPES[rc.dtype] if rc is not None else 0, _DTYPES[wc.dtype], _DTYPES[out_dtype], _stream())
        ctx.save_for_backward(s, wc, rstd)
        ctx.has_res = rc is not None
        ctx.res_dtype = None if rc is None else rc.dtype
        if rc is not None:
            return s, y
        return y

    @staticmethod
    def backward(ctx, *grads):
        s, w, rstd = ctx.saved_tensors
        ds, dy = (grads if ctx.has_res else (None, grads[0]))
        dim = s.shape[-1]
        rows = s.numel() // dim if dim else 0
        if dy is None:  # only the running sum was used downstream
            return (ds, None if ds is None else ds.to(ctx.res_dtype), None, None, None) if ctx.has_res else (None,) * 5
        dyc = _c(dy)
        dsc = None if ds is None else _c(ds.to(s.dtype))
        dx = torch.empty_like(s)
        need_dw = ctx.needs_input_grad[2]
        dw = torch.empty_like(w) if need_dw else None
        part = None
        if need_dw:
            part = torch.empty((int(_lib.load().nsa_rmsnorm_partials(rows)), dim), dtype=torch.float32, device=s.device)
        _call("nsa_rmsnorm_bwd", _ptr(dyc), _ptr(s), _ptr(w), _ptr(rstd), _ptr(dsc), _ptr(dx), _ptr(dw), _ptr(part), rows, dim,
              _DTYPES[s.dtype], _DTYPES[w.dtype], _DTYPES[dyc.dtype], _stream())
        return dx, (dx.to(ctx.res_dtype) if ctx.has_res else None), dw, None, None


def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float = 1e-6, *, residual: Optional[torch.Tensor] = None,
            out_dtype: Optional[torch.dtype] = None):
    """RMSNorm (nsa/model/llama_block_nsa.py:13-22) in one pass, optionally fused with the residual add that feeds it
    (llama_block_nsa.py:102-106) and emitting the dtype its consumers want (under autocast the projections cast the fp32 norm output
    to bf16 once each; rounding the same fp32 value here is bit-equal and is done once).
    residual=None -> y;  residual=r -> (s, y) with s = x + r (dtype of x) and y = norm(s)."""
    _require_cuda(x, weight, residual)
    if x.dtype not in _DTYPES or weight.dtype not in _DTYPES:
        raise RuntimeError(f"rmsnorm: unsupported dtype {x.dtype} / {weight.dtype}")
    if x.shape[-1] % 4 != 0 or x.shape[-1] != weight.shape[-1]:
        raise RuntimeError(f"rmsnorm: dim {x.shape[-1]} must be a multiple of 4 and match the weight ({tuple(weight.shape)})")
    if residual is not None and (residual.shape != x.shape or residual.dtype not in _DTYPES):
        raise RuntimeError("rmsnorm: residual must have the shape of x and a supported dtype")
    od = out_dtype if out_dtype is not None else x.dtype
    if od not in _DTYPES:
        raise RuntimeError(f"rmsnorm: unsupported output dtype {od}")
    return _RMSNorm.apply(x, residual, weight, eps, od)


# ----------------------------------------------------------------------------------------------------
# observability reductions
# ----------------------------------------------------------------------------------------------------
def _ord_to_float(i: int) -> float:
    import struct
    i = i if i >= 0 else i ^ 0x7FFFFFFF
    return struct.unpack("<f", struct.pack("<i", i))[0]


def stats(gates: Optional[torch.Tensor] = None, ranges: Optional[torch.Tensor] = None) -> dict:
    """Gate-health statistics (nsa_attention.py:127-165) of gates [...,3] and / or selection statistics (:455-507) of ranges
    [...,K,2], reduced on the device and read back with ONE device->host copy of an 80-byte record.  Returns the reference's keys:
    entropy_mean/min, max_gate_mean/max, branch_shares, collapse_fraction, total_gates; k_mean, k_max, rows, pct_at_max."""
    _require_cuda(gates, ranges)
    dev = (gates if gates is not None else ranges).device
    g = None
    if gates is not None:
        g = _c(gates.detach().reshape(-1, 3).float())
    r, K = None, 0
    if ranges is not None:
        K = int(ranges.shape[-2]) if ranges.dim() >= 2 else 0
        r = _c(ranges.detach().reshape(-1, K, 2).to(torch.int32)) if ranges.numel() else None
    n_g = 0 if g is None else int(g.shape[0])
    n_r = 0 if r is None else int(r.shape[0])
    rec = torch.empty(C.sizeof(_lib.Stats), dtype=torch.uint8, device=dev)
    row_len = torch.empty(max(n_r, 1), dtype=torch.int32, device=dev)
    _call("nsa_stats", _ptr(g) if n_g else None, n_g, _ptr(r) if n_r else None, n_r, K, _ptr(row_len), _ptr(rec), _stream())
    st = _lib.Stats.from_buffer_copy(rec.cpu().numpy().tobytes())  # the one transfer (synchronises)
    out = {}
    if gates is not None:
        if n_g:
            out.update(entropy_mean=st.gate_sum[0] / n_g, entropy_min=_ord_to_float(st.entropy_min_ord),
                       max_gate_mean=st.gate_sum[1] / n_g, max_gate_max=_ord_to_float(st.max_gate_max_ord),
                       branch_shares=[st.gate_sum[3] / n_g, st.gate_sum[4] / n_g, st.gate_sum[5] / n_g],
                       collapse_fraction=st.gate_sum[2] / n_g, total_gates=n_g)
        else:
            out.update(entropy_mean=float("nan"), entropy_min=float("nan"), max_gate_mean=float("nan"), max_gate_max=float("nan"),
                       branch_shares=[float("nan")] * 3, collapse_fraction=float("nan"), total_gates=0)
    if ranges is not None:
        if n_r:
            out.update(k_mean=st.k_sum / n_r, k_max=int(st.k_max), rows=n_r, pct_at_max=(st.rows_at_max / n_r) if st.k_max > 0 else 0.0)
        else:
            out.update(k_mean=0.0, k_max=0, rows=0, pct_at_max=0.0)
    return out
