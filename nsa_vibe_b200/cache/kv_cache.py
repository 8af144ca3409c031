"""NSA_KV -- the reference's per-branch cache container (nsa/cache/kv_cache.py:8-65) with the same fields,
shapes and update methods, but appends go into pre-allocated slabs instead of `torch.cat` per token
(the reference copies O(S) bytes per decode step, kv_cache.py:28-49).  The public tensors are views:

    K_sel, V_sel          [B,G,S,D]        every token (K with RoPE)
    K_win, V_win          [B,G,min(w,S),D] last w tokens (a view of the full window slab)
    K_cmp_raw_seq, V_...  [B,G,S,D]        raw (un-rotated) projections feeding phi
    K_cmp, V_cmp          [B,G,S_cmp,D]    emitted compressed tokens

Callers may still construct it with zero-length tensors exactly as the reference's benches/tests do
(bench/bench_decode.py:14-33, nsa/model/llama_block_nsa.py:74-101).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

from ..core.block_index import BlockMeta


@dataclass
class NSA_KV:
    K_sel: torch.Tensor
    V_sel: torch.Tensor
    K_win: torch.Tensor
    V_win: torch.Tensor
    K_cmp_raw_seq: torch.Tensor
    V_cmp_raw_seq: torch.Tensor
    K_cmp: torch.Tensor
    V_cmp: torch.Tensor
    win_ptr: torch.Tensor
    cmp_emit_next: torch.Tensor
    meta: BlockMeta
    reads_pred: torch.Tensor
    reads_act_total: torch.Tensor
    reads_act_sel: torch.Tensor
    reads_act_cmp: torch.Tensor
    reads_act_win: torch.Tensor
    # slab bookkeeping (not part of the reference's dataclass)
    _slabs: Dict[str, torch.Tensor] = field(default_factory=dict, repr=False, compare=False)
    _views: Dict[str, torch.Tensor] = field(default_factory=dict, repr=False, compare=False)
    _lens: Dict[str, int] = field(default_factory=dict, repr=False, compare=False)

    # ---- lazily rebuilt views ------------------------------------------------------------------
    # A decode step only bumps the row counts (commit_*): the public tensors of the fields it touched are dropped from the
    # instance and rebuilt from the slab on the next read (__getattr__ runs only when the attribute is missing), so a step costs
    # a few dictionary operations instead of thirteen narrow() calls.
    _CACHE_FIELDS = ("K_sel", "V_sel", "K_win", "V_win", "K_cmp_raw_seq", "V_cmp_raw_seq", "K_cmp", "V_cmp")

    def __getattr__(self, name: str):
        d = self.__dict__
        slabs = d.get("_slabs")
        if slabs is not None:
            if name in NSA_KV._CACHE_FIELDS and name in slabs:
                self._set_view(name)
                return d[name]
            if name in NSA_KV._COUNTER_FIELDS and "__ctr" in slabs:
                v = slabs["__ctr"][NSA_KV._COUNTER_FIELDS.index(name), :d["_lens"]["__ctr"]]
                d["_views"][name] = v
                d[name] = v
                return v
        raise AttributeError(name)

    def _invalidate(self, name: str) -> None:
        self.__dict__.pop(name, None)
        self._views.pop(name, None)

    def _is_synced(self, name: str) -> bool:
        """The public tensor of `name` is (or, when dropped, will be rebuilt as) the view of its slab."""
        d = self.__dict__
        return name not in d or self._views.get(name) is d[name]

    # ---- slab management -----------------------------------------------------------------------
    def _in_sync(self, name: str) -> bool:
        return name in self._slabs and self._is_synced(name)

    def _ensure(self, name: str, extra: int) -> None:
        """Make `name` slab-backed with room for `extra` more rows, preserving its current content.  A field the
        caller assigned directly (zero-length placeholders, reference-style exact-size tensors) is adopted by copy."""
        cur: torch.Tensor = getattr(self, name)
        sync = self._in_sync(name)
        n = self._lens[name] if sync else cur.shape[2]
        if sync and self._slabs[name].shape[2] >= n + extra:
            return
        keep = self._slabs[name][:, :, :n] if sync else cur
        B, G, _, D = cur.shape
        cap = max(n + extra, 2 * n, 64)
        new = torch.zeros((B, G, cap, D), dtype=cur.dtype, device=cur.device)
        if n:
            new[:, :, :n] = keep
        self._slabs[name] = new
        self._lens[name] = n
        self._set_view(name)

    def _set_view(self, name: str) -> None:
        n = self._lens[name]
        window = self._lens.get("__w_" + name)
        lo = max(0, n - window) if window is not None else 0
        v = self._slabs[name].narrow(2, lo, n - lo)
        self._views[name] = v
        setattr(self, name, v)

    def _append(self, name: str, x: torch.Tensor, window: Optional[int] = None, adopt: bool = False) -> None:
        cur: torch.Tensor = getattr(self, name)
        if cur.dtype != x.dtype or cur.device != x.device:  # zero-length placeholder of another dtype/device
            if cur.shape[2] != 0:
                raise RuntimeError(f"NSA_KV.{name}: dtype/device of the cache does not match the new tokens")
            setattr(self, name, x.new_zeros((x.shape[0], x.shape[1], 0, x.shape[3])))
            self._slabs.pop(name, None)
        if window is not None:
            self._lens["__w_" + name] = int(window)
        s = x.shape[2]
        if adopt and s > 0 and x.is_contiguous():
            # prefill into an empty cache without reserved room: the caller's fresh tensor BECOMES the slab (no zero-filled
            # allocation, no copy: 100 MB of device copies per 64k prefill).  It has no spare capacity, so the next append moves to a
            # larger slab first and the adopted tensor is never written through the cache.
            sync = self._in_sync(name)
            n0 = self._lens[name] if sync else getattr(self, name).shape[2]
            if n0 == 0 and not (sync and self._slabs[name].shape[2] >= s):
                self._slabs[name] = x
                self._lens[name] = s
                self._set_view(name)
                return
        self._ensure(name, s)
        n = self._lens[name]
        self._slabs[name][:, :, n:n + s] = x
        self._lens[name] = n + s
        self._set_view(name)

    def reserve(self, capacity: int) -> "NSA_KV":
        """Pre-allocate every cache for `capacity` tokens (decode then never reallocates)."""
        for name in ("K_sel", "V_sel", "K_win", "V_win", "K_cmp_raw_seq", "V_cmp_raw_seq", "K_cmp", "V_cmp"):
            have = self._lens[name] if self._in_sync(name) else getattr(self, name).shape[2]
            self._ensure(name, max(capacity - have, 0))
        return self

    def slab(self, name: str) -> torch.Tensor:
        """Backing [B,G,cap,D] slab of a cache field (adopting it if needed)."""
        self._ensure(name, 0)
        return self._slabs[name]

    def length(self, name: str) -> int:
        """Rows held in the slab of `name` (for the window slab: its whole history, not just the last w)."""
        self._ensure(name, 0)
        return self._lens[name]

    # ---- the reference's update API ----------------------------------------------------------------
    def update_selection_raw(self, K: torch.Tensor, V: torch.Tensor, adopt: bool = False) -> None:
        """adopt=True: K / V are fresh tensors the caller will not write again (see _append)."""
        self._append("K_sel", K, adopt=adopt)
        self._append("V_sel", V, adopt=adopt)

    def update_window(self, K: torch.Tensor, V: torch.Tensor, w: int, adopt: bool = False) -> None:
        self._append("K_win", K, window=w, adopt=adopt)  # view keeps the last w tokens (kv_cache.py:32-38)
        self._append("V_win", V, window=w, adopt=adopt)

    def update_compressed(self, K_raw_cmp: torch.Tensor, V_raw_cmp: torch.Tensor, l: int, d: int) -> None:
        # prefill: replace with the freshly pooled sequence (kv_cache.py:40-45)
        self.K_cmp = K_raw_cmp
        self.V_cmp = V_raw_cmp
        self._slabs.pop("K_cmp", None)
        self._slabs.pop("V_cmp", None)

    def append_compressed(self, K_new: torch.Tensor, V_new: torch.Tensor) -> None:
        """Decode emission: append one compressed token (nsa_attention.py:598-604 does it with torch.cat)."""
        self._append("K_cmp", K_new)
        self._append("V_cmp", V_new)

    def append_cmp_raw(self, K_raw_tok: torch.Tensor, V_raw_tok: torch.Tensor, adopt: bool = False) -> None:
        self._append("K_cmp_raw_seq", K_raw_tok, adopt=adopt)
        self._append("V_cmp_raw_seq", V_raw_tok, adopt=adopt)

    # ---- in-place decode append (one kernel writes the six token rows, ops.decode_produce) --------------
    _TOKEN_FIELDS = ("K_sel", "V_sel", "K_win", "V_win", "K_cmp_raw_seq", "V_cmp_raw_seq")
    _COUNTER_FIELDS = ("reads_pred", "reads_act_total", "reads_act_sel", "reads_act_cmp", "reads_act_win")

    def token_append_slots(self, like: torch.Tensor):
        """Slabs and row indices where the next token's (K_sel, V_sel, K_win, V_win, K_raw, V_raw) rows go; the caller writes
        them on the device and then calls commit_token_append()."""
        slabs, rows = [], []
        for name in self._TOKEN_FIELDS:
            cur: torch.Tensor = getattr(self, name)
            if cur.dtype != like.dtype or cur.device != like.device:
                if cur.shape[2] != 0:
                    raise RuntimeError(f"NSA_KV.{name}: dtype/device of the cache does not match the new tokens")
                setattr(self, name, like.new_zeros((cur.shape[0], cur.shape[1], 0, cur.shape[3])))
                self._slabs.pop(name, None)
            self._ensure(name, 1)
            slabs.append(self._slabs[name])
            rows.append(self._lens[name])
        return slabs, rows

    def fast_token_rows(self, slabs) -> Optional[list]:
        """Row indices for the next token if the six token caches are still backed by exactly `slabs` (a tuple handed out by
        token_append_slots earlier), in sync with the public views and with a free row each; else None (no tensor op is run)."""
        sl, vw, ln = self._slabs, self._views, self._lens
        rows = []
        for i, name in enumerate(self._TOKEN_FIELDS):
            s = sl.get(name)
            if s is not slabs[i] or not self._is_synced(name):
                return None
            n = ln[name]
            if n >= s.shape[2]:
                return None
            rows.append(n)
        return rows

    def decode_state(self, token_slabs, cmp_slabs, counters) -> Optional[list]:
        """Rows for the next token if this cache is still exactly what a prebuilt decode plan points at: the six token slabs
        (`token_slabs`) with a free row each, the two compressed slabs (`cmp_slabs`) and, when given, the counter slab with a free
        column, all in sync with the public tensors.  None = something was reallocated or replaced: rebuild the plan."""
        rows = self.fast_token_rows(token_slabs)
        if rows is None:
            return None
        if not self.same_compressed_slabs(cmp_slabs):
            return None
        sl, vw = self._slabs, self._views
        if counters is not None and (sl.get("__ctr") is not counters or self._lens["__ctr"] >= counters.shape[1]
                                     or not all(self._is_synced(f) for f in self._COUNTER_FIELDS)):
            return None
        return rows

    def same_compressed_slabs(self, cmp_slabs) -> bool:
        """True if K_cmp / V_cmp are still backed by exactly these slabs (an emission that outgrows a slab reallocates it)."""
        sl = self._slabs
        return (sl.get("K_cmp") is cmp_slabs[0] and sl.get("V_cmp") is cmp_slabs[1] and self._is_synced("K_cmp")
                and self._is_synced("V_cmp"))

    def prepare_decode(self, like: torch.Tensor, with_counters: bool):
        """Make every cache a decode step touches slab-backed with room for one more row (token caches, one compressed token an
        emission may add, one counter column): returns (token_slabs, rows, (K_cmp slab, V_cmp slab), counter slab or None)."""
        slabs, rows = self.token_append_slots(like)
        for name in ("K_cmp", "V_cmp"):
            cur: torch.Tensor = getattr(self, name)
            if cur.dtype != like.dtype or cur.device != like.device:
                if cur.shape[2] != 0:
                    raise RuntimeError(f"NSA_KV.{name}: dtype/device of the cache does not match the new tokens")
                setattr(self, name, like.new_zeros((cur.shape[0], cur.shape[1], 0, cur.shape[3])))
                self._slabs.pop(name, None)
            self._ensure(name, 1)
        ctr = self.counter_slot()[0] if with_counters else None
        return tuple(slabs), rows, (self._slabs["K_cmp"], self._slabs["V_cmp"]), ctr

    def rows_present(self, name: str) -> int:
        """Rows held in the slab of `name` when it is known to be slab-backed and in sync (no tensor op)."""
        return self._lens[name]

    def counter_column(self) -> int:
        return self._lens["__ctr"]

    def commit_token_append(self, w: int) -> None:
        lens = self._lens
        lens["__w_K_win"] = lens["__w_V_win"] = int(w)
        for name in self._TOKEN_FIELDS:
            lens[name] += 1
            self._invalidate(name)

    def commit_compressed_append(self) -> None:
        """One compressed token was written on the device into the next free row of the K_cmp / V_cmp slabs (nsa_decode_emit)."""
        for name in ("K_cmp", "V_cmp"):
            self._lens[name] += 1
            self._invalidate(name)

    def counter_slot(self):
        """([5,cap] int64 slab, column) where the next step's read counters go (rows ordered as _COUNTER_FIELDS); the five
        public tensors become views of it.  commit_counters() publishes the column."""
        sync = "__ctr" in self._slabs and all(self._is_synced(f) for f in self._COUNTER_FIELDS)
        n = self._lens["__ctr"] if sync else int(self.reads_pred.numel())
        if not sync or self._slabs["__ctr"].shape[1] <= n:
            new = torch.zeros((5, max(2 * n, 1024)), dtype=torch.int64, device=self.reads_pred.device)
            for i, f in enumerate(self._COUNTER_FIELDS):
                cur = getattr(self, f)
                m = min(n, int(cur.numel()))
                if m:
                    new[i, :m] = cur[:m]
            self._slabs["__ctr"], self._lens["__ctr"] = new, n
        return self._slabs["__ctr"], n

    def commit_counters(self) -> None:
        self._lens["__ctr"] += 1
        for f in self._COUNTER_FIELDS:
            self._invalidate(f)

    # ---- read counters (kv_cache.py:51-65) --------------------------------------------------------
    @staticmethod
    def _cat(t: torch.Tensor, val: int) -> torch.Tensor:
        v = torch.full((1,), int(val), dtype=torch.int64, device=t.device)
        return torch.cat([t, v], dim=0) if t.numel() else v

    def append_reads_pred(self, value: int) -> None:
        self.reads_pred = self._cat(self.reads_pred, value)

    def append_reads_actual(self, total: int, sel: int, cmp: int, win: int) -> None:
        self.reads_act_total = self._cat(self.reads_act_total, total)
        self.reads_act_sel = self._cat(self.reads_act_sel, sel)
        self.reads_act_cmp = self._cat(self.reads_act_cmp, cmp)
        self.reads_act_win = self._cat(self.reads_act_win, win)


def create_empty_kv(B: int, G: int, d_k: int, d_v: int, meta: BlockMeta, device=None, dtype=torch.float32) -> NSA_KV:
    """Zero-length caches, as bench/bench_decode.py:14-33 builds them."""
    device = device if device is not None else torch.device("cuda")
    z = lambda D: torch.zeros((B, G, 0, D), device=device, dtype=dtype)
    zi = lambda: torch.zeros((0,), dtype=torch.int64, device=device)
    return NSA_KV(K_sel=z(d_k), V_sel=z(d_v), K_win=z(d_k), V_win=z(d_v), K_cmp_raw_seq=z(d_k), V_cmp_raw_seq=z(d_v),
                  K_cmp=z(d_k), V_cmp=z(d_v), win_ptr=torch.zeros((B, G), dtype=torch.int32, device=device),
                  cmp_emit_next=torch.zeros((B, G), dtype=torch.int32, device=device), meta=meta, reads_pred=zi(),
                  reads_act_total=zi(), reads_act_sel=zi(), reads_act_cmp=zi(), reads_act_win=zi())
