"""NSAAttention -- drop-in for the reference module (nsa/core/nsa_attention.py:168-1855).

Same constructor, parameter names (state-dict compatible: W_Q, W_K_sel, W_V_sel, W_K_win, W_V_win, W_K_cmp,
W_V_cmp, out, gate.fc1, gate.fc2), `forward(x, kv, *, prefill) -> (out, kv)`, observability getters and
NSA_* environment flags (snapshotted at construction like the reference, :300-394).  Everything between
"Q/K/V projected + RoPE'd" and "O handed to self.out" runs in hand-written sm_100a kernels through
libnsa_b200.so (ops.prefill_core / ops.decode_core); so do RoPE, the re-layout into the cache format and the phi pooling that
produce the hot path's inputs (ops.rope_shape / ops.phi_avgpool, one pass each); the projections themselves are nn.Linear.

Semantics: all three branches are true softmax over their allowed keys (the reference's default
per-token SDPA routes attend to one key only, F1).  Selection follows the reference's rule for the mode in use
(F2): NSA_PREFILL_BATCHED=1 -> select_topn_ranges_batched; sequential prefill / decode -> select_topn_ranges.
Flags that only choose among the reference's removed back ends (NSA_USE_FA2*, NSA_USE_TRITON_SEL, NSA_SEL_CUDA,
NSA_USE_SEL_PACK/MASK/VARLEN, NSA_FORCE_SEL_MASK, NSA_USE_CMP_MASK, NSA_USE_WIN_MASK, NSA_FORCE_PARITY) are accepted
and ignored: every route is the one CUDA path.  CPU tensors raise -- there is no fallback.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..cache.kv_cache import NSA_KV
from .block_index import build_block_meta


def _env_true(name: str, default: str = "0") -> bool:
    return os.getenv(name, default).lower() in ("1", "true", "yes")


class GateMLP(nn.Module):
    """Gate network (nsa_attention.py:32-82).  The fused kernels evaluate it on-chip from these parameters;
    this torch forward exists for callers that invoke the gate directly."""

    def __init__(self, d_k: int, hidden: Optional[int] = None):
        super().__init__()
        hidden = hidden or max(1, d_k // 2)
        self.fc1 = nn.Linear(d_k, hidden)
        self.fc2 = nn.Linear(hidden, 3)
        nn.init.xavier_uniform_(self.fc2.weight, gain=0.1)
        nn.init.zeros_(self.fc2.bias)
        self._force_uniform_gate = _env_true("NSA_FORCE_UNIFORM_GATE")
        self._force_branch = os.getenv("NSA_FORCE_BRANCH")

    def mode(self) -> int:
        if self._force_uniform_gate:
            return ops.GATE_UNIFORM
        fb = (self._force_branch or "").strip().lower()
        return {"cmp": ops.GATE_CMP, "sel": ops.GATE_SEL, "win": ops.GATE_WIN}.get(fb, ops.GATE_MLP)

    def params(self):
        return (self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias)

    def forward(self, q_group_pooled: torch.Tensor, tau: float = 1.0) -> torch.Tensor:
        m = self.mode()
        shape = (*q_group_pooled.shape[:-1], 3)
        if m == ops.GATE_UNIFORM:
            return torch.full(shape, 1.0 / 3.0, device=q_group_pooled.device, dtype=q_group_pooled.dtype)
        if m != ops.GATE_MLP:
            one = torch.zeros(shape, device=q_group_pooled.device, dtype=q_group_pooled.dtype)
            one[..., m - ops.GATE_CMP] = 1.0
            return one
        g = self.fc2(F.silu(self.fc1(q_group_pooled))) / max(tau, 1e-6)
        p = F.softmax(g, dim=-1)
        top2 = torch.topk(g.detach(), k=2, dim=-1).values
        peaked = (top2[..., 0] - top2[..., 1]) > 50.0
        hard = F.one_hot(torch.argmax(g, dim=-1), 3).to(p.dtype)
        return torch.where(peaked.unsqueeze(-1), hard, p)


def _compute_gate_stats(gates: torch.Tensor) -> dict:
    """Gate health statistics (nsa_attention.py:127-165): one reduction kernel, one device->host transfer (ops.stats)."""
    st = ops.stats(gates=gates)
    return {k: st[k] for k in ("entropy_mean", "entropy_min", "max_gate_mean", "max_gate_max", "branch_shares", "collapse_fraction",
                               "total_gates")}


class NSAAttention(nn.Module):
    def __init__(self, dim: int, n_heads: int, n_kv_groups: int, d_k: int, d_v: int, l: int = 32, d: int = 16,
                 l_sel: int = 64, n_sel: int = 16, w: int = 512, phi: str = "avg", gate_hidden: Optional[int] = None,
                 gate_temp: float = 1.0, rope_impl: str = "llama", use_flash: bool = True,
                 use_triton_sel: bool = False) -> None:
        super().__init__()
        assert n_heads % n_kv_groups == 0, "heads must be divisible by kv groups"
        if l % d != 0 or l_sel % d != 0:
            raise ValueError("M0 requires d|l and d|l_sel; set valid block sizes/stride.")
        self.dim, self.n_heads, self.n_kv_groups = dim, n_heads, n_kv_groups
        self.h_per_group = n_heads // n_kv_groups
        self.d_k, self.d_v = d_k, d_v
        self.l, self.d, self.l_sel, self.n_sel, self.w = l, d, l_sel, n_sel, w
        self.gate_temp = gate_temp
        self.phi_type = (phi or "avg").lower()
        if self.phi_type not in ("avg", "mlp"):
            raise ValueError(f"phi must be 'avg' or 'mlp', got {phi!r}")
        self._last_gates: Optional[torch.Tensor] = None
        self._last_ranges: Optional[torch.Tensor] = None
        self._fallback_counters = {k: 0 for k in (
            "selection_triton_fails", "selection_cuda_fails", "selection_pack_fails", "selection_mask_fails",
            "compressed_fa2_fails", "sliding_fa2_fails", "total_fallbacks")}
        self.W_Q = nn.Linear(dim, n_heads * d_k, bias=False)
        self.W_K_sel = nn.Linear(dim, n_kv_groups * d_k, bias=False)
        self.W_V_sel = nn.Linear(dim, n_kv_groups * d_v, bias=False)
        self.W_K_win = nn.Linear(dim, n_kv_groups * d_k, bias=False)
        self.W_V_win = nn.Linear(dim, n_kv_groups * d_v, bias=False)
        self.W_K_cmp = nn.Linear(dim, n_kv_groups * d_k, bias=False)
        self.W_V_cmp = nn.Linear(dim, n_kv_groups * d_v, bias=False)
        self.out = nn.Linear(n_heads * d_v, dim, bias=False)
        self.gate = GateMLP(d_k, gate_hidden)
        # Learnable phi (nsa_attention.py:275-291): depthwise Conv1d over time, kernel l, stride d, initialised to the average pool.
        # (In the reference the line that clears phi_v_conv is dedented one level too far, :289-291, so phi_v_conv is always None
        # and phi="mlp" asserts on first use; this is the constructor as its comment describes it.)
        self.phi_k_conv: Optional[nn.Conv1d] = None
        self.phi_v_conv: Optional[nn.Conv1d] = None
        if self.phi_type == "mlp":
            self.phi_k_conv = nn.Conv1d(d_k, d_k, kernel_size=l, stride=d, groups=d_k, bias=False)
            self.phi_v_conv = nn.Conv1d(d_v, d_v, kernel_size=l, stride=d, groups=d_v, bias=False)
            with torch.no_grad():
                self.phi_k_conv.weight.fill_(1.0 / float(l))
                self.phi_v_conv.weight.fill_(1.0 / float(l))
        self.use_flash_default = use_flash      # accepted for API compatibility; no flash-attn route exists
        self.use_triton_sel = use_triton_sel    # accepted for API compatibility; no Triton route exists
        self._cache_env_vars()

    # ---- flags (nsa_attention.py:300-394) ---------------------------------------------------------
    def _cache_env_vars(self) -> None:
        self._env_cache = {
            "prefill_batched": _env_true("NSA_PREFILL_BATCHED"),
            "strict_asserts": _env_true("NSA_STRICT_ASSERTS"),
            "stopgrad_gates": _env_true("NSA_STOPGRAD_GATES"),
            "nvtx": _env_true("NSA_NVTX"),
            "disable_aux_stats": _env_true("NSA_DISABLE_AUX_STATS"),
            "pcmp_norm": os.getenv("NSA_PCMP_NORM", "full_row").strip().lower(),  # B200 addition
            # selection_scorer.py:46-56: score fp32 activations with bf16 operands (here: the tcgen05 scorer, fp32 accumulation)
            "pcmp_mixed": os.getenv("NSA_P_CMP_MIXED", "0").lower() in ("1", "true", "yes", "on"),
            # B200 addition: replay the decode step as a CUDA graph (ops.DecodeGraphStep) when the fused tcgen05 decode kernel serves
            # the shape; NSA_DECODE_GRAPH=0 keeps the eager step
            "decode_graph": _env_true("NSA_DECODE_GRAPH", "1"),
        }
        try:
            rs = float(os.getenv("NSA_ROPE_SCALE", "1.0"))
            self.rope_scale = rs if (rs > 0.0 and rs == rs) else 1.0
        except ValueError:
            self.rope_scale = 1.0
        try:
            self.prefill_tile = max(0, int(os.getenv("NSA_PREFILL_TILE", "0")))
        except ValueError:
            self.prefill_tile = 0

    def _cfg(self, *, causal_norm: bool = False) -> ops.NSAConfig:
        norm = ops.NORM_CAUSAL if (causal_norm or self._env_cache["pcmp_norm"] == "causal") else ops.NORM_FULL_ROW
        return ops.NSAConfig(l=self.l, d=self.d, l_sel=self.l_sel, n_sel=self.n_sel, w=self.w,
                             gate_tau=float(self.gate_temp), gate_mode=self.gate.mode(), norm_mode=norm)

    def _shape_q(self, Q: torch.Tensor, B: int, S: int) -> torch.Tensor:
        return Q.view(B, S, self.n_kv_groups, self.h_per_group, self.d_k)

    def _shape_kv(self, X: torch.Tensor, B: int, S: int) -> torch.Tensor:
        return X.view(B, S, self.n_kv_groups, -1).permute(0, 2, 1, 3).contiguous()  # [B,G,S,D*]

    # ---- observability (nsa_attention.py:407-507); stats are reduced on demand, not per step ---------
    def get_gate_stats(self) -> Optional[dict]:
        return None if self._last_gates is None else _compute_gate_stats(self._last_gates)

    def get_fallback_counters(self) -> dict:
        return self._fallback_counters.copy()  # always zero: there is nothing to fall back to

    def reset_fallback_counters(self) -> dict:
        prev = self._fallback_counters.copy()
        for k in self._fallback_counters:
            self._fallback_counters[k] = 0
        return prev

    def get_selection_stats(self) -> Optional[dict]:
        """nsa_attention.py:455-507, reduced on the device (ops.stats)."""
        r = self._last_ranges
        if r is None:
            return None
        base = {"l_sel": int(self.l_sel), "n_sel": int(self.n_sel)}
        if r.numel() == 0:
            return {"k_mean": 0.0, "k_max": 0, "rows": 0, "pct_at_max": 0.0, **base}
        st = ops.stats(ranges=r)
        return {"k_mean": st["k_mean"], "k_max": st["k_max"], "rows": st["rows"], "pct_at_max": st["pct_at_max"], **base}

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, kv: NSA_KV, *, prefill: bool) -> tuple[torch.Tensor, NSA_KV]:
        assert x.dim() == 3, "x must be [B,S,dim]"
        B, S, _ = x.shape
        if not x.is_cuda:
            raise RuntimeError("NSAAttention (B200) needs CUDA tensors: the hot path has no CPU fallback")
        if prefill:
            assert S > 0, f"Prefill mode requires S > 0, got S={S}"
            return self._forward_prefill(x, kv)
        assert S == 1, (f"Decode mode requires S=1 (single token), got S={S}. "
                        f"This ensures proper causal ordering in decode steps.")
        return self._forward_decode(x, kv)

    def _phi_weights(self):
        return (self.phi_k_conv.weight, self.phi_v_conv.weight) if self.phi_type == "mlp" else ()

    def _proj_weights(self):
        return (self.W_Q.weight, self.W_K_sel.weight, self.W_V_sel.weight, self.W_K_win.weight, self.W_V_win.weight,
                self.W_K_cmp.weight, self.W_V_cmp.weight)

    def _project(self, x: torch.Tensor, t0: int):
        """The seven projections as ONE GEMM over the stacked weights (cuBLAS through F.linear; the state dict keeps the seven
        nn.Linear parameters), followed by ONE kernel that applies RoPE and writes Q and the six cache-layout tensors
        (ops.project_split) -- the reference's per-tensor rope + view + permute + contiguous chains (nsa_attention.py:998-1016)."""
        ws = self._proj_weights()
        if torch.is_grad_enabled() and any(w.requires_grad for w in ws):
            w_all = torch.cat(ws, dim=0)          # autograd splits the gradient back onto the seven parameters
        else:
            w_all = self._decode_weights()        # inference: the stacked matrix is cached (rebuilt when a weight changes)
        y = F.linear(x, w_all)
        return ops.project_split(y, H=self.n_heads, G=self.n_kv_groups, Dk=self.d_k, Dv=self.d_v, t0=t0, scale=self.rope_scale)

    def _phi(self, K_raw: torch.Tensor, V_raw: torch.Tensor, t0: int = 0):
        """phi over rows at positions t0.. (compress_pool.py:9-38 / nsa_attention.py:1741-1758): one kernel per stream."""
        if self.phi_type == "mlp":
            return ops.phi_conv(K_raw, V_raw, self.phi_k_conv.weight, self.phi_v_conv.weight, self.l, self.d, t0=t0,
                                rope_scale=self.rope_scale)
        return ops.phi_avgpool(K_raw, V_raw, self.l, self.d, t0=t0)

    def _nvtx(self, name: Optional[str]):
        if self._env_cache["nvtx"]:
            if name is None:
                torch.cuda.nvtx.range_pop()
            else:
                torch.cuda.nvtx.range_push(name)

    def _forward_prefill(self, x: torch.Tensor, kv: NSA_KV) -> tuple[torch.Tensor, NSA_KV]:
        """All three reference prefill routes (batched :978-1448, sequential :1521-1723, via-decode :1507-1519)
        collapse onto one fused pass; they differ only in the selection rule and the p_cmp normaliser."""
        B, S, _ = x.shape
        t0 = int(kv.K_sel.shape[2])
        self._nvtx("projections+rope")
        Q, K_sel, V_sel, K_win, V_win, K_raw, V_raw = self._project(x, t0)
        self._nvtx(None)
        # the six streams are fresh outputs of the producer kernel: an empty cache adopts them as its slabs instead of copying
        kv.update_selection_raw(K_sel.detach(), V_sel.detach(), adopt=True)
        kv.meta = build_block_meta(seq_len=t0 + S, l=self.l, d=self.d, l_sel=self.l_sel, n_sel=self.n_sel, w=self.w)
        kv.update_window(K_win.detach(), V_win.detach(), self.w, adopt=True)
        # NOTE: unlike the reference (whose prefill never fills K_cmp_raw_seq, so a following decode restarts the
        # emission count at zero) the raw stream is recorded, keeping "emit every d after warm-up l" absolute.
        kv.append_cmp_raw(K_raw.detach(), V_raw.detach(), adopt=True)
        if t0 == 0:
            K_cmp, V_cmp = self._phi(K_raw, V_raw)
        else:
            K_cmp, V_cmp = self._phi(kv.K_cmp_raw_seq, kv.V_cmp_raw_seq)
        kv.update_compressed(K_cmp.detach(), V_cmp.detach(), self.l, self.d)

        via_decode = self.prefill_tile > 0
        sel_mode = 0 if (self._env_cache["prefill_batched"] and not via_decode) else 1
        cfg = self._cfg(causal_norm=via_decode)
        gate = self.gate.params() if cfg.gate_mode == ops.GATE_MLP else None
        # The reference's ranges "pcmp_all", "map_pcmp_to_pslc", "topk+ranges" and "branch_attn+gate" (nsa_attention.py:1070, :1077,
        # :1103, :1112) cover ONE C-ABI call here (scoring with the Eq.9 / Eq.10 folds, top-n + range merging, the three branches
        # and the gated combine; for long prefill the scorer's second pass even shares a kernel with the compressed branch), so
        # the four names are pushed nested around it.
        for name in ("pcmp_all", "map_pcmp_to_pslc", "topk+ranges", "branch_attn+gate"):
            self._nvtx(name)
        pre_ranges = None
        if self._env_cache["pcmp_mixed"] and Q.dtype == torch.float32:
            # NSA_P_CMP_MIXED (selection_scorer.py:46-56): the scores that drive the selection come from bf16 operands; the three
            # branches still attend in fp32 over the ranges chosen that way
            pre_ranges = ops.score_select(Q.detach().bfloat16(), K_cmp.detach().bfloat16(), cfg, mode=sel_mode, t0=t0)
        if t0 == 0:
            O, ranges, gates = ops.prefill_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate, cfg, sel_mode=sel_mode,
                                                t0=0, stopgrad_gates=self._env_cache["stopgrad_gates"], ranges=pre_ranges,
                                                ranges_trusted=True)
        else:  # chunked prefill continues on the cache slabs (inference only)
            if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
                raise RuntimeError("chunked prefill (kv already holds tokens) attends over the detached cache slabs: gradients "
                                   "would reach Q and the gate only.  Run it under torch.no_grad(), or prefill in one call.")
            n = t0 + S
            O, ranges, gates = ops.prefill_core(
                Q, kv.slab("K_sel"), kv.slab("V_sel"), kv.slab("K_win"), kv.slab("V_win"), K_cmp, V_cmp, gate, cfg,
                sel_mode=sel_mode, t0=t0, S_sel_kv=n, S_win_kv=kv.length("K_win"), win_off=n - kv.length("K_win"), ranges=pre_ranges,
                stopgrad_gates=self._env_cache["stopgrad_gates"], ranges_trusted=True)
        for _ in range(4):
            self._nvtx(None)
        if self._env_cache["strict_asserts"] and ranges.numel() > 0:
            tpos = torch.arange(t0, t0 + S, device=x.device).view(1, S, 1, 1)
            assert bool((ranges[..., 1] <= tpos + 1).all()), "Selection ranges cannot access future tokens."
        self._last_gates, self._last_ranges = gates, ranges
        out = self.out(O.reshape(B, S, self.n_heads * self.d_v))
        return out, kv

    def _decode_weights(self) -> torch.Tensor:
        """The seven projection weights stacked as one [H*Dk + G*(3Dk+3Dv), dim] matrix (rebuilt when any of them changes), so
        that a decode token needs ONE GEMM instead of seven M=B GEMVs."""
        ws = self._proj_weights()
        key = tuple((w.data_ptr(), w._version, w.dtype) for w in ws)
        cached = getattr(self, "_wcat", None)
        if cached is None or cached[0] != key:
            with torch.no_grad():
                cached = (key, torch.cat([w.detach() for w in ws], dim=0).contiguous())
            self._wcat = cached
        return cached[1]

    def _decode_plan(self, y: torch.Tensor, kv: NSA_KV, aux: bool):
        """(plan, rows): the prebuilt argument blocks of this module's decode step on `kv` (ops.DecodeStepPlan), rebuilt whenever a
        cache slab was reallocated or replaced, the batch / dtype changed, or gate / rope settings moved."""
        cfg_key = (float(self.gate_temp), self.gate.mode(), float(self.rope_scale), aux, y.dtype, y.shape[0],
                   tuple((p.data_ptr(), p._version) for p in self.gate.params()))
        plan = getattr(kv, "_decode_plan", None)
        if plan is not None and plan[0] is self and plan[1] == cfg_key:
            rows = kv.decode_state(plan[2].slabs, plan[2].cmp_slabs, plan[2].counters if aux else None)
            if rows is not None:
                return plan[2], rows
        slabs, rows, cmp_slabs, ctr = kv.prepare_decode(y, aux)
        p = self._new_plan(y, slabs, cmp_slabs, ctr)
        kv._decode_plan = (self, cfg_key, p)
        return p, rows

    def _new_plan(self, y, slabs, cmp_slabs, ctr) -> "ops.DecodeStepPlan":
        cfg = self._cfg()
        gate = self.gate.params() if cfg.gate_mode == ops.GATE_MLP else None
        return ops.DecodeStepPlan(y, slabs, cmp_slabs, ctr, H=self.n_heads, G=self.n_kv_groups, Dk=self.d_k, Dv=self.d_v, cfg=cfg,
                                  gate=gate, rope_scale=self.rope_scale)

    # ---- decode as a replayed CUDA graph ----------------------------------------------------------------------------
    def _graph_step(self, x: torch.Tensor, kv: NSA_KV, aux: bool):
        """The captured decode step of this module on `kv` (ops.DecodeGraphStep), or None when the shape has no fused tcgen05 decode
        kernel (fp32, other head dims, capacities beyond the fused scorer): those run the eager step.  Rebuilt (re-captured) when a
        cache slab was reallocated or replaced, the batch / dtype changed, or a weight / gate / rope setting moved."""
        ws = self._proj_weights()
        cfg_key = (float(self.gate_temp), self.gate.mode(), float(self.rope_scale), aux, x.dtype, x.shape[0],
                   tuple((p.data_ptr(), p._version) for p in (*self.gate.params(), *ws, self.out.weight, *self._phi_weights())))
        ent = getattr(kv, "_decode_graph", None)
        if ent is not None and ent[0] is self and ent[1] == cfg_key:
            gs = ent[2]
            if gs is None:
                return None
            if kv.decode_state(gs.slabs, gs.cmp_slabs, gs.counters if aux else None) is not None:
                # an emission needs a free row in the compressed slabs
                if kv.rows_present("K_cmp") < gs.cmp_slabs[0].shape[2]:
                    return gs
        like = torch.empty((x.shape[0], 1), dtype=x.dtype, device=x.device)
        slabs, _, cmp_slabs, ctr = kv.prepare_decode(like, aux)
        # room to run: re-capturing every few steps would cost more than the graph saves
        need = max(int(kv.rows_present("K_sel")) + 256, 2 * self.l_sel)
        if any(s.shape[2] < kv.rows_present(n) + 64 for s, n in zip(slabs, kv._TOKEN_FIELDS)) or cmp_slabs[0].shape[2] < kv.rows_present("K_cmp") + 8:
            kv.reserve(need)
            slabs, _, cmp_slabs, ctr = kv.prepare_decode(like, aux)
        cfg = self._cfg()
        gate = self.gate.params() if cfg.gate_mode == ops.GATE_MLP else None
        hid = int(self.gate.fc1.weight.shape[0]) if gate is not None else 0
        gs = None
        if ops.DecodeGraphStep.supported(x, slabs, cmp_slabs, cfg, H=self.n_heads, G=self.n_kv_groups, Dk=self.d_k, Dv=self.d_v,
                                         gate_hidden=hid):
            gs = ops.DecodeGraphStep(x.reshape(x.shape[0], self.dim), self._decode_weights(), self.out.weight.detach(), slabs, cmp_slabs,
                                     ctr, H=self.n_heads, G=self.n_kv_groups, Dk=self.d_k, Dv=self.d_v, cfg=cfg, gate=gate,
                                     rope_scale=self.rope_scale, phi_w=self._phi_weights() or None)
            st = (kv.rows_present("K_sel"), kv.rows_present("K_win"), kv.rows_present("K_cmp_raw_seq"), kv.rows_present("K_cmp"),
                  kv.counter_column() if aux else 0)
            gs.set_state(*st)
            gs.capture()
            gs.set_state(*st)
        kv._decode_graph = (self, cfg_key, gs)
        return gs

    def _forward_decode_graph(self, x: torch.Tensor, kv: NSA_KV, gs, aux: bool) -> tuple[torch.Tensor, NSA_KV]:
        B = x.shape[0]
        t, row_win, row_raw = kv.rows_present("K_sel"), kv.rows_present("K_win"), kv.rows_present("K_cmp_raw_seq")
        st = (t, row_win, row_raw, kv.rows_present("K_cmp"), kv.counter_column() if aux else 0)
        if gs.expected != st:  # the cache moved without us (a prefill in between, another module): re-seed the device record
            gs.set_state(*st)
        gs.x.copy_(x.reshape(B, self.dim))
        gs.replay()
        # mirror the step in the host bookkeeping while the GPU runs
        S_raw = row_raw + 1
        emitted = S_raw >= self.l and (S_raw - self.l) % self.d == 0
        kv.commit_token_append(self.w)
        if aux:
            kv.commit_counters()
        if emitted:
            kv.commit_compressed_append()
        gs.expected = (t + 1, row_win + 1, row_raw + 1, st[3] + (1 if emitted else 0), st[4] + (1 if aux else 0))
        if not aux:
            gs.expected = gs.expected[:4] + (0,)
            gs.state_host[4] = 0
        if getattr(kv, "meta", None) is None or kv.meta.sel_starts.numel() * self.l_sel < t + 1 or kv.meta.sel_starts.numel() == 0:
            kv.meta = build_block_meta(seq_len=max(t + 1, self.l_sel), l=self.l, d=self.d, l_sel=self.l_sel, n_sel=self.n_sel, w=self.w)
        self._last_ranges = gs.ranges
        return gs.out.clone(), kv

    def _forward_decode(self, x: torch.Tensor, kv: NSA_KV) -> tuple[torch.Tensor, NSA_KV]:
        """One decode step (nsa_attention.py:545-976): one GEMM for the seven projections, one kernel that rotates Q/K and writes
        the token's six cache rows in place (nsa_decode_produce), phi on emission steps, the fused decode kernel, the out GEMM.
        Where the fused tcgen05 decode kernel serves the shape the whole step is ONE replayed CUDA graph driven by a device-side
        step record (ops.DecodeGraphStep); otherwise the two C-ABI calls go through a per-cache DecodeStepPlan (argument blocks
        built once, scalars updated per step)."""
        B = x.shape[0]
        G = self.n_kv_groups
        if self._env_cache["decode_graph"] and not self._env_cache["strict_asserts"] and x.dtype in (torch.bfloat16, torch.float16):
            with torch.no_grad():
                aux = not self._env_cache["disable_aux_stats"]
                gs = self._graph_step(x, kv, aux)
                if gs is not None:
                    return self._forward_decode_graph(x, kv, gs, aux)
        with torch.no_grad():
            y = F.linear(x.reshape(B, self.dim), self._decode_weights())
            aux = not self._env_cache["disable_aux_stats"]
            plan, rows = self._decode_plan(y, kv, aux)
            t = rows[0]            # position of the new token = tokens already in K_sel
            S_raw = rows[4] + 1
            num_cmp = 0 if S_raw < self.l else (S_raw - self.l) // self.d + 1
            n_win_read = min(self.w, S_raw)
            reads = num_cmp + self.n_sel * self.l_sel + n_win_read  # :634-638
            plan.produce(y, t, rows, kv.counter_column() if aux else 0, (reads, reads, self.n_sel * self.l_sel, num_cmp, n_win_read))
            kv.commit_token_append(self.w)
            if aux:
                kv.commit_counters()
            if S_raw >= self.l and (S_raw - self.l) % self.d == 0:  # emission schedule (:587-604)
                K_new, V_new = self._phi(kv.K_cmp_raw_seq[:, :, S_raw - self.l:S_raw], kv.V_cmp_raw_seq[:, :, S_raw - self.l:S_raw],
                                         t0=S_raw - self.l)
                kv.append_compressed(K_new, V_new)
                if not kv.same_compressed_slabs(plan.cmp_slabs):
                    # the emission outgrew the compressed slab: finish this step on the new slabs (the step's Q, already produced
                    # into the old plan's buffer, is carried over) and let the next step build its plan afresh
                    slabs, _, cmp_slabs, _ = kv.prepare_decode(y, False)
                    grown = self._new_plan(y, slabs, cmp_slabs, None)
                    grown.Q.copy_(plan.Q)
                    plan, kv._decode_plan = grown, None
            if getattr(kv, "meta", None) is None or kv.meta.sel_starts.numel() * self.l_sel < t + 1 or kv.meta.sel_starts.numel() == 0:
                kv.meta = build_block_meta(seq_len=max(t + 1, self.l_sel), l=self.l, d=self.d, l_sel=self.l_sel, n_sel=self.n_sel, w=self.w)
            ranges = torch.empty((B, G, self.n_sel, 2), dtype=torch.int32, device=x.device)
            O = plan.attend(t, kv.rows_present("K_win"), kv.rows_present("K_cmp"), ranges)
            if self._env_cache["strict_asserts"]:
                assert int(ranges[..., 1].max()) <= t + 1, "Selection must not access future tokens."
            self._last_ranges = ranges
            out = self.out(O.reshape(B, 1, self.n_heads * self.d_v))
        return out, kv
