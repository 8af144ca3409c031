"""Host-side pipeline for the NSA hot path: pinned host buffers in, pinned host results out.

`PrefillEngine.run` overlaps, across consecutive batches, the host->device copy of batch i+1, the kernels of batch i and
the device->host copy of result i-1 on three CUDA streams (double-buffered device inputs).  Every batch is still copied
in and its result copied out -- nothing is cached between steps -- so the end-to-end rate is bounded by
max(kernels, H2D, D2H) per batch instead of their sum.  PyTorch is plumbing here (streams, events, pinned memory).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import ops

_KEYS = ("Q", "K_sel", "V_sel", "K_win", "V_win", "K_cmp", "V_cmp")


class PrefillEngine:
    def __init__(self, cfg: ops.NSAConfig, gate, device, depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be >= 2 (double buffering)")
        self.cfg, self.gate, self.dev, self.depth = cfg, gate, torch.device(device), depth
        self.s_in = torch.cuda.Stream(device=self.dev)
        self.s_cmp = torch.cuda.Stream(device=self.dev)
        self.s_out = torch.cuda.Stream(device=self.dev)
        self._in: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._out: List[Optional[torch.Tensor]] = [None] * depth

    def _slot_inputs(self, slot: int, host: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        cur = self._in[slot]
        if cur is None or any(cur[k].shape != host[k].shape or cur[k].dtype != host[k].dtype for k in _KEYS):
            cur = {k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.dev) for k in _KEYS}
            self._in[slot] = cur
        return cur

    def step_kernels(self, d: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Scoring + selection + three-branch attention + gated combine for one resident batch."""
        ranges = ops.score_select(d["Q"], d["K_cmp"], self.cfg, mode=0)
        O, _, _ = ops.prefill_core(d["Q"], d["K_sel"], d["V_sel"], d["K_win"], d["V_win"], d["K_cmp"], d["V_cmp"], self.gate,
                                   self.cfg, sel_mode=0, ranges=ranges, ranges_trusted=True)
        return O

    @torch.no_grad()
    def run(self, host_batches: Sequence[Dict[str, torch.Tensor]], host_outs: Sequence[torch.Tensor]) -> None:
        """host_batches[i]: pinned tensors keyed Q,K_sel,...; host_outs[i]: pinned [B,S,G,h,Dv] result buffers."""
        n = len(host_batches)
        main = torch.cuda.current_stream(self.dev)
        start = torch.cuda.Event()
        start.record(main)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_event(start)
        in_done = [torch.cuda.Event() for _ in range(n)]
        cmp_done = [torch.cuda.Event() for _ in range(n)]
        out_done = [torch.cuda.Event() for _ in range(n)]
        for i in range(n):
            slot = i % self.depth
            with torch.cuda.stream(self.s_in):
                if i >= self.depth:
                    self.s_in.wait_event(cmp_done[i - self.depth])  # the kernels that read this slot are done
                d = self._slot_inputs(slot, host_batches[i])
                for k in _KEYS:
                    d[k].copy_(host_batches[i][k], non_blocking=True)
                in_done[i].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(in_done[i])
                if i >= self.depth:
                    self.s_cmp.wait_event(out_done[i - self.depth])  # its result buffer has been drained
                O = self.step_kernels(d)
                if self._out[slot] is None or self._out[slot].shape != O.shape:
                    self._out[slot] = torch.empty_like(O)
                self._out[slot].copy_(O)
                cmp_done[i].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(cmp_done[i])
                host_outs[i].copy_(self._out[slot], non_blocking=True)
                out_done[i].record(self.s_out)
        for s in (self.s_in, self.s_cmp, self.s_out):
            main.wait_stream(s)
