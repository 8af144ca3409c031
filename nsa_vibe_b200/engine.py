"""Host-side pipelines for the NSA hot path: pinned host buffers in, pinned host results out.

`Pipeline.run` overlaps, across consecutive batches, the host->device copy of batch i+1, the kernels of batch i and the
device->host copy of result i-1 on three CUDA streams (double-buffered device inputs).  Every batch is still copied in and
its result copied out -- nothing is cached between steps -- so the end-to-end rate is bounded by max(kernels, H2D, D2H) per
batch instead of their sum.  PyTorch is plumbing here (streams, events, pinned memory).

  PrefillEngine        : the hot path proper, post-projection tensors (Q, six caches) in, O out.
  ModulePrefillEngine  : the reference-facing module call, x [B,S,dim] in, NSAAttention.forward(x, kv, prefill=True) out.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import ops

_KEYS = ("Q", "K_sel", "V_sel", "K_win", "V_win", "K_cmp", "V_cmp")


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that pinned buffers allocated afterwards are
    first-touched on that node and the copy threads run next to it.  Returns the node, or None when the topology does not say
    (single-node hosts report -1) or the call is not permitted."""
    try:
        prop = torch.cuda.get_device_properties(device_index)
        bdf = f"{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


class Pipeline:
    """Three-stream pipeline around `step(device_inputs: dict) -> device tensor`."""

    def __init__(self, step: Callable[[Dict[str, torch.Tensor]], torch.Tensor], keys: Sequence[str], device, depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be >= 2 (double buffering)")
        self.step, self.keys, self.dev, self.depth = step, tuple(keys), torch.device(device), depth
        self.s_in = torch.cuda.Stream(device=self.dev)
        self.s_cmp = torch.cuda.Stream(device=self.dev)
        self.s_out = torch.cuda.Stream(device=self.dev)
        self._in: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._out: List[Optional[torch.Tensor]] = [None] * depth

    def _slot_inputs(self, slot: int, host: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        cur = self._in[slot]
        if cur is None or any(cur[k].shape != host[k].shape or cur[k].dtype != host[k].dtype for k in self.keys):
            cur = {k: torch.empty(host[k].shape, dtype=host[k].dtype, device=self.dev) for k in self.keys}
            self._in[slot] = cur
        return cur

    @torch.no_grad()
    def run(self, host_batches: Sequence[Dict[str, torch.Tensor]], host_outs: Sequence[torch.Tensor]) -> None:
        """host_batches[i]: pinned tensors keyed by self.keys; host_outs[i]: pinned result buffers."""
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
                for k in self.keys:
                    d[k].copy_(host_batches[i][k], non_blocking=True)
                in_done[i].record(self.s_in)
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(in_done[i])
                if i >= self.depth:
                    self.s_cmp.wait_event(out_done[i - self.depth])  # its result buffer has been drained
                O = self.step(d)
                if self._out[slot] is None or self._out[slot].shape != O.shape or self._out[slot].dtype != O.dtype:
                    self._out[slot] = torch.empty_like(O)
                self._out[slot].copy_(O)
                cmp_done[i].record(self.s_cmp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(cmp_done[i])
                host_outs[i].copy_(self._out[slot], non_blocking=True)
                out_done[i].record(self.s_out)
        for s in (self.s_in, self.s_cmp, self.s_out):
            main.wait_stream(s)


class PrefillEngine(Pipeline):
    def __init__(self, cfg: ops.NSAConfig, gate, device, depth: int = 2):
        self.cfg, self.gate = cfg, gate
        super().__init__(self.step_kernels, _KEYS, device, depth)

    def step_kernels(self, d: Dict[str, torch.Tensor]) -> torch.Tensor:
        """Scoring + selection + three-branch attention + gated combine for one resident batch."""
        O, _, _ = ops.prefill_core(d["Q"], d["K_sel"], d["V_sel"], d["K_win"], d["V_win"], d["K_cmp"], d["V_cmp"], self.gate,
                                   self.cfg, sel_mode=0)
        return O


class ModulePrefillEngine(Pipeline):
    """x [B,S,dim] (pinned host) -> NSAAttention.forward(x, fresh NSA_KV, prefill=True) -> out [B,S,dim] (pinned host): the call
    a user of the reference makes (bench/bench_prefill.py:76-85), projections and output projection included."""

    def __init__(self, attn, device, depth: int = 2):
        from .cache.kv_cache import create_empty_kv
        from .core.block_index import build_block_meta
        self.attn, self._mk, self._meta = attn, create_empty_kv, build_block_meta
        super().__init__(self.step_module, ("x",), device, depth)

    def step_module(self, d: Dict[str, torch.Tensor]) -> torch.Tensor:
        a, x = self.attn, d["x"]
        kv = self._mk(x.shape[0], a.n_kv_groups, a.d_k, a.d_v, self._meta(x.shape[1], a.l, a.d, a.l_sel, a.n_sel, a.w),
                      device=x.device, dtype=x.dtype)
        out, _ = a(x, kv, prefill=True)
        return out


class StackPrefillEngine(Pipeline):
    """x [B,S,dim] (pinned host) -> a stack of LlamaBlockNSA layers (RMSNorm, NSAAttention prefill, residual, RMSNorm, SiLU MLP,
    residual; nsa/model/llama_block_nsa.py:33-106) -> hidden states [B,S,dim] (pinned host): ONE host->device copy feeds every
    layer's compute, which is how a model uses the hot path (the per-layer module call moves 200 MB per 64k sequence for ~5 ms of
    kernels and saturates the host side of a multi-GPU box; a 12-layer stack moves the same bytes per ~70 ms)."""

    def __init__(self, blocks, device, depth: int = 2):
        self.blocks = list(blocks)
        super().__init__(self.step_stack, ("x",), device, depth)

    def step_stack(self, d: Dict[str, torch.Tensor]) -> torch.Tensor:
        x, delta = d["x"], None
        for b in self.blocks:  # each block's last residual add runs inside the next norm's kernel
            x, delta = b(x, delta, defer_residual=True)
        return x + delta
