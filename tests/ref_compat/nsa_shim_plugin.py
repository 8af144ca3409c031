"""pytest plugin (test infrastructure): makes `import nsa...` inside the REFERENCE's own test files resolve to nsa_vibe_b200.

The reference's hot-path tests build CPU tensors and CPU modules; the product has no CPU path.  The shim therefore wraps every
entry point in an adapter that moves CPU tensors to cuda:0, calls the product (the hand-written kernels through the C ABI) and
moves the results back -- the tests then run UNCHANGED against the drop-in.  Nothing here computes anything itself.

Loaded with `python -m pytest -p nsa_shim_plugin <reference test files>` by tests/ref_compat/test_reference_tests.py.
"""
from __future__ import annotations

import copy
import functools
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEV = "cuda:0"


def _to(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(dev)
    if isinstance(x, (list, tuple)):
        return type(x)(_to(v, dev) for v in x)
    return x


def on_gpu(fn):
    """CPU tensors in -> cuda -> product call -> results back on the caller's device."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        tens = [a for a in list(args) + list(kwargs.values()) if isinstance(a, torch.Tensor)]
        home = tens[0].device if tens else torch.device("cpu")
        out = fn(*[_to(a, DEV) for a in args], **{k: _to(v, DEV) for k, v in kwargs.items()})
        return _to(out, home)
    return wrapper


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # a package: submodules are looked up in sys.modules only
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], leaf, m)
    return m


def install():
    import nsa_vibe_b200 as P
    from nsa_vibe_b200 import kernels as PK
    from nsa_vibe_b200 import ops
    from nsa_vibe_b200.cache import kv_cache as pkv
    from nsa_vibe_b200.core import attention_kernels as pak
    from nsa_vibe_b200.core import block_index as pbi
    from nsa_vibe_b200.core import nsa_attention as pna
    from nsa_vibe_b200.core import packing as ppk
    from nsa_vibe_b200.core import rope as prope
    from nsa_vibe_b200.core import selection_scorer as pss

    _module("nsa")
    _module("nsa.core")
    _module("nsa.cache")
    _module("nsa.kernels")
    _module("nsa.model")
    _module("nsa.core.block_index", **{k: getattr(pbi, k) for k in ("BlockMeta", "build_block_meta", "build_block_starts", "build_M_csl_csr")})
    _module("nsa.core.packing", **{k: getattr(ppk, k) for k in dir(ppk) if k.startswith("compute_")})
    _module("nsa.core.selection_scorer", **{k: on_gpu(getattr(pss, k)) for k in (
        "compute_pcmp_all", "compute_pgrp_all", "map_pcmp_to_pslc", "map_pcmp_to_pslc_batched", "group_reduce_pslc", "select_topn_ranges",
        "select_topn_ranges_batched", "convert_indices_to_ranges_batched", "convert_indices_to_ranges_batched_v2",
        "convert_indices_to_ranges_batched_dispatch", "validate_selection_determinism")})
    _module("nsa.core.attention_kernels", **{k: on_gpu(getattr(pak, k)) for k in (
        "grouped_selection_attention", "grouped_selection_attention_masked", "grouped_selection_attention_packed",
        "selection_attention_varlen_all", "sliding_window_attention", "batched_causal_attention_compressed")})

    def avg_pool_phi_rope_kv(K_raw, V_raw, l, d, pos=None):  # nsa/core/compress_pool.py:9-38
        t0 = int(pos[0]) if pos is not None and len(pos) else 0
        return ops.phi_avgpool(K_raw, V_raw, l, d, t0=t0)
    _module("nsa.core.compress_pool", avg_pool_phi_rope_kv=on_gpu(avg_pool_phi_rope_kv))
    _module("nsa.core.rope", **{k: getattr(prope, k) for k in dir(prope) if not k.startswith("_")})
    _module("nsa.cache.kv_cache", NSA_KV=pkv.NSA_KV)
    _module("nsa.kernels.flash_wrappers", attention_bgh=on_gpu(PK.attention_bgh), fa2_supported=lambda *a, **k: False,
            is_flash_varlen_available=lambda: False, is_flash_available=lambda: False)
    _module("nsa.kernels.triton_sel_kernel", selection_attention_backward_reference=on_gpu(PK.selection_attention_backward_reference),
            selection_attention_triton=on_gpu(PK.selection_attention_triton))
    _module("nsa.kernels.cuda_sel_kernel", selection_attention_cuda=on_gpu(PK.selection_attention_cuda))

    KV_FIELDS = ("K_sel", "V_sel", "K_win", "V_win", "K_cmp_raw_seq", "V_cmp_raw_seq", "K_cmp", "V_cmp", "win_ptr", "cmp_emit_next",
                 "reads_pred", "reads_act_total", "reads_act_sel", "reads_act_cmp", "reads_act_win")

    class NSAAttention(pna.NSAAttention):
        """The product module with a CPU face: parameters live where the test put them (so `nsa.W_K_cmp(x)` on CPU tensors works);
        forward() on CPU inputs runs a cuda twin carrying the same parameters on a device copy of the cache."""

        def _twin(self):
            tw = self.__dict__.get("_shim_twin")
            if tw is None:
                tw = copy.deepcopy(self)
                tw.__class__ = pna.NSAAttention
                tw.to(DEV)
                self.__dict__["_shim_twin"] = tw
            tw.load_state_dict({k: v.to(DEV) for k, v in self.state_dict().items()})
            tw._env_cache, tw.rope_scale, tw.prefill_tile, tw.gate_temp = self._env_cache, self.rope_scale, self.prefill_tile, self.gate_temp
            tw.gate._force_branch, tw.gate._force_uniform_gate = self.gate._force_branch, self.gate._force_uniform_gate
            return tw

        def forward(self, x, kv, *, prefill):
            if x.is_cuda:
                return super().forward(x, kv, prefill=prefill)
            twin = self._twin()
            kd = kv.__dict__.get("_shim_dev")
            if kd is None:
                kd = pkv.NSA_KV(**{f: getattr(kv, f).to(DEV) for f in KV_FIELDS}, meta=kv.meta)
            out, kd = twin(x.to(DEV), kd, prefill=prefill)
            for f in KV_FIELDS:
                setattr(kv, f, getattr(kd, f).cpu())
            kv.meta = kd.meta
            kv.__dict__["_shim_dev"] = kd
            self._last_gates, self._last_ranges = twin._last_gates, twin._last_ranges
            return out.cpu(), kv

    _module("nsa.core.nsa_attention", NSAAttention=NSAAttention, GateMLP=pna.GateMLP)
    from nsa_vibe_b200.model import llama_block_nsa as pblk
    _module("nsa.model.llama_block_nsa", **{k: getattr(pblk, k) for k in ("LlamaBlockNSA", "RMSNorm", "MLP")})
    return P


install()
