#!/usr/bin/env python3
"""Benchmark of the NSA hot path on B200 (BASELINE.json metric: "NSA prefill tok/s @S=64k, decode us/tok @S=4k;
%roofline; 1/2/4/8 GPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path (scoring -> selection -> three-branch attention -> gated combine,
nsa_attention.py:1066-1398) over one batch of synthetic sequences: S=65536, m7c head dims (H=12, G=2, h=6,
Dk=Dv=64, l=32, d=16, l_sel=64, n_sel=16, w=512), bf16, inputs resident in HBM.  The headline `value` is
tokens/s of that step, whole job (all ranks; the path shards by batch with no collective => weak scaling).
`e2e` is the same step through the public API with HOST (pinned) buffers, copies inside the timed region.
`decode` carries the second half of the metric: us/token of one decode step at S=4096 (batch sharded).
`roofline` is for the dominant kernel of the step; `cpu_baseline` is the oracle port timed on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

M7C = dict(H=12, G=2, h=6, Dk=64, Dv=64, l=32, d=16, l_sel=64, n_sel=16, w=512)
METRIC = "NSA prefill tok/s @S=64k"


# ------------------------------------------------------------------------------------------------------
# algorithmic work, per sequence per layer
# ------------------------------------------------------------------------------------------------------
def num_cmp_at(t, l, d, S_cmp):
    return 0 if t + 1 < l else min((t + 1 - l) // d + 1, S_cmp)


def prefill_flops(S, c=M7C):
    H, Dk, Dv, l, d = c["H"], c["Dk"], c["Dv"], c["l"], c["d"]
    S_cmp = 0 if S < l else (S - l) // d + 1
    sum_cmp = sum(num_cmp_at(t, l, d, S_cmp) for t in range(S))
    sum_sel = sum(min(t + 1, c["n_sel"] * c["l_sel"]) for t in range(S))
    sum_win = sum(min(t + 1, c["w"]) for t in range(S))
    score = 2.0 * S * S_cmp * H * Dk                  # full-row scoring QK^T (reference-prefill parity)
    cmp_pv = 2.0 * H * Dv * sum_cmp                   # cmp P.V (its QK^T is shared with scoring)
    sel = 2.0 * H * (Dk + Dv) * sum_sel
    win = 2.0 * H * (Dk + Dv) * sum_win
    return dict(score=score, cmp_pv=cmp_pv, sel=sel, win=win, total=score + cmp_pv + sel + win,
                sel_gather_bytes=2.0 * c["G"] * (Dk + Dv) * sum_sel)


def scorer_exps(S, c=M7C):
    """Exponentials the split scorer evaluates for one sequence: pass 1 (row max + normaliser) over every compressed key; pass 2
    (probabilities, shared with the compressed branch) per CTA of 4 M-tiles x 4 * (32 // h) tokens up to the larger of the last
    row's selection limit 4 * ((t + 1) // l_sel) and its causal key count num_cmp(t), in whole 64-key tiles (tc_score_cmp.cu)."""
    H, l, d, ls = c["H"], c["l"], c["d"], c["l_sel"]
    S_cmp = 0 if S < l else (S - l) // d + 1
    tok = 4 * 4 * (32 // c["h"])
    p2 = 0
    for s0 in range(0, S, tok):
        s_last = min(S, s0 + tok) - 1
        need = min(S_cmp, (ls // d) * ((s_last + 1) // ls))
        hi = num_cmp_at(s_last, l, d, S_cmp)
        p2 += (min(S, s0 + tok) - s0) * min(S_cmp + 63, -(-max(need, hi) // 64) * 64)
    return float(H) * (S * S_cmp + p2)


def decode_bytes_per_token(S, c=M7C):
    """reads = num_cmp + n_sel*l_sel + min(w,S) (nsa_attention.py:634-638) x G x (Dk+Dv) x 2 B."""
    ncmp = 0 if S < c["l"] else (S - c["l"]) // c["d"] + 1
    reads = ncmp + c["n_sel"] * c["l_sel"] + min(c["w"], S)
    return reads * c["G"] * (c["Dk"] + c["Dv"]) * 2, reads


def num_sel_blocks(S, l_sel):
    return (S + l_sel - 1) // l_sel


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/traffic.json: kernel name -> bytes), or {}."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            j = json.load(f)
        return dict(hbm=j["hbm_gbs"], tf_burst=j["bf16_tflops"], tf_sustained=j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


# ------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def wait_first(self, timeout=1.5):
        t_end = time.time() + timeout
        while self.proc is not None and not self.lines and time.time() < t_end:
            time.sleep(0.01)

    def stop(self, t0=None, t1=None):
        """Median SM clock and throttle reasons over the samples that arrived inside [t0, t1] (the timed region; the sampler is
        started before the warm-up so that nvidia-smi is already streaming when it begins)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [ln for ts, ln in self.lines if (t0 is None or ts >= t0) and (t1 is None or ts <= t1 + 0.06)]
        if not inside:  # a region shorter than one sampling period: take the samples after its start
            inside = [ln for ts, ln in self.lines if t0 is None or ts >= t0] or [ln for _, ln in self.lines[-1:]]
        for ln in inside:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU baselines.  (1) the UNMODIFIED reference package (oracle/_ref, copied by oracle/make_ref.py), timed the way BASELINE.md
# section 3-C4 prescribes for S=64k; (2) the oracle port on a bounded sample of the batched-prefill workload, whose OUTPUT is
# also the parity check of the GPU step at full size (parity_64k).
# ------------------------------------------------------------------------------------------------------
def reference_path():
    p = os.path.join(ROOT, "oracle", "_ref")
    return p if os.path.isdir(os.path.join(p, "nsa")) else None


class ReferenceDecode:
    """The reference's own NSAAttention (m7c dims, fp32, default environment = its stock CPU route) stepping single tokens on an
    NSA_KV that already holds S_ctx tokens of random K/V -- what the reference's NSA_PREFILL_TILE route does per token
    (nsa_attention.py:1507-1519), and the only way the reference can run a 64k context (its batched route needs O(S^2) masks
    and a 12.9 GB p_cmp tensor).  Contexts are spread evenly over the sequence, so tokens / time estimates whole-sequence
    prefill throughput."""

    def __init__(self, S_max, seed=0):
        ref = reference_path()
        if ref not in sys.path:
            sys.path.insert(0, ref)
        from nsa.cache.kv_cache import NSA_KV
        from nsa.core.block_index import build_block_meta
        from nsa.core.nsa_attention import NSAAttention
        c = M7C
        self.NSA_KV, self.c, self.S_max = NSA_KV, c, S_max
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        torch.manual_seed(seed)
        self.attn = NSAAttention(dim=768, n_heads=c["H"], n_kv_groups=c["G"], d_k=c["Dk"], d_v=c["Dv"], l=c["l"], d=c["d"],
                                 l_sel=c["l_sel"], n_sel=c["n_sel"], w=c["w"])
        self.meta = build_block_meta(S_max + c["l_sel"], c["l"], c["d"], c["l_sel"], n_sel=c["n_sel"], w=c["w"])
        g = torch.Generator().manual_seed(seed)
        r = lambda n, D: torch.randn(1, c["G"], n, D, generator=g)
        self.K, self.V = r(S_max, c["Dk"]), r(S_max, c["Dv"])
        self.Kc, self.Vc = r((S_max - c["l"]) // c["d"] + 1, c["Dk"]), r((S_max - c["l"]) // c["d"] + 1, c["Dv"])
        self.g = g

    def kv(self, S):
        c = self.c
        nc = 0 if S < c["l"] else (S - c["l"]) // c["d"] + 1
        z = lambda: torch.zeros((0,), dtype=torch.int64)
        cut = lambda t, n: t[:, :, :n].contiguous()
        lo = max(0, S - c["w"])
        return self.NSA_KV(K_sel=cut(self.K, S), V_sel=cut(self.V, S), K_win=self.K[:, :, lo:S].contiguous(),
                           V_win=self.V[:, :, lo:S].contiguous(), K_cmp_raw_seq=cut(self.K, S), V_cmp_raw_seq=cut(self.V, S),
                           K_cmp=cut(self.Kc, nc), V_cmp=cut(self.Vc, nc), win_ptr=torch.zeros((1, c["G"]), dtype=torch.int32),
                           cmp_emit_next=torch.zeros((1, c["G"]), dtype=torch.int32), reads_pred=z(), reads_act_total=z(),
                           reads_act_sel=z(), reads_act_cmp=z(), reads_act_win=z(), meta=self.meta)

    def steps(self, S_ctx, reps):
        """`reps` decode steps starting from a context of S_ctx tokens; returns seconds (cache construction not timed)."""
        kv = self.kv(S_ctx)
        xs = torch.randn(reps, 1, 1, 768, generator=self.g)
        with torch.no_grad():
            t = time.perf_counter()
            for i in range(reps):
                _, kv = self.attn(xs[i], kv, prefill=False)
            return time.perf_counter() - t


def reference_decode_sample(S, strata, reps, seed=0, rd=None):
    """(tok/s, cores, sample text, ms per token at the last stratum): reps decode steps at each of `strata` contexts spread evenly
    over an S-token sequence (the last one is S - reps: the 65,535-token cache of BASELINE.md section 3-C4)."""
    rd = rd or ReferenceDecode(S, seed)
    rd.steps(min(S - 4, 4096), 2)  # warm-up (first call pays one-off costs: 83 ms vs 35 ms measured)
    tot_t, tot_n, last = 0.0, 0, 0.0
    for i in range(strata):
        ctx = min(S - reps, max(M7C["l"], int((i + 1) / strata * S) - reps))
        dt = rd.steps(ctx, reps)
        tot_t += dt
        tot_n += reps
        last = dt / reps
    return (tot_n / tot_t, rd.cores,
            f"UNMODIFIED reference (oracle/_ref), NSAAttention.forward(prefill=False), fp32, default env: {reps} single-token steps "
            f"at each of {strata} contexts spread evenly up to {S - reps} cached tokens of one S={S} sequence (B=1, m7c dims; "
            f"BASELINE.md section 3-C4)", last * 1e3)


def cpu_prefill_sample(S, rows, seed=0, budget_s=25.0, tensors=None, gate=None):
    """Time the oracle's hot path (batched-prefill semantics) for `rows` consecutive query rows in the middle of an S-token
    sequence (B=1) against the full caches.  tensors: the GPU step's own inputs as fp32 CPU tensors (then the output is
    returned for the parity check).  Returns (tok_per_s, cores, sample description, oracle output dict, t0)."""
    from oracle import nsa_oracle as O
    c = M7C
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    S_cmp = O.num_cmp_blocks(S, c["l"], c["d"])
    bf = lambda *s: torch.randn(*s, generator=g).bfloat16().float()
    t0 = S // 2
    if tensors is None:
        Q = bf(1, rows, c["G"], c["h"], c["Dk"])
        K_sel, V_sel, K_win, V_win = bf(1, c["G"], S, c["Dk"]), bf(1, c["G"], S, c["Dv"]), bf(1, c["G"], S, c["Dk"]), bf(1, c["G"], S, c["Dv"])
        K_cmp, V_cmp = bf(1, c["G"], S_cmp, c["Dk"]), bf(1, c["G"], S_cmp, c["Dv"])
        hid = c["Dk"] // 2
        gate = (torch.randn(hid, c["Dk"], generator=g) * 0.1, torch.zeros(hid), torch.randn(3, hid, generator=g) * 0.1, torch.zeros(3))
    else:
        Q = tensors["Q"][:1, t0:t0 + rows]
        K_sel, V_sel, K_win, V_win, K_cmp, V_cmp = (tensors[k][:1] for k in ("K_sel", "V_sel", "K_win", "V_win", "K_cmp", "V_cmp"))
    kw = dict(l=c["l"], d=c["d"], l_sel=c["l_sel"], n_sel=c["n_sel"], w=c["w"], t0=t0, S_total=S)
    with torch.no_grad():
        O.prefill_core(Q[:, :4], K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate, **kw)  # warm-up
        n, el, out = 0, 0.0, None
        while True:
            t = time.perf_counter()
            out = O.prefill_core(Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate, **kw)
            el += time.perf_counter() - t
            n += 1
            if el > budget_s or n >= 3:
                break
    return (rows * n / el, cores, f"{rows} query rows at t0={t0} of one S={S} sequence (B=1), full K/V caches, {n} repeats, fp32 oracle port, "
            "batched-prefill semantics", out, t0)


def parity_64k(ora, t0, rows, gpu):
    """GPU step vs the oracle on the same inputs at full size.  gpu: dict of the step's device outputs (O, ranges, O_cmp, O_sel,
    O_win).  Outputs are compared on the rows where both selected the same ranges (a different selection is a different
    function); those rows are counted, and for each the fp32 gap between the last kept and the first dropped candidate score is
    reported: a selection that differs from the fp32 oracle's may only do so at a near-tie."""
    from oracle import nsa_oracle as O
    sl = slice(t0, t0 + rows)
    rg = gpu["ranges"][:1, sl].cpu()
    same = torch.tensor([[O.nonempty_ranges(rg[0, s, g].tolist()) == O.nonempty_ranges(ora["ranges"][0, s, g].tolist())
                          for g in range(rg.shape[2])] for s in range(rows)])
    out = {"rows": rows * rg.shape[2], "t0": t0, "rows_with_different_ranges": int((~same).sum())}
    gaps = []
    if ora.get("p_grp") is not None:
        c = M7C
        for s, g in (~same).nonzero().tolist():
            t = t0 + s
            p = ora["p_grp"][0, s, g].clone()
            nvalid = (t + 1) // c["l_sel"]
            cb = t // c["l_sel"]
            comp = p.float() - torch.arange(p.numel(), dtype=torch.float32) * torch.tensor(1e-8, dtype=torch.float32)
            comp[nvalid:] = float("-inf")
            for j in {0, cb, max(cb - 1, 0)}:
                if j < comp.numel():
                    comp[j] = float("-inf")
            v = torch.sort(comp, descending=True).values
            k = c["n_sel"] - 3
            gaps.append(float(v[k - 1] - v[k]) if v.numel() > k else float("nan"))
        out["max_score_gap_at_cut_of_differing_rows"] = max(gaps) if gaps else 0.0
    m = same[:, :, None, None].expand(-1, -1, M7C["h"], M7C["Dv"])
    for k in ("O", "O_sel", "O_cmp", "O_win"):
        if gpu.get(k) is None:
            continue
        d = (gpu[k][:1, sl].float().cpu()[0] - ora[k][0]).abs()
        if k in ("O", "O_sel"):
            d = d[m]
        out[k] = {"max_abs": float(d.max()) if d.numel() else 0.0, "mae": float(d.mean()) if d.numel() else 0.0}
    out["tolerance"] = "bf16 kernels vs the fp32 oracle: max_abs <= 2e-2, mae <= 1e-3 per branch and for the gated O"
    out["ok"] = all(out[k]["max_abs"] <= 2e-2 and out[k]["mae"] <= 1e-3 for k in ("O", "O_sel", "O_cmp", "O_win") if k in out)
    return out


def reference_gpu_block(dev):
    """The reference's best library route on THIS GPU (scripts/run_m7c_1xa100_production.sh:22-31: batched prefill, masked SDPA
    selection): NSAAttention.forward(prefill=True), m7c dims, B=1.  It cannot run at 64k (O(S^2) masks); S=2048 / 4096 are
    what fits.  Apples-to-apples with `module_prefill` of our own arm."""
    out = {"route": "NSA_PREFILL_BATCHED=1 NSA_USE_SEL_MASK=1 NSA_FORCE_SEL_MASK=1 NSA_USE_FA2=0 (run_m7c_1xa100_production.sh)"}
    try:
        ref = reference_path()
        if ref not in sys.path:
            sys.path.insert(0, ref)
        for k, v in dict(NSA_PREFILL_BATCHED="1", NSA_USE_SEL_MASK="1", NSA_FORCE_SEL_MASK="1", NSA_USE_FA2="0", NSA_USE_SEL_VARLEN="0",
                         NSA_USE_TRITON_SEL="0", NSA_USE_SEL_PACK="0").items():
            os.environ[k] = v
        from nsa.cache.kv_cache import NSA_KV
        from nsa.core.block_index import build_block_meta
        from nsa.core.nsa_attention import NSAAttention
        c = M7C
        torch.manual_seed(0)
        for dtype in (torch.bfloat16, torch.float32):
            try:
                attn = NSAAttention(dim=768, n_heads=c["H"], n_kv_groups=c["G"], d_k=c["Dk"], d_v=c["Dv"], l=c["l"], d=c["d"],
                                    l_sel=c["l_sel"], n_sel=c["n_sel"], w=c["w"]).to(dev).to(dtype)
                res = {}
                for S in (2048, 4096):
                    meta = build_block_meta(S, c["l"], c["d"], c["l_sel"], n_sel=c["n_sel"], w=c["w"])
                    x = torch.randn(1, S, 768, device=dev, dtype=dtype)

                    def kv0():
                        z = lambda D: torch.zeros((1, c["G"], 0, D), device=dev, dtype=dtype)
                        zi = lambda: torch.zeros((0,), dtype=torch.int64, device=dev)
                        return NSA_KV(K_sel=z(c["Dk"]), V_sel=z(c["Dv"]), K_win=z(c["Dk"]), V_win=z(c["Dv"]), K_cmp_raw_seq=z(c["Dk"]),
                                      V_cmp_raw_seq=z(c["Dv"]), K_cmp=z(c["Dk"]), V_cmp=z(c["Dv"]),
                                      win_ptr=torch.zeros((1, c["G"]), dtype=torch.int32, device=dev),
                                      cmp_emit_next=torch.zeros((1, c["G"]), dtype=torch.int32, device=dev), reads_pred=zi(),
                                      reads_act_total=zi(), reads_act_sel=zi(), reads_act_cmp=zi(), reads_act_win=zi(), meta=meta)
                    with torch.no_grad():
                        for _ in range(2):
                            attn(x, kv0(), prefill=True)
                        torch.cuda.synchronize()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        n = 5
                        a.record()
                        for _ in range(n):
                            attn(x, kv0(), prefill=True)
                        b.record()
                        torch.cuda.synchronize()
                    ms = a.elapsed_time(b) / n
                    res[f"S={S}"] = {"ms_per_call": ms, "tok_per_s": S / (ms * 1e-3)}
                out.update(dtype=str(dtype).replace("torch.", ""), results=res)
                break
            except Exception as ex:  # e.g. a dtype the reference's route does not take: try the next one
                out[f"error_{str(dtype).replace('torch.', '')}"] = f"{type(ex).__name__}: {str(ex)[:200]}"
        out["note_64k"] = "not runnable: the batched route materialises O(S^2) masks and a 12.9 GB p_cmp tensor at S=65536"
    except Exception as ex:
        out["error"] = f"{type(ex).__name__}: {str(ex)[:300]}"
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same metric and config.  With
    oracle/_ref present (copied by oracle/make_ref.py; it travels with the snapshot) that is the UNMODIFIED reference package
    (kind "reference"); otherwise the oracle port (kind "port").  Rank 0 only."""
    if rank != 0:
        return
    extra = {}
    if reference_path() is not None:
        rd = ReferenceDecode(args.S)
        rd.steps(min(args.S - 4, 4096), 2)
        reps, n_steps = 4, args.warmup + args.steps
        tot_t, tot_n, per = 0.0, 0, []
        for i in range(n_steps):  # step i = `reps` single-token steps at a context from the i-th of K evenly spread positions
            k = max(i - args.warmup, 0)
            ctx = min(args.S - reps, max(M7C["l"], int((k + 1) / max(args.steps, 1) * args.S) - reps))
            dt = rd.steps(ctx, reps)
            if i >= args.warmup:
                tot_t += dt
                tot_n += reps
                per.append(dt)
        val, cores, kind = tot_n / tot_t, rd.cores, "reference"
        ms_step = 1e3 * tot_t / len(per)
        sample = (f"UNMODIFIED reference (oracle/_ref) NSAAttention.forward(prefill=False), fp32, default env: each step = {reps} "
                  f"single-token steps on an NSA_KV holding random K/V, step k at context ~ (k+1)/{args.steps} * {args.S} "
                  f"(the last one on {args.S - reps} cached tokens, BASELINE.md 3-C4); tok/s = tokens / time over the {args.steps} steps")
        extra["ms_per_token_at_64k_context"] = 1e3 * per[-1] / reps
        if torch.cuda.is_available():
            extra["reference_gpu"] = reference_gpu_block(torch.device("cuda", 0))
    else:
        per_step = []
        rows = max(32, args.cpu_rows // 2)  # ~1.5 s per step on 16 cores: a 40-step run ends within a minute
        for i in range(args.warmup + args.steps):
            v, cores, sample, _, _ = cpu_prefill_sample(args.S, rows, seed=i, budget_s=0.0)
            if i >= args.warmup:
                per_step.append(v)
        val, kind = sum(per_step) / len(per_step), "port"
        ms_step = 1e3 * rows / val
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "tok/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": {"value": val, "unit": "tok/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0, **extra}
    print(json.dumps(line))


def workload_config(args, world):
    return {"workload": f"NSA hot path prefill S={args.S}, m7c head dims (H=12,G=2,h=6,Dk=Dv=64,l=32,d=16,l_sel=64,n_sel=16,w=512), "
                        f"B={args.B} sequence(s) per GPU, one layer", "S": args.S, "batch_per_gpu": args.B, "global_batch": args.B * world,
            "selection": "batched-prefill rule, full-row p_cmp normaliser (reference prefill parity)",
            "l2": "inputs+outputs exceed the 126 MB L2 (no flush needed)", "parallelism": f"batch-sharded x{world}, no collective"}


# ------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------
def make_inputs(B, S, dev, seed):
    c = M7C
    g = torch.Generator(device=dev).manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    S_cmp = 0 if S < c["l"] else (S - c["l"]) // c["d"] + 1
    t = dict(Q=r(B, S, c["G"], c["h"], c["Dk"]), K_sel=r(B, c["G"], S, c["Dk"]), V_sel=r(B, c["G"], S, c["Dv"]),
             K_win=r(B, c["G"], S, c["Dk"]), V_win=r(B, c["G"], S, c["Dv"]), K_cmp=r(B, c["G"], S_cmp, c["Dk"]),
             V_cmp=r(B, c["G"], S_cmp, c["Dv"]))
    hid = c["Dk"] // 2
    gate = (torch.randn(hid, c["Dk"], generator=g, device=dev) * 0.1, torch.zeros(hid, device=dev),
            torch.randn(3, hid, generator=g, device=dev) * 0.1, torch.zeros(3, device=dev))
    return t, gate


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--S", type=int, default=65536)
    ap.add_argument("--B", type=int, default=1, help="sequences per GPU in the prefill step")
    ap.add_argument("--decode-S", type=int, default=4096)
    ap.add_argument("--decode-B", type=int, default=592, help="sequences per GPU in the decode step (592 x G=2 = 4 (b,g) rows per resident CTA: 2 CTAs x 148 SMs)")
    ap.add_argument("--cpu-rows", type=int, default=512, help="query rows of the CPU sample (cpu_baseline leg; --impl reference uses half)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the forward+backward block (training shape, config C5)")
    ap.add_argument("--no-train-ddp", action="store_true", help="skip the 12-layer DDP training step (config C5)")
    ap.add_argument("--no-stack", action="store_true", help="skip the end-to-end run through a 12-layer block stack")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch.distributed as dist
    from nsa_vibe_b200 import _lib, ops

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from nsa_vibe_b200.engine import ModulePrefillEngine, PrefillEngine, bind_to_gpu_numa_node
    numa_node = bind_to_gpu_numa_node(local) if world > 1 else None  # pinned e2e buffers are first-touched after this
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    c = M7C
    cfg = ops.NSAConfig(l=c["l"], d=c["d"], l_sel=c["l_sel"], n_sel=c["n_sel"], w=c["w"])
    S, B = args.S, args.B
    args.warmup = max(int(args.warmup), 3)  # timing rule: at least three untimed steps; the JSON reports the count actually run
    inp, gate = make_inputs(B, S, dev, seed=1234 + rank)

    def step(t):
        # ONE C-ABI call (nsa_prefill_full_fwd): scoring, selection, three branches, gated combine
        O, _, _ = ops.prefill_core(t["Q"], t["K_sel"], t["V_sel"], t["K_win"], t["V_win"], t["K_cmp"], t["V_cmp"], gate, cfg, sel_mode=0)
        return O

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            step(inp)
        if rank == 0:
            sampler.wait_first()
        barrier()
        t_wall0 = time.time()
        n0 = lib.nsa_kernel_launches()
        e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_start.record()
        for i in range(args.steps):
            step(inp)
        e_end.record()
        barrier()
        launches = lib.nsa_kernel_launches() - n0
        clocks = sampler.stop(t_wall0, time.time()) if rank == 0 else None
        ms_total = e_start.elapsed_time(e_end)
        tt = torch.tensor([ms_total], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        ms_step = ms_total / args.steps
        value = world * B * S / (ms_step * 1e-3)

        # ---- e2e: host buffers in, host result out, copies inside the timed region ------------------------
        # Public API = the reference-facing module call: x [B,S,768] (pinned host) -> NSAAttention.forward(x, kv, prefill=True)
        # -> out [B,S,768] (pinned host), projections and output projection included (bench/bench_prefill.py:76-85).  Every step
        # copies its input in and its result back; the copies of neighbouring steps overlap the kernels on separate streams
        # (nothing is cached between steps).  `hot_path_only` is the same pipeline around the hot path alone (post-projection
        # Q / K / V tensors in, O out: 270 MB of host traffic per step).
        from nsa_vibe_b200.core.nsa_attention import NSAAttention
        n_e2e = max(4, args.steps)  # the same K steps as the device-resident figure (pipeline fill and drain included)
        os.environ["NSA_PREFILL_BATCHED"] = "1"  # the batched-prefill selection rule of the headline (flags are read at construction)
        torch.manual_seed(11 + rank)
        attn = NSAAttention(dim=768, n_heads=c["H"], n_kv_groups=c["G"], d_k=c["Dk"], d_v=c["Dv"], l=c["l"], d=c["d"], l_sel=c["l_sel"],
                            n_sel=c["n_sel"], w=c["w"]).to(dev).bfloat16()
        x_host = [torch.randn(B, S, 768).bfloat16().pin_memory() for _ in range(2)]
        y_host = [torch.empty((B, S, 768), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
        h2d = x_host[0].numel() * 2
        d2h = y_host[0].numel() * 2
        meng = ModulePrefillEngine(attn, dev)
        meng.run([{"x": x_host[i % 2]} for i in range(3)], [y_host[i % 2] for i in range(3)])  # warm-up
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        meng.run([{"x": x_host[i % 2]} for i in range(n_e2e)], [y_host[i % 2] for i in range(n_e2e)])
        e2.record()
        barrier()
        t2 = torch.tensor([s2.elapsed_time(e2)], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        e2e_val = world * B * S / (float(t2.item()) / n_e2e * 1e-3)
        xd = x_host[(n_e2e - 1) % 2].to(dev)
        e2e_ok = bool(torch.equal(y_host[(n_e2e - 1) % 2].to(dev), meng.step_module({"x": xd})))  # the pipelined result is the plain result
        # device-resident module call (no copies): what the copies cost on top
        ms_mod = None
        try:
            for _ in range(2):
                meng.step_module({"x": xd})
            barrier()
            s3, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s3.record()
            for _ in range(5):
                meng.step_module({"x": xd})
            e3.record()
            barrier()
            ms_mod = s3.elapsed_time(e3) / 5
        except Exception:
            pass
        # the reference's GPU library route runs at S=2048 / 4096 only: the same module call at those sizes for the apples-to-apples line
        module_small = {}
        for S_small in (2048, 4096):
            xs = torch.randn(1, S_small, 768, device=dev).bfloat16()
            for _ in range(3):
                meng.step_module({"x": xs})
            torch.cuda.synchronize()
            s4, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s4.record()
            for _ in range(10):
                meng.step_module({"x": xs})
            e4.record()
            torch.cuda.synchronize()
            module_small[f"S={S_small}"] = {"ms_per_call": s4.elapsed_time(e4) / 10, "tok_per_s": S_small / (s4.elapsed_time(e4) / 10 * 1e-3)}
        del x_host, y_host, meng, attn, xd
        # hot path alone through the same pipeline (round-1 definition of e2e)
        host = {k: v.cpu().pin_memory() for k, v in inp.items()}
        o_host = [torch.empty((B, S, c["G"], c["h"], c["Dv"]), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
        hp_h2d = sum(v.numel() * v.element_size() for v in host.values())
        eng = PrefillEngine(cfg, gate, dev)
        eng.run([host] * 3, [o_host[i % 2] for i in range(3)])  # warm-up
        barrier()
        s5, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s5.record()
        eng.run([host] * n_e2e, [o_host[i % 2] for i in range(n_e2e)])
        e5.record()
        barrier()
        t5 = torch.tensor([s5.elapsed_time(e5)], device=dev)
        if world > 1:
            dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        hp_val = world * B * S / (float(t5.item()) / n_e2e * 1e-3)
        hp_ok = bool(torch.equal(o_host[(n_e2e - 1) % 2].to(dev), step(inp)))
        del host, o_host, eng
        # ---- e2e through a 12-layer LlamaBlockNSA stack: one H2D of x feeds twelve layers of compute ----------------------------
        stack = None
        if not args.no_stack:
            try:
                stack = bench_stack_e2e(args, dev, rank, world, barrier)
            except Exception as ex:
                stack = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}

        # ---- per-kernel device times (outside the timed region; same inputs) -------------------------------
        def t_of(fn, n=3):
            fn()
            torch.cuda.synchronize()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            for _ in range(n):
                fn()
            b_.record()
            torch.cuda.synchronize()
            return a_.elapsed_time(b_) / n
        pg = ops.score_pgrp(inp["Q"], inp["K_cmp"], cfg)
        rg_ = ops.select_ranges_prefill(pg, c["l_sel"], c["n_sel"], S)
        # which selected-branch kernel does nsa_prefill_fwd use for this shape?  (its workspace then includes the index + partials)
        import ctypes as _C
        dm_ = ops.make_dims(inp["Q"], cfg, K_sel=inp["K_sel"], K_win=inp["K_win"], K_cmp=inp["K_cmp"], V=inp["V_sel"], n_ranges=rg_.shape[3], gate_hidden=c["Dk"] // 2)
        staging_ = (3 * inp["Q"].numel() * 2 + 255) // 256 * 256
        sel_blockmajor = int(lib.nsa_workspace_bytes(_C.byref(dm_), _lib.WS_PREFILL)) > staging_
        # The step (nsa_prefill_full_fwd) runs: scorer pass 1 (row statistics) -> pass 2 fused with the compressed branch -> selection
        # -> selected branch (block-major: index, attention, merge) -> sliding branch -> gate + combine.  Each is timed in place
        # through its stand-alone entry point on the same inputs.  "score_select_unfused" / "cmp_unfused" are the kernels the fusion
        # replaced (stand-alone scorer with both passes; dense compressed branch), "score_full_pgrp" the scorer asked for all of p_grp.
        fused_scorer = True
        try:
            st_ = ops.score_stats(inp["Q"], inp["K_cmp"], cfg)
            t_p1 = t_of(lambda: ops.score_stats(inp["Q"], inp["K_cmp"], cfg))
            t_p2 = t_of(lambda: ops.score_cmp(inp["Q"], inp["K_cmp"], inp["V_cmp"], cfg, st_))
            del st_
        except RuntimeError:
            fused_scorer = False
        t_ss = t_of(lambda: ops.score_select(inp["Q"], inp["K_cmp"], cfg, mode=0))
        t_sel = t_of(lambda: ops.select_ranges_prefill(pg, c["l_sel"], c["n_sel"], S))
        t_full = t_of(lambda: ops.score_pgrp(inp["Q"], inp["K_cmp"], cfg))
        t_cmp = t_of(lambda: ops.branch_attention(ops.BR_CMP, inp["Q"], inp["K_cmp"], inp["V_cmp"], cfg))
        kms = ({"score_pass1": t_p1, "score_pass2+cmp": t_p2} if fused_scorer else {"score": max(0.0, t_ss - t_sel), "cmp": t_cmp})
        kms.update({"select": t_sel,
                    "sel": t_of((lambda: ops.sel_attention_blockmajor(inp["Q"], inp["K_sel"], inp["V_sel"], cfg, rg_, ranges_trusted=True)) if sel_blockmajor else
                                (lambda: ops.branch_attention(ops.BR_SEL, inp["Q"], inp["K_sel"], inp["V_sel"], cfg, rg_, ranges_trusted=True))),
                    "win": t_of(lambda: ops.branch_attention(ops.BR_WIN, inp["Q"], inp["K_win"], inp["V_win"], cfg))})
        # The parts are timed stand-alone; inside the step the sliding branch and the GateMLP run on a side stream next to the
        # scorer, and the gated blend is folded into the selected branch's merge, so the parts add up to about the step and this
        # residual is what is left of the combine (round 1: a separate 0.11-0.14 ms pass).
        kms["gate_combine_and_rest"] = max(0.0, ms_step - sum(kms.values()))
        kms["note"] = "parts timed stand-alone; in the step win + gate overlap the scorer on a side stream and the blend is folded into the merge"
        kms["score_full_pgrp"] = t_full  # not part of the step
        kms["score_select_unfused"] = t_ss  # not part of the step
        kms["cmp_unfused"] = t_cmp  # not part of the step
        del pg, rg_
        # ---- the step's outputs for the rows the CPU leg recomputes (parity at full size) ------------------------
        gpu_rows = None
        if rank == 0 and not args.no_cpu:
            t0p = S // 2
            O_step, rg_step, _ = ops.prefill_core(inp["Q"], inp["K_sel"], inp["V_sel"], inp["K_win"], inp["V_win"], inp["K_cmp"], inp["V_cmp"],
                                                  gate, cfg, sel_mode=0)
            sl = slice(0, t0p + args.cpu_rows)
            if fused_scorer:  # O_cmp as the step computes it (the fused kernel), not the stand-alone dense kernel
                o_cmp_step = ops.score_cmp(inp["Q"], inp["K_cmp"], inp["V_cmp"], cfg, ops.score_stats(inp["Q"], inp["K_cmp"], cfg))[1]
            else:
                o_cmp_step = ops.branch_attention(ops.BR_CMP, inp["Q"], inp["K_cmp"], inp["V_cmp"], cfg)
            gpu_rows = {"ranges": rg_step[:1, sl].clone(), "O": O_step[:1, sl].clone(),
                        "O_cmp": o_cmp_step[:1, sl].clone(),
                        "O_win": ops.branch_attention(ops.BR_WIN, inp["Q"], inp["K_win"], inp["V_win"], cfg)[:1, sl].clone(),
                        "O_sel": (ops.sel_attention_blockmajor(inp["Q"], inp["K_sel"], inp["V_sel"], cfg, rg_step, ranges_trusted=True)
                                  if sel_blockmajor else
                                  ops.branch_attention(ops.BR_SEL, inp["Q"], inp["K_sel"], inp["V_sel"], cfg, rg_step, ranges_trusted=True))[:1, sl].clone()}
            del rg_step, O_step, o_cmp_step

        # ---- decode @S=4096 -------------------------------------------------------------------------------
        decode = None
        if not args.no_decode:
            decode = bench_decode(args, ops, cfg, dev, rank, world, barrier)
    train = None
    if not args.no_train:
        train = bench_train_core(ops, cfg, dev, rank, world, barrier)
    train_ddp = None
    if not args.no_train_ddp:
        try:
            train_ddp = bench_train_ddp(dev, rank, world, barrier)
        except Exception as ex:  # additional evidence: never lose the headline over it
            train_ddp = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}

    if rank != 0:
        _leave(world, train_ddp)
        return
    # ---- roofline of the dominant kernel -------------------------------------------------------------------
    pk = measured_peaks()
    fl = prefill_flops(S)
    traffic = load_traffic()
    # (name, ms, bound, algorithmic work per launch [FLOP or bytes], how the work is counted)
    if "score_pass1" in kms:
        # pass 1 carries the QK^T of every (row, key); pass 2 recomputes it up to the causal limit and adds the branch's P.V
        cand = [
            ("score_tc_kernel[pass 1: row statistics]", kms["score_pass1"], "tensor", B * fl["score"],
             "2*S*S_cmp*H*Dk (full-row scoring QK^T; pass 2's recomputation is not counted as work)"),
            ("score_cmp_tc_kernel[pass 2 + compressed branch]", kms["score_pass2+cmp"], "tensor", B * fl["cmp_pv"],
             "2*H*Dv*sum_t num_cmp(t) (the branch's P.V; the QK^T it shares with scoring is counted under pass 1)"),
        ]
    else:
        cand = [
            ("score_tc_kernel", kms["score"], "tensor", B * fl["score"], "2*S*S_cmp*H*Dk (full-row scoring QK^T)"),
            ("dense_attn_tc_kernel[cmp]", kms["cmp"], "tensor", B * fl["cmp_pv"], "2*H*Dv*sum_t num_cmp(t) (P.V; its QK^T is counted under scoring)"),
        ]
    cand += [
        ("select_kernel", kms["select"], "hbm", B * S * c["G"] * (4.0 * num_sel_blocks(S, c["l_sel"]) + 8.0 * c["n_sel"]),
         "4*S_sel B read + 8*n_sel B written per (b,t,g) row"),
        ("sel2_attn_kernel+index+merge[sel, KV-block-major]" if sel_blockmajor else "gather_attn_tc_kernel[sel]", kms["sel"], "tensor",
         B * fl["sel"], "2*H*(Dk+Dv)*sum_t min(t+1, n_sel*l_sel) (QK^T + PV of the selected keys)"),
        ("dense_attn_tc_kernel[win]", kms["win"], "tensor", B * fl["win"], "2*H*(Dk+Dv)*sum_t min(t+1, w)"),
    ]
    ms_scorer = kms["score_pass1"] + kms["score_pass2+cmp"] if "score_pass1" in kms else kms["score"]
    # dominant KERNEL: the selected branch's entry is a group of five launches (index x3, attention, merge; split in the committed
    # launch list profiles/r2_launches_prefill_bench_v2.csv: 0.12 + 0.68 + 0.40 ms), so it is listed in per_kernel but the headline
    # roofline is that of the longest single kernel
    singles = [x for x in cand if "+index+merge" not in x[0]]
    kname, k_ms, bound, work, how = max(singles, key=lambda x: x[1])
    if bound == "tensor":
        achieved, peak, unit, src = work / (k_ms * 1e-3) / 1e12, pk["tf_sustained"], "TFLOP/s", pk["src"] + " (bf16 sustained)"
    else:
        achieved, peak, unit, src = work / (k_ms * 1e-3) / 1e9, pk["hbm"], "GB/s", pk["src"] + " (HBM copy)"
    roofline = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                "traffic": traffic.get(kname), "kernel": kname, "kernel_ms": k_ms, "algorithmic_work": how, "peak_source": src,
                "exp_bound": {"what": "MUFU ex2 throughput also bounds the scorer: S*S_cmp*H exponentials in pass 1 plus pass 2 up to each CTA's causal limit",
                              "ex2_per_launch": B * scorer_exps(S), "ex2_per_s": B * scorer_exps(S) / (ms_scorer * 1e-3),
                              "peak_ex2_per_s": 16 * 148 * 1.965e9, "peak_source": "tools/ubench/mufu.cu on this pool: 15.9 ex2/clk/SM",
                              "frac": B * scorer_exps(S) / (ms_scorer * 1e-3) / (16 * 148 * 1.965e9)},
                "step_tflops": B * fl["total"] / (ms_step * 1e-3) / 1e12,
                "step_frac_of_tensor_peak": B * fl["total"] / (ms_step * 1e-3) / 1e12 / pk["tf_sustained"],
                "kernel_ms_breakdown": kms,
                "algorithmic_gflop_per_seq": {k: v / 1e9 for k, v in fl.items() if k != "sel_gather_bytes"},
                "per_kernel": [{"kernel": n_, "ms": m_, "bound": b_, "achieved": (w_ / (m_ * 1e-3) / (1e12 if b_ == "tensor" else 1e9)),
                                "unit": "TFLOP/s" if b_ == "tensor" else "GB/s",
                                "frac": (w_ / (m_ * 1e-3) / (1e12 if b_ == "tensor" else 1e9)) / (pk["tf_sustained"] if b_ == "tensor" else pk["hbm"])}
                               for n_, m_, b_, w_, _ in cand]}
    # the same 64k step under the semantics the reference itself can run at this length (per-token decode steps, i.e.
    # NSA_PREFILL_TILE: decode selection rule, p_cmp normalised over the visible compressed keys only) -- reported beside the headline,
    # which keeps the batched-prefill rule with the full-row normaliser (more exponentials)
    alt = None
    try:
        cfg_c = ops.NSAConfig(l=c["l"], d=c["d"], l_sel=c["l_sel"], n_sel=c["n_sel"], w=c["w"], norm_mode=ops.NORM_CAUSAL)

        def step_causal():
            rg2 = ops.score_select(inp["Q"], inp["K_cmp"], cfg_c, mode=1)
            ops.prefill_core(inp["Q"], inp["K_sel"], inp["V_sel"], inp["K_win"], inp["V_win"], inp["K_cmp"], inp["V_cmp"], gate, cfg_c,
                             sel_mode=1, ranges=rg2, ranges_trusted=True)
        with torch.no_grad():
            ms_alt = t_of(step_causal, n=5)
        alt = {"selection": "decode rule (select_topn_ranges), causal p_cmp normaliser: what NSA_PREFILL_TILE / stepwise decode computes",
               "ms_per_step": ms_alt, "tok_per_s_per_gpu": B * S / (ms_alt * 1e-3)}
    except Exception as ex:
        alt = {"error": f"{type(ex).__name__}: {ex}"}
    roofline["alt_semantics"] = alt
    cpu, cpu_port, parity = None, None, None
    if not args.no_cpu:
        # (1) oracle port on the GPU step's own inputs: a bounded sample of the batched-prefill workload AND the parity check
        host_in = {k: v[:1].float().cpu() for k, v in inp.items()}
        v, cores, sample, ora, t0p = cpu_prefill_sample(S, args.cpu_rows, tensors=host_in, gate=tuple(t.float().cpu() for t in gate))
        cpu_port = {"value": v, "unit": "tok/s", "cores": cores, "kind": "port", "sample": sample}
        try:
            parity = parity_64k(ora, t0p, args.cpu_rows, gpu_rows)
        except Exception as ex:
            parity = {"error": f"{type(ex).__name__}: {ex}"}
        del host_in, ora
        cpu = cpu_port
        # (2) the UNMODIFIED reference on the same host cores (>= 128 decode steps up to a 65,535-token cache)
        if reference_path() is not None:
            try:
                v, cores, sample, ms_last = reference_decode_sample(S, strata=16, reps=8)
                cpu = {"value": v, "unit": "tok/s", "cores": cores, "kind": "reference", "sample": sample,
                       "ms_per_token_at_64k_context": ms_last}
            except Exception as ex:
                cpu_port["reference_error"] = f"{type(ex).__name__}: {str(ex)[:200]}"
    line = {"metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": workload_config(args, world), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "tok/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "how": "ModulePrefillEngine.run: x [B,S,768] bf16 from pinned host memory -> NSAAttention.forward(x, kv, prefill=True) "
                           "(projections, RoPE, cache build, hot path, output projection) -> out [B,S,768] to pinned host memory, every "
                           "step; H2D / kernels / D2H of neighbouring steps on 3 streams",
                    "matches_device_path": e2e_ok, "module_ms_device_resident": ms_mod, "numa_node_bound": numa_node,
                    "hot_path_only": {"value": hp_val, "unit": "tok/s", "h2d_bytes_per_step": hp_h2d, "d2h_bytes_per_step": B * S * c["H"] * c["Dv"] * 2,
                                      "how": "PrefillEngine.run: post-projection Q + six K/V tensors in, O out (round-1 e2e definition)",
                                      "matches_device_path": hp_ok}},
            "e2e_stack12": stack,
            "module_prefill": {"what": "NSAAttention.forward(prefill=True), B=1, bf16, device-resident x: the sizes the reference's GPU route can run "
                                       "(see reference_gpu in the --impl reference line)", **module_small},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "cpu_port": cpu_port, "parity_64k": parity,
            "decode": decode, "train_core": train, "train_ddp": train_ddp}
    print(json.dumps(line))
    _leave(world, train_ddp)


def bench_stack_e2e(args, dev, rank, world, barrier, layers=12, steps=6):
    """End to end through a stack: x [B,S,768] bf16 in pinned host memory -> 12 LlamaBlockNSA layers (m7c dims; RMSNorm kernels,
    NSAAttention prefill with the batched selection rule, SiLU MLP) -> hidden states back to pinned host memory, every step, with
    the copies of neighbouring steps overlapped (StackPrefillEngine).  Same bytes per step as the one-layer e2e, twelve times the
    compute: what the host side of the box can feed."""
    import torch.distributed as dist
    from nsa_vibe_b200.engine import StackPrefillEngine
    from nsa_vibe_b200.model.llama_block_nsa import LlamaBlockNSA
    c = M7C
    os.environ["NSA_PREFILL_BATCHED"] = "1"
    torch.manual_seed(21 + rank)
    blocks = [LlamaBlockNSA(768, c["H"], c["G"], c["Dk"], c["Dv"], c["l"], c["d"], c["l_sel"], c["n_sel"], c["w"]).to(dev).bfloat16()
              for _ in range(layers)]
    B, S = args.B, args.S
    x_host = [(torch.randn(B, S, 768) * 0.5).bfloat16().pin_memory() for _ in range(2)]
    y_host = [torch.empty((B, S, 768), dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    eng = StackPrefillEngine(blocks, dev)
    with torch.no_grad():
        eng.run([{"x": x_host[i % 2]} for i in range(2)], [y_host[i % 2] for i in range(2)])
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.run([{"x": x_host[i % 2]} for i in range(steps)], [y_host[i % 2] for i in range(steps)])
        e.record()
        barrier()
    tt = torch.tensor([s.elapsed_time(e) / steps], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    finite = bool(torch.isfinite(y_host[(steps - 1) % 2].float()).all())
    return {"what": f"{layers} LlamaBlockNSA layers, x [B,S,768] bf16 from / hidden states to pinned host memory every step", "layers": layers,
            "ms_per_step": ms, "value": world * B * S / (ms * 1e-3), "unit": "tok/s (through the whole stack)",
            "h2d_bytes_per_step": B * S * 768 * 2, "d2h_bytes_per_step": B * S * 768 * 2, "finite": finite}


def _leave(world, train_ddp):
    """A process group whose collectives sit in a live CUDA graph does not tear down cleanly (destroy_process_group hung on two
    B200s until the launcher's timeout): flush, meet the other ranks and leave without the NCCL destructor."""
    import torch.distributed as dist
    if world <= 1:
        return
    if train_ddp and train_ddp.get("graph"):
        sys.stdout.flush()
        sys.stderr.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    dist.destroy_process_group()


def bench_train_ddp(dev, rank, world, barrier, S=2048, B=8, layers=12, steps=8):
    """Config C5: one training step of the 12-layer m7c TinyLM (78.3 M parameters; NSA forward / backward kernels,
    RMSNorm kernels, bf16 autocast, fused AdamW) with the reference's gradient exchange -- divide by world, round to bf16,
    all_reduce over NCCL, copy back (scripts/train_showcase.py:654-665) -- issued per layer bucket from inside backward so it
    overlaps the backward of the earlier layers (nsa_vibe_b200.dist.OverlappedGradExchange), the whole step replayed as ONE CUDA
    graph.  B sequences of S tokens per GPU, synthetic byte tokens: weak scaling over GPUs."""
    import torch.distributed as dist
    import torch.nn.functional as F
    from nsa_vibe_b200 import dist as nd
    from nsa_vibe_b200.model.tiny_lm import m7c_tiny_lm
    os.environ["NSA_PREFILL_BATCHED"] = "1"
    torch.manual_seed(1337 + rank)
    model = m7c_tiny_lm(layers).to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    check = None
    if world > 1:
        nd.broadcast_parameters(model, 0)
        check = nd.check_bf16_exchange(dev)
    exch = nd.OverlappedGradExchange(model.grad_buckets()) if world > 1 else None
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4, capturable=True, fused=True)
    ids = torch.randint(0, 256, (B, S + 1), device=dev)

    def body():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = model(ids[:, :-1])
        loss = F.cross_entropy(logits.float().reshape(-1, 256), ids[:, 1:].reshape(-1))
        loss.backward()
        if exch is not None:
            exch.finish()
        opt.step()
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            opt.zero_grad(set_to_none=True)
            body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    holder = {}
    with torch.cuda.graph(graph):
        holder["loss"] = body()
    for _ in range(3):
        graph.replay()
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        graph.replay()
    e.record()
    barrier()
    tt = torch.tensor([s.elapsed_time(e) / steps], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    loss = float(holder["loss"].detach())
    return {"what": "config C5: 12-layer m7c TinyLM training step (forward + backward + bf16-compressed gradient exchange + fused AdamW), "
                    "one CUDA-graph replay per step", "layers": layers, "params": n_params, "S": S, "batch_per_gpu": B, "n_gpus": world,
            "ms_per_step": ms, "tokens_per_s": world * B * S / (ms * 1e-3), "loss": loss, "graph": True,
            "grad_exchange": ("per-layer bf16 buckets, async NCCL all_reduce issued from inside backward (overlapped), captured in the graph"
                              if world > 1 else "none (single GPU)"),
            "allreduce_bytes_per_step": 2 * n_params if world > 1 else 0, "exchange_numerics": check}


def bench_train_core(ops, cfg, dev, rank, world, barrier, S=2048, B=8):
    """Forward + analytical backward of the hot path at the training shape of config C5 (m7c dims, S=2048, bf16), one layer,
    B sequences per GPU: scoring + selection + three branches + gated combine, then dQ / dK / dV of every branch and the
    gate MLP gradients (tensor-core backward, tc_bwd.cu)."""
    import torch.distributed as dist
    c = M7C
    inp, gate = make_inputs(B, S, dev, seed=77 + rank)
    leaves = [inp[k].requires_grad_(True) for k in ("Q", "K_sel", "V_sel", "K_win", "V_win", "K_cmp", "V_cmp")]
    gate = tuple(t.requires_grad_(True) for t in gate)
    dO = torch.randn((B, S, c["G"], c["h"], c["Dv"]), device=dev).to(torch.bfloat16)

    def fwd():
        return ops.prefill_core(*leaves, gate, cfg, sel_mode=0)[0]

    def timed(fn, n=5):
        for _ in range(2):
            fn()
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record()
        barrier()
        tt = torch.tensor([s.elapsed_time(e) / n], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    with torch.no_grad():
        ms_f = timed(fwd)
    O = fwd()
    ms_b = timed(lambda: torch.autograd.grad(O, leaves + list(gate), dO, retain_graph=True))
    return {"what": "NSA hot path forward + backward, one layer, training shape (config C5)", "S": S, "batch_per_gpu": B,
            "fwd_ms": ms_f, "bwd_ms": ms_b, "tokens_per_s_fwd_bwd": world * B * S / ((ms_f + ms_b) * 1e-3),
            "bwd_kernels": "bwd_delta16 + bwd_tc_kernel[cmp|win|sel] (tcgen05, KV-tile-major, dK/dV in TMEM, dQ by TMA bulk "
                           "reductions) + sel2 index + gate_bwd_fast; bwd_ms includes torch's fp32 zero-fills and casts"}


def bench_decode(args, ops, cfg, dev, rank, world, barrier):
    """us/token of one decode step at context S (caches resident, pre-allocated): score the emitted compressed keys,
    select, attend over cmp + selected + window, gate, combine -- for decode_B sequences per GPU."""
    import torch.distributed as dist
    c = M7C
    S, Bd = args.decode_S, args.decode_B
    cap = S + 64
    g = torch.Generator(device=dev).manual_seed(99 + rank)
    r = lambda *s: torch.randn(*s, generator=g, device=dev, dtype=torch.float32).to(torch.bfloat16)
    K_sel, V_sel, K_win, V_win = r(Bd, c["G"], cap, c["Dk"]), r(Bd, c["G"], cap, c["Dv"]), r(Bd, c["G"], cap, c["Dk"]), r(Bd, c["G"], cap, c["Dv"])
    S_cmp = (S - c["l"]) // c["d"] + 1
    K_cmp, V_cmp = r(Bd, c["G"], S_cmp + 8, c["Dk"]), r(Bd, c["G"], S_cmp + 8, c["Dv"])
    q = r(Bd, 1, c["G"], c["h"], c["Dk"])
    hid = c["Dk"] // 2
    gate = (torch.randn(hid, c["Dk"], generator=g, device=dev) * 0.1, torch.zeros(hid, device=dev),
            torch.randn(3, hid, generator=g, device=dev) * 0.1, torch.zeros(3, device=dev))
    gate_cache = ops._gate_struct(gate, dev)
    out = torch.empty((Bd, 1, c["G"], c["h"], c["Dv"]), dtype=torch.bfloat16, device=dev)
    rg = torch.empty((Bd, c["G"], c["n_sel"], 2), dtype=torch.int32, device=dev)

    def dstep():
        ops.decode_core(q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, gate, cfg, t=S - 1, S_sel_kv=S, S_win_kv=S, win_off=0,
                        S_cmp=S_cmp, ranges_out=rg, out=out, gate_cache=gate_cache)

    for _ in range(3):
        dstep()
    barrier()
    n = 20
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        dstep()
    e.record()
    barrier()
    tt = torch.tensor([s.elapsed_time(e) / n], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    # single-sequence latency (launch-bound: 0.9 MB per step cannot fill the machine)
    q1, o1, r1 = q[:1], out[:1], rg[:1]
    def d1():
        ops.decode_core(q1, K_sel[:1], V_sel[:1], K_win[:1], V_win[:1], K_cmp[:1], V_cmp[:1], gate, cfg, t=S - 1, S_sel_kv=S,
                        S_win_kv=S, win_off=0, S_cmp=S_cmp, ranges_out=r1, out=o1, gate_cache=gate_cache)
    for _ in range(3):
        d1()
    torch.cuda.synchronize()
    s1, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s1.record()
    for _ in range(n):
        d1()
    e1.record()
    torch.cuda.synchronize()
    b1_us = s1.elapsed_time(e1) / n * 1e3
    # module-level line: timed locally (no collective inside the try, so a rank that fails cannot leave the others waiting);
    # the max over ranks is taken outside
    module, mod_ms = None, float("nan")
    try:
        module = bench_decode_module(S, Bd, dev, rank)
        mod_ms = module["us_per_step"] * 1e-3
        module["b1_us_per_step"] = bench_decode_module(S, 1, dev, rank)["us_per_step"]
    except Exception as ex:  # the module-level line is additional evidence; never lose the kernel line over it
        module = {"error": f"{type(ex).__name__}: {ex}"}
    if world > 1:
        mt = torch.tensor([mod_ms if mod_ms == mod_ms else 1e30], device=dev)
        dist.all_reduce(mt, op=dist.ReduceOp.MAX)
        if "error" not in module and float(mt.item()) < 1e29:
            module["us_per_step"] = float(mt.item()) * 1e3
            module["us_per_token"] = module["us_per_step"] / (Bd * world)
        elif "error" not in module:
            module["note"] += "; another rank failed, figures are this rank's"
    byts, reads = decode_bytes_per_token(S)
    pk = measured_peaks()
    ach = Bd * byts / (ms * 1e-3) / 1e9
    return {"metric": "NSA decode us/tok @S=4k", "value": ms * 1e3 / (Bd * world), "unit": "us/token (step latency / global batch)",
            "S": S, "batch_per_gpu": Bd, "ms_per_step": ms, "tokens_per_s": world * Bd / (ms * 1e-3), "b1_step_latency_us": b1_us,
            "kernel": "gather_attn_tc_kernel (fused decode step: scoring, selection, gate, cmp+sel+win attention, combine; 1 launch)",
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s", "frac": ach / pk["hbm"],
                         "traffic": load_traffic().get("gather_attn_tc_kernel[decode]"), "algorithmic_bytes_per_token": byts, "reads_per_token": reads, "peak_source": pk["src"]},
            "module": module}


def bench_decode_module(S, Bd, dev, rank, n=24):
    """The same step through the reference-facing module API: NSAAttention.forward(x [B,1,dim], kv, prefill=False) at context S
    (bench/bench_decode.py:113-136) -- one GEMM for the seven projections, the produce kernel (RoPE + in-place cache rows +
    read counters), phi on emission steps, the fused decode kernel, the output projection.  Caches are built directly from random
    tensors of length S - n (no prefill needed for timing), bf16, m7c dims."""
    from nsa_vibe_b200.cache.kv_cache import create_empty_kv
    from nsa_vibe_b200.core.block_index import build_block_meta
    from nsa_vibe_b200.core.nsa_attention import NSAAttention
    c = M7C
    torch.manual_seed(5 + rank)
    attn = NSAAttention(dim=768, n_heads=c["H"], n_kv_groups=c["G"], d_k=c["Dk"], d_v=c["Dv"], l=c["l"], d=c["d"], l_sel=c["l_sel"],
                        n_sel=c["n_sel"], w=c["w"]).to(dev).bfloat16()
    S0 = S - n - 8
    kv = create_empty_kv(Bd, c["G"], c["Dk"], c["Dv"], build_block_meta(S + 64, c["l"], c["d"], c["l_sel"], c["n_sel"], c["w"]),
                         device=dev, dtype=torch.bfloat16)
    r = lambda rows, D: torch.randn(Bd, c["G"], rows, D, device=dev, dtype=torch.bfloat16)
    kv.update_selection_raw(r(S0, c["Dk"]), r(S0, c["Dv"]))
    kv.update_window(r(S0, c["Dk"]), r(S0, c["Dv"]), c["w"])
    kv.append_cmp_raw(r(S0, c["Dk"]), r(S0, c["Dv"]))
    kv.append_compressed(r((S0 - c["l"]) // c["d"] + 1, c["Dk"]), r((S0 - c["l"]) // c["d"] + 1, c["Dv"]))
    kv.reserve(S + 64)
    x1 = torch.randn(Bd, 1, 768, device=dev, dtype=torch.bfloat16)
    with torch.no_grad():
        for _ in range(8):
            attn(x1, kv, prefill=False)
        torch.cuda.synchronize()
        k0 = ops_launches()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            attn(x1, kv, prefill=False)
        e.record()
        torch.cuda.synchronize()
        k1 = ops_launches()
    ms = s.elapsed_time(e) / n
    return {"what": "NSAAttention.forward(prefill=False): projections + cache append + fused decode step + output projection",
            "us_per_step": ms * 1e3, "us_per_token": ms * 1e3 / Bd, "batch_per_gpu": Bd, "context": int(kv.K_sel.shape[2]),
            "nsa_launches_per_step": (k1 - k0) / n, "graph": getattr(kv, "_decode_graph", (None, None, None))[2] is not None,
            "hbm_frac": (Bd * decode_bytes_per_token(S)[0] / (ms * 1e-3) / 1e9) / measured_peaks()["hbm"],
            "note": "one replayed CUDA graph per step (ops.DecodeGraphStep: GEMM, produce, emit, fused decode, GEMM, advance; position and "
                    "row counts in a device record), token copied in and result cloned out on the host side"}


def ops_launches():
    from nsa_vibe_b200 import ops
    return ops.launch_count


if __name__ == "__main__":
    main()
