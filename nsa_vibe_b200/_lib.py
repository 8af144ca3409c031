"""ctypes binding of libnsa_b200.so -- the C ABI declared in include/nsa_b200.h.

There is deliberately no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NSA_B200_LIB", os.path.join(_HERE, "lib", "libnsa_b200.so"))

NSA_F32, NSA_BF16, NSA_F16 = 0, 1, 2
NORM_FULL_ROW, NORM_CAUSAL = 0, 1
GATE_MLP, GATE_UNIFORM, GATE_CMP, GATE_SEL, GATE_WIN = 0, 1, 2, 3, 4
IMPL_AUTO, IMPL_SIMT, IMPL_TC = 0, 1, 2
WS_SCORE_SELECT, WS_DECODE, WS_PREFILL, WS_SEL_BLOCKMAJOR, WS_BWD, WS_PREFILL_FULL = 0, 1, 2, 3, 4, 5
MAX_SEL_BLOCKS = 16  # NSA_MAX_SEL_BLOCKS


class Dims(C.Structure):
    """struct nsa_dims (include/nsa_b200.h)."""

    _fields_ = [(n, C.c_int32) for n in (
        "B", "S", "G", "h", "Dk", "Dv", "l", "d", "l_sel", "n_sel", "w", "t0",
        "S_sel_kv", "cap_sel", "S_win_kv", "cap_win", "win_off", "S_cmp", "cap_cmp",
        "n_ranges", "dtype", "gate_mode", "gate_hidden", "norm_mode", "impl")] + [
        ("gate_tau", C.c_float), ("scale", C.c_float)]


class GateParams(C.Structure):
    """struct nsa_gate_params (include/nsa_b200.h)."""

    _fields_ = [("fc1_w", C.c_void_p), ("fc1_b", C.c_void_p), ("fc2_w", C.c_void_p), ("fc2_b", C.c_void_p)]


class DecodeProduce(C.Structure):
    """struct nsa_decode_produce (include/nsa_b200.h)."""

    _fields_ = [("y", C.c_void_p), ("q_out", C.c_void_p), ("slab", C.c_void_p * 6), ("cap", C.c_int32 * 6), ("row", C.c_int32 * 6),
                ("counters", C.c_void_p), ("counters_cap", C.c_int32), ("counters_idx", C.c_int32), ("counter_val", C.c_int64 * 5),
                ("B", C.c_int32), ("H", C.c_int32), ("G", C.c_int32), ("Dk", C.c_int32), ("Dv", C.c_int32), ("t", C.c_int32),
                ("base", C.c_float), ("scale", C.c_float), ("dtype", C.c_int32), ("S", C.c_int32), ("inverse", C.c_int32),
                ("state", C.c_void_p), ("l", C.c_int32), ("d", C.c_int32), ("l_sel", C.c_int32), ("n_sel", C.c_int32), ("w", C.c_int32),
                ("rope_q", C.c_void_p), ("rope_k", C.c_void_p), ("rope_t0", C.c_int32), ("rope_rows", C.c_int32)]


class DecodeState(C.Structure):
    """struct nsa_decode_state (include/nsa_b200.h); lives in device memory, this mirror is for sizes and host-side initialisation."""

    _fields_ = [("t", C.c_int32), ("row_win", C.c_int32), ("row_raw", C.c_int32), ("S_cmp", C.c_int32), ("ctr_idx", C.c_int32),
                ("pad_", C.c_int32 * 3)]


class DecodeEmit(C.Structure):
    """struct nsa_decode_emit (include/nsa_b200.h)."""

    _fields_ = [("state", C.c_void_p), ("K_raw", C.c_void_p), ("V_raw", C.c_void_p), ("K_cmp", C.c_void_p), ("V_cmp", C.c_void_p),
                ("BG", C.c_int32), ("cap_raw", C.c_int32), ("cap_cmp", C.c_int32), ("Dk", C.c_int32), ("Dv", C.c_int32),
                ("l", C.c_int32), ("d", C.c_int32), ("base", C.c_float), ("scale", C.c_float), ("dtype", C.c_int32),
                ("w_k", C.c_void_p), ("w_v", C.c_void_p)]


class Stats(C.Structure):
    """struct nsa_stats (include/nsa_b200.h)."""

    _fields_ = [("gate_sum", C.c_double * 6), ("entropy_min_ord", C.c_int32), ("max_gate_max_ord", C.c_int32), ("k_sum", C.c_int64),
                ("k_max", C.c_int32), ("pad_", C.c_int32), ("rows_at_max", C.c_int64)]


_P, _I, _I64 = C.c_void_p, C.c_int, C.c_int64
_DP, _GP = C.POINTER(Dims), C.POINTER(GateParams)

# name -> (restype, argtypes); must list every symbol include/nsa_b200.h declares
SIGNATURES = {
    "nsa_version": (C.c_char_p, []),
    "nsa_last_error": (C.c_char_p, []),
    "nsa_kernel_launches": (_I64, []),
    "nsa_prefill_range_cols": (_I, [_I, _I, _I]),
    "nsa_prefill_range_cols_ex": (_I, [_I, _I, _I, _I, _I]),
    "nsa_select_ranges_prefill": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "nsa_select_ranges_decode": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "nsa_pcmp_all": (_I, [_DP, _P, _P, _P, _P]),
    "nsa_map_pcmp_to_pslc": (_I, [_P, _I64, _I, _I, _I, _I, _I, _P, _P]),
    "nsa_indices_to_ranges": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "nsa_score": (_I, [_DP, _P, _P, _I, _P, _P]),
    "nsa_score_select": (_I, [_DP, _P, _P, _I, _I, _I, _P, _P, _P]),
    "nsa_branch_attn_fwd": (_I, [_DP, _I, _P, _P, _P, _P, _P, _P, _P]),
    "nsa_sel_attn_fwd_blockmajor": (_I, [_DP, _P, _P, _P, _P, _P, _P, _P, _P]),
    "nsa_branch_attn_bwd": (_I, [_DP, _I] + [_P] * 12),
    "nsa_gate_fwd": (_I, [_DP, _P, _GP, _P, _P]),
    "nsa_gate_bwd": (_I, [_DP, _P, _GP, _P, _P, _P, _P, _P, _P, _P]),
    "nsa_prefill_fwd": (_I, [_DP] + [_P] * 8 + [_GP] + [_P] * 6),
    "nsa_prefill_full_fwd": (_I, [_DP] + [_P] * 7 + [_GP, _I, _I, _I] + [_P] * 7),
    "nsa_score_stats": (_I, [_DP, _P, _P, _P, _P]),
    "nsa_score_cmp": (_I, [_DP, _P, _P, _P, _I, _P, _P, _P, _P, _P]),
    "nsa_prefill_bwd": (_I, [_DP] + [_P] * 22),
    "nsa_decode_fwd": (_I, [_DP] + [_P] * 7 + [_GP] + [_P] * 4),
    "nsa_rope_shape": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, C.c_float, C.c_float, _I, _I, _P]),
    "nsa_phi_avgpool": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, C.c_float, C.c_float, _I, _I, _P]),
    "nsa_phi_conv": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, C.c_float, C.c_float, _I, _I, _P]),
    "nsa_decode_produce": (_I, [C.c_void_p, _P]),
    "nsa_rope_table": (_I, [_I, _I, _I, _I, C.c_float, C.c_float, _I, _P, _P]),
    "nsa_decode_emit": (_I, [C.c_void_p, _P]),
    "nsa_decode_advance": (_I, [_P, _I, _I, _P]),
    "nsa_decode_stepped_supported": (_I, [_DP]),
    "nsa_decode_fwd_stepped": (_I, [_DP] + [_P] * 7 + [_GP] + [_P] * 4),
    "nsa_rmsnorm_fwd": (_I, [_P] * 6 + [_I, _I, C.c_float, _I, _I, _I, _I, _P]),
    "nsa_rmsnorm_bwd": (_I, [_P] * 8 + [_I, _I, _I, _I, _I, _P]),
    "nsa_rmsnorm_partials": (_I, [_I]),
    "nsa_stats": (_I, [_P, _I64, _P, _I64, _I, _P, _P, _P]),
    "nsa_ranges_max_blocks": (_I, [_P, _I64, _I, _I, _P, _P]),
    "nsa_workspace_bytes": (_I64, [_DP, _I]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load the shared library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"libnsa_b200.so not found at {LIB_PATH}: build it with `python -m nsa_vibe_b200.build` "
                "(there is no CPU or PyTorch fallback for the NSA hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().nsa_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")
