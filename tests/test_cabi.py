"""The C-ABI library loads and exports exactly what include/nsa_b200.h declares (CPU: no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nsa_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nsa_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from nsa_vibe_b200 import _lib
    from nsa_vibe_b200.build import build
    build()  # no-op when up to date; nvcc cross-compiles without a GPU
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in nsa_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES and the header disagree"
    assert b"sm_100a" in _lib.load().nsa_version()


def test_invalid_arguments_return_codes_not_crashes():
    from nsa_vibe_b200 import _lib
    lib = _lib.load()
    assert lib.nsa_prefill_range_cols(200, 64, 16) == 4      # 3 forced + min(13, 4) -> n >= S_sel -> S_sel
    assert lib.nsa_prefill_range_cols(2048, 64, 16) == 16
    assert lib.nsa_prefill_range_cols(40, 64, 16) == 1
    rc = lib.nsa_select_ranges_prefill(None, 1, 1, 1, 4, 64, 16, 200, 0, 4, 1, 2, None, None)
    assert rc == -1 and b"NULL" in lib.nsa_last_error()
    dm = _lib.Dims()
    dm.B, dm.S, dm.G, dm.h, dm.Dk, dm.Dv = 1, 1, 1, 1, 16, 16
    dm.l, dm.d, dm.l_sel = 32, 12, 64  # d does not divide l
    rc = lib.nsa_score(ctypes.byref(dm), None, None, 1, None, None)
    assert rc == -1 and b"d|l" in lib.nsa_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "nsa_score")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from nsa_vibe_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()
