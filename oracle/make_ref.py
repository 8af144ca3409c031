#!/usr/bin/env python3
"""Recipe for oracle/_ref: a verbatim, git-ignored copy of the reference's own `nsa/` package (pure Python, no build step)
taken from /root/reference where it lies, so that the UNMODIFIED reference travels to the GPU box with the snapshot and
`bench.py --impl reference` / the `cpu_baseline` leg can time it on that box's host cores, and tests/ref_compat can read
the reference's own test files.  TEST / BASELINE INFRASTRUCTURE: nothing under nsa_vibe_b200/ imports it, no reference
source enters the git history (oracle/_ref/ is in .gitignore, not in .gpurunignore).

    python oracle/make_ref.py            # no-op when /root/reference is absent (the GPU box uses the copy it was sent)
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("NSA_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")
# the package itself, plus the two stand-alone scripts BASELINE.md section 3 names
ITEMS = ("nsa", os.path.join("bench", "bench_decode.py"), os.path.join("bench", "needle_64k_smoke.py"))


def make_ref(verbose: bool = True) -> str | None:
    if not os.path.isdir(os.path.join(SRC, "nsa")):
        if verbose:
            print(f"[make_ref] {SRC}/nsa not present; keeping {DST} as it is" + ("" if os.path.isdir(DST) else " (absent)"))
        return DST if os.path.isdir(os.path.join(DST, "nsa")) else None
    os.makedirs(DST, exist_ok=True)
    for item in ITEMS:
        s, d = os.path.join(SRC, item), os.path.join(DST, item)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc", "*.so"))
        elif os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Verbatim copy of /root/reference/{nsa,bench/bench_decode.py,bench/needle_64k_smoke.py} made by oracle/make_ref.py.\n"
                "Git-ignored; baseline / test infrastructure only.\n")
    if verbose:
        print(f"[make_ref] copied the reference package to {DST}")
    return DST


def ref_path() -> str | None:
    """Directory to put on sys.path to `import nsa` (the reference), or None."""
    return DST if os.path.isdir(os.path.join(DST, "nsa")) else None


if __name__ == "__main__":
    sys.exit(0 if make_ref() is not None or True else 1)
