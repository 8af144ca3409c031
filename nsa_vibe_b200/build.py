"""Builds libnsa_b200.so (hand-written sm_100a CUDA behind a C ABI) in-tree with nvcc.

    python -m nsa_vibe_b200.build [--force]

The library is plain CUDA runtime code: it links against nothing from torch, so it can be loaded by any
host (ctypes here; cgo/JNI/etc. elsewhere).  The built .so is git-ignored but travels with `gpurun`.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libnsa_b200.so")
SOURCES = ["api.cu", "select.cu", "generic.cu", "tc_dispatch.cu", "tc_score.cu", "tc_score_cmp.cu", "tc_dense.cu", "tc_gather.cu", "tc_sel2.cu", "tc_bwd.cu", "producers.cu", "block_ops.cu", "stats.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
] + os.environ.get("NSA_B200_NVCC_FLAGS", "").split()  # e.g. -DNSA_DENSE_DBG for the debug timelines


def _deps():
    out = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    out.append(os.path.join(os.path.dirname(HERE), "include", "nsa_b200.h"))
    return out


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(p) > t for p in _deps())


def _compile(src: str) -> str:
    obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
    cmd = [NVCC, *FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(LIBDIR, src.replace(".cu", ".ptxas.log"))
    with open(log, "w") as f:
        f.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stderr[-4000:]}")
    return obj


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and not needs_build():
        return LIB
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build libnsa_b200.so")
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stderr[-4000:]}")
    if verbose:
        print(f"built {LIB}")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
