#!/usr/bin/env python3
"""Per-kernel SASS evidence of libnsa_b200.so: counts of the Blackwell mnemonics (UTCHMMA = tcgen05.mma, LDTM / STTM =
tcgen05.ld / st, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP / UBLKRED = bulk copy / bulk reduce, MUFU.EX2) from `cuobjdump -sass`, and
registers / spills / shared memory from the `-Xptxas -v` logs the build keeps next to the objects.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
from __future__ import annotations

import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nsa_vibe_b200", "lib")
MNEMONICS = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKRED", "MUFU.EX2", "SYNCS")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def ptxas_info():
    """mangled name -> (registers, spill stores, spill loads, smem bytes)."""
    info = {}
    for log in glob.glob(os.path.join(LIB, "*.ptxas.log")):
        txt = open(log).read()
        for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'(.*?)(?=ptxas info\s+: Compiling|\Z)", txt, re.S):
            body = m.group(2)
            regs = re.search(r"Used (\d+) registers", body)
            sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", body)
            sm = re.search(r"(\d+) bytes smem", body)
            info[m.group(1)] = (int(regs.group(1)) if regs else -1, int(sp.group(1)) if sp else 0, int(sp.group(2)) if sp else 0,
                                int(sm.group(1)) if sm else 0)
    return info


def main():
    so = os.path.join(LIB, "libnsa_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    counts, cur = {}, None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = dict.fromkeys(MNEMONICS, 0)
            counts[cur]["instr"] = 0
            continue
        if cur is None or "/*" not in ln:
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["instr"] += 1
        for k in MNEMONICS:
            if op.startswith(k):
                counts[cur][k] += 1
    info = ptxas_info()
    names = demangle(list(counts))
    print("# libnsa_b200.so -- per-kernel SASS mnemonic counts (cuobjdump -sass) and ptxas resources (-Xptxas -v)")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = cp.async.bulk.tensor load/store (TMA), UBLKCP/UBLKRED = bulk copy/reduce")
    hdr = f"{'kernel':<78} {'instr':>6} " + " ".join(f"{k:>8}" for k in MNEMONICS) + f" {'regs':>5} {'spill_st':>8} {'spill_ld':>8}"
    print(hdr)
    for fn in sorted(counts, key=lambda f: names[f]):
        c = counts[fn]
        nm = re.sub(r"\(.*", "", names[fn]).replace("void ", "").replace("nsa::", "").replace("(anonymous namespace)::", "")
        r = info.get(fn, (-1, 0, 0, 0))
        print(f"{nm[:78]:<78} {c['instr']:>6} " + " ".join(f"{c[k]:>8}" for k in MNEMONICS) + f" {r[0]:>5} {r[1]:>8} {r[2]:>8}")
    tc = [f for f in counts if counts[f]["UTCHMMA"]]
    print(f"# {len(counts)} kernels, {len(tc)} issue tcgen05.mma; any spill: {any(info.get(f, (0, 0, 0, 0))[1] for f in counts)}")


if __name__ == "__main__":
    sys.exit(main())
