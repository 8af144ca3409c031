"""Print the timeline of one CTA of the win backward walk from a -DNSA_BWD_DBG build:
    NSA_B200_NVCC_FLAGS=-DNSA_BWD_DBG python -m nsa_vibe_b200.build --force && python tools/prof_bwd.py 2> log && python tools/dbg_bwd.py log
tags: 1/2 producer wait/got stage; 10-12 MMA S/dP (wait Q/dO, wait S free, issue); 13-16 MMA gradients (wait P/dS, wait dQ free,
issue, issued); 20-23 softmax (wait S, wait P buffer, start, done); 30-32 drain (wait dQ, start, done)."""
import sys

NAMES = {1: "prod wait", 2: "prod go", 10: "mma sdp: wait qdo", 11: "mma sdp: wait s_empty", 12: "mma sdp: issue",
         13: "mma grads: wait pds", 14: "mma grads: wait dq_empty", 15: "mma grads: issue", 16: "mma grads: issued",
         20: "soft: wait s_full", 21: "soft: wait pds_empty", 22: "soft: start", 23: "soft: done",
         30: "drain: wait dq_full", 31: "drain: start", 32: "drain: done"}
if len(sys.argv) > 4 and sys.argv[4] == "sel2":  # -DNSA_SEL2_DBG build: the block-major forward (tc_sel2.cu)
    NAMES = {1: "prod wait q_empty", 2: "prod go", 10: "mma qk: wait q_full", 11: "mma qk: wait s_empty", 12: "mma qk: issue",
             13: "mma pv: wait p_full", 14: "mma pv: wait o_empty", 15: "mma pv: issue", 20: "soft: wait s_full", 21: "soft: start",
             22: "soft: P ready, wait p_empty", 23: "soft: P published", 24: "soft: O arrived", 25: "soft: partial stored",
             30: "epi: wait o_full", 31: "epi: O arrived", 32: "epi: store issued"}
ev = []
for line in open(sys.argv[1]):
    if line.startswith("BDBG "):
        _, tag, it, clk = line.split()
        ev.append((int(clk), int(tag), int(it)))
ev.sort()
t0 = ev[0][0] if ev else 0
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (4, 9)
for clk, tag, it in ev:
    if lo <= it <= hi:
        print(f"{clk - t0:8d}  tile {it:3d}  {NAMES.get(tag, tag)}")
