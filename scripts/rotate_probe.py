#!/usr/bin/env python3
"""Tuning aid: the fused set call over FOUR resident 50 MB texts in rotation (200 MB > the 126 MB L2: the text is
cold, the kernel's code, parameters and tables stay warm) next to the same call after an L2 flush."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rejit_b200 as rj
from rejit_b200 import workloads as W
rs = rj.RegejSet(W.DNA_PATTERNS)
texts = [rj.DeviceText(W.fasta_sequence(5_000_000, seed=42 + i)) for i in range(4)]
st = rj.Stats()
for mode in (sys.argv[1:] or ["rotate", "flush"]):
    t = []
    for i in range(24):
        if mode == "flush":
            rj.lib().rejit_b200_flush_l2(0)
        rs.match_all_device(texts[i % 4], stats=st)
        t.append((st.scan_ms, st.total_ms))
    t = sorted(t[4:])
    print(mode, "kernel us: min %.1f med %.1f | pipeline us: med %.1f" % (t[0][0] * 1e3, t[len(t) // 2][0] * 1e3, sorted(x[1] for x in t)[len(t) // 2] * 1e3), flush=True)
