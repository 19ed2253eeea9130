#!/usr/bin/env python3
"""Tuning aid: one case of the single-pass scan, a few device-resident calls (the target of an ncu capture).
   usage: emit_probe.py literal|c4|hat|strip|b|c3hits|kmer625"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import rejit_b200 as rj  # noqa: E402
from rejit_b200 import workloads as W  # noqa: E402

case = sys.argv[1]
if case == "literal":
    pat, text = W.LITERAL_PATTERN, W.random_ascii(200_000_000, seed=21)
elif case == "c4":
    pat, text = W.JREP_PATTERN, W.source_text_range(0, 200_000_000).numpy()
elif case == "hat":
    pat, text = "^", W.source_text_range(0, 100_000_000).numpy()
elif case == "strip":
    pat, text = W.STRIP_PATTERN, np.frombuffer(W.fasta_file(10_000_000), dtype=np.uint8)
elif case == "b":
    pat, text = "B", W.fasta_sequence(5_000_000)
elif case == "c3hits":
    text = W.random_ascii(200_000_000, seed=21)
    W.plant(text, W.COMPLEX_HITS, every=10_007)
    pat = W.COMPLEX_PATTERN
elif case == "kmer625":
    rs = rj.RegejSet(W.DNA_PATTERNS)
    dt = rj.DeviceText(np.tile(W.fasta_sequence(6_250_000), 10))
    st = rj.Stats()
    for _ in range(4):
        rj.lib().rejit_b200_flush_l2(0)
        print(sum(rs.match_all_device(dt, stats=st)), st.scan_ms)
    sys.exit(0)
r = rj.Regej(pat)
dt = rj.DeviceText(text)
st = rj.Stats()
for _ in range(4):
    rj.lib().rejit_b200_flush_l2(0)
    print(r.match_all_device(dt, stats=st), st.scan_ms, st.launches)
