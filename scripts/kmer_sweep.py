#!/usr/bin/env python3
"""Tuning aid: k_set_kmer's scan time against the text size (is the per-byte cost flat?)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import rejit_b200 as rj  # noqa: E402
from rejit_b200 import workloads as W  # noqa: E402

rs = rj.RegejSet(W.DNA_PATTERNS)
base = W.fasta_sequence(6_250_000)
for mb in (25, 50, 100, 150, 200, 250, 300, 400, 500, 625, 1000):
    text = np.tile(base, (mb * 1_000_000) // len(base) + 1)[:mb * 1_000_000]
    dt = rj.DeviceText(text)
    st = rj.Stats()
    for _ in range(3):
        rs.match_all_device(dt, stats=st)
    tot = 0.0
    for _ in range(5):
        rj.lib().rejit_b200_flush_l2(0)
        rs.match_all_device(dt, stats=st)
        tot += st.scan_ms
    print(json.dumps({"mb": mb, "scan_us": round(tot / 5 * 1e3, 1), "us_per_100mb": round(tot / 5 * 1e3 / mb * 100, 2),
                      "tbs": round(mb / (tot / 5) / 1e3, 3)}), flush=True)
    dt.free()
