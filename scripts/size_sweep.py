#!/usr/bin/env python3
"""Tuning aid: the literal scan over texts of growing size (how much of the 500 MB figure is wave quantisation)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import rejit_b200 as rj  # noqa: E402
from rejit_b200 import workloads as W  # noqa: E402

base = W.random_ascii(500_000_000, seed=21)
r = rj.Regej(W.LITERAL_PATTERN)
for mult in (1, 2, 4, 8):
    text = np.tile(base, mult) if mult > 1 else base
    dt = rj.DeviceText(text)
    st = rj.Stats()
    best = 1e9
    for i in range(6):
        rj.lib().rejit_b200_flush_l2(0)
        cnt = r.match_all_device(dt, stats=st)
        if i >= 2:
            best = min(best, st.scan_ms)
    print("literal %5d MB: matches %d best %.4f ms  %.0f GB/s" % (len(text) // 1000000, cnt, best, len(text) / best / 1e6), flush=True)
    dt.free()
    del text
