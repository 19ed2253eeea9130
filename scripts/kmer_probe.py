#!/usr/bin/env python3
"""Tuning aid (needs a build with RJ_NVCC_EXTRA=-DRJ_KMER_PROBE): when the warps of k_set_kmer stop streaming,
per warp id and per CTA, for the 50 MB and the 625 MB text.  Prints the engine's [kmer probe] lines."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import rejit_b200 as rj  # noqa: E402
from rejit_b200 import workloads as W  # noqa: E402

rs = rj.RegejSet(W.DNA_PATTERNS)
for n in (5_000_000, 62_500_000):
    seq = W.fasta_sequence(5_000_000 if n == 5_000_000 else 6_250_000)
    text = seq if n == 5_000_000 else np.tile(seq, 10)
    dt = rj.DeviceText(text)
    st = rj.Stats()
    for _ in range(4):
        rj.lib().rejit_b200_flush_l2(0)
        rs.match_all_device(dt, stats=st)
    print("bytes", len(text), "scan_ms", st.scan_ms, flush=True)
    dt.free()
