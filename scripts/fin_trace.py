#!/usr/bin/env python3
"""Debugging aid: phase times of the in-kernel finish of k_set_tma (RJ_FIN_TRACE=1)."""
import os, sys
os.environ["RJ_FIN_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rejit_b200 as rj
from rejit_b200 import workloads as W
seq = W.fasta_sequence(5_000_000)
rs = rj.RegejSet(W.DNA_PATTERNS)
dt = rj.DeviceText(seq)
st = rj.Stats()
for i in range(6):
    if not os.environ.get("RJ_NOFLUSH"):
        rj.lib().rejit_b200_flush_l2(0)
    rs.match_all_device(dt, stats=st)
    print("total_ms", round(st.total_ms, 4), "scan_ms", round(st.scan_ms, 4), file=sys.stderr)
