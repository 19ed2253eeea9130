#!/usr/bin/env python3
"""Tuning aid: device time of k_set_kmer cut after each phase (RJ_KMER_STOP=1..4, 0 = whole kernel)."""
import os, subprocess, sys
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import rejit_b200 as rj
    from rejit_b200 import workloads as W
    seq = W.fasta_sequence(5_000_000)
    rs = rj.RegejSet(W.DNA_PATTERNS)
    dt = rj.DeviceText(seq)
    st = rj.Stats()
    t = []
    for i in range(12):
        rj.lib().rejit_b200_flush_l2(0)
        rs.match_all_device(dt, stats=st)
        t.append(st.scan_ms)
    t = sorted(t[2:])
    print("stop", os.environ.get("RJ_KMER_STOP", "0"), "kernel us: min %.1f med %.1f" % (t[0] * 1e3, t[len(t) // 2] * 1e3))
else:
    for stop in ("1", "2", "3", "4", "5", "6", "0"):
        env = dict(os.environ, RJ_KMER_STOP=stop)
        subprocess.run([sys.executable, __file__, "child"], env=env)
