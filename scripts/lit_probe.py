#!/usr/bin/env python3
"""Literal-scan cases only (k_lit_scan): dense and sparse hits, text resident, L2 flushed; one JSON line per case."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rejit_b200 as rj
from rejit_b200 import workloads as W


def run(name, pattern, text, reps=6):
    r = rj.Regej(pattern)
    dt = rj.DeviceText(text)
    st = rj.Stats()
    for _ in range(3):
        cnt = r.match_all_device(dt, stats=st)
    tot = scan = 0.0
    for _ in range(reps):
        rj.lib().rejit_b200_flush_l2(0)
        cnt = r.match_all_device(dt, stats=st)
        tot += st.total_ms
        scan += st.scan_ms
    n = len(text)
    print(json.dumps({"case": name, "bytes": n, "matches": cnt, "pipeline_ms": round(tot / reps, 4), "scan_ms": round(scan / reps, 4),
                      "pipeline_gbs": round(n / (tot / reps) / 1e6, 1), "scan_gbs": round(n / (scan / reps) / 1e6, 1),
                      "launches": st.launches}), flush=True)
    dt.free()


seq = W.fasta_sequence(5_000_000)
run("C5 IUB 'B' (300 k hits in 50 MB)", "B", seq)
run("C2 literal 'agggtaaa'", "agggtaaa", seq)
text = W.random_ascii(200_000_000, seed=21)
run("literal 'regexp', 200 MB random, no hits", W.LITERAL_PATTERN, text)
run("single byte 'q', 200 MB random (2.7 M hits)", "q", text)
blob = W.source_blob(100_000_000, seed=3)
run("C4 jrep literal ';\\n}' over 100 MB source blob", W.JREP_PATTERN, blob)
