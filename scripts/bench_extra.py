#!/usr/bin/env python3
"""Secondary measurements (not the contract line of bench.py): the other
BASELINE.json configurations with the text resident in HBM, one JSON line per
case — device pipeline time, scan-kernel time, GB/s and fraction of the measured
HBM bandwidth.  Used to fill the table in DESIGN.md / profiles/."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import rejit_b200 as rj  # noqa: E402
from rejit_b200 import workloads as W  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def run(name, pattern, text, reps=int(os.environ.get("RJ_EXTRA_REPS", "10")), flush=True):
    r = rj.Regej(pattern)
    dt = rj.DeviceText(text)
    st = rj.Stats()
    for _ in range(3):
        r.match_all_device(dt, stats=st)
    tot = scan = 0.0
    for _ in range(reps):
        if flush:
            rj.lib().rejit_b200_flush_l2(0)
        cnt = r.match_all_device(dt, stats=st)
        tot += st.total_ms
        scan += st.scan_ms
    n = len(text)
    line = {"case": name, "pattern": pattern if len(pattern) < 70 else pattern[:67] + "...", "bytes": n,
            "matches": cnt, "strategy": r.describe().split(";")[0],
            "pipeline_ms": round(tot / reps, 4), "scan_ms": round(scan / reps, 4),
            "pipeline_gbs": round(n / (tot / reps) / 1e6, 1), "scan_gbs": round(n / (scan / reps) / 1e6, 1),
            "scan_frac_of_hbm": round(n / (scan / reps) / 1e6 / peak(), 4), "launches": st.launches}
    print(json.dumps(line), flush=True)
    dt.free()


def run_set(name, patterns, text, reps=int(os.environ.get("RJ_EXTRA_REPS", "10"))):
    rs = rj.RegejSet(patterns)
    dt = rj.DeviceText(text)
    st = rj.Stats()
    for _ in range(3):
        counts = rs.match_all_device(dt, stats=st)
    tot = scan = 0.0
    for _ in range(reps):
        rj.lib().rejit_b200_flush_l2(0)
        counts = rs.match_all_device(dt, stats=st)
        tot += st.total_ms
        scan += st.scan_ms
    n = len(text)
    print(json.dumps({"case": name, "bytes": n, "matches": sum(counts), "strategy": rs.describe()[:60],
                      "pipeline_ms": round(tot / reps, 4), "scan_ms": round(scan / reps, 4),
                      "pipeline_gbs": round(n / (tot / reps) / 1e6, 1), "scan_gbs": round(n / (scan / reps) / 1e6, 1),
                      "scan_frac_of_hbm": round(n / (scan / reps) / 1e6 / peak(), 4), "launches": st.launches}), flush=True)
    dt.free()


def main():
    big = int(os.environ.get("RJ_EXTRA_BYTES", "500000000"))
    seq50 = W.fasta_sequence(5_000_000)
    run_set("C2 nine patterns fused, 50 MB FASTA", W.DNA_PATTERNS, seq50)
    run_set("C5 slab: nine patterns fused, 625 MB FASTA (the 62.5 MB sequence tiled)", W.DNA_PATTERNS,
            np.tile(W.fasta_sequence(6_250_000), 10)[:min(625_000_000, big * 5 // 4)])
    t = W.random_ascii(1 << 20, seed=1)
    run("C1 literal, 1 MiB random ASCII (L2 resident)", W.LITERAL_PATTERN, t, flush=False)
    text = W.random_ascii(big, seed=21)
    run("C3 complex regex, random text, no hits", W.COMPLEX_PATTERN, text)
    run("literal 'regexp', same text", W.LITERAL_PATTERN, text)
    W.plant(text, W.COMPLEX_HITS, every=10_007)
    run("C3 complex regex, hits every ~10 kB", W.COMPLEX_PATTERN, text)
    blob = W.source_blob(big, seed=3)
    run("C4 jrep literal ';\\n}' over source blob", W.JREP_PATTERN, blob)
    run("jrep line index '^'", "^", blob[:100_000_000])
    seq = W.fasta_sequence(5_000_000)
    run("C2 dna #1", W.DNA_PATTERNS[0], seq)
    run("C5 IUB 'B'", "B", seq)
    fa = np.frombuffer(W.fasta_file(2_000_000), dtype=np.uint8)
    run("C5 strip '>.*\\n|\\n' over 20 MB FASTA file", W.STRIP_PATTERN, fa)


def regexdna_chain(n_lines=2_000_000):
    """C5 shape (sample/regexdna.cc:49-91) with everything on the device: upload the
    FASTA file once, strip headers/newlines, count the nine variants in one fused
    scan, apply the eleven IUB substitutions; only counts and lengths come back."""
    fa = W.fasta_file(n_lines)
    strip = rj.Regej(W.STRIP_PATTERN)
    iub = [(rj.Regej(c), a.encode()) for c, a in W.IUB_SUBSTITUTIONS]
    rs = rj.RegejSet(W.DNA_PATTERNS)
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        cur = rj.Text(fa)
        t1 = time.perf_counter()
        st = rj.Stats()
        nxt, n_strip = strip.replace_all_text(cur, b"", stats=st)
        strip_dev = st.total_ms
        cur.free(); cur = nxt
        stripped = len(cur)
        t2 = time.perf_counter()
        counts = rs.match_all_text(cur)
        t3 = time.perf_counter()
        nxt, _ = rj.replace_all_set_text([r for r, _ in iub], cur, [a for _, a in iub], stats=st)
        iub_dev = st.total_ms
        cur.free(); cur = nxt
        final = len(cur)
        t4 = time.perf_counter()
        cur.free()
        line = {"case": "C5 regex-dna chain on device", "bytes": len(fa), "stripped": stripped, "final": final,
                "counts": counts, "upload_ms": round((t1 - t0) * 1e3, 3), "strip_ms": round((t2 - t1) * 1e3, 3),
                "count9_ms": round((t3 - t2) * 1e3, 3), "iub11_ms": round((t4 - t3) * 1e3, 3),
                "strip_device_ms": round(strip_dev, 3), "iub11_device_ms": round(iub_dev, 3),
                "total_ms": round((t4 - t0) * 1e3, 3), "gbs_of_input": round(len(fa) / (t4 - t0) / 1e9, 2)}
        if best is None or line["total_ms"] < best["total_ms"]:
            best = line
    print(json.dumps(best), flush=True)


if __name__ == "__main__":
    main()
    regexdna_chain()
    regexdna_chain(int(os.environ.get("RJ_EXTRA_CHAIN_LINES", "50000000")))
