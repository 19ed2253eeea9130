#!/usr/bin/env python3
"""Tuning aid: a few device-resident cases, one short line each (select the library with RJ_LIB=...).
   usage: ab_run.py [case ...]   cases: lit c3 c3hits c4 b hat strip kmer50 kmer625 iub striprep"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import rejit_b200 as rj  # noqa: E402
from rejit_b200 import workloads as W  # noqa: E402

REPS = int(os.environ.get("RJ_AB_REPS", "8"))
cases = sys.argv[1:] or ["lit", "c3", "c3hits", "c4", "b", "hat", "strip", "kmer50", "kmer625"]
_cache = {}


def text_of(kind):
    if kind not in _cache:
        if kind == "rand":
            _cache[kind] = W.random_ascii(500_000_000, seed=21)
        elif kind == "randhits":
            t = W.random_ascii(500_000_000, seed=21)
            _cache[kind] = W.plant(t, W.COMPLEX_HITS, every=10_007)
        elif kind == "src":
            _cache[kind] = W.source_text_range(0, 500_000_000).numpy()
        elif kind == "src100":
            _cache[kind] = W.source_text_range(0, 100_000_000).numpy()
        elif kind == "seq50":
            _cache[kind] = W.fasta_sequence(5_000_000)
        elif kind == "seq625":
            _cache[kind] = np.tile(W.fasta_sequence(6_250_000), 10)
        elif kind == "file":
            _cache[kind] = np.frombuffer(W.fasta_file(10_000_000), dtype=np.uint8)
    return _cache[kind]


def timed(call, dt):
    st = rj.Stats()
    for _ in range(3):
        call(dt, st)
    best = 1e9
    tot = 0.0
    for _ in range(REPS):
        rj.lib().rejit_b200_flush_l2(0)
        out = call(dt, st)
        tot += st.scan_ms
        best = min(best, st.scan_ms)
    return out, tot / REPS, best, st.launches


SPEC = {"lit": (W.LITERAL_PATTERN, "rand"), "c3": (W.COMPLEX_PATTERN, "rand"), "c3hits": (W.COMPLEX_PATTERN, "randhits"),
        "c4": (W.JREP_PATTERN, "src"), "b": ("B", "seq50"), "hat": ("^", "src100"), "strip": (W.STRIP_PATTERN, "file")}
for c in cases:
    if c in SPEC:
        pat, kind = SPEC[c]
        text = text_of(kind)
        r = rj.Regej(pat)
        dt = rj.DeviceText(text)
        out, avg, best, la = timed(lambda d, s: r.match_all_device(d, stats=s), dt)
        n = len(text)
    elif c in ("kmer50", "kmer625"):
        text = text_of("seq50" if c == "kmer50" else "seq625")
        rs = rj.RegejSet(W.DNA_PATTERNS)
        dt = rj.DeviceText(text)
        out, avg, best, la = timed(lambda d, s: sum(rs.match_all_device(d, stats=s)), dt)
        n = len(text)
    elif c == "iub":
        text = text_of("seq625")
        n = len(text)
        tx = rj.Text(text)
        regs = [rj.Regej(p) for p, _ in W.IUB_SUBSTITUTIONS]
        withs = [w.encode() for _, w in W.IUB_SUBSTITUTIONS]
        best, tot = 1e9, 0.0
        for i in range(REPS + 2):
            st = rj.Stats()
            res, _counts = rj.replace_all_set_text(regs, tx, withs, stats=st)
            out = len(res)
            res.free()
            if i >= 2:
                tot += st.total_ms
                best = min(best, st.total_ms)
        avg, la = tot / REPS, st.launches
        dt = tx
    elif c in ("c4big", "litbig"):
        import torch
        n = 5_000_000_000
        t = (W.source_text_range(0, n + 64, device="cuda") if c == "c4big" else W.random_ascii_range(0, n + 64, device="cuda"))
        r = rj.Regej(W.JREP_PATTERN if c == "c4big" else W.LITERAL_PATTERN)
        dt = rj.DeviceText(nbytes=n, device=0, borrowed_ptr=t.data_ptr())
        out, avg, best, la = timed(lambda d, s: r.match_all_device(d, stats=s), dt)
    elif c == "striprep":
        text = np.frombuffer(W.fasta_file(50_000_000), dtype=np.uint8)
        n = len(text)
        tx = rj.Text(text)
        r = rj.Regej(W.STRIP_PATTERN)
        best, tot = 1e9, 0.0
        for i in range(REPS + 2):
            st = rj.Stats()
            res, _m = r.replace_all_text(tx, b"", stats=st)
            out = len(res)
            res.free()
            if i >= 2:
                tot += st.total_ms
                best = min(best, st.total_ms)
        avg, la = tot / REPS, st.launches
        dt = tx
    else:
        continue
    print("%-8s n=%d out=%s avg %.4f ms best %.4f ms  %.0f GB/s (best %.0f)  launches %d" %
          (c, n, out, avg, best, n / avg / 1e6, n / best / 1e6, la), flush=True)
    dt.free()
