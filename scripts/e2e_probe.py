#!/usr/bin/env python3
"""Times the individual steps of the upload-once end-to-end path (diagnostic)."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rejit_b200 as rj
from rejit_b200 import workloads as W
L = rj.lib()
seq = W.fasta_sequence(5_000_000)
n = len(seq)
pinned = L.rejit_b200_pinned_alloc(n)
ctypes.memmove(pinned, seq.ctypes.data, n)
regs = [rj.Regej(p) for p in W.DNA_PATTERNS]
for r in regs: r.compile()
err = ctypes.create_string_buffer(256)
for rep in range(4):
    t0 = time.perf_counter()
    h = L.rejit_b200_text_upload(0, pinned, n, err, 256)
    t1 = time.perf_counter()
    per = []
    for r in regs:
        pairs = ctypes.POINTER(ctypes.c_uint64)()
        ta = time.perf_counter()
        k = L.rejit_b200_match_all_text(r._prog, h, ctypes.byref(pairs), None, err, 256)
        per.append(round((time.perf_counter() - ta) * 1e6))
        L.rejit_b200_free(pairs)
    t2 = time.perf_counter()
    L.rejit_b200_text_free(h)
    t3 = time.perf_counter()
    print("rep", rep, "upload_us", round((t1 - t0) * 1e6), "calls_us", per, "free_us", round((t3 - t2) * 1e6), flush=True)
dt = rj.DeviceText(seq)
st = rj.Stats()
for rep in range(3):
    per = []
    for r in regs:
        ta = time.perf_counter(); r.match_all_device(dt, stats=st); per.append((round((time.perf_counter() - ta) * 1e6), round(st.total_ms * 1e3), round(st.scan_ms * 1e3)))
    print("device path (wall_us, pipeline_us, scan_us):", per, flush=True)
for rep in range(2):
    per = []
    for r in regs:
        ta = time.perf_counter(); r.match_all_device(dt); per.append(round((time.perf_counter() - ta) * 1e6))
    print("device path no stats wall_us:", per, flush=True)
