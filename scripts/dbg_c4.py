import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import rejit_b200 as rj
from rejit_b200 import workloads as W
for n in (1_000_000, 20_000_000, 200_000_000):
    text = W.source_text_range(0, n).numpy()
    for pat in (W.JREP_PATTERN, "B", ";"):
        r = rj.Regej(pat)
        dt = rj.DeviceText(text)
        for i in range(3):
            st = rj.Stats()
            c = r.match_all_device(dt, stats=st)
            print(n, repr(pat), "call", i, "matches", c, "launches", st.launches, "reruns", st.reruns, "large", st.large_path, "scan_ms %.4f" % st.scan_ms, flush=True)
        dt.free()
